"""Python loader + timing harness for the C port of the reference's dominant CPU loops (oracle/cpu_kernels.c).

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by tests (checked against the NumPy oracle) and by bench.py's cpu_baseline and
`--impl reference` legs.  /root/reference cannot be compiled in this image (no Fortran compiler), so the
kind of this baseline is "port": same loop nests, same OpenMP decomposition, same -O3 -ffast-math -fopenmp, and the
ladder contraction through dgemm of the image's OpenBLAS (numpy), as the reference does at src/ccsd.f90:1669.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

from . import afesp_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libafesp_cpu_port.so")
_dp = C.POINTER(C.c_double)


def load():
    src = os.path.join(HERE, "cpu_kernels.c")
    if not os.path.exists(SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(SO)):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
    lib = C.CDLL(SO)
    lib.afesp_ref_ring.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int]
    lib.afesp_ref_ring.restype = None
    lib.afesp_ref_triples.argtypes = [C.c_int, C.c_int] + [_dp] * 7 + [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, _dp]
    lib.afesp_ref_triples.restype = None
    lib.afesp_ref_triples_bounded.argtypes = ([C.c_int, C.c_int] + [_dp] * 7 +
                                              [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, _dp])
    lib.afesp_ref_triples_bounded.restype = None
    lib.afesp_ref_threads.restype = C.c_int
    return lib


def _F(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def ring(lib, t2, I_ovov, asym, I_voov, bmax=None):
    """tmp_t2 contribution of src/ccsd.f90:1680-1695; returns (array, seconds).  bmax: only b < bmax (sample)."""
    o, v = t2.shape[0], t2.shape[2]
    out = np.zeros((o, o, v, v), order="F")
    a, b, c, d = _F(t2), _F(I_ovov), _F(asym), _F(I_voov)
    t0 = time.perf_counter()
    lib.afesp_ref_ring(o, v, _p(a), _p(b), _p(c), _p(d), _p(out), v if bmax is None else int(bmax))
    return out, time.perf_counter() - t0


def triples(lib, t1, t2, v_oovv, v_vvov, v_oovo, eps, ijk, paren, renorm, amax=None):
    """(e_T, e_TT, D_T, D_TT) over the listed ordered triples, reference loop (src/ccsd.f90:2152-2233).
    amax: only the slab a < amax of the outer virtual loop (timing sample; the sums are then partial)."""
    o, v = t1.shape
    t2r = _F(t2.transpose(3, 2, 1, 0))
    vvovv = _F(v_vvov.transpose(3, 2, 1, 0))
    vovoo = _F(v_oovo.transpose(3, 2, 1, 0))
    a1, a2, a3 = _F(t1), _F(t2), _F(v_oovv)
    e = np.ascontiguousarray(eps, dtype=np.float64)
    tri = np.ascontiguousarray(np.asarray(ijk, dtype=np.int32).reshape(-1, 3))
    out = np.zeros(4)
    t0 = time.perf_counter()
    lib.afesp_ref_triples_bounded(o, v, _p(a1), _p(a2), _p(t2r), _p(vvovv), _p(vovoo), _p(a3), _p(e), tri.shape[0],
                                  tri.ctypes.data_as(C.POINTER(C.c_int)), int(paren), int(renorm),
                                  v if amax is None else int(amax), _p(out))
    return out, time.perf_counter() - t0


def ladder(c_oovv, v_vvvv):
    """1/2 c(ij,ef) v_vvvv(ef,ab) through OpenBLAS dgemm (src/ccsd.f90:1669); returns (array, seconds)."""
    o, v = c_oovv.shape[0], c_oovv.shape[2]
    A = _F(c_oovv).reshape((o * o, v * v), order="F")
    B = _F(v_vvvv).reshape((v * v, v * v), order="F")
    t0 = time.perf_counter()
    X = 0.5 * (A @ B)
    dt = time.perf_counter() - t0
    return X.reshape((o, o, v, v), order="F"), dt
