"""Python loader + timing harness for the C port of the reference's dominant CPU loops (oracle/cpu_kernels.c).

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by tests (checked against the NumPy oracle) and by bench.py's cpu_baseline and
`--impl reference` legs.  /root/reference cannot be compiled in this image (no Fortran compiler), so the
kind of this baseline is "port": same loop nests, same OpenMP decomposition, same -O3 -ffast-math -fopenmp, and the
ladder contraction through dgemm of the image's OpenBLAS (numpy), as the reference does at src/ccsd.f90:1669.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
import time

import numpy as np

from . import afesp_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libafesp_cpu_port.so")
_dp = C.POINTER(C.c_double)


_blas_keepalive = []


def _find_dgemm():
    """Fortran-interface dgemm of the image's OpenBLAS (the library numpy / scipy ship): (address, ilp64, path)."""
    import numpy
    import scipy
    for pkg, pat, sym, ilp64 in ((numpy, "numpy.libs/libscipy_openblas64_*.so", "scipy_dgemm_64_", 1),
                                 (scipy, "scipy.libs/libscipy_openblas*.so", "scipy_dgemm_", 0)):
        base = os.path.dirname(os.path.dirname(pkg.__file__))
        for path in sorted(glob.glob(os.path.join(base, pat))):
            try:
                bl = C.CDLL(path)
                fn = getattr(bl, sym)
            except (OSError, AttributeError):
                continue
            _blas_keepalive.append(bl)
            return C.cast(fn, C.c_void_p).value, ilp64, path, bl
    raise RuntimeError("no OpenBLAS dgemm found in numpy.libs / scipy.libs")


def set_threads(lib, nthreads=None):
    """Use `nthreads` (default: every host core) for the OpenMP loops AND the OpenBLAS dgemms, whatever OMP_NUM_THREADS
    says (torchrun exports OMP_NUM_THREADS=1).  Returns the thread count in force."""
    n = int(nthreads or os.cpu_count() or 1)
    lib.afesp_ref_set_threads(n)
    bl = lib._blas
    for sym in ("scipy_openblas_set_num_threads64_", "scipy_openblas_set_num_threads", "openblas_set_num_threads"):
        if hasattr(bl, sym):
            getattr(bl, sym)(C.c_int(n))
            break
    try:  # numpy's own matmul (the ladder helper below) goes through the same library; keep it in step
        from threadpoolctl import threadpool_limits
        lib._tp_limit = threadpool_limits(limits=n)
    except Exception:
        pass
    return n


def load():
    srcs = [os.path.join(HERE, f) for f in ("cpu_kernels.c", "cpu_ccsd.c")]
    if not os.path.exists(SO) or any(os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
        subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
    lib = C.CDLL(SO)
    addr, ilp64, path, bl = _find_dgemm()
    lib.afesp_ref_set_dgemm.argtypes = [C.c_void_p, C.c_int]
    lib.afesp_ref_set_dgemm(addr, ilp64)
    lib._blas, lib._blas_path = bl, path
    lib.afesp_ref_set_threads.argtypes = [C.c_int]
    _ip = C.POINTER(C.c_int)
    lib.afesp_ref_ccsd_iter.argtypes = [C.c_int, C.c_int] + [_dp] * 6 + [C.c_int, _dp, _dp, _dp, C.c_int, C.c_int, _dp]
    lib.afesp_ref_ccsd_iter.restype = None
    lib.afesp_ref_ao2mo.argtypes = [C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, _dp]
    lib.afesp_ref_ao2mo.restype = None
    lib.afesp_ref_slice.argtypes = [_dp] + [C.c_int] * 8 + [_dp]
    lib.afesp_ref_slice.restype = None
    lib.afesp_ref_reshape.argtypes = [_dp, _dp, _ip, C.c_char_p, C.c_int, C.c_double]
    lib.afesp_ref_reshape.restype = None
    lib.afesp_ref_dgemm.argtypes = [C.c_char, C.c_char, C.c_long, C.c_long, C.c_long, _dp, _dp, _dp, C.c_double, C.c_double]
    lib.afesp_ref_dgemm.restype = None
    lib.afesp_ref_ring.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_int]
    lib.afesp_ref_ring.restype = None
    lib.afesp_ref_triples.argtypes = [C.c_int, C.c_int] + [_dp] * 7 + [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, _dp]
    lib.afesp_ref_triples.restype = None
    lib.afesp_orbit_T_epilogue.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_double, _dp]
    lib.afesp_orbit_T_epilogue.restype = C.c_double
    lib.afesp_ref_triples_bounded.argtypes = ([C.c_int, C.c_int] + [_dp] * 7 +
                                              [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, _dp])
    lib.afesp_ref_triples_bounded.restype = None
    lib.afesp_ref_threads.restype = C.c_int
    lib.afesp_ref_triples_cr.argtypes = ([C.c_int, C.c_int] + [_dp] * 9 + [C.c_int, C.POINTER(C.c_int), C.c_int, _dp])
    lib.afesp_ref_triples_cr.restype = None
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


def triples_cr(lib, t1, t2, v_oovv, v_vvov, v_oovo, I_vovv_pp, I_ooov_pp, eps, ijk, paren):
    """(e_T, e_TT, D_T, D_TT, e_CR, e_CRT) over the listed ordered triples with the completely renormalised part
    (src/ccsd.f90:2152-2233 with doing_CR); returns (sums[6], seconds)."""
    o, v = t1.shape
    t2r = _F(t2.transpose(3, 2, 1, 0))
    vvovv = _F(v_vvov.transpose(3, 2, 1, 0))
    vovoo = _F(v_oovo.transpose(3, 2, 1, 0))
    a1, a2, a3, a4, a5 = _F(t1), _F(t2), _F(v_oovv), _F(I_vovv_pp), _F(I_ooov_pp)
    e = np.ascontiguousarray(eps, dtype=np.float64)
    tri = np.ascontiguousarray(np.asarray(ijk, dtype=np.int32).reshape(-1, 3))
    out = np.zeros(6)
    t0 = time.perf_counter()
    lib.afesp_ref_triples_cr(o, v, _p(a1), _p(a2), _p(t2r), _p(vvovv), _p(vovoo), _p(a3), _p(a4), _p(a5), _p(e),
                             tri.shape[0], tri.ctypes.data_as(C.POINTER(C.c_int)), int(paren), _p(out))
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


CCSD_ITER_PARTS = ("intermediates: dgemms + reshapes", "I_ovov loop :1170-1182", "I_voov / I_vovv_p / x_voov loops",
                   "T1 equations", "ladder dgemm :1669", "ring loop :1680-1695", "other T2 terms", "P(ia/jb) + divide")


def slices(lib, eri_mo, n, o, vvvv_cols=None):
    """init_cc slices (src/ccsd.f90:496-512) from the packed MO integrals, in the reference's layouts.
    vvvv_cols=(s0, ns): only columns d in [s0, s0+ns) of v_vvvv(a,b,c,d) (the dense slice is 134 GB at nbf=400)."""
    v = n - o
    e = np.ascontiguousarray(eri_mo, dtype=np.float64)

    def cut(lp, np_, lq, nq, lr, nr, ls, ns):
        out = np.empty((np_, nq, nr, ns), order="F")
        lib.afesp_ref_slice(_p(e), lp, np_, lq, nq, lr, nr, ls, ns, _p(out))
        return out

    s0, ns = (0, v) if vvvv_cols is None else vvvv_cols
    return {"v_oovv": cut(0, o, 0, o, o, v, o, v), "v_ovov": cut(0, o, o, v, 0, o, o, v),
            "v_vvov": cut(o, v, o, v, 0, o, o, v), "v_oovo": cut(0, o, 0, o, o, v, 0, o),
            "v_oooo": cut(0, o, 0, o, 0, o, 0, o), "v_vvvv": cut(o, v, o, v, o, v, o + s0, ns)}


def ccsd_iter(lib, V, eps, t1, t2, ring_bmax=0, iovov_amax=0):
    """One spin-free CCSD iteration exactly as the reference issues it (update_restricted_intermediates +
    update_amplitudes_restricted, src/ccsd.f90:1040-1312, 1538-1732).  Returns (t1_new, t2_new, seconds[8], wall).
    V: dict of column-major slices (see slices()); V["v_vvvv"] may be a block of its last axis (timing sample)."""
    o, v = t1.shape
    a1, a2 = _F(t1).copy(order="F"), _F(t2).copy(order="F")
    vv = V["v_vvvv"]
    ncol = v * vv.shape[3]
    work = {k: _F(V[k]) for k in ("v_oovv", "v_ovov", "v_vvov", "v_oovo", "v_oooo")}
    e = np.ascontiguousarray(eps, dtype=np.float64)
    times = np.zeros(8)
    t0 = time.perf_counter()
    lib.afesp_ref_ccsd_iter(o, v, _p(work["v_oovv"]), _p(work["v_ovov"]), _p(work["v_vvov"]), _p(work["v_oovo"]),
                            _p(work["v_oooo"]), _p(_F(vv)), ncol, _p(e), _p(a1), _p(a2), int(ring_bmax),
                            int(iovov_amax), _p(times))
    return a1, a2, times, time.perf_counter() - t0


def ao2mo(lib, eri_ao, Cmo, lmax=0, smax=0, want_result=True):
    """AO->MO transform as do_mp2_spatial does it (src/mp2.f90:321-410): four O(n^5) OpenMP loop nests + serial repack.
    Returns (eri_mo or None, seconds[5]).  lmax/smax < n: slab sample (see oracle/cpu_ccsd.c)."""
    n = Cmo.shape[0]
    full = (lmax in (0, n)) and (smax in (0, n))
    e = np.ascontiguousarray(eri_ao, dtype=np.float64)
    Cf = _F(Cmo)
    out = np.empty_like(e) if (want_result and full) else None
    times = np.zeros(5)
    lib.afesp_ref_ao2mo(n, _p(e), _p(Cf), int(lmax), int(smax), _p(out) if out is not None else None, _p(times))
    return out, times


def synthetic_mo_integrals(nbf, nocc, seed=20260):
    """Packed MO integrals of the synthetic workload (afesp_b200/synthetic.py) WITHOUT the O(n^5) transform, from the
    factored form: (pq|rs) = sum_P Bmo[pq,P] Bmo[rs,P], Bmo^P = C B^P C^T.  Used to feed the CPU timing legs with the same
    system the GPU arm runs (the CPU AO->MO itself is timed separately, on a slab).  Returns (eri_mo_packed, C, eps)."""
    from afesp_b200 import synthetic

    B, Cmo, eps = synthetic.make_factors(nbf, nocc, seed)
    n = nbf
    ii, jj = np.tril_indices(n)
    Bmo = np.empty_like(B)
    full = np.empty((n, n))
    for P in range(B.shape[1]):
        full[ii, jj] = B[:, P]
        full[jj, ii] = B[:, P]
        Bmo[:, P] = (Cmo @ full @ Cmo.T)[ii, jj]
    G = Bmo @ Bmo.T
    npair = ii.size
    packed = np.empty(npair * (npair + 1) // 2)
    pos = 0
    for r in range(npair):
        packed[pos:pos + r + 1] = G[r, :r + 1]
        pos += r + 1
    return packed, Cmo, eps


def orbit_T_fast(lib, t2, iv_block, v_oovo, eps, triples, progress=None, checkpoint=None):
    """sum over the listed unique (i <= j <= k) triples of their [T] orbit contributions -- the same quantity as
    oracle.afesp_oracle.triples_bracket_T_orbit_form, organised for large shapes (nbf=400: 11480 orbits of v^3 = 4.7e7
    elements): operands pre-arranged once in column-major blocks, the twelve products of an orbit accumulated by dgemm
    (beta = 1) into three buffers by row label, one C pass for the permuted sum and the energy expression
    (cpu_kernels.c: afesp_orbit_T_epilogue).  t2: (o,o,v,v) C-order; iv_block(k) -> [d,(b,c)] matrix of v_vovv(d,k,b,c);
    v_oovo: (o,o,v,o) C-order.  checkpoint = (path, every): partial sums are saved and a rerun resumes."""
    import json
    import os

    o, v = t2.shape[0], t2.shape[2]
    v3 = v ** 3
    eo, ev = np.ascontiguousarray(eps[:o]), np.ascontiguousarray(eps[o:])
    t2 = np.ascontiguousarray(t2)
    v_oovo = np.ascontiguousarray(v_oovo)
    used_k = sorted({x for t in triples for x in t})
    VF = np.empty((v, v, v, o), order="F")      # VF[d,y,z,k] = v_vovv(d,k,y,z);  VFT[d,y,z,k] = v_vovv(d,k,z,y)
    VFT = np.empty((v, v, v, o), order="F")
    for k in used_k:
        m = np.asarray(iv_block(k)).reshape(v, v, v)
        VF[:, :, :, k] = m
        VFT[:, :, :, k] = m.transpose(0, 2, 1)
    TF = np.asfortranarray(t2.transpose(0, 2, 3, 1))      # TF[l,y,z,A] = t2[l,A,y,z]
    TFT = np.asfortranarray(t2.transpose(0, 3, 2, 1))     # TFT[l,y,z,A] = t2[l,A,z,y]
    Y = [np.empty(v3) for _ in range(3)]
    W = np.empty(v3)
    S3 = [(0, 1, 2), (1, 0, 2), (2, 1, 0), (0, 2, 1), (1, 2, 0), (2, 0, 1)]
    ptr = lambda arr, off: C.cast(arr.ctypes.data + 8 * off, _dp)
    total, start = 0.0, 0
    if checkpoint and os.path.exists(checkpoint[0]):
        ck = json.load(open(checkpoint[0]))
        if ck.get("ntriples") == len(triples):
            total, start = ck["total"], ck["next"]
    t0 = time.perf_counter()
    for n_done in range(start, len(triples)):
        i, j, k = triples[n_done]
        idx = (i, j, k)
        first = [True, True, True]
        for p in S3:
            A, B, Cc = idx[p[0]], idx[p[1]], idx[p[2]]
            # particle: Y[p0](x; y,z) += sum_d t2[A,B](x,d) v_vovv(d,Cc; j_p1, j_p2)   (columns in increasing label position)
            cls = p[0]
            Vsrc = VF if p[1] < p[2] else VFT
            lib.afesp_ref_dgemm(b"T", b"N", v, v * v, v, ptr(t2, (A * o + B) * v * v), ptr(Vsrc, Cc * v3), ptr(Y[cls], 0), 1.0,
                                0.0 if first[cls] else 1.0)
            first[cls] = False
        for p in S3:
            A, B, Cc = idx[p[0]], idx[p[1]], idx[p[2]]
            # hole: Y[p2](x; y,z) -= sum_l v_oovo[Cc,B](x,l) t2[l,A; j_p1, j_p0]
            cls = p[2]
            Tsrc = TF if p[1] < p[0] else TFT
            lib.afesp_ref_dgemm(b"T", b"N", v, v * v, o, ptr(v_oovo, (Cc * o + B) * v * o), ptr(Tsrc, A * o * v * v), ptr(Y[cls], 0),
                                -1.0, 1.0)
        e = lib.afesp_orbit_T_epilogue(v, ptr(Y[0], 0), ptr(Y[1], 0), ptr(Y[2], 0), ptr(ev, 0), float(eo[i] + eo[j] + eo[k]),
                                       ptr(W, 0))
        total += e * len({(i, j, k), (i, k, j), (j, i, k), (j, k, i), (k, i, j), (k, j, i)}) / 6.0
        if progress and (n_done + 1) % progress == 0:
            dt = time.perf_counter() - t0
            print(f"[orbit_T_fast] {n_done + 1} / {len(triples)} orbits, {dt / (n_done + 1 - start):.2f} s per orbit", flush=True)
        if checkpoint and (n_done + 1) % checkpoint[1] == 0:
            json.dump({"ntriples": len(triples), "next": n_done + 1, "total": total}, open(checkpoint[0], "w"))
    return total
