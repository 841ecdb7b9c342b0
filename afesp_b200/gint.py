"""Gaussian-integral generator for the reference's input files (s.dat, t.dat, v.dat, eri.dat) -- SURVEY.md section 8 f-4.

The reference takes its integrals from Psi4 (tools/generate_integrals.py in the reference tree) and ships no eri.dat for
sample_data/h2o-cc-pvtz (listed in .MISSING_LARGE_BLOBS), so that sample -- BASELINE.json configs[1], the only one with
published reference timings -- cannot be run from the checkout.  This module regenerates the four files from geom.dat
and the public basis-set parameters with the C integral code in host/gint.c (McMurchie-Davidson, Psi4 conventions: unit
self-overlap, pure functions m = 0, +1, -1, ..., shells in basis-file order atom by atom).  Host-side tool: no GPU.

    python -m afesp_b200.gint <dir with geom.dat> cc-pvtz     # writes s.dat t.dat v.dat eri.dat next to geom.dat
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(os.path.dirname(HERE), "host", "libafesp_gint.so")

# Correlation-consistent basis sets (Dunning 1989), parameters as published (Basis Set Exchange / Psi4 *.gbs):
# element -> list of shells (l, [(exponent, coefficient), ...]); general contractions are listed as separate shells.
_O_S_DZ = [11720.0, 1759.0, 400.8, 113.7, 37.03, 13.27, 5.025, 1.013]
_O_S_TZ = [15330.0, 2299.0, 522.4, 147.3, 47.55, 16.76, 6.207, 0.6882]
_N_S_DZ = [9046.0, 1357.0, 309.3, 87.73, 28.56, 10.21, 3.838, 0.7466]
_F_S_DZ = [14710.0, 2207.0, 502.8, 142.6, 46.47, 16.70, 6.356, 1.316]
BASIS = {
    "cc-pvdz": {
        1: [(0, [(13.01, 0.019685), (1.962, 0.137977), (0.4446, 0.478148)]), (0, [(0.122, 1.0)]), (1, [(0.727, 1.0)])],
        8: [(0, list(zip(_O_S_DZ, [0.000710, 0.005470, 0.027837, 0.104800, 0.283062, 0.448719, 0.270952, 0.015458]))),
            (0, list(zip(_O_S_DZ, [-0.000160, -0.001263, -0.006267, -0.025716, -0.070924, -0.165411, -0.116955, 0.557368]))),
            (0, [(0.3023, 1.0)]),
            (1, [(17.70, 0.043018), (3.854, 0.228913), (1.046, 0.508728)]), (1, [(0.2753, 1.0)]),
            (2, [(1.185, 1.0)])],
        # nitrogen and fluorine: reproduce every shipped integral of sample_data/n2-cc-pvdz and f2-cc-pvdz (tests/test_gint.py)
        7: [(0, list(zip(_N_S_DZ, [0.000700, 0.005389, 0.027406, 0.103207, 0.278723, 0.448540, 0.278238, 0.015440]))),
            (0, list(zip(_N_S_DZ, [-0.000153, -0.001208, -0.005992, -0.024544, -0.067459, -0.158078, -0.121831, 0.549003]))),
            (0, [(0.2248, 1.0)]),
            (1, [(13.55, 0.039919), (2.917, 0.217169), (0.7973, 0.510319)]), (1, [(0.2185, 1.0)]),
            (2, [(0.817, 1.0)])],
        9: [(0, list(zip(_F_S_DZ, [0.000721, 0.005553, 0.028267, 0.106444, 0.286814, 0.448641, 0.264761, 0.015333]))),
            (0, list(zip(_F_S_DZ, [-0.000165, -0.001308, -0.006495, -0.026691, -0.073690, -0.170776, -0.112327, 0.562814]))),
            (0, [(0.3897, 1.0)]),
            (1, [(22.67, 0.044878), (4.977, 0.235718), (1.347, 0.508521)]), (1, [(0.3471, 1.0)]),
            (2, [(1.640, 1.0)])],
    },
    # Weigend & Ahlrichs 2005.  The reference's `sample_data/h2o-cc-pvdz` files were in fact generated with this basis
    # (kinetic diagonal of its t.dat: d exponent 1.2, hydrogen p exponent 0.8), see tests/test_gint.py.
    "def2-svp": {
        1: [(0, [(13.0107010, 0.19682158e-01), (1.9622572, 0.13796524), (0.44453796, 0.47831935)]),
            (0, [(0.12194962, 1.0)]), (1, [(0.8, 1.0)])],
        8: [(0, [(2266.1767785, -0.53431809926e-02), (340.87010191, -0.39890039230e-01), (77.363135167, -0.17853911985),
                 (21.479644940, -0.46427684959), (6.6589433124, -0.44309745172)]),
            (0, [(0.80975975668, 1.0)]), (0, [(0.25530772234, 1.0)]),
            (1, [(17.721504317, 0.43394573193e-01), (3.8635505440, 0.23094120765), (1.0480920883, 0.51375311064)]),
            (1, [(0.27641544411, 1.0)]), (2, [(1.2, 1.0)])],
    },
    "cc-pvtz": {
        1: [(0, [(33.87, 0.006068), (5.095, 0.045308), (1.159, 0.202822)]), (0, [(0.3258, 1.0)]), (0, [(0.1027, 1.0)]),
            (1, [(1.407, 1.0)]), (1, [(0.388, 1.0)]), (2, [(1.057, 1.0)])],
        8: [(0, list(zip(_O_S_TZ, [0.000508, 0.003929, 0.020243, 0.079181, 0.230687, 0.433118, 0.350260, -0.008154]))),
            (0, list(zip(_O_S_TZ, [-0.000115, -0.000895, -0.004636, -0.018724, -0.058463, -0.136463, -0.175740, 0.603418]))),
            (0, [(1.752, 1.0)]), (0, [(0.2384, 1.0)]),
            (1, [(34.46, 0.015928), (7.749, 0.099740), (2.280, 0.310492)]), (1, [(0.7156, 1.0)]), (1, [(0.2140, 1.0)]),
            (2, [(2.314, 1.0)]), (2, [(0.645, 1.0)]),
            (3, [(1.428, 1.0)])],
    },
}

_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(os.path.dirname(HERE), "host", "gint.c")
        if not os.path.exists(SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(SO)):
            subprocess.run(["make", "-C", os.path.dirname(SO), "libafesp_gint.so"], check=True, capture_output=True)
        _lib = C.CDLL(SO)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        _lib.afesp_gint_compute.argtypes = [C.c_int, dp, dp, C.c_int, ip, ip, ip, ip, dp, dp, dp, dp, dp, dp]
        _lib.afesp_gint_compute.restype = C.c_int
    return _lib


def compute(Z, xyz, basis, want_eri=True, shells=None):
    """Z[natom] nuclear charges, xyz[natom,3] in bohr, basis = name in BASIS (or `shells`: per-atom list of shell lists).
    Returns dict(s, t, v: dense nbf x nbf; eri: packed in the reference's 8-fold canonical order, or None)."""
    lib = load()
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
    sh_atom, sh_l, sh_np, sh_off, exps, coefs = [], [], [], [], [], []
    for at, z in enumerate(Z):
        for (l, prims) in (shells[at] if shells is not None else BASIS[basis.lower()][int(round(z))]):
            sh_atom.append(at); sh_l.append(l); sh_np.append(len(prims)); sh_off.append(len(exps))
            exps += [p[0] for p in prims]; coefs += [p[1] for p in prims]
    nbf = sum(2 * l + 1 for l in sh_l)
    S, T, V = (np.zeros((nbf, nbf)) for _ in range(3))
    npair = nbf * (nbf + 1) // 2
    eri = np.zeros(npair * (npair + 1) // 2) if want_eri else None
    ia = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    a1, a2, a3, a4 = ia(sh_atom), ia(sh_l), ia(sh_np), ia(sh_off)
    e, c = np.ascontiguousarray(exps, dtype=np.float64), np.ascontiguousarray(coefs, dtype=np.float64)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    P = lambda a: a.ctypes.data_as(dp)
    n = lib.afesp_gint_compute(len(Z), P(Z), P(xyz), len(sh_l), a1.ctypes.data_as(ip), a2.ctypes.data_as(ip),
                               a3.ctypes.data_as(ip), a4.ctypes.data_as(ip), P(e), P(c), P(S), P(T), P(V),
                               P(eri) if want_eri else None)
    if n != nbf:
        raise RuntimeError("afesp_gint_compute rejected the basis description")
    return {"s": S, "t": T, "v": V, "eri": eri, "nbf": nbf}


def read_geom(path):
    """geom.dat (src/geometry.f90:8-50): atom count, then `Z x y z` per atom in bohr."""
    rows = [ln.split() for ln in open(path).read().strip().splitlines()]
    n = int(rows[0][0])
    Z = np.array([float(r[0]) for r in rows[1:1 + n]])
    xyz = np.array([[float(x) for x in r[1:4]] for r in rows[1:1 + n]])
    return Z, xyz


def write_dat_files(dirpath, res, threshold=0.0):
    """s.dat / t.dat / v.dat (`i j value`, lower triangle) and eri.dat (`i j k l value`, canonical order), 1-based, in the
    layout the reference's readers expect (src/integrals.f90:48-165)."""
    n = res["nbf"]
    for name, key in (("s.dat", "s"), ("t.dat", "t"), ("v.dat", "v")):
        with open(os.path.join(dirpath, name), "w") as f:
            for i in range(n):
                for j in range(i + 1):
                    f.write("%d\t%d\t%.15f\n" % (i + 1, j + 1, res[key][i, j]))
    if res["eri"] is not None:
        with open(os.path.join(dirpath, "eri.dat"), "w") as f:
            pos = 0
            for i in range(n):
                for j in range(i + 1):
                    for k in range(i + 1):
                        for l in range((j if k == i else k) + 1):
                            val = res["eri"][pos]
                            pos += 1
                            if abs(val) > threshold:
                                f.write("%d\t%d\t%d\t%d\t%.15f\n" % (i + 1, j + 1, k + 1, l + 1, val))


if __name__ == "__main__":
    d, basis = sys.argv[1], sys.argv[2]
    Z, xyz = read_geom(os.path.join(d, "geom.dat"))
    out = compute(Z, xyz, basis)
    write_dat_files(d, out)
    print("wrote s.dat t.dat v.dat eri.dat for %d basis functions in %s" % (out["nbf"], d))
