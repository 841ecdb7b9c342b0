"""CPU oracle for the AFESP coupled-cluster hot path (TEST INFRASTRUCTURE ONLY).

This module is a NumPy restatement of the reference algorithm for the path
named in BASELINE.json: input readers -> RHF (+Pulay DIIS) -> AO->MO ERI
transform -> MP2 -> {spin-free, spin-orbital} CCSD with CC-DIIS -> the
(T)/[T]/R/CR triples family.  It is the *checker* for the CUDA product path:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
legs may import it.  Nothing in afesp_b200/ imports it and the product path
never falls back to it.

Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
the reference's own shipped outputs (tests/golden/golden.json, parsed from
sample_data/*/els.out and ref_out by tests/golden/make_fixtures.py): every
SCF and CCSD iteration line, MP2, nine triples-family energies, D[T], D(T)
and the T1 diagnostic for N2 and F2; the spin-orbital CCSD table of the
older H2O ref_out (Stanton-correct F_mi, see q1 switch).

Every function cites the reference lines it follows (paths relative to the
reference root, src/...).  Arrays use the reference's index order, e.g.
t2[i,j,a,b]; the memory layout is irrelevant here.  Index helpers are
0-based; the packed ERI order equals the reference's (eri_ind - 1).

Reference quirks reproduced: Q1 transposed F_oo dgemm in
the spin-orbital build_F; Q2 plain CCSD(T)_spatial gives E(T)=E[T]; Q3a the
truncated `do e = 1, nocc` loop in I_ooov_pp; Q3b stale I_vo/asym_t2 in the
CR intermediates; Q4 "delta RMS T2" is the squared Frobenius norm.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

# --------------------------------------------------------------------------
# packed ERI indexing -- src/integrals.f90:196-210 (eri_ind), 0-based
# --------------------------------------------------------------------------


def tri(i, j):
    """eri_ind (src/integrals.f90:196-210), 0-based: max*(max+1)/2 + min."""
    i = np.asarray(i, dtype=np.int64)
    j = np.asarray(j, dtype=np.int64)
    hi = np.maximum(i, j)
    lo = np.minimum(i, j)
    return hi * (hi + 1) // 2 + lo


def npair(n):
    return n * (n + 1) // 2


def npacked(n):
    m = npair(n)
    return m * (m + 1) // 2


def unpack_eri(packed, n):
    """Packed 8-fold (ij|kl) -> dense chemist array g[i,j,k,l]."""
    idx = np.arange(n)
    ij = tri(idx[:, None], idx[None, :])  # (n,n)
    ijkl = tri(ij[:, :, None, None], ij[None, None, :, :])
    return np.asarray(packed)[ijkl]


def pack_eri(g):
    """Dense chemist g[p,q,r,s] -> packed canonical order.

    Follows the repack loop src/mp2.f90:388-410: p>=q, r<=p, s<=(q if r==p
    else r); the running counter there equals eri_ind(eri_ind(p,q),
    eri_ind(r,s)).
    """
    n = g.shape[0]
    m = npair(n)
    p, q = np.tril_indices(n)  # row-major order of lower triangle == pair index order
    assert np.array_equal(tri(p, q), np.arange(m))
    gp = g[p[:, None], q[:, None], p[None, :], q[None, :]]  # (m,m) pair matrix
    a, b = np.tril_indices(m)
    return np.ascontiguousarray(gp[a, b])


# --------------------------------------------------------------------------
# input files -- src/integrals.f90:48-165, src/geometry.f90:8-50,74-95,
#                src/system.f90:81-167, src/hf.f90:153-170
# --------------------------------------------------------------------------

CALC_TYPES = {
    # calc_type string -> (level, restricted, paren, renorm, comp_renorm); src/system.f90:116-165
    "RHF": ("HF", True, False, False, False),
    "UHF": ("HF", False, False, False, False),
    "MP2_spinorb": ("MP2", False, False, False, False),
    "MP2_spatial": ("MP2", True, False, False, False),
    "CCSD_spinorb": ("CCSD", False, False, False, False),
    "CCSD_spatial": ("CCSD", True, False, False, False),
    "CCSD(T)_spinorb": ("CCSD(T)", False, False, False, False),
    "CCSD(T)_spatial": ("CCSD(T)", True, True, False, False),
    "CCSD[T]_spatial": ("CCSD(T)", True, False, False, False),
    "RCCSD(T)_spatial": ("CCSD(T)", True, True, True, False),
    "RCCSD[T]_spatial": ("CCSD(T)", True, False, True, False),
    "CRCCSD(T)_spatial": ("CCSD(T)", True, True, False, True),
    "CRCCSD[T]_spatial": ("CCSD(T)", True, False, False, True),
}


@dataclass
class System:
    """Mirror of system_t (src/system.f90:10-69) plus the integral store."""

    nbasis: int = 0
    nel: int = 0
    nocc: int = 0  # spatial occupied count (nel/2)
    e_nuc: float = 0.0
    ovlp: np.ndarray | None = None
    hcore: np.ndarray | None = None
    eri: np.ndarray | None = None  # packed AO ERIs
    guess: np.ndarray | None = None  # AO Fock guess (guess_in.dat) or None
    calc_type: str = "CCSD(T)_spatial"
    scf_e_tol: float = 1e-6
    scf_d_tol: float = 1e-6
    scf_diis_n_errmat: int = 6
    ccsd_e_tol: float = 1e-6
    ccsd_t_tol: float = 1e-6
    ccsd_diis_n_errmat: int = 8
    scf_maxiter: int = 50
    ccsd_maxiter: int = 50
    scf_read_guess: bool = False
    # results
    e_hf: float = 0.0
    coeff: np.ndarray | None = None  # C[mo, ao]  (src/hf.f90:102,127)
    eps: np.ndarray | None = None
    eri_mo: np.ndarray | None = None  # packed MO ERIs
    e_mp2: float = 0.0
    log: dict = field(default_factory=dict)


def parse_els_in(text):
    """Namelist &elsinput reader (src/system.f90:96-114)."""
    out = {}
    for raw in text.splitlines():
        line = raw.strip().rstrip(",")
        if not line or line.startswith("&") or line.startswith("/"):
            continue
        if "=" not in line:
            continue
        k, v = [s.strip() for s in line.split("=", 1)]
        v = v.rstrip(",").strip()
        if v.startswith('"') or v.startswith("'"):
            out[k] = v.strip("\"'")
        elif v.lower() in (".true.", "t", ".t."):
            out[k] = True
        elif v.lower() in (".false.", "f", ".f."):
            out[k] = False
        else:
            try:
                out[k] = int(v)
            except ValueError:
                out[k] = float(v.lower().replace("d", "e"))
    return out


def _read_sym(path, n):
    m = np.zeros((n, n))
    d = np.loadtxt(path, ndmin=2)
    i = d[:, 0].astype(int) - 1
    j = d[:, 1].astype(int) - 1
    m[i, j] = d[:, 2]
    m[j, i] = d[:, 2]
    return m


def read_system(dirpath):
    """read_system_in + read_integrals_in + read_geometry_in (main.F90:36-39)."""
    sysm = System()
    with open(os.path.join(dirpath, "els.in")) as f:
        nml = parse_els_in(f.read())
    for k, v in nml.items():
        if hasattr(sysm, k):
            setattr(sysm, k, v)
    s = np.loadtxt(os.path.join(dirpath, "s.dat"), ndmin=2)
    n = int(max(s[:, 0].max(), s[:, 1].max()))  # src/integrals.f90:84-91
    sysm.nbasis = n
    sysm.ovlp = _read_sym(os.path.join(dirpath, "s.dat"), n)
    ke = _read_sym(os.path.join(dirpath, "t.dat"), n)
    en = _read_sym(os.path.join(dirpath, "v.dat"), n)
    sysm.hcore = ke + en  # src/integrals.f90:137
    d = np.loadtxt(os.path.join(dirpath, "eri.dat"), ndmin=2)
    ii = d[:, :4].astype(np.int64) - 1
    packed = np.zeros(npacked(n))
    packed[tri(tri(ii[:, 0], ii[:, 1]), tri(ii[:, 2], ii[:, 3]))] = d[:, 4]  # :152-160
    sysm.eri = packed
    with open(os.path.join(dirpath, "geom.dat")) as f:
        toks = f.read().split()
    nat = int(toks[0])
    g = np.array(toks[1 : 1 + 4 * nat], dtype=float).reshape(nat, 4)
    z = g[:, 0].astype(int)  # charges(i) = int(charge), src/geometry.f90:33
    xyz = g[:, 1:]
    sysm.nel = int(z.sum())
    sysm.nocc = sysm.nel // 2
    e_nuc = 0.0
    for j in range(1, nat):  # src/geometry.f90:85-89
        for i in range(j):
            e_nuc += z[i] * z[j] / np.linalg.norm(xyz[i] - xyz[j])
    sysm.e_nuc = e_nuc
    if sysm.scf_read_guess:
        gpath = os.path.join(dirpath, "guess_in.dat")
        d = np.loadtxt(gpath, ndmin=2)
        gm = np.zeros((n, n))
        gm[d[:, 0].astype(int) - 1, d[:, 1].astype(int) - 1] = d[:, 2]
        sysm.guess = gm
    return sysm


# --------------------------------------------------------------------------
# small dense solver used by both DIIS procedures -- src/linalg.fpp:38-56
# --------------------------------------------------------------------------


def diis_solve(B_lower, n):
    """Solve the (n+1)x(n+1) Pulay system from its lower triangle.

    The reference calls dsysv('L') (src/linalg.fpp:51) which reads only the
    lower triangle; we mirror it into a full symmetric matrix and use LU.
    """
    B = np.tril(B_lower) + np.tril(B_lower, -1).T
    rhs = np.zeros(n + 1)
    rhs[n] = -1.0
    return np.linalg.solve(B, rhs)


# --------------------------------------------------------------------------
# RHF -- src/hf.f90:21-151 (do_rhf), 197-242 (update_diis), 319-385
# --------------------------------------------------------------------------


def build_fock(hcore, g, D):
    """src/hf.f90:349-385: F = h + sum_kl D(k,l) [2 (ij|kl) - (ik|jl)]."""
    return hcore + 2.0 * np.einsum("ijkl,kl->ij", g, D) - np.einsum("ikjl,kl->ij", g, D)


def do_rhf(sysm: System):
    n, nocc = sysm.nbasis, sysm.nel // 2
    S, h = sysm.ovlp, sysm.hcore
    g = unpack_eri(sysm.eri, n)
    w, U = np.linalg.eigh(S)  # src/hf.f90:53-68
    X = U @ np.diag(1.0 / np.sqrt(w)) @ U.T
    F = sysm.guess.copy() if (sysm.scf_read_guess and sysm.guess is not None) else h.copy()
    nerr = sysm.scf_diis_n_errmat
    use_diis = nerr >= 2
    Fs = np.zeros((nerr, n, n)) if use_diis else None
    Es = np.zeros((nerr, n, n)) if use_diis else None
    it_slot, n_active = 0, 0
    energy, D_old = 0.0, np.zeros((n, n))
    table = []
    conv = False
    for it in range(1, sysm.scf_maxiter + 1):
        Fp = X.T @ F @ X  # :95
        eps, Cp = np.linalg.eigh(Fp)  # :98
        C = (X @ Cp).T  # :102  C[mo, ao]
        D = C[:nocc].T @ C[:nocc]  # :317
        e_old = energy
        energy = float(np.sum(D * (h + F)))  # :341
        rms = float(np.sqrt(np.sum((D - D_old) ** 2)))
        if rms < sysm.scf_d_tol and abs(energy - e_old) < sysm.scf_e_tol:
            conv = True
        D_old = D
        table.append((it, energy, energy - e_old, rms))
        if conv:
            break
        F = build_fock(h, g, D)  # :137
        if use_diis:  # :197-242
            it_slot += 1
            if it_slot > nerr:
                it_slot -= nerr
            if n_active < nerr:
                n_active += 1
            Fs[it_slot - 1] = F
            Es[it_slot - 1] = F @ D @ S - S @ D @ F
            na = n_active
            if na > 1:
                B = np.zeros((na + 1, na + 1))
                B[na, :] = -1.0
                B[na, na] = 0.0
                for i in range(na):
                    for j in range(i + 1):
                        B[i, j] = np.sum(Es[i] * Es[j])
                c = diis_solve(B, na)
                F = np.tensordot(c[:na], Fs[:na], axes=(0, 0))
    sysm.e_hf = energy
    sysm.coeff = C
    sysm.eps = eps
    sysm.log["scf"] = table
    sysm.log["scf_converged"] = conv
    return sysm


# --------------------------------------------------------------------------
# AO->MO transform + MP2 -- src/mp2.f90:261-449 (do_mp2_spatial)
# --------------------------------------------------------------------------


def ao2mo_dense(g_ao, C):
    """Four quarter transforms (src/mp2.f90:321-385); C[mo, ao]."""
    t = np.einsum("pi,ijkl->pjkl", C, g_ao, optimize=True)
    t = np.einsum("qj,pjkl->pqkl", C, t, optimize=True)
    t = np.einsum("rk,pqkl->pqrl", C, t, optimize=True)
    t = np.einsum("sl,pqrl->pqrs", C, t, optimize=True)
    return t


def ao2mo_packed(eri_packed, C):
    n = C.shape[0]
    g = ao2mo_dense(unpack_eri(eri_packed, n), C)
    return pack_eri(g)  # src/mp2.f90:388-410


def mp2_energy(eri_mo_packed, eps, nocc):
    """src/mp2.f90:418-438."""
    n = len(eps)
    g = unpack_eri(eri_mo_packed, n)
    o, v = slice(0, nocc), slice(nocc, n)
    iajb = g[o, v, o, v]  # (ia|jb)
    D = eps[o, None, None, None] + eps[None, None, o, None] - eps[None, v, None, None] - eps[None, None, None, v]
    return float(np.sum(iajb * (2.0 * iajb - iajb.transpose(0, 3, 2, 1)) / D))


def do_mp2_spatial(sysm: System):
    sysm.eri_mo = ao2mo_packed(sysm.eri, sysm.coeff)
    sysm.e_mp2 = mp2_energy(sysm.eri_mo, sysm.eps, sysm.nel // 2)
    return sysm


# --------------------------------------------------------------------------
# CC-DIIS -- src/ccsd.f90:577-676 (init_diis_cc_t, update_diis_cc)
# --------------------------------------------------------------------------


class CCDiis:
    def __init__(self, nerr, t1_shape, t2_shape):
        self.nerr = nerr
        self.use = nerr >= 2  # :593-595
        self.slot = 0
        self.n_active = 0
        if self.use:
            self.t1 = np.zeros((nerr,) + t1_shape)
            self.e1 = np.zeros((nerr,) + t1_shape)
            self.t2 = np.zeros((nerr,) + t2_shape)
            self.e2 = np.zeros((nerr,) + t2_shape)
        self.t1_s = None
        self.t2_s = None
        self.last_c = None

    def stash(self, t1, t2):  # :342-343 / :232-236
        if self.use:
            self.t1_s = t1.copy()
            self.t2_s = t2.copy()

    def update(self, t1, t2):
        if not self.use:
            return t1, t2
        self.slot += 1
        if self.slot > self.nerr:
            self.slot -= self.nerr
        if self.n_active < self.nerr:
            self.n_active += 1
        s = self.slot - 1
        self.t1[s] = t1
        self.t2[s] = t2
        self.e1[s] = t1 - self.t1_s
        self.e2[s] = t2 - self.t2_s
        n = self.n_active
        B = np.zeros((n + 1, n + 1))
        B[n, :] = -1.0
        B[n, n] = 0.0
        for i in range(n):
            for j in range(i + 1):
                B[i, j] = np.sum(self.e1[i] * self.e1[j]) + np.sum(self.e2[i] * self.e2[j])
        c = diis_solve(B, n)
        self.last_c = c
        t1n = np.tensordot(c[:n], self.t1[:n], axes=(0, 0))
        t2n = np.tensordot(c[:n], self.t2[:n], axes=(0, 0))
        return t1n, t2n


# --------------------------------------------------------------------------
# spin-free CCSD -- src/ccsd.f90:279-402, 404-575, 1040-1312, 1538-1732, 1734-1810
# --------------------------------------------------------------------------


def spatial_slices(eri_mo_packed, n, nocc):
    """init_cc slices (src/ccsd.f90:496-512): physicist <pq|rs> = (pr|qs)."""
    g = unpack_eri(eri_mo_packed, n)
    phys = g.transpose(0, 2, 1, 3)  # <pq|rs>[p,q,r,s] = (pr|qs)
    o, v = slice(0, nocc), slice(nocc, n)
    return {
        "v_oovv": np.ascontiguousarray(phys[o, o, v, v]),
        "v_ovov": np.ascontiguousarray(phys[o, v, o, v]),
        "v_vvov": np.ascontiguousarray(phys[v, v, o, v]),
        "v_oovo": np.ascontiguousarray(phys[o, o, v, o]),
        "v_oooo": np.ascontiguousarray(phys[o, o, o, o]),
        "v_vvvv": np.ascontiguousarray(phys[v, v, v, v]),
    }


def denominators(eps, nocc):
    eo, ev = eps[:nocc], eps[nocc:]
    D1 = eo[:, None] - ev[None, :]
    D2 = eo[:, None, None, None] + eo[None, :, None, None] - ev[None, None, :, None] - ev[None, None, None, :]
    return D1, D2


def restricted_intermediates(t1, t2, V):
    """update_restricted_intermediates (src/ccsd.f90:1040-1312), Piecuch Table 1."""
    v_oovv, v_ovov, v_vvov, v_oovo, v_oooo = V["v_oovv"], V["v_ovov"], V["v_vvov"], V["v_oovo"], V["v_oooo"]
    I = {}
    asym_t2 = 2.0 * t2 - t2.transpose(1, 0, 2, 3)  # :1063-1064
    c = t2 + np.einsum("ia,jb->ijab", t1, t1)  # :1071-1079
    A = 2.0 * v_oovv - v_oovv.transpose(0, 1, 3, 2)  # antisymmetrise '1243' :1089
    # I_vo(a,i) = sum_me A(i,m,a,e) t1(m,e)                               :1089-1092
    I["I_vo"] = np.einsum("imae,me->ai", A, t1)
    # I_vv(b,a): v_vvov antisym '2134' -> 2v(i,j,k,l)-v(j,i,k,l); reshape '2431'  :1101-1111
    Avv = 2.0 * v_vvov - v_vvov.transpose(1, 0, 2, 3)
    # reshape_tmp(j,l,k,i)=Avv(i,j,k,l): I_vv(b,a) = sum_{m,e} Avv(e,b,m,a) t1(m,e)
    I_vv = np.einsum("ebma,me->ba", Avv, t1)
    # reshape '4123': out(l,i,j,k)=A(i,j,k,l); I_vv(b,a) -= sum_{mne} A(m,n,e,b) c(m,n,e,a)
    I_vv = I_vv - np.einsum("mneb,mnea->ba", A, c)
    I["I_vv"] = I_vv
    # I_oo_p(j,i): v_oovo antisym '2134'; reshape '4213': out(l,j,i,k)=in(i,j,k,l)      :1121-1131
    Aoo = 2.0 * v_oovo - v_oovo.transpose(1, 0, 2, 3)
    # out(j',i',m,e)=Aoo(m,i',e,j'): I_oo_p(j,i) = sum_me Aoo(m,i,e,j) t1(m,e)
    I_oo_p = np.einsum("miej,me->ji", Aoo, t1)
    # reshape '1432': out(i,l,k,j)=v_oovv(i,j,k,l) -> tmp(m,f,e,i)=v_oovv(m,i,e,f);
    # dgemm: I_oo_p(j,i) += sum_{m,f,e}... A-matrix is asym_t2 viewed (o, o v v): asym_t2(j, m,f',e')
    # C(j,i) = sum_{m,x,y} asym_t2(j,m,x,y) tmp(m,x,y,i) = sum asym_t2(j,m,f,e) v_oovv(m,i,e,f)
    I_oo_p = I_oo_p + np.einsum("jmfe,mief->ji", asym_t2, v_oovv)
    I["I_oo_p"] = I_oo_p
    # I_oo(j,i) = I_oo_p + sum_e t1(j,e) I_vo(e,i)                                   :1136-1137
    I["I_oo"] = I_oo_p + t1 @ I["I_vo"]
    # I_oooo(k,l,i,j)                                                                :1143-1155
    I_oooo = v_oooo + np.einsum("klef,ijef->klij", c, v_oovv)
    scr = np.einsum("ke,ilej->klij", t1, v_oovo)  # scratch(k,j',i',l') with reshape '3214'
    I_oooo = I_oooo + scr + scr.transpose(1, 0, 3, 2)
    I["I_oooo"] = I_oooo
    # I_ovov(j,b,i,a)                                                                :1165-1191
    I_ovov = v_ovov - 0.5 * np.einsum("mibe,mjae->jbia", v_oovv, c)
    I_ovov = I_ovov - np.einsum("mibj,ma->jbia", v_oovo, t1)
    I_ovov = I_ovov + np.einsum("je,ebia->jbia", t1, v_vvov)
    I["I_ovov"] = I_ovov
    # I_voov(b,j,i,a)                                                                :1205-1252
    Av = 2.0 * v_oovv - v_oovv.transpose(0, 1, 3, 2)  # 2v(i,m,b,e)-v(i,m,e,b)
    I_voov = 0.5 * np.einsum("imbe,mjea->bjia", Av, t2)
    I_voov = I_voov - 0.5 * np.einsum("imbe,mjae->bjia", v_oovv, c)
    I_voov = I_voov - np.einsum("imbj,ma->bjia", v_oovo, t1)
    I_voov = I_voov + v_oovv.transpose(3, 0, 1, 2) + np.einsum("beia,je->bjia", v_vvov, t1)
    I["I_voov"] = I_voov
    # I_vovv_p(c,i,a,b)                                                              :1261-1272,1297-1299
    I_vovv_p = v_vvov.transpose(3, 2, 1, 0) - np.einsum("micb,ma->ciab", v_oovv, t1)
    I_vovv_p = I_vovv_p - np.einsum("maic,mb->ciab", v_ovov, t1)
    I["I_vovv_p"] = I_vovv_p
    # x_voov(b,j,i,a) = v_vvov(b,e,i,a) t1(j,e)                                      :1279-1290
    x_voov = np.einsum("beia,je->bjia", v_vvov, t1)
    I["x_voov"] = x_voov
    # I_ooov_p(j,k,i,a)                                                              :1306-1308
    I_ooov_p = v_oovo.transpose(1, 0, 3, 2) + np.einsum("jkef,efia->jkia", t2, v_vvov)
    I_ooov_p = I_ooov_p + np.einsum("je,ekia->jkia", t1, x_voov)
    I["I_ooov_p"] = I_ooov_p
    I["asym_t2"] = asym_t2
    I["c_oovv"] = c
    return I


def restricted_amplitudes(t1, t2, V, I, D1, D2):
    """update_amplitudes_restricted (src/ccsd.f90:1538-1732), Piecuch Eqs. 43-44."""
    v_oovv, v_ovov, v_vvov, v_oovo, v_vvvv = V["v_oovv"], V["v_ovov"], V["v_vvov"], V["v_oovo"], V["v_vvvv"]
    asym_t2, c = I["asym_t2"], I["c_oovv"]
    r1 = t1 @ I["I_vv"] - I["I_oo_p"] @ t1  # :1571-1572
    r1 = r1 + np.einsum("em,miea->ia", I["I_vo"], asym_t2)  # :1580-1589
    r1 = r1 + np.einsum("me,miea->ia", t1, 2.0 * v_oovv) - np.einsum("me,maie->ia", t1, v_ovov)
    # reshape '2143': tmp(j,i,l,k)=v_oovo(i,j,k,l); r1(i,a) -= sum tmp(i,m,n,e) asym_t2(m,n,e,a)   :1606-1607
    r1 = r1 - np.einsum("mien,mnea->ia", v_oovo, asym_t2)
    r1 = r1 + np.einsum("efma,mief->ia", v_vvov, asym_t2)  # :1618-1630
    X = np.einsum("ijae,eb->ijab", t2, I["I_vv"])  # :1647
    X = X - np.einsum("miba,jm->ijab", t2, I["I_oo"])  # :1654-1664
    X = X + 0.5 * np.einsum("ijef,efab->ijab", c, v_vvvv)  # :1669
    X = X + 0.5 * np.einsum("ijmn,mnab->ijab", I["I_oooo"], c)  # :1673
    X = X - np.einsum("mjae,iemb->ijab", t2, I["I_ovov"])  # :1680-1695
    X = X - np.einsum("iema,mjeb->ijab", I["I_ovov"], t2)
    X = X + np.einsum("miea,ejmb->ijab", asym_t2, I["I_voov"])
    X = X + np.einsum("ie,ejab->ijab", t1, I["I_vovv_p"])  # :1700
    X = X - np.einsum("ma,ijmb->ijab", t1, I["I_ooov_p"])  # :1705-1715
    X = X + X.transpose(1, 0, 3, 2) + v_oovv  # :1721-1722
    return r1 / D1, X / D2  # :1727-1728


def restricted_energy(t1, t2, v_oovv):
    """src/ccsd.f90:1774."""
    c = t2 + np.einsum("ia,jb->ijab", t1, t1)
    return float(np.sum((2.0 * v_oovv - v_oovv.transpose(0, 1, 3, 2)) * c))


def cr_intermediates(t1, t2, V, I_vo_stale, asym_t2_stale, q3a=True):
    """build_cr_ccsd_t_intermediates (src/ccsd.f90:2338-2551).

    q3a=True keeps the truncated `do e = 1, nocc` loop at :2535.  The stale
    I_vo / asym_t2 (Q3b) are whatever the caller passes.
    """
    v_oovv, v_ovov, v_vvov, v_oovo, v_oooo, v_vvvv = (
        V["v_oovv"], V["v_ovov"], V["v_vvov"], V["v_oovo"], V["v_oooo"], V["v_vvvv"])
    nocc, nvirt = t1.shape
    x_vvvo_p = v_vvov.transpose(1, 0, 3, 2) - 0.5 * np.einsum("ma,mibc->bcai", t1, v_oovv)  # :2425-2435
    x_ovov_p = v_ovov - 0.5 * np.einsum("mibj,ma->jbia", v_oovo, t1) + np.einsum("je,beai->jbia", t1, x_vvvo_p)
    x_voov_p = v_oovv.transpose(2, 1, 0, 3) - 0.5 * np.einsum("imbj,ma->bjia", v_oovo, t1) \
        + np.einsum("ebai,je->bjia", x_vvvo_p, t1)
    x_vvvo = x_vvvo_p - 0.5 * np.einsum("ma,mibc->bcai", t1, v_oovv)  # :2461-2471
    x_ovoo = v_oovo.transpose(3, 2, 1, 0) + np.einsum("ke,ijea->kaij", t1, v_oovv)  # :2473-2483
    x_ovov_pp = v_ovov - np.einsum("mibj,ma->jbia", v_oovo, t1) + 0.5 * np.einsum("je,beai->jbia", t1, x_vvvo)
    x_voov_pp = v_oovv.transpose(2, 1, 0, 3) - np.einsum("imbj,ma->bjia", v_oovo, t1) \
        + 0.5 * np.einsum("ebai,je->bjia", x_vvvo, t1)
    # I_vovv_pp(c,i,a,b)  :2509-2525
    Ivv = v_vvov.transpose(3, 2, 1, 0) + np.einsum("ecba,ie->ciab", v_vvvv, t1)
    Ivv = Ivv - np.einsum("icma,mb->ciab", x_ovov_p, t1) - np.einsum("ma,cimb->ciab", t1, x_voov_p)
    Ivv = Ivv - np.einsum("cm,miab->ciab", I_vo_stale, t2) + np.einsum("mnba,icmn->ciab", t2, x_ovoo)
    Ivv = Ivv + np.einsum("ceam,imbe->ciab", x_vvvo, asym_t2_stale)
    Ivv = Ivv - np.einsum("ecam,mieb->ciab", x_vvvo, t2) - np.einsum("miae,ecbm->ciab", t2, x_vvvo)
    # I_ooov_pp(j,k,i,a)  :2527-2544
    Ioo = v_oovo.transpose(1, 0, 3, 2) - np.einsum("mikj,ma->jkia", v_oooo, t1)
    Ioo = Ioo + np.einsum("jeia,ke->jkia", x_ovov_pp, t1) + np.einsum("je,ekia->jkia", t1, x_voov_pp)
    Ioo = Ioo + np.einsum("kjef,efai->jkia", t2, x_vvvo)
    ne = nocc if q3a else nvirt  # Q3a: `do e = 1, nocc` runs a *virtual* index only to nocc
    es = slice(0, min(ne, nvirt))
    Ioo = Ioo + np.einsum("jeim,mkea->jkia", x_ovoo[:, es], asym_t2_stale[:, :, es, :])
    Ioo = Ioo - np.einsum("jemi,mkea->jkia", x_ovoo[:, es], t2[:, :, es, :])
    Ioo = Ioo - np.einsum("mjae,kemi->jkia", t2[:, :, :, es], x_ovoo[:, es])
    return Ivv, Ioo


def t1_diagnostic(t1, nel):
    return float(np.sqrt(np.sum(t1 ** 2)) / np.sqrt(float(nel)))  # src/ccsd.f90:372


def ccsd_spatial(eri_mo_packed, eps, nocc, e_tol=1e-6, t_tol=1e-7, diis_n=8, maxiter=50,
                 want_cr=False, q3a=True, q3b=True, callback=None):
    """do_ccsd_spatial (src/ccsd.f90:279-402).  Returns a dict with the table and final state."""
    n = len(eps)
    V = spatial_slices(eri_mo_packed, n, nocc)
    D1, D2 = denominators(eps, nocc)
    t1 = np.zeros_like(D1)
    t2 = V["v_oovv"] / D2  # :520-521
    diis = CCDiis(diis_n, t1.shape, t2.shape)
    t2_old = np.zeros_like(t2)
    e_old = 0.0
    energy = restricted_energy(t1, t2, V["v_oovv"])
    rms = float(np.sum((t2 - t2_old) ** 2))
    t2_old = t2.copy()
    table = [("MP1", energy, energy - e_old, rms)]
    conv = False
    I = None
    for it in range(1, maxiter + 1):
        diis.stash(t1, t2)
        I = restricted_intermediates(t1, t2, V)
        t1, t2 = restricted_amplitudes(t1, t2, V, I, D1, D2)
        e_old = energy
        energy = restricted_energy(t1, t2, V["v_oovv"])
        rms = float(np.sum((t2 - t2_old) ** 2))  # Q4: squared norm printed
        t2_old = t2.copy()
        table.append((it, energy, energy - e_old, rms))
        if callback:
            callback(it, t1, t2, I)
        if np.sqrt(rms) < t_tol and abs(energy - e_old) < e_tol:  # :1805
            conv = True
            break
        t1, t2 = diis.update(t1, t2)
    out = {"table": table, "converged": conv, "e_ccsd": energy, "t1": t1, "t2": t2, "V": V,
           "t1_diag": t1_diagnostic(t1, 2 * nocc), "iterations": len(table) - 1}
    if want_cr:
        if q3b:
            I_vo_s, asym_s = I["I_vo"], I["asym_t2"]
        else:
            fresh = restricted_intermediates(t1, t2, V)
            I_vo_s, asym_s = fresh["I_vo"], fresh["asym_t2"]
        out["I_vovv_pp"], out["I_ooov_pp"] = cr_intermediates(t1, t2, V, I_vo_s, asym_s, q3a=q3a)
    return out


# --------------------------------------------------------------------------
# spin-free triples -- src/ccsd.f90:2018-2293 (do_ccsd_t_spatial), 2295-2336 (make_x_bar)
# --------------------------------------------------------------------------


def x_bar(x):
    """make_x_bar (src/ccsd.f90:2314-2318) on the last three axes: 4/3 x(abc) - 2 x(acb) + 2/3 x(cab)."""
    return 4.0 * x / 3.0 - 2.0 * np.swapaxes(x, -1, -2) + 2.0 * np.moveaxis(x, -3, -1) / 3.0


_S3 = [(0, 1, 2), (1, 0, 2), (2, 1, 0), (0, 2, 1), (1, 2, 0), (2, 0, 1)]


def _perm6(X):
    """Sum over the six simultaneous permutations of (i,a),(j,b),(k,c) (src/ccsd.f90:2168-2173)."""
    W = np.zeros_like(X)
    for p in _S3:
        W += X.transpose(p[0], p[1], p[2], 3 + p[0], 3 + p[1], 3 + p[2])
    return W


def triples_spatial_sums(t1, t2, v_oovv, v_vvov, v_oovo, eps, paren, renorm, comp_renorm,
                         I_vovv_pp=None, I_ooov_pp=None, triples=None):
    """The six accumulators of do_ccsd_t_spatial, for all o^3 (i,j,k) or a given list.

    Returns (e_T, e_TT, D_T, D_TT, e_CR, e_CRT) *without* the `1 + 2 sum t1^2 + ...`
    term of :2243 (see triples_spatial for the assembled energies).
    """
    nocc, nvirt = t1.shape
    eo, ev = eps[:nocc], eps[nocc:]
    doing_T, doing_R, doing_CR = paren, renorm, comp_renorm
    if triples is None:
        triples = [(i, j, k) for i in range(nocc) for j in range(nocc) for k in range(nocc)]
    acc = np.zeros(6)
    Dabc = -(ev[:, None, None] + ev[None, :, None] + ev[None, None, :])

    def conn(i, j, k, Iv, Io_is_v_oovo, Io):
        # Xc(abc) = sum_d t2(i,j,a,d) Iv(d,k,b,c) - sum_l t2(l,i,b,a) Io(...)
        x = np.einsum("ad,dbc->abc", t2[i, j], Iv[:, k])
        if Io_is_v_oovo:
            x -= np.einsum("lba,cl->abc", t2[:, i], Io[k, j])  # v_oovo(k,j,c,l)
        else:
            x -= np.einsum("lba,lc->abc", t2[:, i], Io[j, k])  # I_ooov_pp(j,k,l,c)
        return x

    v_vovv = v_vvov.transpose(3, 2, 1, 0)  # v_vovv(d,k,b,c)=v_vvov(c,b,k,d)  :2061

    def six(i, j, k, Iv, flag, Io):
        idx = (i, j, k)
        W = np.zeros((nvirt,) * 3)
        for p in _S3:
            pi, pj, pk = idx[p[0]], idx[p[1]], idx[p[2]]
            x = conn(pi, pj, pk, Iv, flag, Io)  # indexed (a',b',c') = permuted virtual labels
            # W(a,b,c) += x(perm(a,b,c)): x axes are (p0,p1,p2)-th of (a,b,c)
            W += x.transpose(np.argsort(p))
        return W

    for (i, j, k) in triples:
        D3 = eo[i] + eo[j] + eo[k] + Dabc
        W = six(i, j, k, v_vovv, True, v_oovo)
        t3 = W / D3
        tb = x_bar(t3)
        zb = None
        if doing_T:
            z3 = (t1[i][:, None, None] * v_oovv[j, k][None, :, :]
                  + t1[j][None, :, None] * v_oovv[i, k][:, None, :]
                  + t1[k][None, None, :] * v_oovv[i, j][:, :, None]) / D3
            # Q2: z3_bar only formed when (T) *and* (R or CR)  :2211-2215
            zb = x_bar(z3) if (doing_R or doing_CR) else np.zeros_like(z3)
        tmp = np.sum(tb * W)
        acc[0] += tmp
        if doing_T:
            acc[1] += tmp + np.sum(zb * W)
        if doing_CR:
            M = six(i, j, k, I_vovv_pp, False, I_ooov_pp)
            tmp = np.sum(tb * M)
            acc[4] += tmp
            if doing_T:
                acc[5] += tmp + np.sum(zb * M)
        if doing_R or doing_CR:
            y = (t1[i][:, None, None] * t1[j][None, :, None] * t1[k][None, None, :]
                 + t1[i][:, None, None] * t2[j, k][None, :, :]
                 + t1[j][None, :, None] * t2[i, k][:, None, :]
                 + t1[k][None, None, :] * t2[i, j][:, :, None])
            tmp = np.sum(tb * y)
            acc[2] += tmp
            if doing_T:
                acc[3] += tmp + np.sum(zb * y)
    return tuple(acc)


def triples_denominator_constant(t1, t2):
    """1 + 2 sum t1^2 + sum asym_t2 * c  (src/ccsd.f90:2243)."""
    asym = 2.0 * t2 - t2.transpose(1, 0, 2, 3)
    c = t2 + np.einsum("ia,jb->ijab", t1, t1)
    return float(1.0 + 2.0 * np.sum(t1 ** 2) + np.sum(asym * c))


def assemble_triples(e_ccsd, sums, const, paren, renorm, comp_renorm):
    """Energy assembly, src/ccsd.f90:2239-2276.  Returns dict of correlation energies."""
    e_T, e_TT, D_T, D_TT, e_CR, e_CRT = sums
    out = {}
    if renorm or comp_renorm:
        D_T += const
        if paren:
            D_TT += const
    out["e_ccsd_t"] = e_ccsd + e_T
    if paren:
        out["e_ccsd_tt"] = e_ccsd + e_TT
    if renorm or comp_renorm:
        out["e_rccsd_t"] = e_ccsd + e_T / D_T
        out["D_T"] = D_T
        if paren:
            out["e_rccsd_tt"] = e_ccsd + e_TT / D_TT
        if comp_renorm:
            out["e_crccsd_t"] = e_ccsd + e_CR / D_T
            out["D_TT"] = D_TT
            if paren:
                out["e_crccsd_tt"] = e_ccsd + e_CRT / D_TT
    return out


def triples_spatial(cc, eps, paren, renorm, comp_renorm):
    V = cc["V"]
    sums = triples_spatial_sums(cc["t1"], cc["t2"], V["v_oovv"], V["v_vvov"], V["v_oovo"], eps,
                                paren, renorm, comp_renorm, cc.get("I_vovv_pp"), cc.get("I_ooov_pp"))
    const = triples_denominator_constant(cc["t1"], cc["t2"]) if (renorm or comp_renorm) else 0.0
    return assemble_triples(cc["e_ccsd"], sums, const, paren, renorm, comp_renorm), sums


# --------------------------------------------------------------------------
# spin-orbital CCSD -- src/ccsd.f90:71-277, 678-1038 ; (T) 1812-1922
# --------------------------------------------------------------------------


def spinorb_antisym(eri_mo_packed, n):
    """<pq||rs> over 2n spin-orbitals, alpha=even index, beta=odd (src/ccsd.f90:111-143)."""
    g = unpack_eri(eri_mo_packed, n)
    phys = g.transpose(0, 2, 1, 3)  # <pq|rs> = (pr|qs)
    big = np.repeat(np.repeat(np.repeat(np.repeat(phys, 2, 0), 2, 1), 2, 2), 2, 3)
    s = np.arange(2 * n) % 2
    d = (s[:, None] == s[None, :]).astype(float)
    direct = big * d[:, None, :, None] * d[None, :, None, :]  # delta(p,r) delta(q,s)
    return direct - direct.transpose(0, 1, 3, 2)


def spinorb_slices(asym, nocc):
    o, v = slice(0, nocc), slice(nocc, asym.shape[0])
    names = ["oooo", "ooov", "ovoo", "oovo", "oovv", "ovvo", "ovvv", "vovv", "vvvv"]  # :182-194
    sl = {"o": o, "v": v}
    return {nm: np.ascontiguousarray(asym[tuple(sl[ch] for ch in nm)]) for nm in names}


def spinorb_symmetry_error(asym):
    """The run-time check at src/ccsd.f90:150-167, evaluated over the full index range (a superset)."""
    a = asym
    return float(np.abs(a + a.transpose(0, 1, 3, 2)).sum() + np.abs(a - a.transpose(2, 3, 0, 1)).sum()
                 + np.abs(a + a.transpose(3, 2, 0, 1)).sum() + np.abs(a - a.transpose(3, 2, 1, 0)).sum())


def spinorb_iteration(t1, t2, G, D1, D2, q1=True):
    """build_tau + build_F + build_W + update_amplitudes (src/ccsd.f90:678-1038)."""
    oooo, ooov, ovoo, oovo, oovv, ovvo, ovvv, vovv, vvvv = (
        G[k] for k in ["oooo", "ooov", "ovoo", "oovo", "oovv", "ovvo", "ovvv", "vovv", "vvvv"])
    x = np.einsum("ia,jb->ijab", t1, t1) - np.einsum("ib,ja->ijab", t1, t1)  # :701-711
    tau_t = t2 + 0.5 * x
    tau = t2 + x
    # build_F :717-797
    F_vv = np.einsum("mf,mafe->ae", t1, ovvv) + 0.5 * np.einsum("mnaf,mnfe->ae", tau_t, oovv)
    F_oo = -np.einsum("ne,nmie->mi", t1, ooov)
    dg = 0.5 * np.einsum("rnef,cnef->rc", tau_t, oovv)  # dgemm :795 lands at (row=i, col=m)
    F_oo = F_oo + (dg if q1 else dg.T)  # Q1: as coded the term is transposed w.r.t. Stanton Eq. 4
    F_ov = np.einsum("nf,mnef->me", t1, oovv)
    # build_W :799-905
    scr = np.einsum("mnie,je->mnij", ooov, t1)
    W_oooo = oooo + scr - scr.transpose(0, 1, 3, 2) + 0.5 * np.einsum("mnef,ijef->mnij", oovv, tau)
    W_ijmn = W_oooo.transpose(2, 3, 0, 1)  # stored (i,j,m,n) :841-842
    scr = np.einsum("mb,maef->baef", t1, ovvv)  # scratch(b,a,e,f) :850
    W_vvvv = vvvv + scr.transpose(1, 0, 2, 3) - scr  # :854
    W_efab = W_vvvv.transpose(2, 3, 0, 1)  # :856-857
    W_ovvo = np.einsum("mbef,jf->mbej", ovvv, t1) + ovvo  # :864-865
    W_ovvo = W_ovvo + np.einsum("nb,nmej->mbej", t1, oovo)  # :870-873
    scr = 0.5 * t2.transpose(1, 2, 0, 3) + np.einsum("jf,nb->nfjb", t1, t1)  # scratch(n,f,j,b) :885-893
    W_ovvo = W_ovvo - np.einsum("mnef,nfjb->mbej", oovv, scr)  # :897-901
    # update_amplitudes :907-1038
    r1 = np.einsum("ie,ae->ia", t1, F_vv) - np.einsum("mi,ma->ia", F_oo, t1)  # :939-941
    r1 = r1 + np.einsum("me,maei->ia", t1, ovvo) + np.einsum("miea,me->ia", t2, F_ov)
    r1 = r1 + 0.5 * np.einsum("mife,mafe->ia", t2, ovvv) - 0.5 * np.einsum("mnea,mnei->ia", t2, oovo)
    r1 = r1 / D1
    S = -np.einsum("ie,ma,mbej->ijab", t1, t1, ovvo, optimize=True) + np.einsum("miea,mbej->ijab", t2, W_ovvo)
    r2 = oovv + S - S.transpose(1, 0, 2, 3) - S.transpose(0, 1, 3, 2) + S.transpose(1, 0, 3, 2)  # :995-1000
    S = np.einsum("ijae,be->ijab", t2, F_vv)  # :1003
    r2 = r2 + S - S.transpose(0, 1, 3, 2)
    S = np.einsum("ijae,be->ijab", t2, t1.T @ F_ov)  # :1007
    r2 = r2 - 0.5 * (S - S.transpose(0, 1, 3, 2))
    S = np.einsum("im,mjab->ijab", t1 @ F_ov.T, t2)  # :1011
    r2 = r2 - 0.5 * (S - S.transpose(1, 0, 2, 3))
    S = np.einsum("ie,ejab->ijab", t1, vovv)  # :1015
    r2 = r2 + S - S.transpose(1, 0, 2, 3)
    S = np.einsum("ijam,mb->ijab", oovo, t1)  # :1019  tmp(i,j,b',a') with b' third
    r2 = r2 + S.transpose(0, 1, 3, 2) - S
    S = np.einsum("mi,mjab->ijab", F_oo, t2)  # :1024
    r2 = r2 - S + S.transpose(1, 0, 2, 3)
    r2 = r2 + 0.5 * np.einsum("ijmn,mnab->ijab", W_ijmn, tau)  # :1028
    r2 = r2 + 0.5 * np.einsum("ijef,efab->ijab", tau, W_efab)  # :1030
    return r1, r2 / D2


def spinorb_energy(t1, t2, oovv):
    """src/ccsd.f90:1793."""
    return float(0.25 * np.sum(oovv * (t2 + 2.0 * np.einsum("ia,jb->ijab", t1, t1))))


def ccsd_spinorb(eri_mo_packed, eps, nel, e_tol=1e-6, t_tol=1e-7, diis_n=8, maxiter=50, q1=True):
    """do_ccsd_spinorb (src/ccsd.f90:71-277).  nel = number of occupied spin-orbitals."""
    n = len(eps)
    asym = spinorb_antisym(eri_mo_packed, n)
    G = spinorb_slices(asym, nel)
    es = np.repeat(eps, 2)  # canon_levels_spinorb :460-463
    D1, D2 = denominators(es, nel)
    t1 = np.zeros_like(D1)
    t2 = G["oovv"] / D2  # :523
    diis = CCDiis(diis_n, t1.shape, t2.shape)
    t2_old = np.zeros_like(t2)
    e_old = 0.0
    energy = spinorb_energy(t1, t2, G["oovv"])
    rms = float(np.sum((t2 - t2_old) ** 2))
    t2_old = t2.copy()
    table = [("MP1", energy, energy - e_old, rms)]
    conv = False
    for it in range(1, maxiter + 1):
        diis.stash(t1, t2)
        t1, t2 = spinorb_iteration(t1, t2, G, D1, D2, q1=q1)
        e_old = energy
        energy = spinorb_energy(t1, t2, G["oovv"])
        rms = float(np.sum((t2 - t2_old) ** 2))
        t2_old = t2.copy()
        table.append((it, energy, energy - e_old, rms))
        if np.sqrt(rms) < t_tol and abs(energy - e_old) < e_tol:
            conv = True
            break
        t1, t2 = diis.update(t1, t2)
    return {"table": table, "converged": conv, "e_ccsd": energy, "t1": t1, "t2": t2, "G": G,
            "eps_so": es, "iterations": len(table) - 1}


def triples_spinorb(t1, t2, oovv, vovv, ovoo, eps_so, triples=None):
    """do_ccsd_t_spinorb (src/ccsd.f90:1812-1922): e_T over all o^3 (i,j,k), /36."""
    nocc, nvirt = t1.shape
    eo, ev = eps_so[:nocc], eps_so[nocc:]
    Dabc = -(ev[:, None, None] + ev[None, :, None] + ev[None, None, :])
    if triples is None:
        triples = [(i, j, k) for i in range(nocc) for j in range(nocc) for k in range(nocc)]

    def pabc(x):  # x - x(bac) - x(cba)  :1897-1907
        return x - x.transpose(1, 0, 2) - x.transpose(2, 1, 0)

    e_T = 0.0
    for (i, j, k) in triples:
        D3 = eo[i] + eo[j] + eo[k] + Dabc
        # vvoo(b,c,j,k) = oovv(j,k,b,c)
        t3d = (t1[i][:, None, None] * oovv[j, k][None] - t1[j][:, None, None] * oovv[i, k][None]
               - t1[k][:, None, None] * oovv[j, i][None]) / D3
        # t2_reshape(f,a,k,j) = t2(j,k,a,f)
        t3c = (np.einsum("fbc,af->abc", vovv[:, i], t2[j, k]) - np.einsum("fbc,af->abc", vovv[:, j], t2[i, k])
               - np.einsum("fbc,af->abc", vovv[:, k], t2[j, i]))
        t3c += (-np.einsum("mcb,ma->abc", t2[:, i], ovoo[:, :, j, k]) + np.einsum("mcb,ma->abc", t2[:, j], ovoo[:, :, i, k])
                + np.einsum("mcb,ma->abc", t2[:, k], ovoo[:, :, j, i]))
        t3cd = t3c / D3
        e_T += np.sum(pabc(t3c) * (pabc(t3cd) + pabc(t3d))) / 36.0
    return float(e_T)


# --------------------------------------------------------------------------
# whole program -- src/main.F90:36-120
# --------------------------------------------------------------------------


def run(sysm: System, q1=True, q3a=True, q3b=True):
    level, restricted, paren, renorm, comp_renorm = CALC_TYPES[sysm.calc_type]
    do_rhf(sysm)
    res = {"e_hf": sysm.e_hf, "e_nuc": sysm.e_nuc, "scf": sysm.log["scf"]}
    if level == "HF":
        return res
    do_mp2_spatial(sysm)
    res["e_mp2"] = sysm.e_mp2
    if level == "MP2":
        return res
    nocc = sysm.nel // 2
    if restricted:
        cc = ccsd_spatial(sysm.eri_mo, sysm.eps, nocc, sysm.ccsd_e_tol, sysm.ccsd_t_tol, sysm.ccsd_diis_n_errmat,
                          sysm.ccsd_maxiter, want_cr=comp_renorm, q3a=q3a, q3b=q3b)
        res.update(ccsd=cc["table"], e_ccsd=cc["e_ccsd"], t1_diag=cc["t1_diag"])
        if level == "CCSD(T)":
            en, sums = triples_spatial(cc, sysm.eps, paren, renorm, comp_renorm)
            res.update(en)
            res["triples_sums"] = sums
    else:
        cc = ccsd_spinorb(sysm.eri_mo, sysm.eps, sysm.nel, sysm.ccsd_e_tol, sysm.ccsd_t_tol,
                          sysm.ccsd_diis_n_errmat, sysm.ccsd_maxiter, q1=q1)
        res.update(ccsd=cc["table"], e_ccsd=cc["e_ccsd"])
        if level == "CCSD(T)":
            G = cc["G"]
            e_t = triples_spinorb(cc["t1"], cc["t2"], G["oovv"], G["vovv"], G["ovoo"], cc["eps_so"])
            res["e_ccsd_t"] = cc["e_ccsd"] + e_t
    res["cc"] = cc
    return res
