"""Seeded synthetic inputs of a named (nbf, nocc) shape for the hot path.

    (ij|kl) = sum_P B_ij^P B_kl^P,  B_ij^P = B_ji^P ~ N(0, s^2) exp(-|i-j|/w),  naux = nbf

gives 8-fold symmetric, positive semi-definite two-electron integrals in the reference's packed order.  The MO basis is
a seeded random orthogonal C(mo,ao) (so the AO->MO transform does real work) and the orbital energies are o levels in
[-2,-0.5] and v levels in [0.5,3] (gap >= 1 Eh).  s is chosen so that |E_MP2|/nocc is of order 0.02 Eh, which keeps
the CCSD iteration convergent.  No SCF is run: the CC equations only see (eri, C, eps), exactly what the hot path
receives from the host program.
"""
from __future__ import annotations

import numpy as np


def make_factors(nbf: int, nocc: int, seed: int = 20260, w: float = 6.0, target_emp2_per_occ: float = 0.02):
    """Returns (B[npair, naux], C[mo,ao], eps): the same system as make(), with the ERIs left in factored form
    (ij|kl) = sum_P B[ij,P] B[kl,P] so that large shapes can be expanded on the device (afesp_gpu_synth_eri_ao)."""
    rng = np.random.default_rng(seed)
    n, o, v = nbf, nocc, nbf - nocc
    naux = n
    ii, jj = np.tril_indices(n)
    damp = np.exp(-(ii - jj) / w)
    B = rng.standard_normal((ii.size, naux)) * damp[:, None]
    full_ms = (2.0 * np.sum(B * B) - np.sum(B[ii == jj] ** 2)) / (n * n * naux)
    sigma4 = 3.0 * target_emp2_per_occ / (o * v * v * naux)
    B *= (sigma4 ** 0.25) / np.sqrt(full_ms)
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    eps = np.concatenate([np.linspace(-2.0, -0.5, o), np.linspace(0.5, 3.0, v)])
    eps += 0.01 * rng.standard_normal(n)
    eps = np.sort(eps)
    return B, np.ascontiguousarray(q.T), eps


def make(nbf: int, nocc: int, seed: int = 20260, w: float = 6.0, target_emp2_per_occ: float = 0.02):
    """Returns (eri_ao_packed, C[mo,ao], eps)."""
    rng = np.random.default_rng(seed)
    n, o, v = nbf, nocc, nbf - nocc
    naux = n
    ii, jj = np.tril_indices(n)  # pair order p = i(i+1)/2 + j, i >= j (row-major lower triangle)
    npair = ii.size
    damp = np.exp(-(ii - jj) / w)
    B = rng.standard_normal((npair, naux)) * damp[:, None]
    # sigma_eff^2 = mean square of the full symmetric B_ij^P; E_MP2/o ~ o v^2 naux sigma_eff^4 / 3
    full_ms = (2.0 * np.sum(B * B) - np.sum(B[ii == jj] ** 2)) / (n * n * naux)
    sigma4 = 3.0 * target_emp2_per_occ / (o * v * v * naux)
    B *= (sigma4 ** 0.25) / np.sqrt(full_ms)
    G = B @ B.T
    packed = np.empty(npair * (npair + 1) // 2)
    pos = 0
    for r in range(npair):
        packed[pos:pos + r + 1] = G[r, :r + 1]
        pos += r + 1
    del G
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    eps = np.concatenate([np.linspace(-2.0, -0.5, o), np.linspace(0.5, 3.0, v)])
    eps += 0.01 * rng.standard_normal(n)
    eps = np.sort(eps)
    return packed, np.ascontiguousarray(q.T), eps
