"""The algebra behind the single-GEMM form of the spin-orbital (T) (DESIGN.md section 9, item 1;
tools/patches/spinorb_triples_three_segment_gemm.patch), checked on the CPU with the oracle's tensors.

The reference builds the connected block from six products (src/ccsd.f90:1881-1889)

    X(a,b,c) = sum_f [ t2(j,k,a,f) <fi||bc> - t2(i,k,a,f) <fj||bc> - t2(j,i,a,f) <fk||bc> ]
             + sum_m [ <ma||jk> t2(m,i,b,c) - <ma||ik> t2(m,j,b,c) - <ma||ji> t2(m,k,b,c) ].

With the K-concatenated operands Acat(a,[f|m];p,q) = [t2(p,q,a,f) | <ma||pq>] and Bcat([f|m],(b,c);r) =
[<fr||bc> ; t2(m,r,b,c)], both minus signs are absorbed by the antisymmetry of Acat in (p,q), so that
X = Acat(j,k) Bcat(i) + Acat(k,i) Bcat(j) + Acat(i,j) Bcat(k): one GEMM of depth 3 (v + o), written once."""
import numpy as np

from afesp_b200 import synthetic
from oracle import afesp_oracle as orc


def test_three_segment_form_equals_the_reference_six_term_form():
    n, nocc = 8, 2
    eri, Cm, eps = synthetic.make(n, nocc, seed=9)
    mo = orc.ao2mo_packed(eri, Cm)
    G = orc.spinorb_slices(orc.spinorb_antisym(mo, n), 2 * nocc)
    o, v = 2 * nocc, 2 * (n - nocc)
    rng = np.random.default_rng(1)
    t2 = rng.standard_normal((o, o, v, v))
    t2 = t2 - t2.transpose(1, 0, 2, 3)
    t2 = t2 - t2.transpose(0, 1, 3, 2)          # antisymmetric in (i,j) and in (a,b), like converged amplitudes
    vovv, ovoo = G["vovv"], G["ovoo"]
    Acat = np.concatenate([t2.transpose(2, 3, 0, 1), ovoo.transpose(1, 0, 2, 3)], axis=1)          # (a, [f|m], p, q)
    Bcat = np.concatenate([vovv.transpose(0, 2, 3, 1), t2.transpose(0, 2, 3, 1)], axis=0)          # ([f|m], b, c, r)
    assert np.max(np.abs(Acat + Acat.transpose(0, 1, 3, 2))) < 1e-14                               # antisymmetry in (p,q)
    for (i, j, k) in [(0, 1, 2), (1, 2, 3), (0, 2, 3), (0, 1, 3)]:
        X = (np.einsum("af,fbc->abc", t2[j, k], vovv[:, i]) - np.einsum("af,fbc->abc", t2[i, k], vovv[:, j])
             - np.einsum("af,fbc->abc", t2[j, i], vovv[:, k])
             + np.einsum("ma,mbc->abc", ovoo[:, :, j, k], t2[:, i]) - np.einsum("ma,mbc->abc", ovoo[:, :, i, k], t2[:, j])
             - np.einsum("ma,mbc->abc", ovoo[:, :, j, i], t2[:, k]))
        Y = (np.einsum("ax,xbc->abc", Acat[:, :, j, k], Bcat[..., i]) + np.einsum("ax,xbc->abc", Acat[:, :, k, i], Bcat[..., j])
             + np.einsum("ax,xbc->abc", Acat[:, :, i, j], Bcat[..., k]))
        assert np.max(np.abs(X - Y)) < 1e-13
