"""The C port of the reference's dominant CPU loops (oracle/cpu_kernels.c, the timed CPU baseline) against the
NumPy oracle on a small synthetic system."""
import numpy as np

from afesp_b200 import synthetic
from oracle import afesp_oracle as orc
from oracle import cpu_port


def test_cpu_port_ring_and_triples_match_numpy_oracle():
    n, o = 14, 3
    eri, Cm, eps = synthetic.make(n, o, seed=3)
    mo = orc.ao2mo_packed(eri, Cm)
    cc = orc.ccsd_spatial(mo, eps, o, 1e-9, 1e-10, 8, 50)
    V, t1, t2 = cc["V"], cc["t1"], cc["t2"]
    I = orc.restricted_intermediates(t1, t2, V)
    lib = cpu_port.load()
    got, _ = cpu_port.ring(lib, t2, I["I_ovov"], I["asym_t2"], I["I_voov"])
    want = (-np.einsum("mjae,iemb->ijab", t2, I["I_ovov"]) - np.einsum("iema,mjeb->ijab", I["I_ovov"], t2)
            + np.einsum("miea,ejmb->ijab", I["asym_t2"], I["I_voov"]))
    assert np.max(np.abs(got - want)) < 1e-13
    lad, _ = cpu_port.ladder(I["c_oovv"], V["v_vvvv"])
    assert np.max(np.abs(lad - 0.5 * np.einsum("ijef,efab->ijab", I["c_oovv"], V["v_vvvv"]))) < 1e-13
    ijk = [(i, j, k) for i in range(o) for j in range(o) for k in range(o)]
    sums, _ = cpu_port.triples(lib, t1, t2, V["v_oovv"], V["v_vvov"], V["v_oovo"], eps, ijk, True, True)
    ref = orc.triples_spatial_sums(t1, t2, V["v_oovv"], V["v_vvov"], V["v_oovo"], eps, True, True, False)
    assert np.max(np.abs(sums - np.array(ref[:4]))) < 1e-13
