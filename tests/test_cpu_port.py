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


def test_cpu_port_full_ccsd_iteration_and_ao2mo_match_numpy_oracle():
    """The complete spin-free iteration (same dgemm calls / reshapes / loops as src/ccsd.f90:1040-1312, 1538-1732) and
    the four quarter transforms + repack of src/mp2.f90:321-410."""
    n, o = 14, 3
    eri, Cm, eps = synthetic.make(n, o, seed=3)
    lib = cpu_port.load()
    assert cpu_port.set_threads(lib, 2) == 2
    mo_ref = orc.ao2mo_packed(eri, Cm)
    mo, times = cpu_port.ao2mo(lib, eri, Cm)
    assert np.max(np.abs(mo - mo_ref)) < 1e-13 and times.shape == (5,)
    V = orc.spatial_slices(mo_ref, n, o)
    Vc = cpu_port.slices(lib, mo_ref, n, o)
    for k in V:
        assert np.array_equal(np.asarray(Vc[k]), V[k]), k
    D1, D2 = orc.denominators(eps, o)
    rng = np.random.default_rng(0)
    t1 = rng.standard_normal((o, n - o)) * 1e-2
    t2 = V["v_oovv"] / D2
    I = orc.restricted_intermediates(t1, t2, V)
    r1, r2 = orc.restricted_amplitudes(t1, t2, V, I, D1, D2)
    g1, g2, parts, wall = cpu_port.ccsd_iter(lib, Vc, eps, t1, t2)
    assert np.max(np.abs(g1 - r1)) < 1e-13 and np.max(np.abs(g2 - r2)) < 1e-13
    assert len(parts) == len(cpu_port.CCSD_ITER_PARTS) and wall > 0
    # the integral slices are antisymmetrised / restored in place inside the call (as in the reference, to rounding)
    g1b, g2b, _, _ = cpu_port.ccsd_iter(lib, Vc, eps, t1, t2)
    assert np.max(np.abs(g1 - g1b)) < 1e-15 and np.max(np.abs(g2 - g2b)) < 1e-15


def test_cpu_port_cr_triples_match_numpy_oracle():
    n, o = 12, 3
    eri, Cm, eps = synthetic.make(n, o, seed=5)
    mo = orc.ao2mo_packed(eri, Cm)
    cc = orc.ccsd_spatial(mo, eps, o, 1e-9, 1e-10, 8, 50, want_cr=True)
    V = cc["V"]
    lib = cpu_port.load()
    ijk = [(i, j, k) for i in range(o) for j in range(o) for k in range(o)]
    sums, _ = cpu_port.triples_cr(lib, cc["t1"], cc["t2"], V["v_oovv"], V["v_vvov"], V["v_oovo"], cc["I_vovv_pp"],
                                  cc["I_ooov_pp"], eps, ijk, True)
    ref = orc.triples_spatial_sums(cc["t1"], cc["t2"], V["v_oovv"], V["v_vvov"], V["v_oovo"], eps, True, False, True,
                                   cc["I_vovv_pp"], cc["I_ooov_pp"])
    assert np.max(np.abs(sums - np.array(ref))) < 1e-13


def test_pin_recipes_from_the_factored_integrals_equal_the_general_oracle():
    """tests/golden/make_bench_pins.py computes CPU pins for shapes where the dense slices cannot be built (nbf=400) straight
    from the factored form of the synthetic integrals: the first CCSD iteration with the t1 = 0 terms dropped and the ladder
    integrals in slabs, MP2 from the (ia|jb) block, and single-orbit (T) values on the MP1 amplitudes.  At a small shape each
    recipe must equal the general oracle functions on the slices of the packed MO integrals."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_bench_pins as M

    n, o = 37, 5
    mo, C, eps = cpu_port.synthetic_mo_integrals(n, o)
    V = orc.spatial_slices(mo, n, o)
    D1, D2 = orc.denominators(eps, o)
    t2 = V["v_oovv"] / D2
    t1 = np.zeros_like(D1)
    I = orc.restricted_intermediates(t1, t2, V)
    t1n, t2n = orc.restricted_amplitudes(t1, t2, V, I, D1, D2)
    e, rms, e_mp2 = M.ccsd_iter1_from_factors(n, o, ladder_block=5)
    assert abs(e - orc.restricted_energy(t1n, t2n, V["v_oovv"])) < 1e-14
    assert abs(rms - float(np.sum((t2n - t2) ** 2))) < 1e-15
    assert abs(e_mp2 - orc.mp2_energy(mo, eps, o)) < 1e-14 and abs(M.mp2_from_factors(n, o) - e_mp2) < 1e-14
    picks = [(0, 0, 0), (0, 0, 3), (1, 2, 4), (2, 4, 4)]
    vals, _ = M.mp1_triples_from_factors(n, o, picks)
    for t, x in zip(picks, vals):
        assert abs(x - orc.triples_bracket_T_orbit_form(t2, V["v_vvov"], V["v_oovo"], eps, triples=[t])) < 1e-16
    assert M.unique_triples(3) == [(0, 0, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1), (0, 1, 2), (0, 2, 2), (1, 1, 1), (1, 1, 2),
                                   (1, 2, 2), (2, 2, 2)]
