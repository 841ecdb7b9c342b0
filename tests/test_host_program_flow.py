"""CPU tests of the host-side driver (afesp_b200/host.py: src/main.F90 + the iteration loops of src/ccsd.f90) with the
engine replaced by a test double over the NumPy oracle (tests/_oracle_engine.py).

What is under test is everything the HOST does once the hot path is entered -- call order, the convergence test of
src/ccsd.f90:1805, when DIIS is (not) called, the printed iteration table, the assembly of the nine triples energies
(:2239-2276), the 'Final energy breakdown' block that utils/els_wrapper.py:99-128 parses -- against the reference's own
shipped els.out.  The GPU parity tests (tests/test_gpu_parity.py) run the same host code over the CUDA library.
"""
import numpy as np
import pytest

from afesp_b200 import host
from tests._fixtures import compare_els_out, golden, golden_els_out, load_els_input
from tests._oracle_engine import OracleEngine


def _wrapper_energies(stdout):
    """The parsing rule of utils/els_wrapper.py:99-128 (`run_els`): substring match, last blank-separated token."""
    keys = ["RHF energy:", "MP2 energy:", " CCSD energy:", " CCSD[T] energy:", " CCSD(T) energy:", " R-CCSD[T] energy:",
            " R-CCSD(T) energy:", " CR-CCSD[T] energy:", " CR-CCSD(T) energy:", " T1 diagnostic:", " D[T]:", " D(T):"]
    energy = np.zeros(12)
    for line in stdout.split("\n"):
        for k, key in enumerate(keys):
            if key in line:
                energy[k] = float(line.split(" ")[-1])
    return energy


@pytest.fixture(scope="module")
def runs():
    cache = {}

    def get(name, calc_type=None, **kw):
        key = (name, calc_type, tuple(sorted(kw.items())))
        if key not in cache:
            inp = load_els_input(name, calc_type)
            for k, v in kw.items():
                setattr(inp, k, v)
            eng = OracleEngine()
            cache[key] = (inp, eng, host.run(inp, gpu=eng))
        return cache[key]

    return get


@pytest.mark.parametrize("name", ["n2", "f2"])
def test_whole_program_output_on_the_cpu_double_matches_shipped_els_out(runs, name):
    """CRCCSD(T)_spatial as shipped: every line of els.out (dates / times masked), numbers within 2 units of the last
    printed digit or 1e-9 Eh."""
    inp, eng, res = runs(name)
    diffs = compare_els_out(res.stdout, golden_els_out(name), ulps=2.0, abs_tol=1e-9)
    assert not diffs, "\n".join(diffs[:20])


@pytest.mark.parametrize("name", ["n2", "f2"])
def test_host_drives_the_engine_in_the_reference_order(runs, name):
    """src/ccsd.f90:339-396: iterate, test convergence, and only if NOT converged extrapolate; finalize once; (T) once."""
    inp, eng, res = runs(name)
    g = golden()[name]
    n_it = len(g["ccsd"]) - 1   # table rows minus the MP1 line
    expect = ["ao2mo", "mp2_energy", "ccsd_init"] + ["ccsd_iterate", "ccsd_diis"] * (n_it - 1) + \
             ["ccsd_iterate", "ccsd_finalize", "ccsd_t_spatial"]
    assert eng.calls == expect
    assert res.ccsd_converged and len(res.ccsd_table) == n_it + 1


@pytest.mark.parametrize("name", ["n2", "f2"])
def test_final_block_is_parsed_by_the_reference_wrapper_rule(runs, name):
    inp, eng, res = runs(name)
    mine = _wrapper_energies(res.stdout)
    ref = _wrapper_energies(golden_els_out(name))
    assert np.all(ref[:9] < 0) and ref[9] > 0 and ref[10] > 1 and ref[11] > 1   # the rule picks up all twelve numbers
    assert np.max(np.abs(mine - ref)) < 1e-9


def test_ccsd_that_does_not_converge_skips_the_triples(runs):
    """ccsd_maxiter exhausted: the loop of src/ccsd.f90:339-396 ends after a DIIS call, no 'Convergence reached' line, and
    the host does not enter (T) with unconverged amplitudes (the reference would read an unallocated int_store_cc)."""
    inp, eng, res = runs("n2", None, ccsd_maxiter=3)
    assert not res.ccsd_converged
    assert eng.calls == ["ao2mo", "mp2_energy", "ccsd_init"] + ["ccsd_iterate", "ccsd_diis"] * 3 + ["ccsd_finalize"]
    cc_part = res.stdout.split(" CCSD\n ----------\n", 1)[1]
    assert "Convergence reached" not in cc_part and "Restricted completely renormalised" not in cc_part
    assert len(res.ccsd_table) == 4


@pytest.mark.parametrize("calc,labels", [
    ("MP2_spatial", ["MP2"]),
    ("CCSD_spatial", ["MP2", "CCSD"]),
    ("CCSD[T]_spatial", ["MP2", "CCSD", "CCSD[T]"]),
    ("CCSD(T)_spatial", ["MP2", "CCSD", "CCSD[T]", "CCSD(T)"]),
    ("RCCSD[T]_spatial", ["MP2", "CCSD", "CCSD[T]", "R-CCSD[T]"]),
    ("RCCSD(T)_spatial", ["MP2", "CCSD", "CCSD[T]", "CCSD(T)", "R-CCSD[T]", "R-CCSD(T)"]),
    ("CRCCSD[T]_spatial", ["MP2", "CCSD", "CCSD[T]", "R-CCSD[T]", "CR-CCSD[T]"]),
])
def test_calc_types_print_the_blocks_main_F90_prints(runs, oracle_runs, calc, labels):
    """src/main.F90:123-175 per calc_type on the water sample: which energy pairs appear, in which order, and their values
    against the oracle's own whole-program run."""
    inp, eng, res = runs("h2o", calc)
    block = res.stdout.split(" Final energy breakdown\n", 1)[1]
    found = [ln.split(" correlation energy:")[0].strip() for ln in block.splitlines() if " correlation energy:" in ln]
    assert found == labels
    _, ref = oracle_runs("h2o", calc)
    assert abs(res.e_mp2 - ref["e_mp2"]) < 1e-12
    if "CCSD" in labels:
        assert abs(res.e_ccsd - ref["e_ccsd"]) < 1e-12 and len(res.ccsd_table) == len(ref["ccsd"])
        assert ("T1 diagnostic:" in block)
    for k in ["e_ccsd_t", "e_ccsd_tt", "e_rccsd_t", "e_rccsd_tt", "e_crccsd_t", "D_T"]:
        if k in ref:
            assert abs(res.energies[k] - ref[k]) < 1e-12, k
    assert ("D[T]:" in block) == any(l.startswith(("R-", "CR-")) for l in labels)
    if calc == "MP2_spatial":
        assert eng.calls == ["ao2mo", "mp2_energy"]


def test_spin_orbital_program_flow(runs, oracle_runs):
    """CCSD(T)_spinorb on the water sample (BASELINE.json configs[0]): banners of src/ccsd.f90:106-220 incl. the symmetry
    check, the unrestricted labels of main.F90:66-80,157-159, energies against the oracle's whole-program run."""
    inp, eng, res = runs("h2o", "CCSD(T)_spinorb")
    _, ref = oracle_runs("h2o", "CCSD(T)_spinorb")
    assert len(res.ccsd_table) == len(ref["ccsd"]) and abs(res.e_ccsd - ref["e_ccsd"]) < 1e-12
    assert abs(res.energies["e_ccsd_t"] - ref["e_ccsd_t"]) < 1e-12
    for text in [" Forming antisymmetrised spinorbital ERIs...",
                 " Checking that the permuational symmetry of the antisymmetrised integrals hold...",
                 " Forming slices of antisymmetrised spinorbital ERIs",
                 " Unrestricted CCSD(T) correlation energy (Hartree):", " Time taken for unrestricted CCSD:",
                 " Time taken for unrestricted CCSD(T):"]:
        assert text in res.stdout, text
    block = res.stdout.split(" Final energy breakdown\n", 1)[1]
    assert "CCSD[T]" not in block and "T1 diagnostic" not in block and " CCSD(T) energy:" in block
    assert eng.calls[-2:] == ["ccsd_finalize", "ccsd_t_spinorb"]


def test_whole_program_spin_orbital_output_matches_the_reference_els_cpu_out(runs):
    """sample_data/h2o-cc-pvtz/2.00_104.45/els_cpu.out (current code version, CCSD(T)_spinorb, 116 spin-orbitals; two-electron
    integrals regenerated by afesp_b200/gint.py): every line of the program output -- the spin-orbital counts of the system
    block (src/geometry.f90:40-46), the banners of src/ccsd.f90:106-220, the 19-row iteration table, E[CCSD(T)], the
    unrestricted final block and the 'Total execution time' line -- with numbers within 2 units of the last printed digit
    or 1e-9 Eh.  About 40 s."""
    import os

    from tests._fixtures import GOLDEN_DIR

    inp, eng, res = runs("h2o_tz", "CCSD(T)_spinorb")
    ref = open(os.path.join(GOLDEN_DIR, "h2o_tz_els_cpu_out.txt")).read()
    diffs = compare_els_out(res.stdout, ref, ulps=2.0, abs_tol=1e-9)
    assert not diffs, "\n".join(diffs[:20])
    assert " Number of occupied orbitals: 10\n Number of virtual orbitals: 106\n" in res.stdout
    assert res.stdout.rstrip().splitlines()[-1].startswith(" Total execution time:")
    mine, want = _wrapper_energies(res.stdout), _wrapper_energies(ref)
    assert want[4] < 0 and np.max(np.abs(mine - want)) < 1e-9


def test_symmetry_assertion_block_is_printed_in_the_reference_format():
    """Status 5 from ccsd_init: the banners, then 'Permutational symmetry error:' in Fortran E15.6 (mantissa in [0.1,1):
    0.350000E-06, not C's 3.500000E-07), nothing after it (src/ccsd.f90:161-167); the exception carries the output so far."""
    from afesp_b200.capi import AfespError

    inp = load_els_input("h2o", "CCSD_spinorb")
    with pytest.raises(AfespError) as ei:
        host.run(inp, gpu=OracleEngine(force_symmetry_error=3.5e-7))
    assert ei.value.code == 5
    tail = ei.value.stdout.splitlines()[-4:]
    assert tail[0].startswith(" Time taken:") and tail[1] == ""
    assert tail[2] == " Checking that the permuational symmetry of the antisymmetrised integrals hold..."
    assert tail[3] == " Permutational symmetry error:    0.350000E-06"
