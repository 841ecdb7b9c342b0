"""CPU tests of the C++ host program (host/els_host.cpp, the stand-in for the Fortran els.x) beyond the SCF section.

els_host links the C ABI of include/afesp_gpu.h.  Here a TEST DOUBLE of the handful of entry points it calls
(tests/_double/afesp_gpu_double.c: every call is handed to the NumPy oracle through an embedded interpreter) is built into
a temporary directory and LD_PRELOADed in front of libafesp_gpu.so, so that the host program's own logic -- the
reference's iteration loop, convergence test and DIIS call order (src/ccsd.f90:339-396), every printed line of
src/main.F90 / src/ccsd.f90, the assembly of the triples energies (:2239-2276), the error block
(src/error_handling.f90:6-20) -- is compared with the reference's shipped outputs without a GPU.  The GPU suite runs the
same binary over the real library (tests/test_gpu_parity.py::test_els_host_*).
"""
import os
import subprocess

import pytest

from tests._fixtures import GOLDEN_DIR, compare_els_out, els_host_binary, golden, golden_els_out, write_sample_dir

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env, path, **extra):
    e = dict(env, **extra)
    return subprocess.run([els_host_binary(), str(path)], capture_output=True, text=True, timeout=900, env=e)


@pytest.mark.parametrize("name", ["n2", "f2"])
def test_els_host_whole_program_on_the_cpu_double_matches_shipped_els_out(double_env, name, tmp_path):
    """CRCCSD(T)_spatial as shipped: every line of els.out (dates / times masked; numbers within 2 units of the last printed
    digit or 1e-9 Eh), the engine driven in the reference's order, guess_out.dat written as the reference does."""
    text = write_sample_dir(name, str(tmp_path))
    log = tmp_path / "calls.log"
    r = _run(double_env, tmp_path, AFESP_DOUBLE_CALL_LOG=str(log))
    assert r.returncode == 0, r.stderr[-3000:]
    diffs = compare_els_out(r.stdout, golden_els_out(name), ulps=2.0, abs_tol=1e-9)
    assert diffs == [], "\n".join(diffs[:20])
    n_it = len(golden()[name]["ccsd"]) - 1
    calls = log.read_text().split()
    assert calls == ["ao2mo", "mp2_energy", "ccsd_init"] + ["ccsd_iterate", "ccsd_diis"] * (n_it - 1) + \
        ["ccsd_iterate", "ccsd_finalize", "ccsd_t_spatial"]
    wants_guess = "scf_write_guess = .true." in text or "scf_write_guess=.true." in text
    assert (tmp_path / "guess_out.dat").exists() == wants_guess


def test_els_host_spin_orbital_whole_program_matches_the_reference_els_cpu_out(double_env, tmp_path):
    """The reference's cc-pVTZ water directory as shipped (no eri.dat: generated on the fly), calc_type CCSD(T)_spinorb:
    the complete output against the reference's own els_cpu.out (current code version), line by line."""
    from tests.test_gint import _write_tz_dir

    _write_tz_dir(tmp_path, calc_type="CCSD(T)_spinorb")
    r = _run(double_env, tmp_path)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = open(os.path.join(GOLDEN_DIR, "h2o_tz_els_cpu_out.txt")).read()
    diffs = compare_els_out(r.stdout, ref, ulps=2.0, abs_tol=1e-9)
    assert diffs == [], "\n".join(diffs[:20])


def test_els_host_symmetry_assertion_prints_the_reference_block_and_stops(double_env, tmp_path):
    """Status 5 from afesp_gpu_ccsd_init: the two banners, 'Permutational symmetry error:' (E15.6) and the reference's error
    block on stderr with a non-zero stop (src/ccsd.f90:150-167, src/error_handling.f90:6-20)."""
    write_sample_dir("h2o", str(tmp_path), calc_type="CCSD_spinorb")
    r = _run(double_env, tmp_path, AFESP_DOUBLE_SYMMETRY_ERROR="3.5e-7")
    assert r.returncode != 0
    tail = r.stdout.splitlines()[-4:]
    assert tail[0].startswith(" Time taken:") and tail[1] == ""
    assert tail[2] == " Checking that the permuational symmetry of the antisymmetrised integrals hold..."
    assert tail[3] == " Permutational symmetry error:    0.350000E-06"
    assert "ccsd::do_ccsd" in r.stderr and "Permutational symmetry of antisymmetrised integrals does not hold" in r.stderr


@pytest.mark.parametrize("calc,last_call", [("MP2_spatial", "mp2_energy"), ("MP2_spinorb", "mp2_energy"),
                                            ("CCSD_spinorb", "ccsd_finalize"), ("CCSD_spatial", "ccsd_finalize"),
                                            ("RCCSD[T]_spatial", "ccsd_t_spatial"), ("CCSD(T)_spinorb", "ccsd_t_spinorb")])
def test_els_host_and_python_host_print_the_same_program_output(double_env, calc, last_call, tmp_path):
    """Both hosts over the same double on the water sample: identical text (times masked) for calc_types without a shipped
    log, and the engine is left after the right stage."""
    from afesp_b200 import host
    from tests._fixtures import load_els_input
    from tests._oracle_engine import OracleEngine

    write_sample_dir("h2o", str(tmp_path), calc_type=calc)
    log = tmp_path / "calls.log"
    r = _run(double_env, tmp_path, AFESP_DOUBLE_CALL_LOG=str(log))
    assert r.returncode == 0, r.stderr[-3000:]
    assert log.read_text().split()[-1] == last_call
    inp = host.read_inputs(str(tmp_path))
    res = host.run(inp, gpu=OracleEngine())
    diffs = compare_els_out(r.stdout, res.stdout, ulps=1.0, abs_tol=1e-10)
    assert diffs == [], "\n".join(diffs[:20])


def test_els_host_writes_the_fcidump_the_python_writer_writes(double_env, tmp_path):
    """write_fcidump = .true. (src/mp2.f90:451-487): the file els_host writes from the MO integrals the engine returns is the
    file the Python writer produces from the oracle's transform (sign freedom of the eigenvectors aside), and the two
    banner lines are printed around it."""
    import re

    from afesp_b200 import host
    from oracle import afesp_oracle as orc
    from tests._fixtures import load_els_input

    text = write_sample_dir("h2o", str(tmp_path), calc_type="MP2_spatial")
    text = re.sub(r"write_fcidump\s*=\s*\.false\.", "write_fcidump = .true.", text)
    assert "write_fcidump = .true." in text
    (tmp_path / "els.in").write_text(text)
    r = _run(double_env, tmp_path)
    assert r.returncode == 0, r.stderr[-3000:]
    out = r.stdout.splitlines()
    k = out.index(" Writing FCIDUMP file...")
    assert out[k + 1] == " Done writing FCIDUMP file!" and out[k - 1].startswith(" MP2 correlation energy (Hartree):")
    inp = load_els_input("h2o", "MP2_spatial")
    _, C, eps, _, conv = host.rhf(inp)
    ref_path = tmp_path / "FCIDUMP.ref"
    host.write_fcidump(str(ref_path), orc.ao2mo_packed(inp.eri, C), inp.nbasis)
    mine, ref = (tmp_path / "FCIDUMP").read_text().splitlines(), ref_path.read_text().splitlines()
    key = lambda ln: (ln[:12], round(abs(float(ln[12:])), 7))
    big = lambda lines: {key(ln) for ln in lines if abs(float(ln[12:])) > 1e-5}
    assert len(big(ref)) > 1000 and big(mine) == big(ref)
