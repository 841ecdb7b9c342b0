"""The C++ host program (host/els_host.cpp, stand-in for the Fortran els.x) on the reference's sample inputs:
CPU part here (readers, namelist, RHF, guess_out.dat, error behaviour); the full CRCCSD(T) runs are in
tests/test_gpu_parity.py::test_els_host_*."""
import os
import subprocess

import numpy as np

from tests._fixtures import GOLDEN_DIR, compare_els_out, els_host_binary, golden_els_out, write_sample_dir


def test_els_host_scf_section_is_byte_identical_to_the_shipped_output(tmp_path):
    write_sample_dir("n2", str(tmp_path), calc_type="RHF")
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    mine = r.stdout.splitlines()
    ref = golden_els_out("n2").replace('calc_type="CRCCSD(T)_spatial"', 'calc_type="RHF"').splitlines()
    stop = next(i for i, ln in enumerate(ref) if ln.startswith(" Time taken for restricted Hartree-Fock")) + 1
    assert compare_els_out("\n".join(mine[:stop]), "\n".join(ref[:stop]), ulps=0.0) == []
    assert " RHF energy:                     -108.3305827541" in mine
    # guess_out.dat: same text as the file the reference wrote (src/hf.f90:172-191)
    z = np.load(os.path.join(GOLDEN_DIR, "n2.npz"))
    lines = (tmp_path / "guess_out.dat").read_text().splitlines()
    for mine_ln, ref_ln in zip(lines, str(z["guess_out_head"]).splitlines()):
        assert len(mine_ln) == len(ref_ln) and mine_ln.split()[:2] == ref_ln.split()[:2]
        if abs(float(ref_ln.split()[2])) > 1e-6:      # entries that are numerical noise (1e-12) differ in every run
            assert mine_ln == ref_ln
    got = np.array([float(x.split()[2]) for x in lines]).reshape(z["guess_out"].shape)
    assert np.max(np.abs(got - z["guess_out"])) < 5e-9


def test_els_host_f2_scf_table(tmp_path):
    write_sample_dir("f2", str(tmp_path), calc_type="RHF")
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    ref = golden_els_out("f2").replace('calc_type="CRCCSD(T)_spatial"', 'calc_type="RHF"').splitlines()
    stop = next(i for i, ln in enumerate(ref) if ln.startswith(" Time taken for restricted Hartree-Fock")) + 1
    assert compare_els_out("\n".join(r.stdout.splitlines()[:stop]), "\n".join(ref[:stop]), ulps=1.0) == []


def test_els_host_error_behaviour(tmp_path):
    """src/error_handling.f90:6-20: error block on stderr, non-zero stop."""
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0
    assert " ERROR." in r.stderr and "system::read_system_in" in r.stderr and "els.in does not exist" in r.stderr
    write_sample_dir("n2", str(tmp_path), calc_type="CCSD(Q)_spatial")
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "Unrecognised calculation type!" in r.stderr


def test_binary_packed_integrals_are_read_by_both_hosts(tmp_path):
    """eri.bin (the packed array as little-endian doubles, for basis sets whose text eri.dat
    would be billions of lines) gives the same SCF as eri.dat, in the C++ host and in the Python host."""
    from afesp_b200 import host

    write_sample_dir("f2", str(tmp_path), calc_type="RHF")
    r1 = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=300)
    z = np.load(os.path.join(GOLDEN_DIR, "f2.npz"))
    z["eri"].astype("<f8").tofile(str(tmp_path / "eri.bin"))
    os.remove(tmp_path / "eri.dat")
    r2 = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r1.returncode == 0 and r2.returncode == 0, r2.stderr
    assert [ln for ln in r1.stdout.splitlines() if "energy" in ln] == [ln for ln in r2.stdout.splitlines() if "energy" in ln]
    inp = host.read_inputs(str(tmp_path))
    assert np.array_equal(inp.eri, z["eri"])
    # a truncated file is refused with the reference's error block
    z["eri"][:-3].astype("<f8").tofile(str(tmp_path / "eri.bin"))
    r3 = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r3.returncode != 0 and "eri.bin" in r3.stderr


NAMELIST_CASES = [
    # (els.in text, expected (calc_type, scf_maxiter, scf_e_tol) or None = the reference's 'invalid input file format!')
    ('&elsinput\ncalc_type="RHF",\nscf_e_tol=1e-6,\nscf_maxiter = 77,\n/\n', ("RHF", 77, 1e-6)),
    # several assignments per line, upper case, single quotes, d exponent, comments, text after the closing slash
    ("! header comment\n &ELSINPUT CALC_TYPE = 'RHF' SCF_E_TOL=1d-8, scf_maxiter=5 ! trailing\n write_fcidump=T /\n junk = 1\n",
     ("RHF", 5, 1e-8)),
    ('&elsinput calc_type="RHF" scf_maxiter=3 /', ("RHF", 3, None)),
    ('&elsinput\ncalc_type="RHF"\nscf_read_guess = .FALSE.\nscf_write_guess=.f.\n&end\n', ("RHF", None, None)),
    ('calc_type="RHF"\n/\n', None),                         # no group
    ('&elsinput\ncalc_type="RHF",\nbogus_key=1\n/\n', None),  # not a member of the namelist
    ('&elsinput\ncalc_type="RHF",\nscf_maxiter=many\n/\n', None),
    ('&elsinput\ncalc_type="RHF",\nscf_maxiter=5\n', None),   # group never closed
]


def test_namelist_reader_follows_fortran_rules_in_both_hosts(tmp_path):
    """`read(unit=ir, nml=elsinput)` (src/system.f90:96-107): what a Fortran runtime accepts is accepted by both hosts with
    the same values, what it refuses stops both with 'invalid input file format!'."""
    import re

    from afesp_b200 import host

    write_sample_dir("f2", str(tmp_path), calc_type="RHF")
    for text, want in NAMELIST_CASES:
        (tmp_path / "els.in").write_text(text)
        r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=300)
        if want is None:
            assert r.returncode != 0 and "invalid input file format!" in r.stderr, text
            try:
                host.parse_namelist(text)
                raise AssertionError("the Python host accepted: " + text)
            except ValueError as ex:
                assert "invalid input file format!" in str(ex)
            continue
        assert r.returncode == 0, (text, r.stderr)
        got = host.parse_namelist(text)
        assert got["calc_type"] == want[0]
        if want[1] is not None:
            assert got["scf_maxiter"] == want[1]
            assert re.search(r"Maximum number of SCF iterations: %d\n" % want[1], r.stdout)
        if want[2] is not None:
            assert got["scf_e_tol"] == want[2]
            assert (" scf_e_tol: %8.2E" % want[2]) in r.stdout


def test_calc_type_uhf_runs_the_restricted_scf_as_the_reference_does(tmp_path):
    """src/main.F90:49-56: every unrestricted calc_type starts with do_rhf (do_uhf is a stub, src/hf.f90:193); "UHF" then
    goes straight to the final table.  The system block counts spin-orbitals (src/geometry.f90:40-46).  Both hosts."""
    from afesp_b200 import host

    write_sample_dir("f2", str(tmp_path), calc_type="UHF")
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert " Number of occupied orbitals: 18\n Number of virtual orbitals: 38\n" in r.stdout
    assert " Time taken for restricted Hartree-Fock:" in r.stdout and " MP2" not in r.stdout
    res = host.run(host.read_inputs(str(tmp_path)))
    assert compare_els_out(r.stdout, res.stdout, ulps=1.0) == []
    ref = golden_els_out("f2").splitlines()
    rhf = next(ln for ln in ref if ln.startswith(" RHF energy:"))
    assert rhf in r.stdout.splitlines()
    block = r.stdout.split(" Final energy breakdown\n", 1)[1].splitlines()
    assert [ln.split(":")[0].strip() for ln in block[:5]] == ["RHF energy", "-" * 47, "Total electronic energy",
                                                               "Nuclear repulsion", "Total energy"]
