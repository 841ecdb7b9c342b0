"""tools/els_wrapper.py -- the Psi4-free counterpart of the reference's scan driver utils/els_wrapper.py -- on the CPU.

From nothing but a molecule name, a basis-set name and a bond length it has to rebuild the sample directories the reference
ships (geometry in Psi4's frame, the four integral files), run the host program there the way the reference's wrapper runs
els.x (guess chaining between scan points), and scrape the same twelve numbers into els_energy.dat / binding_data_els.dat.
"""
import os

import numpy as np
import pytest

from tests._fixtures import GOLDEN_DIR, els_host_binary, golden
from tools import els_wrapper as W


@pytest.mark.parametrize("fixture,mol,bl,ang", [("h2o_tz", "h2o", 2.00, 104.45), ("h2o", "h2o", 1.80, 104.45),
                                                ("f2", "f2", 1.75, 0.0), ("n2", "n2", 2.00, 0.0)])
def test_geometry_generator_reproduces_the_shipped_geom_dat(fixture, mol, bl, ang):
    """The directory names of sample_data (`<bl>_<ang>`) are the wrapper's inputs: Psi4's frame and unit constant give the
    shipped coordinates to the last printed digit (%17.15f)."""
    z = np.load(os.path.join(GOLDEN_DIR, f"{fixture}.npz"))
    Z, xyz = W.generate_molecule(mol, bl, ang)
    assert np.array_equal(Z, z["geom"][:, 0])
    assert np.max(np.abs(xyz - z["geom"][:, 1:])) < 2e-15


def test_wrapper_rebuilds_the_shipped_f2_directory_from_the_bond_length(tmp_path):
    """geom.dat, s.dat, t.dat, v.dat and eri.dat of sample_data/f2-cc-pvdz/1.75_0.00 from ("f2", "cc-pvdz", 1.75): file
    formats of generate_dat_psi (utils/els_wrapper.py:38-69), every integral within 1e-13 of the shipped one."""
    from afesp_b200 import host

    d = tmp_path / "f2-cc-pvdz" / "1.75_0.00"
    Z, xyz = W.generate_molecule("f2", 1.75, 0.0)
    assert W.generate_dat(str(d), Z, xyz, "cc-pvdz") == 28
    assert (d / "geom.dat").read_text().splitlines()[1] == "9\t0.000000000000000\t0.000000000000000\t-1.653510359775600"
    (d / "els.in").write_text(W.els_in_text("RHF", read_guess=False))
    inp = host.read_inputs(str(d))
    z = np.load(os.path.join(GOLDEN_DIR, "f2.npz"))
    assert np.max(np.abs(inp.ovlp - z["ovlp"])) < 1e-13
    assert np.max(np.abs(inp.core_hamil - (z["ke"] + z["en"]))) < 2e-13
    assert np.max(np.abs(inp.eri - z["eri"])) < 1e-13
    first = (d / "eri.dat").read_text().splitlines()[0].split("\t")
    assert first[:4] == ["1", "1", "1", "1"] and len(first[4].split(".")[1]) == 15


def test_rhf_scan_chains_the_scf_guess_and_writes_the_reference_files(tmp_path):
    """Two scan points at the RHF level with els_host (no GPU involved): the second point starts from the first point's
    guess_out.dat, both leave els.out + els_energy.dat (twelve `label: value` lines), the scan leaves binding_data_els.dat;
    the first point is the shipped F2 geometry, so its HF total is the one in the shipped els_energy.dat."""
    b = W.main("f2", "cc-pvdz", 1.77, 1.75, 0.02, 0.0, [els_host_binary()], True, "RHF", outdir=str(tmp_path), log=lambda *_: None)
    assert b.shape == (2, 14) and list(np.round(b[:, 0], 2)) == [1.75, 1.77]
    d0, d1 = tmp_path / "f2-cc-pvdz" / "1.75_0.00", tmp_path / "f2-cc-pvdz" / "1.77_0.00"
    assert "scf_read_guess = .false." in (d0 / "els.in").read_text() and "scf_read_guess = .true." in (d1 / "els.in").read_text()
    assert (d1 / "guess_in.dat").read_text() == (d0 / "guess_out.dat").read_text()
    assert " Reading previous AO Fock matrix as guess..." in (d1 / "els.out").read_text()
    assert " Reading previous AO Fock matrix as guess..." not in (d0 / "els.out").read_text()
    lines = (d0 / "els_energy.dat").read_text().splitlines()
    assert [ln.split(":")[0] for ln in lines] == W.ENERGY_LABELS
    assert abs(float(lines[0].split(":")[1]) - golden()["f2"]["final"]["RHF energy"]) < 1e-9
    assert all(float(ln.split(":")[1]) == 0.0 for ln in lines[1:])            # nothing beyond RHF was asked for
    assert b[1, 2] > b[0, 2]                                                   # stretching F2 beyond 1.75 A raises E(RHF)
    rows = (tmp_path / "f2-cc-pvdz" / "binding_data_els.dat").read_text().splitlines()
    assert len(rows) == 2 and len(rows[0].split()) == 14 and rows[0].split()[0] == "1.750"


def test_full_scan_point_over_the_cpu_double_gives_the_shipped_els_energy_dat(double_env, tmp_path):
    """("f2", "cc-pvdz", 1.75, CRCCSD(T)_spatial) end to end -- geometry, integrals, host program (over the oracle-backed
    test double of the C ABI), scraping: the twelve numbers of the els_energy.dat the reference shipped for this point."""
    b = W.main("f2", "cc-pvdz", 1.75, 1.75, 0.02, 0.0, [els_host_binary()], False, "CRCCSD(T)_spatial", outdir=str(tmp_path),
               env=double_env, log=lambda *_: None)
    fin = golden()["f2"]["final"]
    want = [fin["RHF energy"], fin["MP2 energy"], fin["CCSD energy"], fin["CCSD[T] energy"], fin["CCSD(T) energy"],
            fin["R-CCSD[T] energy"], fin["R-CCSD(T) energy"], fin["CR-CCSD[T] energy"], fin["CR-CCSD(T) energy"],
            fin["T1 diagnostic"], fin["D[T]"], fin["D(T)"]]
    assert b.shape == (1, 14) and np.max(np.abs(b[0, 2:] - np.array(want))) < 2e-9
    text = (tmp_path / "f2-cc-pvdz" / "1.75_0.00" / "els_energy.dat").read_text()
    assert text.startswith("HF: -198.61595458") and "CR-CCSD(T): -199.08125368" in text
