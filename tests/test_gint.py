"""The Gaussian-integral generator (host/gint.c through afesp_b200/gint.py) against the integral
files the reference ships (produced by Psi4, utils/psi4_integrals_nosym.py of the reference tree).

  * sample_data/h2o-cc-pvtz: s.dat / t.dat / v.dat exist, eri.dat does not (.MISSING_LARGE_BLOBS) -- all three
    one-electron matrices are reproduced to 1e-13 (58 functions, s to f shells);
  * sample_data/h2o-cc-pvdz (whose files were in fact generated with def2-SVP: d exponent 1.2, H p exponent 0.8): s, t, v
    AND all 45150 packed two-electron integrals are reproduced to 1e-13;
  * sample_data/n2-cc-pvdz and f2-cc-pvdz (true cc-pVDZ): s, t, v and all 82621 packed two-electron integrals to 2e-13;
  * the regenerated cc-pVTZ integrals, through the oracle's RHF, give the 22-iteration SCF table, the orbital energies and
    the RHF energy of the reference's own els_cpu.out.
"""
import json
import os

import numpy as np

from afesp_b200 import gint
from oracle import afesp_oracle as orc
from tests._fixtures import GOLDEN_DIR, golden

DEF2_SVP = {  # Weigend & Ahlrichs 2005, as in Psi4's def2-svp.gbs
    1: [(0, [(13.0107010, 0.19682158e-01), (1.9622572, 0.13796524), (0.44453796, 0.47831935)]), (0, [(0.12194962, 1.0)]),
        (1, [(0.8, 1.0)])],
    8: [(0, [(2266.1767785, -0.53431809926e-02), (340.87010191, -0.39890039230e-01), (77.363135167, -0.17853911985),
             (21.479644940, -0.46427684959), (6.6589433124, -0.44309745172)]),
        (0, [(0.80975975668, 1.0)]), (0, [(0.25530772234, 1.0)]),
        (1, [(17.721504317, 0.43394573193e-01), (3.8635505440, 0.23094120765), (1.0480920883, 0.51375311064)]),
        (1, [(0.27641544411, 1.0)]), (2, [(1.2, 1.0)])],
}


def test_one_electron_integrals_match_shipped_cc_pvtz_files():
    z = np.load(os.path.join(GOLDEN_DIR, "h2o_tz.npz"))
    r = gint.compute(z["geom"][:, 0], z["geom"][:, 1:], "cc-pvtz", want_eri=False)
    assert r["nbf"] == 58
    assert np.max(np.abs(r["s"] - z["ovlp"])) < 1e-13
    assert np.max(np.abs(r["t"] - z["ke"])) < 1e-13
    assert np.max(np.abs(r["v"] - z["en"])) < 1e-13


def test_all_integrals_match_shipped_h2o_dz_sample():
    z = np.load(os.path.join(GOLDEN_DIR, "h2o.npz"))
    Z = z["geom"][:, 0]
    r = gint.compute(Z, z["geom"][:, 1:], None, shells=[DEF2_SVP[int(q)] for q in Z])
    assert r["nbf"] == 24
    for key, ref in (("s", z["ovlp"]), ("t", z["ke"]), ("v", z["en"]), ("eri", z["eri"])):
        assert np.max(np.abs(r[key] - ref)) < 1e-13, key
    # 8-fold packed order and the text writer: eri.dat lines are (i j k l value) in canonical order
    assert r["eri"].shape == z["eri"].shape


def test_all_integrals_match_shipped_n2_and_f2_samples():
    """sample_data/n2-cc-pvdz and f2-cc-pvdz (28 functions each, true cc-pVDZ: d exponents 0.817 / 1.640): s, t, v and all
    82621 packed two-electron integrals from the built-in nitrogen and fluorine tables."""
    for name in ("n2", "f2"):
        z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
        r = gint.compute(z["geom"][:, 0], z["geom"][:, 1:], "cc-pvdz")
        assert r["nbf"] == 28 and r["eri"].shape == z["eri"].shape == (82621,)
        for key, ref in (("s", z["ovlp"]), ("t", z["ke"]), ("v", z["en"]), ("eri", z["eri"])):
            assert np.max(np.abs(r[key] - ref)) < 2e-13, (name, key)


def test_host_runs_the_f2_directory_without_eri_dat(tmp_path):
    """The C++ host on the reference's F2 directory with eri.dat removed and the basis named: the SCF section of the
    reference's els.out, number by number (generated integrals instead of Psi4's)."""
    import subprocess

    from tests._fixtures import compare_els_out, els_host_binary, golden_els_out, write_sample_dir

    write_sample_dir("f2", str(tmp_path), calc_type="RHF")
    os.remove(tmp_path / "eri.dat")
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, AFESP_BASIS="cc-pVDZ"))
    assert r.returncode == 0, r.stderr[-2000:]
    ref = golden_els_out("f2").replace('calc_type="CRCCSD(T)_spatial"', 'calc_type="RHF"').splitlines()
    stop = next(i for i, ln in enumerate(ref) if ln.startswith(" Time taken for restricted Hartree-Fock")) + 1
    assert compare_els_out("\n".join(r.stdout.splitlines()[:stop]), "\n".join(ref[:stop]), ulps=1.0) == []


def test_regenerated_cc_pvtz_integrals_reproduce_reference_scf(tmp_path):
    z = np.load(os.path.join(GOLDEN_DIR, "h2o_tz.npz"))
    G = golden()["h2o_tz"]
    r = gint.compute(z["geom"][:, 0], z["geom"][:, 1:], "cc-pvtz")
    # through the reference's file formats: write s/t/v/eri.dat + geom.dat + els.in, read them back with the oracle reader
    with open(tmp_path / "geom.dat", "w") as f:
        f.write("%d\n" % len(z["geom"]))
        for row in z["geom"]:
            f.write("%d\t%.15f\t%.15f\t%.15f\n" % (int(row[0]), row[1], row[2], row[3]))
    (tmp_path / "els.in").write_text(str(z["els_in"]))
    gint.write_dat_files(str(tmp_path), r, threshold=1e-12)   # the reference's script drops |(ij|kl)| <= 1e-12
    sysm = orc.read_system(str(tmp_path))
    assert sysm.nbasis == 58 and np.max(np.abs(sysm.eri - r["eri"])) < 2e-12
    orc.do_rhf(sysm)
    table = sysm.log["scf"]
    assert sysm.log["scf_converged"] and len(table) == len(G["scf"]) == 22
    for (it, e, de, rms), (git, ge, gde, grms) in zip(table, G["scf"]):
        assert it == git and abs(e - ge) < 2e-9 and abs(rms - grms) < 2e-9
    assert np.max(np.abs(np.asarray(sysm.eps) - np.array(G["orbital_energies"]))) < 2e-8
    assert abs(sysm.e_hf + sysm.e_nuc - G["final"]["RHF energy"]) < 2e-9   # table energies are electronic; the final table adds E_nuc


def _write_tz_dir(path, calc_type="RHF", with_basis_file=True):
    """The reference checkout's sample_data/h2o-cc-pvtz directory as shipped: els.in, geom.dat, s/t/v.dat -- and NO eri.dat."""
    import re

    z = np.load(os.path.join(GOLDEN_DIR, "h2o_tz.npz"))
    text = re.sub(r'calc_type\s*=\s*"[^"]*"', f'calc_type="{calc_type}"', str(z["els_in"]))
    (path / "els.in").write_text(text)
    with open(path / "geom.dat", "w") as f:
        f.write("%d\n" % len(z["geom"]))
        for row in z["geom"]:
            f.write("%d\t%.15f\t%.15f\t%.15f\n" % (int(row[0]), row[1], row[2], row[3]))
    n = z["ovlp"].shape[0]
    for name, key in (("s.dat", "ovlp"), ("t.dat", "ke"), ("v.dat", "en")):
        with open(path / name, "w") as f:
            for i in range(n):
                for j in range(i + 1):
                    f.write("%d\t%d\t%.15f\n" % (i + 1, j + 1, z[key][i, j]))
    if with_basis_file:
        (path / "basis.dat").write_text("cc-pvtz\n")


def test_hosts_run_the_cc_pvtz_directory_without_eri_dat(tmp_path):
    """Both host programs take the run directory as the reference checkout ships it (no eri.dat), generate the two-electron
    integrals from geom.dat once the basis set is named (basis.dat / AFESP_BASIS), and print the reference's 22 SCF
    iterations of els_cpu.out.  RHF level only here (no GPU needed); the GPU suite runs the whole CCSD(T)."""
    import subprocess

    from afesp_b200 import host
    from tests._fixtures import els_host_binary

    G = golden()["h2o_tz"]
    _write_tz_dir(tmp_path)
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = [ln.split() for ln in r.stdout.splitlines() if len(ln.split()) == 5 and ln.split()[0].isdigit()]
    assert len(rows) == len(G["scf"]) == 22
    for row, (git, ge, _, grms) in zip(rows, G["scf"]):
        assert int(row[0]) == git and abs(float(row[1]) - ge) < 2e-9 and abs(float(row[3]) - grms) < 2e-9
    # the whole SCF section (header, 22-line table, 58 orbital energies) is the reference's text, number by number
    from tests._fixtures import compare_els_out

    ref = open(os.path.join(GOLDEN_DIR, "h2o_tz_els_cpu_out.txt")).read().replace('calc_type="CCSD(T)_spinorb"',
                                                                                   'calc_type="RHF"').splitlines()
    stop = next(i for i, ln in enumerate(ref) if ln.startswith(" Time taken for restricted Hartree-Fock")) + 1
    diffs = compare_els_out("\n".join(r.stdout.splitlines()[:stop]), "\n".join(ref[:stop]), ulps=1.0)
    # (the reference printed spin-orbital counts for its CCSD(T)_spinorb run; an RHF run prints the spatial ones)
    assert [d for d in diffs if "occupied orbitals" not in d and "virtual orbitals" not in d] == []
    # with the shipped calc_type the system block counts spin-orbitals (src/geometry.f90:40-46): the header up to the first
    # time line is then the reference's text without exception (the run itself needs a GPU from MP2 on)
    sub = tmp_path / "spinorb"
    sub.mkdir()
    _write_tz_dir(sub, calc_type="CCSD(T)_spinorb")
    r2 = subprocess.run([els_host_binary(), str(sub)], capture_output=True, text=True, timeout=600)
    ref2 = open(os.path.join(GOLDEN_DIR, "h2o_tz_els_cpu_out.txt")).read().splitlines()
    stop2 = next(i for i, ln in enumerate(ref2) if ln.startswith(" Time taken for system initialisation")) + 1
    assert compare_els_out("\n".join(r2.stdout.splitlines()[:stop2]), "\n".join(ref2[:stop2]), ulps=0.0) == []
    assert " Number of occupied orbitals: 10" in r2.stdout and " Number of virtual orbitals: 106" in r2.stdout
    # Python host: same directory
    inp = host.read_inputs(str(tmp_path))
    assert inp.nbasis == 58 and inp.eri is not None
    e_hf, _, eps, table, conv = host.rhf(inp)
    assert conv and len(table) == 22 and abs(table[-1][1] - G["scf"][-1][1]) < 2e-9
    # without a named basis the reference's own error stays
    os.remove(tmp_path / "basis.dat")
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "cannot open eri.dat" in r.stderr
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, AFESP_BASIS="cc-pVTZ"))
    assert r.returncode == 0
    # a wrong basis is caught by the overlap check against s.dat
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=60,
                       env=dict(os.environ, AFESP_BASIS="cc-pvdz"))
    assert r.returncode != 0 and "does not match s.dat" in r.stderr
