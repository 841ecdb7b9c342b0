#!/usr/bin/env python
"""Bond-length scan driver around the host program -- the counterpart of the reference's utils/els_wrapper.py without Psi4.

The reference's wrapper builds a molecule with Psi4, dumps its integrals into `<mol>-<basis>/<bl>_<ang>/` (geom.dat, s.dat,
t.dat, v.dat, eri.dat: utils/els_wrapper.py:38-69), runs `els.x` there with the previous point's guess_out.dat as
guess_in.dat (:88-97), scrapes twelve numbers out of the final block of els.out (:99-128) and writes els_energy.dat per
point and binding_data_els.dat per scan (:184-205).  Here the integrals come from afesp_b200/gint.py (Psi4's conventions
and Psi4's molecular frame: centre of mass at the origin, C2 axis = z, molecule in the yz plane, bohr with
1 bohr = 0.52917721067 Angstrom -- the shipped geom.dat files are reproduced to 1e-15) and the program run is
host/els_host (or any els.x-compatible binary).  The Psi4 reference energies (reference.dat, binding_data_psi4.dat) are not
produced: there is no Psi4 in this image.

    python tools/els_wrapper.py --mol f2 --basis cc-pvdz --bl-lower 1.75 --bl-upper 1.79 --bl-step 0.02 \\
        --calc-type "CRCCSD(T)_spatial" --read-in --outdir /tmp/scan
"""
import argparse
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BOHR2ANG = 0.52917721067            # Psi4's constant (CODATA 2014); fixes the shipped geom.dat files to the last digit
MASS = {1: 1.00782503223, 7: 14.00307400443, 8: 15.99491461957, 9: 18.99840316273}   # most abundant isotopes, as Psi4
DIATOMICS = {"n2": 7, "f2": 9, "h2": 1}
ENERGY_KEYS = ["RHF energy:", "MP2 energy:", " CCSD energy:", " CCSD[T] energy:", " CCSD(T) energy:", " R-CCSD[T] energy:",
               " R-CCSD(T) energy:", " CR-CCSD[T] energy:", " CR-CCSD(T) energy:", " T1 diagnostic:", " D[T]:", " D(T):"]
ENERGY_LABELS = ["HF", "MP2", "CCSD", "CCSD[T]", "CCSD(T)", "R-CCSD[T]", "R-CCSD(T)", "CR-CCSD[T]", "CR-CCSD(T)",
                 "T1 diagnostic", "D[T]", "D(T)"]


def generate_water(bl, ang):
    """utils/els_wrapper.py:71-79 (`O / H 1 bl / H 1 bl 2 ang`, Angstrom / degrees) in Psi4's frame.  Returns (Z, xyz[bohr])."""
    half = np.radians(ang) / 2.0
    y, dz = bl * np.sin(half) / BOHR2ANG, bl * np.cos(half) / BOHR2ANG
    z_o = -(2.0 * MASS[1] * dz) / (MASS[8] + 2.0 * MASS[1])
    return np.array([8.0, 1.0, 1.0]), np.array([[0.0, 0.0, z_o], [0.0, -y, z_o + dz], [0.0, y, z_o + dz]])


def generate_diatomic(z, bl):
    """Homonuclear diatomic along z, centre of mass at the origin (the frame of the shipped N2 / F2 directories)."""
    h = 0.5 * bl / BOHR2ANG
    return np.array([float(z), float(z)]), np.array([[0.0, 0.0, -h], [0.0, 0.0, h]])


def generate_molecule(mol, bl, ang):
    mol = mol.lower()
    if mol in ("h2o", "water"):
        return generate_water(bl, ang)
    if mol in DIATOMICS:
        return generate_diatomic(DIATOMICS[mol], bl)
    raise ValueError(f"unknown molecule {mol!r} (h2o, n2, f2, h2)")


def generate_dat(dirname, Z, xyz, basis, eri_threshold=1e-12):
    """geom.dat + the four integral files in the formats of generate_dat_psi (utils/els_wrapper.py:38-69): atom count, then
    `Z<TAB>x<TAB>y<TAB>z` with %17.15f in bohr; integrals `i<TAB>j<TAB>value` / `i j k l value`, |eri| > 1e-12 only (:34)."""
    from afesp_b200 import gint

    os.makedirs(dirname, exist_ok=True)
    with open(os.path.join(dirname, "geom.dat"), "w") as f:
        f.write("%d\n" % len(Z))
        for z, r in zip(Z, xyz):
            f.write("%1d\t%17.15f\t%17.15f\t%17.15f\n" % (int(z), r[0], r[1], r[2]))
    res = gint.compute(Z, xyz, basis)
    gint.write_dat_files(dirname, res, threshold=eri_threshold)
    return res["nbf"]


def els_in_text(calc_type, read_guess, write_guess=True, scf_maxiter=150, ccsd_maxiter=200, write_fcidump=False):
    """utils/els.in / utils/els_noread.in of the reference with the calc_type and the guess switches filled in."""
    b = lambda x: ".true." if x else ".false."
    return ("&elsinput\n"
            f'calc_type="{calc_type}",\nscf_e_tol=1e-6,\nscf_d_tol=1e-7,\nscf_diis_n_errmat=6,\nccsd_e_tol=1e-6,\n'
            f"ccsd_t_tol=1e-7,\nccsd_diis_n_errmat=8,\nscf_maxiter = {scf_maxiter},\nccsd_maxiter = {ccsd_maxiter},\n"
            f"write_fcidump = {b(write_fcidump)},\nscf_read_guess = {b(read_guess)},\nscf_write_guess = {b(write_guess)}\n/\n")


def parse_energies(lines):
    """The scraping rule of run_els (utils/els_wrapper.py:99-128): substring match, last blank-separated token."""
    energy = np.zeros(12)
    for line in lines:
        for k, key in enumerate(ENERGY_KEYS):
            if key in line:
                energy[k] = float(line.split(" ")[-1])
    return energy


def run_els(els_cmd, directory, previous_directory, read_in, calc_type, env=None):
    """run_els (utils/els_wrapper.py:88-128): els.in, guess chaining, run in the directory, keep els.out, scrape."""
    with open(os.path.join(directory, "els.in"), "w") as f:
        f.write(els_in_text(calc_type, read_guess=bool(read_in)))
    if read_in:
        shutil.copy(os.path.join(previous_directory, "guess_out.dat"), os.path.join(directory, "guess_in.dat"))
    r = subprocess.run(els_cmd, cwd=directory, capture_output=True, text=True, env=env)
    with open(os.path.join(directory, "els.out"), "w") as f:
        f.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError(f"els failure in {directory}: {r.stderr[-600:]}")
    return parse_energies(r.stdout.split("\n"))


def write_els_energy(directory, e):
    with open(os.path.join(directory, "els_energy.dat"), "w") as f:   # utils/els_wrapper.py:190-203
        for label, val in zip(ENERGY_LABELS, e):
            f.write(f"{label}: {val}\n")


def main(molname, basis, bl_upper, bl_lower, bl_step, ang, els_cmd, read_in, calc_type, outdir=".", env=None, log=print):
    """main (utils/els_wrapper.py:130-208) minus the Psi4 reference run.  Returns binding_data_els (num_points x 14)."""
    top = os.path.join(outdir, f"{molname}-{basis}")
    os.makedirs(top, exist_ok=True)
    num_points = int(round((bl_upper - bl_lower) / bl_step + 1))
    binding = np.zeros((num_points, 14))
    prev = ""
    for i, bl in enumerate(np.linspace(bl_lower, bl_upper, num_points)):
        dirname = os.path.join(top, f"{bl:.2f}_{ang:.2f}")
        log(f"Doing calculations in {dirname}")
        Z, xyz = generate_molecule(molname, bl, ang)
        generate_dat(dirname, Z, xyz, basis)
        try:
            e = run_els(els_cmd, dirname, prev, read_in=(read_in and i > 0), calc_type=calc_type, env=env)
        except RuntimeError as ex:
            log(f"els failure: {ex}")
            binding = binding[:i]
            break
        prev = dirname
        binding[i, :2] = [bl, ang]
        binding[i, 2:] = e
        write_els_energy(dirname, e)
    np.savetxt(os.path.join(top, "binding_data_els.dat"), binding, ["%5.3f", "%6.3f"] + ["%17.15f"] * 12)
    return binding


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--mol", default="h2o")
    ap.add_argument("--basis", default="cc-pvdz")
    ap.add_argument("--bl-lower", type=float, default=2.00)
    ap.add_argument("--bl-upper", type=float, default=2.20)
    ap.add_argument("--bl-step", type=float, default=0.02)
    ap.add_argument("--ang", type=float, default=104.45, help="degrees (0 for diatomics, as the shipped directory names)")
    ap.add_argument("--calc-type", default="CRCCSD(T)_spatial")
    ap.add_argument("--read-in", action="store_true", help="chain guess_out.dat -> guess_in.dat between points")
    ap.add_argument("--els", default=os.path.join(ROOT, "host", "els_host"), help="els.x-compatible binary")
    ap.add_argument("--outdir", default=".")
    a = ap.parse_args()
    main(a.mol, a.basis, a.bl_upper, a.bl_lower, a.bl_step, a.ang, [a.els], a.read_in, a.calc_type, a.outdir)
