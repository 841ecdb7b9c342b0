"""Build the committed parity fixtures from the reference's sample_data (run in the build container only).

    python tests/golden/make_fixtures.py [/root/reference]

Writes, next to this file:
  <mol>.npz   inputs of one sample directory (overlap, kinetic, nuclear attraction, packed AO ERIs,
              geometry, optional guess_in Fock matrix, the els.in text) -- data, not reference source;
  golden.json numbers parsed from the reference's own shipped outputs (els.out / ref_out):
              SCF table, MP2, CCSD iteration table, final-energy table.
/root/reference does not exist on the GPU box, so tests read only these files.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import afesp_oracle as orc  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
MOLS = {
    "n2": ("sample_data/n2-cc-pvdz/2.00_0.00", "els.out"),
    "f2": ("sample_data/f2-cc-pvdz/1.75_0.00", "els.out"),
    "h2o": ("sample_data/h2o-cc-pvdz/1.80_104.45", "ref_out"),
}
FLOAT = r"[-+]?\d+\.\d+(?:[EeDd][-+]?\d+)?"


def parse_out(text):
    g = {"scf": [], "ccsd": [], "final": {}}
    mode = None
    for line in text.splitlines():
        if "Restricted Hartree-Fock" in line:
            mode = "scf"
        elif line.strip() == "CCSD":
            mode = "ccsd"
        elif "Final energy breakdown" in line:
            mode = "final"
        m = re.match(rf"^\s+(\d+|MP1)\s+({FLOAT})\s+({FLOAT})\s+({FLOAT})(?:\s+({FLOAT}))?\s*$", line)
        if m and mode in ("scf", "ccsd"):
            it = m.group(1)
            g[mode].append([it if it == "MP1" else int(it), float(m.group(2)), float(m.group(3)), float(m.group(4))])
            continue
        m = re.match(rf"^\s*Iteration\s+(\d+)\s+({FLOAT})\s+{FLOAT} s", line)  # older ref_out format
        if m and mode == "ccsd":
            g["ccsd"].append([int(m.group(1)), float(m.group(2))])
            continue
        m = re.match(rf"^\s*Final CCSD Energy \(Hartree\):\s+({FLOAT})", line)
        if m:
            g["e_ccsd_12"] = float(m.group(1))
        m = re.match(rf"^\s*MP2 correlation energy \(Hartree\):\s+({FLOAT})", line)
        if m:
            g["e_mp2_8"] = float(m.group(1))
        if mode == "final":
            m = re.match(rf"^\s*(.+?):\s+({FLOAT})\s*$", line)
            if m:
                g["final"][m.group(1).strip()] = float(m.group(2))
    return g


def main():
    golden = {}
    for name, (rel, outname) in MOLS.items():
        d = os.path.join(REF, rel)
        sysm = orc.read_system(d)
        s = np.loadtxt(os.path.join(d, "s.dat"), ndmin=2)
        n = sysm.nbasis
        ke = orc._read_sym(os.path.join(d, "t.dat"), n)
        en = orc._read_sym(os.path.join(d, "v.dat"), n)
        with open(os.path.join(d, "geom.dat")) as f:
            toks = f.read().split()
        nat = int(toks[0])
        geom = np.array(toks[1:1 + 4 * nat], dtype=float).reshape(nat, 4)
        guess = np.zeros((0, 0))
        gpath = os.path.join(d, "guess_in.dat")
        if os.path.exists(gpath):
            gd = np.loadtxt(gpath, ndmin=2)
            guess = np.zeros((n, n))
            guess[gd[:, 0].astype(int) - 1, gd[:, 1].astype(int) - 1] = gd[:, 2]
        with open(os.path.join(d, "els.in")) as f:
            els_in = f.read()
        # guess_out.dat as the reference wrote it (src/hf.f90:172-191): values + the first lines verbatim (format pin)
        guess_out, guess_out_head = np.zeros((0, 0)), ""
        gopath = os.path.join(d, "guess_out.dat")
        if os.path.exists(gopath):
            gd = np.loadtxt(gopath, ndmin=2)
            guess_out = np.zeros((n, n))
            guess_out[gd[:, 0].astype(int) - 1, gd[:, 1].astype(int) - 1] = gd[:, 2]
            with open(gopath) as f:
                guess_out_head = "".join(f.readlines()[:5])
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), ovlp=sysm.ovlp, ke=ke, en=en, eri=sysm.eri,
                            geom=geom, guess=guess, els_in=np.array(els_in), guess_out=guess_out,
                            guess_out_head=np.array(guess_out_head))
        with open(os.path.join(d, outname)) as f:
            out_text = f.read()
        golden[name] = parse_out(out_text)
        if outname == "els.out":   # program output of the current code version, kept verbatim for the layout tests
            with open(os.path.join(HERE, f"{name}_els_out.txt"), "w") as f:
                f.write(out_text)
        golden[name]["source"] = f"{rel}/{outname}"
        print(name, n, len(golden[name]["scf"]), len(golden[name]["ccsd"]), sorted(golden[name]["final"])[:3])
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1)


if __name__ == "__main__":
    main()
