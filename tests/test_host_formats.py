"""On-disk formats either side of the hot path, host logic only (no GPU):
guess_out.dat (src/hf.f90:172-191) against the file the reference itself wrote for N2, FCIDUMP
(src/mp2.f90:451-487) against the format and ordering the reference's loop produces."""
import os
import re

import numpy as np

from afesp_b200 import host
from oracle import afesp_oracle as orc
from tests._fixtures import GOLDEN_DIR, load_els_input


def test_guess_out_matches_the_file_the_reference_wrote(tmp_path):
    z = np.load(os.path.join(GOLDEN_DIR, "n2.npz"))
    inp = load_els_input("n2")
    e, C, eps, table, conv = host.rhf(inp)
    assert conv and len(table) == 12                      # sample_data/n2-cc-pvdz/2.00_0.00/els.out:57-68
    path = tmp_path / "guess_out.dat"
    host.write_scf_guess(str(path), inp.fock_final)
    lines = path.read_text().splitlines()
    n = inp.nbasis
    assert len(lines) == n * n
    # format (I0, 1X, I0, 1X, ES16.9): pinned on the reference's own first lines
    ref_head = str(z["guess_out_head"]).splitlines()
    for mine, ref in zip(lines, ref_head):
        assert re.fullmatch(r"\d+ \d+ [ -]\d\.\d{9}E[+-]\d\d", mine), mine
        assert mine.split()[:2] == ref.split()[:2]
        assert len(mine) == len(ref)
    # values: the SCF is iteration-exact, so the converged Fock matrix agrees to the printed precision
    back = host.read_scf_guess(str(path), n)
    assert np.max(np.abs(back - z["guess_out"])) < 5e-8
    # and it round-trips as the next run's guess_in
    assert np.max(np.abs(back - inp.fock_final)) < 1e-8


def test_fcidump_format_order_and_threshold(tmp_path):
    inp = load_els_input("h2o")
    _, C, eps, _, conv = host.rhf(inp)
    assert conv
    n = inp.nbasis
    mo = orc.ao2mo_packed(inp.eri, C)
    path = tmp_path / "FCIDUMP"
    nwritten = host.write_fcidump(str(path), mo, n)
    lines = path.read_text().splitlines()
    assert len(lines) == nwritten == int(np.sum(np.abs(mo) > np.float32(1e-7)))
    prev = -1
    for ln in lines[:2000] + lines[-2000:]:
        assert re.fullmatch(r"( {0,2}\d{1,3}){4}[ -]{1,2}\d\.\d{9}E[+-]\d\d", ln), ln
        assert len(ln) == 12 + 17
    for ln in lines:
        p, q, r, s = (int(ln[0:3]), int(ln[3:6]), int(ln[6:9]), int(ln[9:12]))
        val = float(ln[12:])
        assert p >= q and p >= r and s <= (q if r == p else r)
        pq, rs = p * (p - 1) // 2 + q, r * (r - 1) // 2 + s
        idx = pq * (pq - 1) // 2 + rs - 1                      # eri_ind, 1-based (src/integrals.f90:196-210)
        assert idx > prev                                      # canonical order, each integral once
        prev = idx
        assert abs(val - mo[idx]) <= 5e-10 * max(1.0, abs(mo[idx]))
    # every loop index of the reference is enumerated exactly once, in packed order
    P, Q, R, S = host.fcidump_indices(n)
    pq = P * (P - 1) // 2 + Q
    rs = R * (R - 1) // 2 + S
    assert np.array_equal(pq * (pq - 1) // 2 + rs - 1, np.arange(mo.size))


def test_program_output_up_to_the_hot_path_matches_the_shipped_els_out():
    """Banner, read-in log, system information, els.in echo, the whole SCF section (12 iteration rows, orbital
    energies) of sample_data/n2-cc-pvdz/2.00_0.00/els.out, byte for byte apart from dates and times."""
    from tests._fixtures import compare_els_out, golden_els_out

    inp = load_els_input("n2", calc_type="RHF")
    inp.els_in_text = inp.els_in_text  # echoed verbatim (calc_type line included as shipped)
    mine = host.run(inp).stdout.splitlines()
    ref = golden_els_out("n2").splitlines()
    stop = next(i for i, ln in enumerate(ref) if ln.startswith(" Time taken for restricted Hartree-Fock")) + 1
    assert compare_els_out("\n".join(mine[:stop]), "\n".join(ref[:stop]), ulps=0.0) == []


def test_python_host_f2_scf_section_matches_the_shipped_output():
    """Same check as for N2 on the second molecule with a shipped els.out (no guess file, 11 SCF iterations)."""
    from tests._fixtures import compare_els_out, golden_els_out

    inp = load_els_input("f2", calc_type="RHF")
    mine = host.run(inp).stdout.splitlines()
    ref = golden_els_out("f2").replace('calc_type="CRCCSD(T)_spatial"', 'calc_type="RHF"').splitlines()
    stop = next(i for i, ln in enumerate(ref) if ln.startswith(" Time taken for restricted Hartree-Fock")) + 1
    # the echoed els.in still names the shipped calc_type: compare from the first line after the echo
    assert compare_els_out("\n".join(mine[:stop]), "\n".join(ref[:stop]).replace('calc_type="RHF"', 'calc_type="CRCCSD(T)_spatial"'),
                           ulps=1.0) == []
