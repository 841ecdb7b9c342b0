"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle and the reference's golden outputs.

Tolerances (BASELINE.json north_star): energies within 1e-9 Eh, converged amplitudes within 1e-8, identical
iteration counts.  Index permutations are bit-exact; GEMMs are compared at 1e-12 relative to the operand scale.
"""
import os

import numpy as np
import pytest

from oracle import afesp_oracle as orc
from tests._fixtures import golden, load_els_input, load_system

pytestmark = pytest.mark.gpu

E_TOL = 1e-9
G = golden()


@pytest.fixture(scope="module")
def gpu():
    from afesp_b200 import AfespGpu

    g = AfespGpu(0)
    yield g
    g.close()


# ---------------------------------------------------------------- operators of linalg.fpp
@pytest.mark.parametrize("ta", ["N", "T"])
@pytest.mark.parametrize("tb", ["N", "T"])
@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 5, 3), (33, 65, 17), (130, 129, 70), (20, 400, 64), (257, 31, 300),
                                   (441, 49, 441), (64, 64, 4096), (21, 21, 3087), (128, 256, 64)])
def test_dgemm_wrapper_matches_blas_semantics(gpu, ta, tb, shape):
    M, N, K = shape
    rng = np.random.default_rng(M * 1000 + N * 10 + K)
    A = rng.standard_normal((K, M) if ta == "T" else (M, K))
    B = rng.standard_normal((N, K) if tb == "T" else (K, N))
    C0 = rng.standard_normal((M, N))
    opA = A.T if ta == "T" else A
    opB = B.T if tb == "T" else B
    for alpha, beta in [(1.0, 0.0), (0.5, 1.0), (-1.0, 0.25)]:
        want = alpha * (opA @ opB) + beta * C0
        got = gpu.dgemm_wrapper(ta, tb, M, N, K, A.ravel(order="F"), B.ravel(order="F"), C0.ravel(order="F"),
                                alpha, beta).reshape((M, N), order="F")
        scale = np.abs(opA) @ np.abs(opB) + np.abs(C0)
        assert np.max(np.abs(got - want) / scale) < 1e-13


def test_dgemm_beta_zero_ignores_nan_in_c(gpu):
    A = np.ones((4, 3)); B = np.ones((3, 5)); Cn = np.full((4, 5), np.nan)
    got = gpu.dgemm_wrapper("N", "N", 4, 5, 3, A.ravel(order="F"), B.ravel(order="F"), Cn.ravel(order="F"), 1.0, 0.0)
    assert np.array_equal(got, np.full(20, 3.0))


def test_omp_reshape_all_24_orders_bit_exact(gpu):
    import itertools

    rng = np.random.default_rng(7)
    x = rng.standard_normal((5, 7, 19, 3))
    for perm in itertools.permutations("1234"):
        order = "".join(perm)
        axes = [int(c) - 1 for c in order]
        want = np.transpose(x, axes)  # out(perm(i,j,k,l)) = in(i,j,k,l)   (src/linalg.fpp:133-147)
        got = gpu.omp_reshape(x, order)
        assert np.array_equal(got, want), order
        y = rng.standard_normal(want.shape)
        got2 = gpu.omp_reshape(x, order, out_arr=y, beta=-0.5)
        assert np.array_equal(got2, -0.5 * y + want), order


def test_omp_reshape_large_transpose_bit_exact(gpu):
    rng = np.random.default_rng(8)
    x = rng.standard_normal((20, 20, 45, 45))
    for order in ["3412", "2143", "4321", "1342", "3124"]:
        axes = [int(c) - 1 for c in order]
        assert np.array_equal(gpu.omp_reshape(x, order), np.transpose(x, axes))


@pytest.mark.parametrize("shape", [(12, 12, 70, 66), (3, 70, 5, 130), (64, 64, 9, 7), (65, 66, 4, 5), (2, 3, 200, 97)])
def test_omp_reshape_every_kernel_family_bit_exact(gpu, shape):
    """All 24 orders (with and without beta) at shapes that reach each permute kernel: contiguous rows of every
    width class, the slab kernel (leading axes shuffled, <= 4096 elements, e.g. 2143 / 2134), the 32x32 and the 64x64
    tile transposes with ragged edges."""
    import itertools

    rng = np.random.default_rng(11)
    x = rng.standard_normal(shape)
    for perm in itertools.permutations("1234"):
        order = "".join(perm)
        axes = [int(c) - 1 for c in order]
        want = np.transpose(x, axes)
        assert np.array_equal(gpu.omp_reshape(x, order), want), (shape, order)
    for order in ["2143", "3412", "1243", "4123"]:
        axes = [int(c) - 1 for c in order]
        want = np.transpose(x, axes)
        y = rng.standard_normal(want.shape)
        assert np.array_equal(gpu.omp_reshape(x, order, out_arr=y, beta=0.25), 0.25 * y + want), (shape, order)


def test_elementwise_oovv_kernels_match_numpy(gpu):
    """Division-free (o,o,v,v) kernels: the MP1 amplitudes / energy of ccsd_init at a ragged shape (o=7, v=23: chunks of
    32 (a,b) pairs do not divide v^2; o^2 = 49 < block size) against NumPy, through the public init call."""
    from afesp_b200 import synthetic

    n, o = 30, 7
    v = n - o
    eri, Cmo, eps = synthetic.make(n, o, seed=4)
    mo = gpu.ao2mo(n, eri, Cmo)
    e_mp2 = gpu.mp2_energy(o, eps)
    e_mp1, rms = gpu.ccsd_init(o, True, eps, 8)
    V = orc.spatial_slices(mo, n, o)
    D1, D2 = orc.denominators(eps, o)
    t2 = V["v_oovv"] / D2
    want_e = float(np.sum((2.0 * V["v_oovv"] - V["v_oovv"].transpose(0, 1, 3, 2)) * t2))
    assert abs(e_mp1 - want_e) < 1e-12 and abs(e_mp1 - e_mp2) < 1e-12
    assert abs(rms - float(np.sum(t2 * t2))) < 1e-12
    _, t1, t2_dev = gpu.ccsd_finalize(want_amplitudes=True)
    assert np.max(np.abs(t2_dev - t2)) < 1e-14


# ---------------------------------------------------------------- AO->MO + MP2
@pytest.mark.parametrize("name", ["n2", "f2", "h2o"])
def test_ao2mo_and_mp2_match_oracle_and_golden(gpu, name, oracle_runs):
    s, r = oracle_runs(name, "MP2_spatial")
    eri_mo = gpu.ao2mo(s.nbasis, s.eri, s.coeff)
    assert np.max(np.abs(eri_mo - s.eri_mo)) < 1e-11
    e = gpu.mp2_energy(s.nel // 2, s.eps)
    assert abs(e - r["e_mp2"]) < E_TOL
    key = "MP2 correlation energy" if name != "h2o" else "E_MP2_corr"
    assert abs(e - G[name]["final"][key]) < E_TOL + 0.5e-10


# ---------------------------------------------------------------- full program through the product host driver
def _check_table(table, gold_rows, tol=E_TOL):
    assert len(table) == len(gold_rows), "iteration count differs from the reference"
    for (it, e, de, rms), row in zip(table, gold_rows):
        assert str(it) == str(row[0])
        assert abs(e - row[1]) < tol, (it, e, row[1])
        if len(row) > 3:
            assert abs(rms - row[3]) < 1e-9


FINAL_MAP = {
    "CCSD[T] correlation energy": "e_ccsd_t", "CCSD(T) correlation energy": "e_ccsd_tt",
    "R-CCSD[T] correlation energy": "e_rccsd_t", "R-CCSD(T) correlation energy": "e_rccsd_tt",
    "CR-CCSD[T] correlation energy": "e_crccsd_t", "CR-CCSD(T) correlation energy": "e_crccsd_tt",
    "D[T]": "D_T", "D(T)": "D_TT",
}


@pytest.mark.parametrize("name", ["n2", "f2"])
def test_crccsd_t_spatial_matches_shipped_els_out(gpu, name):
    from afesp_b200 import host

    inp = load_els_input(name)  # CRCCSD(T)_spatial as shipped
    res = host.run(inp, gpu=gpu)
    g = G[name]
    assert len(res.scf_table) == len(g["scf"])
    assert abs(res.e_hf + res.e_nuc - g["final"]["RHF energy"]) < E_TOL
    assert abs(res.e_mp2 - g["final"]["MP2 correlation energy"]) < E_TOL
    assert res.ccsd_converged
    _check_table(res.ccsd_table, g["ccsd"])
    assert abs(res.e_ccsd - g["e_ccsd_12"]) < E_TOL
    assert abs(res.t1_diagnostic - g["final"]["T1 diagnostic"]) < E_TOL
    for label, key in FINAL_MAP.items():
        assert abs(res.energies[key] - g["final"][label]) < E_TOL + 0.5e-10, (label, res.energies[key], g["final"][label])
    # the printed block keeps the reference's labels
    for label in ["CR-CCSD(T) correlation energy:", "T1 diagnostic:", "D[T]:", "Total energy:"]:
        assert label in res.stdout


@pytest.mark.parametrize("name", ["n2", "f2"])
def test_whole_program_output_matches_shipped_els_out(gpu, name, tmp_path):
    """els.out of the drop-in (Python host + GPU library) against the file the reference shipped for the same
    inputs: same lines, same layout, every printed number equal to within 2 units of its last printed digit or
    1e-9 Eh (the 12-decimal CCSD table); only dates and wall-clock times are masked.
    Also writes guess_out.dat into the run directory as the reference does (scf_write_guess)."""
    import os

    from afesp_b200 import host
    from tests._fixtures import compare_els_out, golden_els_out

    inp = load_els_input(name)
    res = host.run(inp, gpu=gpu, workdir=str(tmp_path))
    diffs = compare_els_out(res.stdout, golden_els_out(name), ulps=2.0, abs_tol=E_TOL)
    assert diffs == [], "\n".join(diffs[:20])
    assert os.path.exists(tmp_path / "guess_out.dat") == bool(inp.scf_write_guess)


@pytest.mark.parametrize("name", ["n2", "f2"])
def test_els_host_whole_program_matches_shipped_els_out(name, tmp_path):
    """The C++ host program (host/els_host.cpp) run in a directory holding the reference's input files, exactly as
    els.x is run: its stdout against the shipped els.out (dates/times masked, numbers within 2 units of the last
    printed digit), plus the guess_out.dat it leaves behind."""
    import subprocess

    from tests._fixtures import compare_els_out, els_host_binary, golden_els_out, write_sample_dir

    text = write_sample_dir(name, str(tmp_path))
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    diffs = compare_els_out(r.stdout, golden_els_out(name), ulps=2.0, abs_tol=E_TOL)
    assert diffs == [], "\n".join(diffs[:20])
    wants_guess = "scf_write_guess = .true." in text or "scf_write_guess=.true." in text
    assert (tmp_path / "guess_out.dat").exists() == wants_guess


def test_els_host_writes_fcidump_and_spinorb_path(tmp_path):
    """write_fcidump = .true. on H2O cc-pVDZ CCSD(T)_spinorb: the FCIDUMP written from the device-resident MO
    integrals equals the Python writer on the oracle's transform; energies equal the oracle's."""
    import re
    import subprocess

    from afesp_b200 import host
    from tests._fixtures import els_host_binary, write_sample_dir

    text = write_sample_dir("h2o", str(tmp_path), calc_type="CCSD(T)_spinorb")
    text = re.sub(r"write_fcidump\s*=\s*\.false\.", "write_fcidump = .true.", text)
    (tmp_path / "els.in").write_text(text)
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert " Writing FCIDUMP file..." in r.stdout and " Done writing FCIDUMP file!" in r.stdout
    inp = load_els_input("h2o", "CCSD(T)_spinorb")
    _, C, eps, _, conv = host.rhf(inp)
    mo = orc.ao2mo_packed(inp.eri, C)
    ref_path = tmp_path / "FCIDUMP.ref"
    host.write_fcidump(str(ref_path), mo, inp.nbasis)
    mine, ref = (tmp_path / "FCIDUMP").read_text().splitlines(), ref_path.read_text().splitlines()
    # eigenvector signs are free, so individual integrals may flip sign; magnitudes and positions must agree
    key = lambda ln: (ln[:12], round(abs(float(ln[12:])), 7))
    big = lambda lines: {key(ln) for ln in lines if abs(float(ln[12:])) > 1e-5}
    assert big(mine) == big(ref)
    m = re.search(r"Unrestricted CCSD\(T\) correlation energy \(Hartree\):\s+(-?\d+\.\d+)", r.stdout)
    sysm = load_system("h2o", "CCSD(T)_spinorb")
    assert abs(float(m.group(1)) - orc.run(sysm)["e_ccsd_t"]) < 2e-9


@pytest.mark.parametrize("calc", ["CCSD(T)_spatial", "CCSD[T]_spatial", "RCCSD(T)_spatial", "RCCSD[T]_spatial",
                                  "CRCCSD[T]_spatial"])
def test_spatial_calc_types_match_oracle(gpu, calc, oracle_runs):
    from afesp_b200 import host

    s, r = oracle_runs("h2o", calc)
    res = host.run(load_els_input("h2o", calc), gpu=gpu)
    assert len(res.ccsd_table) == len(r["ccsd"])
    assert abs(res.e_ccsd - r["e_ccsd"]) < E_TOL
    for key in ["e_ccsd_t", "e_ccsd_tt", "e_rccsd_t", "e_rccsd_tt", "e_crccsd_t", "e_crccsd_tt", "D_T", "D_TT"]:
        assert (key in r) == (key in res.energies), key
        if key in r:
            assert abs(res.energies[key] - r[key]) < E_TOL, (key, res.energies[key], r[key])
    if calc == "CCSD(T)_spatial":  # Q2: E(T) == E[T] as coded
        assert abs(res.energies["e_ccsd_tt"] - res.energies["e_ccsd_t"]) < 1e-13


def test_converged_amplitudes_match_oracle(gpu, oracle_runs):
    s, r = oracle_runs("h2o", "CCSD_spatial")
    gpu.ao2mo(s.nbasis, s.eri, s.coeff, want_result=False)
    from afesp_b200 import host

    table, conv, e, _ = host.ccsd_loop(gpu, s.nel // 2, True, s.eps, s.ccsd_e_tol, s.ccsd_t_tol,
                                       s.ccsd_diis_n_errmat, s.ccsd_maxiter)
    _, t1, t2 = gpu.ccsd_finalize(want_amplitudes=True)
    cc = r["cc"]
    assert conv and len(table) == len(r["ccsd"])
    assert np.sqrt(np.mean((t1 - cc["t1"]) ** 2)) < 1e-8 and np.max(np.abs(t1 - cc["t1"])) < 1e-8
    assert np.sqrt(np.mean((t2 - cc["t2"]) ** 2)) < 1e-8 and np.max(np.abs(t2 - cc["t2"])) < 1e-8


# ---------------------------------------------------------------- spin-orbital path
def test_spinorbital_ccsd_matches_old_ref_out_with_q1_off(gpu):
    from afesp_b200 import host

    gpu.set_option("q1_transposed_foo", 0)
    try:
        res = host.run(load_els_input("h2o", "CCSD_spinorb"), gpu=gpu)
    finally:
        gpu.set_option("q1_transposed_foo", 1)
    rows = res.ccsd_table[1:]
    assert len(rows) == len(G["h2o"]["ccsd"]) == 19
    for (it, e, _, _), (git, ge) in zip(rows, G["h2o"]["ccsd"]):
        assert it == git and abs(e - ge) < E_TOL


def test_spinorbital_ccsd_t_as_coded_matches_oracle(gpu, oracle_runs):
    from afesp_b200 import host

    s, r = oracle_runs("h2o", "CCSD(T)_spinorb", q1=True)
    res = host.run(load_els_input("h2o", "CCSD(T)_spinorb"), gpu=gpu)
    assert len(res.ccsd_table) == len(r["ccsd"])
    for (it, e, _, rms), (oit, oe, _, orms) in zip(res.ccsd_table, r["ccsd"]):
        assert abs(e - oe) < E_TOL and abs(rms - orms) < 1e-9
    assert abs(res.e_ccsd - (-0.311554581875)) < E_TOL
    assert abs(res.energies["e_ccsd_t"] - r["e_ccsd_t"]) < E_TOL


def test_spinorbital_symmetry_assertion_runs_and_aborts(gpu, tmp_path):
    """src/ccsd.f90:150-167: the permutational-symmetry self-check of <pq||rs> is executed on the device (error and its
    device time come back through afesp_gpu_ccsd_init_info, the host prints them); above the threshold the calculation
    stops with the reference's message.  The packed 8-fold storage cannot hold a symmetry-breaking integral, so the abort
    path is driven through the threshold (depsilon, src/const.F90:19) instead of a corrupted input."""
    import subprocess

    from afesp_b200 import host
    from afesp_b200.capi import AfespError
    from tests._fixtures import els_host_binary, write_sample_dir

    inp = load_els_input("h2o", "CCSD_spinorb")
    res = host.run(inp, gpu=gpu)
    info = gpu.ccsd_init_info()
    assert info["symmetry_error"] == 0.0 and info["check_s"] > 0.0 and info["slices_s"] > 0.0
    lines = res.stdout.splitlines()
    k = lines.index(" Checking that the permuational symmetry of the antisymmetrised integrals hold...")
    assert abs(float(lines[k + 1].split()[2]) - info["check_s"]) < 1e-6 and float(lines[k + 1].split()[2]) > 0.0
    # oracle: the same four identities over the full index range
    sysm = load_system("h2o", "CCSD_spinorb")
    orc.do_rhf(sysm)
    asym = orc.spinorb_antisym(orc.ao2mo_packed(sysm.eri, sysm.coeff), sysm.nbasis)
    assert orc.spinorb_symmetry_error(asym) == 0.0
    gpu.set_option("spinorb_symmetry_tol", -1.0)
    try:
        with pytest.raises(AfespError) as ei:
            host.run(inp, gpu=gpu)
    finally:
        gpu.set_option("spinorb_symmetry_tol", 1e-12)
    assert ei.value.code == 5 and "Permutational symmetry of antisymmetrised integrals does not hold" in str(ei.value)
    assert " Permutational symmetry error:" in ei.value.stdout and "Initialisation done" not in ei.value.stdout
    # the C++ host: error block of src/error_handling.f90 naming ccsd::do_ccsd, non-zero stop, nothing after the check
    write_sample_dir("h2o", str(tmp_path), calc_type="CCSD_spinorb")
    env = dict(os.environ, AFESP_GPU_OPTIONS="spinorb_symmetry_tol=-1")
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode != 0 and "ccsd::do_ccsd" in r.stderr and "does not hold" in r.stderr
    assert " Permutational symmetry error:" in r.stdout and "Initialisation done" not in r.stdout
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "Initialisation done" in r.stdout


def test_h2o_cc_pvtz_spinorbital_ccsd_t_matches_reference_els_cpu_out(gpu):
    """BASELINE.json configs[1] input (sample_data/h2o-cc-pvtz, 58 basis functions, 116 spin-orbitals) with the two-electron
    integrals regenerated by the host-side generator (the checkout ships no eri.dat): the whole program reproduces the
    reference's own els_cpu.out -- 22 SCF iterations, MP2, the MP1 line and all 19 spin-orbital CCSD iterations to 1e-9 Eh
    and E[CCSD(T)] = -0.4340327558.  This is the current code version, so it pins Q1 (transposed F_oo term) as coded and
    the spin-orbital (T) against reference output, not only against the oracle."""
    from afesp_b200 import host

    Gz = G["h2o_tz"]
    res = host.run(load_els_input("h2o_tz", "CCSD(T)_spinorb"), gpu=gpu)
    assert len(res.scf_table) == len(Gz["scf"]) == 22
    for (it, e, _, rms), (git, ge, _, grms) in zip(res.scf_table, Gz["scf"]):
        assert it == git and abs(e - ge) < 2e-9
    assert abs(res.e_mp2 - Gz["e_mp2_8"]) < 1e-8
    rows = res.ccsd_table
    assert len(rows) == len(Gz["ccsd"]) == 20 and rows[0][0] == "MP1"
    for (it, e, _, rms), (git, ge, _, grms) in zip(rows, Gz["ccsd"]):
        assert it == git and abs(e - ge) < E_TOL and abs(rms - grms) < 1e-9, (it, e, ge)
    assert abs(res.e_ccsd - Gz["e_ccsd_12"]) < E_TOL
    assert abs(res.energies["e_ccsd_t"] - Gz["final"]["CCSD(T) correlation energy"]) < E_TOL
    assert abs(res.energies["e_ccsd_t"] - Gz["e_ccsd_t_9"]) < E_TOL


# ---------------------------------------------------------------- (T) sharding and symmetry properties
def test_triples_partition_sums_and_symmetry_switch(gpu, oracle_runs):
    from afesp_b200 import host

    s, r = oracle_runs("f2")  # CRCCSD(T)_spatial
    gpu.ao2mo(s.nbasis, s.eri, s.coeff, want_result=False)
    host.ccsd_loop(gpu, s.nel // 2, True, s.eps, s.ccsd_e_tol, s.ccsd_t_tol, s.ccsd_diis_n_errmat, s.ccsd_maxiter)
    gpu.ccsd_finalize(want_cr=True)
    full, const = gpu.ccsd_t_spatial(True, False, True)
    want = np.array(r["triples_sums"])
    assert np.max(np.abs(full - want)) < E_TOL
    # the work dealt to 3 "ranks" adds up to the whole (what the NCCL allreduce does across GPUs)
    parts = []
    for rank in range(3):
        gpu.set_partition(rank, 3)
        parts.append(gpu.ccsd_t_spatial(True, False, True)[0])
    gpu.set_partition(0, 1)
    assert np.max(np.abs(np.sum(parts, axis=0) - full)) < 1e-11
    # all o^3 ordered triples (the reference's loop) give the same sums as the i<=j<=k orbit form
    gpu.set_option("triples_ijk_symmetry", 0)
    try:
        allo3, _ = gpu.ccsd_t_spatial(True, False, True)
    finally:
        gpu.set_option("triples_ijk_symmetry", 1)
    assert np.max(np.abs(allo3 - full)) < 1e-10
    # tiny work buffer -> many batches, same answer
    gpu.set_option("triples_batch_bytes", 1 << 20)
    try:
        small, _ = gpu.ccsd_t_spatial(True, False, True)
    finally:
        gpu.set_option("triples_batch_bytes", 6 << 30)
    assert np.max(np.abs(small - full)) < 1e-11


# ---------------------------------------------------------------- synthetic workload path (device-generated integrals)
def test_synthetic_chain_matches_oracle_and_device_generated_integrals(gpu):
    from afesp_b200 import host, synthetic

    n, o = 26, 4
    eri, Cm, eps = synthetic.make(n, o, seed=11)
    B, C2, eps2 = synthetic.make_factors(n, o, seed=11)
    mo_ref = orc.ao2mo_packed(eri, Cm)
    gpu.synth_eri_ao(n, B, C2)           # integrals expanded on the device from the low-rank factors
    mo = gpu.ao2mo(n)                    # transform of the resident copy
    assert np.max(np.abs(mo - mo_ref)) < 1e-12
    assert np.max(np.abs(gpu.get_eri_mo() - mo)) == 0.0
    assert abs(gpu.mp2_energy(o, eps) - orc.mp2_energy(mo_ref, eps, o)) < E_TOL
    cc = orc.ccsd_spatial(mo_ref, eps, o, 1e-8, 1e-9, 8, 50, want_cr=True)
    table, conv, e, _ = host.ccsd_loop(gpu, o, True, eps, 1e-8, 1e-9, 8, 50)
    assert conv and len(table) == len(cc["table"]) and abs(e - cc["e_ccsd"]) < E_TOL
    gpu.release("eri_ao")
    gpu.ccsd_finalize(want_cr=True)
    sums, const = gpu.ccsd_t_spatial(True, False, True)
    en, osums = orc.triples_spatial(cc, eps, True, False, True)
    assert np.max(np.abs(sums - np.array(osums))) < E_TOL
    got = host.assemble_triples(e, sums, const, True, False, True)
    for k in ["e_ccsd_t", "e_ccsd_tt", "e_rccsd_t", "e_rccsd_tt", "e_crccsd_t", "e_crccsd_tt", "D_T", "D_TT"]:
        assert abs(got[k] - en[k]) < E_TOL, k


@pytest.mark.parametrize("n,o", [(80, 6), (83, 6)])
def test_crccsd_t_chain_at_a_multi_tile_shape_matches_the_cpu(gpu, n, o):
    """The complete CRCCSD(T)_spatial chain -- CCSD to convergence with DIIS, CR intermediates, all six triples sums with the
    M3 / y / z3 branches of the epilogue -- on synthetic integrals with more than one GEMM m-tile per block (v = 74: 64 + a
    ragged 10 on the TMA-staged kernel; v = 77: odd leading dimensions, cp.async kernels), against the CPU: the oracle's
    CCSD and CR intermediates, and the C port of the reference's own triples loop incl. the CR part (oracle/cpu_kernels.c:
    afesp_ref_triples_cr, src/ccsd.f90:2152-2233) over all o^3 ordered triples.  The sample molecules (v <= 53) never leave
    a single tile."""
    from afesp_b200 import host, synthetic
    from oracle import cpu_port

    eri, Cm, eps = synthetic.make(n, o)
    gpu.ao2mo(n, eri, Cm, want_result=False)
    gpu.release("eri_ao")
    table, conv, e, _ = host.ccsd_loop(gpu, o, True, eps, 1e-8, 1e-9, 8, 50)
    t1_diag, _, _ = gpu.ccsd_finalize(want_cr=True)
    sums, const = gpu.ccsd_t_spatial(True, False, True)
    # CPU
    mo, C2, eps2 = cpu_port.synthetic_mo_integrals(n, o)
    assert np.array_equal(eps, eps2)
    cc = orc.ccsd_spatial(mo, eps, o, 1e-8, 1e-9, 8, 50, want_cr=True)
    assert conv and len(table) == len(cc["table"])
    for (it, ee, _, rms), (oit, oe, _, orms) in zip(table, cc["table"]):
        assert it == oit and abs(ee - oe) < E_TOL and abs(rms - orms) < 1e-9
    assert abs(t1_diag - cc["t1_diag"]) < 1e-9
    lib = cpu_port.load()
    cpu_port.set_threads(lib)
    V = cc["V"]
    ijk = [(i, j, k) for i in range(o) for j in range(o) for k in range(o)]
    want, _ = cpu_port.triples_cr(lib, cc["t1"], cc["t2"], V["v_oovv"], V["v_vvov"], V["v_oovo"], cc["I_vovv_pp"],
                                  cc["I_ooov_pp"], eps, ijk, True)
    assert np.max(np.abs(sums - want)) < E_TOL, (sums, want)
    got = host.assemble_triples(e, sums, const, True, False, True)
    ref = orc.assemble_triples(cc["e_ccsd"], tuple(want), orc.triples_denominator_constant(cc["t1"], cc["t2"]), True, False, True)
    for k in ["e_ccsd_t", "e_ccsd_tt", "e_rccsd_t", "e_rccsd_tt", "e_crccsd_t", "e_crccsd_tt", "D_T", "D_TT"]:
        assert abs(got[k] - ref[k]) < E_TOL, k


def test_single_triple_shares_on_mp1_amplitudes_match_cpu_values_from_the_factored_integrals(gpu):
    """The per-triple check bench.py runs at the target shape (bench.mp1_triples_check), here at nbf=64 / nocc=6 where all 56
    unique triples can be gone through: with afesp_gpu_set_partition(r, 56) the handle owns exactly triple number r of the
    (i <= j <= k) list, and on the MP1 amplitudes (state after ccsd_init) its [T] sum equals the orbit value NumPy gets from
    the factored form of the integrals -- for every orbit kind (6, 3 and 1 distinct orderings)."""
    import sys

    from afesp_b200 import synthetic

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_bench_pins as M

    n, o = 64, 6
    uniq = M.unique_triples(o)
    assert len(uniq) == 56
    want, eps = M.mp1_triples_from_factors(n, o, uniq)
    eri, Cm, eps1 = synthetic.make(n, o)
    assert np.array_equal(eps, eps1)
    gpu.ao2mo(n, eri, Cm, want_result=False)
    gpu.release("eri_ao")
    gpu.ccsd_init(o, True, eps, 8)
    gpu.ccsd_finalize()
    try:
        whole, _ = gpu.ccsd_t_spatial(True, False, False)
        got = []
        for r in range(len(uniq)):
            gpu.set_partition(r, len(uniq))
            got.append(gpu.ccsd_t_spatial(True, False, False)[0][0])
    finally:
        gpu.set_partition(0, 1)
    for (i, j, k), g, w in zip(uniq, got, want):
        assert abs(g - w) <= max(1e-9 * abs(w), 1e-15), ((i, j, k), g, w)
    assert abs(sum(got) - whole[0]) < 1e-13 and abs(whole[0] - sum(want)) < 1e-12


@pytest.mark.parametrize("n,o", [(120, 12), (101, 10)])
def test_one_bench_step_at_a_multi_tile_shape_matches_the_cpu(gpu, n, o):
    """One step of bench.py (CCSD iteration, DIIS, (T) on the extrapolated amplitudes) on synthetic integrals large enough for
    several GEMM tiles per block -- v = 108: one full and one ragged m-tile, K = 120: a short K tail, i.e. the EDGE / KTAIL
    variants of the TMA-staged kernel; v = 91: odd leading dimensions, the cp.async kernels -- against a CPU computation of
    the same step: MP2 from the factored integrals, E_CCSD from the CPU port of the reference's iteration
    (oracle/cpu_ccsd.c), the (T) sum over all triples in orbit form (oracle/cpu_port.py: orbit_T_fast, equal to the literal loop).  The same recipe at nbf=200 is what
    tests/golden/bench_pinned.json pins every bench line on."""
    from afesp_b200 import synthetic
    from oracle import cpu_port

    eri, Cm, eps = synthetic.make(n, o)
    gpu.ao2mo(n, eri, Cm, want_result=False)
    gpu.release("eri_ao")
    e_mp2 = gpu.mp2_energy(o, eps)
    e_mp1, _ = gpu.ccsd_init(o, True, eps, 8)
    e1, rms1 = gpu.ccsd_iterate()
    gpu.ccsd_diis()
    gpu.ccsd_finalize()
    sums, _ = gpu.ccsd_t_spatial(True, False, False)
    # CPU: same system through the factored form (no O(n^5) transform), one iteration of the reference's code path
    lib = cpu_port.load()
    cpu_port.set_threads(lib)
    mo, C2, eps2 = cpu_port.synthetic_mo_integrals(n, o)
    assert np.array_equal(eps, eps2) and np.array_equal(Cm, C2)
    V = cpu_port.slices(lib, mo, n, o)
    D1, D2 = orc.denominators(eps, o)
    voovv = np.asarray(V["v_oovv"])
    t2 = np.asfortranarray(voovv / D2)
    want_mp2 = float(np.sum(voovv * (2.0 * voovv - voovv.transpose(0, 1, 3, 2)) / D2))
    assert abs(e_mp2 - want_mp2) < E_TOL and abs(e_mp1 - want_mp2) < E_TOL
    t1n, t2n, _, _ = cpu_port.ccsd_iter(lib, V, eps, np.zeros((o, n - o), order="F"), t2)
    assert abs(e1 - orc.restricted_energy(np.asarray(t1n), np.asarray(t2n), voovv)) < E_TOL
    assert abs(rms1 - float(np.sum((np.asarray(t2n) - np.asarray(t2)) ** 2))) < 1e-9
    vvov, nv = np.ascontiguousarray(V["v_vvov"]), n - o
    want_T = cpu_port.orbit_T_fast(lib, np.ascontiguousarray(t2n),
                                   lambda k: np.ascontiguousarray(vvov[:, :, k, :].transpose(2, 1, 0)).reshape(nv, nv * nv),
                                   np.ascontiguousarray(V["v_oovo"]), eps,
                                   [(i, j, k) for i in range(o) for j in range(i, o) for k in range(j, o)])
    assert abs(sums[0] - want_T) < E_TOL and abs(sums[1] - want_T) < E_TOL      # plain (T): quirk Q2, e_TT = e_T


# ---------------------------------------------------------------- handle / state guards
def test_handle_and_state_guards(gpu):
    """One handle per device and process; DIIS is refused on a finalised state and with an absurd history depth (status
    codes, not crashes)."""
    from afesp_b200 import AfespGpu, synthetic
    from afesp_b200.capi import AfespError

    with pytest.raises(AfespError) as ei:
        AfespGpu(0)
    assert "already has an open handle" in str(ei.value)
    n, o = 12, 2
    eri, Cm, eps = synthetic.make(n, o, seed=2)
    gpu.ao2mo(n, eri, Cm, want_result=False)
    with pytest.raises(AfespError) as ei:
        gpu.ccsd_init(o, True, eps, 65)
    assert "ccsd_diis_n_errmat must lie in 0..64" in str(ei.value)
    gpu.ccsd_init(o, True, eps, 8)
    gpu.ccsd_iterate()
    gpu.ccsd_diis()
    gpu.ccsd_finalize()
    with pytest.raises(AfespError) as ei:
        gpu.ccsd_diis()
    assert ei.value.code == 1 and "finalised" in str(ei.value)
    with pytest.raises(AfespError):
        gpu.ccsd_iterate()
    # finalize_keep_ccsd (benchmark loops) keeps iterating possible -- unless the CR intermediates consumed the work arrays
    gpu.set_option("finalize_keep_ccsd", 1)
    try:
        gpu.ccsd_init(o, True, eps, 8)
        gpu.ccsd_iterate(); gpu.ccsd_diis(); gpu.ccsd_finalize()
        gpu.ccsd_iterate(); gpu.ccsd_diis()
        gpu.ccsd_finalize(want_cr=True)
        with pytest.raises(AfespError) as ei:
            gpu.ccsd_iterate()
        assert "finalised" in str(ei.value)
    finally:
        gpu.set_option("finalize_keep_ccsd", 0)


@pytest.mark.parametrize("calc", ["CCSD(T)_spatial", "RCCSD(T)_spatial"])
def test_h2o_cc_pvtz_spin_free_matches_oracle(gpu, calc):
    """BASELINE.json configs[1]: H2O cc-pVTZ CCSD(T)_spatial (58 basis functions, o 5 / v 53; integrals regenerated, no
    reference log exists for the spin-free calc_type on this molecule): iteration table, E_CCSD and the triples energies
    against the oracle at 1e-9 Eh; RCCSD(T)_spatial adds the true (T) (quirk Q2) and the renormalised pair."""
    from afesp_b200 import host

    sysm = load_system("h2o_tz", calc)
    ref = orc.run(sysm)
    res = host.run(load_els_input("h2o_tz", calc), gpu=gpu)
    assert abs(res.e_mp2 - ref["e_mp2"]) < E_TOL and abs(res.e_mp2 - G["h2o_tz"]["e_mp2_8"]) < 1e-8
    assert len(res.ccsd_table) == len(ref["ccsd"])
    for (it, e, _, rms), (oit, oe, _, orms) in zip(res.ccsd_table, ref["ccsd"]):
        assert it == oit and abs(e - oe) < E_TOL and abs(rms - orms) < 1e-9
    keys = ["e_ccsd_t", "e_ccsd_tt"] + (["e_rccsd_t", "e_rccsd_tt", "D_T"] if calc.startswith("R") else [])
    for k in keys:
        assert abs(res.energies[k] - ref[k]) < E_TOL, k
    if not calc.startswith("R"):
        assert abs(res.energies["e_ccsd_t"] - res.energies["e_ccsd_tt"]) < 1e-12   # Q2: plain CCSD(T)_spatial prints E(T) = E[T]


def test_diis_history_deeper_than_eight_matches_oracle(gpu):
    """ccsd_diis_n_errmat is free in the reference (src/ccsd.f90:577-615 allocates n_errmat ring slots).  Depth 11 on N2:
    the B-matrix row takes two passes of the 8-vector reduction kernel and the extrapolation accumulates the terms beyond
    the eighth -- same table as the oracle at that depth (20 iterations; depth 8 takes the shipped 22)."""
    from afesp_b200 import host

    sysm = load_system("n2", "CCSD_spatial")
    sysm.ccsd_diis_n_errmat = 11
    r = orc.run(sysm)
    inp = load_els_input("n2", "CCSD_spatial")
    inp.ccsd_diis_n_errmat = 11
    res = host.run(inp, gpu=gpu)
    assert len(res.ccsd_table) == len(r["ccsd"]) == 21 and len(G["n2"]["ccsd"]) == 23
    for (it, e, _, rms), (oit, oe, _, orms) in zip(res.ccsd_table, r["ccsd"]):
        assert it == oit and abs(e - oe) < E_TOL and abs(rms - orms) < 1e-9
    assert abs(res.e_ccsd - r["e_ccsd"]) < E_TOL


# ---------------------------------------------------------------- whole-program spin-orbital output (kept last in this file)
def test_whole_program_h2o_cc_pvtz_output_matches_els_cpu_out(gpu, tmp_path):
    """Both hosts on the reference's cc-pVTZ water directory as shipped (no eri.dat; integrals generated on the fly) with
    the shipped calc_type CCSD(T)_spinorb: the complete program output against the reference's own els_cpu.out line by
    line (dates / times masked, numbers within 2 units of the last printed digit or 1e-9 Eh).  The CPU suite runs the same
    comparison for the Python host over the oracle-backed test double (tests/test_host_program_flow.py)."""
    import subprocess

    from afesp_b200 import host
    from tests._fixtures import GOLDEN_DIR, compare_els_out, els_host_binary
    from tests.test_gint import _write_tz_dir

    ref = open(os.path.join(GOLDEN_DIR, "h2o_tz_els_cpu_out.txt")).read()
    res = host.run(load_els_input("h2o_tz", "CCSD(T)_spinorb"), gpu=gpu)
    diffs = compare_els_out(res.stdout, ref, ulps=2.0, abs_tol=E_TOL)
    assert diffs == [], "\n".join(diffs[:20])
    _write_tz_dir(tmp_path, calc_type="CCSD(T)_spinorb")
    r = subprocess.run([els_host_binary(), str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    diffs = compare_els_out(r.stdout, ref, ulps=2.0, abs_tol=E_TOL)
    assert diffs == [], "\n".join(diffs[:20])
