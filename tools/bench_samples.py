"""`python bench.py --workload n2|f2|h2o|h2o-spinorb [--impl ours|reference] [--steps K --warmup W]`

Throughput on the sample_data molecules (BASELINE.json configs[0] and [2]): the WHOLE converged calculation through the
host program (afesp_b200.host.run: RHF on the host, AO->MO + MP2 + CCSD with DIIS to convergence + triples on the GPU),
stage seconds as the reference prints them ("Time taken for restricted MP2 / CCSD / ... CCSD(T)") beside
  * the reference's OWN published stage times, parsed from the shipped els.out (tests/golden/*_els_out.txt: OpenBLAS +
    OpenMP on the author's machine) -- the one place a number of the real reference exists for this path, and
  * `--impl reference`: the CPU port of the same stages run IN FULL on this box's host cores (no extrapolation at
    nbf=28): AO->MO as src/mp2.f90:321-410, every CCSD iteration as src/ccsd.f90:1040-1312 + 1538-1732 (DIIS and the
    energy in NumPy, a few per cent of an iteration), the triples loop of :2152-2233 over all o^3 ordered triples.
Inputs come from the committed fixtures (tests/golden/*.npz) -- /root/reference does not exist on the GPU box.
"""
import os
import re
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKLOADS = {   # name -> (fixture, calc_type, BASELINE.json config)
    "n2": ("n2", "CRCCSD(T)_spatial", "N2 cc-pVDZ CRCCSD(T)_spatial from sample_data (configs[2])"),
    "f2": ("f2", "CRCCSD(T)_spatial", "F2 cc-pVDZ CRCCSD(T)_spatial from sample_data (configs[2])"),
    "h2o": ("h2o", "CCSD(T)_spatial", "H2O cc-pVDZ CCSD(T)_spatial from sample_data inputs"),
    "h2o-spinorb": ("h2o", "CCSD(T)_spinorb", "H2O cc-pVDZ CCSD(T)_spinorb from sample_data (configs[0])"),
}
METRIC = "sample_molecule_ccsd_plus_triples_seconds"


def published_times(fixture):
    """Stage seconds the reference itself printed in the shipped els.out (None where the file has no such line)."""
    path = os.path.join(ROOT, "tests", "golden", f"{fixture}_els_out.txt")
    out = {}
    if not os.path.exists(path):
        return None
    for ln in open(path):
        m = re.match(r"^ Time taken for (.+?):\s+([0-9.]+)s", ln)
        if m:
            out[m.group(1)] = float(m.group(2))
    return out


def stage_times(published):
    if not published:
        return None
    g = lambda pat: next((v for k, v in published.items() if re.search(pat, k)), None)
    return {"mp2_s": g(r"MP2$"), "ccsd_s": g(r"CCSD$"), "triples_s": g(r"CCSD[\[(]T[\])]$"), "rhf_s": g(r"Hartree-Fock")}


def run_ours(args, fixture, calc, emit, log):
    import torch

    from afesp_b200 import AfespGpu, host
    from tests._fixtures import golden, load_els_input

    torch.cuda.set_device(0)
    inp = load_els_input(fixture, calc)
    gpu = AfespGpu(0)
    runs = []
    l0, _ = gpu.counters()
    for s in range(args.warmup + args.steps):
        if s == args.warmup:
            l0, _ = gpu.counters()
        t0 = time.perf_counter()
        res = host.run(inp, gpu=gpu)
        wall = time.perf_counter() - t0
        if s >= args.warmup:
            runs.append({"wall_s": wall, **{k: res.timings.get(k) for k in ("rhf_s", "mp2_s", "ccsd_s", "triples_s")},
                         "ao2mo_device_ms": res.timings.get("ao2mo_device_ms"),
                         "ccsd_iter_device_ms_mean": float(np.mean(res.timings.get("ccsd_iter_device_ms", [0.0]))),
                         "triples_device_ms": res.timings.get("triples_device_ms")})
    l1, _ = gpu.counters()
    med = lambda k: float(np.median([r[k] for r in runs if r[k] is not None])) if any(r[k] is not None for r in runs) else None
    value = (med("ccsd_s") or 0.0) + (med("triples_s") or 0.0)
    G = golden().get(fixture, {})
    parity = {"iterations": len(res.ccsd_table) - 1, "e_ccsd": res.e_ccsd}
    if calc.endswith("_spatial") and calc.startswith("CRCCSD") and "e_ccsd" in G:
        parity.update({"golden_e_ccsd": G["e_ccsd"], "abs_diff_e_ccsd": abs(res.e_ccsd - G["e_ccsd"]),
                       "golden_iterations": len(G.get("ccsd", [])), "ok": bool(abs(res.e_ccsd - G["e_ccsd"]) < 1e-9)})
    pub = published_times(fixture) if calc.startswith("CRCCSD") else None
    line = {"metric": METRIC, "value": value, "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "sample_data fixture (tests/golden/%s.npz)" % fixture,
            "config": {"workload": WORKLOADS[args.workload][2], "calc_type": calc, "nbf": int(inp.nbasis),
                       "nocc": int(inp.nel // 2)},
            "stages_s": {k: med(k) for k in ("rhf_s", "mp2_s", "ccsd_s", "triples_s", "wall_s")},
            "device_ms": {k: med(k) for k in ("ao2mo_device_ms", "ccsd_iter_device_ms_mean", "triples_device_ms")},
            "reference_published_s": stage_times(pub),
            "reference_published_note": "stage times the reference printed in the shipped sample_data els.out (its own "
                                        "OpenBLAS/OpenMP build on the author's machine, core count not stated)" if pub else None,
            "energies": {"e_mp2": res.e_mp2, "e_ccsd": res.e_ccsd,
                         **{k: v for k, v in res.energies.items() if isinstance(v, float)}},
            "parity": parity, "gpu_launches": int(l1 - l0),
            "e2e": {"value": med("wall_s"), "unit": "s", "what": "whole program through host.run from host arrays: RHF on the "
                    "host, H2D of the packed AO integrals, AO->MO, MP2, CCSD to convergence, triples",
                    "h2d_bytes_per_step": int(inp.eri.nbytes + inp.nbasis ** 2 * 8), "d2h_bytes_per_step": 0}}
    emit(line)
    gpu.close()


def _respawn_passive(args):
    """The reference alternates OpenMP loop nests and OpenBLAS dgemms; with libgomp's default spinning workers the two
    thread pools fight for the cores (10x slower iterations at nbf=28).  Re-exec once with OMP_WAIT_POLICY=passive and
    every core for both pools (torchrun exports OMP_NUM_THREADS=1)."""
    if os.environ.get("AFESP_REF_CHILD") == "1":
        return False
    cores = str(len(os.sched_getaffinity(0)))
    env = dict(os.environ, AFESP_REF_CHILD="1", OMP_WAIT_POLICY="passive", OMP_NUM_THREADS=cores, OPENBLAS_NUM_THREADS=cores)
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", args.workload, "--impl", "reference",
                        "--steps", str(args.steps), "--warmup", str(args.warmup), "--gpus", str(args.gpus)], env=env,
                       stdout=subprocess.PIPE, text=True)
    sys.stderr.flush()
    return r.stdout


def run_reference(args, fixture, calc, emit, log):
    sys.path.insert(0, ROOT)
    child_out = _respawn_passive(args)
    if child_out is not False:
        import json
        emit(json.loads(child_out.strip().splitlines()[-1]))
        return
    from oracle import afesp_oracle as orc
    from oracle import cpu_port
    from tests._fixtures import load_system

    restricted = calc.endswith("_spatial")
    if not restricted:
        emit({"impl": "reference", "unavailable": "the CPU port covers the spin-free path only (spin-orbital: NumPy oracle, "
                                                  "not a timed port)"})
        return
    lib = cpu_port.load()
    threads = cpu_port.set_threads(lib, len(os.sched_getaffinity(0)))   # OMP_WAIT_POLICY: see _respawn_passive()
    sysm = load_system(fixture, calc)
    orc.do_rhf(sysm)
    n, o = sysm.nbasis, sysm.nel // 2
    comp_renorm = calc.startswith("CRCCSD")
    renorm = calc.startswith("RCCSD") or comp_renorm
    paren = "(T)" in calc
    runs = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        mo, tq = cpu_port.ao2mo(lib, sysm.eri, sysm.coeff)
        t_ao = time.perf_counter() - t0
        t0 = time.perf_counter()
        V = cpu_port.slices(lib, mo, n, o)
        D1, D2 = orc.denominators(sysm.eps, o)
        t1 = np.zeros((o, n - o), order="F")
        t2 = np.asfortranarray(V["v_oovv"] / D2)
        diis = orc.CCDiis(sysm.ccsd_diis_n_errmat, t1.shape, t2.shape)
        energy = orc.restricted_energy(t1, t2, np.asarray(V["v_oovv"]))
        t2_old = t2.copy()
        iters, t_iter_c = 0, 0.0
        for it in range(1, sysm.ccsd_maxiter + 1):
            diis.stash(np.array(t1), np.array(t2))
            t1, t2, _, w = cpu_port.ccsd_iter(lib, V, sysm.eps, t1, t2)
            t_iter_c += w
            e_old, energy = energy, orc.restricted_energy(t1, t2, np.asarray(V["v_oovv"]))
            rms = float(np.sum((t2 - t2_old) ** 2))
            t2_old = t2.copy()
            iters = it
            if np.sqrt(rms) < sysm.ccsd_t_tol and abs(energy - e_old) < sysm.ccsd_e_tol:
                break
            t1, t2 = diis.update(np.array(t1), np.array(t2))
        t_ccsd = time.perf_counter() - t0
        ijk = [(i, j, k) for i in range(o) for j in range(o) for k in range(o)]
        t0 = time.perf_counter()
        if comp_renorm:   # the CR intermediates come from the NumPy oracle (one-off o v^4 work, not part of the loop timed)
            Vn = {k: np.asarray(x) for k, x in V.items()}
            I = orc.restricted_intermediates(np.array(t1), np.array(t2), Vn)
            Ivv, Ioo = orc.cr_intermediates(np.array(t1), np.array(t2), Vn, I["I_vo"], I["asym_t2"])
            t0 = time.perf_counter()
            sums, _ = cpu_port.triples_cr(lib, t1, t2, V["v_oovv"], V["v_vvov"], V["v_oovo"], Ivv, Ioo, sysm.eps, ijk, paren)
        else:
            sums, _ = cpu_port.triples(lib, t1, t2, V["v_oovv"], V["v_vvov"], V["v_oovo"], sysm.eps, ijk, paren, renorm)
        t_T = time.perf_counter() - t0
        if s >= args.warmup:
            runs.append({"mp2_s": t_ao, "ccsd_s": t_ccsd, "triples_s": t_T, "ccsd_iter_c_s": t_iter_c, "iterations": iters,
                         "e_ccsd": energy, "e_T_sum": float(sums[0])})
        log(f"[reference {fixture}] ao2mo {t_ao:.3f}s ccsd {t_ccsd:.3f}s ({iters} it) triples {t_T:.3f}s E_CCSD {energy:.12f}")
    med = lambda k: float(np.median([r[k] for r in runs]))
    value = med("ccsd_s") + med("triples_s")
    emit({"impl": "reference", "metric": METRIC, "value": value, "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
          "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong",
          "vs_baseline": None, "dtype": "f64", "data": "sample_data fixture (tests/golden/%s.npz)" % fixture,
          "config": {"workload": WORKLOADS[args.workload][2], "calc_type": calc, "nbf": int(n), "nocc": int(o)},
          "stages_s": {k: med(k) for k in ("mp2_s", "ccsd_s", "triples_s")},
          "ccsd_iterations": runs[-1]["iterations"], "energies": {"e_ccsd": runs[-1]["e_ccsd"], "e_T_sum": runs[-1]["e_T_sum"]},
          "reference_published_s": stage_times(published_times(fixture)) if comp_renorm else None,
          "cpu_baseline": {"value": value, "unit": "s", "cores": threads, "kind": "port",
                           "sample": "the whole calculation, no extrapolation: AO->MO (4 quarter transforms + repack), every "
                                     "CCSD iteration to convergence, the triples loop over all o^3 ordered triples"},
          "e2e": {"value": value, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


def run(args, rank, world, local, emit, log):
    if rank != 0:
        return
    if args.workload not in WORKLOADS:
        raise SystemExit(f"unknown workload {args.workload}; choose from {sorted(WORKLOADS)}")
    fixture, calc, _ = WORKLOADS[args.workload]
    sys.path.insert(0, ROOT)
    (run_reference if args.impl == "reference" else run_ours)(args, fixture, calc, emit, log)
