#!/usr/bin/env python
"""Benchmark of the coupled-cluster hot path (BASELINE.json metric: CCSD s/iter & (T) wall-s; % FP64 tensor peak; vs CPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nbf 200 --nocc 20]

One "step" = one pass of the hot path over the synthetic (nbf, nocc) system: one spin-free CCSD iteration
(intermediates + amplitude equations + energy + DIIS extrapolation) followed by the full (T) correction
(calc_type CCSD(T)_spatial).  `value` = seconds per step, device-resident inputs, max over ranks; the breakdown
(ccsd_s_per_iter, t_wall_s, ao2mo_s) is carried in extra keys.  Under torchrun every rank runs the (replicated) CCSD
iteration and its round-robin share of the (i<=j<=k) triples; the six (T) sums are combined by one NCCL allreduce inside
the library ("scaling": "strong": total work fixed).

`--impl reference` times the CPU port of the reference's own loops (oracle/cpu_kernels.c + OpenBLAS dgemm) on the host
cores on a bounded sample of the same workload (the Fortran reference cannot be compiled here: no Fortran compiler).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ccsd_iter_plus_T_seconds"
UNIT = "s"


# Keep stdout clean for the single JSON line: libraries loaded later (NCCL prints its version banner on stdout) write
# to fd 1, so fd 1 is pointed at stderr for the whole run and the JSON goes to the saved original descriptor.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_sample(nbf, nocc, budget_s):
    """Bounded sample of the reference CPU path at shape (nbf, nocc): ladder dgemm (full), ring loop nest on a slice
    of b, reference (T) loop on a handful of ordered triples; each extrapolated to the full loop and stated."""
    from oracle import cpu_port

    lib = cpu_port.load()
    threads = int(lib.afesp_ref_threads())
    o, v = nocc, nbf - nocc
    rng = np.random.default_rng(1)
    t2 = np.asfortranarray(rng.standard_normal((o, o, v, v)) * 1e-3)
    Iov = np.asfortranarray(rng.standard_normal((o, v, o, v)) * 1e-2)
    Ivo = np.asfortranarray(rng.standard_normal((v, o, o, v)) * 1e-2)
    # ladder dgemm through OpenBLAS (src/ccsd.f90:1669): c(o^2 x v^2) . v_vvvv(v^2 x v^2).  The dense v^4 operand is
    # 8.4 GB at nbf=200 and 134 GB at nbf=400, so a column block of at most 4 GB is multiplied and the time scaled.
    ncol = int(max(1, min(v * v, (4 << 30) // (8 * v * v))))
    vv = np.full((v * v, ncol), 1e-3, order="F")
    cm = np.asfortranarray(t2.reshape((o * o, v * v), order="F"))
    t0 = time.perf_counter()
    _ = cm @ vv
    t_ladder = (time.perf_counter() - t0) * (v * v) / ncol
    ladder_note = "in full" if ncol == v * v else f"on {ncol} of {v * v} columns, x{v * v / ncol:.1f}"
    del vv
    # ring loop nest (src/ccsd.f90:1680-1695): calibrate on 1 slice of b, then spend ~budget/3
    _, dt1 = cpu_port.ring(lib, t2, Iov, t2, Ivo, bmax=1)
    bmax = int(max(1, min(v, (budget_s / 3.0) / max(dt1, 1e-6))))
    _, dt = cpu_port.ring(lib, t2, Iov, t2, Ivo, bmax=bmax)
    t_ring = dt * v / bmax
    # (T): reference loop (src/ccsd.f90:2152-2233) on ntri ordered triples, one per thread
    t1 = np.asfortranarray(rng.standard_normal((o, v)) * 1e-2)
    voovv = t2
    vvvov = np.asfortranarray(np.full((v, v, o, v), 1e-3))
    voovo = np.asfortranarray(rng.standard_normal((o, o, v, o)) * 1e-2)
    eps = np.concatenate([np.linspace(-2, -0.5, o), np.linspace(0.5, 3, v)])
    ntri = threads
    ijk = [(int(rng.integers(o)), int(rng.integers(o)), int(rng.integers(o))) for _ in range(ntri)]
    # calibrate on one slab of the outer virtual loop, then spend ~budget/3 (all v slabs when they fit)
    _, dt1 = cpu_port.triples(lib, t1, t2, voovv, vvvov, voovo, eps, ijk, True, False, amax=1)
    amax = int(max(1, min(v, (budget_s / 3.0) / max(dt1, 1e-6))))
    _, dt_t = cpu_port.triples(lib, t1, t2, voovv, vvvov, voovo, eps, ijk, True, False, amax=amax)
    t_T = dt_t * (v / amax) * (o ** 3) / ntri
    sample = (f"ladder dgemm o^2 x v^2 x v^2 {ladder_note} ({t_ladder:.2f}s) + ring loop nest :1680-1695 on b<{bmax} of {v} "
              f"({dt:.2f}s, x{v / bmax:.1f}) [the other per-iteration terms of the reference are smaller dgemms and are "
              f"not timed: lower bound] + reference (T) loop on {ntri} of {o ** 3} ordered triples, outer virtual "
              f"index a<{amax} of {v} ({dt_t:.2f}s, x{(v / amax) * o ** 3 / ntri:.0f})")
    return {"ccsd_s_per_iter": t_ladder + t_ring, "t_wall_s": t_T, "value": t_ladder + t_ring + t_T,
            "cores": threads, "sample": sample, "kind": "port"}


def run_reference(args, rank):
    if rank != 0:
        return
    budget = max(20.0, 150.0 / max(1, args.steps + args.warmup))
    vals = []
    last = None
    for s in range(args.warmup + args.steps):
        last = cpu_sample(args.nbf, args.nocc, budget)
        if s >= args.warmup:
            vals.append(last["value"])
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": v * 1e3, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic integrals nbf={args.nbf} nocc={args.nocc} CCSD(T)_spatial", "nbf": args.nbf,
                   "nocc": args.nocc, "calc_type": "CCSD(T)_spatial"},
        "ccsd_s_per_iter": last["ccsd_s_per_iter"], "t_wall_s": last["t_wall_s"],
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": "port", "sample": last["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local):
    import torch
    import torch.distributed as dist

    from afesp_b200 import AfespGpu, synthetic

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, o = args.nbf, args.nocc
    v = n - o
    big = n > 240   # host-side expansion of the packed ERIs needs npair^2 doubles: expand on the device instead
    t0 = time.perf_counter()
    if big:
        Bfac, Cmo, eps = synthetic.make_factors(n, o)
        eri = None
    else:
        eri, Cmo, eps = synthetic.make(n, o)
    log(f"[rank {rank}] synthetic inputs nbf={n} nocc={o} generated in {time.perf_counter() - t0:.1f}s")
    gpu = AfespGpu(local)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(AfespGpu.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        gpu.comm_init(rank, world, bytes(uid.cpu().tolist()))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.tma >= 0:
        gpu.set_option("gemm_use_tma", args.tma)
    tma_scope, tma_selftest = gpu.tma_status()
    peak = max(gpu.dmma_peak(), gpu.dmma_peak())
    # ---- device-resident leg
    npair = n * (n + 1) // 2
    npk = npair * (npair + 1) // 2
    # host copy of the packed MO integrals (input of the e2e leg): one pinned copy per node, on rank 0; the other
    # ranks receive it over NVLink inside afesp_gpu_set_eri_mo
    src = None
    if rank == 0:
        pinned = torch.empty(npk, dtype=torch.float64).pin_memory()
        src = pinned.numpy()
    if big:
        gpu.synth_eri_ao(n, Bfac, Cmo)
        gpu.ao2mo(n, want_result=False)
    else:
        gpu.ao2mo(n, eri, Cmo, want_result=False)   # also leaves AO integrals + C resident
    gpu.ao2mo(n)                                     # resident repeat, device-timed
    ao2mo_ms = gpu.last_stage_ms()
    if rank == 0:
        gpu.get_eri_mo(src)
    gpu.release("eri_ao")
    e_mp2 = gpu.mp2_energy(o, eps)
    gpu.set_option("finalize_keep_ccsd", 1)
    e_mp1, _ = gpu.ccsd_init(o, True, eps, 8)
    if big:
        gpu.release("eri_mo")
    comp = {"ccsd": [], "diis": [], "t": []}
    gstat = {"ccsd": [0.0, 0.0, 0], "t": [0.0, 0.0, 0]}   # DMMA GEMM (ms, flop, launches) per stage of the timed region
    last = {}

    def step(record):
        e, rms = gpu.ccsd_iterate()
        a = gpu.last_stage_ms()
        gpu.ccsd_diis()
        b = gpu.last_stage_ms()
        gpu.ccsd_finalize()
        if record:
            ms, fl, nl = gpu.gemm_stats()
            gstat["ccsd"][0] += ms; gstat["ccsd"][1] += fl; gstat["ccsd"][2] += nl
        sums, _ = gpu.ccsd_t_spatial(True, False, False)
        c = gpu.last_stage_ms()
        if record:
            ms, fl, nl = gpu.gemm_stats()
            gstat["t"][0] += ms; gstat["t"][1] += fl; gstat["t"][2] += nl
        last.update(e_ccsd=e, rms=rms, e_T=float(sums[0]))
        if record:
            comp["ccsd"].append(a); comp["diis"].append(b); comp["t"].append(c)

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0, f0 = gpu.counters()
    gpu.set_option("gemm_timing", 1)
    gpu.timer_start()                      # CUDA events on the stream the kernels are launched on
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)
    dev_ms = gpu.timer_stop()
    barrier()
    wall = time.perf_counter() - t0
    gpu.set_option("gemm_timing", 0)
    l1, f1 = gpu.counters()
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([dev_ms * 1e-3, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    elapsed, wall = float(tt[0].item()), float(tt[1].item())
    value = elapsed / args.steps

    # ---- end-to-end leg through the C ABI with host buffers (pinned): H2D of the step's MO integrals, D2H of amplitudes
    if big:
        gpu.set_option("finalize_keep_ccsd", 0)   # free the CCSD work arrays before (T): the next step re-initialises
    ksteps = args.steps
    parts = {"h2d_set_eri_mo": 0.0, "ccsd_init": 0.0, "iterate+diis": 0.0, "finalize_d2h": 0.0, "ccsd_t": 0.0}

    def lap(key, t_prev):
        t = time.perf_counter()
        parts[key] += t - t_prev
        return t

    barrier()
    t0 = time.perf_counter()
    for _ in range(ksteps):
        t = time.perf_counter()
        gpu.set_eri_mo(n, src)
        t = lap("h2d_set_eri_mo", t)
        gpu.ccsd_init(o, True, eps, 8)
        if big:
            gpu.release("eri_mo")
        t = lap("ccsd_init", t)
        gpu.ccsd_iterate()
        gpu.ccsd_diis()
        t = lap("iterate+diis", t)
        _, t1h, t2h = gpu.ccsd_finalize(want_amplitudes=True)
        t = lap("finalize_d2h", t)
        gpu.ccsd_t_spatial(True, False, False)
        t = lap("ccsd_t", t)
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    te = torch.tensor([e2e_elapsed], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = float(te.item()) / ksteps
    h2d = int(npk * 8 + n * 8)
    d2h = int((o * o * v * v + o * v) * 8 + 10 * 8)

    # ---- HBM-bound kernels of the path (permute / denominators / energy), device resident, vs the measured copy peak
    hbm = None
    if rank == 0:
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        except Exception:
            hbm_peak, peak_src = 6452.0, "fallback 6452 GB/s (B200_PROFILING.md measured copy)"
        hbm = {"peak_gbs": hbm_peak, "peak_source": peak_src, "kernels": {}}
        for (oo_, vv_) in sorted({(o, v), (40, 360)}):
            for what in ("permute:3412", "permute:2143", "permute_acc:2143", "divide", "energy", "axpby"):
                try:
                    ms, by = gpu.bench_hbm(what, oo_, vv_, reps=20)
                    hbm["kernels"][f"{what} o={oo_} v={vv_}"] = {
                        "ms": ms, "algorithmic_bytes": by, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / hbm_peak}
                except Exception as ex:
                    hbm["kernels"][f"{what} o={oo_} v={vv_}"] = {"error": str(ex)}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            try:
                cpu = cpu_sample(n, o, 20.0)
            except Exception as ex:  # the CPU leg must never take the GPU line down
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_gemm_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch", {}).get(f"nbf{n}", {}).get(
                    "tma" if tma_scope >= 1 else "cpasync")
            except Exception:
                traffic = None
        tot_ms = gstat["ccsd"][0] + gstat["t"][0]
        tot_fl = gstat["ccsd"][1] + gstat["t"][1]
        tf = lambda g: (g[1] / (g[0] * 1e-3) / 1e12) if g[0] > 0 else None
        # dominant kernel: the batched (T) GEMM  C_pqr(x,(y,z)) = Acat(x,[d|l]) Bcat([d|l],(y,z)), M=v, N=v^2, K=nbf
        t_ms, t_fl, t_nl = gstat["t"]
        achieved = tf(gstat["t"])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic integrals nbf={n} nocc={o} CCSD(T)_spatial", "nbf": n, "nocc": o,
                       "calc_type": "CCSD(T)_spatial",
                       "parallelism": (f"(T) ijk round-robin x{world}; CCSD GEMMs column-sharded x{world} with NCCL slab "
                                       f"exchange, V+/- ladder integrals sharded by column block") if world > 1
                       else "single GPU",
                       "l2": "inputs larger than L2 (packed ladder integrals %.1f GB, (T) work buffers %.1f GB)" % (
                           v ** 4 * 4 / 1e9, 6.0),
                       "timing": "CUDA events on the engine's stream around the K steps, max over ranks"},
            "wall_s_per_step": wall / args.steps,
            "ccsd_s_per_iter": (float(np.mean(comp["ccsd"])) + float(np.mean(comp["diis"]))) / 1e3,
            "t_wall_s": float(np.mean(comp["t"])) / 1e3, "ao2mo_s": ao2mo_ms / 1e3,
            "energies": {"e_mp2": e_mp2, "e_mp1": e_mp1, **last},
            "gemm_tflops_executed": {"all": (tot_fl / (tot_ms * 1e-3) / 1e12) if tot_ms > 0 else None,
                                     "ccsd": tf(gstat["ccsd"]), "t": tf(gstat["t"]),
                                     "gemm_share_of_step": tot_ms * 1e-3 / elapsed if elapsed > 0 else None},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": ("gemm_f64_tma<N,N>" if tma_scope >= 1 else "gemm_f64_dmma<64,64,16> (cp.async ring)") +
                                   f" batched over (ijk-permutations): (T) contraction M={v} N={v * v} K={n} per block",
                         "launches": t_nl, "flops_per_launch": (t_fl / t_nl) if t_nl else None,
                         "ms_per_launch": (t_ms / t_nl) if t_nl else None,
                         "share_of_step": t_ms * 1e-3 / elapsed if elapsed > 0 else None,
                         "all_gemms": {"achieved": (tot_fl / (tot_ms * 1e-3) / 1e12) if tot_ms > 0 else None,
                                       "flops_per_step": tot_fl / args.steps, "ms_per_step": tot_ms / args.steps},
                         "peak_source": "in-run register-resident DMMA.8x8x4 issue-rate probe (MEASURED_PEAKS.json has "
                                        "no FP64 entry; vendor FP64 tensor figure 40 TFLOP/s)"},
            "cpu_baseline": ({"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"],
                              "sample": cpu["sample"], "ccsd_s_per_iter": cpu.get("ccsd_s_per_iter"),
                              "t_wall_s": cpu.get("t_wall_s")} if cpu else None),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "set_eri_mo(H2D from pinned host memory on rank 0, NVLink broadcast to the other ranks) + ccsd_init + iterate + diis + finalize(D2H t1,t2) + ccsd_t",
                    "breakdown_s": {k: x / ksteps for k, x in parts.items()}},
            "gpu_launches": int(l1 - l0),
            "clocks": clocks,
            "tma": dict(zip(("scope", "selftest"), gpu.tma_status())),
            "hbm_kernels": hbm,
        }
        if world > 1:
            line["exchange"] = "NCCL grouped broadcasts of GEMM column slabs on the compute stream"
        emit(line)
    gpu.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nbf", type=int, default=200)
    ap.add_argument("--nocc", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--tma", type=int, default=-1, help="gemm_use_tma: -1 library default (1: TMA-staged kernel for the "
                                                        "(T) batches), 0 cp.async kernels only, 2 TMA for every aligned GEMM")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
