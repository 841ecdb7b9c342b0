"""Whole hot path on a large synthetic system with device-generated integrals (nbf >= 200):
AO->MO, MP2, a few CCSD iterations (+DIIS), finalize, (T); prints device times, executed GEMM TFLOP/s and energies."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from afesp_b200 import AfespGpu, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--nbf", type=int, default=400)
ap.add_argument("--nocc", type=int, default=40)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--triples-fraction", type=int, default=1, help="run 1/F of the (T) work units (rank 0 of F) and scale")
ap.add_argument("--out", default="gpurun_out/large.json")
a = ap.parse_args()
n, o = a.nbf, a.nocc
v = n - o
t0 = time.time()
B, C, eps = synthetic.make_factors(n, o)
print("factors %.1fs" % (time.time() - t0), flush=True)
g = AfespGpu(0)
peak = g.dmma_peak()
res = {"nbf": n, "nocc": o, "dmma_peak": peak}
g.set_option("gemm_timing", 1)
g.synth_eri_ao(n, B, C)
res["synth_ms"] = g.last_stage_ms(); g.gemm_time()
g.ao2mo(n, want_result=False)
ms, fl = g.gemm_time()
res["ao2mo_ms"] = g.last_stage_ms(); res["ao2mo_gemm_tflops"] = fl / ms / 1e9
print("ao2mo %.1f ms, gemm %.2f TF/s" % (res["ao2mo_ms"], res["ao2mo_gemm_tflops"]), flush=True)
g.release("eri_ao")
res["e_mp2"] = g.mp2_energy(o, eps)
e1, r1 = g.ccsd_init(o, True, eps, 8)
res["ccsd_init_ms"] = g.last_stage_ms(); g.gemm_time()
g.release("eri_mo")
res["e_mp1"] = e1
its = []
for it in range(a.iters):
    e, r = g.ccsd_iterate()
    t_it = g.last_stage_ms()
    g.ccsd_diis()
    t_d = g.last_stage_ms()
    ms, fl = g.gemm_time()
    its.append({"e": e, "rms": r, "iter_ms": t_it, "diis_ms": t_d, "gemm_ms": ms, "gemm_tflops": fl / ms / 1e9})
    print("iter", it + 1, its[-1], flush=True)
res["iters"] = its
g.ccsd_finalize()
if a.triples_fraction > 1:
    g.set_partition(0, a.triples_fraction)
sums, _ = g.ccsd_t_spatial(True, False, False)
ms, fl = g.gemm_time()
res["t_ms"] = g.last_stage_ms() * a.triples_fraction
res["t_measured_fraction"] = 1.0 / a.triples_fraction
res["t_gemm_tflops"] = fl / ms / 1e9
res["t_gemm_share"] = ms / g.last_stage_ms()
res["e_T_partial" if a.triples_fraction > 1 else "e_T"] = float(sums[0])
launches, flops = g.counters()
res["launches"] = launches
print(json.dumps(res), flush=True)
os.makedirs(os.path.dirname(a.out), exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
g.close()
