"""Time every tile config of the DMMA GEMM on the shapes that dominate the hot path (device-timed)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afesp_b200 import AfespGpu

g = AfespGpu(0)
peak = g.dmma_peak()
print("dmma peak", peak)
shapes = [
    ("T1 particle n200 x16", "N", "N", 180, 32400 * 16, 180, 0.0),
    ("T2 hole n200 x16", "T", "N", 32400 * 16, 180, 20, 1.0),
    ("ladder n200", "N", "N", 400, 32400, 32400, 1.0),
    ("ring n200", "N", "N", 3600, 3600, 3600, 1.0),
    ("square 4096", "N", "N", 4096, 4096, 4096, 0.0),
    ("T1 particle n400 x2", "N", "N", 360, 129600 * 2, 360, 0.0),
    ("ladder n400-ish", "N", "N", 1600, 20000, 20000, 1.0),
]
res = []
for name, ta, tb, M, N, K, beta in shapes:
    row = {"shape": name, "M": M, "N": N, "K": K, "cfg": {}}
    for cfg in [-1, 0, 1, 2, 3, 4, 5, 6]:
        g.set_option("gemm_force_config", cfg)
        try:
            ms = g.bench_dgemm(ta, tb, M, N, K, reps=2, beta=beta)
        except Exception as e:
            print(name, cfg, "failed", e); continue
        tf = 2.0 * M * N * K / ms / 1e9
        row["cfg"][cfg] = {"ms": ms, "tflops": tf}
        print("%-24s cfg %2d  %9.3f ms  %6.2f TF/s  (%.0f%% of DMMA peak)" % (name, cfg, ms, tf, 100 * tf / peak), flush=True)
    res.append(row)
g.set_option("gemm_force_config", -1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"peak": peak, "rows": res}, open("gpurun_out/gemm_tune.json", "w"), indent=1)
