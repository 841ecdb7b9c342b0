"""First-contact probe on the B200: DMMA issue-rate peak and a few GEMM shapes (device-timed)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afesp_b200 import AfespGpu

g = AfespGpu(0)
peak = g.dmma_peak()
peak2 = g.dmma_peak()
out = {"dmma_peak_tflops": [peak, peak2], "gemm": []}
shapes = [("N", "N", 4096, 4096, 4096), ("N", "N", 400, 32400, 32400), ("T", "N", 4096, 4096, 4096),
          ("N", "T", 4096, 4096, 4096), ("T", "T", 4096, 4096, 4096), ("N", "N", 180, 32400, 180),
          ("T", "N", 32400, 180, 20), ("N", "N", 3600, 3600, 3600), ("N", "N", 8192, 8192, 8192),
          ("N", "N", 20, 648000, 180), ("N", "N", 180, 180, 72000)]
for ta, tb, M, N, K in shapes:
    ms = g.bench_dgemm(ta, tb, M, N, K, reps=3)
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    out["gemm"].append({"t": ta + tb, "M": M, "N": N, "K": K, "ms": ms, "tflops": tf})
    print(ta + tb, M, N, K, "%.3f ms  %.2f TFLOP/s" % (ms, tf), flush=True)
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
