"""HBM-bound kernels of the path, device resident, GB/s against the measured copy peak (see afesp_gpu_bench_hbm)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afesp_b200 import AfespGpu
g = AfespGpu(0)
peak = 6452.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
res = {}
for (o, v) in [(20, 180), (40, 360)]:
    for what in ["permute:3412", "permute:2143", "permute_acc:2143", "permute:1243", "permute:1324", "permute:4321",
                 "permute:2134", "divide", "divide_probe", "energy", "axpby"]:
        ms, by = g.bench_hbm(what, o, v, reps=20)
        res[f"{what} o={o} v={v}"] = {"ms": ms, "gbs": by / ms / 1e6, "frac": by / ms / 1e6 / peak}
        print("%-28s o=%d v=%d  %8.3f ms  %7.0f GB/s  %5.1f%%" % (what, o, v, ms, by / ms / 1e6, 100 * by / ms / 1e6 / peak), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"peak_gbs": peak, "kernels": res}, open("gpurun_out/hbm_probe.json", "w"), indent=1)
