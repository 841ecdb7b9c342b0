"""How far do two runs that differ only in floating-point summation order drift apart during the DIIS-accelerated
CCSD iterations, and do they meet again at convergence?  (One GPU: TMA kernel vs cp.async kernel, which sum the 16 k's
of a tile in different orders.)  Context: tools/dist_check.py compares mid-iteration amplitudes of sharded and
replicated runs; with 8 ranks the slab widths select other tile shapes than the full-width GEMMs."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afesp_b200 import AfespGpu, synthetic

n, o = int(os.environ.get("NBF", 128)), int(os.environ.get("NOCC", 12))
eri, Cmo, eps = synthetic.make(n, o)
g = AfespGpu(0)
out = {}
for iters in (6, 40):
    runs = []
    for tma in (1, 0):
        g.set_option("gemm_use_tma", 2 * tma)
        g.ao2mo(n, eri, Cmo, want_result=False)
        e0, _ = g.ccsd_init(o, True, eps, 8)
        es = [e0]
        for _ in range(iters):
            e, r = g.ccsd_iterate()
            es.append(e)
            if np.sqrt(r) < 1e-10 and abs(es[-1] - es[-2]) < 1e-12:
                break
            g.ccsd_diis()
        _, t1, t2 = g.ccsd_finalize(want_amplitudes=True)
        runs.append((np.array(es), t2))
    m = min(len(runs[0][0]), len(runs[1][0]))
    out[f"iters<={iters}"] = {"iterations": [len(r[0]) - 1 for r in runs],
                              "max |dE| over iterations": float(np.max(np.abs(runs[0][0][:m] - runs[1][0][:m]))),
                              "|dE| last": float(abs(runs[0][0][m - 1] - runs[1][0][m - 1])),
                              "max |dT2|": float(np.max(np.abs(runs[0][1] - runs[1][1])))}
    print(iters, out[f"iters<={iters}"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/rounding_sensitivity.json", "w"), indent=1)
