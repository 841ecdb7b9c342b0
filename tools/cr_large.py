"""CR-CCSD(T) at the target shape: CR intermediates without the dense v^4 slice + the doubled (T) GEMM stream, on a
fraction of the triples (PART) to keep the run short.  Prints stage times and the device memory high-water mark."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from afesp_b200 import AfespGpu, synthetic
n, o = int(os.environ.get("NBF", 400)), int(os.environ.get("NOCC", 40))
part = int(os.environ.get("PART", 32))
g = AfespGpu(0)
Bfac, Cmo, eps = synthetic.make_factors(n, o)
g.synth_eri_ao(n, Bfac, Cmo)
g.ao2mo(n, want_result=False)
g.release("eri_ao")
out = {"nbf": n, "nocc": o, "part": part, "ao2mo_ms": g.last_stage_ms()}
g.ccsd_init(o, True, eps, 8)
for it in range(2):
    e, r = g.ccsd_iterate(); g.ccsd_diis()
out["ccsd_iter_ms"] = g.last_stage_ms(); out["e_ccsd_2it"] = e
t0 = time.perf_counter()
d, _, _ = g.ccsd_finalize(want_cr=True)
out["finalize_cr_s"] = time.perf_counter() - t0
g.set_partition(0, part)
t0 = time.perf_counter()
sums, const = g.ccsd_t_spatial(True, False, True)
out["crccsd_t_partial_s"] = time.perf_counter() - t0
out["t_stage_ms"] = g.last_stage_ms()
out["sums_partial"] = [float(x) for x in sums]; out["D_const"] = const
free, total = torch.cuda.mem_get_info(0)
out["device_mem_used_GB_after"] = (total - free) / 1e9
print(json.dumps(out), flush=True)
g.close()
