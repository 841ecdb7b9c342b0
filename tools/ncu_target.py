"""Short program for ncu captures on the synthetic system (NBF/NOCC env, default 200/20).

PROFILE=T     cudaProfilerStart right before the (T) call: the batched (T) GEMM (gemm_f64_tma), the fused
              combine+energy kernel and the partial-sum finish are the first kernels captured.
PROFILE=CCSD  cudaProfilerStart right before one CCSD iteration (ladder, rings, permutes, divide).
Run under `ncu --profile-from-start off ...`; without ncu it just runs the chain."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from afesp_b200 import AfespGpu, synthetic

n, o = int(os.environ.get("NBF", 200)), int(os.environ.get("NOCC", 20))
v = n - o
what = os.environ.get("PROFILE", "T")
rt = torch.cuda.cudart()
g = AfespGpu(0)
if os.environ.get("TMA_SCOPE"):
    g.set_option("gemm_use_tma", int(os.environ["TMA_SCOPE"]))
if n > 240:   # the packed AO integrals are expanded on the device from the low-rank factors
    Bf, C, eps = synthetic.make_factors(n, o)
    g.synth_eri_ao(n, Bf, C)
    g.ao2mo(n, want_result=False)
else:
    eri, C, eps = synthetic.make(n, o)
    g.ao2mo(n, eri, C, want_result=False)
g.release("eri_ao")
g.ccsd_init(o, True, eps, 8)
if n > 240:
    g.release("eri_mo")
g.ccsd_iterate()
g.ccsd_diis()
if what == "CCSD":
    # the two GEMM shapes that carry the CCSD iteration, as the iteration issues them: the (+)-packed ladder
    # S(o^2 x P+) . V+(P+ x P+) and an o^3 v^3 ring (ov x ov x ov); one launch each inside the profiled range
    Pp = v * (v + 1) // 2
    g.bench_dgemm("N", "N", o * o, Pp, Pp, reps=1, beta=0.0)   # warm-up inside bench_dgemm is outside the range below
    rt.cudaProfilerStart()
    print("ladder ms", g.bench_dgemm("N", "N", o * o, Pp, Pp, reps=1, beta=0.0))
    print("ring ms", g.bench_dgemm("N", "N", o * v, o * v, o * v, reps=1, beta=1.0))
    rt.cudaProfilerStop()
g.ccsd_finalize()
if what == "T":
    g.set_option("triples_batch_bytes", float(6 << 30))
    rt.cudaProfilerStart()
sums, _ = g.ccsd_t_spatial(True, False, False)
if what == "T":
    rt.cudaProfilerStop()
print("T ms", g.last_stage_ms(), sums[:2])
g.close()
