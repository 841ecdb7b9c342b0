"""Short program for ncu captures: the ladder-shaped and (T)-shaped DMMA GEMMs, then one full (T) on the synthetic
nbf=200/nocc=20 system (so the fused epilogue kernel runs at its real shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afesp_b200 import AfespGpu, synthetic

n, o = int(os.environ.get("NBF", 200)), int(os.environ.get("NOCC", 20))
v = n - o
g = AfespGpu(0)
print("ladder", g.bench_dgemm("N", "N", o * o, v * v, v * v, reps=2, beta=1.0))
print("T1", g.bench_dgemm("N", "N", v, v * v * 8, v, reps=2))
print("T2", g.bench_dgemm("T", "N", v * v * 8, v, o, reps=2, beta=1.0))
eri, C, eps = synthetic.make(n, o)
g.ao2mo(n, eri, C, want_result=False)
g.ccsd_init(o, True, eps, 8)
g.ccsd_iterate()
g.ccsd_finalize()
sums, _ = g.ccsd_t_spatial(True, False, False)
print("T ms", g.last_stage_ms(), sums[:2])
g.close()
