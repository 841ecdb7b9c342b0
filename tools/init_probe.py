"""Where does afesp_gpu_ccsd_init spend its time? (AFESP_TRACE=1)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afesp_b200 import AfespGpu, synthetic
n, o = int(os.environ.get("NBF", 200)), int(os.environ.get("NOCC", 20))
g = AfespGpu(0)
eri, C, eps = synthetic.make(n, o)
g.ao2mo(n, eri, C, want_result=False)
g.release("eri_ao")
for keep in (1, 0):
    g.set_option("finalize_keep_ccsd", keep)
    for rep in range(3):
        t0 = time.perf_counter(); g.ccsd_init(o, True, eps, 8); t1 = time.perf_counter()
        print(f"keep={keep} rep={rep} ccsd_init wall {1e3*(t1-t0):.1f} ms, stage {g.last_stage_ms():.1f} ms", file=sys.stderr, flush=True)
        g.ccsd_iterate(); g.ccsd_diis(); g.ccsd_finalize()
        t0 = time.perf_counter(); g.ccsd_t_spatial(True, False, False); t1 = time.perf_counter()
        print(f"   ccsd_t wall {1e3*(t1-t0):.1f} ms", file=sys.stderr, flush=True)
g.close()
