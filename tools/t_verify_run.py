"""AFESP_T_VERIFY=1 python tools/t_verify_run.py : every (T) batch GEMM through both kernels, compared on the device."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afesp_b200 import AfespGpu, synthetic
n, o = int(os.environ.get("NBF", 200)), int(os.environ.get("NOCC", 20))
g = AfespGpu(0)
g.set_option("gemm_use_tma", int(os.environ.get("TMA", 1)))   # the TMA path is opt-in; this tool exists to check it
print("tma status (scope, selftest):", g.tma_status(), flush=True)
if n > 240:
    Bfac, Cmo, eps = synthetic.make_factors(n, o)
    g.synth_eri_ao(n, Bfac, Cmo)
    g.ao2mo(n, want_result=False)
else:
    eri, Cmo, eps = synthetic.make(n, o)
    g.ao2mo(n, eri, Cmo, want_result=False)
g.release("eri_ao")
part = int(os.environ.get("PART", 1))
if part > 1:
    g.set_partition(0, part)      # only every part-th triple (keeps a large shape short)
g.set_option("finalize_keep_ccsd", 1)
g.ccsd_init(o, True, eps, 8)
if n > 240:
    g.release("eri_mo")
for it in range(2):
    g.ccsd_iterate(); g.ccsd_diis()
g.ccsd_finalize()
for rep in range(int(os.environ.get("REPS", 3))):
    sums, _ = g.ccsd_t_spatial(True, False, False)
    print("rep", rep, "e_T", repr(float(sums[0])), flush=True)
g.close()
