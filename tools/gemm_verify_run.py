"""AFESP_GEMM_VERIFY=1 python tools/gemm_verify_run.py : one CCSD chain at NBF/NOCC with every unbatched TMA GEMM cross-checked."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afesp_b200 import AfespGpu, synthetic
n, o = int(os.environ.get("NBF", 128)), int(os.environ.get("NOCC", 12))
eri, Cmo, eps = synthetic.make(n, o)
g = AfespGpu(0)
g.ao2mo(n, eri, Cmo, want_result=False)
g.ccsd_init(o, True, eps, 8)
for it in range(2):
    print("iter", it, g.ccsd_iterate(), flush=True)
    g.ccsd_diis()
g.close()
