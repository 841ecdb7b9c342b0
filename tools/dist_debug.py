"""Isolate a sharded-vs-replicated discrepancy: one CCSD iteration under (dist on/off) x (TMA on/off) x (einsum sharding on/off)."""
import os, sys, itertools
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_REAL = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
import torch, torch.distributed as dist
from afesp_b200 import AfespGpu, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, o = int(os.environ.get("NBF", 128)), int(os.environ.get("NOCC", 12))
eri, Cmo, eps = synthetic.make(n, o)
gpu = AfespGpu(local)
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.tensor(list(AfespGpu.comm_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
gpu.comm_init(rank, world, bytes(uid.cpu().tolist()))

def run(dist_on, tma, minflops):
    gpu.set_option("dist_ccsd", dist_on); gpu.set_option("gemm_use_tma", 2 * tma); gpu.set_option("dist_min_flops", minflops)
    mo = gpu.ao2mo(n, eri, Cmo, want_result=True)
    gpu.ccsd_init(o, True, eps, 8)
    e, r = gpu.ccsd_iterate()
    _, t1, t2 = gpu.ccsd_finalize(want_amplitudes=True)
    return mo, t1, t2, e

cases = {"rep_tma": (0, 1, 0.0), "rep_cpasync": (0, 0, 0.0), "shard_all_tma": (1, 1, 0.0), "shard_all_cpasync": (1, 0, 0.0),
         "shard_ladder_only_tma": (1, 1, 1e30), "shard_ladder_only_cpasync": (1, 0, 1e30), "shard_all_tma_again": (1, 1, 0.0)}
res = {k: run(*v) for k, v in cases.items()}
base = res["rep_cpasync"]
lines = []
for k, v in res.items():
    d = [float(np.max(np.abs(a - b))) for a, b in zip(v[:3], base[:3])] + [abs(v[3] - base[3])]
    t = torch.tensor(d, dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lines.append("%-28s vs rep_cpasync: d_eri_mo %.2e d_t1 %.2e d_t2 %.2e d_e %.2e" % ((k,) + tuple(t.tolist())))
if rank == 0:
    _REAL.write("\n".join(lines) + "\n"); _REAL.flush()
gpu.close(); dist.destroy_process_group()
