"""Multi-rank (T) sharding logic on the CPU: world_size 2 over gloo.

Each rank takes its round-robin share of the i<=j<=k triples (the same deal the device code uses), evaluates the
reference's ordered-triple loop on the orbits of its share with the NumPy oracle, and the six sums are combined with an
all_reduce(SUM) -- the collective the library performs with NCCL on GPUs.  The result must equal the single-rank sums."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from afesp_b200 import AfespGpu, partition, synthetic
from oracle import afesp_oracle as orc


def _worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, o = 12, 3
    eri, Cm, eps = synthetic.make(n, o, seed=5)
    mo = orc.ao2mo_packed(eri, Cm)
    cc = orc.ccsd_spatial(mo, eps, o, 1e-9, 1e-10, 8, 50, want_cr=True)
    V = cc["V"]
    mine = partition.my_triples(o, rank, world)
    ordered = [t for (i, j, k, w) in mine for t in partition.orbit(i, j, k)]
    sums = np.array(orc.triples_spatial_sums(cc["t1"], cc["t2"], V["v_oovv"], V["v_vvov"], V["v_oovo"], eps, True, False,
                                             True, cc["I_vovv_pp"], cc["I_ooov_pp"], triples=ordered))
    t = torch.from_numpy(sums.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        full = np.array(orc.triples_spatial_sums(cc["t1"], cc["t2"], V["v_oovv"], V["v_vvov"], V["v_oovo"], eps, True,
                                                 False, True, cc["I_vovv_pp"], cc["I_ooov_pp"]))
        out_q.put((t.numpy().tolist(), full.tolist(), len(ordered)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_triples_allreduce_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, full, nmine = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.max(np.abs(np.array(got) - np.array(full))) < 1e-13
    assert 0 < nmine < 27


def test_python_partition_mirrors_the_library():
    for o in [1, 3, 7, 20]:
        for nranks in [1, 2, 4, 8]:
            lib = AfespGpu.triples_partition(o, nranks)
            py = [len(partition.my_triples(o, r, nranks)) for r in range(nranks)]
            assert lib == py
            weights = sum(w for r in range(nranks) for (_, _, _, w) in partition.my_triples(o, r, nranks))
            assert weights == o ** 3  # the orbits tile the reference's full o^3 loop
