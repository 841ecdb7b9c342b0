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

from afesp_b200 import AfespGpu, synthetic
from tests import _partition as partition
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


# ---------------------------------------------------------------------------------------------------------------
# Column-sharded GEMM and the AO->MO all-to-all plan (afesp_b200/csrc/contract.cu: dgemm_sharded,
# integrals.cu: ao2mo_packed with a communicator), restated on the CPU over gloo with the same partition.
def test_column_partition_tiles_the_range():
    for ncols in [0, 1, 63, 64, 65, 1000, 16290, 64980]:
        for nranks in [1, 2, 3, 8]:
            for gran in [16, 64]:
                lib = AfespGpu.column_partition(ncols, nranks, gran)
                py = [partition.column_range(ncols, r, nranks, gran) for r in range(nranks)]
                assert lib == py
                assert lib[0][0] == 0 and lib[-1][1] == ncols
                for (a, b), (c, d) in zip(lib[:-1], lib[1:]):
                    assert b == c and a <= b
                assert all(a % gran == 0 for a, _ in lib if a < ncols)


def _pair_list(n):
    return [(i, j) for i in range(n) for j in range(i + 1)]


def _dist_worker(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(11)
    # --- sharded GEMM: every rank computes its column slab, slabs are broadcast by their owners
    M, N, K = 37, 300, 53
    A, B = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    Cm = np.zeros((M, N))
    lo, hi = partition.column_range(N, rank, world, 64)
    Cm[:, lo:hi] = A @ B[:, lo:hi]
    for r in range(world):
        a, b = partition.column_range(N, r, world, 64)
        if b > a:
            t = torch.from_numpy(np.ascontiguousarray(Cm[:, a:b]))
            dist.broadcast(t, r)
            Cm[:, a:b] = t.numpy()
    gemm_err = float(np.max(np.abs(Cm - A @ B)))
    # --- AO->MO: phase 1 on my kl pairs, exchange, phase 2 on my pq pairs, owners broadcast their packed rows
    n, o = 8, 2
    eri, Cmo, _ = synthetic.make(n, o, seed=3)
    g = orc.unpack_eri(eri, n)                       # g[i,j,k,l] = (ij|kl)
    pairs = _pair_list(n)
    npair = len(pairs)
    rows = [partition.column_range(npair, r, world, 16) for r in range(world)]
    k0, k1 = rows[rank]
    Hloc = np.zeros((k1 - k0, npair))                # H(kl in mine, pq)
    for x, (k, l) in enumerate(pairs[k0:k1]):
        half = Cmo @ g[:, :, k, l] @ Cmo.T           # (pq) <- sum_ij C(p,i) C(q,j) (ij|kl)
        Hloc[x, :] = [half[p, q] for (p, q) in pairs]
    Hcols = np.zeros((npair, k1 - k0))               # H(all kl, pq in mine) after the exchange
    for src in range(world):
        a, b = rows[src]
        for dst in range(world):
            c, d = rows[dst]
            if src == rank and dst == rank:
                Hcols[a:b, :] = Hloc[:, c:d]
            elif src == rank:
                dist.send(torch.from_numpy(np.ascontiguousarray(Hloc[:, c:d])), dst)
            elif dst == rank:
                t = torch.zeros((b - a, k1 - k0), dtype=torch.float64)
                dist.recv(t, src)
                Hcols[a:b, :] = t.numpy()
    packed = np.zeros(npair * (npair + 1) // 2)
    for x, (p, q) in enumerate(pairs[k0:k1]):
        full = np.zeros((n, n))
        for y, (k, l) in enumerate(pairs):
            full[k, l] = full[l, k] = Hcols[y, x]
        mo = Cmo @ full @ Cmo.T
        pq = k0 + x
        for rs, (r_, s_) in enumerate(pairs[: pq + 1]):
            packed[pq * (pq + 1) // 2 + rs] = mo[r_, s_]
    for r in range(world):
        a, b = rows[r]
        a, b = a * (a + 1) // 2, b * (b + 1) // 2
        if b > a:
            t = torch.from_numpy(packed[a:b].copy())
            dist.broadcast(t, r)
            packed[a:b] = t.numpy()
    ao2mo_err = float(np.max(np.abs(packed - orc.ao2mo_packed(eri, Cmo))))
    if rank == 0:
        out_q.put((gemm_err, ao2mo_err))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_gemm_and_ao2mo_exchange_match_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gemm_err, ao2mo_err = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gemm_err < 1e-12
    assert ao2mo_err < 1e-12


# ---------------------------------------------------------------------------------------------------------------
# afesp_gpu_set_eri_mo with a communicator (capi.cu): every rank uploads the slice [np*r/R, np*(r+1)/R) of the packed
# array over its own PCIe link, then each slice is broadcast from its owner -- restated over gloo; and the in-place
# sharded GEMM with beta (contract.cu: every rank updates ITS column slab of the replicated C in place, beta included,
# then the slabs are broadcast): all ranks must end with the same, correct matrix.
def _worker_slices(rank, world, port, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    npk = 1003                                    # not divisible by the world size
    host = np.arange(npk, dtype=np.float64) * 0.5 + 1.0      # the array every rank can read (shared host copy)
    dev = np.full(npk, np.nan)                    # this rank's "device" copy
    ranges = [(npk * r // world, npk * (r + 1) // world) for r in range(world)]
    lo, hi = ranges[rank]
    dev[lo:hi] = host[lo:hi]                      # own share only
    for r, (a, b) in enumerate(ranges):
        t = torch.from_numpy(dev[a:b].copy())
        dist.broadcast(t, src=r)
        dev[a:b] = t.numpy()
    ok_upload = bool(np.array_equal(dev, host)) and ranges[0][0] == 0 and ranges[-1][1] == npk
    # column-sharded C = alpha A B + beta C, in place
    rng = np.random.default_rng(3)
    M, N, K, alpha, beta = 7, 200, 5, 0.5, -0.25
    A, B, C0 = rng.standard_normal((M, K)), rng.standard_normal((K, N)), rng.standard_normal((M, N))
    C = C0.copy()
    cols = [partition.column_range(N, r, world, 64) for r in range(world)]
    c0, c1 = cols[rank]
    C[:, c0:c1] = alpha * A @ B[:, c0:c1] + beta * C[:, c0:c1]
    for r, (a, b) in enumerate(cols):
        if b > a:
            t = torch.from_numpy(np.ascontiguousarray(C[:, a:b]))
            dist.broadcast(t, src=r)
            C[:, a:b] = t.numpy()
    ok_gemm = bool(np.max(np.abs(C - (alpha * A @ B + beta * C0))) < 1e-14)
    flag = torch.tensor([1.0 if (ok_upload and ok_gemm) else 0.0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out_q.put(float(flag.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sliced_upload_and_in_place_sharded_gemm():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_slices, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok == 1.0
