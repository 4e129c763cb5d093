"""GPU test of the sharded path (needs >= 2 GPUs; skipped otherwise): two NCCL ranks, byte-balanced list
ownership, local scans + all-gather + merge kernel == the unsharded index == the oracle."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from parity import check_search

HERE = os.path.dirname(os.path.abspath(__file__))
WORLD = 2


def _worker(rank, port, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
    sharded = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.sharded")
    dim, nlist, n, nq, nprobe, k = 64, 48, 20000, 40, 12, 10
    x = O.gaussian(31, n + nq, dim)
    db, q = x[:n], x[n:]
    ix = sharded.ShardedIVFFlatIndex(pkg, pkg.Config(dimension=dim, nlist=nlist, device=rank))
    ix.train(db[:4000])
    owners = ix.local.owners()
    # first half replicated add (every rank sees the rows), second half data-parallel add (each rank a slice)
    half = n // 2
    ix.add(db[:half], np.arange(half, dtype=np.uint64))
    mine = np.arange(half + rank, n, WORLD)
    ix.add_distributed(torch.from_numpy(db[mine]).cuda(), torch.from_numpy(mine.astype(np.int64)).cuda())
    assert ix.get_total_vectors() == n
    assert ix.exchange is not None, "an NCCL group of 2 ranks must take the peer-memory exchange"
    D, I = ix.search(q, nprobe, k)
    # the portable exchange (two all-gathers + merge kernel) must give the same bits, call after call
    qd = torch.from_numpy(q).cuda()
    for rep in range(5):
        Dp, Ip = ix.search_device(qd, nprobe, k)
        ex, ix.exchange = ix.exchange, None
        Dn, In = ix.search_device(qd, nprobe, k)
        ix.exchange = ex
        assert torch.equal(Dp, Dn) and torch.equal(Ip, In), f"p2p and nccl exchange differ (rep {rep})"
    assert np.array_equal(Dp.cpu().numpy(), D) and np.array_equal(Ip.cpu().numpy().view(np.uint64), I)
    # ragged shapes through the mailbox: fewer queries, larger k (padding travels too)
    D2, I2 = ix.search_device(qd[:7], nlist, 64)
    ex, ix.exchange = ix.exchange, None
    D3, I3 = ix.search_device(qd[:7], nlist, 64)
    ix.exchange = ex
    assert torch.equal(D2, D3) and torch.equal(I2, I3)
    # publish / collect as two launches with the next batch's search in between (one batch in flight)
    ex = ix.exchange
    st = torch.cuda.current_stream().cuda_stream
    batches = [qd[lo:lo + 10] for lo in (0, 10, 20, 30)]
    want = [ix.search_device(b, nprobe, k) for b in batches]
    got, Dl, Il = [], [None, None], [None, None]
    for i, b in enumerate(batches):
        Dl[i & 1] = torch.empty((10, k), dtype=torch.float32, device="cuda")
        Il[i & 1] = torch.empty((10, k), dtype=torch.int64, device="cuda")
        ix.local.search_async(b, nprobe, k, Dl[i & 1], Il[i & 1], st)
        if i:
            Do = torch.empty((10, k), dtype=torch.float32, device="cuda")
            Io = torch.empty((10, k), dtype=torch.int64, device="cuda")
            ex.collect_into(Do, Io, st)
            got.append((Do, Io))
        ex.publish(Dl[i & 1], Il[i & 1], st)
    Do = torch.empty((10, k), dtype=torch.float32, device="cuda")
    Io = torch.empty((10, k), dtype=torch.int64, device="cuda")
    ex.collect_into(Do, Io, st)
    got.append((Do, Io))
    torch.cuda.synchronize()
    for (Dw, Iw), (Dg_, Ig_) in zip(want, got):
        assert torch.equal(Dw, Dg_) and torch.equal(Iw, Ig_), "pipelined exchange differs from the one-launch form"
    sizes = ix.local.list_sizes()
    assert (sizes[owners != rank] == 0).all()
    tot = torch.tensor([int(sizes.sum())], device="cuda")
    dist.all_reduce(tot)
    if rank == 0:
        ora = O.OracleIndex(dim, nlist)
        ora.train(db[:4000])
        ora.add(db)
        Dr, Ir = ora.search(q, nprobe, k)
        try:
            assert int(tot.item()) == n
            check_search(D, I, Dr, Ir)
            full = ora.list_sizes().astype(np.int64)
            load = [int(full[owners == r].sum()) for r in range(WORLD)]
            assert max(load) - min(load) <= max(full.max(), n // 20), load  # byte-balanced ownership
            ret.put("ok")
        except AssertionError as e:
            ret.put(f"FAIL {e}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_gpu_sharded_search_matches_oracle():
    if torch.cuda.device_count() < WORLD:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, port, ret)) for r in range(WORLD)]
    for p in procs:
        p.start()
    msg = ret.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
    assert msg == "ok", msg
