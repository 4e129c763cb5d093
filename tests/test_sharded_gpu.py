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
from parity import check_search, ip_scale

HERE = os.path.dirname(os.path.abspath(__file__))
WORLD = 2


def _nccl_answer(ix, qd, nprobe, k):
    """the portable exchange (local search, two all-gathers, merge kernel) on the same shards"""
    ex, ix.exchange = ix.exchange, None
    try:
        return ix.search_device(qd, nprobe, k)
    finally:
        ix.exchange = ex


def _case(pkg, sharded, rank, name, metric, dim, nlist, n, nq, nprobe, k, ntrain, centroids=None):
    x = O.gaussian(31 + dim, n + nq, dim)
    db, q = x[:n], x[n:]
    ix = sharded.ShardedIVFFlatIndex(pkg, pkg.Config(dimension=dim, nlist=nlist, device=rank, metric=pkg.Metric(metric)))
    if centroids is None:
        ix.train(db[:ntrain])
    else:
        ix.local.centroids = centroids
    owners = ix.local.owners()
    # first half replicated add (every rank sees the rows), second half data-parallel add (each rank a slice)
    half = n // 2
    ix.add(db[:half], np.arange(half, dtype=np.uint64))
    mine = np.arange(half + rank, n, WORLD)
    ix.add_distributed(torch.from_numpy(db[mine]).cuda(), torch.from_numpy(mine.astype(np.int64)).cuda())
    assert ix.get_total_vectors() == n
    assert ix.exchange is not None, "an NCCL group of 2 ranks must take the peer-memory exchange"
    D, I = ix.search(q.astype(np.float64), nprobe, k)  # any float dtype in: coerced, not reinterpreted
    qd = torch.from_numpy(q).cuda()
    for rep in range(3):
        Dp, Ip = ix.search_device(qd, nprobe, k)
        Dn, In = _nccl_answer(ix, qd, nprobe, k)
        assert torch.equal(Dp, Dn) and torch.equal(Ip, In), f"{name}: p2p and nccl exchange differ (rep {rep})"
    assert np.array_equal(Dp.cpu().numpy(), D) and np.array_equal(Ip.cpu().numpy().view(np.uint64), I)
    # pipelined collective search: batches in flight, device and host buffers
    bs = 8
    starts = list(range(0, nq - bs + 1, bs))
    Dd = [torch.empty((bs, k), dtype=torch.float32, device="cuda") for _ in starts]
    Id = [torch.empty((bs, k), dtype=torch.int64, device="cuda") for _ in starts]
    tickets = [ix.search_submit(qd[lo:lo + bs], nprobe, k, Dd[i], Id[i]) for i, lo in enumerate(starts)]
    for t in tickets:
        ix.search_wait(t)
    for i, lo in enumerate(starts):
        assert torch.equal(Dd[i], Dp[lo:lo + bs]) and torch.equal(Id[i], Ip[lo:lo + bs]), f"{name}: pipelined differs"
    Dh = np.empty((bs, k), np.float32)
    Ih = np.empty((bs, k), np.uint64)
    ix.search_wait(ix.search_submit(np.ascontiguousarray(q[:bs]), nprobe, k, Dh, Ih))
    assert np.array_equal(Dh, D[:bs]) and np.array_equal(Ih, I[:bs])
    sizes = ix.local.list_sizes()
    assert (sizes[owners != rank] == 0).all()
    tot = torch.tensor([int(sizes.sum())], device="cuda")
    dist.all_reduce(tot)
    assert int(tot.item()) == n
    return ix, db, q, D, I, owners


def _worker(rank, port, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), NCCL_DEBUG="WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
    sharded = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.sharded")
    try:
        # --- L2, trained on the device
        dim, nlist, n, nq, nprobe, k = 64, 48, 20000, 40, 12, 10
        ix, db, q, D, I, owners = _case(pkg, sharded, rank, "l2", O.METRIC_L2, dim, nlist, n, nq, nprobe, k, 4000)
        qd = torch.from_numpy(q).cuda()
        # ragged shapes through the mailbox: fewer queries, larger k (padding travels too)
        D2, I2 = ix.search_device(qd[:7], nlist, 64)
        D3, I3 = _nccl_answer(ix, qd[:7], nlist, 64)
        assert torch.equal(D2, D3) and torch.equal(I2, I3)
        # beyond the mailbox (k > max_k): falls back to the all-gather path and still agrees with the oracle below
        D4, I4 = ix.search_device(qd[:5], nprobe, 100)
        if rank == 0:
            ora = O.OracleIndex(dim, nlist)
            ora.train(db[:4000])
            ora.add(db)
            Dr, Ir = ora.search(q, nprobe, k)
            check_search(D, I, Dr, Ir)
            Dr4, Ir4 = ora.search(q[:5], nprobe, 100)
            check_search(D4.cpu().numpy(), I4.cpu().numpy().view(np.uint64), Dr4, Ir4)
            full = ora.list_sizes().astype(np.int64)
            load = [int(full[owners == r].sum()) for r in range(WORLD)]
            assert max(load) - min(load) <= max(full.max(), n // 20), load  # byte-balanced ownership
        ix.close()
        # --- inner product (BASELINE configs[3] metric), sharded
        dim, nlist, n, nq, nprobe, k = 96, 64, 16000, 24, 16, 10
        ix, db, q, D, I, owners = _case(pkg, sharded, rank, "ip", O.METRIC_IP, dim, nlist, n, nq, nprobe, k, 3000)
        if rank == 0:
            ora = O.OracleIndex(dim, nlist, O.METRIC_IP)
            ora.train(db[:3000])
            ora.add(db)
            Dr, Ir = ora.search(q, nprobe, k)
            check_search(D, I, Dr, Ir, ip_scale(q, db))
        ix.close()
        # --- configs[3]-shaped but small: inner product, nlist 16384, nprobe 64 (centroids = data rows; the
        # oracle's O(nlist^2) seeding is infeasible at this nlist)
        dim, nlist, n, nq, nprobe, k = 16, 16384, 30000, 16, 64, 10
        cent = O.gaussian(31 + dim, n + nq, dim)[:nlist].copy()
        ix, db, q, D, I, owners = _case(pkg, sharded, rank, "c4-shaped", O.METRIC_IP, dim, nlist, n, nq, nprobe, k, 0,
                                        centroids=cent)
        if rank == 0:
            ora = O.OracleIndex(dim, nlist, O.METRIC_IP)
            ora.centroids = cent
            ora.add(db)
            Dr, Ir = ora.search(q, nprobe, k, 8)
            check_search(D, I, Dr, Ir, ip_scale(q, db))
        ix.close()
        if rank == 0:
            ret.put("ok")
    except Exception as e:  # noqa: BLE001
        import traceback
        if rank == 0:
            ret.put(f"FAIL rank {rank}: {e}\n{traceback.format_exc()}")
        raise
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
