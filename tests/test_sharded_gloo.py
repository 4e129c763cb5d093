"""CPU test of the N>1 host logic with world_size 2 over gloo: list ownership,
the all-gather layout, and that merging per-shard top-k's by (distance, id)
reproduces the unsharded reference answer -- i.e. the result does not depend on
the number of GPUs.  The shards' local searches are played by the oracle here
(the CUDA path needs a GPU); the exchange code is the product's."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O

HERE = os.path.dirname(os.path.abspath(__file__))
WORLD = 2


def _worker(rank, port, ret):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    sharded = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.sharded")
    dim, nlist, n, nq, nprobe, k = 24, 10, 3000, 16, 6, 10
    x = O.gaussian(77, n + nq, dim)
    db, q = x[:n], x[n:]
    full = O.OracleIndex(dim, nlist)
    full.train(db[:1000])
    asg = full.assign(db)
    # this rank's shard: the rows of the lists it owns, same centroids
    mine = np.array([sharded.owner_of(l, WORLD) == rank for l in asg])
    shard = O.OracleIndex(dim, nlist)
    shard.centroids = full.centroids
    shard.load_assigned(db[mine], np.nonzero(mine)[0].astype(np.uint64), asg[mine])
    D, I = shard.search(q, nprobe, k)
    Dg, Ig = sharded.gather_topk(torch.from_numpy(D), torch.from_numpy(I.view(np.int64)))
    assert Dg.shape == (WORLD, nq, k)
    assert torch.equal(Dg[rank], torch.from_numpy(D))  # rank order is the gather order
    if rank == 0:
        full.add(db)
        Dr, Ir = full.search(q, nprobe, k)
        Dg, Ig = Dg.numpy(), Ig.numpy().view(np.uint64)
        ok = True
        for qi in range(nq):
            cand = sorted((float(d), int(i)) for p in range(WORLD) for d, i in zip(Dg[p, qi], Ig[p, qi])
                          if i != O.ID_PAD)[:k]
            ok &= [c[1] for c in cand] == [int(v) for v in Ir[qi]]
            ok &= np.array_equal(np.array([c[0] for c in cand], np.float32), Dr[qi])
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_shards_merge_to_the_unsharded_answer():
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, port, ret)) for r in range(WORLD)]
    for p in procs:
        p.start()
    ok = ret.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_owner_of_partitions_all_lists():
    sharded = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.sharded")
    for world in (1, 2, 4, 8):
        owners = [sharded.owner_of(l, world) for l in range(4096)]
        assert set(owners) == set(range(world))
        counts = np.bincount(owners)
        assert counts.max() - counts.min() <= 1
