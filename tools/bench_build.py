"""BASELINE.json configs[4]: k-means IVF training + batched add() of 10M x 768D at nlist 16384 on N GPUs.

    python tools/bench_build.py                                    # 1 GPU
    python tools/bench_build.py --devices 0,1,2,3,4,5,6,7           # ONE process, one shard per device
                                                                   # (vdb_index_create_sharded): data-parallel,
                                                                   # bit-exact training; every shard assigns
                                                                   # the batch and keeps the lists it owns
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_build.py [--mode distributed|replicated]

Training is replicated (every rank runs the same bit-exact k-means on the same sample: the 16384 sequential
seeding steps do not shard).  add() is data-parallel in `distributed` mode: each rank generates and assigns 1/N of
every batch on the tensor cores and one NCCL all-to-all routes the rows to the ranks that own their lists;
`replicated` mode has every rank assign every row and keep its own lists (no exchange).  Rank 0 prints one JSON line.
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def single_process(a):
    """vdb_index_create_sharded over --devices: train() is data-parallel over the devices (rows sliced for the
    distance updates and the assignment, clusters partitioned for the sums) and bit-identical to one GPU --
    `centroid_sha256` is the same at every device count; add() runs on all shards concurrently."""
    import hashlib
    import numpy as np
    pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
    devs = tuple(int(d) for d in a.devices.split(","))
    torch.cuda.set_device(devs[0])
    dev = torch.device("cuda", devs[0])
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=a.dim, nlist=a.nlist, devices=devs if len(devs) > 1 else (),
                                     device=devs[0]))
    gen = torch.Generator(device=dev).manual_seed(12345)
    xt = torch.randn(a.ntrain, a.dim, generator=gen, device=dev)
    torch.cuda.synchronize()
    t = time.perf_counter()
    ix.train(xt)
    t_train = time.perf_counter() - t
    del xt
    sha = hashlib.sha256(np.ascontiguousarray(ix.centroids).tobytes()).hexdigest()
    gen = torch.Generator(device=dev).manual_seed(777)
    t_add = 0.0
    for lo in range(0, a.rows, a.batch_rows):
        nb = min(a.batch_rows, a.rows - lo)
        x = torch.randn(nb, a.dim, generator=gen, device=dev)
        ids = torch.arange(lo, lo + nb, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        t = time.perf_counter()
        ix.add(x, ids)
        t_add += time.perf_counter() - t
        del x, ids
    st = ix.stats()
    sizes = ix.list_sizes()
    print(json.dumps({"config": f"configs[4] train({a.ntrain}) + add({a.rows}) {a.dim}D nlist={a.nlist}",
                      "n_gpus": len(devs), "mode": "single process, one shard per device" if len(devs) > 1 else "single",
                      "train_s": round(t_train, 3), "add_s": round(t_add, 3), "add_rows_per_s": a.rows / max(t_add, 1e-9),
                      "rows_held_all_shards": int(sizes.sum()), "centroid_sha256": sha,
                      "index_gb_all_shards": round(st.gpu_memory_bytes / 2**30, 2)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nlist", type=int, default=16384)
    ap.add_argument("--ntrain", type=int, default=1_000_000, help="SURVEY 8d: train on the first 1M rows")
    ap.add_argument("--devices", default="", help="comma-separated device list: single-process sharded index")
    ap.add_argument("--batch-rows", type=int, default=1_000_000)
    ap.add_argument("--mode", default="distributed", choices=["distributed", "replicated"])
    a = ap.parse_args()
    if a.devices:
        return single_process(a)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
    sh = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.sharded")
    ix = sh.ShardedIVFFlatIndex(pkg, pkg.Config(dimension=a.dim, nlist=a.nlist, device=local)) if world > 1 else None
    loc = ix.local if ix else pkg.IVFFlatIndex(pkg.Config(dimension=a.dim, nlist=a.nlist))
    gen = torch.Generator(device=dev).manual_seed(12345)
    xt = torch.randn(a.ntrain, a.dim, generator=gen, device=dev)  # same sample on every rank

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if world > 1:  # NCCL builds its channels on first use: keep that out of the timed regions
        w = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(torch.empty_like(w), w)
        dist.all_reduce(w)
    sync()
    t = time.perf_counter()
    loc.train(xt)
    sync()
    t_train = time.perf_counter() - t
    del xt
    # rank-private generators: in distributed mode every rank contributes different rows
    gen = torch.Generator(device=dev).manual_seed(777 + (rank if a.mode == "distributed" else 0))
    t_add = 0.0
    for lo in range(0, a.rows, a.batch_rows):
        nb = min(a.batch_rows, a.rows - lo)
        if world > 1 and a.mode == "distributed":
            per = (nb + world - 1) // world
            mine = max(0, min(per, nb - rank * per))
            x = torch.randn(mine, a.dim, generator=gen, device=dev)
            ids = torch.arange(lo + rank * per, lo + rank * per + mine, dtype=torch.int64, device=dev)
            sync()
            t = time.perf_counter()
            ix.add_distributed(x, ids)
        else:
            x = torch.randn(nb, a.dim, generator=gen, device=dev)
            ids = torch.arange(lo, lo + nb, dtype=torch.int64, device=dev)
            sync()
            t = time.perf_counter()
            loc.add(x, ids)
        sync()
        t_add += time.perf_counter() - t
        del x, ids
    tot = torch.tensor([loc.get_total_vectors() if a.mode == "replicated" or world == 1 else 0,
                        loc.get_gpu_memory_usage()], dtype=torch.int64, device=dev)
    held = torch.tensor([int(loc.list_sizes().sum())], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(held)
    if rank == 0:
        print(json.dumps({"config": f"configs[4] train({a.ntrain}) + add({a.rows}) {a.dim}D nlist={a.nlist}",
                          "n_gpus": world, "mode": a.mode if world > 1 else "single", "train_s": round(t_train, 3),
                          "add_s": round(t_add, 3), "add_rows_per_s": a.rows / t_add,
                          "rows_held_all_ranks": int(held.item()),
                          "rank0_index_gb": round(int(tot[1].item()) / 2**30, 2)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
