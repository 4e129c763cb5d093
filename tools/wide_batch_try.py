"""One wide batch (the reference's bench/benchmark.cpp searches 10 000 queries in one call) on an index with and
without the shadow: with it the batch is answered in pipelined 64-query chunks by the tensor-core screen, without it
by one pass of the fp32 scan kernel.  python tools/wide_batch_try.py [rows] [queries]"""
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dim, nlist, nprobe, k = 768, max(64, n // 2441), 32, 10
res = {}
for mirror in ("2", "0"):
    os.environ["VDB_SCAN_MIRROR"] = mirror
    g = torch.Generator(device="cuda").manual_seed(1)
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    x = torch.randn(262144, dim, generator=g, device="cuda")
    torch.cuda.synchronize()
    ix.train(x)
    for lo in range(0, n, 500_000):
        y = torch.randn(min(500_000, n - lo), dim, generator=g, device="cuda")
        torch.cuda.synchronize()
        ix.add(y)
    q = torch.randn(nq, dim, generator=g, device="cuda")
    torch.cuda.synchronize()
    ix.search(q[:64], nprobe, k)
    t = time.perf_counter()
    D, I = ix.search(q, nprobe, k)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    res[mirror] = (D.cpu(), I.cpu())
    print(f"{nq} queries in one call, {n} rows, shadow {'int8' if mirror == '2' else 'none'}: {dt * 1e3:.1f} ms = {nq / dt:.0f} QPS", flush=True)
    del ix
    torch.cuda.empty_cache()
print("identical:", bool(torch.equal(res["2"][0], res["0"][0]) and torch.equal(res["2"][1], res["0"][1])))
