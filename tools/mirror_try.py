"""Quick look at the bf16 tensor-core screen (csrc/screen.cuh) on one GPU: result equality against the fp32 scan on a
few shapes, then the two scans timed on a headline-shaped shard.  Development aid (tests/test_scan_mirror.py is the
parity test proper):  python tools/mirror_try.py [rows_for_timing]"""
import importlib
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")


KIND = int(os.environ.get("MIRROR_KIND", "1"))  # 1 = bf16, 2 = int8


def make(dim, nlist, metric, mirror):
    os.environ["VDB_SCAN_EXACT"] = "0" if mirror else "1"
    os.environ["VDB_SCAN_MIRROR"] = str(KIND) if mirror else "0"
    return pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist, metric=metric))


def check(dim, nlist, n, nq, nprobe, k, metric):
    g = torch.Generator(device="cuda").manual_seed(dim + n)
    db = torch.randn(n, dim, generator=g, device="cuda")
    q = torch.randn(nq, dim, generator=g, device="cuda")
    torch.cuda.synchronize()
    res = []
    for mirror in (True, False):
        ix = make(dim, nlist, metric, mirror)
        ix.centroids = db[:nlist].cpu().numpy().copy()
        ix.add(db[: n // 3])
        ix.add(db[n // 3:])
        t = time.perf_counter()
        D, I = ix.search(q, nprobe, k)
        torch.cuda.synchronize()
        res.append((D.cpu().numpy(), I.cpu().numpy(), time.perf_counter() - t))
    (Da, Ia, ta), (Db, Ib, tb) = res
    same = np.array_equal(Da, Db) and np.array_equal(Ia, Ib)
    print(f"dim {dim} nlist {nlist} n {n} nq {nq} nprobe {nprobe} k {k} metric {int(metric)}: "
          f"{'IDENTICAL' if same else 'DIFFERENT'}  (mirror {ta * 1e3:.1f} ms, fp32 {tb * 1e3:.1f} ms)", flush=True)
    if not same:
        bad = np.nonzero((Da != Db).any(1) | (Ia != Ib).any(1))[0]
        print("  queries that differ:", bad[:10], "of", nq)
        for qi in bad[:2]:
            print("  mirror", Da[qi][:6], Ia[qi][:6])
            print("  fp32  ", Db[qi][:6], Ib[qi][:6])
    return same


def timing(n):
    dim, nlist, nq, nprobe, k = 768, max(64, n // 2441), 64, 32, 10
    g = torch.Generator(device="cuda").manual_seed(1)
    out = {}
    for mirror in (True, False):
        ix = make(dim, nlist, pkg.Metric.L2, mirror)
        x = torch.randn(262144, dim, generator=g, device="cuda")
        torch.cuda.synchronize()
        ix.train(x)
        for lo in range(0, n, 500_000):
            y = torch.randn(min(500_000, n - lo), dim, generator=g, device="cuda")
            torch.cuda.synchronize()
            ix.add(y)
        qs = [torch.randn(nq, dim, generator=g, device="cuda") for _ in range(8)]
        torch.cuda.synchronize()
        D = torch.empty(nq, k, device="cuda")
        I = torch.empty(nq, k, dtype=torch.int64, device="cuda")
        for _ in range(3):
            ix.search(qs[0], nprobe, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tickets = []
        for i in range(40):
            tickets.append(ix.search_submit(qs[i % 8], nprobe, k, D, I))
            if len(tickets) > 3:
                ix.search_wait(tickets.pop(0))
        for t in tickets:
            ix.search_wait(t)
        e1.record()
        torch.cuda.synchronize()
        out[mirror] = e0.elapsed_time(e1) / 40
        print(f"timing n={n} nlist={nlist} mirror={mirror}: {out[mirror]:.3f} ms per batch, "
              f"index {ix.stats().gpu_memory_bytes / 2**30:.2f} GiB", flush=True)
        del ix
        torch.cuda.empty_cache()
    print(f"speed-up {out[False] / out[True]:.2f}x")


if __name__ == "__main__":
    ok = True
    ok &= check(768, 16, 9000, 8, 16, 10, pkg.Metric.L2)
    ok &= check(768, 64, 40000, 64, 16, 10, pkg.Metric.L2)
    ok &= check(128, 32, 50000, 40, 32, 25, pkg.Metric.InnerProduct)
    ok &= check(1024, 8, 6000, 9, 8, 5, pkg.Metric.L2)
    if ok and len(sys.argv) > 1:
        timing(int(sys.argv[1]))
    sys.exit(0 if ok else 1)
