"""Secondary measurements for the other BASELINE.json configs (parity-test cases, not the bench line).

    python tools/bench_configs.py c1 c2 c5 [--c5-ntrain N]

c1: IVF-Flat 100K x 128D, nlist 128, nprobe 16, k 10, batch 64 (the reference's gpu_vs_cpu_test shape),
    index is L2-resident (52 MB) -> latency bound; the reference's CPU path is timed beside it.
c2: exact brute force 1M x 768D, 1024 queries, k 100 (scan kernel over a flat view, exact fp32).
c5: k-means training (k-means++ + 10 Lloyd) + batched add() of 10M x 768D at nlist 16384.
"""
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")


def ev_time(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def c1():
    import oracle_lib as O
    n, dim, nlist, nprobe, k, nq = 100_000, 128, 128, 16, 10, 64
    x = O.gaussian(12345, n + nq, dim)
    db, q = x[:n], x[n:]
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    t = time.perf_counter(); ix.train(db[:10_000]); t_train = time.perf_counter() - t
    t = time.perf_counter(); ix.add(db); t_add = time.perf_counter() - t
    qd = torch.from_numpy(q).cuda()
    D = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    I = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ms = ev_time(lambda: ix.search_async(qd, nprobe, k, D, I, s), 200)
    t = time.perf_counter()
    for _ in range(50):
        ix.search(q, nprobe, k)
    e2e_ms = (time.perf_counter() - t) / 50 * 1e3
    out = {"config": "c1 IVF-Flat 100Kx128 nlist=128 nprobe=16 k=10 batch=64", "gpu_ms_per_batch": ms,
           "gpu_qps": nq / ms * 1e3, "e2e_ms_per_batch": e2e_ms, "e2e_qps": nq / e2e_ms * 1e3,
           "train_s": t_train, "add_s": t_add}
    if O.ref_lib() is not None:
        ref = O.RefIndex(dim, nlist)
        ref.centroids = ix.centroids
        ref.load_assigned(db, np.arange(n, dtype=np.uint64), ix.assign(db))
        ref.search(q[:4], nprobe, k)
        t = time.perf_counter(); ref.search(q, nprobe, k, 1); out["ref_cpu_1thread_qps"] = nq / (time.perf_counter() - t)
        nt = os.cpu_count()
        t = time.perf_counter(); ref.search(q, nprobe, k, nt); out[f"ref_cpu_{nt}threads_qps"] = nq / (time.perf_counter() - t)
    print(json.dumps(out), flush=True)


def c2():
    n, dim, nq, k = 1_000_000, 768, 1024, 100
    g = torch.Generator(device="cuda").manual_seed(7)
    db = torch.randn(n, dim, generator=g, device="cuda")
    q = torch.randn(nq, dim, generator=g, device="cuda")
    ms = ev_time(lambda: pkg.bruteforce_search(db, q, k), 3, warm=1)
    D, I = pkg.bruteforce_search(db, q, k)
    # property check of exactness on a few queries against a torch fp32 reference (ids up to ties)
    ref = torch.cdist(q[:8], db).pow(2).topk(k, largest=False)
    agree = float((ref.indices == I[:8]).float().mean())
    print(json.dumps({"config": "c2 brute force 1Mx768, 1024 queries, k=100", "ms_per_batch": ms,
                      "qps": nq / ms * 1e3, "tflops_equiv": 2.0 * n * nq * dim / ms / 1e9,
                      "topk_id_agreement_vs_torch_fp32": agree}), flush=True)


def c5(ntrain):
    n, dim, nlist = 10_000_000, 768, 16384
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    g = torch.Generator(device="cuda").manual_seed(12345)
    x = torch.randn(1_000_000, dim, generator=g, device="cuda")
    t = time.perf_counter(); ix.train(x[:ntrain]); t_train = time.perf_counter() - t
    t_add = 0.0
    for lo in range(0, n, 1_000_000):
        if lo:
            x = torch.randn(1_000_000, dim, generator=g, device="cuda")
        torch.cuda.synchronize()
        t = time.perf_counter(); ix.add(x); t_add += time.perf_counter() - t
    sizes = ix.list_sizes()
    # exactness spot check of the tensor-core assignment against a float64 torch argmin
    cent = torch.from_numpy(ix.centroids).cuda().double()
    xs = x[:20000]
    d2 = (cent * cent).sum(1)[None, :] - 2.0 * xs.double() @ cent.T
    agree = float((d2.argmin(1).int() == ix.assign_device(xs)).float().mean())
    print(json.dumps({"config": f"c5 train({ntrain}) + add(10M) 768D nlist=16384", "train_s": t_train,
                      "assign_agreement_vs_torch_f64": agree,
                      "add_s": t_add, "add_rows_per_s": n / t_add,
                      "add_assign_tflops": 2.0 * n * nlist * dim / t_add / 1e12,
                      "list_min_med_max": [int(sizes.min()), int(np.median(sizes)), int(sizes.max())]}), flush=True)


def assign():
    """nearest-centroid assignment alone (the contraction inside add() and every Lloyd iteration)"""
    n, dim, nlist = 1_000_000, 768, 16384
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, dim, generator=g, device="cuda")
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    ix.centroids = torch.randn(nlist, dim, generator=g, device="cuda").cpu().numpy()
    ms = ev_time(lambda: ix.assign_device(x), 3, warm=1)
    print(json.dumps({"config": "assign 1Mx768 vs 16384 centroids", "ms": ms,
                      "tflops_equiv": 2.0 * n * nlist * dim / ms / 1e9}), flush=True)


if __name__ == "__main__":
    args = sys.argv[1:] or ["c1", "c2"]
    ntrain = 262144
    if "--c5-ntrain" in args:
        ntrain = int(args[args.index("--c5-ntrain") + 1])
    if "c1" in args:
        c1()
    if "c2" in args:
        c2()
    if "assign" in args:
        assign()
    if "c5" in args:
        c5(ntrain)
