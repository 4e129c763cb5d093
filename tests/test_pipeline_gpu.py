"""GPU tests of the pipelined search (vdb_index_search_submit / wait): batches in flight on the index's own
streams must return exactly what the synchronous call returns, which is what the oracle returns; plus the
error paths the round-1 review asked for (query-batch chunking, bad assignments, the HBM budget)."""
import importlib
import os
import threading

import numpy as np
import pytest

import oracle_lib as O
from parity import check_search

pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed, n, dim, nlist, ntrain, nq, nprobe, k, metric = (int(v) for v in g["params"])
    x = O.gaussian(seed, n + nq, dim)
    return g, x[:n], x[n:], dict(dim=dim, nlist=nlist, ntrain=ntrain, nprobe=nprobe, k=k, metric=metric)


def golden_index(name, **kw):
    g, db, q, p = load_case(name)
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=p["dim"], nlist=p["nlist"], metric=pkg.Metric(p["metric"]), **kw))
    ix.centroids = g["centroids"]
    ix.add(db)
    return g, q, p, ix


@pytest.mark.parametrize("depth", [1, 2, 4, 8])
def test_pipelined_device_batches_match_the_golden_results(depth):
    import torch
    g, q, p, ix = golden_index("config1", pipeline_depth=depth)
    qd = torch.from_numpy(q).cuda()
    nq, k, bs = q.shape[0], p["k"], 16
    ix.reserve_search(bs, p["nprobe"], k)
    nb = 3 * depth + 2
    starts = [(7 * i) % (nq - bs) for i in range(nb)]
    D = [torch.empty((bs, k), dtype=torch.float32, device="cuda") for _ in range(nb)]
    I = [torch.empty((bs, k), dtype=torch.int64, device="cuda") for _ in range(nb)]
    torch.cuda.synchronize()
    tickets = [ix.search_submit(qd[lo:lo + bs], p["nprobe"], k, D[i], I[i]) for i, lo in enumerate(starts)]
    # device-side wait for the odd tickets, host wait for the even ones
    for i, t in enumerate(tickets):
        if i % 2:
            ix.search_wait_stream(t, torch.cuda.current_stream().cuda_stream)
        else:
            ix.search_wait(t)
    torch.cuda.synchronize()
    for i, lo in enumerate(starts):
        check_search(D[i].cpu().numpy(), I[i].cpu().numpy().view(np.uint64), g["D"][lo:lo + bs], g["I"][lo:lo + bs])
    # and bit-equal to the synchronous call
    Ds, Is = ix.search(qd, p["nprobe"], k)
    for i, lo in enumerate(starts):
        assert torch.equal(D[i], Ds[lo:lo + bs]) and torch.equal(I[i], Is[lo:lo + bs])


def test_pipelined_host_buffers_pageable_and_pinned():
    import torch
    g, q, p, ix = golden_index("ctest_gpu_vs_cpu")
    nq, k, bs = q.shape[0], p["k"], 20
    outs = []
    tickets = []
    for i, lo in enumerate(range(0, nq - bs + 1, bs)):
        if i % 2:  # pinned: copied straight by the stream
            qb = torch.from_numpy(q[lo:lo + bs].copy()).pin_memory()
            D = torch.empty((bs, k), dtype=torch.float32).pin_memory()
            I = torch.empty((bs, k), dtype=torch.int64).pin_memory()
        else:      # pageable numpy: staged through the slot's pinned buffers, delivered at wait()
            qb = np.ascontiguousarray(q[lo:lo + bs])
            D = np.full((bs, k), -1, np.float32)
            I = np.zeros((bs, k), np.uint64)
        tickets.append(ix.search_submit(qb, p["nprobe"], k, D, I))
        outs.append((lo, qb, D, I))
    for t in reversed(tickets):  # any order
        ix.search_wait(t)
    for lo, _, D, I in outs:
        Dn = D.numpy() if hasattr(D, "numpy") else D
        In = I.numpy().view(np.uint64) if hasattr(I, "numpy") else I
        check_search(Dn, In, g["D"][lo:lo + bs], g["I"][lo:lo + bs])


def test_unwaited_tickets_are_delivered_when_their_slot_is_recycled():
    g, q, p, ix = golden_index("ctest_gpu_vs_cpu", pipeline_depth=2)
    k, bs = p["k"], 10
    outs = []
    for lo in range(0, 60, bs):  # six submits over two slots, nobody waits in between
        D = np.full((bs, k), -1, np.float32)
        I = np.zeros((bs, k), np.uint64)
        t = ix.search_submit(np.ascontiguousarray(q[lo:lo + bs]), p["nprobe"], k, D, I)
        outs.append((lo, t, D, I))
    for lo, t, D, I in outs:
        ix.search_wait(t)
        check_search(D, I, g["D"][lo:lo + bs], g["I"][lo:lo + bs])


def test_many_host_threads_pipeline_through_one_index():
    g, q, p, ix = golden_index("ctest_gpu_vs_cpu")
    Dg, Ig = np.array(g["D"]), np.array(g["I"])
    errs = []

    def worker(lo):
        try:
            for rep in range(8):
                if rep % 2:
                    D, I = ix.search(q[lo:lo + 12], pkg.SearchParams(nprobe=p["nprobe"], k=p["k"]))
                else:
                    D = np.empty((12, p["k"]), np.float32)
                    I = np.empty((12, p["k"]), np.uint64)
                    ix.search_wait(ix.search_submit(np.ascontiguousarray(q[lo:lo + 12]), p["nprobe"], p["k"], D, I))
                check_search(D, I, Dg[lo:lo + 12], Ig[lo:lo + 12])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(lo,)) for lo in range(0, 96, 12)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_query_batch_too_large_for_one_pass_is_chunked():
    """nq * nprobe * k * 12 bytes of partial results above the 1 GiB cap: round 1 doubled pages-per-item until it
    wrapped to zero and divided by it (ADVICE, high).  Now the batch is split over passes."""
    dim, nlist, n, nq, nprobe, k = 16, 128, 20000, 3000, 128, 256
    x = O.gaussian(5, n + nq, dim)
    db, q = x[:n], x[n:]
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    ix.train(db[:2000])
    ix.add(db)
    D, I = ix.search(q, nprobe, k)
    assert (np.diff(D, axis=1) >= 0).all()
    # every slice searched on its own gives the same bits
    for lo in (0, 1500, 2730, 2990):
        Ds, Is = ix.search(q[lo:lo + 10], nprobe, k)
        assert np.array_equal(Ds, D[lo:lo + 10]) and np.array_equal(Is, I[lo:lo + 10])
    # nprobe = nlist: exhaustive, so the flat oracle is the truth
    Dr, Ir = O.flat_search(db, q[:8], k)
    check_search(D[:8], I[:8], Dr, Ir)
    # the stream-ordered form has no room to chunk: it must refuse, not crash
    import torch
    qd = torch.from_numpy(q).cuda()
    Dd = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    Id = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    with pytest.raises(ValueError):
        ix.search_async(qd, nprobe, k, Dd, Id, 0)


def test_add_assigned_rejects_assignments_outside_the_list_table():
    import torch
    dim, nlist, n = 32, 16, 1000
    x = torch.from_numpy(O.gaussian(9, n, dim)).cuda()
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    ix.train(x[:500])
    a = ix.assign_device(x)
    ids = torch.arange(n, dtype=torch.int64, device="cuda")
    bad = a.clone()
    bad[17] = nlist + 3
    with pytest.raises(ValueError):
        ix.add_assigned(x, ids, bad, n)
    assert ix.get_total_vectors() == 0 and int(ix.list_sizes().sum()) == 0
    ix.add_assigned(x, ids, a, n)  # the index is still usable
    assert int(ix.list_sizes().sum()) == n
    D, I = ix.search(x[:4], nlist, 1)
    assert I[:, 0].tolist() == [0, 1, 2, 3]


def test_max_gpu_memory_is_a_hard_budget():
    """IVFFlatIndex::Config::max_gpu_memory (ivf_flat_index.h:21): growth beyond it is OUT_OF_MEMORY, what was
    added before stays searchable."""
    import torch
    dim, nlist = 64, 8
    xh = O.gaussian(3, 300_000, dim)
    x = torch.from_numpy(xh).cuda()  # device rows are scattered in place: no staging buffer in the index's account
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist, max_gpu_memory=96 << 20))
    ix.train(x[:2000])
    ix.add(x[:50_000])  # 12.8 MB of rows: first 64 MiB slab
    with pytest.raises(MemoryError):
        ix.add(x[50_000:])  # 64 MB more needs a second slab: over budget
    assert ix.get_gpu_memory_usage() <= 96 << 20
    D, I = ix.search(xh[:5], nlist, 1)
    assert I[:, 0].tolist() == [0, 1, 2, 3, 4]


def test_float64_and_strided_queries_are_coerced_not_reinterpreted():
    g, q, p, ix = golden_index("simple_test")
    q64 = q.astype(np.float64)
    D, I = ix.search(q64, p["nprobe"], p["k"])
    check_search(D, I, g["D"], g["I"])
    wide = np.zeros((q.shape[0], 2 * q.shape[1]), np.float32)
    wide[:, ::2] = q
    D, I = ix.search(wide[:, ::2], p["nprobe"], p["k"])
    check_search(D, I, g["D"], g["I"])
    import torch
    with pytest.raises(ValueError):
        ix.search_async(torch.from_numpy(q64).cuda(), p["nprobe"], p["k"],
                        torch.empty((q.shape[0], p["k"]), device="cuda"),
                        torch.empty((q.shape[0], p["k"]), dtype=torch.int64, device="cuda"), 0)


def test_index_takes_its_list_memory_from_the_transfer_manager_pool():
    """the reference's index allocates through the TransferManager it is given (ivf_flat_index.cpp:424-433): with an
    arena attached the list slabs come out of the pool; what the pool cannot hold is allocated directly"""
    import ctypes as C
    l = pkg.lib()
    a = C.c_void_p()
    assert l.vdb_arena_create(0, 200 << 20, 1 << 20, 2, C.byref(a)) == 0
    st = (C.c_uint64 * 4)()
    g, db, q, p = load_case("config1")
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=p["dim"], nlist=p["nlist"]))
    pkg._check(l.vdb_index_set_arena(ix._h, a, 0))
    ix.centroids = g["centroids"]
    ix.add(db)  # 51 MB of rows: the first 64 MiB slab
    l.vdb_arena_stats(a, st)
    assert st[0] >= 64 << 20 and st[3] >= 1, list(st)
    D, I = ix.search(q, p["nprobe"], p["k"])
    check_search(D, I, g["D"], g["I"])
    big = O.gaussian(8, 400_000, p["dim"])  # 205 MB more: beyond the pool -> direct allocations, still one index
    ix.add(big, np.arange(400_000, dtype=np.uint64) + 10**7)
    assert ix.get_total_vectors() == db.shape[0] + 400_000
    D2, I2 = ix.search(big[:3], p["nlist"], 1)
    assert I2[:, 0].tolist() == [10**7, 10**7 + 1, 10**7 + 2] and np.all(D2[:, 0] == 0)
    ix.close()  # pooled slabs go back to the arena
    l.vdb_arena_stats(a, st)
    assert st[0] == 0 and st[3] == 0, list(st)
    assert l.vdb_arena_destroy(a) == 0


def test_transfer_callback_runs_behind_the_copy_without_blocking_the_caller():
    """TransferManager::enqueue_transfer with a callback (transfer_manager.cpp:218-261: cudaLaunchHostFunc)"""
    import ctypes as C
    import threading
    l = pkg.lib()
    a = C.c_void_p()
    assert l.vdb_arena_create(0, 64 << 20, 32 << 20, 2, C.byref(a)) == 0
    n = 4 << 20
    h = l.vdb_arena_allocate_pinned(a, n)
    d = l.vdb_arena_allocate_device(a, n)
    back = l.vdb_arena_allocate_pinned(a, n)
    C.memset(h, 7, n)
    done = threading.Event()
    seen = []
    CB = C.CFUNCTYPE(None, C.c_void_p)

    def fn(user):
        seen.append((C.c_ubyte * 4).from_address(back)[:])  # the D2H copy in front of the callback has landed
        done.set()

    cb = CB(fn)
    s = l.vdb_arena_get_stream(a)
    assert l.vdb_arena_enqueue_transfer(a, d, h, n, 1, s) == 0
    assert l.vdb_arena_enqueue_transfer_cb(a, back, d, n, 2, s, cb, None) == 0
    assert done.wait(30)
    assert seen == [[7, 7, 7, 7]]
    l.vdb_arena_return_stream(a, s)
    assert l.vdb_arena_destroy(a) == 0
