"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes
through the C ABI (libvdb_b200.so); the oracle is only the checker."""
import importlib
import os

import numpy as np
import pytest

import oracle_lib as O
from parity import check_search, check_probes, ip_scale, FLT_MAX, ID_PAD

pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["simple_test", "ctest_gpu_vs_cpu", "small_ip", "config1"]


def load_case(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed, n, dim, nlist, ntrain, nq, nprobe, k, metric = (int(v) for v in g["params"])
    x = O.gaussian(seed, n + nq, dim)
    return g, x[:n], x[n:], dict(dim=dim, nlist=nlist, ntrain=ntrain, nprobe=nprobe, k=k, metric=metric)


def new_index(dim, nlist, metric=0, **kw):
    return pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist, metric=pkg.Metric(metric), **kw))


def scale_for(metric, q, db):
    return ip_scale(q, db) if metric == O.METRIC_IP else None


@pytest.mark.parametrize("name", CASES)
def test_search_with_reference_centroids_matches_golden(name):
    """centroids imported from the reference, so coarse + add + scan + merge are isolated from training"""
    g, db, q, p = load_case(name)
    ix = new_index(p["dim"], p["nlist"], p["metric"])
    ix.centroids = g["centroids"]
    ix.add(db)
    assert np.array_equal(ix.list_sizes(), g["list_sizes"]), "assignment differs from the reference"
    assert ix.get_total_vectors() == db.shape[0]
    probes = ix.select_nprobe(q, p["nprobe"])
    # coarse order is (dist, list id); a swap is only legal between near-equal centroid distances
    check_probes(probes, g["probes"], q, g["centroids"], p["metric"])
    D, I = ix.search(q, pkg.SearchParams(nprobe=p["nprobe"], k=p["k"]))
    ties = check_search(D, I, g["D"], g["I"], scale=scale_for(p["metric"], q, db))
    assert ties <= 2


@pytest.mark.parametrize("name", CASES)
def test_train_is_bit_exact(name):
    """k-means++ (mt19937(42)) + 10 Lloyd iterations on the device == the reference's CPU train()"""
    g, db, q, p = load_case(name)
    ix = new_index(p["dim"], p["nlist"], p["metric"])
    ix.train(db[: p["ntrain"]])
    c = ix.centroids
    assert np.array_equal(c, g["centroids"]), f"max |diff| {np.abs(c - g['centroids']).max()}"


def test_full_pipeline_config1():
    """BASELINE.json configs[0]: train + add + search all on the GPU vs the reference's outputs"""
    g, db, q, p = load_case("config1")
    ix = new_index(p["dim"], p["nlist"], p["metric"])
    ix.train(db[: p["ntrain"]])
    ix.add(db, np.arange(db.shape[0], dtype=np.uint64))
    D, I = ix.search(q, pkg.SearchParams(nprobe=p["nprobe"], k=p["k"]))
    check_search(D, I, g["D"], g["I"])
    st = ix.last_search_stats()
    sizes = g["list_sizes"].astype(np.int64)
    assert st.algorithmic_rows == int(sizes[g["probes"]].sum())
    assert st.unique_rows == int(sizes[np.unique(g["probes"])].sum())
    assert ix.get_gpu_memory_usage() > db.nbytes


@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
@pytest.mark.parametrize("dim", [3, 17, 100, 260])
def test_odd_dimensions_and_multiple_adds(metric, dim):
    rng = np.random.default_rng(dim)
    n, nlist, nq = 3000, 12, 20
    x = O.gaussian(100 + dim, n + nq, dim)
    db, q = x[:n], x[n:]
    ora = O.OracleIndex(dim, nlist, metric)
    ora.train(db[:1500])
    ix = new_index(dim, nlist, metric)
    ix.train(db[:1500])
    assert np.array_equal(ix.centroids, ora.centroids)
    ids = (rng.permutation(n) + 10_000_000_000).astype(np.uint64)  # ids beyond 32 bits
    for lo in range(0, n, 700):
        ora.add(db[lo:lo + 700], ids[lo:lo + 700])
        ix.add(db[lo:lo + 700], ids[lo:lo + 700])
    assert np.array_equal(ix.list_sizes(), ora.list_sizes())
    for l in range(nlist):
        assert np.array_equal(np.sort(ix.list_ids(l)), np.sort(ora.list_ids(l)))
    assert np.array_equal(ix.assign(q), ora.assign(q))
    for nprobe, k in [(1, 1), (4, 10), (nlist, 33), (nlist + 7, 100)]:
        Dr, Ir = ora.search(q, nprobe, k)
        D, I = ix.search(q, nprobe, k)
        check_search(D, I, Dr, Ir, scale=scale_for(metric, q, db))


def test_edge_cases():
    dim, nlist = 8, 6
    x = O.gaussian(3, 40, dim)
    cent = O.gaussian(4, nlist, dim)
    ora = O.OracleIndex(dim, nlist)
    ora.centroids = cent
    ix = new_index(dim, nlist)
    ix.centroids = cent
    # empty index: everything padded (ivf_flat_index.cpp:514-517)
    D, I = ix.search(x[:2], 3, 4)
    assert (I == ID_PAD).all() and (D == FLT_MAX).all()
    # duplicate ids, inside one list and across lists: merge_results de-duplicates (:496-504)
    ids = np.arange(40, dtype=np.uint64) % 13
    ora.add(x, ids)
    ix.add(x, ids)
    for nprobe, k in [(6, 5), (6, 40), (2, 3), (1, 1)]:
        Dr, Ir = ora.search(x[:7], nprobe, k)
        D, I = ix.search(x[:7], nprobe, k)
        check_search(D, I, Dr, Ir)
    with pytest.raises(ValueError):
        ix.search(x[:1], 3, 0)
    with pytest.raises(ValueError):
        ix.search(x[:1], 3, 5000)


def test_large_k_and_skewed_lists():
    """k = 1000 (the server's topk cap, query_service.cpp:77) on lists spanning several pages"""
    dim, nlist, n, nq = 32, 4, 9000, 6
    x = O.clustered(11, n + nq, dim, n_centers=3, spread=0.3)
    db, q = x[:n], x[n:]
    ora = O.OracleIndex(dim, nlist)
    ora.train(db[:2000])
    ix = new_index(dim, nlist, page_rows=256)
    ix.centroids = ora.centroids
    ora.add(db)
    ix.add(db)
    assert np.array_equal(ix.list_sizes(), ora.list_sizes())
    for nprobe, k in [(2, 1000), (4, 300), (4, 2048)]:
        Dr, Ir = ora.search(q, nprobe, k)
        D, I = ix.search(q, nprobe, k)
        check_search(D, I, Dr, Ir)


def test_clustered_small_distances():
    """distances much smaller than the vector norms: the exact (q-v)^2 form must keep 1e-5"""
    dim, nlist, n, nq = 64, 16, 20000, 32
    x = O.clustered(5, n + nq, dim, n_centers=40, spread=0.01)
    db, q = x[:n], x[n:]
    ora = O.OracleIndex(dim, nlist)
    ora.train(db[:4000])
    ix = new_index(dim, nlist)
    ix.centroids = ora.centroids
    ora.add(db)
    ix.add(db)
    Dr, Ir = ora.search(q, 4, 10)
    D, I = ix.search(q, 4, 10)
    check_search(D, I, Dr, Ir)


@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
def test_bruteforce_matches_flat_oracle(metric):
    """launch_bruteforce_search replacement: exact top-k, k = 100 > the reference kernel's cap of 32"""
    n, dim, nq, k = 10007, 96, 37, 100
    x = O.gaussian(21, n + nq, dim)
    db, q = x[:n], x[n:]
    ids = (np.arange(n, dtype=np.uint64) * 3 + 7)
    Dr, Ir = O.flat_search(db, q, k, metric, ids)
    D, I = pkg.bruteforce_search(db, q, k, pkg.Metric(metric), ids)
    check_search(D, I, Dr, Ir, scale=scale_for(metric, q, db))
    Dr, Ir = O.flat_search(db, q[:5], 7, metric)
    D, I = pkg.bruteforce_search(db, q[:5], 7, pkg.Metric(metric))
    check_search(D, I, Dr, Ir, scale=scale_for(metric, q[:5], db))


def test_bruteforce_768d_many_queries():
    """configs[1] shape at reduced N: 768-D, 256 queries, k = 100"""
    n, dim, nq, k = 20000, 768, 256, 100
    x = O.gaussian(33, n + nq, dim)
    db, q = x[:n], x[n:]
    Dr, Ir = O.flat_search(db, q[:24], k)
    D, I = pkg.bruteforce_search(db, q, k)
    check_search(D[:24], I[:24], Dr, Ir)
    # every row sorted by (distance, id)
    assert (np.diff(D, axis=1) >= 0).all()


@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
@pytest.mark.parametrize("n,dim,nq,k", [(70001, 96, 37, 100), (131072, 100, 130, 10), (90000, 768, 64, 33),
                                        (65536, 64, 16, 1000), (300000, 32, 257, 1)])
def test_bruteforce_tensor_path_matches_flat_oracle(metric, n, dim, nq, k):
    """n >= 65536 and >= 16 queries: TF32 contraction over nested samples + exact fp32 re-scoring (bruteforce_tc.cu)
    must return what the exact scan returns -- ids bit-exact up to ties, distances to 1e-5"""
    x = O.gaussian(77 + dim, n + nq, dim)
    db, q = x[:n], x[n:]
    ids = (np.arange(n, dtype=np.uint64) * 5 + 11) if dim == 96 else None
    nchk = min(nq, 24)
    Dr, Ir = O.flat_search(db, q[:nchk], k, metric, ids)
    D, I = pkg.bruteforce_search(db, q, k, pkg.Metric(metric), ids)
    check_search(D[:nchk], I[:nchk], Dr, Ir, scale=scale_for(metric, q[:nchk], db))
    assert (np.diff(D, axis=1) >= 0).all()
    # the small-batch route (< 16 queries) is the exact scan: both routes must agree on every query they share
    D2, I2 = pkg.bruteforce_search(db, q[:8], k, pkg.Metric(metric), ids)
    check_search(D[:8], I[:8], D2, I2, scale=scale_for(metric, q[:8], db))


def test_bruteforce_tensor_path_heavy_ties_falls_back():
    """every row identical: the candidate lists overflow and the call must still answer exactly (scan fallback)"""
    import torch
    n, dim, nq, k = 70000, 64, 32, 10
    db = torch.ones(n, dim, device="cuda")
    db[12345] = 0.5
    q = torch.full((nq, dim), 0.5, device="cuda")
    D, I = pkg.bruteforce_search(db, q, k)
    assert (I[:, 0] == 12345).all() and (D[:, 0] == 0).all()
    assert torch.equal(I[:, 1:].cpu(), torch.arange(0, k - 1).expand(nq, k - 1))
    assert torch.allclose(D[:, 1:], torch.full((nq, k - 1), 16.0, device="cuda"))


def test_bruteforce_tensor_path_sorted_database():
    """database ordered by cluster: the strided samples still see every cluster"""
    import torch
    n, dim, nq, k = 100000, 128, 48, 20
    gen = torch.Generator(device="cuda").manual_seed(5)
    centers = torch.randn(50, dim, generator=gen, device="cuda") * 4
    lab = torch.arange(n, device="cuda") // (n // 50)
    db = centers[lab] + torch.randn(n, dim, generator=gen, device="cuda")
    q = centers[torch.arange(nq, device="cuda") % 50] + torch.randn(nq, dim, generator=gen, device="cuda")
    D, I = pkg.bruteforce_search(db, q, k)
    ref = torch.cdist(q.double(), db.double()).pow(2).topk(k, largest=False)
    assert torch.equal(ref.indices, I.long())
    assert torch.allclose(ref.values.float(), D, rtol=1e-5)


def test_kmeans_assign_entry_point():
    n, dim, nc = 5000, 40, 50
    x = O.gaussian(8, n + nc, dim)
    v, c = x[:n], x[n:]
    for metric in (O.METRIC_L2, O.METRIC_IP):
        ora = O.OracleIndex(dim, nc, metric)
        ora.centroids = c
        a, d = pkg.kmeans_assign(v, c, pkg.Metric(metric), want_distances=True)
        assert np.array_equal(a, ora.assign(v))
        assert np.isfinite(d).all()


def test_ivf_full_probe_equals_bruteforce_768d():
    """size-independent property: nprobe = nlist makes IVF exact; 768-D rows, lists over many pages"""
    import torch
    n, dim, nlist, nq, k = 60000, 768, 24, 64, 10
    gen = torch.Generator(device="cuda").manual_seed(1)
    db = torch.randn(n, dim, generator=gen, device="cuda")
    q = torch.randn(nq, dim, generator=gen, device="cuda")
    ix = new_index(dim, nlist)
    ix.train(db[:3000])
    ix.add(db)  # device pointers straight in, ids default to the row number
    assert int(ix.list_sizes().sum()) == n
    D, I = ix.search(q, nlist, k)
    Db, Ib = pkg.bruteforce_search(db, q, k)
    assert torch.equal(I, Ib)
    assert torch.allclose(D, Db, rtol=1e-6, atol=0)
    # oracle spot check on 4 queries
    Dr, Ir = O.flat_search(db.cpu().numpy(), q[:4].cpu().numpy(), k)
    check_search(D[:4].cpu().numpy(), I[:4].cpu().numpy(), Dr, Ir)
    # fewer probes: a subset of the candidates, so distances can only grow
    D8, _ = ix.search(q, 8, k)
    assert (D8 >= D - 1e-3).all()


@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
@pytest.mark.parametrize("dim,nlist,nq,nprobe", [(768, 1024, 70, 32), (100, 300, 9, 64), (128, 4096, 64, 16)])
def test_tensor_core_coarse_matches_oracle(metric, dim, nlist, nq, nprobe):
    """tcgen05 TF32 contraction + fp32 re-check == select_nprobe_lists; must not depend on tensor rounding"""
    x = O.gaussian(1000 + dim, nlist + nq, dim)
    cent, q = x[:nlist] * np.linspace(0.2, 1.5, nlist, dtype=np.float32)[:, None], x[nlist:]
    ora = O.OracleIndex(dim, nlist, metric)
    ora.centroids = cent
    ref = np.stack([ora.select_nprobe(q[i], nprobe) for i in range(nq)])
    got = {}
    for mode in (1, 2):  # SIMT scan, tensor cores
        ix = new_index(dim, nlist, metric, coarse_mode=mode)
        ix.centroids = cent
        got[mode] = ix.select_nprobe(q, nprobe)
        ties = check_probes(got[mode], ref, q, cent, metric)
        assert ties <= nq // 8 + 2, f"mode {mode}: {ties} tie-explained differences"


def test_search_same_under_both_coarse_modes():
    g, db, q, p = load_case("config1")
    out = {}
    for mode in (1, 2):
        ix = new_index(p["dim"], p["nlist"], p["metric"], coarse_mode=mode)
        ix.centroids = g["centroids"]
        ix.add(db)
        out[mode] = ix.search(q, pkg.SearchParams(nprobe=p["nprobe"], k=p["k"]))
        check_search(out[mode][0], out[mode][1], g["D"], g["I"])


@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
@pytest.mark.parametrize("dim,nlist,n", [(96, 300, 5000), (768, 512, 3000), (50, 1000, 2500)])
def test_tensor_core_assignment_is_bit_exact(metric, dim, nlist, n):
    """tcgen05 GEMM + bounds + reference-order re-check == assign_to_lists (nlist >= 256 takes this path)"""
    x = O.gaussian(500 + dim, n + nlist, dim)
    v, cent = x[:n], x[n:] * np.linspace(0.3, 1.2, nlist, dtype=np.float32)[:, None]
    ora = O.OracleIndex(dim, nlist, metric)
    ora.centroids = cent
    ref = ora.assign(v)
    for mode in (pkg.TrainMode.AUTO, pkg.TrainMode.EXACT):  # tensor cores, scalar kernel
        ix = new_index(dim, nlist, metric, train_mode=mode)
        ix.centroids = cent
        assert np.array_equal(ix.assign(v), ref), f"train_mode {mode}"
    # near-duplicate centroids: many candidates per row, ties resolved by the lowest index
    cent2 = np.repeat(cent[: nlist // 4], 4, axis=0)
    cent2[1::4] += 1e-7
    ora.centroids = cent2
    ix = new_index(dim, nlist, metric)
    ix.centroids = cent2
    assert np.array_equal(ix.assign(v[:600]), ora.assign(v[:600]))


def test_train_bit_exact_with_tensor_core_assignment():
    dim, nlist, n = 32, 256, 4096
    x = O.gaussian(77, n, dim)
    ora = O.OracleIndex(dim, nlist)
    ora.train(x)
    ix = new_index(dim, nlist)
    ix.train(x)
    assert np.array_equal(ix.centroids, ora.centroids)
    ora.add(x)
    ix.add(x)
    assert np.array_equal(ix.list_sizes(), ora.list_sizes())


def test_merge_topk_entry_point():
    import torch
    parts, nq, k = 4, 9, 10
    rng = np.random.default_rng(0)
    d = np.sort(rng.random((parts, nq, k)).astype(np.float32), axis=2)
    i = rng.integers(0, 50, (parts, nq, k)).astype(np.uint64)  # duplicates across parts on purpose
    d[1, :, 7:] = FLT_MAX
    i[1, :, 7:] = ID_PAD
    D, I = pkg.merge_topk(torch.from_numpy(d).cuda(), torch.from_numpy(i.view(np.int64)).cuda())
    D, I = D.cpu().numpy(), I.cpu().numpy().view(np.uint64)
    for q in range(nq):
        cand = sorted((float(dd), int(ii)) for p in range(parts) for dd, ii in zip(d[p, q], i[p, q]) if ii != ID_PAD)
        seen, out = set(), []
        for dd, ii in cand:
            if ii not in seen:
                seen.add(ii)
                out.append((dd, ii))
        out = out[:k]
        assert [int(v) for v in I[q, :len(out)]] == [o[1] for o in out]
        assert np.allclose(D[q, :len(out)], [o[0] for o in out])


def test_arena():
    import ctypes as C
    l = pkg.lib()
    a = C.c_void_p()
    assert l.vdb_arena_create(0, 64 << 20, 16 << 20, 4, C.byref(a)) == 0
    p1 = l.vdb_arena_allocate_device(a, 1000)
    p2 = l.vdb_arena_allocate_device(a, 1 << 20)
    h = l.vdb_arena_allocate_pinned(a, 4096)
    assert p1 and p2 and h and p1 % 256 == 0 and p2 % 256 == 0
    assert l.vdb_arena_allocate_device(a, 1 << 30) is None  # exhausted pool -> nullptr
    src = (C.c_float * 256)(*range(256))
    C.memmove(h, src, 1024)
    s = l.vdb_arena_get_stream(a)
    assert l.vdb_arena_enqueue_transfer(a, p1, h, 1024, 1, s) == 0
    back = l.vdb_arena_allocate_pinned(a, 4096)
    assert l.vdb_arena_enqueue_transfer(a, back, p1, 1024, 2, s) == 0
    assert l.vdb_arena_synchronize_stream(a, s) == 0
    assert l.vdb_arena_return_stream(a, s) == 0
    out = (C.c_float * 256).from_address(back)
    assert list(out) == list(range(256))
    st = (C.c_uint64 * 4)()
    l.vdb_arena_stats(a, st)
    assert st[0] == 1024 + (1 << 20) and st[3] == 4
    assert l.vdb_arena_free_device(a, p1) == 0 and l.vdb_arena_free_device(a, p2) == 0
    assert l.vdb_arena_free_device(a, p2) != 0  # double free is reported
    p3 = l.vdb_arena_allocate_device(a, 60 << 20)  # coalesced back into one block
    assert p3
    assert l.vdb_arena_destroy(a) == 0


def test_cpp_host_mirror_simple_test():
    """the C++ vdb::IVFFlatIndex mirror (host/ivf_flat_index.h) driven like test/simple_test.cpp"""
    import subprocess
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "host", "simple_test")
    if not os.path.exists(exe):
        pkg.build()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "PASSED" in out.stdout, out.stdout + out.stderr
    g = np.load(os.path.join(GOLD, "simple_test.npz"))
    rows = [l.split() for l in out.stdout.splitlines() if l.startswith("R ")]
    I = np.array([int(r[2]) for r in rows], np.uint64).reshape(g["I"].shape)
    D = np.array([float(r[3]) for r in rows], np.float32).reshape(g["D"].shape)
    check_search(D, I, g["D"], g["I"])


def test_concurrent_searches_from_host_threads():
    """the gRPC server calls Search from several poller threads on one index (server/main.cpp:92-94)"""
    import threading
    g, db, q, p = load_case("ctest_gpu_vs_cpu")
    ix = new_index(p["dim"], p["nlist"], p["metric"])
    ix.centroids = g["centroids"]
    ix.add(db)
    errs = []
    Dg, Ig = np.array(g["D"]), np.array(g["I"])  # NpzFile members are not safe to read from several threads

    def worker(lo):
        try:
            for _ in range(5):
                D, I = ix.search(q[lo:lo + 25], pkg.SearchParams(nprobe=p["nprobe"], k=p["k"]))
                check_search(D, I, Dg[lo:lo + 25], Ig[lo:lo + 25])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(lo,)) for lo in (0, 25, 50, 75)]
    [t_.start() for t_ in th]
    [t_.join() for t_ in th]
    assert not errs, errs


def test_async_searches_on_different_streams_do_not_share_a_workspace_in_time():
    """search_async from alternating streams without any host synchronisation: the index orders the searches on
    the device (one workspace), so every result must equal the single-stream answer"""
    import torch
    g, db, q, p = load_case("config1")
    ix = new_index(p["dim"], p["nlist"], p["metric"])
    ix.centroids = g["centroids"]
    ix.add(db)
    qd = torch.from_numpy(q).cuda()
    nq, k = qd.shape[0], p["k"]
    Dref, Iref = ix.search(qd, p["nprobe"], k)
    streams = [torch.cuda.Stream() for _ in range(3)]
    outs = []
    torch.cuda.synchronize()
    for rep in range(12):
        s = streams[rep % 3]
        lo = (rep * 5) % (nq - 16)
        D = torch.empty((16, k), dtype=torch.float32, device="cuda")
        I = torch.empty((16, k), dtype=torch.int64, device="cuda")
        ix.search_async(qd[lo:lo + 16], p["nprobe"], k, D, I, s.cuda_stream)
        outs.append((lo, D, I))
    torch.cuda.synchronize()
    for lo, D, I in outs:
        assert torch.equal(I, Iref[lo:lo + 16]) and torch.equal(D, Dref[lo:lo + 16])


def test_single_query_and_k1_and_nlist1():
    x = O.gaussian(2, 3000, 20)
    ora = O.OracleIndex(20, 1)
    ora.train(x[:50])
    ora.add(x[:2900])
    ix = new_index(20, 1)
    ix.train(x[:50])
    ix.add(x[:2900])
    for nq, k in [(1, 1), (1, 17), (100, 1)]:
        Dr, Ir = ora.search(x[2900:2900 + nq], 1, k)
        D, I = ix.search(x[2900:2900 + nq], 1, k)
        check_search(D, I, Dr, Ir)


@pytest.mark.parametrize("case", ["gaussian768", "integer_ties", "subnormal", "duplicates", "tiny_n"])
def test_parallel_exact_sampler_matches_sequential(case):
    """AUTO evaluates the reference's sequential fp32 D^2-sampling sums with a block-wide scan of parity-dependent
    integer maps (seed_sample_par_kernel); EXACT runs the literal one-lane chain.  One differing pick would change
    every later centroid, so equal centroids == equal sums and picks, bit for bit."""
    rng = np.random.default_rng(17)
    if case == "gaussian768":
        n, dim, nlist = 100_000, 768, 40
        x = O.gaussian(91, n, dim)
    elif case == "integer_ties":  # integer distances in the thousands: running sums beyond 2^24 round with exact ties
        n, dim, nlist = 120_000, 64, 48
        x = rng.integers(0, 16, (n, dim)).astype(np.float32)
    elif case == "subnormal":  # squared distances around 1e-41: subnormal terms and sums
        n, dim, nlist = 50_000, 16, 32
        x = (rng.standard_normal((n, dim)) * 1e-21).astype(np.float32)
    elif case == "duplicates":  # many zero terms, a few huge ones (binade jumps of many bits at once)
        n, dim, nlist = 70_000, 8, 24
        x = np.repeat(rng.standard_normal((700, dim)).astype(np.float32), 100, axis=0)
        x[::997] *= 1e6
    else:
        n, dim, nlist = 37, 5, 9
        x = O.gaussian(3, n, dim)
    a = new_index(dim, nlist)  # AUTO
    a.train(x)
    b = new_index(dim, nlist, train_mode=pkg.TrainMode.EXACT)
    b.train(x)
    assert np.array_equal(a.centroids, b.centroids), f"max |diff| {np.abs(a.centroids - b.centroids).max()}"


def test_fast_train_mode_is_statistically_equivalent():
    """FAST k-means++ sampling (parallel sums): same RNG stream, nearly always the same seeds; the clustering
    quality (inertia) must match the reference's within 1 %"""
    dim, nlist, n = 24, 64, 6000
    x = O.clustered(9, n, dim, n_centers=20, spread=0.5)
    ora = O.OracleIndex(dim, nlist)
    ora.train(x)
    ix = new_index(dim, nlist, train_mode=pkg.TrainMode.FAST)
    ix.train(x)

    def inertia(c):
        d = ((x[:, None, :].astype(np.float64) - c[None, :, :]) ** 2).sum(-1)
        return d.min(1).sum()

    a, b = inertia(ix.centroids), inertia(ora.centroids)
    assert abs(a - b) <= 0.01 * b, (a, b)


def test_kmeans_accumulate_and_finalize_building_blocks():
    """the data-parallel Lloyd pieces of the ABI: per-cluster sums in input order, counts, division"""
    import ctypes as C
    import torch
    n, dim, nc = 3000, 20, 17
    x = O.gaussian(4, n, dim)
    rng = np.random.default_rng(0)
    a = rng.integers(0, nc - 2, n).astype(np.uint32)  # the last two clusters stay empty
    xd, ad = torch.from_numpy(x).cuda(), torch.from_numpy(a.astype(np.int32)).cuda()
    sums = torch.zeros((nc, dim), dtype=torch.float32, device="cuda")
    counts = torch.zeros(nc, dtype=torch.int32, device="cuda")
    l = pkg.lib()
    assert l.vdb_kmeans_accumulate(xd.data_ptr(), ad.data_ptr(), n, nc, dim, sums.data_ptr(), counts.data_ptr(), None) == 0
    ref_s = np.zeros((nc, dim), np.float32)
    for v in range(n):  # fp32, input order, like ivf_flat_index.cpp:123-131
        ref_s[a[v]] += x[v]
    assert np.array_equal(sums.cpu().numpy(), ref_s)
    assert np.array_equal(counts.cpu().numpy(), np.bincount(a, minlength=nc))
    cent = torch.full((nc, dim), 7.0, dtype=torch.float32, device="cuda")
    assert l.vdb_kmeans_finalize(sums.data_ptr(), counts.data_ptr(), cent.data_ptr(), nc, dim, None) == 0
    torch.cuda.synchronize()
    c = cent.cpu().numpy()
    cnt = np.bincount(a, minlength=nc)
    assert np.array_equal(c[cnt > 0], ref_s[cnt > 0] / cnt[cnt > 0, None].astype(np.float32))
    assert (c[cnt == 0] == 7.0).all()  # an empty cluster keeps its centroid (:134-141)


@pytest.mark.parametrize("dim", [200, 256, 512, 1000, 1024, 1536, 2048])
def test_every_row_width_instantiation(dim):
    """one case per register-tile instantiation of the scan kernel (NJ = 2, 4, 8, 12, 16; full and masked widths)"""
    n, nq, k, nlist = 1500, 11, 7, 5
    x = O.gaussian(900 + dim, n + nq, dim)
    db, q = x[:n], x[n:]
    Dr, Ir = O.flat_search(db, q, k)
    D, I = pkg.bruteforce_search(db, q, k)
    check_search(D, I, Dr, Ir)
    ora = O.OracleIndex(dim, nlist)
    ora.train(db[:300])
    ora.add(db)
    ix = new_index(dim, nlist)
    ix.centroids = ora.centroids
    ix.add(db)
    Dr, Ir = ora.search(q, 3, k)
    D, I = ix.search(q, 3, k)
    check_search(D, I, Dr, Ir)


def test_large_query_batch_and_k():
    """2000 queries x k = 100 in one call (many query tiles per list, large partial-result buffers)"""
    dim, nlist, n, nq, k, nprobe = 48, 64, 40000, 2000, 100, 8
    x = O.gaussian(71, n + nq, dim)
    db, q = x[:n], x[n:]
    ora = O.OracleIndex(dim, nlist)
    ora.train(db[:4000])
    ora.add(db)
    ix = new_index(dim, nlist)
    ix.centroids = ora.centroids
    ix.add(db)
    Dr, Ir = ora.search(q, nprobe, k, nthreads=8)
    D, I = ix.search(q, nprobe, k)
    check_search(D, I, Dr, Ir)


def test_nprobe_beyond_the_topk_machinery():
    """nprobe > 2048 (the select kernels' pool): full sort of the centroid table per query"""
    dim, nlist, n, nq, k = 16, 3000, 30000, 9, 10
    x = O.gaussian(55, n + nq, dim)
    db, q = x[:n], x[n:]
    cent = db[:nlist].copy()
    ora = O.OracleIndex(dim, nlist)
    ora.centroids = cent
    ora.add(db)
    ix = new_index(dim, nlist)
    ix.centroids = cent
    ix.add(db)
    for nprobe in (2049, 2500, 3000, 5000):
        Dr, Ir = ora.search(q, nprobe, k)
        D, I = ix.search(q, nprobe, k)
        check_search(D, I, Dr, Ir)
        ref = np.stack([ora.select_nprobe(q[i], nprobe) for i in range(nq)])
        check_probes(ix.select_nprobe(q, nprobe), ref, q, cent, 0)


def test_full_size_config2_bruteforce_against_float64():
    """BASELINE.json configs[1] at FULL size (1M x 768, 1024 queries, k = 100): every result row sorted by
    (distance, id); 32 of the queries checked against a float64 evaluation of all 1M distances (ids equal except
    ties inside the fp32 tolerance); the result does not depend on how the query batch is split (the tensor path
    takes >= 16 queries, the scan path fewer)."""
    import torch
    n, dim, nq, k = 1_000_000, 768, 1024, 100
    gen = torch.Generator(device="cuda").manual_seed(7)
    db = torch.randn(n, dim, generator=gen, device="cuda")
    q = torch.randn(nq, dim, generator=gen, device="cuda")
    D, I = pkg.bruteforce_search(db, q, k)
    assert (D[:, 1:] >= D[:, :-1]).all() and (I >= 0).all() and (I < n).all()
    same_d = D[:, 1:] == D[:, :-1]
    assert (I[:, 1:][same_d] > I[:, :-1][same_d]).all()
    sub = torch.arange(0, nq, 32, device="cuda")
    q64, d64 = q[sub].double(), None
    d64 = (q64 * q64).sum(1)[:, None] - 2.0 * q64 @ db.double().T
    vn = torch.empty(n, dtype=torch.float64, device="cuda")
    for lo in range(0, n, 100_000):
        vn[lo:lo + 100_000] = (db[lo:lo + 100_000].double() ** 2).sum(1)
    d64 += vn[None, :]
    ref = d64.topk(k, dim=1, largest=False)
    check_search(D[sub].cpu().numpy(), I[sub].cpu().numpy(), ref.values.float().cpu().numpy(),
                 ref.indices.cpu().numpy().astype(np.uint64))
    del d64, vn
    # the reference's own arithmetic on a 64-query subset -- the oracle as a flat index (nlist = 1,
    # nprobe = 1: search_list_cpu over all 1M rows, strict left-to-right fp32), every host core on its own queries
    sub64 = np.arange(0, nq, 16)
    dbh = db.cpu().numpy()
    ora = O.OracleIndex(dim, 1)
    ora.centroids = np.zeros((1, dim), np.float32)
    ora.load_assigned(dbh, np.arange(n, dtype=np.uint64), np.zeros(n, np.uint32))
    Dr, Ir = ora.search(q[torch.from_numpy(sub64).cuda()].cpu().numpy(), 1, k, os.cpu_count() or 1)
    check_search(D[sub64].cpu().numpy(), I[sub64].cpu().numpy(), Dr, Ir)
    ora.close()
    del dbh
    D8, I8 = pkg.bruteforce_search(db, q[:8], k)   # < 16 queries: the exact scan kernel
    check_search(D[:8].cpu().numpy(), I[:8].cpu().numpy(), D8.cpu().numpy(), I8.cpu().numpy())


def test_full_size_config4_training_is_bit_identical_to_the_literal_restatement():
    """BASELINE.json configs[4] training at FULL size (262 144 x 768 sample, nlist 16384): the production path
    (parallel exact k-means++ sums, TMA distance update, tensor-core assignment with reference-order re-check) must
    produce the same centroids, bit for bit, as train_mode = EXACT (one-lane sequential sums, scalar fp32 assignment
    kernel) -- the configuration that is itself pinned to the reference on the golden cases."""
    import torch
    dim, nlist, n = 768, 16384, 262144
    gen = torch.Generator(device="cuda").manual_seed(12345)
    x = torch.randn(n, dim, generator=gen, device="cuda")
    a = new_index(dim, nlist)
    a.train(x)
    b = new_index(dim, nlist, train_mode=pkg.TrainMode.EXACT)
    b.train(x)
    ca, cb = a.centroids, b.centroids
    assert np.array_equal(ca, cb), f"{int((ca != cb).any(1).sum())} of {nlist} centroids differ"
    # and add() on top: tensor-core assignment == scalar assignment on fresh rows
    y = torch.randn(200_000, dim, generator=gen, device="cuda")
    assert torch.equal(a.assign_device(y), b.assign_device(y))


def test_full_size_config3_ivf_equals_bruteforce_and_is_monotone():
    """BASELINE.json configs[2] at FULL size (10M x 768, nlist 4096) through size-independent properties: with
    nprobe = nlist the IVF path (coarse + grouped list scan + merge over 34 GB of pages) must return exactly what
    the independent tensor-core brute-force path returns on the same rows; fewer probes can only raise distances;
    every id sits in the list its row was assigned to; the lists hold every row exactly once."""
    import torch
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 90 * 2**30:
        pytest.skip("needs ~70 GB of HBM")
    n, dim, nlist, nq, k = 10_000_000, 768, 4096, 32, 10
    gen = torch.Generator(device="cuda").manual_seed(12345)
    db = torch.empty((n, dim), dtype=torch.float32, device="cuda")
    for lo in range(0, n, 1_000_000):
        db[lo:lo + 1_000_000].normal_(generator=gen)
    q = torch.randn(nq, dim, generator=gen, device="cuda")
    ix = new_index(dim, nlist)
    ix.train(db[:262144])
    for lo in range(0, n, 1_000_000):
        ix.add(db[lo:lo + 1_000_000], torch.arange(lo, lo + 1_000_000, dtype=torch.int64, device="cuda"))
    sizes = ix.list_sizes().astype(np.int64)
    assert int(sizes.sum()) == n and ix.get_total_vectors() == n
    Dall, Iall = ix.search(q, nlist, k)            # exhaustive IVF
    Db, Ib = pkg.bruteforce_search(db, q, k)       # tensor-core brute force over the flat copy
    assert torch.equal(Iall, Ib)
    assert torch.allclose(Dall, Db, rtol=1e-6, atol=0)
    D32, I32 = ix.search(q, 32, k)                 # the headline setting: a subset of the candidates
    assert (D32 >= Dall * (1 - 1e-6)).all()
    recall = float((I32[:, :, None] == Iall[:, None, :]).any(-1).float().mean())
    assert recall > 0.03, recall                   # pure noise has no cluster structure: low (~0.1), chance is 32/4096
    # every returned id is a row of one of the probed lists, at the distance the flat copy gives
    probes = torch.from_numpy(ix.select_nprobe(q.cpu().numpy(), 32).astype(np.int64)).cuda()
    owner = ix.assign_device(db[I32.reshape(-1)]).long().reshape(nq, k)
    assert (owner[:, :, None] == probes[:, None, :]).any(-1).all()
    d_chk = ((db[I32.reshape(-1)].reshape(nq, k, dim) - q[:, None, :]) ** 2).sum(-1)
    assert torch.allclose(d_chk, D32, rtol=1e-5)
    # COMPLETENESS of the headline setting, from an independent truth: assign all 10M rows, keep the rows whose list
    # one of the query's 32 probes names, evaluate their distances in float64 with torch, take the top-k.  No row of
    # a probed list may have been dropped by the scan's bound pruning (global per-query threshold, item pruning).
    assign = torch.cat([ix.assign_device(db[lo:lo + 1_000_000]) for lo in range(0, n, 1_000_000)]).long()
    assert torch.equal(torch.bincount(assign, minlength=nlist).cpu(), torch.from_numpy(sizes))
    Dt = np.empty((nq, k), np.float32)
    It = np.empty((nq, k), np.uint64)
    for qi in range(nq):
        rows = torch.isin(assign, probes[qi]).nonzero().squeeze(1)
        d = torch.empty(rows.numel(), dtype=torch.float64, device="cuda")
        for lo in range(0, rows.numel(), 131072):
            r = rows[lo:lo + 131072]
            d[lo:lo + 131072] = ((db[r].double() - q[qi].double()[None, :]) ** 2).sum(1)
        # top-k by (distance, id): a stable sort of ids-ascending rows by distance
        order = torch.argsort(d, stable=True)[:k]
        Dt[qi] = d[order].float().cpu().numpy()
        It[qi] = rows[order].cpu().numpy().astype(np.uint64)
    check_search(D32.cpu().numpy(), I32.cpu().numpy().view(np.uint64), Dt, It)
