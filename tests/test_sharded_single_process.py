"""GPU tests of vdb_index_create_sharded: ONE process, one list shard per entry of `devices`, merged through the
root shard's mailbox.  A device may be listed several times, so the whole path -- ownership, replicated centroids,
per-shard scans, merge kernels publishing into the root's mailbox, the collect kernel, pipelined tickets -- is
exercised on a single GPU too (the round-end box has one); with >= 2 GPUs the shards sit on different devices and
the mailbox stores cross NVLink."""
import importlib
import os

import numpy as np
import pytest

import oracle_lib as O
from parity import check_search, ip_scale

pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed, n, dim, nlist, ntrain, nq, nprobe, k, metric = (int(v) for v in g["params"])
    x = O.gaussian(seed, n + nq, dim)
    return g, x[:n], x[n:], dict(dim=dim, nlist=nlist, ntrain=ntrain, nprobe=nprobe, k=k, metric=metric)


def device_lists():
    import torch
    n = torch.cuda.device_count()
    out = [(0, 0), (0, 0, 0, 0)]
    if n >= 2:
        out.append(tuple(range(min(n, 8))))
    return out


@pytest.mark.parametrize("name", ["config1", "small_ip", "ctest_gpu_vs_cpu"])
def test_sharded_handle_reproduces_the_golden_results(name):
    """train + add + search through the one handle == the reference's results (golden fixtures), for every
    device list this box allows; the shards partition the rows and the owners are byte-balanced"""
    g, db, q, p = load_case(name)
    for devs in device_lists():
        ix = pkg.IVFFlatIndex(pkg.Config(dimension=p["dim"], nlist=p["nlist"], metric=pkg.Metric(p["metric"]),
                                         devices=devs))
        ix.train(db[:p["ntrain"]])
        assert np.array_equal(ix.centroids, g["centroids"]), "sharded train differs from the reference centroids"
        half = db.shape[0] // 2
        ix.add(db[:half])
        ix.add(db[half:], np.arange(half, db.shape[0], dtype=np.uint64))
        assert ix.get_total_vectors() == db.shape[0]
        assert np.array_equal(ix.list_sizes(), g["list_sizes"])
        owners = ix.owners()
        assert set(owners.tolist()) == set(range(len(devs)))
        load = np.array([int(g["list_sizes"][owners == r].sum()) for r in range(len(devs))])
        assert load.max() - load.min() <= max(int(g["list_sizes"].max()), db.shape[0] // 10), load
        scale = ip_scale(q, db) if p["metric"] == O.METRIC_IP else None
        D, I = ix.search(q, p["nprobe"], p["k"])
        check_search(D, I, g["D"], g["I"], scale)
        # same bits as the unsharded index
        one = pkg.IVFFlatIndex(pkg.Config(dimension=p["dim"], nlist=p["nlist"], metric=pkg.Metric(p["metric"])))
        one.centroids = g["centroids"]
        one.add(db)
        D1, I1 = one.search(q, p["nprobe"], p["k"])
        assert np.array_equal(D, D1) and np.array_equal(I, I1), f"devices={devs}: sharded != unsharded"
        # the ids of a list come from the shard that owns it
        l = int(np.argmax(g["list_sizes"]))
        assert sorted(ix.list_ids(l).tolist()) == sorted(one.list_ids(l).tolist())
        ix.close()
        one.close()


def test_sharded_handle_pipelines_batches_and_threads():
    import threading
    import torch
    g, db, q, p = load_case("config1")
    devs = device_lists()[-1]
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=p["dim"], nlist=p["nlist"], devices=devs, pipeline_depth=3))
    ix.centroids = g["centroids"]
    ix.add(torch.from_numpy(db).cuda())  # device rows on device 0: the other devices stage them over NVLink
    ix.reserve_search(16, p["nprobe"], p["k"])
    qd = torch.from_numpy(q).cuda()
    nq, k, bs = q.shape[0], p["k"], 16
    starts = [(5 * i) % (nq - bs) for i in range(11)]
    D = [torch.empty((bs, k), dtype=torch.float32, device="cuda:0") for _ in starts]
    I = [torch.empty((bs, k), dtype=torch.int64, device="cuda:0") for _ in starts]
    tickets = [ix.search_submit(qd[lo:lo + bs], p["nprobe"], k, D[i], I[i]) for i, lo in enumerate(starts)]
    for t in tickets:
        ix.search_wait(t)
    for i, lo in enumerate(starts):
        check_search(D[i].cpu().numpy(), I[i].cpu().numpy().view(np.uint64), g["D"][lo:lo + bs], g["I"][lo:lo + bs])
    # host threads through the synchronous call
    Dg, Ig = np.array(g["D"]), np.array(g["I"])
    errs = []

    def worker(lo):
        try:
            for _ in range(4):
                Dh, Ih = ix.search(q[lo:lo + 8], p["nprobe"], k)
                check_search(Dh, Ih, Dg[lo:lo + 8], Ig[lo:lo + 8])
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=worker, args=(lo,)) for lo in (0, 8, 16, 24, 32, 40)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    # larger k than the first mailbox: it is rebuilt between searches
    Dk, Ik = ix.search(q[:5], p["nprobe"], 100)
    one = pkg.IVFFlatIndex(pkg.Config(dimension=p["dim"], nlist=p["nlist"]))
    one.centroids = g["centroids"]
    one.add(db)
    D1, I1 = one.search(q[:5], p["nprobe"], 100)
    assert np.array_equal(Dk, D1) and np.array_equal(Ik, I1)
    with pytest.raises(ValueError):  # stream-ordered form is refused on the composite
        ix.search_async(qd[:4], 4, 4, D[0][:4, :4].contiguous(), I[0][:4, :4].contiguous(), 0)


def test_sharded_c4_shaped_inner_product():
    """BASELINE configs[3] in miniature: inner product, nlist 16384, nprobe 64 (centroids = data rows: the
    reference's O(nlist^2) seeding is infeasible at this nlist), 4 shards"""
    dim, nlist, n, nq, nprobe, k = 16, 16384, 30000, 16, 64, 10
    x = O.gaussian(77, n + nq, dim)
    db, q = x[:n], x[n:]
    cent = db[:nlist].copy()
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist, metric=pkg.Metric.InnerProduct,
                                     devices=device_lists()[-1] if len(device_lists()) > 2 else (0, 0, 0, 0)))
    ix.centroids = cent
    ix.add(db)
    D, I = ix.search(q, nprobe, k)
    ora = O.OracleIndex(dim, nlist, O.METRIC_IP)
    ora.centroids = cent
    ora.add(db)
    Dr, Ir = ora.search(q, nprobe, k, 8)
    check_search(D, I, Dr, Ir, ip_scale(q, db))


def test_sharded_add_of_device_rows_staged_in_chunks():
    """rows that live on device 0 reach the other shards chunk by chunk over NVLink (more than one 256 MB chunk):
    every row must land on exactly one shard, in the list the unsharded index puts it in"""
    import torch
    dim, nlist, n = 128, 64, 600_000
    gen = torch.Generator(device="cuda:0").manual_seed(3)
    x = torch.randn(n, dim, generator=gen, device="cuda:0")
    cent = x[:nlist].cpu().numpy()
    one = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    one.centroids = cent
    one.add(x)
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist, devices=device_lists()[-1]))
    ix.centroids = cent
    ix.add(x)
    assert ix.get_total_vectors() == n
    assert np.array_equal(ix.list_sizes(), one.list_sizes()) and int(ix.list_sizes().sum()) == n
    l = int(np.argmin(one.list_sizes()))
    assert sorted(ix.list_ids(l).tolist()) == sorted(one.list_ids(l).tolist())
    q = x[:9].cpu().numpy()
    Ds, Is = ix.search(q, nlist, 3)
    Do, Io = one.search(q, nlist, 3)
    assert np.array_equal(Ds, Do) and np.array_equal(Is, Io) and Is[:, 0].tolist() == list(range(9))


def test_cpp_mirror_spans_several_gpus_and_matches_the_golden_fixture(tmp_path):
    """the C++ vdb::IVFFlatIndex mirror with Config::devices (host/sharded_test.cpp): train / add / search / pipelined
    submit / save / load_from_epoch on a multi-device index, results == tests/golden/config1.npz (the reference's own
    CPU results on the same std::mt19937 data)"""
    import subprocess
    import torch
    exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "host", "sharded_test")
    if not os.path.exists(exe):
        pkg.build()
    devs = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    out = subprocess.run([exe, devs, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "PASSED" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    g = np.load(os.path.join(GOLD, "config1.npz"))
    rows = [l.split() for l in out.stdout.splitlines() if l.startswith("R ")]
    I = np.array([int(r[2]) for r in rows], np.uint64).reshape(g["I"].shape)
    D = np.array([float(r[3]) for r in rows], np.float32).reshape(g["D"].shape)
    check_search(D, I, g["D"], g["I"])
