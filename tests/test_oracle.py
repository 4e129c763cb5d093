"""CPU tests: the oracle restatement (oracle/ivf_oracle.c) against
(a) the committed golden fixtures generated from the reference's own CPU path,
(b) the reference itself (oracle/_ref/libvdbref.so) when it is built here.
Bit-exact: ids, distances, centroids, probe lists."""
import os

import numpy as np
import pytest

import oracle_lib as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["simple_test", "ctest_gpu_vs_cpu", "small_ip", "config1"]


def load_case(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed, n, dim, nlist, ntrain, nq, nprobe, k, metric = (int(v) for v in g["params"])
    x = O.gaussian(seed, n + nq, dim)
    return g, x[:n], x[n:], dict(dim=dim, nlist=nlist, ntrain=ntrain, nprobe=nprobe, k=k, metric=metric)


@pytest.mark.parametrize("name", CASES)
def test_port_matches_golden(name):
    g, db, q, p = load_case(name)
    assert np.array_equal(db[:4], g["db_head"]) and np.array_equal(q[:2], g["q_head"])
    ix = O.OracleIndex(p["dim"], p["nlist"], p["metric"])
    ix.train(db[: p["ntrain"]])
    assert np.array_equal(ix.centroids, g["centroids"])
    ix.add(db)
    assert np.array_equal(ix.list_sizes(), g["list_sizes"])
    assert ix.ntotal == db.shape[0]
    for i in range(q.shape[0]):
        assert np.array_equal(ix.select_nprobe(q[i], p["nprobe"]), g["probes"][i])
    D, I = ix.search(q, p["nprobe"], p["k"])
    assert np.array_equal(I, g["I"])
    assert np.array_equal(D, g["D"])
    # threads only partition queries
    D2, I2 = ix.search(q, p["nprobe"], p["k"], nthreads=4)
    assert np.array_equal(I2, I) and np.array_equal(D2, D)


needs_ref = pytest.mark.skipif(O.ref_lib() is None, reason="reference not built (no /root/reference)")


@needs_ref
@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP, O.METRIC_COSINE])
@pytest.mark.parametrize("seed", [1, 7])
def test_port_matches_reference_random(metric, seed):
    rng = np.random.default_rng(seed)
    n, dim, nlist, nq = int(rng.integers(300, 1500)), int(rng.choice([3, 17, 32, 100])), int(rng.integers(2, 40)), 9
    x = O.gaussian(seed, n + nq, dim)
    db, q = x[:n], x[n:]
    a, b = O.OracleIndex(dim, nlist, metric), O.RefIndex(dim, nlist, metric)
    a.train(db[: n // 2])
    b.train(db[: n // 2])
    assert np.array_equal(a.centroids, b.centroids)
    assert np.array_equal(a.assign(db), b.assign(db))
    ids = rng.permutation(n).astype(np.uint64) + 1000
    for lo in range(0, n, 400):  # several add() calls append
        a.add(db[lo:lo + 400], ids[lo:lo + 400])
        b.add(db[lo:lo + 400], ids[lo:lo + 400])
    assert np.array_equal(a.list_sizes(), b.list_sizes())
    for l in range(nlist):
        assert np.array_equal(a.list_ids(l), b.list_ids(l))
    for nprobe, k in [(1, 1), (3, 10), (nlist, 7), (nlist + 5, 50)]:
        Da, Ia = a.search(q, nprobe, k)
        Db, Ib = b.search(q, nprobe, k)
        assert np.array_equal(Ia, Ib) and np.array_equal(Da, Db)


@needs_ref
def test_edge_cases_match_reference():
    dim, nlist = 8, 6
    x = O.gaussian(3, 40, dim)
    a, b = O.OracleIndex(dim, nlist), O.RefIndex(dim, nlist)
    cent = O.gaussian(4, nlist, dim)
    a.centroids = cent
    b.centroids = cent
    # untrained/empty index: everything padded
    Da, Ia = a.search(x[:2], 3, 4)
    Db, Ib = b.search(x[:2], 3, 4)
    assert np.array_equal(Ia, Ib) and (Ia == O.ID_PAD).all() and (Da == O.FLT_MAX).all() and np.array_equal(Da, Db)
    # duplicate ids (same id in two lists and twice in one list) are de-duplicated by merge_results
    ids = np.arange(40, dtype=np.uint64) % 13
    a.add(x, ids)
    b.add(x, ids)
    for nprobe, k in [(6, 5), (6, 40), (2, 3)]:
        Da, Ia = a.search(x[:7], nprobe, k)
        Db, Ib = b.search(x[:7], nprobe, k)
        assert np.array_equal(Ia, Ib) and np.array_equal(Da, Db)
        for row in Ia:
            real = row[row != O.ID_PAD]
            assert len(set(real.tolist())) == len(real)


def test_flat_search_is_nlist1():
    x = O.gaussian(5, 520, 24)
    db, q = x[:500], x[500:]
    ix = O.OracleIndex(24, 1)
    ix.centroids = np.zeros((1, 24), np.float32)
    ix.add(db)
    D1, I1 = ix.search(q, 1, 100)
    D2, I2 = O.flat_search(db, q, 100)
    assert np.array_equal(I1, I2) and np.array_equal(D1, D2)


def test_invalid_config_rejected():
    with pytest.raises(ValueError):
        O.OracleIndex(0, 4)
    with pytest.raises(ValueError):
        O.OracleIndex(4, 0)


@needs_ref
def test_load_assigned_equals_add():
    x = O.gaussian(9, 600, 12)
    for cls in (O.OracleIndex, O.RefIndex):
        a, b = cls(12, 5), cls(12, 5)
        cent = O.gaussian(10, 5, 12)
        a.centroids = cent
        b.centroids = cent
        ids = np.arange(600, dtype=np.uint64) + 5
        a.add(x, ids)
        b.load_assigned(x, ids, a.assign(x))
        Da, Ia = a.search(x[:9], 3, 10)
        Db, Ib = b.search(x[:9], 3, 10)
        assert np.array_equal(Ia, Ib) and np.array_equal(Da, Db) and a.ntotal == b.ntotal
