"""The tensor-core screen of the list scan (csrc/screen.cuh): an index that keeps a low-precision shadow of its pages
(bf16, or int8 with a scale per row) streams a half / a quarter of the bytes, bounds every (row, query) distance from
below with the tensor-core dot product of the shadow operands and re-scores only the admitted pairs -- with the fp32
scan's own arithmetic.  The screen may only ever admit
MORE than the exact test would, so results must be BIT-IDENTICAL to the plain fp32 scan (VDB_SCAN_EXACT=1) and agree
with the oracle, for both metrics, every supported row width, ragged pages, tiny and huge k, and where the bound is
tight (large norms, tiny distances) or useless (huge offsets: everything is re-scored)."""
import importlib
import os

import numpy as np
import pytest

import oracle_lib as O
from parity import check_search

pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
pytestmark = pytest.mark.gpu


def build(dim, nlist, metric, cent, db, mirror, chunks=2):
    old = {k_: os.environ.get(k_) for k_ in ("VDB_SCAN_EXACT", "VDB_SCAN_MIRROR")}
    os.environ["VDB_SCAN_EXACT"] = "0" if mirror else "1"
    os.environ["VDB_SCAN_MIRROR"] = str(int(mirror))  # 0 = none (and the fp32 scan), 1 = bf16, 2 = int8
    try:
        ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist, metric=metric))
    finally:
        for k_, v in old.items():
            if v is None:
                os.environ.pop(k_, None)
            else:
                os.environ[k_] = v
    ix.centroids = cent
    step = (db.shape[0] + chunks - 1) // chunks
    for lo in range(0, db.shape[0], step):  # several adds: rows land behind a partly filled tail page
        ix.add(db[lo:lo + step])
    return ix


CASES = {
    # name: (dim, nlist, n, nq, nprobe, k, generator)
    "gaussian768": (768, 64, 40000, 64, 16, 10, lambda n, d: O.gaussian(1, n, d)),
    "gaussian768_few_queries": (768, 16, 9000, 3, 16, 10, lambda n, d: O.gaussian(2, n, d)),
    "w128": (128, 32, 50000, 40, 32, 25, lambda n, d: O.gaussian(3, n, d)),
    "w256_k1": (256, 16, 20000, 17, 4, 1, lambda n, d: O.gaussian(4, n, d)),
    "w512": (512, 24, 20000, 33, 24, 40, lambda n, d: O.gaussian(5, n, d)),
    "w1024": (1024, 8, 6000, 9, 8, 5, lambda n, d: O.gaussian(6, n, d)),
    "large_k": (128, 16, 30000, 12, 16, 500, lambda n, d: O.gaussian(7, n, d)),
    # a far-away cloud: the bf16 rounding error of coordinates ~100 dwarfs the distances, every pair is re-scored
    "huge_offset": (128, 16, 20000, 20, 16, 10, lambda n, d: O.gaussian(8, n, d) * 0.01 + 100.0),
    "tight_clusters": (128, 16, 20000, 20, 8, 10, lambda n, d: O.clustered(9, n, d, 16, 0.002)),
    "duplicates_and_zeros": (128, 8, 8000, 16, 8, 20,
                             lambda n, d: np.concatenate([np.zeros((n // 4, d), np.float32),
                                                          np.repeat(O.gaussian(10, n // 8, d), 6, axis=0)])[:n]),
    # lists of a handful of rows: every tile is ragged
    "tiny_lists": (256, 200, 1500, 30, 50, 10, lambda n, d: O.gaussian(11, n, d)),
    # values that bf16 represents exactly (small integers): zero rounding error, heavy distance ties
    "integers": (128, 8, 12000, 16, 8, 30, lambda n, d: np.round(O.gaussian(12, n, d) * 2.0).astype(np.float32)),
}


BF16, I8 = 1, 2


@pytest.mark.parametrize("kind", [BF16, I8])
@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
@pytest.mark.parametrize("name", sorted(CASES))
def test_mirror_scan_is_bit_identical_to_the_fp32_scan_and_matches_the_oracle(name, metric, kind):
    dim, nlist, n, nq, nprobe, k, gen = CASES[name]
    x = np.ascontiguousarray(gen(n + nq, dim), np.float32)
    db, q = x[:n], x[n:]
    cent = db[:: n // nlist][:nlist].copy()
    a = build(dim, nlist, metric, cent, db, mirror=kind, chunks=3)
    b = build(dim, nlist, metric, cent, db, mirror=0)
    for np_, k_ in ((nprobe, k), (nlist, k), (1, 3)):
        Da, Ia = a.search(q, np_, k_)
        # the screen kernel really ran
        assert a.last_search_stats().streamed_bytes_per_row == (2 * dim + 8 if kind == BF16 else dim + 12)
        Db, Ib = b.search(q, np_, k_)
        assert b.last_search_stats().streamed_bytes_per_row == 4 * dim + 8
        assert np.array_equal(Da, Db) and np.array_equal(Ia, Ib), f"{name}: the screen changed the result (nprobe {np_})"
    ora = O.OracleIndex(dim, nlist, metric)
    ora.centroids = cent
    ora.add(db)
    Dr, Ir = ora.search(q, nprobe, k, 8)
    Da, Ia = a.search(q, nprobe, k)
    if name == "huge_offset":
        # the oracle's own fp32 rounding of sums of ~1e4-sized terms: compare on the scale of the magnitudes summed
        #  -- |v|^2 ~ 1.3e6 for the inner product, 1e-3 of it for the differences of the L2 form
        scale = np.full(nq, float((db.astype(np.float64) ** 2).sum(1).max()) * (1e-3 if metric == O.METRIC_L2 else 1.0))
        check_search(Da, Ia, Dr, Ir, scale)
    else:
        check_search(Da, Ia, Dr, Ir)


@pytest.mark.parametrize("kind", [BF16, I8])
def test_batches_wider_than_the_screen_are_chunked_and_agree(kind):
    dim, nlist, n = 128, 32, 30000
    x = O.gaussian(21, n + 200, dim)
    db, q = x[:n], x[n:]
    cent = db[:nlist].copy()
    a = build(dim, nlist, O.METRIC_L2, cent, db, mirror=kind)
    b = build(dim, nlist, O.METRIC_L2, cent, db, mirror=0)
    Da, Ia = a.search(q, 8, 10)          # 200 queries: four pipelined chunks of <= 64 through the screen kernel
    Db, Ib = b.search(q, 8, 10)
    assert np.array_equal(Da, Db) and np.array_equal(Ia, Ib)
    for lo in range(0, 200, 64):         # 64 at a time: the screen kernel
        D, I = a.search(q[lo:lo + 64], 8, 10)
        assert np.array_equal(D, Db[lo:lo + 64]) and np.array_equal(I, Ib[lo:lo + 64])


@pytest.mark.parametrize("kind", [BF16, I8])
def test_loaded_epoch_gets_its_shadow(tmp_path, kind):
    """rows that arrive through vdb_index_append_list (epoch load) get norms, error norms and the bf16 shadow from
    page_norms_kernel, not the scatter kernel"""
    storage = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.storage")
    dim, nlist, n = 256, 10, 9000
    x = O.gaussian(31, n + 12, dim)
    db, q = x[:n], x[n:]
    cent = db[:nlist].copy()
    src = build(dim, nlist, O.METRIC_L2, cent, db, mirror=0)
    d = os.path.join(tmp_path, "ep")
    storage.save_epoch(src, d)
    os.environ["VDB_SCAN_MIRROR"] = str(kind)
    try:
        scr = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    finally:
        os.environ.pop("VDB_SCAN_MIRROR", None)
    storage.load_epoch(scr, d)
    D0, I0 = src.search(q, 6, 10)
    D1, I1 = scr.search(q, 6, 10)
    assert np.array_equal(D0, D1) and np.array_equal(I0, I1)
    scr.add(O.gaussian(32, 3000, dim))    # and keeps growing behind the loaded rows
    src.add(O.gaussian(32, 3000, dim))
    D0, I0 = src.search(q, 10, 10)
    D1, I1 = scr.search(q, 10, 10)
    assert np.array_equal(D0, D1) and np.array_equal(I0, I1)


def test_unsupported_shapes_refuse_an_explicit_shadow_and_auto_skips_it():
    for want in (2, 3):
        with pytest.raises(Exception):
            pkg.IVFFlatIndex(pkg.Config(dimension=100, nlist=4, scan_mirror=want))
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=100, nlist=4))  # auto: no shadow, fp32 scan
    x = O.gaussian(41, 3000, 100)
    ix.centroids = x[:4].copy()
    ix.add(x)
    D, I = ix.search(x[:5], 4, 3)
    assert (I[:, 0] == np.arange(5)).all()
