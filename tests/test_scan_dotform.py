"""The L2 list scan screens (row, query) pairs with |q|^2 + |v|^2 - 2 q.v minus a proven rounding bound and computes
the exact (q - v)^2 only for the pairs that pass.  The screen may only ever admit MORE than the exact test would:
results must be bit-identical to the unscreened kernel (VDB_SCAN_EXACT=1, read when an index is created) and agree
with the oracle -- including where the bound is tight (large norms, tiny distances) or loose (huge offsets)."""
import importlib
import os

import numpy as np
import pytest

import oracle_lib as O
from parity import check_search

pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
pytestmark = pytest.mark.gpu


def build(dim, nlist, cent, db, exact):
    # both knobs are read when an index is created.  The screen is normally reserved for long scans (>= 12 items per
    # CTA); the small cases here force it on.
    old = {k_: os.environ.get(k_) for k_ in ("VDB_SCAN_EXACT", "VDB_SCAN_DOT_MIN_ROWS")}
    os.environ["VDB_SCAN_EXACT"] = "1" if exact else "0"
    os.environ["VDB_SCAN_DOT_MIN_ROWS"] = "0"
    try:
        ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    finally:
        for k_, v in old.items():
            if v is None:
                os.environ.pop(k_, None)
            else:
                os.environ[k_] = v
    ix.centroids = cent
    half = db.shape[0] // 2
    ix.add(db[:half])
    ix.add(db[half:])
    return ix


CASES = {
    # name: (dim, nlist, n, nq, nprobe, k, generator)
    "gaussian768": (768, 64, 30000, 40, 16, 10, lambda n, d: O.gaussian(1, n, d)),
    "odd_width_100": (100, 32, 20000, 33, 32, 25, lambda n, d: O.gaussian(2, n, d)),
    "wide_2048": (2048, 8, 3000, 9, 8, 5, lambda n, d: O.gaussian(3, n, d)),
    # a far-away cloud: |q|^2 + |v|^2 ~ 1e6 x the distances, so the slack dwarfs them and every pair is re-scored
    "huge_offset": (64, 16, 20000, 20, 16, 10, lambda n, d: O.gaussian(4, n, d) * 0.01 + 100.0),
    # tight clusters: distances of 1e-3 .. 1e-2 between points of norm ~ 8
    "tight_clusters": (64, 16, 20000, 20, 8, 10, lambda n, d: O.clustered(5, n, d, 16, 0.002)),
    # exact duplicates and zero vectors: distance 0 must survive a screen that can go negative
    "duplicates_and_zeros": (32, 8, 8000, 16, 8, 20,
                             lambda n, d: np.concatenate([np.zeros((n // 4, d), np.float32),
                                                          np.repeat(O.gaussian(6, n // 8, d), 6, axis=0)])[:n]),
    "large_k": (128, 16, 30000, 12, 16, 500, lambda n, d: O.gaussian(7, n, d)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_screened_scan_is_bit_identical_to_the_exact_scan_and_matches_the_oracle(name):
    dim, nlist, n, nq, nprobe, k, gen = CASES[name]
    x = np.ascontiguousarray(gen(n + nq, dim), np.float32)
    db, q = x[:n], x[n:]
    cent = db[:: n // nlist][:nlist].copy()
    a = build(dim, nlist, cent, db, exact=False)
    b = build(dim, nlist, cent, db, exact=True)
    for np_, k_ in ((nprobe, k), (nlist, k), (1, 3)):
        Da, Ia = a.search(q, np_, k_)
        Db, Ib = b.search(q, np_, k_)
        assert np.array_equal(Da, Db) and np.array_equal(Ia, Ib), f"{name}: screen changed the result (nprobe {np_})"
    ora = O.OracleIndex(dim, nlist)
    ora.centroids = cent
    ora.add(db)
    Dr, Ir = ora.search(q, nprobe, k, 8)
    Da, Ia = a.search(q, nprobe, k)
    if name in ("huge_offset",):
        # the oracle's own fp32 rounding of distances ~1e-2 computed from coordinates ~100 is ~1e-5 absolute: compare
        # on the coordinate scale (tolerance is relative to the magnitudes summed)
        scale = np.full(nq, float((db.astype(np.float64) ** 2).sum(1).max()) * 1e-3)
        check_search(Da, Ia, Dr, Ir, scale)
    else:
        check_search(Da, Ia, Dr, Ir)


def test_loaded_epoch_has_norms_too(tmp_path):
    """rows that arrive through vdb_index_append_list (epoch load) get their norms from page_norms_kernel, not the
    scatter kernel: the screened scan over them must equal the exact scan"""
    storage = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.storage")
    dim, nlist, n = 48, 10, 9000
    x = O.gaussian(11, n + 12, dim)
    db, q = x[:n], x[n:]
    cent = db[:nlist].copy()
    src = build(dim, nlist, cent, db, exact=True)
    d = os.path.join(tmp_path, "ep")
    storage.save_epoch(src, d)
    os.environ["VDB_SCAN_DOT_MIN_ROWS"] = "0"
    try:
        scr = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))  # screened
    finally:
        os.environ.pop("VDB_SCAN_DOT_MIN_ROWS", None)
    storage.load_epoch(scr, d)
    D0, I0 = src.search(q, 6, 10)
    D1, I1 = scr.search(q, 6, 10)
    assert np.array_equal(D0, D1) and np.array_equal(I0, I1)
