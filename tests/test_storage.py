"""Arrow vector files (the reference's format/storage.cpp layout): CPU round trip; GPU load into HBM pages."""
import importlib
import os

import numpy as np
import pytest

import oracle_lib as O

storage = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.storage")
pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")


def test_arrow_round_trip_is_zero_copy(tmp_path):
    x = O.gaussian(3, 500, 24)
    ids = (np.arange(500, dtype=np.uint64) * 7 + 1)
    p = os.path.join(tmp_path, "vectors.arrow")
    storage.write_vectors(p, x, ids)
    import pyarrow as pa
    with pa.memory_map(p, "r") as src:
        t = pa.ipc.open_file(src).read_all()
    assert t.schema.names == ["id", "vector"] and str(t.schema.field("vector").type) == "list<item: float>"
    (rid, rv, _k), = storage.read_vectors(p)
    assert np.array_equal(rid, ids) and np.array_equal(rv, x)
    assert not rv.flags.owndata  # a view of the mapped file, not a copy


@pytest.mark.gpu
def test_arrow_file_loads_into_hbm_index(tmp_path):
    dim, nlist, n = 32, 8, 4000
    x = O.gaussian(5, n + 10, dim)
    db, q = x[:n], x[n:]
    ids = np.arange(n, dtype=np.uint64) + 100
    p = os.path.join(tmp_path, "shard0.arrow")
    storage.write_vectors(p, db, ids)
    ora = O.OracleIndex(dim, nlist)
    ora.train(db[:1000])
    ora.add(db, ids)
    cp = os.path.join(tmp_path, "centroids.arrow")
    storage.write_vectors(cp, ora.centroids)
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist))
    storage.load_centroids(ix, cp)
    assert storage.load_into_index(ix, p) == n
    assert np.array_equal(ix.list_sizes(), ora.list_sizes())
    D, I = ix.search(q, 4, 10)
    Dr, Ir = ora.search(q, 4, 10)
    from parity import check_search
    check_search(D, I, Dr, Ir)
