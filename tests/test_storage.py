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


def test_cpp_writer_and_pyarrow_agree_on_the_reference_format(tmp_path):
    """libvdb_b200_storage.so (Arrow C++) and pyarrow both produce / consume ArrowStorage::write_vectors files
    (format/storage.cpp:183-226): {id: uint64, vector: list<float32>}, one record batch"""
    import ctypes as C
    import pyarrow as pa
    x = O.gaussian(11, 300, 17)
    ids = np.arange(300, dtype=np.uint64) * 3 + 5
    lib = storage.storage_lib()
    p1 = os.path.join(tmp_path, "cpp.arrow")
    storage._scheck(lib.vdb_storage_write_vectors(os.fsencode(p1), x.ctypes.data, ids.ctypes.data, 300, 17))
    with pa.memory_map(p1, "r") as src:
        rd = pa.ipc.open_file(src)
        assert rd.num_record_batches == 1
        t = rd.read_all()
    assert t.schema.names == ["id", "vector"] and str(t.schema.field("vector").type) == "list<item: float>"
    assert t.column("id").to_pylist() == ids.tolist()
    assert np.array_equal(np.array(t.column("vector").to_pylist(), np.float32), x)
    (rid, rv, _k), = storage.read_vectors(p1)  # the Python reader on the C++ file
    assert np.array_equal(rid, ids) and np.array_equal(rv, x)
    p2 = os.path.join(tmp_path, "py.arrow")
    storage.write_vectors(p2, x, ids)          # the C++ reader on the pyarrow file
    n, dim = C.c_uint64(), C.c_uint32()
    storage._scheck(lib.vdb_storage_read_vectors(os.fsencode(p2), None, None, C.byref(n), C.byref(dim)))
    assert (n.value, dim.value) == (300, 17)
    v2, i2 = np.empty((300, 17), np.float32), np.empty(300, np.uint64)
    storage._scheck(lib.vdb_storage_read_vectors(os.fsencode(p2), v2.ctypes.data, i2.ctypes.data, C.byref(n), C.byref(dim)))
    assert np.array_equal(v2, x) and np.array_equal(i2, ids)
    with pytest.raises(ValueError):
        storage._scheck(lib.vdb_storage_read_vectors(os.fsencode(os.path.join(tmp_path, "missing.arrow")), None, None,
                                                     C.byref(n), C.byref(dim)))


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [(), (0, 0, 0)])
def test_epoch_save_and_load_round_trip(tmp_path, devices):
    """IVFFlatIndex::save / load (ivf_flat_index.h:66-67) through the reference's epoch layout: manifest.json with
    the reference's fields, centroids.arrow, one list file per non-empty list; a loaded index -- unsharded or a
    single-process sharded one, which re-balances ownership from the manifest -- answers exactly like the saved one"""
    import json
    from parity import check_search
    dim, nlist, n = 40, 12, 6000
    x = O.gaussian(21, n + 16, dim)
    db, q = x[:n], x[n:]
    ids = np.arange(n, dtype=np.uint64) * 2 + 7
    a = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist, metric=pkg.Metric.InnerProduct))
    a.train(db[:1500])
    a.add(db, ids)
    d = os.path.join(tmp_path, "epoch_000001")
    storage.save_epoch(a, d, "demo", "epoch_000001")
    m = json.load(open(os.path.join(d, "manifest.json")))
    assert (m["index_name"], m["epoch"], m["dimension"], m["nlist"], m["metric"]) == ("demo", "epoch_000001", dim, nlist, "InnerProduct")
    assert m["pq_params"] == {"m": 0, "nbits": 8} and m["created_at"] > 0
    sizes = a.list_sizes()
    assert {s["list_id"]: s["num_vectors"] for s in m["shards"]} == {l: int(c) for l, c in enumerate(sizes) if c}
    for s in m["shards"]:
        assert os.path.getsize(os.path.join(d, s["path"])) == s["file_size"]
        (rid, rv, _k), = storage.read_vectors(os.path.join(d, s["path"]))
        assert sorted(rid.tolist()) == sorted(a.list_ids(s["list_id"]).tolist()) and rv.shape == (s["num_vectors"], dim)
    b = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist, metric=pkg.Metric.InnerProduct, devices=devices))
    storage.load_epoch(b, d)
    assert b.get_total_vectors() == n and np.array_equal(b.list_sizes(), sizes)
    assert np.array_equal(b.centroids, a.centroids)
    Da, Ia = a.search(q, 5, 10)
    Db, Ib = b.search(q, 5, 10)
    assert np.array_equal(Da, Db) and np.array_equal(Ia, Ib)
    ora = O.OracleIndex(dim, nlist, O.METRIC_IP)
    ora.train(db[:1500])
    ora.add(db, ids)
    Dr, Ir = ora.search(q, 5, 10)
    from parity import ip_scale
    check_search(Db, Ib, Dr, Ir, ip_scale(q, db))
    b.add(db[:100], np.arange(100, dtype=np.uint64) + 10**6)  # a loaded index keeps growing
    assert b.get_total_vectors() == n + 100
    # mismatching target: refused, nothing loaded
    c = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist + 1, metric=pkg.Metric.InnerProduct))
    with pytest.raises(ValueError):
        storage.load_epoch(c, d)
    with pytest.raises(ValueError):
        storage.load_epoch(b, d)  # not empty


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
