"""The gRPC contract (proto/vdb.proto) served over the C ABI.  CPU part: wire compatibility of the dynamically
built messages and the reference's ErrorHandling / CreateIndex / GetStats / Warmup cases
(test/integration/grpc_integration_test.cpp:80-321), none of which touch the GPU.  GPU part: BuildEpoch from an
Arrow file, Search parity against the oracle, request coalescing under concurrent clients."""
import importlib
import os
import threading

import grpc
import numpy as np
import pytest

import oracle_lib as O

pkg = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")
srv = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.server")
storage = importlib.import_module("cuda-acceleratedvectordatabaseengine_b200.storage")


@pytest.fixture()
def endpoint():
    server, servicer = srv.serve(pkg, address="127.0.0.1:0")
    client = srv.Client(f"127.0.0.1:{server.bound_port}")
    yield client, servicer
    client.close()
    server.stop(0)
    servicer.coalescer.close()


def code(fn, req):
    try:
        fn(req, timeout=30)
        return grpc.StatusCode.OK
    except grpc.RpcError as e:
        return e.code()


def test_wire_format_matches_vdb_proto():
    # hand-encoded SearchRequest: queries[0]{id=7, values=[1.0]}, topk=3, nprobe=2, index="a"
    raw = bytes([0x0A, 0x08, 0x08, 0x07, 0x12, 0x04, 0x00, 0x00, 0x80, 0x3F, 0x10, 0x03, 0x18, 0x02, 0x22, 0x01, 0x61])
    r = srv.SearchRequest.FromString(raw)
    assert r.queries[0].id == 7 and list(r.queries[0].values) == [1.0] and r.topk == 3 and r.nprobe == 2 and r.index == "a"
    assert r.SerializeToString() == raw
    s = srv.StatsResponse(total_vectors=5, gpu_memory_used=1.5)
    assert s.SerializeToString() == bytes([0x08, 0x05, 0x25, 0x00, 0x00, 0xC0, 0x3F])


def test_error_handling_cases(endpoint):
    c, _ = endpoint
    q = srv.SearchRequest(index="x", topk=5)
    assert code(c.Search, q) == grpc.StatusCode.INVALID_ARGUMENT  # no queries
    q.queries.add().values.extend([0.0] * 8)
    q.topk = 0
    assert code(c.Search, q) == grpc.StatusCode.INVALID_ARGUMENT  # topk = 0
    q.topk = 1001
    assert code(c.Search, q) == grpc.StatusCode.INVALID_ARGUMENT
    q.topk, q.index = 5, ""
    assert code(c.Search, q) == grpc.StatusCode.INVALID_ARGUMENT  # no index name
    q.index = "missing"
    assert code(c.Search, q) == grpc.StatusCode.NOT_FOUND


def test_create_index_stats_warmup(endpoint):
    c, _ = endpoint
    req = srv.CreateIndexRequest(name="idx", dimension=16, metric="L2", nlist=4)
    assert code(c.CreateIndex, req) == grpc.StatusCode.OK
    assert code(c.CreateIndex, req) == grpc.StatusCode.ALREADY_EXISTS
    st = c.GetStats(srv.StatsRequest(index="idx"))
    assert st.total_vectors == 0 and st.gpu_memory_used == 0.0
    assert code(c.GetStats, srv.StatsRequest(index="nope")) == grpc.StatusCode.NOT_FOUND
    assert code(c.Warmup, srv.WarmupRequest(index="idx")) == grpc.StatusCode.OK
    assert code(c.Warmup, srv.WarmupRequest(index="nope")) == grpc.StatusCode.NOT_FOUND
    s = srv.SearchRequest(index="idx", topk=3)
    s.queries.add().values.extend([0.0] * 16)
    assert code(c.Search, s) in (grpc.StatusCode.OK, grpc.StatusCode.NOT_FOUND, grpc.StatusCode.FAILED_PRECONDITION)
    assert code(c.CreateIndex, srv.CreateIndexRequest(name="c", dimension=4, metric="Cosine", nlist=2)) == \
        grpc.StatusCode.INVALID_ARGUMENT


def test_coalescer_batches_and_splits():
    calls = []

    def fake(key, q):
        calls.append(q.shape[0])
        return q[:, :2] * 2, np.arange(q.shape[0] * 2).reshape(-1, 2)

    co = srv.RequestCoalescer(fake, batch_size=8, window_ms=50)
    out = [None] * 6
    th = [threading.Thread(target=lambda i=i: out.__setitem__(i, co.submit("k", np.full((2, 3), i, np.float32))))
          for i in range(6)]
    [t.start() for t in th]
    [t.join() for t in th]
    co.close()
    assert sum(calls) == 12 and len(calls) <= 3  # 12 queries served by at most 3 searches
    for i in range(6):
        assert out[i][0].shape == (2, 2) and (out[i][0] == 2 * i).all()


@pytest.mark.gpu
def test_build_epoch_and_search_parity(endpoint, tmp_path):
    c, servicer = endpoint
    dim, nlist, n, nq = 32, 8, 5000, 12
    x = O.gaussian(17, n + nq, dim)
    db, q = x[:n], x[n:]
    path = os.path.join(tmp_path, "src.arrow")
    storage.write_vectors(path, db)
    assert code(c.CreateIndex, srv.CreateIndexRequest(name="e2e", dimension=dim, metric="L2", nlist=nlist)) == grpc.StatusCode.OK
    assert code(c.BuildEpoch, srv.BuildEpochRequest(index="e2e", source_path=path)) == grpc.StatusCode.OK
    st = c.GetStats(srv.StatsRequest(index="e2e"))
    assert st.total_vectors == n and st.gpu_memory_used > 0 and st.current_epoch
    ora = O.OracleIndex(dim, nlist)
    ora.train(db)
    ora.add(db)
    Dr, Ir = ora.search(q, 8, 5)  # nprobe defaults to 8 when the request leaves it 0

    def one(i, res):
        r = srv.SearchRequest(index="e2e", topk=5)
        r.queries.add().values.extend(q[i].tolist())
        res[i] = c.Search(r, timeout=60)

    res = [None] * nq
    th = [threading.Thread(target=one, args=(i, res)) for i in range(nq)]
    [t.start() for t in th]
    [t.join() for t in th]
    from parity import check_search
    D = np.array([[nb.distance for nb in r.results[0].neighbors] for r in res], np.float32)
    I = np.array([[nb.id for nb in r.results[0].neighbors] for r in res], np.uint64)
    check_search(D, I, Dr, Ir)
    bad = srv.SearchRequest(index="e2e", topk=5)
    bad.queries.add().values.extend([0.0] * (dim + 1))
    assert code(c.Search, bad) == grpc.StatusCode.INVALID_ARGUMENT  # dimension mismatch
    assert "vdb_searches_total" in servicer.metrics_text()


@pytest.mark.gpu
@pytest.mark.parametrize("devices", [(), (0, 0)])
def test_epochs_are_persisted_activated_and_measured(tmp_path, devices):
    """N3 + N4: BuildEpoch writes the reference's epoch directory, ActivateEpoch / LoadIndex swap a stored epoch
    back in (query_service.cpp:515-519, 232-257), GetStats names the serving epoch, and /metrics renders the
    reference's four Prometheus series from real counters (query_service.cpp:748-780)."""
    import json
    import urllib.request
    server, servicer = srv.serve(pkg, address="127.0.0.1:0", data_dir=str(tmp_path / "data"), devices=devices)
    c = srv.Client(f"127.0.0.1:{server.bound_port}")
    try:
        dim, nlist, n = 24, 6, 3000
        a, b = O.gaussian(1, n, dim), O.gaussian(2, n, dim) + 3.0
        pa_, pb_ = os.path.join(tmp_path, "a.arrow"), os.path.join(tmp_path, "b.arrow")
        storage.write_vectors(pa_, a)
        storage.write_vectors(pb_, b, np.arange(n, dtype=np.uint64) + 10**6)
        assert code(c.CreateIndex, srv.CreateIndexRequest(name="ix", dimension=dim, metric="L2", nlist=nlist)) == grpc.StatusCode.OK
        assert code(c.BuildEpoch, srv.BuildEpochRequest(index="ix", source_path=pa_)) == grpc.StatusCode.OK
        e1 = c.GetStats(srv.StatsRequest(index="ix")).current_epoch
        assert code(c.BuildEpoch, srv.BuildEpochRequest(index="ix", source_path=pb_)) == grpc.StatusCode.OK
        e2 = c.GetStats(srv.StatsRequest(index="ix")).current_epoch
        assert e1 and e2 and e1 != e2
        for e in (e1, e2):  # both epochs are complete directories in the reference's layout
            m = json.load(open(os.path.join(tmp_path, "data", "ix", e, "manifest.json")))
            assert m["index_name"] == "ix" and m["epoch"] == e and m["dimension"] == dim and m["nlist"] == nlist
            assert sum(s["num_vectors"] for s in m["shards"]) == n

        def nearest(vec):
            r = srv.SearchRequest(index="ix", topk=1, nprobe=nlist)
            r.queries.add().values.extend(vec.tolist())
            return c.Search(r, timeout=60).results[0].neighbors[0]

        assert nearest(b[5]).id == 10**6 + 5 and nearest(b[5]).distance == 0.0  # epoch 2 serves
        assert code(c.ActivateEpoch, srv.ActivateEpochRequest(index="ix", epoch=e1)) == grpc.StatusCode.OK
        assert c.GetStats(srv.StatsRequest(index="ix")).current_epoch == e1
        assert nearest(a[7]).id == 7 and nearest(a[7]).distance == 0.0           # epoch 1 again, loaded from disk
        assert code(c.LoadIndex, srv.LoadIndexRequest(index="ix", epoch=e2)) == grpc.StatusCode.OK
        assert nearest(b[9]).id == 10**6 + 9
        assert code(c.ActivateEpoch, srv.ActivateEpochRequest(index="ix", epoch="epoch_0_none")) == grpc.StatusCode.NOT_FOUND
        assert code(c.ActivateEpoch, srv.ActivateEpochRequest(index="nope", epoch=e1)) == grpc.StatusCode.NOT_FOUND
        httpd, port = servicer.serve_metrics()
        text = urllib.request.urlopen(f"http://127.0.0.1:{port}/metrics", timeout=10).read().decode()
        httpd.shutdown()
        for series in ("vdb_search_duration_milliseconds", "vdb_searches_total", "vdb_gpu_memory_bytes",
                       "vdb_queries_per_second"):
            assert f"# TYPE {series}" in text
        vals = dict(line.rsplit(" ", 1) for line in text.splitlines() if line and not line.startswith("#"))
        assert float(vals['vdb_searches_total{index="ix"}']) == 5
        assert float(vals["vdb_gpu_memory_bytes"]) > 0 and float(vals["vdb_queries_per_second"]) > 0
        assert float(vals['vdb_search_duration_milliseconds{index="ix",quantile="0.99"}']) > 0
        assert float(vals['vdb_hbm_bytes_scanned_total{index="ix"}']) > 0
    finally:
        c.close()
        server.stop(0)
        servicer.coalescer.close()
