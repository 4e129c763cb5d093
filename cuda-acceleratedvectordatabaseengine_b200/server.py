"""gRPC front for the B200 IVF-Flat path: the reference's wire contract
(proto/vdb.proto:10-109 -- vdb.QueryService {Search, Warmup, LoadIndex}, vdb.AdminService {CreateIndex,
BuildEpoch, ActivateEpoch, GetStats}) served from Python over the C ABI.

* There is no protoc / grpc_tools in this image, so the message classes are built at import time from a
  FileDescriptorProto that restates vdb.proto field by field (same names, numbers and types => byte-compatible).
* Validation, defaults and status codes follow server/query_service.cpp: no queries / topk outside [1,1000] /
  empty index name / dimension mismatch -> INVALID_ARGUMENT (:72-85,:117-120), unknown index -> NOT_FOUND (:88-92),
  nprobe <= 0 -> 8 (:97), padded (UINT64_MAX) results dropped (:150), engine exceptions -> INTERNAL (:164-167),
  CreateIndex twice -> ALREADY_EXISTS, GetStats.gpu_memory_used in GiB (:539-540).
* What the reference only sketches is real here: BuildEpoch (its worker is placeholders, :549-584) reads an Arrow
  vector file (format/storage.cpp layout), trains on a prefix of at most 100 000 rows (bench/benchmark.cpp:69),
  adds everything and activates the epoch; and the request coalescer (query_service.h:26-27: 64 queries / 2 ms,
  the queue nothing ever feeds, query_service.cpp:267-285) batches concurrent Search calls into one GPU search.
"""
import threading
import time
from concurrent import futures

import grpc
import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, empty_pb2, message_factory

_F = descriptor_pb2.FieldDescriptorProto


def _file_descriptor():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "vdb.proto", "vdb", "proto3"
    fd.dependency.append("google/protobuf/empty.proto")

    def msg(name, fields):
        m = fd.message_type.add()
        m.name = name
        for fname, num, ftype, label, tname in fields:
            f = m.field.add()
            f.name, f.number, f.type, f.label = fname, num, ftype, label
            if tname:
                f.type_name = tname
    O, R = _F.LABEL_OPTIONAL, _F.LABEL_REPEATED
    msg("Vector", [("id", 1, _F.TYPE_UINT64, O, None), ("values", 2, _F.TYPE_FLOAT, R, None)])
    msg("SearchRequest", [("queries", 1, _F.TYPE_MESSAGE, R, ".vdb.Vector"), ("topk", 2, _F.TYPE_INT32, O, None),
                          ("nprobe", 3, _F.TYPE_INT32, O, None), ("index", 4, _F.TYPE_STRING, O, None),
                          ("metric", 5, _F.TYPE_STRING, O, None), ("rerank_exact", 6, _F.TYPE_BOOL, O, None)])
    msg("Neighbor", [("id", 1, _F.TYPE_UINT64, O, None), ("distance", 2, _F.TYPE_FLOAT, O, None)])
    msg("SearchResult", [("neighbors", 1, _F.TYPE_MESSAGE, R, ".vdb.Neighbor")])
    msg("SearchResponse", [("results", 1, _F.TYPE_MESSAGE, R, ".vdb.SearchResult")])
    msg("WarmupRequest", [("index", 1, _F.TYPE_STRING, O, None), ("lists", 2, _F.TYPE_INT32, R, None)])
    msg("LoadIndexRequest", [("index", 1, _F.TYPE_STRING, O, None), ("epoch", 2, _F.TYPE_STRING, O, None)])
    msg("CreateIndexRequest", [("name", 1, _F.TYPE_STRING, O, None), ("dimension", 2, _F.TYPE_INT32, O, None),
                               ("metric", 3, _F.TYPE_STRING, O, None), ("nlist", 4, _F.TYPE_INT32, O, None),
                               ("m", 5, _F.TYPE_INT32, O, None), ("nbits", 6, _F.TYPE_INT32, O, None)])
    msg("BuildEpochRequest", [("index", 1, _F.TYPE_STRING, O, None), ("source_path", 2, _F.TYPE_STRING, O, None)])
    msg("ActivateEpochRequest", [("index", 1, _F.TYPE_STRING, O, None), ("epoch", 2, _F.TYPE_STRING, O, None)])
    msg("StatsRequest", [("index", 1, _F.TYPE_STRING, O, None)])
    msg("StatsResponse", [("total_vectors", 1, _F.TYPE_UINT64, O, None), ("indexed_vectors", 2, _F.TYPE_UINT64, O, None),
                          ("current_epoch", 3, _F.TYPE_STRING, O, None), ("gpu_memory_used", 4, _F.TYPE_FLOAT, O, None),
                          ("nvme_usage", 5, _F.TYPE_FLOAT, O, None)])
    return fd


_pool = descriptor_pool.Default()
try:
    _pool.FindFileByName("vdb.proto")
except KeyError:
    _pool.Add(_file_descriptor())


def _cls(name):
    return message_factory.GetMessageClass(_pool.FindMessageTypeByName("vdb." + name))


Vector, SearchRequest, Neighbor, SearchResult, SearchResponse = (_cls(n) for n in (
    "Vector", "SearchRequest", "Neighbor", "SearchResult", "SearchResponse"))
WarmupRequest, LoadIndexRequest, CreateIndexRequest, BuildEpochRequest, ActivateEpochRequest, StatsRequest, \
    StatsResponse = (_cls(n) for n in ("WarmupRequest", "LoadIndexRequest", "CreateIndexRequest", "BuildEpochRequest",
                                      "ActivateEpochRequest", "StatsRequest", "StatsResponse"))
Empty = empty_pb2.Empty
ID_PAD = 0xFFFFFFFFFFFFFFFF


class RequestCoalescer:
    """Batches concurrent Search calls on one (index, nprobe, k) into a single GPU search: a batch closes when it
    holds `batch_size` queries or `window_ms` after its first request (query_service.h:26-27)."""

    def __init__(self, search_fn, batch_size=64, window_ms=2.0):
        self.search_fn, self.batch_size, self.window = search_fn, batch_size, window_ms / 1e3
        self.lock = threading.Condition()
        self.pending = {}  # key -> list of [queries, event, result]
        self.deadline = {}
        self.stop = False
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def submit(self, key, queries):
        slot = [queries, threading.Event(), None]
        with self.lock:
            self.pending.setdefault(key, []).append(slot)
            self.deadline.setdefault(key, time.monotonic() + self.window)
            self.lock.notify()
        slot[1].wait()
        if isinstance(slot[2], Exception):
            raise slot[2]
        return slot[2]

    def _run(self):
        while True:
            with self.lock:
                while not self.stop:
                    now = time.monotonic()
                    ready = [k for k, v in self.pending.items()
                             if sum(s[0].shape[0] for s in v) >= self.batch_size or self.deadline[k] <= now]
                    if ready:
                        break
                    timeout = min((d - now for d in self.deadline.values()), default=None)
                    self.lock.wait(timeout)
                if self.stop:
                    return
                key = ready[0]
                slots = self.pending.pop(key)
                self.deadline.pop(key)
            try:
                q = np.concatenate([s[0] for s in slots], axis=0)
                D, I = self.search_fn(key, q)
                lo = 0
                for s in slots:
                    n = s[0].shape[0]
                    s[2] = (D[lo:lo + n], I[lo:lo + n])
                    lo += n
            except Exception as e:  # noqa: BLE001 -- handed to every waiter, each maps it to INTERNAL
                for s in slots:
                    s[2] = e
            for s in slots:
                s[1].set()

    def close(self):
        with self.lock:
            self.stop = True
            self.lock.notify()


class VdbServicer:
    """data_dir: where BuildEpoch persists epochs -- <data_dir>/<index>/<epoch id>/ in the reference's epoch layout
    (format/storage.cpp:318-348: manifest.json + Arrow IPC files); ActivateEpoch / LoadIndex load such a directory
    into a fresh HBM index and swap it in (server/query_service.cpp:515-519, 232-257).  devices: more than one entry
    makes every index a single-process sharded one (one list shard per GPU)."""

    def __init__(self, pkg, device=0, coalesce=True, batch_size=64, window_ms=2.0, data_dir=None, devices=()):
        self.pkg, self.device, self.devices, self.data_dir = pkg, device, tuple(devices), data_dir
        self.lock = threading.RLock()
        self.specs = {}    # name -> dict(dimension, metric, nlist)
        self.indices = {}  # name -> (IVFFlatIndex, epoch id)
        self.epochs = {}   # name -> {epoch id: directory or None (not persisted)}
        self.coalescer = RequestCoalescer(self._batched_search, batch_size, window_ms) if coalesce else None
        self.searches = 0
        self.queries = 0
        self.per_index = {}  # name -> [search count, latencies]
        self.started = time.monotonic()
        self.latencies_ms = []

    # ---- QueryService --------------------------------------------------------------------------------
    def Search(self, req, ctx):
        if len(req.queries) == 0:
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "No queries provided")
        if req.topk <= 0 or req.topk > 1000:
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "Invalid topk value")
        if not req.index:
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "Index name required")
        with self.lock:
            known = req.index in self.specs
            entry = self.indices.get(req.index)
        if not known:
            ctx.abort(grpc.StatusCode.NOT_FOUND, "Index not found: " + req.index)
        if entry is None:
            ctx.abort(grpc.StatusCode.FAILED_PRECONDITION, "Index has no active epoch: " + req.index)
        nprobe = req.nprobe if req.nprobe > 0 else 8
        dim = entry[0].get_dimension()
        q = np.empty((len(req.queries), dim), np.float32)
        for i, v in enumerate(req.queries):
            if len(v.values) != dim:
                ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "Query dimension mismatch")
            q[i] = v.values
        t0 = time.perf_counter()
        try:
            key = (req.index, nprobe, req.topk)
            D, I = self.coalescer.submit(key, q) if self.coalescer else self._batched_search(key, q)
        except Exception as e:  # noqa: BLE001
            ctx.abort(grpc.StatusCode.INTERNAL, "Search failed: " + str(e))
        with self.lock:
            ms = (time.perf_counter() - t0) * 1e3
            self.searches += 1
            self.queries += q.shape[0]
            self.latencies_ms = (self.latencies_ms + [ms])[-10000:]
            pi = self.per_index.setdefault(req.index, [0, []])
            pi[0] += 1
            pi[1] = (pi[1] + [ms])[-10000:]
        resp = SearchResponse()
        for qi in range(q.shape[0]):
            r = resp.results.add()
            for j in range(req.topk):
                if int(I[qi, j]) != ID_PAD:
                    nb = r.neighbors.add()
                    nb.id, nb.distance = int(I[qi, j]), float(D[qi, j])
        return resp

    def _batched_search(self, key, q):
        name, nprobe, k = key
        with self.lock:
            ix = self.indices[name][0]
        return ix.search(q, nprobe, k)

    def Warmup(self, req, ctx):
        with self.lock:
            if req.index not in self.specs:
                ctx.abort(grpc.StatusCode.NOT_FOUND, "Index not found: " + req.index)
            entry = self.indices.get(req.index)
        if entry is not None:
            try:
                entry[0].warmup_lists(list(req.lists)) if len(req.lists) else entry[0].warmup_all()
            except ValueError as e:
                ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, str(e))
        return Empty()

    def _new_index(self, spec):
        return self.pkg.IVFFlatIndex(self.pkg.Config(dimension=spec["dimension"], nlist=spec["nlist"],
                                                     metric=self.pkg.Metric(spec["metric"]), device=self.device,
                                                     devices=self.devices))

    def _load_index_internal(self, name, epoch, ctx):
        """load_index_internal (query_service.cpp:232-257): epoch directory -> new index -> swap under the lock"""
        from . import storage
        with self.lock:
            spec = self.specs.get(name)
            active = self.indices.get(name)
            known = self.epochs.get(name, {})
        if spec is None:
            ctx.abort(grpc.StatusCode.NOT_FOUND, "Index not found: " + name)
        if active is not None and (not epoch or active[1] == epoch):
            return Empty()  # already serving that epoch
        if epoch not in known:
            ctx.abort(grpc.StatusCode.NOT_FOUND, "Epoch not found: " + epoch)
        if known[epoch] is None:
            ctx.abort(grpc.StatusCode.FAILED_PRECONDITION, "Epoch was not persisted (server has no data_dir): " + epoch)
        try:
            ix = self._new_index(spec)
            storage.load_epoch(ix, known[epoch])
        except Exception as e:  # noqa: BLE001
            ctx.abort(grpc.StatusCode.INTERNAL, "Load failed: " + str(e))
        with self.lock:
            old = self.indices.get(name)
            self.indices[name] = (ix, epoch)
        del old  # the previous epoch's HBM is released once the searches holding it have returned
        return Empty()

    def LoadIndex(self, req, ctx):
        return self._load_index_internal(req.index, req.epoch, ctx)

    # ---- AdminService --------------------------------------------------------------------------------
    def CreateIndex(self, req, ctx):
        if not req.name or req.dimension <= 0 or req.nlist <= 0:
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "name, dimension and nlist are required")
        metric = {"L2": 0, "": 0, "InnerProduct": 1}.get(req.metric)
        if metric is None:
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "Unsupported metric: " + req.metric)
        with self.lock:
            if req.name in self.specs:
                ctx.abort(grpc.StatusCode.ALREADY_EXISTS, "Index already exists: " + req.name)
            self.specs[req.name] = dict(dimension=req.dimension, nlist=req.nlist, metric=metric)
        return Empty()

    def BuildEpoch(self, req, ctx):
        from . import storage
        with self.lock:
            spec = self.specs.get(req.index)
        if spec is None:
            ctx.abort(grpc.StatusCode.NOT_FOUND, "Index not found: " + req.index)
        try:
            parts = storage.read_vectors(req.source_path)
        except Exception as e:  # noqa: BLE001
            ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, f"cannot read {req.source_path}: {e}")
        try:
            ix = self._new_index(spec)
            first = parts[0][1]
            if first.shape[1] != spec["dimension"]:
                ctx.abort(grpc.StatusCode.INVALID_ARGUMENT, "source dimension mismatch")
            ix.train(first[: min(100_000, first.shape[0])])
            for ids, vecs, _keep in parts:
                ix.add(vecs, ids)
        except grpc.RpcError:
            raise
        except Exception as e:  # noqa: BLE001
            ctx.abort(grpc.StatusCode.INTERNAL, "Build failed: " + str(e))
        # build_index_worker (query_service.cpp:549-584): ... -> save epoch -> activate
        with self.lock:
            epoch = f"epoch_{int(time.time())}_{len(self.epochs.get(req.index, {}))}"
        path = None
        if self.data_dir:
            import os
            path = os.path.join(self.data_dir, req.index, epoch)
            try:
                os.makedirs(os.path.dirname(path), exist_ok=True)
                storage.save_epoch(ix, path, req.index, epoch)
            except Exception as e:  # noqa: BLE001
                ctx.abort(grpc.StatusCode.INTERNAL, "Build failed while writing the epoch: " + str(e))
        with self.lock:
            self.epochs.setdefault(req.index, {})[epoch] = path
            self.indices[req.index] = (ix, epoch)
        return Empty()

    def ActivateEpoch(self, req, ctx):
        """AdminServiceImpl::ActivateEpoch = load_index_internal(index, epoch) (query_service.cpp:515-519)"""
        return self._load_index_internal(req.index, req.epoch, ctx)

    def GetStats(self, req, ctx):
        with self.lock:
            if req.index not in self.specs:
                ctx.abort(grpc.StatusCode.NOT_FOUND, "Index not found: " + req.index)
            entry = self.indices.get(req.index)
        out = StatsResponse()
        if entry is not None:
            st = entry[0].stats()
            out.total_vectors = out.indexed_vectors = st.total_vectors
            out.current_epoch = entry[1]
            out.gpu_memory_used = st.gpu_memory_bytes / float(1 << 30)
        return out

    def metrics_text(self):
        """MetricsCollector::prometheus_format (query_service.cpp:748-780): the reference's four series, same names,
        HELP / TYPE lines and labels, fed by real counters."""
        with self.lock:
            per = {k: (v[0], sorted(v[1])) for k, v in self.per_index.items()}
            stats = {name: e[0].stats() for name, e in self.indices.items()}
            mem = sum(st.gpu_memory_bytes for st in stats.values())
            elapsed = max(time.monotonic() - self.started, 1e-9)
            total = self.searches
        q = lambda lat, p: lat[min(len(lat) - 1, int(p * len(lat)))] if lat else 0.0  # noqa: E731
        out = ["# HELP vdb_search_duration_milliseconds Search latency in milliseconds",
               "# TYPE vdb_search_duration_milliseconds histogram"]
        for name, (_n, lat) in sorted(per.items()):
            out += [f'vdb_search_duration_milliseconds{{index="{name}",quantile="{p}"}} {q(lat, p):.3f}'
                    for p in (0.5, 0.95, 0.99)]
        out += ["# HELP vdb_searches_total Total number of searches", "# TYPE vdb_searches_total counter"]
        out += [f'vdb_searches_total{{index="{name}"}} {n}' for name, (n, _l) in sorted(per.items())]
        out += ["# HELP vdb_gpu_memory_bytes GPU memory usage in bytes", "# TYPE vdb_gpu_memory_bytes gauge",
                f"vdb_gpu_memory_bytes {mem}",
                "# HELP vdb_queries_per_second Current queries per second", "# TYPE vdb_queries_per_second gauge",
                f"vdb_queries_per_second {total / elapsed:.3f}",
                # (not in the reference) what the hot path is bound by: distinct list bytes streamed from HBM
                "# HELP vdb_hbm_bytes_scanned_total Inverted-list bytes streamed from HBM by searches",
                "# TYPE vdb_hbm_bytes_scanned_total counter"]
        out += [f'vdb_hbm_bytes_scanned_total{{index="{name}"}} {st.scanned_bytes}' for name, st in sorted(stats.items())]
        return "\n".join(out) + "\n"

    def serve_metrics(self, port=0, host="127.0.0.1"):
        """GET /metrics over HTTP (the reference's metrics endpoint); returns (http server, bound port)"""
        import http.server
        servicer = self

        class H(http.server.BaseHTTPRequestHandler):
            def do_GET(self):  # noqa: N802
                body = servicer.metrics_text().encode() if self.path.startswith("/metrics") else b"not found\n"
                self.send_response(200 if self.path.startswith("/metrics") else 404)
                self.send_header("Content-Type", "text/plain; version=0.0.4")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def log_message(self, *args):
                pass

        httpd = http.server.ThreadingHTTPServer((host, port), H)
        threading.Thread(target=httpd.serve_forever, daemon=True).start()
        return httpd, httpd.server_address[1]


def _handlers(servicer):
    def unary(fn, req_cls):
        return grpc.unary_unary_rpc_method_handler(fn, request_deserializer=req_cls.FromString,
                                                   response_serializer=lambda m: m.SerializeToString())
    q = grpc.method_handlers_generic_handler("vdb.QueryService", {
        "Search": unary(servicer.Search, SearchRequest), "Warmup": unary(servicer.Warmup, WarmupRequest),
        "LoadIndex": unary(servicer.LoadIndex, LoadIndexRequest)})
    a = grpc.method_handlers_generic_handler("vdb.AdminService", {
        "CreateIndex": unary(servicer.CreateIndex, CreateIndexRequest),
        "BuildEpoch": unary(servicer.BuildEpoch, BuildEpochRequest),
        "ActivateEpoch": unary(servicer.ActivateEpoch, ActivateEpochRequest),
        "GetStats": unary(servicer.GetStats, StatsRequest)})
    return q, a


def serve(pkg, address="127.0.0.1:50051", device=0, max_workers=8, **kw):
    """Start the server (server/main.cpp:88-94: 100 MB messages, 2-8 pollers); returns (grpc server, servicer)."""
    servicer = VdbServicer(pkg, device=device, **kw)
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers),
                         options=[("grpc.max_receive_message_length", 100 << 20),
                                  ("grpc.max_send_message_length", 100 << 20)])
    server.add_generic_rpc_handlers(_handlers(servicer))
    port = server.add_insecure_port(address)
    server.start()
    server.bound_port = port
    return server, servicer


class Client:
    """Minimal client over the same dynamically built messages (what test/integration/*.cpp does with stubs)."""

    def __init__(self, target):
        self.ch = grpc.insecure_channel(target, options=[("grpc.max_receive_message_length", 100 << 20)])

        def m(path, req, resp):
            return self.ch.unary_unary(path, request_serializer=lambda x: x.SerializeToString(),
                                       response_deserializer=resp.FromString)
        self.Search = m("/vdb.QueryService/Search", SearchRequest, SearchResponse)
        self.Warmup = m("/vdb.QueryService/Warmup", WarmupRequest, Empty)
        self.LoadIndex = m("/vdb.QueryService/LoadIndex", LoadIndexRequest, Empty)
        self.CreateIndex = m("/vdb.AdminService/CreateIndex", CreateIndexRequest, Empty)
        self.BuildEpoch = m("/vdb.AdminService/BuildEpoch", BuildEpochRequest, Empty)
        self.ActivateEpoch = m("/vdb.AdminService/ActivateEpoch", ActivateEpochRequest, Empty)
        self.GetStats = m("/vdb.AdminService/GetStats", StatsRequest, StatsResponse)

    def close(self):
        self.ch.close()
