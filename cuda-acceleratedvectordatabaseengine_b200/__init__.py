"""B200-native IVF-Flat hot path: Python host binding over the C ABI.

Mirrors the reference's ``vdb::IVFFlatIndex`` surface (engine/ivf_flat_index.h:14-67)
-- ``Config`` / ``SearchParams`` field names and defaults, ``train``, ``add``,
``search``, ``get_gpu_memory_usage``, ``get_total_vectors`` -- plus the methods
the reference's server calls but its engine lacks (``get_dimension``,
``warmup_lists``, ``warmup_all``; server/query_service.cpp:112,191,195).

Everything is computed by ``libvdb_b200.so`` (hand-written sm_100a CUDA behind
``include/vdb_b200.h``).  There is NO CPU path: importing works without a GPU
(so the symbol table can be checked), but creating an index raises.

The directory name contains '-', so load it with
``importlib.import_module("cuda-acceleratedvectordatabaseengine_b200")``.
"""
import ctypes as C
import enum
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvdb_b200.so")

FLT_MAX = np.float32(3.4028234663852886e38)
ID_PAD = np.uint64(0xFFFFFFFFFFFFFFFF)

VDB_OK, VDB_INVALID_ARGUMENT, VDB_OUT_OF_MEMORY, VDB_CUDA_ERROR, VDB_NCCL_ERROR, VDB_NOT_TRAINED, VDB_INTERNAL = range(7)


class Metric(enum.IntEnum):
    """kernels.cuh:24-28"""
    L2 = 0
    InnerProduct = 1
    Cosine = 2


class TrainMode(enum.IntEnum):
    AUTO = 0
    EXACT = 1
    FAST = 2


class _Config(C.Structure):
    _fields_ = [("dimension", C.c_uint32), ("nlist", C.c_uint32), ("metric", C.c_int32), ("device", C.c_int32),
                ("max_gpu_memory", C.c_uint64), ("train_mode", C.c_int32), ("coarse_mode", C.c_int32),
                ("page_rows", C.c_uint32), ("shard_rank", C.c_uint32), ("shard_count", C.c_uint32),
                ("pipeline_depth", C.c_uint32), ("reserve_sms", C.c_uint32), ("scan_mirror", C.c_uint32),
                ("reserved", C.c_uint32 * 2)]


class _Stats(C.Structure):
    _fields_ = [("total_vectors", C.c_uint64), ("local_vectors", C.c_uint64), ("gpu_memory_bytes", C.c_uint64),
                ("pages", C.c_uint64), ("dimension", C.c_uint32), ("nlist", C.c_uint32),
                ("row_stride", C.c_uint32), ("page_rows", C.c_uint32), ("trained", C.c_int32),
                ("metric", C.c_int32), ("scanned_bytes", C.c_uint64)]


class _SearchStats(C.Structure):
    _fields_ = [("algorithmic_rows", C.c_uint64), ("unique_rows", C.c_uint64), ("scan_items", C.c_uint64),
                ("bytes_per_row", C.c_uint64), ("scan_ctas", C.c_uint64),
                ("streamed_bytes_per_row", C.c_uint64), ("rescored_pairs", C.c_uint64)]


# every symbol include/vdb_b200.h declares: name -> (restype, argtypes)
_vp, _u32, _u64, _i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
ABI = {
    "vdb_last_error_string": (C.c_char_p, []),
    "vdb_status_string": (C.c_char_p, [_i32]),
    "vdb_version": (_i32, []),
    "vdb_config_default": (None, [C.POINTER(_Config)]),
    "vdb_index_create": (_i32, [C.POINTER(_Config), C.POINTER(_vp)]),
    "vdb_index_create_sharded": (_i32, [C.POINTER(_Config), C.POINTER(_i32), _i32, C.POINTER(_vp)]),
    "vdb_index_destroy": (_i32, [_vp]),
    "vdb_index_train": (_i32, [_vp, _vp, _u64]),
    "vdb_index_add": (_i32, [_vp, _vp, _vp, _u64]),
    "vdb_index_add_assigned": (_i32, [_vp, _vp, _vp, _vp, _u64, _u64]),
    "vdb_index_search": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    "vdb_index_search_async": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "vdb_index_search_submit": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, C.POINTER(_u64)]),
    "vdb_index_search_wait": (_i32, [_vp, _u64]),
    "vdb_index_search_wait_stream": (_i32, [_vp, _u64, _vp]),
    "vdb_index_set_arena": (_i32, [_vp, _vp, _i32]),
    "vdb_index_reserve_search": (_i32, [_vp, _u32, _u32, _u32]),
    "vdb_index_attach_exchange": (_i32, [_vp, _vp]),
    "vdb_index_select_nprobe": (_i32, [_vp, _vp, _u32, _u32, _vp]),
    "vdb_index_assign": (_i32, [_vp, _vp, _u64, _vp]),
    "vdb_index_get_centroids": (_i32, [_vp, _vp]),
    "vdb_index_set_centroids": (_i32, [_vp, _vp]),
    "vdb_index_get_owners": (_i32, [_vp, _vp]),
    "vdb_index_set_owners": (_i32, [_vp, _vp]),
    "vdb_index_list_sizes": (_i32, [_vp, _vp]),
    "vdb_index_list_ids": (_i32, [_vp, _u32, _vp]),
    "vdb_index_list_vectors": (_i32, [_vp, _u32, _vp]),
    "vdb_index_append_list": (_i32, [_vp, _u32, _vp, _vp, _u64]),
    "vdb_index_finish_load": (_i32, [_vp]),
    "vdb_index_balance_owners": (_i32, [_vp, _vp]),
    "vdb_index_stats": (_i32, [_vp, C.POINTER(_Stats)]),
    "vdb_index_last_search_stats": (_i32, [_vp, C.POINTER(_SearchStats)]),
    "vdb_index_warmup": (_i32, [_vp, _vp, _u32]),
    "vdb_index_set_profiling": (_i32, [_vp, _i32]),
    "vdb_index_read_profile": (_i32, [_vp, _vp, _vp]),
    "vdb_bruteforce_search": (_i32, [_vp, _vp, _vp, _u64, _u32, _u32, _u32, _vp, _vp, _i32, _vp]),
    "vdb_kmeans_assign": (_i32, [_vp, _vp, _vp, _vp, _u64, _u32, _u32, _i32, _vp]),
    "vdb_kmeans_accumulate": (_i32, [_vp, _vp, _u64, _u32, _u32, _vp, _vp, _vp]),
    "vdb_kmeans_finalize": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp]),
    "vdb_merge_topk": (_i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    "vdb_exchange_create": (_i32, [_i32, _u32, _u32, _u32, _u32, C.POINTER(_vp)]),
    "vdb_exchange_handle": (_i32, [_vp, _vp]),
    "vdb_exchange_connect": (_i32, [_vp, _vp]),
    "vdb_exchange_merge_topk": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "vdb_exchange_publish": (_i32, [_vp, _vp, _vp, _u32, _u32, _vp]),
    "vdb_exchange_collect": (_i32, [_vp, _vp, _vp, _vp]),
    "vdb_exchange_connect_local": (_i32, [_vp, _u32, _u32]),
    "vdb_exchange_status": (_i32, [_vp]),
    "vdb_exchange_reset": (_i32, [_vp]),
    "vdb_exchange_destroy": (_i32, [_vp]),
    "vdb_arena_create": (_i32, [_i32, _u64, _u64, _i32, C.POINTER(_vp)]),
    "vdb_arena_destroy": (_i32, [_vp]),
    "vdb_arena_allocate_device": (_vp, [_vp, _u64]),
    "vdb_arena_free_device": (_i32, [_vp, _vp]),
    "vdb_arena_allocate_pinned": (_vp, [_vp, _u64]),
    "vdb_arena_free_pinned": (_i32, [_vp, _vp]),
    "vdb_arena_get_stream": (_vp, [_vp]),
    "vdb_arena_return_stream": (_i32, [_vp, _vp]),
    "vdb_arena_enqueue_transfer": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp]),
    "vdb_arena_enqueue_transfer_cb": (_i32, [_vp, _vp, _vp, _u64, _i32, _vp, _vp, _vp]),
    "vdb_arena_synchronize": (_i32, [_vp]),
    "vdb_arena_synchronize_stream": (_i32, [_vp, _vp]),
    "vdb_arena_stats": (_i32, [_vp, _vp]),
}

_LIB = None


def build(force=False):
    """Compile libvdb_b200.so for sm_100a in tree (nvcc cross-compiles without a GPU)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-j8"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    """The loaded C-ABI library.  Fails loudly when the CUDA extension is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is not built: run __graft_entry__.build() (there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in ABI.items():
            fn = getattr(l, name)  # AttributeError = header/library drift
            fn.restype, fn.argtypes = res, args
        _LIB = l
    return _LIB


class VdbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"{msg} [{status}]")
        self.status = status


def _check(st):
    if st == VDB_OK:
        return
    l = lib()
    msg = l.vdb_last_error_string().decode() or l.vdb_status_string(st).decode()
    if st == VDB_INVALID_ARGUMENT:
        raise ValueError(msg)  # std::invalid_argument in the reference's ctor
    if st == VDB_OUT_OF_MEMORY:
        raise MemoryError(msg)
    raise VdbError(st, msg)


def _ptr(x):
    """Device or host address of a numpy array / torch tensor / int."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return int(x)


def _is_torch(x):
    return hasattr(x, "data_ptr")


@dataclass
class Config:
    """IVFFlatIndex::Config, ivf_flat_index.h:16-22 (+ placement fields)."""
    dimension: int = 0
    nlist: int = 0
    metric: Metric = Metric.L2
    use_gpu: bool = True          # accepted either way and without effect: there is no CPU path, the GPU path
                                  # returns what the reference's CPU path returns (test/simple_test.cpp:114 sets False)
    max_gpu_memory: int = 0       # 0 = uncapped; the reference default of 8 GiB would not hold the headline index
    device: int = 0
    train_mode: TrainMode = TrainMode.AUTO
    coarse_mode: int = 0
    page_rows: int = 0
    shard_rank: int = 0
    shard_count: int = 1
    devices: tuple = ()           # more than one entry: ONE process, one list shard per device (create_sharded)
    pipeline_depth: int = 0       # searches in flight (search_submit), 0 = 4
    reserve_sms: int = 0          # SMs a pipelined scan leaves to the neighbouring batches' small kernels, 0 = 8
    scan_mirror: int = 0          # shadow of the lists for the tensor-core screen: 0 = auto, 1 = off, 2 = bf16, 3 = int8


@dataclass
class SearchParams:
    """IVFFlatIndex::SearchParams, ivf_flat_index.h:38-42."""
    nprobe: int = 10
    k: int = 10
    use_exact_rerank: bool = False


class IVFFlatIndex:
    def __init__(self, config, tm=None):
        l = lib()
        c = _Config()
        l.vdb_config_default(C.byref(c))
        c.dimension, c.nlist, c.metric, c.device = config.dimension, config.nlist, int(config.metric), config.device
        c.max_gpu_memory, c.train_mode, c.coarse_mode = config.max_gpu_memory, int(config.train_mode), config.coarse_mode
        c.page_rows, c.shard_rank, c.shard_count = config.page_rows, config.shard_rank, config.shard_count
        c.pipeline_depth, c.reserve_sms = config.pipeline_depth, config.reserve_sms
        c.scan_mirror = config.scan_mirror
        self.config = config
        self._tm = tm  # borrowed, like the reference's TransferManager*
        self._h = _vp()
        if len(config.devices) > 1:
            devs = (_i32 * len(config.devices))(*config.devices)
            _check(l.vdb_index_create_sharded(C.byref(c), devs, len(config.devices), C.byref(self._h)))
        else:
            if len(config.devices) == 1:
                c.device = config.devices[0]
            _check(l.vdb_index_create(C.byref(c), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().vdb_index_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- reference surface ---------------------------------------------------
    def train(self, vectors, n_vectors=None):
        v = self._rows(vectors)
        n = v.shape[0] if n_vectors is None else n_vectors
        _check(lib().vdb_index_train(self._h, _ptr(v), n))

    def add(self, vectors, ids=None, n_vectors=None):
        v = self._rows(vectors)
        n = v.shape[0] if n_vectors is None else n_vectors
        if ids is not None and not _is_torch(ids):
            ids = np.ascontiguousarray(ids, np.uint64)
        _check(lib().vdb_index_add(self._h, _ptr(v), _ptr(ids), n))

    def add_assigned(self, vectors, ids, assignments, global_n):
        """device tensors: rows whose lists are already known (data-parallel add of a sharded index)"""
        n = vectors.shape[0]
        _check(lib().vdb_index_add_assigned(self._h, _ptr(vectors) if n else None, _ptr(ids) if n else None,
                                            _ptr(assignments) if n else None, n, global_n))

    def assign_device(self, vectors):
        """CUDA tensor [n][dim] -> CUDA int32 tensor [n] of list ids"""
        import torch
        v = self._rows(vectors)
        out = torch.empty(v.shape[0], dtype=torch.int32, device=v.device)
        if v.shape[0]:
            _check(lib().vdb_index_assign(self._h, _ptr(v), v.shape[0], _ptr(out)))
        return out

    def search(self, queries, params_or_nprobe=None, k=None, distances=None, indices=None):
        """search(queries, SearchParams) -> (distances [nq][k] f32, indices [nq][k] u64).
        numpy in -> numpy out; CUDA torch tensors in -> CUDA torch tensors out (int64 view of the u64 ids)."""
        if isinstance(params_or_nprobe, SearchParams):
            nprobe, k = params_or_nprobe.nprobe, params_or_nprobe.k
        else:
            nprobe = SearchParams().nprobe if params_or_nprobe is None else params_or_nprobe
            k = SearchParams().k if k is None else k
        q = self._rows(queries)
        nq = q.shape[0]
        if _is_torch(q):
            import torch
            D = torch.empty((nq, k), dtype=torch.float32, device=q.device) if distances is None else distances
            I = torch.empty((nq, k), dtype=torch.int64, device=q.device) if indices is None else indices
        else:
            D = np.empty((nq, k), np.float32) if distances is None else distances
            I = np.empty((nq, k), np.uint64) if indices is None else indices
        _check(lib().vdb_index_search(self._h, _ptr(q), nq, nprobe, k, _ptr(D), _ptr(I)))
        return D, I

    def search_async(self, queries, nprobe, k, distances, indices, stream=0):
        """Device tensors only, enqueued on `stream` (int handle), no host sync."""
        self._check_device_args(queries, distances, indices, k)
        _check(lib().vdb_index_search_async(self._h, _ptr(queries), queries.shape[0], nprobe, k, _ptr(distances),
                                            _ptr(indices), stream))

    def _check_device_args(self, queries, distances, indices, k):
        import torch
        nq = queries.shape[0]
        if not (queries.is_cuda and queries.dtype == torch.float32 and queries.is_contiguous() and
                queries.shape[-1] == self.config.dimension):
            raise ValueError("queries must be a contiguous float32 CUDA tensor [nq][dimension]")
        if not (distances.is_cuda and distances.dtype == torch.float32 and distances.is_contiguous() and
                distances.numel() == nq * k):
            raise ValueError("distances must be a contiguous float32 CUDA tensor [nq][k]")
        if not (indices.is_cuda and indices.dtype in (torch.int64, torch.uint64) and indices.is_contiguous() and
                indices.numel() == nq * k):
            raise ValueError("indices must be a contiguous int64 CUDA tensor [nq][k]")

    def search_submit(self, queries, nprobe, k, distances, indices):
        """Pipelined search (vdb_index_search_submit): enqueue one batch on the index's own streams and return a
        ticket; up to pipeline_depth batches overlap.  queries / outputs: float32 numpy arrays or torch tensors
        (host, pinned or CUDA), kept alive and untouched by the caller until search_wait(ticket)."""
        if isinstance(queries, np.ndarray):
            assert queries.dtype == np.float32 and queries.flags.c_contiguous
        nq = queries.shape[0]
        t = _u64()
        _check(lib().vdb_index_search_submit(self._h, _ptr(queries), nq, nprobe, k, _ptr(distances), _ptr(indices),
                                             C.byref(t)))
        return t.value

    def search_wait(self, ticket):
        _check(lib().vdb_index_search_wait(self._h, ticket))

    def search_wait_stream(self, ticket, stream=0):
        """make `stream` (int handle) wait for the ticket instead of the host"""
        _check(lib().vdb_index_search_wait_stream(self._h, ticket, stream))

    def reserve_search(self, max_nq, max_nprobe, max_k):
        _check(lib().vdb_index_reserve_search(self._h, max_nq, max_nprobe, max_k))

    def attach_exchange(self, exchange_handle):
        """one process per GPU: searches become collective and return the merged result of all shards"""
        _check(lib().vdb_index_attach_exchange(self._h, exchange_handle))

    def get_gpu_memory_usage(self):
        return self.stats().gpu_memory_bytes

    def get_total_vectors(self):
        return self.stats().total_vectors

    def get_dimension(self):
        return self.config.dimension

    def warmup_lists(self, list_ids):
        a = np.ascontiguousarray(list_ids, np.uint32)
        _check(lib().vdb_index_warmup(self._h, _ptr(a), a.size))

    def warmup_all(self):
        self.warmup_lists(np.arange(self.config.nlist, dtype=np.uint32))

    # -- hooks ---------------------------------------------------------------
    def stats(self):
        s = _Stats()
        _check(lib().vdb_index_stats(self._h, C.byref(s)))
        return s

    def last_search_stats(self):
        s = _SearchStats()
        _check(lib().vdb_index_last_search_stats(self._h, C.byref(s)))
        return s

    def set_profiling(self, enable=True):
        _check(lib().vdb_index_set_profiling(self._h, int(enable)))

    def read_profile(self):
        """-> dict of summed ms (coarse, group, scan, merge, collect), the scan streams' busy span, and the number
        of searches covered"""
        ms = (C.c_float * 8)()
        n = C.c_uint32()
        _check(lib().vdb_index_read_profile(self._h, ms, C.byref(n)))
        return {"coarse_ms": ms[0], "group_ms": ms[1], "scan_ms": ms[2], "merge_ms": ms[3], "collect_ms": ms[4],
                "scan_span_ms": ms[5], "searches": n.value}

    @property
    def centroids(self):
        out = np.empty((self.config.nlist, self.config.dimension), np.float32)
        _check(lib().vdb_index_get_centroids(self._h, _ptr(out)))
        return out

    @centroids.setter
    def centroids(self, c):
        c = np.ascontiguousarray(c, np.float32)
        assert c.shape == (self.config.nlist, self.config.dimension)
        _check(lib().vdb_index_set_centroids(self._h, _ptr(c)))

    def set_centroids_device(self, t):
        assert tuple(t.shape) == (self.config.nlist, self.config.dimension) and t.is_contiguous()
        _check(lib().vdb_index_set_centroids(self._h, _ptr(t)))

    def owners(self):
        """rank that holds each inverted list (all zeros on an unsharded index)"""
        out = np.empty(self.config.nlist, np.uint8)
        _check(lib().vdb_index_get_owners(self._h, _ptr(out)))
        return out

    def set_owners(self, owners):
        o = np.ascontiguousarray(owners, np.uint8)
        assert o.shape == (self.config.nlist,)
        _check(lib().vdb_index_set_owners(self._h, _ptr(o)))

    def list_sizes(self):
        out = np.empty(self.config.nlist, np.uint64)
        _check(lib().vdb_index_list_sizes(self._h, _ptr(out)))
        return out

    def list_ids(self, l):
        n = int(self.list_sizes()[l])
        out = np.empty(n, np.uint64)
        if n:
            _check(lib().vdb_index_list_ids(self._h, l, _ptr(out)))
        return out

    def list_vectors(self, l):
        n = int(self.list_sizes()[l])
        out = np.empty((n, self.config.dimension), np.float32)
        if n:
            _check(lib().vdb_index_list_vectors(self._h, l, _ptr(out)))
        return out

    def select_nprobe(self, queries, nprobe):
        q = self._rows(queries)
        np_eff = min(nprobe, self.config.nlist)
        out = np.empty((q.shape[0], np_eff), np.uint32)
        _check(lib().vdb_index_select_nprobe(self._h, _ptr(q), q.shape[0], nprobe, _ptr(out)))
        return out

    def assign(self, vectors):
        v = self._rows(vectors)
        out = np.empty(v.shape[0], np.uint32)
        _check(lib().vdb_index_assign(self._h, _ptr(v), v.shape[0], _ptr(out)))
        return out

    def _rows(self, x):
        d = self.config.dimension
        if _is_torch(x):
            import torch
            assert x.dtype == torch.float32
            x = x.reshape(-1, d).contiguous()
            if x.is_cuda:
                # the library works on its own (non-blocking) streams, which do not wait for torch's: hand it rows
                # whose producing kernels have finished
                torch.cuda.current_stream(x.device).synchronize()
            return x
        return np.ascontiguousarray(x, np.float32).reshape(-1, d)


def bruteforce_search(database, queries, k, metric=Metric.L2, ids=None, stream=0):
    """kernels::launch_bruteforce_search<float> (kernels.cu:13-43): exact top-k, any k <= 2048."""
    tor = _is_torch(database)
    if tor:
        import torch
        db, q = database.contiguous(), queries.contiguous()
        D = torch.empty((q.shape[0], k), dtype=torch.float32, device=q.device)
        I = torch.empty((q.shape[0], k), dtype=torch.int64, device=q.device)
    else:
        db = np.ascontiguousarray(database, np.float32)
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, db.shape[1])
        D = np.empty((q.shape[0], k), np.float32)
        I = np.empty((q.shape[0], k), np.uint64)
        if ids is not None:
            ids = np.ascontiguousarray(ids, np.uint64)
    _check(lib().vdb_bruteforce_search(_ptr(db), _ptr(q), _ptr(ids), db.shape[0], q.shape[0], db.shape[1], k,
                                       _ptr(D), _ptr(I), int(metric), stream))
    return D, I


def kmeans_assign(vectors, centroids, metric=Metric.L2, want_distances=False, stream=0):
    """kernels::launch_kmeans_assign<float> (kernels.cu:80-92)."""
    v = np.ascontiguousarray(vectors, np.float32)
    c = np.ascontiguousarray(centroids, np.float32)
    a = np.empty(v.shape[0], np.uint32)
    d = np.empty(v.shape[0], np.float32) if want_distances else None
    _check(lib().vdb_kmeans_assign(_ptr(v), _ptr(c), _ptr(a), _ptr(d), v.shape[0], c.shape[0], v.shape[1],
                                   int(metric), stream))
    return (a, d) if want_distances else a


def merge_topk(dist_parts, id_parts, stream=0):
    """merge_results across shards: [parts][nq][k] device tensors -> ([nq][k], [nq][k])."""
    import torch
    parts, nq, k = dist_parts.shape
    D = torch.empty((nq, k), dtype=torch.float32, device=dist_parts.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=dist_parts.device)
    _check(lib().vdb_merge_topk(_ptr(dist_parts), _ptr(id_parts), parts, nq, k, _ptr(D), _ptr(I), stream))
    return D, I
