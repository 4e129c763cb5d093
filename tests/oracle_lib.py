"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when built,
the reference's own CPU path (oracle/_ref/libvdbref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under the product
package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libvdbref.so")

METRIC_L2, METRIC_IP, METRIC_COSINE = 0, 1, 2
FLT_MAX = np.float32(3.4028234663852886e38)
ID_PAD = np.uint64(0xFFFFFFFFFFFFFFFF)

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build_port():
    if not os.path.exists(PORT_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return PORT_SO


def build_ref(reference="/root/reference"):
    """Compile the reference's CPU path in place; returns None when the
    reference checkout is absent (the GPU box) and no prebuilt .so travelled."""
    if os.path.exists(REF_SO):
        return REF_SO
    if not os.path.isdir(os.path.join(reference, "engine")):
        return None
    subprocess.check_call(["make", "-C", ORACLE_DIR, "ref", f"REF={reference}"], stdout=subprocess.DEVNULL)
    return REF_SO


def _load_port():
    lib = C.CDLL(build_port())
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_create.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_train.argtypes = [C.c_void_p, _f32p, C.c_uint64]
    lib.oracle_add.argtypes = [C.c_void_p, _f32p, _u64p, C.c_uint64]
    lib.oracle_assign.argtypes = [C.c_void_p, _f32p, C.c_uint64, _u32p]
    lib.oracle_load_assigned.argtypes = [C.c_void_p, _f32p, _u64p, _u32p, C.c_uint64]
    lib.oracle_select_nprobe.restype = C.c_uint32
    lib.oracle_select_nprobe.argtypes = [C.c_void_p, _f32p, C.c_uint32, _u32p]
    lib.oracle_search.argtypes = [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _u64p, C.c_int]
    lib.oracle_get_centroids.argtypes = [C.c_void_p, _f32p]
    lib.oracle_set_centroids.argtypes = [C.c_void_p, _f32p]
    lib.oracle_list_sizes.argtypes = [C.c_void_p, _u64p]
    lib.oracle_list_ids.argtypes = [C.c_void_p, C.c_uint32, _u64p]
    lib.oracle_total_vectors.restype = C.c_uint64
    lib.oracle_total_vectors.argtypes = [C.c_void_p]
    lib.oracle_flat_search.argtypes = [_f32p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, _f32p, C.c_uint32,
                                       C.c_uint32, _f32p, _u64p]
    lib.gen_gaussian.argtypes = [C.c_uint32, _f32p, C.c_uint64]
    lib.gen_clustered.argtypes = [C.c_uint32, _f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_float]
    return lib


def _load_ref():
    so = build_ref()
    if so is None:
        return None
    lib = C.CDLL(so)
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_uint32, C.c_uint32, C.c_int]
    lib.ref_destroy.argtypes = [C.c_void_p]
    lib.ref_train.argtypes = [C.c_void_p, _f32p, C.c_uint64]
    lib.ref_add.argtypes = [C.c_void_p, _f32p, _u64p, C.c_uint64]
    lib.ref_assign.argtypes = [C.c_void_p, _f32p, C.c_uint64, _u32p]
    lib.ref_load_assigned.argtypes = [C.c_void_p, _f32p, _u64p, _u32p, C.c_uint64]
    lib.ref_select_nprobe.argtypes = [C.c_void_p, _f32p, C.c_uint32, _u32p]
    lib.ref_search.argtypes = [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _u64p, C.c_int]
    lib.ref_search_batched.argtypes = [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, _f32p, _u64p]
    lib.ref_get_centroids.argtypes = [C.c_void_p, _f32p]
    lib.ref_set_centroids.argtypes = [C.c_void_p, _f32p]
    lib.ref_list_sizes.argtypes = [C.c_void_p, _u64p]
    lib.ref_list_ids.argtypes = [C.c_void_p, C.c_uint32, _u64p]
    lib.ref_total_vectors.restype = C.c_uint64
    lib.ref_total_vectors.argtypes = [C.c_void_p]
    return lib


_PORT = None
_REF = False


def port_lib():
    global _PORT
    if _PORT is None:
        _PORT = _load_port()
    return _PORT


def ref_lib():
    global _REF
    if _REF is False:
        _REF = _load_ref()
    return _REF


def gaussian(seed, n, dim):
    """n x dim fp32 from std::mt19937(seed) + std::normal_distribution<float>."""
    out = np.empty(n * dim, np.float32)
    port_lib().gen_gaussian(seed, out, out.size)
    return out.reshape(n, dim)


def clustered(seed, n, dim, n_centers=32, spread=0.05):
    out = np.empty(n * dim, np.float32)
    port_lib().gen_clustered(seed, out, n, dim, n_centers, spread)
    return out.reshape(n, dim)


class _Index:
    """Common surface over the port ('oracle_*') and the reference ('ref_*')."""

    def __init__(self, lib, prefix, dim, nlist, metric=METRIC_L2):
        self._lib, self._p = lib, prefix
        self.dim, self.nlist, self.metric = dim, nlist, metric
        self._h = self._f("create")(dim, nlist, metric)
        if not self._h:
            raise ValueError("Invalid configuration: dimension and nlist must be > 0")

    def _f(self, name):
        return getattr(self._lib, f"{self._p}_{name}")

    def close(self):
        if self._h:
            self._f("destroy")(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def train(self, x):
        x = np.ascontiguousarray(x, np.float32)
        self._f("train")(self._h, x, x.shape[0])

    def add(self, x, ids=None):
        x = np.ascontiguousarray(x, np.float32)
        if ids is None:
            ids = np.arange(x.shape[0], dtype=np.uint64)
        self._f("add")(self._h, x, np.ascontiguousarray(ids, np.uint64), x.shape[0])

    def load_assigned(self, x, ids, assign):
        """bench set-up: append rows with caller-provided list assignments (add() minus its assign step)"""
        x = np.ascontiguousarray(x, np.float32)
        self._f("load_assigned")(self._h, x, np.ascontiguousarray(ids, np.uint64),
                                 np.ascontiguousarray(assign, np.uint32), x.shape[0])

    def assign(self, x):
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty(x.shape[0], np.uint32)
        self._f("assign")(self._h, x, x.shape[0], out)
        return out

    def select_nprobe(self, q, nprobe):
        nprobe = min(nprobe, self.nlist)
        out = np.empty(nprobe, np.uint32)
        self._f("select_nprobe")(self._h, np.ascontiguousarray(q, np.float32), nprobe, out)
        return out

    def search(self, q, nprobe, k, nthreads=1):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.dim)
        D = np.empty((q.shape[0], k), np.float32)
        I = np.empty((q.shape[0], k), np.uint64)
        self._f("search")(self._h, q, q.shape[0], nprobe, k, D, I, nthreads)
        return D, I

    @property
    def centroids(self):
        out = np.empty((self.nlist, self.dim), np.float32)
        self._f("get_centroids")(self._h, out)
        return out

    @centroids.setter
    def centroids(self, c):
        c = np.ascontiguousarray(c, np.float32)
        assert c.shape == (self.nlist, self.dim)
        self._f("set_centroids")(self._h, c)

    def list_sizes(self):
        out = np.empty(self.nlist, np.uint64)
        self._f("list_sizes")(self._h, out)
        return out

    def list_ids(self, l):
        n = int(self.list_sizes()[l])
        out = np.empty(n, np.uint64)
        if n:
            self._f("list_ids")(self._h, l, out)
        return out

    @property
    def ntotal(self):
        return int(self._f("total_vectors")(self._h))


class OracleIndex(_Index):
    """The C restatement (oracle/ivf_oracle.c)."""

    def __init__(self, dim, nlist, metric=METRIC_L2):
        super().__init__(port_lib(), "oracle", dim, nlist, metric)


class RefIndex(_Index):
    """The reference's own unmodified CPU path (oracle/_ref/libvdbref.so)."""

    def __init__(self, dim, nlist, metric=METRIC_L2):
        lib = ref_lib()
        if lib is None:
            raise RuntimeError("oracle/_ref/libvdbref.so not built and /root/reference absent")
        super().__init__(lib, "ref", dim, nlist, metric)

    def search_batched(self, q, nprobe, k):
        q = np.ascontiguousarray(q, np.float32).reshape(-1, self.dim)
        D = np.empty((q.shape[0], k), np.float32)
        I = np.empty((q.shape[0], k), np.uint64)
        self._lib.ref_search_batched(self._h, q, q.shape[0], nprobe, k, D, I)
        return D, I


def flat_search(db, q, k, metric=METRIC_L2, ids=None):
    db = np.ascontiguousarray(db, np.float32)
    q = np.ascontiguousarray(q, np.float32).reshape(-1, db.shape[1])
    D = np.empty((q.shape[0], k), np.float32)
    I = np.empty((q.shape[0], k), np.uint64)
    idp = None
    if ids is not None:
        ids = np.ascontiguousarray(ids, np.uint64)
        idp = ids.ctypes.data_as(C.c_void_p)
    port_lib().oracle_flat_search(db, idp, db.shape[0], db.shape[1], metric, q, q.shape[0], k, D, I)
    return D, I
