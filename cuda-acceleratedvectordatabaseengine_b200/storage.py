"""Arrow-backed loader that places shards directly in HBM.

File format = the reference's ArrowStorage (format/storage.cpp:183-226, 287-292): one Arrow IPC *file* holding one
record batch with the schema ``{id: uint64, vector: list<float32>}``.  The reference reads such a file into an
``arrow::Table`` and never hands it to the engine (``load_from_epoch`` does not exist, server/query_service.cpp:245).
Here the file is memory-mapped; the ``vector`` column's values buffer is already one contiguous row-major fp32
array and the ``id`` column one contiguous u64 array, so both go to ``vdb_index_add`` as zero-copy host pointers:
assignment runs on the GPU and every row lands in the HBM page of its inverted list -- no intermediate host copy,
no per-row builder loop.  A sharded index (``Config.shard_rank/shard_count``) keeps only the lists it owns.
"""
import ctypes as C
import os

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
STORAGE_LIB_PATH = os.path.join(_HERE, "libvdb_b200_storage.so")
_SLIB = None

# every symbol include/vdb_b200_storage.h declares: name -> (restype, argtypes)
_vp, _u32, _u64, _i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
STORAGE_ABI = {
    "vdb_index_save_epoch": (_i32, [_vp, C.c_char_p, C.c_char_p, C.c_char_p]),
    "vdb_index_save": (_i32, [_vp, C.c_char_p]),
    "vdb_index_load": (_i32, [_vp, C.c_char_p]),
    "vdb_storage_write_vectors": (_i32, [C.c_char_p, _vp, _vp, _u64, _u32]),
    "vdb_storage_read_vectors": (_i32, [C.c_char_p, _vp, _vp, C.POINTER(_u64), C.POINTER(_u32)]),
    "vdb_storage_last_error": (C.c_char_p, []),
}


def storage_lib():
    """libvdb_b200_storage.so (C++ on Apache Arrow): the epoch directory of format/storage.cpp <-> the HBM index."""
    global _SLIB
    if _SLIB is None:
        if not os.path.exists(STORAGE_LIB_PATH):
            raise RuntimeError(f"{STORAGE_LIB_PATH} is not built: run __graft_entry__.build()")
        l = C.CDLL(STORAGE_LIB_PATH)
        for name, (res, args) in STORAGE_ABI.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _SLIB = l
    return _SLIB


def _scheck(st):
    if st == 0:
        return
    msg = storage_lib().vdb_storage_last_error().decode()
    if st == 1:
        raise ValueError(msg)
    raise RuntimeError(f"{msg} [{st}]")


def save_epoch(index, directory, index_name="", epoch=""):
    """IVFFlatIndex::save (ivf_flat_index.h:66): manifest.json + centroids.arrow + one list_<id>.arrow per list"""
    _scheck(storage_lib().vdb_index_save_epoch(index._h, os.fsencode(directory), index_name.encode(), epoch.encode()))


def load_epoch(index, directory):
    """IVFFlatIndex::load / load_from_epoch (server/query_service.cpp:245): the list files are memory-mapped and
    copied straight from the mapping into the owning GPU's HBM pages"""
    _scheck(storage_lib().vdb_index_load(index._h, os.fsencode(directory)))


def vector_schema():
    """create_vector_schema, format/storage.cpp:287-292"""
    return pa.schema([pa.field("id", pa.uint64()), pa.field("vector", pa.list_(pa.float32()))])


def write_vectors(path, vectors, ids=None):
    """ArrowStorage::write_vectors (format/storage.cpp:183-226): one record batch, IPC file format."""
    v = np.ascontiguousarray(vectors, np.float32)
    n, dim = v.shape
    ids = np.arange(n, dtype=np.uint64) if ids is None else np.ascontiguousarray(ids, np.uint64)
    offsets = pa.array(np.arange(0, (n + 1) * dim, dim, dtype=np.int32))
    col = pa.ListArray.from_arrays(offsets, pa.array(v.reshape(-1)))
    batch = pa.RecordBatch.from_arrays([pa.array(ids), col], schema=vector_schema())
    with pa.OSFile(path, "wb") as f, pa.ipc.new_file(f, vector_schema()) as w:
        w.write_batch(batch)


def read_vectors(path):
    """-> (ids [n] u64, vectors [n][dim] f32) as zero-copy views of the memory-mapped file."""
    src = pa.memory_map(path, "r")
    reader = pa.ipc.open_file(src)
    if reader.schema.names != ["id", "vector"]:
        raise ValueError(f"{path}: not a vdb vector file (schema {reader.schema})")
    parts = []
    for b in range(reader.num_record_batches):
        batch = reader.get_batch(b)
        idc, vc = batch.column(0), batch.column(1)
        n = batch.num_rows
        off = vc.offsets.to_numpy()
        if n == 0:
            continue
        dim = int(off[1] - off[0])
        if not np.all(np.diff(off) == dim):
            raise ValueError(f"{path}: ragged vectors")
        vals = vc.values.to_numpy(zero_copy_only=True)[int(off[0]): int(off[0]) + n * dim]
        parts.append((idc.to_numpy(zero_copy_only=True), vals.reshape(n, dim), src))
    return parts


def load_into_index(index, path):
    """Stream an Arrow vector file straight into the index's HBM pages; returns the number of rows added."""
    total = 0
    for ids, vecs, _keepalive in read_vectors(path):
        if vecs.shape[1] != index.config.dimension:
            raise ValueError(f"{path}: dimension {vecs.shape[1]} != index dimension {index.config.dimension}")
        index.add(vecs, ids)
        total += vecs.shape[0]
    return total


def load_centroids(index, path):
    """read_centroids (format/storage.cpp:228-232): same file format, ids are the centroid numbers."""
    (ids, vecs, _k), = read_vectors(path)
    order = np.argsort(ids, kind="stable")
    index.centroids = vecs[order]
