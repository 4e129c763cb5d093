/*
 * vdb_b200_storage.h -- C ABI of the epoch persistence of the B200 IVF-Flat index (libvdb_b200_storage.so).
 *
 * Kept apart from vdb_b200.h / libvdb_b200.so because it links Apache Arrow (the reference's storage layer does,
 * format/storage.h:3-6); it uses the index only through the public C ABI.  On-disk layout = the reference's epoch
 * directory (format/storage.cpp):
 *   <dir>/manifest.json     IndexManifest::to_json, :22-56  (index_name, epoch, dimension, nlist, metric,
 *                           pq_params{m,nbits}, shards[{list_id, path, num_vectors, file_size}], created_at ns)
 *   <dir>/centroids.arrow   ArrowStorage::write_centroids, :234-246  ({id: uint64, vector: list<float32>}, one batch)
 *   <dir>/list_<id>.arrow   ArrowStorage::write_vectors, :183-226, one file per non-empty inverted list
 * Replaces IVFFlatIndex::save / load (ivf_flat_index.h:66-67, declared and never defined) and gives the server's
 * index->load_from_epoch(epoch) (server/query_service.cpp:245) something to call.
 */
#ifndef VDB_B200_STORAGE_H
#define VDB_B200_STORAGE_H

#include "vdb_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Write the index (centroids + every list's rows and ids) as an epoch directory.  index_name / epoch may be NULL. */
int32_t vdb_index_save_epoch(vdb_index* ix, const char* dir, const char* index_name, const char* epoch);
int32_t vdb_index_save(vdb_index* ix, const char* dir);
/* Load an epoch directory into an EMPTY index of the same dimension / nlist / metric: centroids are set, a sharded
 * index re-balances its list ownership from the manifest's list sizes, and every list file is memory-mapped and
 * copied straight from the mapping into the owner's HBM pages (no assignment, no host staging copy). */
int32_t vdb_index_load(vdb_index* ix, const char* dir);
/* ArrowStorage::write_vectors / read_vectors for callers that hold plain arrays: rows [n][dim], ids [n].
 * read: call with vectors == NULL to get n and dim, then with buffers of that size. */
int32_t vdb_storage_write_vectors(const char* path, const float* vectors, const uint64_t* ids, uint64_t n,
                                  uint32_t dim);
int32_t vdb_storage_read_vectors(const char* path, float* vectors, uint64_t* ids, uint64_t* n, uint32_t* dim);
const char* vdb_storage_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* VDB_B200_STORAGE_H */
