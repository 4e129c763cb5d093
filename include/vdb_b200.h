/*
 * vdb_b200.h -- C ABI of the B200-native IVF-Flat hot path.
 *
 * This is the drop-in boundary: plain C, opaque handles, POD arguments, an
 * int32 status return, no exceptions and no torch / C++ types in any
 * signature.  Each entry point names the reference interface it replaces
 * (wedevxer/CUDA-AcceleratedVectorDatabaseEngine, file:line).  The reference's
 * host object links against exactly two kernel launchers and seven
 * TransferManager methods; the C++ mirror of vdb::IVFFlatIndex
 * in cuda-acceleratedvectordatabaseengine_b200/host/ and the ctypes binding in
 * cuda-acceleratedvectordatabaseengine_b200/__init__.py sit on top of this
 * header and nothing else.
 *
 * Pointer rules: every `const float*` / `const uint64_t*` input and every
 * output array may be HOST or DEVICE memory unless the name says `_dev`
 * (cudaPointerGetAttributes decides); inputs are borrowed for the duration of
 * the call, outputs are caller-allocated, the library owns all HBM it
 * allocates.  There is no CPU fallback anywhere: a missing device or a failed
 * launch is reported as a status code.
 */
#ifndef VDB_B200_H
#define VDB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    VDB_OK = 0,
    VDB_INVALID_ARGUMENT = 1, /* ctor's std::invalid_argument, ivf_flat_index.cpp:17-19 */
    VDB_OUT_OF_MEMORY = 2,    /* pool alloc returning nullptr, transfer_manager.cpp:125 */
    VDB_CUDA_ERROR = 3,       /* cudaGetLastError polling, ivf_flat_index.cpp:575 */
    VDB_NCCL_ERROR = 4,
    VDB_NOT_TRAINED = 5,
    VDB_INTERNAL = 6
} vdb_status;

/* kernels.cuh:24-28.  L2 is the SQUARED distance, InnerProduct is -dot. */
typedef enum { VDB_METRIC_L2 = 0, VDB_METRIC_IP = 1, VDB_METRIC_COSINE = 2 } vdb_metric;

/* Assignment and centroid update are always bit-identical to the reference
 * (the tensor-core assignment re-checks its survivors in reference order).
 * AUTO (default) is bit-identical throughout: the reference's sequential fp32
 * sums of the k-means++ sampling are evaluated exactly by a parallel scan
 * (seed_sample_par_kernel), tensor-core assignment when nlist >= 256.
 * EXACT = the literal restatement -- one-lane sequential sums, scalar
 * assignment kernel -- kept as the checker for AUTO.  FAST samples with
 * parallel double-precision sums (same RNG stream, a pick can differ when the
 * target lies within rounding distance of a row boundary). */
typedef enum { VDB_TRAIN_AUTO = 0, VDB_TRAIN_EXACT = 1, VDB_TRAIN_FAST = 2 } vdb_train_mode;

/* How the query x centroid coarse distances are produced:
 * SIMT   exact fp32 scan kernel (same kernel as the list scan);
 * TENSOR tcgen05 TF32 contraction + exact fp32 re-check of every candidate
 *        within the rounding-error bound of the nprobe-th distance. */
typedef enum { VDB_COARSE_AUTO = 0, VDB_COARSE_SIMT = 1, VDB_COARSE_TENSOR = 2 } vdb_coarse_mode;

/* IVFFlatIndex::Config, ivf_flat_index.h:16-22 (dimension, nlist, metric,
 * max_gpu_memory keep their meaning; use_gpu has no counterpart: there is no
 * CPU path).  Zero-initialise and call vdb_config_default() first. */
typedef struct {
    uint32_t dimension;
    uint32_t nlist;
    int32_t metric;          /* vdb_metric */
    int32_t device;          /* CUDA device ordinal */
    uint64_t max_gpu_memory; /* cap on HBM owned by this index, 0 = no cap (ivf_flat_index.h:21) */
    int32_t train_mode;      /* vdb_train_mode */
    int32_t coarse_mode;     /* vdb_coarse_mode */
    uint32_t page_rows;      /* rows per inverted-list page, 0 = auto */
    uint32_t shard_rank;     /* this process holds the lists whose owner is shard_rank (see vdb_index_get_owners) */
    uint32_t shard_count;    /* 1 = unsharded, at most 255 */
    uint32_t pipeline_depth; /* searches in flight (vdb_index_search_submit), 0 = 4, at most 8 */
    uint32_t reserve_sms;    /* SMs a pipelined list scan leaves to the coarse / merge kernels of the neighbouring
                                batches, 0 = 8, 0xffffffff = none */
    uint32_t scan_mirror;    /* low-precision shadow of the inverted lists for the tensor-core screen of the list scan
                                (results unchanged: admitted pairs are re-scored in fp32): 0 = auto (int8 where
                                supported: row stride 128 * {1,2,4,6,8} floats), 1 = off, 2 = bf16 (+50 % HBM, half the
                                bytes streamed per search), 3 = int8 with a scale per row (+25 % HBM, a quarter of the
                                bytes); 2 and 3 are refused where unsupported */
    uint32_t reserved[2];
} vdb_config;

typedef struct {
    uint64_t total_vectors;      /* get_total_vectors(), ivf_flat_index.h:64 (all shards' adds seen by this rank) */
    uint64_t local_vectors;      /* rows resident on this device */
    uint64_t gpu_memory_bytes;   /* get_gpu_memory_usage(), ivf_flat_index.cpp:707-709 */
    uint64_t pages;              /* inverted-list pages in use */
    uint32_t dimension, nlist, row_stride, page_rows;
    int32_t trained;
    int32_t metric;              /* vdb_metric */
    uint64_t scanned_bytes;      /* distinct inverted-list bytes streamed from HBM by all searches so far */
} vdb_stats;

/* Byte accounting of the most recent search: probed rows summed
 * per query (algorithmic) and over distinct probed lists (unique). */
typedef struct {
    uint64_t algorithmic_rows;
    uint64_t unique_rows;
    uint64_t scan_items;
    uint64_t bytes_per_row; /* 4*dim + 8 */
    uint64_t scan_ctas;     /* persistent CTAs of the list-scan launch */
    uint64_t streamed_bytes_per_row; /* what the scan kernel streams per distinct row: bytes_per_row for the fp32
                                        scan; 2*row_stride + 8 (bf16) or row_stride + 12 (int8) when the screen
                                        ran (vdb_config.scan_mirror) */
    uint64_t rescored_pairs; /* (row, query) pairs the screen admitted and re-scored exactly in fp32 (each one
                                reads the row's fp32 copy: 4*row_stride more bytes); 0 for the fp32 scan */
} vdb_search_stats;

typedef struct vdb_index vdb_index;
typedef struct vdb_arena vdb_arena;
typedef struct vdb_exchange vdb_exchange;

const char* vdb_last_error_string(void);
const char* vdb_status_string(int32_t status);
int32_t vdb_version(void);

void vdb_config_default(vdb_config* cfg);

/* IVFFlatIndex::IVFFlatIndex(const Config&, TransferManager*), ivf_flat_index.cpp:13-33 */
int32_t vdb_index_create(const vdb_config* cfg, vdb_index** out);
/* The same object over several GPUs of ONE process (the reference server constructs one IVFFlatIndex in one
 * process, server/query_service.cpp:232-245): shard r lives on devices[r] and holds the lists whose owner is r
 * (vdb_index_get_owners; byte-balanced by train()); every vdb_index_* call below works on the returned handle.
 * cfg->device / shard_rank / shard_count are ignored.  Searches run on the shards' own pipelines -- submit / wait /
 * the synchronous vdb_index_search; vdb_index_search_async and vdb_index_add_assigned are refused -- and each
 * shard's merge kernel stores its top-k into the root shard's mailbox over NVLink peer access, where a collect
 * kernel merges them.  Results are identical to the unsharded index.  A device may be listed more than once. */
int32_t vdb_index_create_sharded(const vdb_config* cfg, const int32_t* devices, int32_t ndev, vdb_index** out);
/* IVFFlatIndex::~IVFFlatIndex, ivf_flat_index.cpp:36-46 */
int32_t vdb_index_destroy(vdb_index* ix);

/* IVFFlatIndex::train, ivf_flat_index.cpp:49-145: k-means++ (mt19937(42)) +
 * 10 Lloyd iterations, all on the device. */
int32_t vdb_index_train(vdb_index* ix, const float* vectors, uint64_t n);
/* IVFFlatIndex::add, ivf_flat_index.cpp:148-202: assign + append. */
int32_t vdb_index_add(vdb_index* ix, const float* vectors, const uint64_t* ids, uint64_t n);
/* Data-parallel add for a sharded index: the rows' lists were already computed (vdb_index_assign on
 * the rank that held them) and the rows routed to this rank; only rows of lists this rank owns are
 * kept.  ids are required.  global_n = rows added across all ranks by this collective step (what
 * get_total_vectors() grows by). */
int32_t vdb_index_add_assigned(vdb_index* ix, const float* vectors, const uint64_t* ids,
                               const uint32_t* assignments_dev, uint64_t n, uint64_t global_n);
/* IVFFlatIndex::search, ivf_flat_index.cpp:205-256.  distances/indices are
 * [nq][k]; missing results are padded FLT_MAX / UINT64_MAX (:380-383,:514-517).
 * nprobe is clamped to nlist.  Thread-safe against concurrent searches. */
int32_t vdb_index_search(vdb_index* ix, const float* queries, uint32_t nq, uint32_t nprobe, uint32_t k,
                         float* distances, uint64_t* indices);
/* Same, device pointers only, every launch enqueued on `stream` (a cudaStream_t)
 * in stream order, without a host synchronisation of this call (the call may
 * block on the search issued pipeline_depth calls earlier, whose scratch it
 * reuses).  Searches enqueued on different streams overlap. */
int32_t vdb_index_search_async(vdb_index* ix, const float* queries_dev, uint32_t nq, uint32_t nprobe,
                               uint32_t k, float* distances_dev, uint64_t* indices_dev, void* stream);
/* Pipelined search: the serving form (the reference server's intended 64-query
 * batches in flight, query_service.h:26-27).  submit() enqueues the batch on the
 * index's own streams and returns a ticket at once; up to pipeline_depth batches
 * are in flight and overlap: while batch i's list scan streams from HBM, batch
 * i+1's coarse selection / probe grouping and batch i-1's merge (and cross-GPU
 * exchange) run on the SMs the scan leaves free, and batch i+1's scan CTAs start
 * on the SMs batch i's tail releases.  queries / outputs may be host or device
 * memory; a device query array must already hold its values and, like the
 * outputs, stay valid until the ticket is waited for.  wait() blocks the host
 * until the batch is complete (host outputs are filled in by then);
 * wait_stream() instead makes `stream` wait for it (device outputs).  Results
 * are identical to vdb_index_search; vdb_index_search is submit + wait, so
 * concurrent host threads overlap the same way. */
int32_t vdb_index_search_submit(vdb_index* ix, const float* queries, uint32_t nq, uint32_t nprobe, uint32_t k,
                                float* distances, uint64_t* indices, uint64_t* ticket);
int32_t vdb_index_search_wait(vdb_index* ix, uint64_t ticket);
int32_t vdb_index_search_wait_stream(vdb_index* ix, uint64_t ticket, void* stream);
/* The reference's index takes its device memory from the TransferManager it is constructed with
 * (ivf_flat_index.cpp:13, :424-433).  Same here: with an arena attached (before the first add), list slabs come from
 * vdb_arena_allocate_device on `arena_device`; a slab the pool cannot hold is allocated directly instead (the
 * reference would leave the list on the host and search it on the CPU).  The arena is borrowed and must outlive
 * the index. */
int32_t vdb_index_set_arena(vdb_index* ix, vdb_arena* arena, int32_t arena_device);
/* Allocate every search buffer for batches of up to (max_nq, max_nprobe, max_k)
 * now (and again after each add()), so that no allocation happens inside a search. */
int32_t vdb_index_reserve_search(vdb_index* ix, uint32_t max_nq, uint32_t max_nprobe, uint32_t max_k);
/* One process per GPU: attach this rank's connected vdb_exchange (borrowed, may
 * be NULL to detach).  From then on every search is collective -- same calls in
 * the same order on every rank -- and returns the merged result of all shards:
 * the merge kernel stores the shard's top-k straight into the peers' mailboxes
 * and a small collect kernel merges the world's blocks. */
int32_t vdb_index_attach_exchange(vdb_index* ix, vdb_exchange* ex);
/* select_nprobe_lists, ivf_flat_index.cpp:298-336: [nq][min(nprobe,nlist)] list ids. */
int32_t vdb_index_select_nprobe(vdb_index* ix, const float* queries, uint32_t nq, uint32_t nprobe,
                                uint32_t* lists);
/* assign_to_lists, ivf_flat_index.cpp:259-295. */
int32_t vdb_index_assign(vdb_index* ix, const float* vectors, uint64_t n, uint32_t* lists);

/* Test hooks / persistence seam: centroids_ is [nlist][dimension] fp32 row-major. */
int32_t vdb_index_get_centroids(vdb_index* ix, float* out);
int32_t vdb_index_set_centroids(vdb_index* ix, const float* in);
/* List ownership of a sharded index: owner[l] = rank holding list l.  Starts as l % shard_count;
 * vdb_index_train() re-balances it by bytes (greedy, largest list first, from the training sample's
 * list sizes -- deterministic, identical on every rank); may be set explicitly while the index is empty. */
int32_t vdb_index_get_owners(vdb_index* ix, uint8_t* out /* [nlist] */);
int32_t vdb_index_set_owners(vdb_index* ix, const uint8_t* in /* [nlist] */);
int32_t vdb_index_list_sizes(vdb_index* ix, uint64_t* out /* [nlist] */);
int32_t vdb_index_list_ids(vdb_index* ix, uint32_t list, uint64_t* out /* [list size] */);
/* Persistence seam (libvdb_b200_storage.so writes / reads the reference's epoch layout on top of these):
 * the rows of one list in storage order, [list size][dimension] (host or device memory); append rows to a list
 * without assignment, copied straight from the caller's memory (e.g. a memory-mapped Arrow values buffer) into the
 * list's HBM pages -- a sharded index stores them only on the owner; finish_load publishes the appended lists to the
 * search tables once at the end; balance_owners re-computes the byte-balanced ownership of an EMPTY sharded index
 * from per-list row counts (what train() does from its sample). */
int32_t vdb_index_list_vectors(vdb_index* ix, uint32_t list, float* out /* [list size][dimension] */);
int32_t vdb_index_append_list(vdb_index* ix, uint32_t list, const float* vectors, const uint64_t* ids, uint64_t n);
int32_t vdb_index_finish_load(vdb_index* ix);
int32_t vdb_index_balance_owners(vdb_index* ix, const uint64_t* list_sizes /* [nlist] */);
int32_t vdb_index_stats(vdb_index* ix, vdb_stats* out);
int32_t vdb_index_last_search_stats(vdb_index* ix, vdb_search_stats* out);
/* Profiling for the roofline report: when enabled, CUDA events on the launching
 * streams bracket, for every search, [0] the coarse step, [1] probe grouping,
 * [2] the list-scan kernel, [3] the merge (+ publish), [4] the cross-GPU collect;
 * [5] = first scan start .. last scan end (the scan streams' busy time: with
 * batches in flight consecutive scans overlap, so this, not the sum of [2], is
 * the time the scans took).  read_profile synchronises, returns the summed
 * milliseconds since the last call and the number of searches. */
int32_t vdb_index_set_profiling(vdb_index* ix, int32_t enable);
int32_t vdb_index_read_profile(vdb_index* ix, float* out_ms /* [8] */, uint32_t* searches);
/* warmup_lists()/warmup_all() of the server contract (query_service.cpp:191,195):
 * lists are always HBM-resident here, so this only validates ids. */
int32_t vdb_index_warmup(vdb_index* ix, const uint32_t* lists, uint32_t n);

/* kernels::launch_bruteforce_search<float>, kernels.cu:13-43 / kernels.cuh:391-396:
 * exact top-k of nq queries against a flat [n][dim] row-major database.
 * ids may be NULL (row number is the id).  k is not capped at 32. */
int32_t vdb_bruteforce_search(const float* database, const float* queries, const uint64_t* ids, uint64_t n,
                              uint32_t nq, uint32_t dim, uint32_t k, float* distances, uint64_t* indices,
                              int32_t metric, void* stream);
/* kernels::launch_kmeans_assign<float>, kernels.cu:80-92 / kernels.cuh:410-414.
 * The reference kernel is L2-only; `metric` follows the CPU path instead
 * (ivf_flat_index.cpp:275-285).  distances may be NULL. */
int32_t vdb_kmeans_assign(const float* vectors, const float* centroids, uint32_t* assignments, float* distances,
                          uint64_t n, uint32_t n_centroids, uint32_t dim, int32_t metric, void* stream);
/* Data-parallel k-means building blocks (device pointers): per-cluster fp32
 * sums / counts of this rank's rows, and the division step after the caller
 * has all-reduced them (ivf_flat_index.cpp:123-141). */
int32_t vdb_kmeans_accumulate(const float* vectors_dev, const uint32_t* assignments_dev, uint64_t n,
                              uint32_t n_centroids, uint32_t dim, float* sums_dev, uint32_t* counts_dev,
                              void* stream);
int32_t vdb_kmeans_finalize(const float* sums_dev, const uint32_t* counts_dev, float* centroids_dev,
                            uint32_t n_centroids, uint32_t dim, void* stream);
/* merge_results, ivf_flat_index.cpp:474-518, across shards: `parts` blocks of
 * [nq][k] (as produced by an all-gather of every rank's local top-k) -> [nq][k]
 * by (distance, id), duplicates removed, padded.  Device pointers. */
int32_t vdb_merge_topk(const float* dist_parts_dev, const uint64_t* id_parts_dev, uint32_t parts, uint32_t nq,
                       uint32_t k, float* distances_dev, uint64_t* indices_dev, void* stream);

/* The same merge fused with the exchange, over NVLink peer memory (csrc/exchange.cu): one kernel per rank stores
 * the rank's local [nq][k] block into every peer's mailbox, waits for the peers' blocks and merges -- instead of
 * two all-gathers plus vdb_merge_topk.  One process per GPU: create on every rank, pass the 64-byte handle of
 * every rank (rank order; e.g. from an all_gather of vdb_exchange_handle) to connect, then call merge_topk
 * collectively (same order and shapes on every rank).  world * max_k <= 4096.  A peer that never arrives (20 s
 * timeout) makes the call's results padded; vdb_exchange_status() -- after synchronising with the stream -- and
 * every later call return VDB_NCCL_ERROR until vdb_exchange_reset(). */
int32_t vdb_exchange_create(int32_t device, uint32_t rank, uint32_t world, uint32_t max_nq, uint32_t max_k,
                            vdb_exchange** out);
int32_t vdb_exchange_handle(vdb_exchange* ex, uint8_t* out64);
int32_t vdb_exchange_connect(vdb_exchange* ex, const uint8_t* handles);
int32_t vdb_exchange_merge_topk(vdb_exchange* ex, const float* local_dist_dev, const uint64_t* local_ids_dev,
                                uint32_t nq, uint32_t k, float* distances_dev, uint64_t* indices_dev, void* stream);
/* The same in two launches, so that other work of the stream (the next batch's scan) runs between them and nobody
 * waits for the slowest rank of a batch: publish(i) ... collect(i) -> merged results of batch i.  One batch in
 * flight: every rank calls collect(i) before publish(i + 1). */
int32_t vdb_exchange_publish(vdb_exchange* ex, const float* local_dist_dev, const uint64_t* local_ids_dev, uint32_t nq,
                             uint32_t k, void* stream);
int32_t vdb_exchange_collect(vdb_exchange* ex, float* distances_dev, uint64_t* indices_dev, void* stream);
/* Exchanges of ONE process (a single-process sharded index): `all` in rank order.  No IPC: the root's mailbox is
 * addressed directly through peer access, every rank publishes into it and the root alone collects. */
int32_t vdb_exchange_connect_local(vdb_exchange** all, uint32_t world, uint32_t root);
/* VDB_NCCL_ERROR once a wait on a peer has timed out (valid after synchronising with the stream of the call);
 * reset clears that state (all ranks must call it, then continue in step). */
int32_t vdb_exchange_status(vdb_exchange* ex);
int32_t vdb_exchange_reset(vdb_exchange* ex);
int32_t vdb_exchange_destroy(vdb_exchange* ex);

/* TransferManager rewrite (transfer_manager.h:42-88): one HBM slab and one
 * pinned slab carved by a best-fit allocator, a stream pool, async copies. */
int32_t vdb_arena_create(int32_t device, uint64_t device_bytes, uint64_t pinned_bytes, int32_t num_streams,
                         vdb_arena** out);
int32_t vdb_arena_destroy(vdb_arena* a);
void* vdb_arena_allocate_device(vdb_arena* a, uint64_t bytes);  /* allocate_device */
int32_t vdb_arena_free_device(vdb_arena* a, void* p);           /* free_device */
void* vdb_arena_allocate_pinned(vdb_arena* a, uint64_t bytes);  /* allocate_pinned */
int32_t vdb_arena_free_pinned(vdb_arena* a, void* p);           /* free_pinned */
void* vdb_arena_get_stream(vdb_arena* a);                       /* get_stream */
int32_t vdb_arena_return_stream(vdb_arena* a, void* stream);    /* return_stream */
/* enqueue_transfer: kind 1 = H2D, 2 = D2H, 3 = D2D (cudaMemcpyKind values) */
int32_t vdb_arena_enqueue_transfer(vdb_arena* a, void* dst, const void* src, uint64_t bytes, int32_t kind,
                                   void* stream);
/* the same with Transfer::callback: a host function enqueued behind the copy (cudaLaunchHostFunc), caller not blocked */
int32_t vdb_arena_enqueue_transfer_cb(vdb_arena* a, void* dst, const void* src, uint64_t bytes, int32_t kind,
                                      void* stream, void (*callback)(void*), void* user);
int32_t vdb_arena_synchronize(vdb_arena* a);                     /* synchronize */
int32_t vdb_arena_synchronize_stream(vdb_arena* a, void* stream); /* synchronize_stream */
/* out[0..3] = device bytes in use, device peak, pinned in use, live allocations */
int32_t vdb_arena_stats(vdb_arena* a, uint64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* VDB_B200_H */
