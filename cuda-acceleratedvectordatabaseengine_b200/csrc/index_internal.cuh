// Internal layout of a vdb_index (index.cu, persist.cu, sharded_index.cu share it; never crosses the C ABI).
#pragma once
#include <mutex>
#include <vector>

#include "coarse.cuh"
#include "common.cuh"
#include "kmeans.cuh"
#include "scan.cuh"

struct vdb_exchange;

namespace vdb {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    int32_t reserve(size_t n) {
        if (n <= cap) return VDB_OK;
        cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 16;
        VDB_CUDA_TRY(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
        return VDB_OK;
    }
    void release() {
        cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    size_t bytes() const { return cap * sizeof(T); }
};

// pinned host staging (query upload / result download of host-side callers)
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    int32_t reserve(size_t bytes) {
        if (bytes <= cap) return VDB_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        VDB_CUDA_TRY(cudaMallocHost(&p, want));
        cap = want;
        return VDB_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

constexpr uint32_t MAX_SEARCH_SLOTS = 8;
constexpr uint32_t SLOT_TIMERS = 8;

// One search in flight.  A search is three phases that only meet through events, so consecutive batches overlap:
//   front  (high-priority stream)  query staging, coarse selection (tcgen05 GEMM + select), probe grouping
//   scan   (two alternating streams) the persistent list-scan kernel; the next batch's CTAs start on the SMs the
//          previous batch's tail frees
//   back   (high-priority stream)  merge (+ publish to the peers' mailboxes, + collect), result download
// Each slot owns every buffer one search touches, so `depth` batches are in flight without sharing scratch.
struct SearchSlot {
    ScanWorkspace ws_scan, ws_coarse;
    DevBuf<float> q_buf, q_raw, dots, coarse_d, out_d;  // q_raw: queries peer-copied from another device
    DevBuf<uint64_t> coarse_i, out_i;
    DevBuf<uint32_t> probes, zero_probes;
    PinnedBuf h_q, h_d, h_i;
    cudaEvent_t ev_front = nullptr, ev_scan = nullptr, ev_done = nullptr;
    // profiling marks: [0] front start [1] coarse end [2] grouping end | [3] scan start [4] scan end |
    // [5] back start [6] merge (+publish) end [7] collect end
    cudaEvent_t tm[SLOT_TIMERS] = {nullptr};
    bool timed = false;
    uint64_t ticket = 0;  // the search this slot serves / served last (0 = never used)
    bool busy = false;    // enqueued and not yet known to be complete
    // host outputs are downloaded into h_d / h_i and copied to the caller's arrays when the search is waited for
    float* user_d = nullptr;
    uint64_t* user_i = nullptr;
    size_t out_elems = 0;
    bool deliver = false;
    bool used_exchange = false;
    // where the merged result goes: device arrays, and (host callers) the pinned or staged host arrays behind them
    float* dev_d = nullptr;
    uint64_t* dev_i = nullptr;
    float* host_d = nullptr;
    uint64_t* host_i = nullptr;
    // sharded search: the publish went out, the collect is enqueued behind the NEXT batch's scan (or on demand)
    bool collect_deferred = false;
    cudaStream_t collect_stream = nullptr;
    ScanLaunchInfo info{};

    uint64_t bytes() const {
        return ws_scan.bytes + ws_coarse.bytes + q_buf.bytes() + dots.bytes() + coarse_d.bytes() + out_d.bytes() +
               coarse_i.bytes() + out_i.bytes() + probes.bytes() + zero_probes.bytes() + q_raw.bytes();
    }
};

struct SearchStreams {
    cudaStream_t front, scan, back;
    bool split;  // the three differ: order them with the slot's events
    cudaEvent_t back_wait = nullptr;  // the back phase also waits for this (mailbox reuse of a root-only exchange)
};

struct Composite;  // sharded_index.cu: the shards of a single-process multi-device index

}  // namespace vdb

struct vdb_index {
    // a single-process multi-device index (vdb_index_create_sharded) is a handle whose work is done by one
    // vdb_index per device; none of the fields below the config are used on it
    vdb::Composite* composite = nullptr;
    vdb_config cfg{};
    uint32_t dim = 0, ld = 0, nlist = 0, page_rows = 0;
    uint64_t page_bytes = 0, ids_off = 0;
    uint32_t mirror_off = 0, mirror_kind = 0;  // low-precision shadow of every page (ListTable::mirror_off), 0 = none
    int device = 0;
    cudaStream_t stream = nullptr;  // train / add / bookkeeping
    bool trained = false;
    std::mutex mu;

    vdb::DevBuf<float> centroids;  // [nlist][ld]
    vdb::DevBuf<float> cnorm;      // [nlist] |c|^2 (tensor-core coarse path)
    vdb::DevBuf<uint32_t> cmax_bits;
    // flat paged view of the centroid table (the SIMT coarse step scans it like a list)
    vdb::DevBuf<uint32_t> c_rows, c_page_off;
    vdb::DevBuf<uint64_t> c_page_vec, c_page_ids;
    uint32_t c_npages = 0;

    // inverted lists: host mirror of the page chains + device tables
    std::vector<uint32_t> h_rows;
    std::vector<std::vector<uint32_t>> h_pages;
    std::vector<void*> slabs;
    std::vector<bool> slab_pooled;     // slab i came from the arena (TransferManager pool), not from cudaMalloc
    vdb_arena* arena = nullptr;        // borrowed (vdb_index_set_arena): must outlive the index, like the reference's tm_
    int arena_device = 0;
    unsigned long long* d_scanned = nullptr;  // device counter: distinct list rows streamed by all searches
    std::vector<uint64_t> page_addr;
    uint32_t pages_per_slab = 0, pages_used = 0;
    vdb::DevBuf<uint32_t> d_rows, d_page_off;
    vdb::DevBuf<uint64_t> d_page_vec, d_page_ids;
    std::vector<uint32_t> npages_desc;  // page counts, descending (search slot bound)
    bool tables_dirty = false;          // h_rows / h_pages changed without an upload (vdb_index_append_list)

    uint64_t total_vectors = 0, local_vectors = 0, slab_bytes_total = 0;
    // sharding: owner[l] = rank that holds list l (empty on an unsharded index)
    std::vector<uint8_t> h_owner;
    vdb::DevBuf<uint8_t> d_owner;

    // search pipeline
    vdb::SearchSlot slots[vdb::MAX_SEARCH_SLOTS];
    vdb::SearchSlot aux_slot;    // vdb_index_select_nprobe (synchronous, outside the ring)
    uint32_t depth = 4;          // slots in use
    uint32_t reserve_sms = 8;    // SMs a pipelined scan leaves to the front / back kernels of its neighbours
    uint32_t ppi_override = 0;   // pages per scan item (0 = heuristic)
    uint32_t dot_min_rows = 20000; // the L2 scan screens by dot product when a launch streams >= this many distinct rows per CTA
    bool scan_exact = false;     // VDB_SCAN_EXACT=1: L2 scan without the dot-form screen (A/B measurements)
    uint64_t next_ticket = 0;
    int last_slot = -1;
    cudaStream_t s_front = nullptr, s_scan[2] = {nullptr, nullptr}, s_back = nullptr;
    vdb_exchange* exchange = nullptr;  // borrowed: attached => searches return the merged result of all shards
    vdb::SearchSlot* deferred = nullptr;  // the slot whose collect is still owed (at most one)
    uint32_t rs_nq = 0, rs_np = 0, rs_k = 0;  // shape the slots were pre-reserved for (0 = none)

    vdb::DevBuf<uint32_t> assign_buf, hist_buf, fill_buf;
    vdb::DevBuf<float> stage_buf;
    vdb::DevBuf<uint64_t> ids_stage;
    vdb::AssignTcScratch tc_assign;

    // profiling (vdb_index_set_profiling): per-phase sums harvested from the slots' timing events
    bool profiling = false;
    double prof_ms[6] = {0, 0, 0, 0, 0, 0};  // coarse, group, scan, merge(+publish), collect, -
    uint32_t prof_searches = 0;
    cudaEvent_t span_start = nullptr, span_end = nullptr;  // first scan start .. last scan end
    bool span_open = false;

    uint64_t hbm_bytes() const {
        uint64_t b = slab_bytes_total + centroids.bytes() + cnorm.bytes() + c_rows.bytes() + c_page_off.bytes() +
                     c_page_vec.bytes() + c_page_ids.bytes() + d_rows.bytes() + d_page_off.bytes() +
                     d_page_vec.bytes() + d_page_ids.bytes() + assign_buf.bytes() + hist_buf.bytes() +
                     fill_buf.bytes() + stage_buf.bytes() + ids_stage.bytes();
        for (uint32_t i = 0; i < vdb::MAX_SEARCH_SLOTS; ++i) b += slots[i].bytes();
        return b;
    }
};

namespace vdb {

// index.cu internals used by the sibling translation units
bool is_device_ptr(const void* p);
bool is_pinned_ptr(const void* p);
int32_t index_upload_list_tables(vdb_index* ix);
int32_t index_alloc_page(vdb_index* ix, uint32_t* page);
int32_t index_refresh_centroids(vdb_index* ix);  // norms + flat view after the centroid table changed
int32_t index_upload_owners(vdb_index* ix);
int32_t index_assign_rows(vdb_index* ix, const float* x, uint64_t n, uint32_t* out, cudaStream_t stream);
void index_balance_owners(vdb_index* ix, const std::vector<uint32_t>& counts);  // fills ix->h_owner
// enqueue one search of `ix` into slot `s` (see SearchSlot); queries may be host or device memory.
// collect = false: the merged local result is only published (non-root shard of a single-process sharded index)
int32_t index_enqueue_search(vdb_index* ix, SearchSlot& s, const float* queries, uint32_t nq, uint32_t nprobe,
                             uint32_t k, float* distances, uint64_t* indices, const SearchStreams& st, bool collect);
int32_t index_acquire_slot(vdb_index* ix, SearchSlot** out, uint64_t* ticket);
int32_t index_finish_slot(vdb_index* ix, SearchSlot& s);  // host-wait + deliver + harvest timings
int32_t index_flush_deferred(vdb_index* ix, SearchSlot& s);  // enqueue the slot's owed collect (no-op otherwise)
SearchStreams index_pipeline_streams(vdb_index* ix, uint64_t ticket);
// sum of the partial-result bytes one query needs (0 chunks => fits): nq_chunk < nq means "split the batch"
void index_choose_ppi(const vdb_index* ix, uint32_t nq, uint32_t np, uint32_t k, uint32_t* ppi, uint32_t* nq_chunk);

// sharded_index.cu: the composite's side of every C-ABI entry point
int32_t composite_destroy(vdb_index* ix);
int32_t composite_train(vdb_index* ix, const float* vectors, uint64_t n);
int32_t composite_add(vdb_index* ix, const float* vectors, const uint64_t* ids, uint64_t n);
int32_t composite_submit(vdb_index* ix, const float* queries, uint32_t nq, uint32_t nprobe, uint32_t k,
                         float* distances, uint64_t* indices, uint64_t* ticket);
int32_t composite_wait(vdb_index* ix, uint64_t ticket);
int32_t composite_wait_stream(vdb_index* ix, uint64_t ticket, cudaStream_t stream);
vdb_index* composite_root(vdb_index* ix);
uint32_t composite_size(vdb_index* ix);
vdb_index* composite_shard(vdb_index* ix, uint32_t r);
int32_t composite_set_centroids(vdb_index* ix, const float* in);
int32_t composite_set_owners(vdb_index* ix, const uint8_t* in);
int32_t composite_list_sizes(vdb_index* ix, uint64_t* out);
int32_t composite_list_ids(vdb_index* ix, uint32_t list, uint64_t* out);
int32_t composite_stats(vdb_index* ix, vdb_stats* out);
int32_t composite_last_search_stats(vdb_index* ix, vdb_search_stats* out);
int32_t composite_set_profiling(vdb_index* ix, int32_t enable);
int32_t composite_read_profile(vdb_index* ix, float* out_ms, uint32_t* searches);
int32_t composite_reserve_search(vdb_index* ix, uint32_t nq, uint32_t np, uint32_t k);
int32_t composite_note_added(vdb_index* ix, uint64_t n);

}  // namespace vdb
