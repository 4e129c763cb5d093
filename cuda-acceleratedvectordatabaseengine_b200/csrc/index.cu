// vdb_index: the device-resident IVF-Flat index behind the C ABI.
//
// Host orchestration of IVFFlatIndex::{train, add, search}
// (ivf_flat_index.cpp:49-256) with every arithmetic step on the GPU.
// HBM layout: centroids [nlist][ld] fp32; every inverted list is a chain of
// fixed-size pages ([page_rows][ld] fp32 rows + [page_rows] u64 ids) carved
// from geometrically growing slabs (64 MiB .. 2 GiB), so add() appends without ever moving resident rows and a
// page is the unit of scan work.  There is no host copy of the vectors and no
// CPU fallback.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <numeric>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "bruteforce_tc.cuh"
#include "coarse.cuh"
#include "common.cuh"
#include "exchange.cuh"
#include "index_internal.cuh"
#include "kmeans.cuh"
#include "scan.cuh"

namespace vdb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

bool is_pinned_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

namespace {

constexpr uint64_t SLAB_BYTES = 64ull << 20;
constexpr uint32_t CENTROID_PAGE_ROWS = 64;
constexpr uint64_t ADD_CHUNK_BYTES = 1ull << 30;
constexpr uint64_t PARTIAL_BYTES_CAP = 1ull << 30;  // partial-result buffer of one search

}  // namespace
}  // namespace vdb

using namespace vdb;

static void balance_owners_impl(vdb_index* ix, const std::vector<uint32_t>& counts);

namespace {

// assign_to_lists: tensor cores + exact re-check for large centroid tables (bit-identical to the scalar
// kernel), the scalar order-exact kernel otherwise or when train_mode = EXACT asks for it
int32_t assign_rows(vdb_index* ix, const float* x, uint64_t n, uint32_t* out, cudaStream_t stream) {
    if (ix->cfg.train_mode != VDB_TRAIN_EXACT && n >= 256 && assign_tensor_supported(ix->nlist, ix->ld))
        return kmeans_assign_tensor(x, n, ix->ld, ix->centroids.p, ix->nlist, ix->ld, ix->dim, ix->cfg.metric, out,
                                    ix->tc_assign, stream);
    return kmeans_assign_exact(x, n, ix->ld, ix->centroids.p, ix->nlist, ix->ld, ix->dim, ix->cfg.metric, out, nullptr,
                               stream);
}

}  // namespace

int32_t vdb::index_assign_rows(vdb_index* ix, const float* x, uint64_t n, uint32_t* out, cudaStream_t stream) {
    return assign_rows(ix, x, n, out, stream);
}

void vdb::index_balance_owners(vdb_index* ix, const std::vector<uint32_t>& counts) { balance_owners_impl(ix, counts); }

namespace {

ListTable centroid_table(vdb_index* ix) {
    ListTable lt;
    lt.rows = ix->c_rows.p;
    lt.page_off = ix->c_page_off.p;
    lt.page_vec = ix->c_page_vec.p;
    lt.page_ids = ix->c_page_ids.p;
    lt.ids_flat = nullptr;
    lt.nlist = 1;
    lt.page_rows = CENTROID_PAGE_ROWS;
    lt.ld = ix->ld;
    return lt;
}

ListTable list_table(vdb_index* ix) {
    ListTable lt;
    lt.rows = ix->d_rows.p;
    lt.page_off = ix->d_page_off.p;
    lt.page_vec = ix->d_page_vec.p;
    lt.page_ids = ix->d_page_ids.p;
    lt.ids_flat = nullptr;
    lt.nlist = ix->nlist;
    lt.page_rows = ix->page_rows;
    lt.ld = ix->ld;
    lt.mirror_off = ix->mirror_off;
    lt.mirror_kind = ix->mirror_kind;
    return lt;
}

// Describe a flat [n][ld] device array as a one-list paged view.
int32_t build_flat_view(const float* base, uint64_t n, uint32_t ld, uint32_t page_rows, DevBuf<uint32_t>& rows,
                        DevBuf<uint32_t>& page_off, DevBuf<uint64_t>& page_vec, DevBuf<uint64_t>& page_ids,
                        uint32_t* npages_out, cudaStream_t stream) {
    const uint32_t npages = (uint32_t)((n + page_rows - 1) / page_rows);
    std::vector<uint64_t> pv(std::max(npages, 1u)), pi(std::max(npages, 1u), 0);
    for (uint32_t i = 0; i < npages; ++i) pv[i] = (uint64_t)(uintptr_t)(base + (size_t)i * page_rows * ld);
    uint32_t hrows = (uint32_t)n, hoff[2] = {0, npages};
    VDB_TRY(rows.reserve(1));
    VDB_TRY(page_off.reserve(2));
    VDB_TRY(page_vec.reserve(pv.size()));
    VDB_TRY(page_ids.reserve(pi.size()));
    VDB_CUDA_TRY(cudaMemcpyAsync(rows.p, &hrows, 4, cudaMemcpyHostToDevice, stream));
    VDB_CUDA_TRY(cudaMemcpyAsync(page_off.p, hoff, 8, cudaMemcpyHostToDevice, stream));
    VDB_CUDA_TRY(cudaMemcpyAsync(page_vec.p, pv.data(), pv.size() * 8, cudaMemcpyHostToDevice, stream));
    VDB_CUDA_TRY(cudaMemcpyAsync(page_ids.p, pi.data(), pi.size() * 8, cudaMemcpyHostToDevice, stream));
    VDB_CUDA_TRY(cudaStreamSynchronize(stream));  // the host vectors die here
    *npages_out = npages;
    return VDB_OK;
}

}  // namespace

int32_t vdb::index_upload_owners(vdb_index* ix) {
    if (ix->cfg.shard_count <= 1) return VDB_OK;
    VDB_TRY(ix->d_owner.reserve(ix->nlist));
    VDB_CUDA_TRY(cudaMemcpyAsync(ix->d_owner.p, ix->h_owner.data(), ix->nlist, cudaMemcpyHostToDevice, ix->stream));
    VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return VDB_OK;
}

namespace {

// Greedy (largest first) byte balancing of the lists over the shards from per-list row counts: iid data
// clusters very unevenly, so `l % world` can leave one GPU with far more to scan than another.
// Deterministic, so every rank computes the same table from the same counts.
void balance_owners(vdb_index* ix, const std::vector<uint32_t>& counts) { balance_owners_impl(ix, counts); }

}  // namespace

static void balance_owners_impl(vdb_index* ix, const std::vector<uint32_t>& counts) {
    const uint32_t world = ix->cfg.shard_count;
    std::vector<uint32_t> order(ix->nlist);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return counts[a] > counts[b]; });
    std::vector<uint64_t> load(world, 0);
    for (uint32_t l : order) {
        uint32_t best = 0;
        for (uint32_t r = 1; r < world; ++r)
            if (load[r] < load[best]) best = r;
        ix->h_owner[l] = (uint8_t)best;
        load[best] += (uint64_t)counts[l] + 1;  // +1 spreads the empty lists too
    }
}

namespace {

int32_t refresh_centroid_aux(vdb_index* ix) {
    VDB_TRY(ix->cnorm.reserve(ix->nlist));
    VDB_TRY(ix->cmax_bits.reserve(1));
    return centroid_norms(ix->centroids.p, ix->nlist, ix->ld, ix->cnorm.p, ix->cmax_bits.p, ix->stream);
}

int32_t refresh_centroid_view(vdb_index* ix) {
    return build_flat_view(ix->centroids.p, ix->nlist, ix->ld, CENTROID_PAGE_ROWS, ix->c_rows, ix->c_page_off,
                           ix->c_page_vec, ix->c_page_ids, &ix->c_npages, ix->stream);
}

}  // namespace

int32_t vdb::index_refresh_centroids(vdb_index* ix) {
    VDB_TRY(refresh_centroid_view(ix));
    return refresh_centroid_aux(ix);
}

int32_t vdb::index_alloc_page(vdb_index* ix, uint32_t* page) {
    if (ix->pages_used == ix->page_addr.size()) {
        // slabs grow geometrically: 64 MiB first, doubling the resident total, capped at 2 GiB
        const uint64_t want = std::min<uint64_t>(2ull << 30, std::max<uint64_t>(SLAB_BYTES, ix->slab_bytes_total));
        const uint32_t slab_pages = (uint32_t)std::max<uint64_t>(1, want / ix->page_bytes);
        const uint64_t slab = (uint64_t)slab_pages * ix->page_bytes;
        if (ix->cfg.max_gpu_memory && ix->hbm_bytes() + slab > ix->cfg.max_gpu_memory) {
            set_last_error("max_gpu_memory exceeded while growing the inverted lists");
            return VDB_OUT_OF_MEMORY;
        }
        // TransferManager::allocate_device first (the reference's index takes its device memory from the pool it is
        // given, ivf_flat_index.cpp:424-433); a pool that cannot hold the slab -- the reference then skips the GPU and
        // searches on the CPU -- is bypassed with a plain allocation
        void* p = nullptr;
        bool pooled = false;
        if (ix->arena && ix->arena_device == ix->device) {
            p = vdb_arena_allocate_device(ix->arena, slab);
            pooled = p != nullptr;
        }
        if (!p) VDB_CUDA_TRY(cudaMalloc(&p, slab));
        ix->slabs.push_back(p);
        ix->slab_pooled.push_back(pooled);
        ix->slab_bytes_total += slab;
        for (uint32_t i = 0; i < slab_pages; ++i)
            ix->page_addr.push_back((uint64_t)(uintptr_t)p + (uint64_t)i * ix->page_bytes);
    }
    *page = ix->pages_used++;
    return VDB_OK;
}

int32_t vdb::index_upload_list_tables(vdb_index* ix) {
    const uint32_t nlist = ix->nlist;
    std::vector<uint32_t> off(nlist + 1, 0);
    for (uint32_t l = 0; l < nlist; ++l) off[l + 1] = off[l] + (uint32_t)ix->h_pages[l].size();
    const uint32_t npages = off[nlist];
    std::vector<uint64_t> pv(std::max(npages, 1u), 0), pi(std::max(npages, 1u), 0);
    for (uint32_t l = 0; l < nlist; ++l)
        for (size_t j = 0; j < ix->h_pages[l].size(); ++j) {
            const uint64_t a = ix->page_addr[ix->h_pages[l][j]];
            pv[off[l] + j] = a;
            pi[off[l] + j] = a + ix->ids_off;
        }
    VDB_TRY(ix->d_rows.reserve(nlist));
    VDB_TRY(ix->d_page_off.reserve(nlist + 1));
    VDB_TRY(ix->d_page_vec.reserve(pv.size()));
    VDB_TRY(ix->d_page_ids.reserve(pi.size()));
    VDB_CUDA_TRY(cudaMemcpyAsync(ix->d_rows.p, ix->h_rows.data(), nlist * 4, cudaMemcpyHostToDevice, ix->stream));
    VDB_CUDA_TRY(cudaMemcpyAsync(ix->d_page_off.p, off.data(), (nlist + 1) * 4, cudaMemcpyHostToDevice, ix->stream));
    VDB_CUDA_TRY(cudaMemcpyAsync(ix->d_page_vec.p, pv.data(), pv.size() * 8, cudaMemcpyHostToDevice, ix->stream));
    VDB_CUDA_TRY(cudaMemcpyAsync(ix->d_page_ids.p, pi.data(), pi.size() * 8, cudaMemcpyHostToDevice, ix->stream));
    VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
    ix->npages_desc.resize(nlist);
    for (uint32_t l = 0; l < nlist; ++l) ix->npages_desc[l] = (uint32_t)ix->h_pages[l].size();
    std::sort(ix->npages_desc.begin(), ix->npages_desc.end(), std::greater<uint32_t>());
    return VDB_OK;
}

namespace {

// rows [n][dim] at `src` (host or device) -> device [n][ld], zero padded; returns the device pointer
// (src itself when it already is a device array with dim == ld)
int32_t stage_rows(vdb_index* ix, const float* src, uint64_t n, DevBuf<float>& buf, const float** out,
                   cudaStream_t stream) {
    const bool dev = is_device_ptr(src);
    if (dev && ix->dim == ix->ld && !((uintptr_t)src & 15)) {
        *out = src;
        return VDB_OK;
    }
    VDB_TRY(buf.reserve((size_t)n * ix->ld));
    if (dev) {
        VDB_TRY(launch_pad_rows(src, ix->dim, ix->dim, buf.p, ix->ld, n, stream));
    } else if (ix->dim == ix->ld) {
        VDB_CUDA_TRY(cudaMemcpyAsync(buf.p, src, (size_t)n * ix->ld * 4, cudaMemcpyHostToDevice, stream));
    } else {
        VDB_CUDA_TRY(cudaMemsetAsync(buf.p, 0, (size_t)n * ix->ld * 4, stream));
        VDB_CUDA_TRY(cudaMemcpy2DAsync(buf.p, (size_t)ix->ld * 4, src, (size_t)ix->dim * 4, (size_t)ix->dim * 4, n,
                                       cudaMemcpyHostToDevice, stream));
    }
    *out = buf.p;
    return VDB_OK;
}

uint64_t slot_bound(const vdb_index* ix, uint32_t nq, uint32_t np, uint32_t ppi) {
    uint64_t s = 0;
    for (uint32_t i = 0; i < np && i < ix->npages_desc.size(); ++i) s += (ix->npages_desc[i] + ppi - 1) / ppi;
    return s * nq;
}

// pages per scan item and the number of queries one pass may take so that the partial-result buffer
// (slots x k x 12 bytes) stays under PARTIAL_BYTES_CAP.  Longer runs of a list amortise the per-item costs (tile
// announcement, barriers, final selection, partial write-out, merge input) over up to ~3 MB.  Measured: 4 pages
// beat 1 and 2 even on a 1/8 shard of the headline index and 8 loses to the longer tail, so 4 it is whenever that
// still leaves a handful of items per SM.  Widening stops at the longest list (beyond it the bound no longer
// falls: every non-empty probed list keeps one range); what still does not fit is split over query chunks.
void choose_ppi(const vdb_index* ix, uint32_t nq, uint32_t np, uint32_t k, uint32_t* ppi_out, uint32_t* nq_chunk) {
    uint32_t ppi = std::max<uint32_t>(1, std::min<uint32_t>(4, ix->pages_used / (NUM_SMS_B200 * 4u)));
    if (ppi == 3) ppi = 2;
    if (ix->ppi_override) ppi = ix->ppi_override;
    const uint32_t longest = ix->npages_desc.empty() ? 1u : std::max(1u, ix->npages_desc[0]);
    while (slot_bound(ix, nq, np, ppi) * k * 12 > PARTIAL_BYTES_CAP && ppi < longest) ppi *= 2;
    const uint64_t per_query = std::max<uint64_t>(1, slot_bound(ix, 1, np, ppi)) * k * 12;
    const uint64_t fit = std::max<uint64_t>(1, PARTIAL_BYTES_CAP / per_query);
    *ppi_out = ppi;
    *nq_chunk = (uint32_t)std::min<uint64_t>(nq, fit);
}

// coarse: top-np centroids of every query = select_nprobe_lists (ivf_flat_index.cpp:298-336) -> s.probes
int32_t slot_coarse_select(vdb_index* ix, SearchSlot& s, const float* q_dev, uint32_t nq, uint32_t np,
                      cudaStream_t stream) {
    const bool tensor = ix->cfg.coarse_mode == VDB_COARSE_TENSOR ||
                        (ix->cfg.coarse_mode == VDB_COARSE_AUTO && ix->nlist >= 256);
    VDB_TRY(s.coarse_d.reserve((size_t)nq * np));
    VDB_TRY(s.probes.reserve((size_t)nq * np));
    if (tensor && coarse_tensor_supported(ix->nlist, ix->ld, np)) {
        const uint32_t ldd = round_up(ix->nlist, 4);
        VDB_TRY(s.dots.reserve((size_t)nq * ldd));
        VDB_TRY(score_gemm(q_dev, nq, ix->ld, ix->centroids.p, ix->nlist, ix->ld, ix->ld, s.dots.p, ldd, stream));
        return vdb::coarse_select(s.dots.p, ldd, q_dev, nq, ix->centroids.p, ix->cnorm.p, ix->cmax_bits.p, ix->nlist,
                                  ix->ld, np, ix->cfg.metric, s.probes.p, s.coarse_d.p, nullptr, stream);
    }
    if (ix->cfg.coarse_mode == VDB_COARSE_TENSOR) {
        set_last_error("coarse_mode TENSOR requested but the tensor-core path does not support this shape/driver");
        return VDB_INVALID_ARGUMENT;
    }
    if (np > (uint32_t)scan_max_k()) {  // wider than the top-k machinery of the other two modes
        VDB_REQUIRE(coarse_wide_supported(ix->nlist, ix->ld), "nprobe > 2048 needs nlist <= 16384");
        return coarse_select_wide(q_dev, nq, ix->centroids.p, ix->nlist, ix->ld, np, ix->cfg.metric, s.probes.p,
                                  s.coarse_d.p, stream);
    }
    VDB_TRY(s.zero_probes.reserve(nq));
    VDB_CUDA_TRY(cudaMemsetAsync(s.zero_probes.p, 0, (size_t)nq * 4, stream));
    VDB_TRY(s.coarse_i.reserve((size_t)nq * np));
    // keep the partial-result buffer below ~64 MiB by widening the page range per item (one range per query at most)
    uint32_t ppi = 1;
    while ((uint64_t)nq * ((ix->c_npages + ppi - 1) / ppi) * np * 12 > (64ull << 20) && ppi < ix->c_npages) ppi *= 2;
    const uint64_t slots = (uint64_t)nq * ((ix->c_npages + ppi - 1) / ppi);
    return scan_search(centroid_table(ix), q_dev, nq, s.zero_probes.p, 1, np, ix->cfg.metric, ppi, slots,
                       s.ws_coarse, false, s.coarse_d.p, s.coarse_i.p, s.probes.p, stream);
}

int32_t ensure_slot_events(SearchSlot& s) {
    if (!s.ev_done) {
        VDB_CUDA_TRY(cudaEventCreateWithFlags(&s.ev_front, cudaEventDisableTiming));
        VDB_CUDA_TRY(cudaEventCreateWithFlags(&s.ev_scan, cudaEventDisableTiming));
        VDB_CUDA_TRY(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    }
    return VDB_OK;
}

int32_t ensure_slot_timers(SearchSlot& s) {
    if (!s.tm[0])
        for (uint32_t i = 0; i < SLOT_TIMERS; ++i) VDB_CUDA_TRY(cudaEventCreate(&s.tm[i]));
    return VDB_OK;
}

// add the phase times of a completed search to the index's sums
void harvest_timers(vdb_index* ix, SearchSlot& s) {
    if (!s.timed) return;
    s.timed = false;
    static const int pairs[5][2] = {{0, 1}, {1, 2}, {3, 4}, {5, 6}, {6, 7}};
    for (int i = 0; i < 5; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.tm[pairs[i][0]], s.tm[pairs[i][1]]) == cudaSuccess) ix->prof_ms[i] += ms;
        else cudaGetLastError();
    }
    ++ix->prof_searches;
}

}  // namespace

// results of slot `s` (device) -> the caller's host arrays or the slot's pinned staging, behind everything on `st`
static int32_t enqueue_delivery(SearchSlot& s, cudaStream_t st) {
    if (s.host_d) {
        VDB_CUDA_TRY(cudaMemcpyAsync(s.host_d, s.dev_d, s.out_elems * 4, cudaMemcpyDeviceToHost, st));
        VDB_CUDA_TRY(cudaMemcpyAsync(s.host_i, s.dev_i, s.out_elems * 8, cudaMemcpyDeviceToHost, st));
    }
    return VDB_OK;
}

// enqueue the collect of the batch whose publish went out last (if it is still owed), its delivery and its
// completion event
static int32_t flush_deferred_collect(vdb_index* ix, cudaStream_t st) {
    SearchSlot* d = ix->deferred;
    if (!d) return VDB_OK;
    ix->deferred = nullptr;
    d->collect_deferred = false;
    VDB_TRY(exchange_collect(ix->exchange, d->dev_d, d->dev_i, st));
    if (d->timed) cudaEventRecord(d->tm[7], st);
    VDB_TRY(enqueue_delivery(*d, st));
    VDB_CUDA_TRY(cudaEventRecord(d->ev_done, st));
    return VDB_OK;
}

int32_t vdb::index_flush_deferred(vdb_index* ix, SearchSlot& s) {
    if (!s.collect_deferred) return VDB_OK;
    DeviceGuard g(ix->device);
    return flush_deferred_collect(ix, s.collect_stream);
}

// Block the host until the search in slot `s` is complete, hand host results to the caller's arrays, report a
// peer-exchange timeout of this search, and make the slot reusable.  Caller holds ix->mu or owns the ticket.
int32_t vdb::index_finish_slot(vdb_index* ix, SearchSlot& s) {
    if (!s.busy) return VDB_OK;
    DeviceGuard g(ix->device);
    if (s.collect_deferred) VDB_TRY(flush_deferred_collect(ix, s.collect_stream));  // nobody submitted after it
    VDB_CUDA_TRY(cudaEventSynchronize(s.ev_done));
    s.busy = false;
    harvest_timers(ix, s);
    if (s.deliver) {
        std::memcpy(s.user_d, s.h_d.p, s.out_elems * 4);
        std::memcpy(s.user_i, s.h_i.p, s.out_elems * 8);
        s.deliver = false;
    }
    if (s.used_exchange && ix->exchange) return vdb_exchange_status(ix->exchange);
    return VDB_OK;
}

int32_t vdb::index_acquire_slot(vdb_index* ix, SearchSlot** out, uint64_t* ticket) {
    if (ix->tables_dirty) {
        ix->tables_dirty = false;
        VDB_TRY(index_upload_list_tables(ix));
    }
    const uint64_t t = ++ix->next_ticket;
    SearchSlot& s = ix->slots[t % ix->depth];
    VDB_TRY(ensure_slot_events(s));
    const int32_t st = index_finish_slot(ix, s);  // the search that used this slot `depth` tickets ago
    s.ticket = t;
    ix->last_slot = (int)(t % ix->depth);
    *out = &s;
    *ticket = t;
    return st;
}

// One search, three phases (SearchSlot).  `queries`: host or device, [nq][dim]; a device array must already hold
// its values (or be ordered before st.front by the caller).  distances / indices: host or device, or null when
// the merged local result is only published (collect = false).
int32_t vdb::index_enqueue_search(vdb_index* ix, SearchSlot& s, const float* queries, uint32_t nq, uint32_t nprobe,
                                  uint32_t k, float* distances, uint64_t* indices, const SearchStreams& st,
                                  bool collect) {
    const uint32_t np = std::min(nprobe, ix->nlist);  // the reference reads past probe_lists instead (:221-222)
    const bool prof = ix->profiling;
    if (prof) VDB_TRY(ensure_slot_timers(s));
    vdb_exchange* ex = ix->exchange;
    const bool out_dev = distances ? is_device_ptr(distances) : true;
    if (distances) VDB_REQUIRE(out_dev == is_device_ptr(indices), "search: distances and indices must live on the same side");
    VDB_REQUIRE(collect || ex, "search: publishing without an exchange");

    // ---- front: queries -> [nq][ld] on the device, coarse selection, probe grouping
    nvtxRangePushA("vdb.search.front");
    if (prof) cudaEventRecord(s.tm[0], st.front);
    const float* q = nullptr;
    const bool q_dev = is_device_ptr(queries);
    if (q_dev && ix->dim == ix->ld && !((uintptr_t)queries & 15)) {
        q = queries;
    } else {
        VDB_TRY(s.q_buf.reserve((size_t)nq * ix->ld));
        if (q_dev) {
            VDB_TRY(launch_pad_rows(queries, ix->dim, ix->dim, s.q_buf.p, ix->ld, nq, st.front));
        } else {
            const float* src = queries;
            if (!is_pinned_ptr(queries)) {  // pageable memory would make the copy synchronous: stage it
                VDB_TRY(s.h_q.reserve((size_t)nq * ix->dim * 4));
                std::memcpy(s.h_q.p, queries, (size_t)nq * ix->dim * 4);
                src = static_cast<const float*>(s.h_q.p);
            }
            if (ix->dim == ix->ld) {
                VDB_CUDA_TRY(cudaMemcpyAsync(s.q_buf.p, src, (size_t)nq * ix->ld * 4, cudaMemcpyHostToDevice, st.front));
            } else {
                VDB_CUDA_TRY(cudaMemsetAsync(s.q_buf.p, 0, (size_t)nq * ix->ld * 4, st.front));
                VDB_CUDA_TRY(cudaMemcpy2DAsync(s.q_buf.p, (size_t)ix->ld * 4, src, (size_t)ix->dim * 4,
                                               (size_t)ix->dim * 4, nq, cudaMemcpyHostToDevice, st.front));
            }
        }
        q = s.q_buf.p;
    }
    VDB_TRY(slot_coarse_select(ix, s, q, nq, np, st.front));
    if (prof) cudaEventRecord(s.tm[1], st.front);
    uint32_t ppi = 1, nq_chunk = nq;
    choose_ppi(ix, nq, np, k, &ppi, &nq_chunk);
    VDB_REQUIRE(nq_chunk == nq, "search: internal error, batch not chunked");
    // a pipelined scan leaves a few SMs to the front / back kernels of the neighbouring batches
    int sms = NUM_SMS_B200;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
    const uint32_t max_ctas = st.split && (uint32_t)sms > 2 * ix->reserve_sms ? (uint32_t)sms - ix->reserve_sms : 0u;
    ScanPlan plan;
    VDB_TRY(scan_plan(list_table(ix), q, nq, s.probes.p, np, k, ix->cfg.metric, ppi, slot_bound(ix, nq, np, ppi), true,
                      max_ctas, s.ws_scan, &plan));
    plan.has_norms = !ix->scan_exact;
    if (ix->scan_exact) {  // VDB_SCAN_EXACT=1: the plain fp32 scan (A/B and parity checks)
        plan.mirror = false;
        plan.info.mirror = 0;
    }
    plan.lifetime_rows = ix->d_scanned;
    plan.dot_min_rows = ix->dot_min_rows;
    if (!ix->ppi_override) plan.ppi_max = std::max(ppi, 16u);  // an explicit VDB_SCAN_PPI is taken literally
    VDB_TRY(scan_enqueue_groups(plan, s.ws_scan, st.front));
    if (prof) cudaEventRecord(s.tm[2], st.front);
    if (st.split) {
        VDB_CUDA_TRY(cudaEventRecord(s.ev_front, st.front));
        VDB_CUDA_TRY(cudaStreamWaitEvent(st.scan, s.ev_front, 0));
    }
    nvtxRangePop();

    // ---- scan
    nvtxRangePushA("vdb.search.scan");
    if (prof) {
        if (!ix->span_open) {
            cudaEventRecord(ix->span_start, st.scan);
            ix->span_open = true;
        }
        cudaEventRecord(s.tm[3], st.scan);
    }
    VDB_TRY(scan_enqueue_scan(plan, s.ws_scan, st.scan));
    if (prof) {
        cudaEventRecord(s.tm[4], st.scan);
        cudaEventRecord(ix->span_end, st.scan);
    }
    if (st.split) {
        VDB_CUDA_TRY(cudaEventRecord(s.ev_scan, st.scan));
        VDB_CUDA_TRY(cudaStreamWaitEvent(st.back, s.ev_scan, 0));
    }
    nvtxRangePop();

    // ---- back: merge (+ publish into the peers' mailboxes, + collect), results to the caller
    nvtxRangePushA("vdb.search.back");
    if (st.back_wait) VDB_CUDA_TRY(cudaStreamWaitEvent(st.back, st.back_wait, 0));
    if (prof) cudaEventRecord(s.tm[5], st.back);
    float* od = distances;
    uint64_t* oi = indices;
    if (collect && !out_dev) {
        VDB_TRY(s.out_d.reserve((size_t)nq * k));
        VDB_TRY(s.out_i.reserve((size_t)nq * k));
        od = s.out_d.p;
        oi = s.out_i.p;
    }
    s.used_exchange = ex != nullptr;
    s.deliver = false;
    s.dev_d = od;
    s.dev_i = oi;
    s.host_d = nullptr;
    s.host_i = nullptr;
    s.out_elems = (size_t)nq * k;
    if (collect && !out_dev) {
        s.host_d = distances;
        s.host_i = indices;
        if (!is_pinned_ptr(distances) || !is_pinned_ptr(indices)) {
            VDB_TRY(s.h_d.reserve(s.out_elems * 4));
            VDB_TRY(s.h_i.reserve(s.out_elems * 8));
            s.host_d = static_cast<float*>(s.h_d.p);
            s.host_i = static_cast<uint64_t*>(s.h_i.p);
            s.user_d = distances;
            s.user_i = indices;
            s.deliver = true;
        }
    }
    s.collect_deferred = false;
    if (ex) {
        // The previous batch's collect goes HERE, behind this batch's scan: by now every peer has long published it,
        // so the collect kernel finds its flags set instead of spinning on the SMs the front / back kernels need
        // (measured at 8 GPUs: collecting right behind the publish kept 16 CTAs polling for 0.3 ms of every 0.57 ms
        // step).  It still precedes this batch's publish on every rank, which is what keeps two mailbox halves enough.
        VDB_TRY(flush_deferred_collect(ix, st.back));
        PublishTarget pub;
        VDB_TRY(exchange_begin_publish(ex, nq, k, &pub));
        VDB_TRY(scan_enqueue_merge(plan, s.ws_scan, nullptr, nullptr, nullptr, &pub, st.back));
        if (prof) cudaEventRecord(s.tm[6], st.back);
        s.busy = true;
        s.timed = prof;
        s.info = plan.info;
        nvtxRangePop();
        if (!collect) {  // a non-root shard of a single-process index: published, nothing to collect
            if (prof) cudaEventRecord(s.tm[7], st.back);
            VDB_CUDA_TRY(cudaEventRecord(s.ev_done, st.back));
            return VDB_OK;
        }
        s.collect_deferred = true;
        s.collect_stream = st.back;
        ix->deferred = &s;
        // stream-ordered callers (vdb_index_search_async) get their result in stream order: no deferral
        if (!st.split) VDB_TRY(flush_deferred_collect(ix, st.back));
        return VDB_OK;
    }
    VDB_TRY(scan_enqueue_merge(plan, s.ws_scan, od, oi, nullptr, nullptr, st.back));
    if (prof) {
        cudaEventRecord(s.tm[6], st.back);
        cudaEventRecord(s.tm[7], st.back);
    }
    VDB_TRY(enqueue_delivery(s, st.back));
    VDB_CUDA_TRY(cudaEventRecord(s.ev_done, st.back));
    nvtxRangePop();
    s.busy = true;
    s.timed = prof;
    s.info = plan.info;
    return VDB_OK;
}

// vdb_index_append_list defers the upload of the list tables (one per stored list would be quadratic)
static int32_t flush_tables(vdb_index* ix) {
    if (!ix->tables_dirty) return VDB_OK;
    ix->tables_dirty = false;
    return index_upload_list_tables(ix);
}

SearchStreams vdb::index_pipeline_streams(vdb_index* ix, uint64_t ticket) {
    return SearchStreams{ix->s_front, ix->s_scan[ticket & 1], ix->s_back, true, nullptr};
}

void vdb::index_choose_ppi(const vdb_index* ix, uint32_t nq, uint32_t np, uint32_t k, uint32_t* ppi,
                           uint32_t* nq_chunk) {
    choose_ppi(ix, nq, np, k, ppi, nq_chunk);
}

namespace {

// every slot's buffers for searches of up to (rs_nq, rs_np, rs_k), so that no cudaMalloc (an implicit device
// synchronisation) happens inside a search; add() repeats it because longer lists mean more partial-result slots
int32_t reserve_slots(vdb_index* ix) {
    if (!ix->rs_nq) return VDB_OK;
    const uint32_t nq = ix->rs_nq, np = std::min(ix->rs_np, ix->nlist), k = ix->rs_k;
    uint32_t ppi = 1, nq_chunk = nq;
    choose_ppi(ix, nq, np, k, &ppi, &nq_chunk);
    const uint64_t slots = std::max<uint64_t>(1, slot_bound(ix, nq_chunk, np, ppi));
    for (uint32_t i = 0; i < ix->depth; ++i) {
        SearchSlot& s = ix->slots[i];
        VDB_TRY(index_finish_slot(ix, s));
        VDB_TRY(ensure_slot_events(s));
        VDB_TRY(ensure_slot_timers(s));
        VDB_TRY(s.q_buf.reserve((size_t)nq * ix->ld));
        VDB_TRY(s.dots.reserve((size_t)nq * round_up(ix->nlist, 4)));
        VDB_TRY(s.coarse_d.reserve((size_t)nq * np));
        VDB_TRY(s.probes.reserve((size_t)nq * np));
        VDB_TRY(s.out_d.reserve((size_t)nq * k));
        VDB_TRY(s.out_i.reserve((size_t)nq * k));
        VDB_TRY(s.h_q.reserve((size_t)nq * ix->dim * 4));
        VDB_TRY(s.h_d.reserve((size_t)nq * k * 4));
        VDB_TRY(s.h_i.reserve((size_t)nq * k * 8));
        VDB_TRY(s.ws_scan.reserve(ix->nlist, nq_chunk * np, slots, k, nq_chunk));
    }
    return VDB_OK;
}

// the shapes one pass cannot (partial-result buffer) or should not (below) take are split over query chunks, each a
// pass of its own; the chunks ride the search pipeline like separately submitted batches (`depth` in flight)
int32_t search_chunked(vdb_index* ix, const float* queries, uint32_t nq, uint32_t nprobe, uint32_t k, float* distances,
                       uint64_t* indices, uint32_t nq_chunk) {
    int32_t st = VDB_OK;
    for (uint32_t lo = 0; lo < nq && st == VDB_OK; lo += nq_chunk) {
        const uint32_t m = std::min(nq_chunk, nq - lo);
        SearchSlot* s = nullptr;
        uint64_t t = 0;
        st = index_acquire_slot(ix, &s, &t);  // waits for (and delivers) the chunk that used this slot `depth` ago
        if (st == VDB_OK)
            st = index_enqueue_search(ix, *s, queries + (size_t)lo * ix->dim, m, nprobe, k, distances + (size_t)lo * k,
                                      indices + (size_t)lo * k, index_pipeline_streams(ix, t), true);
    }
    for (uint32_t i = 0; i < ix->depth; ++i) {  // the chunks still in flight (also after an error: nothing may stay busy)
        const int32_t fs = index_finish_slot(ix, ix->slots[i]);
        if (st == VDB_OK) st = fs;
    }
    return st;
}

// An index with a shadow answers a batch wider than the screen kernel's 64 queries chunk by chunk: the
// chunks stream the int8 / bf16 shadow and do their dot products on the tensor cores, where one wide pass would take
// the fp32 kernel and be compute-bound (10 000 queries -- the reference's bench/benchmark.cpp -- are 157 chunks of
// ~1.1 ms against tens of seconds of fp32 FMAs).  Same results either way.
uint32_t screen_chunk(const vdb_index* ix, uint32_t nq, uint32_t k, uint32_t nq_chunk) {
    if (ix->mirror_off && !ix->scan_exact && nq > screen_max_batch() && k <= screen_max_k())
        return std::min(nq_chunk, screen_max_batch());
    return nq_chunk;
}

int32_t check_search_args(vdb_index* ix, const void* q, const void* d, const void* i, uint32_t nq, uint32_t nprobe,
                          uint32_t k) {
    VDB_REQUIRE(q && d && i, "search: null buffer");
    VDB_REQUIRE(nq >= 1 && k >= 1 && nprobe >= 1, "search: nq, k and nprobe must be >= 1");
    VDB_REQUIRE(k <= (uint32_t)scan_max_k(), "search: k must be <= 2048");
    (void)ix;
    return VDB_OK;
}

int32_t check_index(vdb_index* ix) {
    if (!ix) {
        set_last_error("null index handle");
        return VDB_INVALID_ARGUMENT;
    }
    return VDB_OK;
}

}  // namespace

extern "C" {

const char* vdb_last_error_string(void) { return g_last_error.c_str(); }

const char* vdb_status_string(int32_t s) {
    switch (s) {
        case VDB_OK: return "OK";
        case VDB_INVALID_ARGUMENT: return "INVALID_ARGUMENT";
        case VDB_OUT_OF_MEMORY: return "OUT_OF_MEMORY";
        case VDB_CUDA_ERROR: return "CUDA_ERROR";
        case VDB_NCCL_ERROR: return "NCCL_ERROR";
        case VDB_NOT_TRAINED: return "NOT_TRAINED";
        default: return "INTERNAL";
    }
}

int32_t vdb_version(void) { return 100; }

void vdb_config_default(vdb_config* cfg) {
    if (!cfg) return;
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->metric = VDB_METRIC_L2;
    cfg->max_gpu_memory = 0;
    cfg->train_mode = VDB_TRAIN_AUTO;
    cfg->coarse_mode = VDB_COARSE_AUTO;
    cfg->shard_count = 1;
}

int32_t vdb_index_create(const vdb_config* cfg, vdb_index** out) {
    VDB_REQUIRE(cfg && out, "null config or output handle");
    // ctor contract: dimension and nlist must be > 0 (ivf_flat_index.cpp:17-19)
    VDB_REQUIRE(cfg->dimension > 0 && cfg->nlist > 0, "Invalid configuration: dimension and nlist must be > 0");
    VDB_REQUIRE(cfg->dimension <= 2048, "dimension must be <= 2048 (kernels.cuh MAX_DIM)");
    VDB_REQUIRE(cfg->metric == VDB_METRIC_L2 || cfg->metric == VDB_METRIC_IP,
                "metric must be L2 or InnerProduct (the reference CPU path computes 0 for Cosine)");
    VDB_REQUIRE(cfg->shard_count >= 1 && cfg->shard_count <= 255 && cfg->shard_rank < cfg->shard_count,
                "bad shard rank/count");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_last_error("no CUDA device: this library has no CPU path");
        return VDB_CUDA_ERROR;
    }
    VDB_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "device ordinal out of range");
    std::unique_ptr<vdb_index> ix(new vdb_index());
    ix->cfg = *cfg;
    ix->dim = cfg->dimension;
    ix->ld = round_up(cfg->dimension, 4);
    ix->nlist = cfg->nlist;
    ix->device = cfg->device;
    uint32_t pr = cfg->page_rows;
    if (pr == 0) {
        const uint64_t target = ix->ld >= 256 ? (768ull << 10) : (256ull << 10);
        pr = (uint32_t)std::max<uint64_t>(16, target / (ix->ld * 4ull) / 16 * 16);
        // row widths the bf16 screen supports get whole 128-row operand tiles (1024-D: 192 -> 256 rows)
        if (screen_supported(ix->ld, MIRROR_TILE_ROWS, cfg->metric)) pr = round_up(pr, MIRROR_TILE_ROWS);
    }
    VDB_REQUIRE(pr % 16 == 0 && pr <= 65536, "page_rows must be a multiple of 16, <= 65536");
    ix->page_rows = pr;
    ix->ids_off = (uint64_t)pr * ix->ld * 4;
    // page = [pr][ld] fp32 rows | [pr] u64 ids | [pr] fp32 |row|^2 (the dot-form screen of the L2 scan)
    ix->page_bytes = (ix->ids_off + (uint64_t)pr * 12 + 255) / 256 * 256;
    // ... | [pr] fp32 |row - shadow(row)| | [pr] fp32 row scales | the rows once more in low precision, as tensor-core
    // operand tiles: the scan's screen (screen.cuh) streams the shadow instead of the fp32 rows.  scan_mirror: 0 = auto
    // (int8 where the screen kernel supports the shape), 1 = off, 2 = bf16 (+50 % HBM, half the bytes per search),
    // 3 = int8 with one scale per row (+25 % HBM, a quarter of the bytes); 2 / 3 are refused where unsupported.
    // VDB_SCAN_MIRROR overrides: 0 = off, 1 = bf16, 2 = int8.
    {
        uint32_t want = cfg->scan_mirror;
        if (const char* e = std::getenv("VDB_SCAN_MIRROR")) {
            const int v = std::atoi(e);
            want = v == 0 ? 1u : v == 2 ? 3u : 2u;
        }
        const bool can = screen_supported(ix->ld, pr, cfg->metric);
        VDB_REQUIRE(want <= 3, "scan_mirror must be 0 (auto), 1 (off), 2 (bf16) or 3 (int8)");
        VDB_REQUIRE(want < 2 || can, "scan_mirror: the screen needs a row stride of 128 * {1,2,4,6,8} floats and page_rows % 128 == 0");
        if (can && want != 1) {
            ix->mirror_kind = want == 2 ? MIRROR_BF16 : MIRROR_I8;  // auto: int8 (less memory, fewer bytes per search)
            ix->mirror_off = (uint32_t)((ix->ids_off + (uint64_t)pr * 20 + 1023) / 1024 * 1024);
            ix->page_bytes = ((uint64_t)ix->mirror_off + (uint64_t)pr * ix->ld * mirror_elem_bytes(ix->mirror_kind) + 255) / 256 * 256;
        }
    }
    ix->pages_per_slab = (uint32_t)std::max<uint64_t>(1, SLAB_BYTES / ix->page_bytes);
    ix->h_rows.assign(ix->nlist, 0);
    ix->h_pages.resize(ix->nlist);
    DeviceGuard g(ix->device);
    VDB_CUDA_TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    // search pipeline: the short front / back kernels run at high priority so that they take the first SM a
    // scan CTA frees; consecutive scans alternate between two streams so that batch i+1 fills the tail of batch i
    int prio_lo = 0, prio_hi = 0;
    VDB_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    VDB_CUDA_TRY(cudaStreamCreateWithPriority(&ix->s_front, cudaStreamNonBlocking, prio_hi));
    VDB_CUDA_TRY(cudaStreamCreateWithPriority(&ix->s_back, cudaStreamNonBlocking, prio_hi));
    VDB_CUDA_TRY(cudaStreamCreateWithPriority(&ix->s_scan[0], cudaStreamNonBlocking, prio_lo));
    VDB_CUDA_TRY(cudaStreamCreateWithPriority(&ix->s_scan[1], cudaStreamNonBlocking, prio_lo));
    VDB_CUDA_TRY(cudaMalloc(&ix->d_scanned, 8));
    VDB_CUDA_TRY(cudaMemset(ix->d_scanned, 0, 8));
    VDB_CUDA_TRY(cudaEventCreate(&ix->span_start));
    VDB_CUDA_TRY(cudaEventCreate(&ix->span_end));
    ix->depth = cfg->pipeline_depth ? cfg->pipeline_depth : 4;
    ix->reserve_sms = cfg->reserve_sms == 0xffffffffu ? 0 : cfg->reserve_sms ? cfg->reserve_sms : 8;
    // tuning knobs, read once: pages per scan item, pipeline depth, SMs left to the side kernels
    if (const char* e = std::getenv("VDB_SCAN_PPI")) {
        const int v = std::atoi(e);
        if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32 || v == 64) ix->ppi_override = (uint32_t)v;
    }
    if (const char* e = std::getenv("VDB_PIPELINE_DEPTH")) ix->depth = (uint32_t)std::atoi(e);
    if (const char* e = std::getenv("VDB_SCAN_EXACT")) ix->scan_exact = std::atoi(e) != 0;  // A/B: no dot-form screen
    if (const char* e = std::getenv("VDB_SCAN_DOT_MIN_ROWS")) ix->dot_min_rows = (uint32_t)std::atoi(e);  // tests: 0
    if (const char* e = std::getenv("VDB_RESERVE_SMS")) ix->reserve_sms = (uint32_t)std::atoi(e);
    VDB_REQUIRE(ix->depth >= 1 && ix->depth <= MAX_SEARCH_SLOTS, "pipeline_depth must be in [1, 8]");
    VDB_REQUIRE(ix->reserve_sms <= 64, "reserve_sms must be <= 64");
    VDB_TRY(ix->centroids.reserve((size_t)ix->nlist * ix->ld));
    VDB_CUDA_TRY(cudaMemsetAsync(ix->centroids.p, 0, ix->centroids.bytes(), ix->stream));  // centroids_ value-init (:22)
    if (cfg->shard_count > 1) {
        ix->h_owner.resize(ix->nlist);
        for (uint32_t l = 0; l < ix->nlist; ++l) ix->h_owner[l] = (uint8_t)(l % cfg->shard_count);
        VDB_TRY(index_upload_owners(ix.get()));
    }
    VDB_TRY(refresh_centroid_view(ix.get()));
    VDB_TRY(refresh_centroid_aux(ix.get()));
    VDB_TRY(index_upload_list_tables(ix.get()));
    *out = ix.release();
    return VDB_OK;
}

int32_t vdb_index_destroy(vdb_index* ix) {
    if (ix && ix->composite) return composite_destroy(ix);
    if (!ix) return VDB_OK;
    {
        DeviceGuard g(ix->device);
        cudaDeviceSynchronize();
        for (size_t i = 0; i < ix->slabs.size(); ++i) {
            if (ix->slab_pooled[i]) vdb_arena_free_device(ix->arena, ix->slabs[i]);
            else cudaFree(ix->slabs[i]);
        }
        ix->centroids.release(); ix->cnorm.release(); ix->cmax_bits.release(); ix->c_rows.release();
        ix->c_page_off.release(); ix->c_page_vec.release(); ix->c_page_ids.release(); ix->d_rows.release();
        ix->d_page_off.release(); ix->d_page_vec.release(); ix->d_page_ids.release();
        ix->assign_buf.release(); ix->hist_buf.release(); ix->fill_buf.release(); ix->stage_buf.release();
        ix->ids_stage.release(); ix->d_owner.release();
        ix->tc_assign.release();
        SearchSlot* all_slots[MAX_SEARCH_SLOTS + 1];
        for (uint32_t i = 0; i < MAX_SEARCH_SLOTS; ++i) all_slots[i] = &ix->slots[i];
        all_slots[MAX_SEARCH_SLOTS] = &ix->aux_slot;
        for (SearchSlot* sp : all_slots) {
            SearchSlot& s = *sp;
            s.q_raw.release();
            s.ws_scan.release(); s.ws_coarse.release();
            s.q_buf.release(); s.dots.release(); s.coarse_d.release(); s.out_d.release();
            s.coarse_i.release(); s.out_i.release(); s.probes.release(); s.zero_probes.release();
            s.h_q.release(); s.h_d.release(); s.h_i.release();
            if (s.ev_front) cudaEventDestroy(s.ev_front);
            if (s.ev_scan) cudaEventDestroy(s.ev_scan);
            if (s.ev_done) cudaEventDestroy(s.ev_done);
            for (auto e : s.tm)
                if (e) cudaEventDestroy(e);
        }
        cudaFree(ix->d_scanned);
        if (ix->span_start) cudaEventDestroy(ix->span_start);
        if (ix->span_end) cudaEventDestroy(ix->span_end);
        for (cudaStream_t st : {ix->stream, ix->s_front, ix->s_back, ix->s_scan[0], ix->s_scan[1]})
            if (st) cudaStreamDestroy(st);
    }
    delete ix;
    return VDB_OK;
}

int32_t vdb_index_train(vdb_index* ix, const float* vectors, uint64_t n) {
    if (ix && ix->composite) {
        VDB_REQUIRE(vectors && n >= 1, "train: no vectors");
        return composite_train(ix, vectors, n);
    }
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(vectors && n >= 1, "train: no vectors");
    VDB_REQUIRE(n < 0xffffffffull, "train: at most 2^32-2 training vectors");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    const float* x = nullptr;
    DevBuf<float> train_buf;
    int32_t st = stage_rows(ix, vectors, n, train_buf, &x, ix->stream);
    KMeansScratch sc;
    if (st == VDB_OK) st = sc.reserve((uint32_t)n, ix->nlist, ix->ld);
    const uint32_t ldx = ix->ld;
    // k-means++ seeding, always L2 (ivf_flat_index.cpp:63-104)
    // AUTO: the reference's sequential fp32 sampling sums, evaluated in parallel (bit-exact at any size);
    // EXACT: the literal one-lane chain (and the scalar assignment kernel) -- the checker for AUTO; FAST: fp64 sums
    const SeedSampler sampler = ix->cfg.train_mode == VDB_TRAIN_EXACT  ? SeedSampler::Sequential
                                : ix->cfg.train_mode == VDB_TRAIN_FAST ? SeedSampler::Fast
                                                                       : SeedSampler::ExactParallel;
    if (st == VDB_OK)
        st = kmeanspp_seed(x, (uint32_t)n, ldx, ix->dim, ix->ld, ix->nlist, ix->centroids.p, sc, sampler, ix->stream);
    // exactly 10 Lloyd iterations; assignment honours the index metric (:109-142, :275-285)
    for (int iter = 0; iter < 10 && st == VDB_OK; ++iter) {
        st = assign_rows(ix, x, n, sc.assign, ix->stream);
        if (st == VDB_OK)
            st = kmeans_update_exact(x, (uint32_t)n, ldx, sc.assign, ix->nlist, ix->ld, ix->centroids.p, sc,
                                     ix->stream);
    }
    if (st == VDB_OK) st = refresh_centroid_aux(ix);
    if (st == VDB_OK && ix->cfg.shard_count > 1 && ix->total_vectors == 0) {
        // list sizes of the training sample under the final centroids -> byte-balanced list ownership
        std::vector<uint32_t> counts(ix->nlist, 0);
        st = assign_rows(ix, x, n, sc.assign, ix->stream);
        if (st == VDB_OK) st = ix->hist_buf.reserve(ix->nlist + 1);
        if (st == VDB_OK && cudaMemsetAsync(ix->hist_buf.p, 0, (ix->nlist + 1) * 4, ix->stream) != cudaSuccess) st = VDB_CUDA_ERROR;
        if (st == VDB_OK) st = launch_hist(sc.assign, n, ix->nlist, 0, nullptr, ix->hist_buf.p, ix->stream);
        if (st == VDB_OK && cudaMemcpyAsync(counts.data(), ix->hist_buf.p, ix->nlist * 4, cudaMemcpyDeviceToHost,
                                            ix->stream) != cudaSuccess)
            st = VDB_CUDA_ERROR;
        if (st == VDB_OK && cudaStreamSynchronize(ix->stream) != cudaSuccess) st = VDB_CUDA_ERROR;
        if (st == VDB_OK) {
            balance_owners(ix, counts);
            st = index_upload_owners(ix);
        }
    }
    if (st == VDB_OK && cudaStreamSynchronize(ix->stream) != cudaSuccess) {
        set_last_error(std::string("train: ") + cudaGetErrorString(cudaGetLastError()));
        st = VDB_CUDA_ERROR;
    }
    sc.release();
    train_buf.release();
    if (st == VDB_OK) ix->trained = true;
    return st;
}

// add() and add_assigned() share everything but the assignment step
static int32_t add_impl(vdb_index* ix, const float* vectors, const uint64_t* ids, const uint32_t* assigned,
                        uint64_t n, uint64_t counted) {
    ix->tables_dirty = false;  // every chunk below uploads the tables anyway
    const uint64_t chunk_rows = std::max<uint64_t>(1024, ADD_CHUNK_BYTES / (ix->ld * 4ull));
    const bool ids_dev = is_device_ptr(ids);
    for (uint64_t lo = 0; lo < n; lo += chunk_rows) {
        const uint64_t m = std::min(chunk_rows, n - lo);
        const float* x = nullptr;
        VDB_TRY(stage_rows(ix, vectors + lo * ix->dim, m, ix->stage_buf, &x, ix->stream));
        const uint64_t* dids = nullptr;
        if (ids) {
            if (ids_dev) {
                dids = ids + lo;
            } else {
                VDB_TRY(ix->ids_stage.reserve(m));
                VDB_CUDA_TRY(cudaMemcpyAsync(ix->ids_stage.p, ids + lo, m * 8, cudaMemcpyHostToDevice, ix->stream));
                dids = ix->ids_stage.p;
            }
        }
        // assign (ivf_flat_index.cpp:151-157), unless the caller already did
        const uint32_t* asg = assigned ? assigned + lo : nullptr;
        if (!asg) {
            VDB_TRY(ix->assign_buf.reserve(m));
            VDB_TRY(assign_rows(ix, x, m, ix->assign_buf.p, ix->stream));
            asg = ix->assign_buf.p;
        }
        // per-list counts -> grow the page chains
        VDB_TRY(ix->hist_buf.reserve(ix->nlist + 1));
        VDB_TRY(ix->fill_buf.reserve(ix->nlist));
        VDB_CUDA_TRY(cudaMemsetAsync(ix->hist_buf.p, 0, (ix->nlist + 1) * 4, ix->stream));
        VDB_CUDA_TRY(cudaMemsetAsync(ix->fill_buf.p, 0, ix->nlist * 4, ix->stream));
        VDB_TRY(launch_hist(asg, m, ix->nlist, ix->cfg.shard_rank, ix->d_owner.p, ix->hist_buf.p, ix->stream));
        std::vector<uint32_t> hist(ix->nlist + 1);
        VDB_CUDA_TRY(cudaMemcpyAsync(hist.data(), ix->hist_buf.p, (ix->nlist + 1) * 4, cudaMemcpyDeviceToHost, ix->stream));
        VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
        // caller-supplied assignments (add_assigned) naming no list: nothing of this chunk has been stored yet
        VDB_REQUIRE(hist[ix->nlist] == 0, "add: an assignment is >= nlist");
        std::vector<uint32_t> new_rows = ix->h_rows;
        uint64_t added = 0;
        for (uint32_t l = 0; l < ix->nlist; ++l) {
            if (!hist[l]) continue;
            VDB_REQUIRE((uint64_t)new_rows[l] + hist[l] < 0xffffffffull, "add: a list would exceed 2^32 rows");
            new_rows[l] += hist[l];
            added += hist[l];
            const uint32_t need = (new_rows[l] + ix->page_rows - 1) / ix->page_rows;
            while (ix->h_pages[l].size() < need) {
                uint32_t pg;
                VDB_TRY(index_alloc_page(ix, &pg));
                ix->h_pages[l].push_back(pg);
            }
        }
        // device tables: page chains now, old row counts as the append base
        VDB_TRY(index_upload_list_tables(ix));  // uploads h_rows (= old counts) and the grown chains
        VDB_TRY(launch_scatter_rows(x, ix->ld, dids, ix->total_vectors + lo, m, asg, ix->d_rows.p, ix->fill_buf.p,
                                    ix->d_page_off.p, ix->d_page_vec.p, ix->d_page_ids.p, ix->page_rows, ix->ld,
                                    ix->nlist, ix->cfg.shard_rank, ix->d_owner.p, ix->mirror_off, ix->mirror_kind, ix->stream));
        ix->h_rows = new_rows;
        VDB_CUDA_TRY(cudaMemcpyAsync(ix->d_rows.p, ix->h_rows.data(), ix->nlist * 4, cudaMemcpyHostToDevice,
                                     ix->stream));
        VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
        ix->local_vectors += added;
    }
    ix->total_vectors += counted;  // total_vectors_ += n_vectors (:200)
    return reserve_slots(ix);
}

int32_t vdb_index_add(vdb_index* ix, const float* vectors, const uint64_t* ids, uint64_t n) {
    if (ix && ix->composite) {
        if (n == 0) return VDB_OK;
        VDB_REQUIRE(vectors, "add: null vectors");
        return composite_add(ix, vectors, ids, n);
    }
    VDB_TRY(check_index(ix));
    if (n == 0) return VDB_OK;
    VDB_REQUIRE(vectors, "add: null vectors");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    return add_impl(ix, vectors, ids, nullptr, n, n);
}

int32_t vdb_index_add_assigned(vdb_index* ix, const float* vectors, const uint64_t* ids,
                               const uint32_t* assignments_dev, uint64_t n, uint64_t global_n) {
    VDB_REQUIRE(!(ix && ix->composite), "add_assigned: address the shards of a single-process sharded index through vdb_index_add");
    VDB_TRY(check_index(ix));
    if (n == 0 && global_n == 0) return VDB_OK;
    VDB_REQUIRE(n == 0 || (vectors && ids && assignments_dev), "add_assigned: null buffer");
    VDB_REQUIRE(n == 0 || is_device_ptr(assignments_dev), "add_assigned: assignments must be a device array");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    return add_impl(ix, vectors, ids, assignments_dev, n, global_n);
}

int32_t vdb_index_search_async(vdb_index* ix, const float* queries_dev, uint32_t nq, uint32_t nprobe, uint32_t k,
                               float* distances_dev, uint64_t* indices_dev, void* stream) {
    VDB_REQUIRE(!(ix && ix->composite), "search_async: a single-process sharded index runs on its own streams; use vdb_index_search_submit");
    VDB_TRY(check_index(ix));
    VDB_TRY(check_search_args(ix, queries_dev, distances_dev, indices_dev, nq, nprobe, k));
    VDB_REQUIRE(is_device_ptr(distances_dev) && is_device_ptr(indices_dev), "search_async: outputs must be device arrays");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    uint32_t ppi = 1, nq_chunk = nq;
    choose_ppi(ix, nq, std::min(nprobe, ix->nlist), k, &ppi, &nq_chunk);
    VDB_REQUIRE(nq_chunk == nq, "search_async: nq * nprobe * k too large for one pass; use vdb_index_search");
    SearchSlot* s = nullptr;
    uint64_t t = 0;
    VDB_TRY(index_acquire_slot(ix, &s, &t));
    cudaStream_t st = (cudaStream_t)stream;
    return index_enqueue_search(ix, *s, queries_dev, nq, nprobe, k, distances_dev, indices_dev,
                                SearchStreams{st, st, st, false, nullptr}, true);
}

int32_t vdb_index_search_submit(vdb_index* ix, const float* queries, uint32_t nq, uint32_t nprobe, uint32_t k,
                                float* distances, uint64_t* indices, uint64_t* ticket) {
    if (ix && ix->composite) {
        VDB_TRY(check_search_args(ix, queries, distances, indices, nq, nprobe, k));
        VDB_REQUIRE(ticket, "search_submit: null ticket");
        return composite_submit(ix, queries, nq, nprobe, k, distances, indices, ticket);
    }
    VDB_TRY(check_index(ix));
    VDB_TRY(check_search_args(ix, queries, distances, indices, nq, nprobe, k));
    VDB_REQUIRE(ticket, "search_submit: null ticket");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    uint32_t ppi = 1, nq_chunk = nq;
    choose_ppi(ix, nq, std::min(nprobe, ix->nlist), k, &ppi, &nq_chunk);
    nq_chunk = screen_chunk(ix, nq, k, nq_chunk);
    if (nq_chunk < nq) {  // too large for one pass: done chunk by chunk right here, the ticket is already complete
        VDB_TRY(search_chunked(ix, queries, nq, nprobe, k, distances, indices, nq_chunk));
        *ticket = ix->next_ticket;
        return VDB_OK;
    }
    SearchSlot* s = nullptr;
    VDB_TRY(index_acquire_slot(ix, &s, ticket));
    return index_enqueue_search(ix, *s, queries, nq, nprobe, k, distances, indices, index_pipeline_streams(ix, *ticket), true);
}

int32_t vdb_index_search_wait(vdb_index* ix, uint64_t ticket) {
    if (ix && ix->composite) return composite_wait(ix, ticket);
    VDB_TRY(check_index(ix));
    SearchSlot* s = nullptr;
    {
        std::lock_guard<std::mutex> lock(ix->mu);
        VDB_REQUIRE(ticket >= 1 && ticket <= ix->next_ticket, "search_wait: unknown ticket");
        s = &ix->slots[ticket % ix->depth];
        if (s->ticket != ticket || !s->busy) return VDB_OK;  // finished (and delivered) when its slot was recycled
        if (s->collect_deferred) VDB_TRY(flush_deferred_collect(ix, s->collect_stream));
    }
    {
        DeviceGuard g(ix->device);
        VDB_CUDA_TRY(cudaEventSynchronize(s->ev_done));  // without the lock: other threads keep submitting
    }
    std::lock_guard<std::mutex> lock(ix->mu);
    if (s->ticket != ticket) return VDB_OK;
    return index_finish_slot(ix, *s);
}

int32_t vdb_index_search_wait_stream(vdb_index* ix, uint64_t ticket, void* stream) {
    if (ix && ix->composite) return composite_wait_stream(ix, ticket, (cudaStream_t)stream);
    VDB_TRY(check_index(ix));
    std::lock_guard<std::mutex> lock(ix->mu);
    VDB_REQUIRE(ticket >= 1 && ticket <= ix->next_ticket, "search_wait_stream: unknown ticket");
    SearchSlot& s = ix->slots[ticket % ix->depth];
    if (s.ticket != ticket || !s.busy) return VDB_OK;
    DeviceGuard g(ix->device);
    if (s.collect_deferred) VDB_TRY(flush_deferred_collect(ix, s.collect_stream));
    VDB_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, s.ev_done, 0));
    return VDB_OK;
}

int32_t vdb_index_search(vdb_index* ix, const float* queries, uint32_t nq, uint32_t nprobe, uint32_t k,
                         float* distances, uint64_t* indices) {
    uint64_t ticket = 0;
    VDB_TRY(vdb_index_search_submit(ix, queries, nq, nprobe, k, distances, indices, &ticket));
    return vdb_index_search_wait(ix, ticket);
}

int32_t vdb_index_set_arena(vdb_index* ix, vdb_arena* arena, int32_t arena_device) {
    VDB_TRY(check_index(ix));
    if (ix->composite) {
        for (uint32_t r = 0; r < composite_size(ix); ++r) VDB_TRY(vdb_index_set_arena(composite_shard(ix, r), arena, arena_device));
        return VDB_OK;
    }
    std::lock_guard<std::mutex> lock(ix->mu);
    VDB_REQUIRE(ix->slabs.empty() || arena == ix->arena, "set_arena: the index already holds list pages");
    ix->arena = arena;
    ix->arena_device = arena_device;
    return VDB_OK;
}

int32_t vdb_index_reserve_search(vdb_index* ix, uint32_t max_nq, uint32_t max_nprobe, uint32_t max_k) {
    if (ix && ix->composite) return composite_reserve_search(ix, max_nq, max_nprobe, max_k);
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(max_nq >= 1 && max_nprobe >= 1 && max_k >= 1 && max_k <= (uint32_t)scan_max_k(), "reserve_search: bad shape");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    ix->rs_nq = max_nq; ix->rs_np = max_nprobe; ix->rs_k = max_k;
    return reserve_slots(ix);
}

int32_t vdb_index_attach_exchange(vdb_index* ix, vdb_exchange* ex) {
    VDB_REQUIRE(!(ix && ix->composite), "attach_exchange: a single-process sharded index owns its exchange");
    VDB_TRY(check_index(ix));
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    for (uint32_t i = 0; i < ix->depth; ++i) VDB_TRY(index_finish_slot(ix, ix->slots[i]));
    ix->exchange = ex;
    return VDB_OK;
}

int32_t vdb_index_select_nprobe(vdb_index* ix, const float* queries, uint32_t nq, uint32_t nprobe, uint32_t* lists) {
    if (ix && ix->composite) ix = composite_root(ix);
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(queries && lists && nq >= 1 && nprobe >= 1, "select_nprobe: bad arguments");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    const uint32_t np = std::min(nprobe, ix->nlist);
    SearchSlot* s = &ix->aux_slot;  // not a slot of the search ring: takes no ticket
    const float* q = nullptr;
    VDB_TRY(stage_rows(ix, queries, nq, s->q_buf, &q, ix->stream));
    VDB_TRY(slot_coarse_select(ix, *s, q, nq, np, ix->stream));
    VDB_CUDA_TRY(cudaMemcpyAsync(lists, s->probes.p, (size_t)nq * np * 4,
                                 is_device_ptr(lists) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                 ix->stream));
    VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return VDB_OK;
}

int32_t vdb_index_assign(vdb_index* ix, const float* vectors, uint64_t n, uint32_t* lists) {
    if (ix && ix->composite) ix = composite_root(ix);
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(vectors && lists, "assign: null buffer");
    if (n == 0) return VDB_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    const float* x = nullptr;
    VDB_TRY(stage_rows(ix, vectors, n, ix->stage_buf, &x, ix->stream));
    VDB_TRY(ix->assign_buf.reserve(n));
    VDB_TRY(assign_rows(ix, x, n, ix->assign_buf.p, ix->stream));
    VDB_CUDA_TRY(cudaMemcpyAsync(lists, ix->assign_buf.p, n * 4,
                                 is_device_ptr(lists) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                 ix->stream));
    VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return VDB_OK;
}

int32_t vdb_index_get_centroids(vdb_index* ix, float* out) {
    if (ix && ix->composite) ix = composite_root(ix);
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(out, "get_centroids: null buffer");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    VDB_CUDA_TRY(cudaMemcpy2DAsync(out, (size_t)ix->dim * 4, ix->centroids.p, (size_t)ix->ld * 4, (size_t)ix->dim * 4,
                                   ix->nlist, cudaMemcpyDefault, ix->stream));
    VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return VDB_OK;
}

int32_t vdb_index_set_centroids(vdb_index* ix, const float* in) {
    if (ix && ix->composite) {
        VDB_REQUIRE(in, "set_centroids: null buffer");
        return composite_set_centroids(ix, in);
    }
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(in, "set_centroids: null buffer");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    VDB_CUDA_TRY(cudaMemsetAsync(ix->centroids.p, 0, (size_t)ix->nlist * ix->ld * 4, ix->stream));
    VDB_CUDA_TRY(cudaMemcpy2DAsync(ix->centroids.p, (size_t)ix->ld * 4, in, (size_t)ix->dim * 4, (size_t)ix->dim * 4,
                                   ix->nlist, cudaMemcpyDefault, ix->stream));
    VDB_TRY(refresh_centroid_aux(ix));
    VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
    ix->trained = true;
    return VDB_OK;
}

int32_t vdb_index_get_owners(vdb_index* ix, uint8_t* out) {
    if (ix && ix->composite) ix = composite_root(ix);
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(out, "get_owners: null buffer");
    std::lock_guard<std::mutex> lock(ix->mu);
    for (uint32_t l = 0; l < ix->nlist; ++l) out[l] = ix->cfg.shard_count > 1 ? ix->h_owner[l] : 0;
    return VDB_OK;
}

int32_t vdb_index_set_owners(vdb_index* ix, const uint8_t* in) {
    if (ix && ix->composite) {
        VDB_REQUIRE(in, "set_owners: null buffer");
        return composite_set_owners(ix, in);
    }
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(in, "set_owners: null buffer");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    VDB_REQUIRE(ix->total_vectors == 0, "set_owners: list ownership can only change while the index is empty");
    if (ix->cfg.shard_count <= 1) return VDB_OK;
    for (uint32_t l = 0; l < ix->nlist; ++l) {
        VDB_REQUIRE(in[l] < ix->cfg.shard_count, "set_owners: rank out of range");
        ix->h_owner[l] = in[l];
    }
    return index_upload_owners(ix);
}

int32_t vdb_index_list_sizes(vdb_index* ix, uint64_t* out) {
    if (ix && ix->composite) {
        VDB_REQUIRE(out, "list_sizes: null buffer");
        return composite_list_sizes(ix, out);
    }
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(out, "list_sizes: null buffer");
    std::lock_guard<std::mutex> lock(ix->mu);
    for (uint32_t l = 0; l < ix->nlist; ++l) out[l] = ix->h_rows[l];
    return VDB_OK;
}

int32_t vdb_index_list_ids(vdb_index* ix, uint32_t list, uint64_t* out) {
    if (ix && ix->composite) {
        VDB_REQUIRE(out, "list_ids: bad arguments");
        return composite_list_ids(ix, list, out);
    }
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(out && list < ix->nlist, "list_ids: bad arguments");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    uint32_t left = ix->h_rows[list];
    for (size_t j = 0; j < ix->h_pages[list].size() && left; ++j) {
        const uint32_t take = std::min(left, ix->page_rows);
        const void* src = (const void*)(uintptr_t)(ix->page_addr[ix->h_pages[list][j]] + ix->ids_off);
        VDB_CUDA_TRY(cudaMemcpyAsync(out + (size_t)j * ix->page_rows, src, (size_t)take * 8, cudaMemcpyDefault,
                                     ix->stream));
        left -= take;
    }
    VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return VDB_OK;
}

int32_t vdb_index_list_vectors(vdb_index* ix, uint32_t list, float* out) {
    VDB_TRY(check_index(ix));
    if (ix->composite) {
        VDB_REQUIRE(out && list < ix->cfg.nlist, "list_vectors: bad arguments");
        std::vector<uint8_t> owners(ix->cfg.nlist);
        VDB_TRY(vdb_index_get_owners(composite_root(ix), owners.data()));
        return vdb_index_list_vectors(composite_shard(ix, owners[list]), list, out);
    }
    VDB_REQUIRE(out && list < ix->nlist, "list_vectors: bad arguments");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    uint32_t left = ix->h_rows[list];
    for (size_t j = 0; j < ix->h_pages[list].size() && left; ++j) {
        const uint32_t take = std::min(left, ix->page_rows);
        const void* src = (const void*)(uintptr_t)ix->page_addr[ix->h_pages[list][j]];
        VDB_CUDA_TRY(cudaMemcpy2DAsync(out + (size_t)j * ix->page_rows * ix->dim, (size_t)ix->dim * 4, src,
                                       (size_t)ix->ld * 4, (size_t)ix->dim * 4, take, cudaMemcpyDefault, ix->stream));
        left -= take;
    }
    VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return VDB_OK;
}

// Append rows whose list is known (persistence: one stored list file) -- no assignment, and no staging either: the
// copies go from the caller's memory (e.g. a memory-mapped Arrow values buffer) straight into the list's HBM pages.
int32_t vdb_index_append_list(vdb_index* ix, uint32_t list, const float* vectors, const uint64_t* ids, uint64_t n) {
    VDB_TRY(check_index(ix));
    if (n == 0) return VDB_OK;
    if (ix->composite) {
        VDB_REQUIRE(vectors && ids && list < ix->cfg.nlist, "append_list: bad arguments");
        std::vector<uint8_t> owners(ix->cfg.nlist);
        VDB_TRY(vdb_index_get_owners(composite_root(ix), owners.data()));
        // every shard counts the rows (get_total_vectors), the owner stores them
        for (uint32_t r = 0; r < composite_size(ix); ++r) {
            vdb_index* s = composite_shard(ix, r);
            if (r == owners[list]) VDB_TRY(vdb_index_append_list(s, list, vectors, ids, n));
            else {
                std::lock_guard<std::mutex> lock(s->mu);
                s->total_vectors += n;
            }
        }
        return composite_note_added(ix, n);
    }
    VDB_REQUIRE(vectors && ids && list < ix->nlist, "append_list: bad arguments");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    ix->total_vectors += n;
    if (ix->cfg.shard_count > 1 && ix->h_owner[list] != ix->cfg.shard_rank) return VDB_OK;  // another shard's list
    VDB_REQUIRE((uint64_t)ix->h_rows[list] + n < 0xffffffffull, "append_list: the list would exceed 2^32 rows");
    const uint32_t old = ix->h_rows[list];
    const uint32_t need = (uint32_t)((old + n + ix->page_rows - 1) / ix->page_rows);
    while (ix->h_pages[list].size() < need) {
        uint32_t pg;
        VDB_TRY(index_alloc_page(ix, &pg));
        ix->h_pages[list].push_back(pg);
    }
    uint64_t done = 0;
    while (done < n) {
        const uint32_t pos = old + (uint32_t)done, pg = pos / ix->page_rows, r = pos % ix->page_rows;
        const uint64_t take = std::min<uint64_t>(ix->page_rows - r, n - done);
        uint8_t* page = (uint8_t*)(uintptr_t)ix->page_addr[ix->h_pages[list][pg]];
        if (ix->dim != ix->ld)
            VDB_CUDA_TRY(cudaMemsetAsync(page + (size_t)r * ix->ld * 4, 0, take * ix->ld * 4, ix->stream));
        VDB_CUDA_TRY(cudaMemcpy2DAsync(page + (size_t)r * ix->ld * 4, (size_t)ix->ld * 4, vectors + done * ix->dim,
                                       (size_t)ix->dim * 4, (size_t)ix->dim * 4, take, cudaMemcpyDefault, ix->stream));
        VDB_CUDA_TRY(cudaMemcpyAsync(page + ix->ids_off + (size_t)r * 8, ids + done, take * 8, cudaMemcpyDefault,
                                     ix->stream));
        VDB_TRY(launch_page_norms(reinterpret_cast<const float*>(page), ix->ld,
                                  reinterpret_cast<float*>(page + ix->ids_off + (size_t)ix->page_rows * 8), r,
                                  (uint32_t)take, ix->mirror_off, ix->mirror_kind, ix->page_rows, ix->stream));
        done += take;
    }
    VDB_CUDA_TRY(cudaStreamSynchronize(ix->stream));
    ix->h_rows[list] = old + (uint32_t)n;
    ix->local_vectors += n;
    ix->tables_dirty = true;  // the device tables are refreshed once, by the next search / add / finish_load
    return VDB_OK;
}

int32_t vdb_index_finish_load(vdb_index* ix) {
    VDB_TRY(check_index(ix));
    if (ix->composite) {
        for (uint32_t r = 0; r < composite_size(ix); ++r) VDB_TRY(vdb_index_finish_load(composite_shard(ix, r)));
        return VDB_OK;
    }
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    VDB_TRY(flush_tables(ix));
    return reserve_slots(ix);
}

int32_t vdb_index_balance_owners(vdb_index* ix, const uint64_t* list_sizes) {
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(list_sizes, "balance_owners: null buffer");
    if (ix->composite) {
        for (uint32_t r = 0; r < composite_size(ix); ++r) VDB_TRY(vdb_index_balance_owners(composite_shard(ix, r), list_sizes));
        return VDB_OK;
    }
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    VDB_REQUIRE(ix->total_vectors == 0, "balance_owners: list ownership can only change while the index is empty");
    if (ix->cfg.shard_count <= 1) return VDB_OK;
    std::vector<uint32_t> counts(ix->nlist);
    for (uint32_t l = 0; l < ix->nlist; ++l) counts[l] = (uint32_t)std::min<uint64_t>(list_sizes[l], 0xffffffffull);
    balance_owners(ix, counts);
    return index_upload_owners(ix);
}

int32_t vdb_index_stats(vdb_index* ix, vdb_stats* out) {
    if (ix && ix->composite) {
        VDB_REQUIRE(out, "stats: null buffer");
        return composite_stats(ix, out);
    }
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(out, "stats: null buffer");
    std::lock_guard<std::mutex> lock(ix->mu);
    std::memset(out, 0, sizeof(*out));
    out->total_vectors = ix->total_vectors;
    out->local_vectors = ix->local_vectors;
    out->gpu_memory_bytes = ix->hbm_bytes();
    out->pages = ix->pages_used;
    out->dimension = ix->dim;
    out->nlist = ix->nlist;
    out->row_stride = ix->ld;
    out->page_rows = ix->page_rows;
    out->trained = ix->trained ? 1 : 0;
    out->metric = ix->cfg.metric;
    if (ix->d_scanned) {  // distinct list rows streamed by every search so far (counted by the grouping kernel)
        DeviceGuard g(ix->device);
        unsigned long long rows = 0;
        VDB_CUDA_TRY(cudaMemcpy(&rows, ix->d_scanned, 8, cudaMemcpyDeviceToHost));
        out->scanned_bytes = rows * (4ull * ix->dim + 8);
    }
    return VDB_OK;
}

int32_t vdb_index_last_search_stats(vdb_index* ix, vdb_search_stats* out) {
    if (ix && ix->composite) {
        VDB_REQUIRE(out, "search stats: null buffer");
        return composite_last_search_stats(ix, out);
    }
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(out, "search stats: null buffer");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    std::memset(out, 0, sizeof(*out));
    out->bytes_per_row = 4ull * ix->dim + 8;
    out->streamed_bytes_per_row = out->bytes_per_row;
    if (ix->last_slot < 0) return VDB_OK;
    SearchSlot& s = ix->slots[ix->last_slot];
    VDB_TRY(index_finish_slot(ix, s));
    unsigned long long st[3] = {0, 0, 0};
    uint32_t tot[2] = {0, 0};
    VDB_CUDA_TRY(cudaMemcpy(st, s.ws_scan.stats, 24, cudaMemcpyDeviceToHost));
    VDB_CUDA_TRY(cudaMemcpy(tot, s.ws_scan.totals, 8, cudaMemcpyDeviceToHost));
    out->algorithmic_rows = st[0];
    out->unique_rows = st[1];
    out->scan_items = tot[0];
    out->scan_ctas = s.info.grid;
    // the bf16 screen streams the shadow row, |v|^2 and |v - bf16(v)|; ids and fp32 rows only for admitted pairs
    if (s.info.mirror) {
        out->streamed_bytes_per_row = ix->mirror_kind == MIRROR_I8 ? 1ull * ix->ld + 12 : 2ull * ix->ld + 8;
        out->rescored_pairs = st[2];
    }
    return VDB_OK;
}

int32_t vdb_index_set_profiling(vdb_index* ix, int32_t enable) {
    if (ix && ix->composite) return composite_set_profiling(ix, enable);
    VDB_TRY(check_index(ix));
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    for (uint32_t i = 0; i < ix->depth; ++i) VDB_TRY(index_finish_slot(ix, ix->slots[i]));
    ix->profiling = enable != 0;
    for (double& v : ix->prof_ms) v = 0.0;
    ix->prof_searches = 0;
    ix->span_open = false;
    return VDB_OK;
}

int32_t vdb_index_read_profile(vdb_index* ix, float* out_ms, uint32_t* searches) {
    if (ix && ix->composite) {
        VDB_REQUIRE(out_ms && searches, "read_profile: null buffer");
        return composite_read_profile(ix, out_ms, searches);
    }
    VDB_TRY(check_index(ix));
    VDB_REQUIRE(out_ms && searches, "read_profile: null buffer");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    for (uint32_t i = 0; i < ix->depth; ++i) VDB_TRY(index_finish_slot(ix, ix->slots[i]));
    for (int i = 0; i < 5; ++i) out_ms[i] = (float)ix->prof_ms[i];
    out_ms[5] = 0.f;
    if (ix->span_open) {  // first scan start .. last scan end: the scan streams' busy time when batches overlap
        VDB_CUDA_TRY(cudaEventSynchronize(ix->span_end));
        VDB_CUDA_TRY(cudaEventElapsedTime(&out_ms[5], ix->span_start, ix->span_end));
    }
    out_ms[6] = out_ms[7] = 0.f;
    *searches = ix->prof_searches;
    for (double& v : ix->prof_ms) v = 0.0;
    ix->prof_searches = 0;
    ix->span_open = false;
    return VDB_OK;
}

int32_t vdb_index_warmup(vdb_index* ix, const uint32_t* lists, uint32_t n) {
    if (ix && ix->composite) ix = composite_root(ix);
    VDB_TRY(check_index(ix));
    for (uint32_t i = 0; i < n; ++i) VDB_REQUIRE(lists && lists[i] < ix->nlist, "warmup: list id out of range");
    return VDB_OK;  // every list is HBM-resident from add() on; nothing to load (ivf_flat_index.cpp:387-444)
}

int32_t vdb_bruteforce_search(const float* database, const float* queries, const uint64_t* ids, uint64_t n,
                              uint32_t nq, uint32_t dim, uint32_t k, float* distances, uint64_t* indices,
                              int32_t metric, void* stream) {
    VDB_REQUIRE(database && queries && distances && indices, "bruteforce: null buffer");
    VDB_REQUIRE(n >= 1 && n < 0xffffffffull && nq >= 1 && dim >= 1 && dim <= 2048 && k >= 1,
                "bruteforce: bad shape");
    VDB_REQUIRE(metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP, "bruteforce: metric must be L2 or InnerProduct");
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t ld = round_up(dim, 4);
    const bool db_dev = is_device_ptr(database), q_dev = is_device_ptr(queries), o_dev = is_device_ptr(distances);
    VDB_REQUIRE(o_dev == is_device_ptr(indices), "bruteforce: distances and indices must live on the same side");
    DevBuf<float> dbb, qb, od;
    DevBuf<uint64_t> oi, idb, pv, pi;
    DevBuf<uint32_t> rows, poff, zero;
    ScanWorkspace ws;
    BruteTcScratch btc;
    auto cleanup = [&] {
        btc.release();
        dbb.release(); qb.release(); od.release(); oi.release(); idb.release(); pv.release(); pi.release();
        rows.release(); poff.release(); zero.release(); ws.release();
    };
    auto run = [&]() -> int32_t {
        const float* db = database;
        if (!db_dev || dim != ld || ((uintptr_t)database & 15)) {  // bulk TMA needs 16-byte aligned rows
            VDB_TRY(dbb.reserve((size_t)n * ld));
            if (db_dev) {
                VDB_TRY(launch_pad_rows(database, dim, dim, dbb.p, ld, n, s));
            } else {
                VDB_CUDA_TRY(cudaMemsetAsync(dbb.p, 0, (size_t)n * ld * 4, s));
                VDB_CUDA_TRY(cudaMemcpy2DAsync(dbb.p, (size_t)ld * 4, database, (size_t)dim * 4, (size_t)dim * 4, n,
                                               cudaMemcpyHostToDevice, s));
            }
            db = dbb.p;
        }
        const float* q = queries;
        if (!q_dev || dim != ld || ((uintptr_t)queries & 15)) {
            VDB_TRY(qb.reserve((size_t)nq * ld));
            if (q_dev) {
                VDB_TRY(launch_pad_rows(queries, dim, dim, qb.p, ld, nq, s));
            } else {
                VDB_CUDA_TRY(cudaMemsetAsync(qb.p, 0, (size_t)nq * ld * 4, s));
                VDB_CUDA_TRY(cudaMemcpy2DAsync(qb.p, (size_t)ld * 4, queries, (size_t)dim * 4, (size_t)dim * 4, nq,
                                               cudaMemcpyHostToDevice, s));
            }
            q = qb.p;
        }
        const uint64_t* dids = ids;
        if (ids && !is_device_ptr(ids)) {
            VDB_TRY(idb.reserve(n));
            VDB_CUDA_TRY(cudaMemcpyAsync(idb.p, ids, n * 8, cudaMemcpyHostToDevice, s));
            dids = idb.p;
        }
        float* dd = distances;
        uint64_t* di = indices;
        if (!o_dev) {
            VDB_TRY(od.reserve((size_t)nq * k));
            VDB_TRY(oi.reserve((size_t)nq * k));
            dd = od.p;
            di = oi.p;
        }
        auto deliver = [&]() -> int32_t {
            if (!o_dev) {
                VDB_CUDA_TRY(cudaMemcpyAsync(distances, dd, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s));
                VDB_CUDA_TRY(cudaMemcpyAsync(indices, di, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s));
            }
            VDB_CUDA_TRY(cudaStreamSynchronize(s));  // temporaries are freed on return
            return VDB_OK;
        };
        if (bruteforce_tensor_supported(n, nq, ld, k)) {
            // large inputs: contraction on the tensor cores, exact fp32 re-scoring of the few survivors
            int overflowed = 0;
            VDB_TRY(bruteforce_tensor(db, n, ld, dids, q, nq, k, metric, dd, di, btc, &overflowed, s));
            if (!overflowed) return deliver();
        }
        const uint32_t page_rows = 256;
        uint32_t npages = 0;
        VDB_TRY(build_flat_view(db, n, ld, page_rows, rows, poff, pv, pi, &npages, s));
        ListTable lt;
        lt.rows = rows.p; lt.page_off = poff.p; lt.page_vec = pv.p; lt.page_ids = pi.p;
        lt.ids_flat = dids; lt.nlist = 1; lt.page_rows = page_rows; lt.ld = ld;
        VDB_TRY(zero.reserve(nq));
        VDB_CUDA_TRY(cudaMemsetAsync(zero.p, 0, (size_t)nq * 4, s));
        // one item per page range, every query tile scanned inside it: single-page items keep the rows of all
        // concurrently running items within L2 (148 x 768 KB) while the tiles re-read them; widen only when the
        // partial-result buffer (nq x ranges x k entries) would get out of hand
        uint32_t ppi = 1;
        while ((uint64_t)nq * ((npages + ppi - 1) / ppi) * k * 12 > (8ull << 30) && ppi < npages) ppi *= 2;
        // (one range per query is the floor: beyond it widening no longer helps -- the round-1 loop never ended)
        VDB_REQUIRE((uint64_t)nq * ((npages + ppi - 1) / ppi) * k * 12 <= (8ull << 30),
                    "bruteforce: nq * k partial results exceed 8 GiB: split the query batch");
        const uint64_t slots = (uint64_t)nq * ((npages + ppi - 1) / ppi);
        VDB_TRY(scan_search(lt, q, nq, zero.p, 1, k, metric, ppi, slots, ws, false, dd, di, nullptr, s));
        return deliver();
    };
    const int32_t st = run();
    cleanup();
    return st;
}

int32_t vdb_kmeans_assign(const float* vectors, const float* centroids, uint32_t* assignments, float* distances,
                          uint64_t n, uint32_t n_centroids, uint32_t dim, int32_t metric, void* stream) {
    VDB_REQUIRE(vectors && centroids && assignments, "kmeans_assign: null buffer");
    VDB_REQUIRE(metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP, "kmeans_assign: metric must be L2 or InnerProduct");
    cudaStream_t s = (cudaStream_t)stream;
    const bool vdev = is_device_ptr(vectors), cdev = is_device_ptr(centroids), adev = is_device_ptr(assignments);
    DevBuf<float> vb, cb, db;
    DevBuf<uint32_t> ab;
    auto run = [&]() -> int32_t {
        const float* v = vectors;
        const float* c = centroids;
        if (!vdev) {
            VDB_TRY(vb.reserve((size_t)n * dim));
            VDB_CUDA_TRY(cudaMemcpyAsync(vb.p, vectors, (size_t)n * dim * 4, cudaMemcpyHostToDevice, s));
            v = vb.p;
        }
        if (!cdev) {
            VDB_TRY(cb.reserve((size_t)n_centroids * dim));
            VDB_CUDA_TRY(cudaMemcpyAsync(cb.p, centroids, (size_t)n_centroids * dim * 4, cudaMemcpyHostToDevice, s));
            c = cb.p;
        }
        uint32_t* a = assignments;
        float* d = distances;
        if (!adev) {
            VDB_TRY(ab.reserve(n));
            a = ab.p;
            if (distances) {
                VDB_TRY(db.reserve(n));
                d = db.p;
            }
        }
        VDB_TRY(kmeans_assign_exact(v, n, dim, c, n_centroids, dim, dim, metric, a, d, s));
        if (!adev) {
            VDB_CUDA_TRY(cudaMemcpyAsync(assignments, a, n * 4, cudaMemcpyDeviceToHost, s));
            if (distances) VDB_CUDA_TRY(cudaMemcpyAsync(distances, d, n * 4, cudaMemcpyDeviceToHost, s));
        }
        if (!vdev || !cdev || !adev) VDB_CUDA_TRY(cudaStreamSynchronize(s));
        return VDB_OK;
    };
    const int32_t st = run();
    vb.release(); cb.release(); db.release(); ab.release();
    return st;
}

int32_t vdb_kmeans_accumulate(const float* vectors_dev, const uint32_t* assignments_dev, uint64_t n,
                              uint32_t n_centroids, uint32_t dim, float* sums_dev, uint32_t* counts_dev,
                              void* stream) {
    VDB_REQUIRE(vectors_dev && assignments_dev && sums_dev && counts_dev, "kmeans_accumulate: null buffer");
    VDB_REQUIRE(dim % 4 == 0, "kmeans_accumulate: dim must be a multiple of 4 (pad the rows)");
    VDB_REQUIRE(n < 0xffffffffull, "kmeans_accumulate: too many rows");
    cudaStream_t s = (cudaStream_t)stream;
    KMeansScratch sc;
    int32_t st = sc.reserve((uint32_t)n, n_centroids, dim);
    if (st == VDB_OK)
        st = kmeans_cluster_sums(vectors_dev, (uint32_t)n, dim, assignments_dev, n_centroids, dim, sums_dev,
                                 counts_dev, sc, s);
    if (st == VDB_OK && cudaStreamSynchronize(s) != cudaSuccess) st = VDB_CUDA_ERROR;
    sc.release();
    return st;
}

int32_t vdb_kmeans_finalize(const float* sums_dev, const uint32_t* counts_dev, float* centroids_dev,
                            uint32_t n_centroids, uint32_t dim, void* stream) {
    VDB_REQUIRE(sums_dev && counts_dev && centroids_dev, "kmeans_finalize: null buffer");
    return kmeans_divide(sums_dev, counts_dev, n_centroids, dim, centroids_dev, (cudaStream_t)stream);
}

int32_t vdb_merge_topk(const float* dist_parts_dev, const uint64_t* id_parts_dev, uint32_t parts, uint32_t nq,
                       uint32_t k, float* distances_dev, uint64_t* indices_dev, void* stream) {
    VDB_REQUIRE(dist_parts_dev && id_parts_dev && distances_dev && indices_dev, "merge_topk: null buffer");
    return merge_parts(dist_parts_dev, id_parts_dev, parts, nq, k, distances_dev, indices_dev, (cudaStream_t)stream);
}

}  // extern "C"
