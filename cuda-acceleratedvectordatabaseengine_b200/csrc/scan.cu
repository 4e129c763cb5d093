// The inverted-list scan engine for sm_100a.
//
// Replaces, for the whole query batch at once, what the reference does one
// (query, list) pair at a time:
//   select_nprobe_lists  ivf_flat_index.cpp:298-336   (run as a scan over the centroid table)
//   search_list_cpu/gpu  ivf_flat_index.cpp:339-384, 521-617 + bruteforce_search_kernel kernels.cuh:84-185
//   merge_results        ivf_flat_index.cpp:474-518
//
// Three kernels, no host synchronisation in between:
//   1. build_groups_kernel  groups the (query, probe) pairs by list, so that a
//      list probed by several queries of the batch is streamed from HBM once,
//      and cuts the work into items (list page range x tile of <= QT queries).
//   2. scan_kernel          persistent, one CTA per SM.  A producer thread
//      streams the item's rows (and their ids, and the next item's queries)
//      HBM -> shared memory with 1-D bulk TMA copies (cp.async.bulk + mbarrier
//      complete_tx) through an S-stage ring; eight consumer warps keep the
//      tile's queries in REGISTERS for the whole item (lane l owns float4
//      columns l, l+32, ...), read each staged row with 128-bit conflict-free
//      shared loads, accumulate exact fp32 (q-v)^2 or q.v for rows x queries,
//      reduce across lanes with a transposed butterfly (~1 shuffle per pair),
//      and push candidates that beat the running k-th distance into a
//      per-query shared pool that is bitonic-compacted to the best k (the
//      fused top-k).  A per-query bound on the final k-th distance is shared
//      between CTAs through global memory (atomicMin), so later items admit
//      almost nothing.
//   3. merge_kernel         per query: page partials -> per-list top-k
//      (multiset, as search_list_cpu returns it) -> sort by (dist,id), drop
//      duplicate ids, pad: merge_results.
#include <cuda_bf16.h>

#include "scan.cuh"
#include "exchange.cuh"
#include "topk.cuh"

namespace vdb {

namespace {

constexpr int STAGE_ROWS = 16;
constexpr int CONSUMER_WARPS = 8;
constexpr int CONSUMER_THREADS = CONSUMER_WARPS * 32;
constexpr int SCAN_THREADS = CONSUMER_THREADS + 128;  // 2 consumer warpgroups + the producer's warpgroup
constexpr int MAX_QT = 32;  // queries per CTA tile: register tile (tile_queries) x up to 8 warp groups
constexpr uint32_t MAX_K = 2048;
constexpr uint32_t SMEM_BUDGET = 227 * 1024;
constexpr uint32_t ITEM_CLASSES = 320;  // cost classes of the longest-first item order: 40 binades x 8

struct WorkList {
    uint32_t *gcount, *gfill, *goff, *ioff, *gpairs, *pair_slot;
    ScanItem* items;
    uint32_t* totals;
    unsigned long long* stats;
    uint32_t* qthr;  // [nq] running upper bound of each query's k-th distance, as an ordered key
    uint32_t nq;
    unsigned long long* lifetime_rows;  // optional: the index's running total of distinct rows streamed
};

// order-preserving float <-> uint32 so that atomicMin works on distances of either sign
constexpr uint32_t KEY_INF = 0xFF800000u;  // key of +inf
__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---------------------------------------------------------------- utilities

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    // 1024 threads; returns the exclusive prefix of v over the block
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();  // protect s_warp reuse
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t t = s_warp[lane];
        uint32_t u = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, u, o);
            if (lane >= o) u += y;
        }
        s_warp[lane] = u - t;  // exclusive prefix of the warp sums
        if (lane == 31) s_warp[32] = u;
    }
    __syncthreads();
    *total = s_warp[32];
    return s_warp[w] + x - v;
}

__device__ __forceinline__ uint32_t n_ranges(const ListTable& lt, uint32_t l, uint32_t ppi) {
    uint32_t npages = lt.page_off[l + 1] - lt.page_off[l];
    return (npages + ppi - 1) / ppi;
}

// ------------------------------------------------------- 1. probe grouping

// SMEM: the four per-list work arrays live in dynamic shared memory (nlist <= 8192) instead of global
// scratch, which removes most of the kernel's global round trips.
template <bool SMEM>
__global__ void __launch_bounds__(1024) build_groups_kernel(ListTable lt, const uint32_t* __restrict__ probes,
                                                            uint32_t npairs, uint32_t ppi, uint32_t ppi_max,
                                                            uint32_t scan_ctas, WorkList wl) {
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_cls[ITEM_CLASSES];
    __shared__ uint32_t s_cand[8];
    extern __shared__ uint32_t s_lists[];
    const uint32_t tid = threadIdx.x, NT = blockDim.x;
    const uint32_t nlist = lt.nlist;
    if (SMEM) {
        wl.gcount = s_lists;
        wl.gfill = s_lists + (nlist + 1);
        wl.goff = s_lists + 2 * (nlist + 1);
        wl.ioff = s_lists + 3 * (nlist + 1);
    }

    for (uint32_t l = tid; l < nlist; l += NT) {
        wl.gcount[l] = 0;
        wl.gfill[l] = 0;
    }
    if (tid == 0) {
        wl.stats[0] = 0;
        wl.stats[1] = 0;
        wl.stats[2] = 0;  // (row, query) pairs the screen kernel re-scored exactly
    }
    for (uint32_t q = tid; q < wl.nq; q += NT) wl.qthr[q] = KEY_INF;
    __syncthreads();
    for (uint32_t p = tid; p < npairs; p += NT) {
        uint32_t l = probes[p];
        if (l < nlist && lt.rows[l] > 0) atomicAdd(&wl.gcount[l], 1u);
    }
    if (tid < 8) s_cand[tid] = 0;
    __syncthreads();

    // Pages per item, chosen for THIS batch: every item costs a fixed set-up and epilogue (tile announcement, query
    // registers, final selection, the global-bound update), so items should be as long as the load balance allows.
    // The caller's `ppi` sizes the partial-result buffer (the smallest value in play); the kernel doubles it while the
    // batch still yields at least two items per scan CTA, up to ppi_max (16 pages = 12 MB of rows at 768-D: measured
    // best from a 1/8 shard, 0.579 -> 0.527 ms, to the whole index, 3.97 -> 3.89 ms; 32 starves a short scan and
    // gains nothing on a long one).
    if (ppi_max > ppi) {
        const uint32_t chunk = (nlist + NT - 1) / NT;
        const uint32_t lo = min(tid * chunk, nlist), hi = min(lo + chunk, nlist);
        uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t l = lo; l < hi; ++l)
            if (wl.gcount[l]) {
                const uint32_t npages = lt.page_off[l + 1] - lt.page_off[l];
#pragma unroll
                for (int c = 1; c < 8; ++c) {
                    const uint32_t pc = ppi << c;
                    if (pc <= ppi_max) cnt[c] += (npages + pc - 1) / pc;
                }
            }
#pragma unroll
        for (int c = 1; c < 8; ++c)
            if (cnt[c]) atomicAdd(&s_cand[c], cnt[c]);
        __syncthreads();
        uint32_t best = ppi;
#pragma unroll
        for (int c = 1; c < 8; ++c) {
            const uint32_t pc = ppi << c;
            if (pc <= ppi_max && s_cand[c] >= 2u * scan_ctas) best = pc;
        }
        ppi = best;  // the same value in every thread
    }

    // exclusive scans over lists: grouped-pair offsets and item offsets
    {
        const uint32_t chunk = (nlist + NT - 1) / NT;
        const uint32_t lo = min(tid * chunk, nlist), hi = min(lo + chunk, nlist);
        uint32_t g = 0, it = 0;
        unsigned long long alg = 0, uniq = 0;
        for (uint32_t l = lo; l < hi; ++l) {
            uint32_t c = wl.gcount[l];
            if (c) {
                g += c;
                it += n_ranges(lt, l, ppi);
                alg += (unsigned long long)lt.rows[l] * c;
                uniq += lt.rows[l];
            }
        }
        uint32_t tot_g, tot_it;
        uint32_t base_g = block_exclusive_scan(g, s_warp, &tot_g);
        uint32_t base_it = block_exclusive_scan(it, s_warp, &tot_it);
        for (uint32_t l = lo; l < hi; ++l) {
            wl.goff[l] = base_g;
            wl.ioff[l] = base_it;
            uint32_t c = wl.gcount[l];
            if (c) {
                base_g += c;
                base_it += n_ranges(lt, l, ppi);
            }
        }
        if (tid == 0) {
            wl.goff[nlist] = tot_g;
            wl.ioff[nlist] = tot_it;
            wl.totals[0] = tot_it;
            wl.totals[2] = 0;  // dynamic work counter of the scan kernel
        }
        if (alg) atomicAdd(&wl.stats[0], alg);
        if (uniq) atomicAdd(&wl.stats[1], uniq);
        if (uniq && wl.lifetime_rows) atomicAdd(wl.lifetime_rows, uniq);
    }
    // exclusive scan over pairs: first partial-result slot of each pair
    {
        const uint32_t chunk = (npairs + NT - 1) / NT;
        const uint32_t lo = min(tid * chunk, npairs), hi = min(lo + chunk, npairs);
        uint32_t s = 0;
        for (uint32_t p = lo; p < hi; ++p) {
            uint32_t l = probes[p];
            if (l < nlist && lt.rows[l] > 0) s += n_ranges(lt, l, ppi);
        }
        uint32_t tot;
        uint32_t base = block_exclusive_scan(s, s_warp, &tot);
        for (uint32_t p = lo; p < hi; ++p) {
            wl.pair_slot[p] = base;
            uint32_t l = probes[p];
            if (l < nlist && lt.rows[l] > 0) base += n_ranges(lt, l, ppi);
        }
        if (tid == 0) {
            wl.pair_slot[npairs] = tot;
            wl.totals[1] = tot;
        }
    }
    __syncthreads();
    for (uint32_t p = tid; p < npairs; p += NT) {
        uint32_t l = probes[p];
        if (l < nlist && lt.rows[l] > 0) {
            uint32_t pos = wl.goff[l] + atomicAdd(&wl.gfill[l], 1u);
            wl.gpairs[pos] = p;
        }
    }
    // one item per (list, page range); the scan loops over the list's query tiles inside the item, so the
    // second and later tiles re-read the same rows from L2 right after the first brought them in
    auto write_item = [&](uint32_t l, uint32_t r, uint32_t slot) {
        const uint32_t pg_first = lt.page_off[l], npages = lt.page_off[l + 1] - pg_first;
        uint4 lo, hi;
        lo.x = wl.goff[l];                      // gbase
        lo.y = wl.gcount[l];                    // gcount
        lo.z = r;                               // range
        lo.w = pg_first + r * ppi;              // pg0
        hi.x = min(ppi, npages - r * ppi);      // npg
        hi.y = r * ppi * lt.page_rows;          // row_base
        hi.z = lt.rows[l] - hi.y;               // rows_left
        hi.w = l;                               // list
        uint4* dst = reinterpret_cast<uint4*>(wl.items + slot);
        dst[0] = lo;
        dst[1] = hi;
    };
    // Items are handed to the scan in DESCENDING order of estimated cost (longest processing time first): the
    // persistent CTAs claim them dynamically, so the scan ends on the cheapest items and every SM finishes within
    // one small item of the others -- on a 1/8 shard (7 items per SM) the unordered tail cost ~10 % of the kernel.
    // Cost ~ rows x (8 + queries probing the list): HBM time per row plus the per-(row, query) arithmetic.  A
    // counting sort over logarithmic cost classes (3 mantissa bits) is all the precision the schedule needs.
    auto item_class = [&](uint32_t l, uint32_t r) -> uint32_t {
        const uint32_t npages = lt.page_off[l + 1] - lt.page_off[l];
        const uint32_t npg = min(ppi, npages - r * ppi);
        const uint32_t rows = min(npg * lt.page_rows, lt.rows[l] - r * ppi * lt.page_rows);
        const float cost = (float)rows * (float)(8u + wl.gcount[l]);
        int c = (int)(__float_as_uint(cost) >> 20) - (127 << 3);
        c = max(0, min(c, (int)ITEM_CLASSES - 1));
        return (ITEM_CLASSES - 1) - (uint32_t)c;  // class 0 = most expensive
    };
    // f(list, range) for every item, each item visited by exactly one thread (same thread in every pass)
    auto for_each_item = [&](auto&& f) {
        if (SMEM) {
            // thread per item: the owning list is found by binary search in the (shared-memory) item offsets
            const uint32_t total = wl.ioff[nlist];
            for (uint32_t idx = tid; idx < total; idx += NT) {
                uint32_t lo = 0, hi = nlist;  // last l with ioff[l] <= idx (empty lists share their successor's offset)
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1) >> 1;
                    if (wl.ioff[mid] <= idx) lo = mid; else hi = mid - 1;
                }
                f(lo, idx - wl.ioff[lo]);
            }
        } else {
            for (uint32_t l = tid; l < nlist; l += NT) {
                if (!wl.gcount[l]) continue;
                const uint32_t nr = n_ranges(lt, l, ppi);
                for (uint32_t r = 0; r < nr; ++r) f(l, r);
            }
        }
    };
    for (uint32_t c = tid; c < ITEM_CLASSES; c += NT) s_cls[c] = 0;
    __syncthreads();
    for_each_item([&](uint32_t l, uint32_t r) { atomicAdd(&s_cls[item_class(l, r)], 1u); });
    __syncthreads();
    if (tid < 32) {  // exclusive scan of the class counts (ITEM_CLASSES / 32 per lane)
        constexpr uint32_t PER = ITEM_CLASSES / 32;
        uint32_t sum = 0;
        for (uint32_t i = 0; i < PER; ++i) sum += s_cls[tid * PER + i];
        uint32_t x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (tid >= (uint32_t)o) x += y;
        }
        uint32_t run = x - sum;
        for (uint32_t i = 0; i < PER; ++i) {
            const uint32_t c = s_cls[tid * PER + i];
            s_cls[tid * PER + i] = run;
            run += c;
        }
    }
    __syncthreads();
    for_each_item([&](uint32_t l, uint32_t r) { write_item(l, r, atomicAdd(&s_cls[item_class(l, r)], 1u)); });
}

// ------------------------------------------------------------ 2. list scan

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait carries a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~1 ms
// passes) instead of re-issuing the probe, which keeps memory-bound waiting off the issue slots and the
// power budget
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(1000000u)
        : "memory");
}
// 1-D bulk TMA copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same, with an L2 eviction-priority hint (createpolicy) on the global read
__device__ __forceinline__ void tma_bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_THREADS) : "memory"); }
// consumer barrier that also ORs a per-thread flag across the 256 consumer threads
__device__ __forceinline__ bool consumer_bar_or(bool flag) {
    uint32_t r;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 p, %1, 0;\n"
        "bar.red.or.pred q, 1, %2, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(r)
        : "r"((uint32_t)flag), "n"(CONSUMER_THREADS)
        : "memory");
    return r != 0;
}

struct ScanParams {
    ListTable lt;
    const float* queries;  // [nq][ld]
    const ScanItem* items;
    const uint32_t* totals;
    const uint32_t* gpairs;
    const uint32_t* pair_slot;
    float* part_d;
    uint64_t* part_i;
    uint32_t* part_cnt;  // [nslots] valid entries of each partial
    uint32_t* qthr;      // [nq] ordered keys, see f2key
    float* gtop_d;       // [nq][k] each query's running best k over everything scanned so far (bound tightening)
    uint64_t* gtop_i;    // [nq][k]
    uint32_t* glock;     // [nq] spin locks guarding gtop
    uint32_t* work_counter;
    uint32_t k, P, S, np, check_interval, has_ids;
    uint32_t dot_min_rows;  // the screen is used when the launch streams at least this many distinct rows per CTA
    const unsigned long long* stats;  // [1] = distinct probed rows of this launch (build_groups_kernel)
    uint32_t dotform;  // L2 over pages that carry row norms: dot-product screen, exact (q-v)^2 for the survivors
    uint32_t stage_rows;  // rows per ring stage: 16, or 8 for rows wider than 4 KB
    uint32_t qt;  // queries per tile at run time (<= the kernel's register tile)
    int metric;
};

struct ScanSmem {
    float* stages;        // [S][STAGE_ROWS][ld]
    uint64_t* stage_ids;  // [S][STAGE_ROWS] ids of the staged rows (lists that carry ids)
    float* stage_norm;    // [S][STAGE_ROWS] |v|^2 of the staged rows (dotform)
    float* qn;            // [MAX_QT] |q|^2 of the tile's queries (dotform)
    float* sq;            // [2][QT][ld] the tile's queries, double-buffered across items
    uint64_t* pool_i;     // [QT][P]
    float* pool_d;        // [QT][P]
    uint32_t* cnt;
    float* thr;
    uint32_t* spair;
    uint32_t* sqidx;
    uint64_t* full;    // [S]
    uint64_t* empty;   // [S]
    uint64_t* qfull;   // [2]
    uint64_t* qempty;  // [2]
    uint32_t* tile;    // [2][8] {item | END, first pair, queries, range, pages, row_base, rows_left, first page}: producer -> consumers
    uint32_t* tq;      // [2][MAX_QT] query index of each tile slot            (written by the producer with the tile)
    uint32_t* tslot;   // [2][MAX_QT] partial-result slot of each (pair, range)
};

__device__ __forceinline__ ScanSmem carve(uint8_t* base, const ScanParams& p) {
    const uint32_t QT = p.qt;
    ScanSmem s;
    uint8_t* q = base;
    s.stages = (float*)q;
    q += (size_t)p.S * p.stage_rows * p.lt.ld * 4;
    s.stage_ids = (uint64_t*)q;
    q += (size_t)p.S * STAGE_ROWS * 8;
    s.stage_norm = (float*)q;
    q += (size_t)p.S * STAGE_ROWS * 4;
    s.qn = (float*)q;
    q += MAX_QT * 4;
    s.sq = (float*)q;
    q += (size_t)2 * QT * p.lt.ld * 4;
    s.pool_i = (uint64_t*)q;
    q += (size_t)QT * p.P * 8;
    s.pool_d = (float*)q;
    q += (size_t)QT * p.P * 4;
    s.cnt = (uint32_t*)q;
    q += MAX_QT * 4;
    s.thr = (float*)q;
    q += MAX_QT * 4;
    s.spair = (uint32_t*)q;
    q += MAX_QT * 4;
    s.sqidx = (uint32_t*)q;
    q += MAX_QT * 4;
    s.full = (uint64_t*)q;
    q += 8 * 8;
    s.empty = (uint64_t*)q;
    q += 8 * 8;
    s.qfull = (uint64_t*)q;
    q += 2 * 8;
    s.qempty = (uint64_t*)q;
    q += 2 * 8;
    s.tile = (uint32_t*)q;
    q += 2 * 8 * 4;
    s.tq = (uint32_t*)q;
    q += 2 * MAX_QT * 4;
    s.tslot = (uint32_t*)q;
    return s;
}

static uint32_t scan_smem_bytes(uint32_t ld, uint32_t S, uint32_t QT, uint32_t P, uint32_t stage_rows) {
    return S * stage_rows * ld * 4 + S * STAGE_ROWS * 12 + MAX_QT * 4 + 2 * QT * ld * 4 + QT * P * 12 + 4 * MAX_QT * 4 + 20 * 8 + 64 + 4 * MAX_QT * 4;
}

// queries held in registers per tile, by the number of float4 columns a lane owns
__host__ __device__ constexpr int tile_queries(int NJ) { return NJ <= 2 ? 8 : NJ <= 6 ? 4 : NJ <= 12 ? 2 : 1; }

// One warp sorts query j's pool and keeps the best k (multiset: duplicates of
// an id inside one list survive, exactly like search_list_cpu's partial_sort).
// When the best k carry k distinct ids, their k-th distance bounds the query's
// final k-th distance from above whatever the other lists hold, so it is
// published to the per-query global threshold that every later item starts from
// (with duplicate ids among them the bound would be unsound: merge_results
// drops duplicates, ivf_flat_index.cpp:496-504).
__device__ __forceinline__ void compact_pool(const ScanSmem& s, const ScanParams& p, uint32_t j, uint32_t lane) {
    float* d = s.pool_d + (size_t)j * p.P;
    uint64_t* id = s.pool_i + (size_t)j * p.P;
    const uint32_t c = min(s.cnt[j], p.P);
    if (c == 0) return;
    const uint32_t n2 = dev_next_pow2(c);
    for (uint32_t i = c + lane; i < n2; i += 32) {
        d[i] = FLT_MAX;
        id[i] = ID_PAD;
    }
    __syncwarp();
    bitonic_sort_pairs(d, id, n2, lane, 32, [] { __syncwarp(); });
    const uint32_t nc = min(c, p.k);
    bool publish = false;
    if (nc >= p.k) {
        const uint32_t k = p.k;
        bool dup = false;
        if (k <= 64) {
            for (uint32_t i = lane; i < k; i += 32) {
                const uint64_t me = id[i];
                for (uint32_t t = 0; t < i; ++t) dup |= (id[t] == me);
            }
        } else {
            // sort a copy of the k ids in the free upper half of the pool, then compare neighbours
            uint64_t* sc = id + (p.P >> 1);
            const uint32_t m2 = dev_next_pow2(k);
            for (uint32_t i = lane; i < m2; i += 32) sc[i] = i < k ? id[i] : ID_PAD;
            __syncwarp();
            for (uint32_t size = 2; size <= m2; size <<= 1)
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    for (uint32_t t = lane; t < (m2 >> 1); t += 32) {
                        const uint32_t a = ((t / stride) * (stride << 1)) + (t % stride), b = a + stride;
                        const uint64_t x = sc[a], y = sc[b];
                        if (((a & size) == 0) ? (y < x) : (x < y)) {
                            sc[a] = y;
                            sc[b] = x;
                        }
                    }
                    __syncwarp();
                }
            for (uint32_t i = lane; i + 1 < k; i += 32) dup |= (sc[i] == sc[i + 1]);
        }
        publish = !__any_sync(0xffffffffu, dup);
    }
    if (lane == 0) {
        s.cnt[j] = nc;
        if (nc >= p.k) {
            const float kth = d[p.k - 1];
            s.thr[j] = fminf(s.thr[j], kth);
            if (publish) atomicMin(&p.qthr[s.sqidx[j]], f2key(kth));
        }
    }
    __syncwarp();
}

// One warp folds query j's surviving local best (sorted, nc entries at the front of its pool) into the query's
// global running top-k, under a per-query spin lock, and publishes the new k-th distance as the bound.  An
// item-local k-th distance only bounds the answer by "the k-th best of these ~1000 rows"; the running top-k
// over every row scanned so far by any CTA tightens it to what the final answer will be, so that later items
// admit, keep and hand to the merge almost nothing.  (The partial results stay the source of truth for the
// merge -- this list only feeds the bound, which is published only while its k ids are distinct.)
// d / id: P entries of shared memory whose first nc (<= k, sorted) are the contribution; the rest is scratch.
__device__ __forceinline__ void contribute_global_at(float* d, uint64_t* id, uint32_t q, const ScanParams& p,
                                                     uint32_t nc, uint32_t lane) {
    const uint32_t k = p.k;
    if (!(d[0] <= key2f(__ldcg(&p.qthr[q])))) return;  // cannot improve the running top-k (warp-uniform)
    // try-lock: when another CTA is updating this query's list right now, skip -- the bound merely tightens a
    // little later, and no warp ever waits on another CTA
    uint32_t got = 0;
    if (lane == 0) {
        got = atomicCAS(&p.glock[q], 0u, 1u) == 0u;
        __threadfence();
    }
    if (!__shfl_sync(0xffffffffu, got, 0)) return;
    for (uint32_t i = lane; i < k; i += 32) {
        d[nc + i] = __ldcg(&p.gtop_d[(size_t)q * k + i]);
        id[nc + i] = __ldcg(&p.gtop_i[(size_t)q * k + i]);
    }
    const uint32_t m = nc + k, n2 = dev_next_pow2(m);  // m <= 2k <= P
    for (uint32_t i = m + lane; i < n2; i += 32) {
        d[i] = FLT_MAX;
        id[i] = ID_PAD;
    }
    __syncwarp();
    bitonic_sort_pairs(d, id, n2, lane, 32, [] { __syncwarp(); });
    for (uint32_t i = lane; i < k; i += 32) {
        p.gtop_d[(size_t)q * k + i] = d[i];
        p.gtop_i[(size_t)q * k + i] = id[i];
    }
    if (id[k - 1] != ID_PAD) {  // k real entries: their k-th distance is a bound if the ids are distinct
        bool dup = false;
        if (k <= 64) {
            for (uint32_t i = lane; i < k; i += 32) {
                const uint64_t me = id[i];
                for (uint32_t t = 0; t < i; ++t) dup |= (id[t] == me);
            }
        } else {
            uint64_t* sc = id + (p.P >> 1);  // the merged tail beyond k is no longer needed
            const uint32_t m2 = dev_next_pow2(k);
            const float kth_keep = d[k - 1];
            for (uint32_t i = lane; i < m2; i += 32) sc[i] = i < k ? id[i] : ID_PAD;
            __syncwarp();
            for (uint32_t size = 2; size <= m2; size <<= 1)
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    for (uint32_t t = lane; t < (m2 >> 1); t += 32) {
                        const uint32_t a = ((t / stride) * (stride << 1)) + (t % stride), b = a + stride;
                        const uint64_t x = sc[a], y = sc[b];
                        if (((a & size) == 0) ? (y < x) : (x < y)) {
                            sc[a] = y;
                            sc[b] = x;
                        }
                    }
                    __syncwarp();
                }
            for (uint32_t i = lane; i + 1 < k; i += 32) dup |= (sc[i] == sc[i + 1]);
            d[k - 1] = kth_keep;  // (P/2 >= k, so d[k-1] was not touched; kept explicit for clarity)
        }
        if (!__any_sync(0xffffffffu, dup) && lane == 0) atomicMin(&p.qthr[q], f2key(d[k - 1]));
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(&p.glock[q], 0u);
}
__device__ __forceinline__ void contribute_global(const ScanSmem& s, const ScanParams& p, uint32_t j, uint32_t nc,
                                                  uint32_t lane) {
    contribute_global_at(s.pool_d + (size_t)j * p.P, s.pool_i + (size_t)j * p.P, s.sqidx[j], p, nc, lane);
}

// Sum V per-lane partials across the warp: log2(V) "halving" exchanges leave
// each lane with one value (index = its top log2(V) lane bits), the remaining
// butterfly steps complete it.  V + log2(32/V) - 1 shuffles instead of 5 V.
template <int V>
__device__ __forceinline__ float transposed_reduce(float (&x)[V], uint32_t lane) {
#pragma unroll
    for (int half = V / 2, step = 16; half >= 1; half >>= 1, step >>= 1) {
        const bool up = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float send = up ? x[i] : x[i + half];
            const float keep = up ? x[i + half] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    float r = x[0];
#pragma unroll
    for (int step = 16 / V; step >= 1; step >>= 1) r += __shfl_xor_sync(0xffffffffu, r, step);
    return r;
}

// Relative slack of the dot-form screen.  With u = 2^-24 and chains of at most 2*NJ+6 <= 38 roundings (dot, exact
// distance) or 4*NJ+5 <= 69 (norms): |screen - exact| <= (69 + 38 + 4 + 80) u (|q|^2 + |v|^2) < 196 u (|q|^2+|v|^2),
// so lowering the screen by 2e-5 (= 335 u) of |q|^2 + |v|^2 can never reject a pair whose exact distance passes.
constexpr float DOT_SLACK = 2e-5f;

// this lane's share of the exact (q - v)^2 of one staged row: the arithmetic of the L2 branch of score_batch
// (rows come from the page in global memory -- just streamed, so normally still in L2 -- because the stage slot
// has been handed back to the producer by the time a pair is known to pass)
template <int NJ, bool FULLW>
__device__ __forceinline__ float exact_l2_lane(const float4* __restrict__ g4, uint32_t ld4, uint32_t r, uint32_t nr,
                                               uint32_t lane, const float4 (&q)[NJ]) {
    const float2 neg1 = make_float2(-1.f, -1.f);
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const uint32_t c4 = lane + 32 * jj;
        float4 v;
        if (FULLW) v = __ldg(&g4[min(r, nr - 1) * ld4 + c4]);
        else v = (r < nr && c4 < ld4) ? __ldg(&g4[r * ld4 + c4]) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float2 dlo = __ffma2_rn(make_float2(v.x, v.y), neg1, make_float2(q[jj].x, q[jj].y));
        const float2 dhi = __ffma2_rn(make_float2(v.z, v.w), neg1, make_float2(q[jj].z, q[jj].w));
        acc = jj == 0 ? __fmul2_rn(dlo, dlo) : __ffma2_rn(dlo, dlo, acc);
        acc = __ffma2_rn(dhi, dhi, acc);
    }
    return acc.x + acc.y;
}

// Distances of RB staged rows (r_first, r_first + r_stride, ...) to the first QC queries of this warp's
// group, then ONE transposed reduction for the RB x QC per-lane partials and the pushes of the survivors.
// Packed fp32x2 FMAs (sm_100): a (row, query) pair costs 2 instructions per float2 of the row for the exact L2
// form -- d = v * (-1) + q (exactly q - v), acc += d * d -- and 1 for a dot product.
// L2 over index pages (p.dotform) therefore SCREENS with the dot product: |q|^2 + |v|^2 - 2 q.v, lowered by a
// proven bound on its rounding error (DOT_SLACK), is compared with the query's running k-th distance; only a pair
// that passes -- after the first items of a query almost none does -- gets the exact (q - v)^2, computed by the whole
// warp from the row's copy in the page (L2), with the same arithmetic and summation tree as the unscreened
// path: the distances that reach the pools are bit-identical, at half the FP instructions and less power.  While
// some query of the tile has no bound yet (every pair would pass) the exact form runs directly.
// FULLW: the row is exactly 32 * NJ float4 wide (768-D, 128-D, ...), so no column needs masking and rows past
// the end of a short stage are simply read from the last valid row (their results are discarded)
template <int NJ, int QT, int QC, int RB, bool FULLW>
__device__ __forceinline__ void score_batch(const ScanParams& p, const ScanSmem& s, const float4* __restrict__ st4,
                                            uint32_t ld4, uint32_t cur, uint32_t r_first, uint32_t r_stride,
                                            uint32_t nr, uint32_t row0, const float4 (&qv)[QT][NJ], uint32_t myq,
                                            uint32_t g, uint32_t ng, uint32_t lane, uint32_t limit, bool& over,
                                            bool release, uint32_t pg, uint32_t prow, bool screen) {
    constexpr int QCP = QC <= 1 ? 1 : QC <= 2 ? 2 : QC <= 4 ? 4 : 8;  // reduction width: QC padded to a power of two
    constexpr int V = RB * QCP;
    // lane's value after the reduction: index = top log2(V) lane bits = (row in batch, query in group)
    constexpr int COPIES = 32 / V;
    const uint32_t vidx = lane / COPIES;
    const uint32_t b = vidx / QCP, jl = vidx % QCP;
    const uint32_t r = r_first + r_stride * b;
    const uint32_t j = g + ng * jl;  // the query's slot in the CTA tile
    const bool mine = (lane % COPIES) == 0 && r < nr && jl < myq;
    // `screen` (warp-uniform, decided by the caller once per check interval): every query of the tile has a finite
    // bound, so the dot-form screen rejects nearly everything; before that the exact form runs directly
    const bool exact_l2 = p.metric == VDB_METRIC_L2 && !screen;
    // this lane's row norm, read BEFORE the stage slot is handed back (the producer refills ids and norms with it)
    const float vn = (screen && r < nr) ? s.stage_norm[cur * STAGE_ROWS + r] : 0.f;
    float part[V];
    uint64_t rid[RB];
    const float2 neg1 = make_float2(-1.f, -1.f);
#pragma unroll
    for (int bb = 0; bb < RB; ++bb) {
        const uint32_t rr = r_first + r_stride * bb;
        float4 v[NJ];
        if (FULLW) {
            const float4* row4 = st4 + min(rr, nr - 1) * ld4 + lane;
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) v[jj] = row4[32 * jj];
        } else {
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const uint32_t c4 = lane + 32 * jj;
                v[jj] = (rr < nr && c4 < ld4) ? st4[rr * ld4 + c4] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        const uint32_t lr = row0 + rr;  // list-relative row
        rid[bb] = lr;
        if (rr < nr) {
            if (p.has_ids) rid[bb] = s.stage_ids[cur * STAGE_ROWS + rr];
            else if (p.lt.ids_flat) rid[bb] = __ldg(&p.lt.ids_flat[lr]);
        }
        if (release && bb == RB - 1) {
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.empty[cur]);  // slot free: this warp's last row now lives in registers
        }
        float2 acc2[QC];
        if (exact_l2) {
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const float2 vlo = make_float2(v[jj].x, v[jj].y), vhi = make_float2(v[jj].z, v[jj].w);
#pragma unroll
                for (int jq = 0; jq < QC; ++jq) {
                    const float2 dlo = __ffma2_rn(vlo, neg1, make_float2(qv[jq][jj].x, qv[jq][jj].y));
                    const float2 dhi = __ffma2_rn(vhi, neg1, make_float2(qv[jq][jj].z, qv[jq][jj].w));
                    acc2[jq] = jj == 0 ? __fmul2_rn(dlo, dlo) : __ffma2_rn(dlo, dlo, acc2[jq]);
                    acc2[jq] = __ffma2_rn(dhi, dhi, acc2[jq]);
                }
            }
        } else {  // q.v: the inner-product metric, and the screen of the dot-form L2
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const float2 vlo = make_float2(v[jj].x, v[jj].y), vhi = make_float2(v[jj].z, v[jj].w);
#pragma unroll
                for (int jq = 0; jq < QC; ++jq) {
                    const float2 qlo = make_float2(qv[jq][jj].x, qv[jq][jj].y);
                    acc2[jq] = jj == 0 ? __fmul2_rn(qlo, vlo) : __ffma2_rn(qlo, vlo, acc2[jq]);
                    acc2[jq] = __ffma2_rn(make_float2(qv[jq][jj].z, qv[jq][jj].w), vhi, acc2[jq]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < QCP; ++t) part[bb * QCP + t] = t < QC ? acc2[t].x + acc2[t].y : 0.f;
    }
    float tot = transposed_reduce<V>(part, lane);
    const float thr = mine ? s.thr[j] : -INFINITY;
    uint64_t id = rid[0];
#pragma unroll
    for (int t = 1; t < RB; ++t)
        if (b == (uint32_t)t) id = rid[t];
    auto push = [&](float d) {
        const uint32_t pos = atomicAdd(&s.cnt[j], 1u);
        over |= (pos >= limit);
        if (pos < p.P) {
            s.pool_d[(size_t)j * p.P + pos] = d;
            s.pool_i[(size_t)j * p.P + pos] = id;
        }
    };
    if (!screen) {
        if (p.metric != VDB_METRIC_L2) tot = -tot;  // IP distance = -dot, kernels.cuh:59
        if (mine && tot <= thr) push(tot);
        return;
    }
    // screen: a lower bound of the exact distance from the dot product and the two norms
    bool pass = false;
    if (mine) {
        const float sn = s.qn[j] + vn;
        pass = fmaf(-2.f, tot, sn * (1.f - DOT_SLACK)) <= thr;
    }
    uint32_t m = __ballot_sync(0xffffffffu, pass);
    if (m == 0) return;
    // rows of this stage in the page: stage row r = page row prow + r
    const float4* g4 = reinterpret_cast<const float4*>(__ldg(&p.lt.page_vec[pg])) + (size_t)prow * ld4;
    while (m) {  // warp-uniform: one admitted (row, query) pair at a time, exact distance by all 32 lanes
        const uint32_t L = (uint32_t)__ffs((int)m) - 1u;
        m &= m - 1u;
        const uint32_t lv = L / COPIES, vb = lv / QCP, vj = lv % QCP;
        const uint32_t rr = r_first + r_stride * vb;
        float e = 0.f;
#pragma unroll
        for (int t = 0; t < QC; ++t)
            if (vj == (uint32_t)t) e = exact_l2_lane<NJ, FULLW>(g4, ld4, rr, nr, lane, qv[t]);
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) e += __shfl_xor_sync(0xffffffffu, e, step);  // same tree as above
        if (lane == L && e <= thr) push(e);
    }
}

// QC = register-tile width actually needed (next power of two of the group's query count)
template <int NJ, int QT, int RB, bool FULLW>
__device__ __forceinline__ void score_dispatch(const ScanParams& p, const ScanSmem& s, const float4* st4, uint32_t ld4,
                                               uint32_t cur, uint32_t r_first, uint32_t r_stride, uint32_t nr,
                                               uint32_t row0, const float4 (&qv)[QT][NJ], uint32_t myq, uint32_t g,
                                               uint32_t ng, uint32_t lane, uint32_t limit, bool& over, bool release,
                                               uint32_t pg, uint32_t prow, bool screen) {
    if (QT >= 8 && myq > 4)
        score_batch<NJ, QT, (QT >= 8 ? 8 : QT), (RB > 4 ? 4 : RB), FULLW>(p, s, st4, ld4, cur, r_first, r_stride, nr, row0, qv, myq, g, ng, lane, limit, over, release, pg, prow, screen);
    else if (QT >= 4 && myq > 3)
        score_batch<NJ, QT, (QT >= 4 ? 4 : QT), RB, FULLW>(p, s, st4, ld4, cur, r_first, r_stride, nr, row0, qv, myq, g, ng, lane, limit, over, release, pg, prow, screen);
    else if (QT >= 4 && myq > 2)  // 3 queries: a quarter less math than the padded 4-wide tile
        score_batch<NJ, QT, (QT >= 4 ? 3 : QT), RB, FULLW>(p, s, st4, ld4, cur, r_first, r_stride, nr, row0, qv, myq, g, ng, lane, limit, over, release, pg, prow, screen);
    else if (QT >= 2 && myq > 1)
        score_batch<NJ, QT, (QT >= 2 ? 2 : QT), RB, FULLW>(p, s, st4, ld4, cur, r_first, r_stride, nr, row0, qv, myq, g, ng, lane, limit, over, release, pg, prow, screen);
    else
        score_batch<NJ, QT, 1, RB, FULLW>(p, s, st4, ld4, cur, r_first, r_stride, nr, row0, qv, myq, g, ng, lane, limit, over, release, pg, prow, screen);
}

template <int NJ>
__device__ __forceinline__ void consumer_loop(const ScanParams& p, const ScanSmem& s, const bool dotform) {
    constexpr int QT = tile_queries(NJ);
    const uint32_t ctid = threadIdx.x, lane = ctid & 31, warp = ctid >> 5;
    const uint32_t ld = p.lt.ld, ld4 = ld >> 2;
    const uint32_t limit = p.P - p.check_interval * STAGE_ROWS;
    const bool fullw = (ld4 == 32u * NJ);
    uint32_t stage = 0, phase = 0, qbuf = 0, qphase = 0;
    constexpr uint32_t END = 0xffffffffu;

    for (;;) {
        // next tile: announced by the producer together with its queries (staged while the previous tile was scanned)
        mbar_wait(&s.qfull[qbuf], qphase);
        const uint32_t* tw = s.tile + qbuf * 8;
        if (tw[0] == END) break;
        const uint32_t qcount = tw[2];
        struct { uint32_t range, npg, row_base, rows_left; } it = {tw[3], tw[4], tw[5], tw[6]};
        const uint32_t pg0 = tw[7];  // the item's first page in the page tables (exact re-scoring reads rows from there)
        // per-query tile state; the bound is the only global read of the set-up and is issued first so that it
        // overlaps the query loads below
        uint32_t my_q = 0, my_slot = 0;
        float my_thr = INFINITY;
        if (ctid < qcount) {
            my_q = s.tq[qbuf * MAX_QT + ctid];
            my_slot = s.tslot[qbuf * MAX_QT + ctid];
            my_thr = key2f(__ldcg(&p.qthr[my_q]));  // start from what earlier items already proved
        }
        // a tile wider than the register tile is split over warp groups: group g keeps queries g, g+ng, ...
        // and every group reads every staged row, so the rows are still brought in from HBM once
        uint32_t ng = 1;
        while (ng * QT < qcount) ng <<= 1;
        const uint32_t g = warp & (ng - 1), wg = warp / ng, nwg = CONSUMER_WARPS / ng;
        const uint32_t myq = qcount > g ? (qcount - g + ng - 1) / ng : 0;
        const uint32_t rows_per_warp = STAGE_ROWS / nwg;
        float4 qv[QT][NJ];
        {
            const float4* q4 = reinterpret_cast<const float4*>(s.sq) + (size_t)qbuf * p.qt * ld4;
#pragma unroll
            for (int j = 0; j < QT; ++j)
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    const uint32_t c4 = lane + 32 * jj;
                    qv[j][jj] = ((uint32_t)j < myq && c4 < ld4) ? q4[(g + ng * j) * ld4 + c4]
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                }
        }
        if (dotform) {  // |q|^2 of this group's queries (every warp of the group writes the same values)
#pragma unroll
            for (int j = 0; j < QT; ++j) {
                float a = 0.f;
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    a = fmaf(qv[j][jj].x, qv[j][jj].x, a);
                    a = fmaf(qv[j][jj].y, qv[j][jj].y, a);
                    a = fmaf(qv[j][jj].z, qv[j][jj].z, a);
                    a = fmaf(qv[j][jj].w, qv[j][jj].w, a);
                }
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) a += __shfl_xor_sync(0xffffffffu, a, step);
                if (lane == 0 && (uint32_t)j < myq) s.qn[g + ng * j] = a;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.qempty[qbuf]);
        if (++qbuf == 2) {
            qbuf = 0;
            qphase ^= 1;
        }
        if (ctid < qcount) {
            s.cnt[ctid] = 0;
            s.spair[ctid] = my_slot;
            s.sqidx[ctid] = my_q;
            s.thr[ctid] = my_thr;
        }
        consumer_bar();

        uint32_t since_check = 0;
        bool over = false;  // this thread pushed a candidate at or beyond the compaction mark
        // dot-form screen only once every query of the tile has a finite bound (re-evaluated at each check; a stale
        // "not yet" merely keeps the exact form a little longer)
        bool warm = dotform && !__any_sync(0xffffffffu, lane < qcount && s.thr[lane] == INFINITY);
        for (uint32_t pgi = 0; pgi < it.npg; ++pgi) {
            const uint32_t row_base = it.row_base + pgi * p.lt.page_rows;  // list-relative row of the page start
            const uint32_t rows_in_page = min(p.lt.page_rows, it.rows_left - pgi * p.lt.page_rows);
            for (uint32_t r0 = 0; r0 < rows_in_page; r0 += p.stage_rows) {
                const uint32_t nr = min(p.stage_rows, rows_in_page - r0);
                mbar_wait(&s.full[stage], phase);
                // this warp's rows of the stage, one at a time: row slice -> registers (lane owns float4 columns
                // lane, lane+32, ...), score against the tile; the slot is handed back once the last row is read
                const float4* st4 = reinterpret_cast<const float4*>(s.stages + (size_t)stage * p.stage_rows * ld);
                const uint32_t cur = stage;
                if (++stage == p.S) {
                    stage = 0;
                    phase ^= 1;
                }
                // this warp's rows of the stage (wg, wg + nwg, ...), RB at a time
                if (fullw) {
                    if (ng == 1) {
                        score_dispatch<NJ, QT, 2, true>(p, s, st4, ld4, cur, wg, nwg, nr, row_base + r0, qv, myq, g, ng,
                                                        lane, limit, over, true, pg0 + pgi, r0, warm);
                    } else {
#pragma unroll 1
                        for (uint32_t i0 = 0; i0 < rows_per_warp; i0 += 4)
                            score_dispatch<NJ, QT, 4, true>(p, s, st4, ld4, cur, wg + nwg * i0, nwg, nr, row_base + r0,
                                                            qv, myq, g, ng, lane, limit, over, i0 + 4 >= rows_per_warp,
                                                            pg0 + pgi, r0, warm);
                    }
                } else if (ng == 1) {
                    score_dispatch<NJ, QT, 2, false>(p, s, st4, ld4, cur, wg, nwg, nr, row_base + r0, qv, myq, g, ng, lane,
                                                     limit, over, true, pg0 + pgi, r0, warm);
                } else {
#pragma unroll 1
                    for (uint32_t i0 = 0; i0 < rows_per_warp; i0 += 4)
                        score_dispatch<NJ, QT, 4, false>(p, s, st4, ld4, cur, wg + nwg * i0, nwg, nr, row_base + r0, qv,
                                                         myq, g, ng, lane, limit, over, i0 + 4 >= rows_per_warp,
                                                         pg0 + pgi, r0, warm);
                }

                if (++since_check == p.check_interval) {
                    since_check = 0;
                    // counts only grow between checks, so some pool passed the mark iff some thread saw it
                    const bool need = consumer_bar_or(over);
                    over = false;
                    if (need) {
                        for (uint32_t j = warp; j < qcount; j += CONSUMER_WARPS)
                            if (s.cnt[j] > limit) compact_pool(s, p, j, lane);
                        consumer_bar();
                    }
                    // pick up bounds other CTAs published meanwhile (monotone, so a racy read is still a valid bound)
                    if (ctid < qcount) s.thr[ctid] = fminf(s.thr[ctid], key2f(__ldcg(&p.qthr[s.sqidx[ctid]])));
                    if (dotform && !warm) warm = !__any_sync(0xffffffffu, lane < qcount && s.thr[lane] == INFINITY);
                }
            }
        }

        // item done: best k of every query of the tile -> its partial-result slot
        consumer_bar();
        for (uint32_t j = warp; j < qcount; j += CONSUMER_WARPS) {
            compact_pool(s, p, j, lane);
            // entries beyond the bound the whole grid has proved by now cannot reach the final top-k: drop them
            // here, so that most partials leave the kernel (nearly) empty and the merge has little to do
            const float bound = s.thr[j];  // min(the grid-wide bound as last refreshed, this pool's own k-th)
            const uint32_t kept = s.cnt[j];
            uint32_t nc = 0;
            for (uint32_t i0 = 0; i0 < kept; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool ok = i < kept && s.pool_d[(size_t)j * p.P + i] <= bound;
                nc += __popc(__ballot_sync(0xffffffffu, ok));  // sorted ascending: survivors are a prefix
            }
            const size_t slot = s.spair[j];
            for (uint32_t i = lane; i < nc; i += 32) {
                p.part_d[slot * p.k + i] = s.pool_d[(size_t)j * p.P + i];
                p.part_i[slot * p.k + i] = s.pool_i[(size_t)j * p.P + i];
            }
            if (lane == 0) p.part_cnt[slot] = nc;
            if (nc > 0) contribute_global(s, p, j, nc, lane);
        }
        consumer_bar();
    }
}

__device__ __forceinline__ void producer_loop(const ScanParams& p, const ScanSmem& s, const bool dotform) {
    const uint32_t total = *p.totals;
    const uint32_t ld = p.lt.ld;
    constexpr uint32_t END = 0xffffffffu;
    // rows that a later query tile of the same item will read again should stay in L2; the last (or only)
    // pass streams and must not push them out
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
    uint32_t stage = 0, phase = 0, qbuf = 0, qphase = 0;
    // items are handed out dynamically; the next one is claimed and its descriptor fetched while the
    // current one streams, so the ring never waits on that round trip
    uint32_t ii = atomicAdd(p.work_counter, 1u);
    ScanItem it = p.items[min(ii, total ? total - 1 : 0)];
    while (ii < total) {
        const uint32_t ii_next = atomicAdd(p.work_counter, 1u);
        const ScanItem it_next = p.items[min(ii_next, total - 1)];
        const uint64_t src0 = p.lt.page_vec[it.pg0];
        const uint64_t ids0 = p.has_ids ? p.lt.page_ids[it.pg0] : 0;
        for (uint32_t g0 = 0; g0 < it.gcount; g0 += p.qt) {
            const uint32_t qcount = min(p.qt, it.gcount - g0);
            const uint64_t policy = (g0 + p.qt < it.gcount) ? keep : stream;
            // announce the tile and stage its queries (consumers need them before the first row)
            mbar_wait(&s.qempty[qbuf], qphase ^ 1);
            uint32_t* tw = s.tile + qbuf * 8;
            tw[0] = ii;
            tw[1] = it.gbase + g0;
            tw[2] = qcount;
            tw[3] = it.range;
            tw[4] = it.npg;
            tw[5] = it.row_base;
            tw[6] = it.rows_left;
            tw[7] = it.pg0;
            // per query of the tile: its index and the partial-result slot of (pair, range), so that the
            // consumers' tile set-up touches no global memory except the bound
#pragma unroll 4
            for (uint32_t j = 0; j < qcount; ++j) {
                const uint32_t pair = p.gpairs[it.gbase + g0 + j];
                s.tq[qbuf * MAX_QT + j] = pair / p.np;
                s.tslot[qbuf * MAX_QT + j] = p.pair_slot[pair] + it.range;
            }
            mbar_expect_tx(&s.qfull[qbuf], qcount * ld * 4);  // release: publishes the tile words too
            for (uint32_t j = 0; j < qcount; ++j)
                tma_bulk_g2s(s.sq + ((size_t)qbuf * p.qt + j) * ld, p.queries + (size_t)s.tq[qbuf * MAX_QT + j] * ld,
                             ld * 4, &s.qfull[qbuf]);
            if (++qbuf == 2) {
                qbuf = 0;
                qphase ^= 1;
            }
            // the item's rows; from the second tile on they come back from L2
            for (uint32_t pgi = 0; pgi < it.npg; ++pgi) {
                const uint32_t pg = it.pg0 + pgi;
                const uint32_t rows_in_page = min(p.lt.page_rows, it.rows_left - pgi * p.lt.page_rows);
                const float* src = reinterpret_cast<const float*>(pgi ? p.lt.page_vec[pg] : src0);
                const uint64_t* ids =
                    p.has_ids ? reinterpret_cast<const uint64_t*>(pgi ? p.lt.page_ids[pg] : ids0) : nullptr;
                for (uint32_t r0 = 0; r0 < rows_in_page; r0 += p.stage_rows) {
                    const uint32_t nr = min(p.stage_rows, rows_in_page - r0);
                    const uint32_t bytes = nr * ld * 4;
                    // bulk copies move multiples of 16 bytes: an odd tail reads one id slot further, inside the page
                    const uint32_t id_bytes = ids ? ((nr + 1) & ~1u) * 8 : 0;
                    // row norms sit behind the page's ids ([page_rows] u64, then [page_rows] f32)
                    const uint32_t norm_bytes = (ids && dotform) ? ((nr + 3) & ~3u) * 4 : 0;
                    mbar_wait(&s.empty[stage], phase ^ 1);
                    mbar_expect_tx(&s.full[stage], bytes + id_bytes + norm_bytes);
                    tma_bulk_g2s_hint(s.stages + (size_t)stage * p.stage_rows * ld, src + (size_t)r0 * ld, bytes,
                                      &s.full[stage], policy);
                    if (ids) tma_bulk_g2s(s.stage_ids + stage * STAGE_ROWS, ids + r0, id_bytes, &s.full[stage]);
                    if (norm_bytes)
                        tma_bulk_g2s(s.stage_norm + stage * STAGE_ROWS,
                                     reinterpret_cast<const float*>(ids + p.lt.page_rows) + r0, norm_bytes, &s.full[stage]);
                    if (++stage == p.S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        ii = ii_next;
        it = it_next;
    }
    mbar_wait(&s.qempty[qbuf], qphase ^ 1);
    s.tile[qbuf * 8 + 0] = END;
    mbar_arrive(&s.qfull[qbuf]);
}

template <int NJ>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(const __grid_constant__ ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const ScanSmem s = carve(smem_raw, p);
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < p.S; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], CONSUMER_WARPS);
        }
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(&s.qfull[i], 1);
            mbar_init(&s.qempty[i], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    // register reallocation between warpgroups: the producer's warpgroup (one busy thread) shrinks to 40
    // registers per thread, the two consumer warpgroups grow to 232 = (384 * 168 - 128 * 40) / 256
    // The dot-form screen pays where the scan is long and runs into the power cap (whole index: +4.7 %, half: +2.3 %);
    // on a short scan (1/8 shard, full clocks) its admitted pairs -- k ln(rows/k) per query whatever the shard size,
    // each an L2 round trip for one warp -- cost more than the halved FP work saves (-2.4 %), so a launch that streams
    // fewer than 20 000 distinct rows per CTA (between the 1/2 and the 1/4 shard of the headline index) keeps the
    // exact form.  Uniform over the grid: both loops read it here.
    const bool dotform = p.dotform && p.stats[1] >= (unsigned long long)p.dot_min_rows * gridDim.x;
    if (threadIdx.x < CONSUMER_THREADS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(232));
        consumer_loop<NJ>(p, s, dotform);
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(40));
        if (threadIdx.x == CONSUMER_THREADS) producer_loop(p, s, dotform);
    }
}

// ------------------------------------------------- 2b. bf16 tensor-core screen

#include "screen.cuh"

// ------------------------------------------------------------------ 3. merge

struct MergeParams {
    const float* part_d;
    const uint64_t* part_i;
    const uint32_t* pair_slot;  // null: `parts` mode, slot(q, p) = p * nq + q, one slot per pair
    const uint32_t* part_cnt;   // valid entries per slot; null: k, padding skipped
    const uint32_t* qthr;       // per-query bound on the final k-th distance proved by the scan (ordered keys), or null
    uint32_t nq, np, k, P;
    float* out_d;      // may be null when the result is only published
    uint64_t* out_i;
    uint32_t* out_u32;
    PublishTarget pub;  // sharded search: the merged local top-k goes straight into the peers' mailboxes
};

constexpr uint32_t MERGE_W = 256;  // per-warp scratch entries for the per-list selection

__global__ void __launch_bounds__(MERGE_THREADS) merge_kernel(const MergeParams p) {
    extern __shared__ __align__(16) uint8_t msmem[];
    const uint32_t P = p.P, k = p.k;
    uint64_t* i1 = (uint64_t*)msmem;
    uint64_t* i2 = i1 + P;
    uint64_t* i3 = i2 + P;  // scratch of the de-duplicating compaction
    uint64_t* wi = i3 + P;  // [8 warps][MERGE_W]
    float* d1 = (float*)(wi + (MERGE_THREADS / 32) * MERGE_W);
    float* d2 = d1 + P;
    float* d3 = d2 + P;
    float* wd = d3 + P;
    __shared__ uint32_t cnt1, cnt2;
    __shared__ float thr1, thr2;
    __shared__ uint32_t s_scan[MERGE_THREADS / 32 + 1];
    const MergePool L1{d1, i1, &cnt1, &thr1};  // per-list multiset top-k
    const MergePool L2{d2, i2, &cnt2, &thr2};  // across lists, de-duplicated
    const uint32_t q = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        cnt2 = 0;
        thr2 = INFINITY;
    }
    __syncthreads();

    // Pairs are handled 256 at a time.  A: one warp per pair adds up the entries its page ranges
    // produced.  B: pairs that are already a list's top-k (one range, or <= k entries in all) go
    // straight to the cross-list pool, all at once (warp per pair, lane per slot).  C: the few pairs
    // with more than k survivors get the per-list selection first.
    constexpr uint32_t CH = 256;
    __shared__ uint32_t s_ptot[CH];
    const uint32_t lane = tid & 31, warp = tid >> 5;
    auto pair_slots = [&](uint32_t pr, uint32_t& slot0, uint32_t& ns) {
        if (p.pair_slot) {
            const uint32_t pair = q * p.np + pr;
            slot0 = p.pair_slot[pair];
            ns = p.pair_slot[pair + 1] - slot0;
        } else {
            slot0 = pr * p.nq + q;
            ns = 1;
        }
    };
    for (uint32_t pr0 = 0; pr0 < p.np; pr0 += CH) {
        const uint32_t npc = min(CH, p.np - pr0);
        for (uint32_t pl = warp; pl < npc; pl += MERGE_THREADS / 32) {
            uint32_t slot0, ns;
            pair_slots(pr0 + pl, slot0, ns);
            uint32_t t = 0;
            if (p.part_cnt) {
                for (uint32_t sidx = lane; sidx < ns; sidx += 32) t += p.part_cnt[slot0 + sidx];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            } else {
                t = ns * k;
            }
            if (lane == 0) s_ptot[pl] = t;
        }
        __syncthreads();
        // room for every pair's (at most k) contributions of the chunk at once?
        const bool bulk = (uint64_t)npc * k + k <= P;
        if (bulk) {
            if (cnt2 + npc * k > P) pool_compact_block(L2, P, k, true, d3, i3, s_scan);
            // nothing beyond the bound the scan ended with can reach the final top-k (k distinct ids are already
            // at or below it), so most of what the early items left in their partials is dropped right here
            const float thr = fminf(thr2, p.qthr ? key2f(__ldg(&p.qthr[q])) : INFINITY);
            float* wd_ = wd + warp * MERGE_W;
            uint64_t* wi_ = wi + warp * MERGE_W;
            for (uint32_t pl = warp; pl < npc; pl += MERGE_THREADS / 32) {
                uint32_t slot0, ns;
                pair_slots(pr0 + pl, slot0, ns);
                const uint32_t tot = s_ptot[pl];
                if (tot == 0) continue;
                const bool direct = (ns == 1) || (tot <= k);
                if (!direct && tot > MERGE_W) continue;  // left to the block-wide path below
                // the pair's partial slots are contiguous, so (slot, entry) flattens to one coalesced index range;
                // valid entries go straight into the cross-list pool when they already are the list's top-k,
                // else into this warp's scratch for the per-list selection
                uint32_t wcnt = 0;
                const uint32_t span = ns * k;
                const float* sd = p.part_d + (size_t)slot0 * k;
                const uint64_t* si = p.part_i + (size_t)slot0 * k;
                for (uint32_t i0 = 0; i0 < span; i0 += 128) {
                    float dv[4];
                    uint64_t iv[4];
                    bool ok[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t idx = i0 + u * 32 + lane;
                        const uint32_t sidx = idx / k, e = idx - sidx * k;
                        ok[u] = idx < span && e < (p.part_cnt ? p.part_cnt[slot0 + sidx] : k);
                        dv[u] = ok[u] ? sd[idx] : FLT_MAX;
                        iv[u] = ok[u] ? si[idx] : ID_PAD;
                        ok[u] = ok[u] && !(iv[u] == ID_PAD && dv[u] == FLT_MAX);  // padding of a short partial
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (direct) {
                            if (ok[u] && dv[u] <= thr) {
                                const uint32_t pos = atomicAdd(&cnt2, 1u);
                                if (pos < P) {
                                    d2[pos] = dv[u];
                                    i2[pos] = iv[u];
                                }
                            }
                        } else {
                            ok[u] = ok[u] && dv[u] <= thr;
                            const uint32_t m = __ballot_sync(0xffffffffu, ok[u]);
                            if (ok[u]) {
                                const uint32_t pos = wcnt + __popc(m & ((1u << lane) - 1u));
                                if (pos < MERGE_W) {
                                    wd_[pos] = dv[u];
                                    wi_[pos] = iv[u];
                                }
                            }
                            wcnt += __popc(m);
                        }
                    }
                }
                wcnt = min(wcnt, MERGE_W);
                if (!direct) {
                    // level 1 by one warp: the list's top-k, duplicates kept (search_list_cpu), then on to the pool
                    const uint32_t n2 = dev_next_pow2(max(wcnt, 1u));
                    for (uint32_t i = wcnt + lane; i < n2; i += 32) {
                        wd_[i] = FLT_MAX;
                        wi_[i] = ID_PAD;
                    }
                    __syncwarp();
                    bitonic_sort_pairs(wd_, wi_, n2, lane, 32, [] { __syncwarp(); });
                    const uint32_t take = min(wcnt, k);
                    for (uint32_t i = lane; i < take; i += 32) {
                        if (wd_[i] <= thr) {
                            const uint32_t pos = atomicAdd(&cnt2, 1u);
                            if (pos < P) {
                                d2[pos] = wd_[i];
                                i2[pos] = wi_[i];
                            }
                        }
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        for (uint32_t pl = 0; pl < npc; ++pl) {
            const uint32_t tot = s_ptot[pl];
            if (tot == 0) continue;
            uint32_t slot0, ns;
            pair_slots(pr0 + pl, slot0, ns);
            const bool direct = (ns == 1) || (tot <= k);
            if (bulk && (direct || tot <= MERGE_W)) continue;  // done by the warps above
            if (direct) {
                // (large k) one pair at a time into the cross-list pool
                for (uint32_t sidx = 0; sidx < ns; ++sidx) {
                    const uint32_t n = p.part_cnt ? p.part_cnt[slot0 + sidx] : k;
                    if (n == 0) continue;
                    if (cnt2 + n > P) pool_compact_block(L2, P, k, true, d3, i3, s_scan);
                    pool_push_block(L2, P, p.part_d + (size_t)(slot0 + sidx) * k, p.part_i + (size_t)(slot0 + sidx) * k, n);
                }
                continue;
            }
            // level 1: the list's page-range partials -> its top-k, duplicates kept (search_list_cpu)
            if (tid == 0) {
                cnt1 = 0;
                thr1 = INFINITY;
            }
            __syncthreads();
            if (tot <= P) {
                for (uint32_t sidx = tid; sidx < ns; sidx += MERGE_THREADS) {
                    const uint32_t n = p.part_cnt ? p.part_cnt[slot0 + sidx] : k;
                    const float* sd = p.part_d + (size_t)(slot0 + sidx) * k;
                    const uint64_t* si = p.part_i + (size_t)(slot0 + sidx) * k;
                    for (uint32_t e = 0; e < n; ++e) {
                        const uint32_t pos = atomicAdd(&cnt1, 1u);
                        if (pos < P) {
                            d1[pos] = sd[e];
                            i1[pos] = si[e];
                        }
                    }
                }
                __syncthreads();
            } else {
                for (uint32_t sidx = 0; sidx < ns; ++sidx) {
                    const uint32_t n = p.part_cnt ? p.part_cnt[slot0 + sidx] : k;
                    if (n == 0) continue;
                    if (cnt1 + n > P) pool_compact_block(L1, P, k, false, nullptr, nullptr, s_scan);
                    pool_push_block(L1, P, p.part_d + (size_t)(slot0 + sidx) * k, p.part_i + (size_t)(slot0 + sidx) * k, n);
                }
            }
            pool_compact_block(L1, P, k, false, nullptr, nullptr, s_scan);
            const uint32_t n = cnt1;
            if (cnt2 + n > P) pool_compact_block(L2, P, k, true, d3, i3, s_scan);
            pool_push_block(L2, P, d1, i1, n);
        }
        __syncthreads();
    }
    pool_compact_block(L2, P, k, true, d3, i3, s_scan);
    const uint32_t nc = cnt2;
    if (p.out_d) {
        for (uint32_t i = tid; i < k; i += MERGE_THREADS) {
            const float d = i < nc ? d2[i] : FLT_MAX;
            const uint64_t id = i < nc ? i2[i] : ID_PAD;
            p.out_d[(size_t)q * k + i] = d;
            p.out_i[(size_t)q * k + i] = id;
            if (p.out_u32) p.out_u32[(size_t)q * k + i] = (id == ID_PAD) ? 0xffffffffu : (uint32_t)id;
        }
    }
    if (p.pub.enabled) publish_query(p.pub, q, k, d2, i2, nc, MERGE_THREADS);
}

uint32_t merge_pool_size(uint32_t k) { return next_pow2(k * 2 < 1024 ? 1024 : k * 2); }

template <int NJ>
int32_t launch_scan(const ScanParams& sp, uint32_t grid, uint32_t smem, cudaStream_t stream) {
    static bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !configured[dev]) {
        VDB_CUDA_TRY(cudaFuncSetAttribute(scan_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)SMEM_BUDGET));
        configured[dev] = true;
    }
    scan_kernel<NJ><<<grid, SCAN_THREADS, smem, stream>>>(sp);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

template <int NJ, bool I8>
int32_t launch_screen(const screen::Params& sp, uint32_t grid, uint32_t smem, cudaStream_t stream) {
    static bool configured[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !configured[dev]) {
        VDB_CUDA_TRY(cudaFuncSetAttribute(screen::screen_kernel<NJ, I8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)SMEM_BUDGET));
        configured[dev] = true;
    }
    screen::screen_kernel<NJ, I8><<<grid, screen::THREADS, smem, stream>>>(sp);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}
template <bool I8>
int32_t launch_screen_width(const screen::Params& mp, uint32_t ld, uint32_t grid, uint32_t smem, cudaStream_t stream) {
    switch (ld) {
        case 128: return launch_screen<1, I8>(mp, grid, smem, stream);
        case 256: return launch_screen<2, I8>(mp, grid, smem, stream);
        case 512: return launch_screen<4, I8>(mp, grid, smem, stream);
        case 768: return launch_screen<6, I8>(mp, grid, smem, stream);
        default: return launch_screen<8, I8>(mp, grid, smem, stream);
    }
}

// pool entries per query of the screen kernel: room for the global-bound merge (2k) and for a useful number of
// pushes between compactions
uint32_t screen_pool(uint32_t k) { return next_pow2(std::max(2 * k, 32u)); }

}  // namespace

int32_t scan_max_k() { return (int32_t)MAX_K; }

uint32_t screen_max_batch() { return (uint32_t)screen::NQ; }
uint32_t screen_max_k() { return screen::POOL_ENTRIES / 2; }  // screen_pool(k) = pow2 >= 2k must fit the pool region

bool screen_supported(uint32_t ld, uint32_t page_rows, int metric) {
    const bool width = ld == 128 || ld == 256 || ld == 512 || ld == 768 || ld == 1024;  // 32 * NJ float4, NJ in {1,2,4,6,8}
    return width && page_rows % MIRROR_TILE_ROWS == 0 && (metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP) &&
           screen::smem_bytes(ld, 3) <= SMEM_BUDGET;
}

int32_t ScanWorkspace::reserve(uint32_t nlists, uint32_t npairs, uint64_t nslots, uint32_t k, uint32_t nq) {
    if ((uint64_t)nq * k > cap_gtop) {
        cudaFree(gtop_d); cudaFree(gtop_i);
        cap_gtop = (uint64_t)nq * k;
        VDB_CUDA_TRY(cudaMalloc(&gtop_d, cap_gtop * 4));
        VDB_CUDA_TRY(cudaMalloc(&gtop_i, cap_gtop * 8));
    }
    if (nq > cap_glock) {
        cudaFree(glock);
        cap_glock = nq;
        VDB_CUDA_TRY(cudaMalloc(&glock, (size_t)cap_glock * 4));
    }
    cudaGetDevice(&device);
    if (nlists + 1 > cap_lists) {
        cudaFree(gcount); cudaFree(gfill); cudaFree(goff); cudaFree(ioff);
        cap_lists = nlists + 1;
        VDB_CUDA_TRY(cudaMalloc(&gcount, cap_lists * 4));
        VDB_CUDA_TRY(cudaMalloc(&gfill, cap_lists * 4));
        VDB_CUDA_TRY(cudaMalloc(&goff, cap_lists * 4));
        VDB_CUDA_TRY(cudaMalloc(&ioff, cap_lists * 4));
    }
    if (npairs + 1 > cap_pairs) {
        cudaFree(gpairs); cudaFree(pair_slot);
        cap_pairs = npairs + 1;
        VDB_CUDA_TRY(cudaMalloc(&gpairs, (size_t)cap_pairs * 4));
        VDB_CUDA_TRY(cudaMalloc(&pair_slot, (size_t)cap_pairs * 4));
    }
    if (nslots > cap_slots) {
        cudaFree(items); cudaFree(part_cnt);
        cap_slots = nslots + nslots / 4 + 64;
        VDB_CUDA_TRY(cudaMalloc(&items, cap_slots * sizeof(ScanItem)));
        VDB_CUDA_TRY(cudaMalloc(&part_cnt, cap_slots * 4));
    }
    const uint32_t nq_need = npairs + 1;  // nq <= npairs
    if (nq_need > cap_q) {
        cudaFree(qthr);
        cap_q = nq_need;
        VDB_CUDA_TRY(cudaMalloc(&qthr, (size_t)cap_q * 4));
    }
    if (cap_slots * k > cap_part) {
        cudaFree(part_d); cudaFree(part_i);
        cap_part = cap_slots * k;
        VDB_CUDA_TRY(cudaMalloc(&part_d, cap_part * 4));
        VDB_CUDA_TRY(cudaMalloc(&part_i, cap_part * 8));
    }
    if (!qconst) {  // bf16 image of a batch's queries for the screen kernel: 64 slots at the widest supported row
        VDB_CUDA_TRY(cudaMalloc(&qimg, (size_t)screen::NQ * 1024 * 2));
        VDB_CUDA_TRY(cudaMalloc(&qconst, (size_t)screen::NQ * 32));
    }
    if (!totals) {
        VDB_CUDA_TRY(cudaMalloc(&totals, 4 * 4));
        VDB_CUDA_TRY(cudaMalloc(&stats, 4 * 8));
        VDB_CUDA_TRY(cudaMemset(stats, 0, 4 * 8));
    }
    bytes = (uint64_t)cap_lists * 16 + (uint64_t)cap_pairs * 8 + cap_slots * sizeof(ScanItem) + cap_part * 12 + 48 +
            (uint64_t)screen::NQ * (1024 * 2 + 32);
    return VDB_OK;
}

void ScanWorkspace::release() {
    cudaFree(gcount); cudaFree(gfill); cudaFree(goff); cudaFree(ioff);
    cudaFree(gpairs); cudaFree(pair_slot); cudaFree(items); cudaFree(part_cnt); cudaFree(qthr);
    cudaFree(gtop_d); cudaFree(gtop_i); cudaFree(glock);
    cudaFree(part_d); cudaFree(part_i); cudaFree(totals); cudaFree(stats);
    cudaFree(qimg); cudaFree(qconst);
    *this = ScanWorkspace();
}

int32_t scan_plan(const ListTable& lt, const float* queries_dev, uint32_t nq, const uint32_t* probes_dev, uint32_t np,
                  uint32_t k, int metric, uint32_t ppi, uint64_t max_slots, bool has_ids, uint32_t max_ctas,
                  ScanWorkspace& ws, ScanPlan* out) {
    VDB_REQUIRE(k >= 1 && k <= MAX_K, "k must be in [1, 2048]");
    VDB_REQUIRE(lt.ld % 4 == 0 && lt.ld >= 4 && lt.ld <= 2048, "row stride must be a multiple of 4 floats, <= 2048");
    VDB_REQUIRE(lt.page_rows % STAGE_ROWS == 0, "page_rows must be a multiple of 16");
    VDB_REQUIRE(metric == VDB_METRIC_L2 || metric == VDB_METRIC_IP, "metric must be L2 or InnerProduct");
    VDB_REQUIRE(nq >= 1 && np >= 1 && ppi >= 1, "empty search");
    const uint64_t npairs64 = (uint64_t)nq * np;
    VDB_REQUIRE(npairs64 < (1ull << 31) && max_slots < (1ull << 31), "search too large for one call");

    // shared-memory plan: pool size P (per query), query tile QT (fixed by the row width), ring depth S
    const uint32_t nj_need = (lt.ld / 4 + 31) / 32;
    const uint32_t NJ = nj_need <= 1 ? 1 : nj_need <= 2 ? 2 : nj_need <= 4 ? 4 : nj_need <= 6 ? 6 : nj_need <= 8 ? 8
                        : nj_need <= 12 ? 12 : 16;
    // CTA tile = register tile (tile_queries) x warp groups, as wide as shared memory allows with a 3-deep
    // ring; shrunk below the register tile when the pools of a large k need the room
    const uint32_t QTreg = (uint32_t)tile_queries((int)NJ);
    const uint32_t P = next_pow2(std::max(k + 64, 2 * k));
    const uint32_t stage_rows = lt.ld > 1024 ? STAGE_ROWS / 2 : STAGE_ROWS;
    auto fits = [&](uint32_t s_, uint32_t qt_) { return scan_smem_bytes(lt.ld, s_, qt_, P, stage_rows) <= SMEM_BUDGET; };
    uint32_t ngroups = CONSUMER_WARPS;
    while (ngroups > 1 && (QTreg * ngroups > MAX_QT || !fits(3, QTreg * ngroups))) ngroups >>= 1;
    uint32_t QT = QTreg * ngroups;
    while (QT > 1 && !fits(3, QT)) QT >>= 1;
    uint32_t S = 4;
    while (S > 2 && !fits(S, QT)) --S;
    VDB_REQUIRE(fits(S, QT), "dimension * k too large for the scan kernel's shared memory");

    VDB_TRY(ws.reserve(lt.nlist, (uint32_t)npairs64, std::max<uint64_t>(max_slots, 1), k, nq));

    int dev = 0, sms = NUM_SMS_B200;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint64_t grid = std::min<uint64_t>((uint64_t)sms, std::max<uint64_t>(max_slots, 1));
    if (max_ctas) grid = std::min<uint64_t>(grid, max_ctas);

    ScanPlan pl;
    pl.lt = lt;
    pl.queries = queries_dev;
    pl.probes = probes_dev;
    pl.nq = nq; pl.np = np; pl.k = k; pl.metric = metric; pl.ppi = ppi;
    pl.has_ids = has_ids;
    pl.dot_min_rows = 20000;
    pl.ppi_max = ppi;  // the caller may allow the grouping kernel to lengthen the items (index scans)
    pl.has_norms = false;  // set by the caller for index pages (ids followed by row norms)
    pl.info = ScanLaunchInfo{QT, P, S, NJ, (uint32_t)grid, scan_smem_bytes(lt.ld, S, QT, P, stage_rows),
                             std::max(1u, std::min(64u, (P - k) / STAGE_ROWS))};
    pl.stage_rows = stage_rows;
    // pages with a low-precision shadow: the tensor-core screen streams a half / a quarter of the bytes (one batch of
    // <= 64 queries per launch; wider batches and unsupported shapes keep the fp32 kernel)
    pl.mirror = lt.mirror_off != 0 && has_ids && nq <= (uint32_t)screen::NQ && screen_pool(k) <= screen::POOL_ENTRIES &&
                screen_supported(lt.ld, lt.page_rows, metric);
    pl.info.mirror = pl.mirror ? 1u : 0u;
    *out = pl;
    return VDB_OK;
}

int32_t scan_enqueue_groups(const ScanPlan& pl, ScanWorkspace& ws, cudaStream_t stream) {
    const ListTable& lt = pl.lt;
    const uint32_t npairs = pl.nq * pl.np;
    WorkList wl{ws.gcount, ws.gfill, ws.goff, ws.ioff, ws.gpairs, ws.pair_slot, ws.items, ws.totals, ws.stats,
                ws.qthr, pl.nq, pl.lifetime_rows};
    // running top-k of every query: "empty" = (3.39e38, UINT64_MAX), i.e. bytes 0x7f / 0xff; locks open
    VDB_CUDA_TRY(cudaMemsetAsync(ws.gtop_d, 0x7f, (size_t)pl.nq * pl.k * 4, stream));
    VDB_CUDA_TRY(cudaMemsetAsync(ws.gtop_i, 0xff, (size_t)pl.nq * pl.k * 8, stream));
    VDB_CUDA_TRY(cudaMemsetAsync(ws.glock, 0, (size_t)pl.nq * 4, stream));
    if (pl.mirror) {
        screen::query_image_kernel<<<screen::NQ * 32 / 256, 256, 0, stream>>>(pl.queries, pl.nq, lt.ld, lt.mirror_kind,
                                                                              ws.qimg, reinterpret_cast<float4*>(ws.qconst));
        VDB_CUDA_TRY(cudaGetLastError());
    }
    if (lt.nlist <= 8192) {
        const uint32_t gsm = 4 * (lt.nlist + 1) * 4;
        static bool gconf[16] = {false};
        int gdev = 0;
        cudaGetDevice(&gdev);
        if (gdev < 16 && !gconf[gdev]) {
            VDB_CUDA_TRY(cudaFuncSetAttribute(build_groups_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              4 * 8193 * 4));
            VDB_CUDA_TRY(cudaFuncSetAttribute(build_groups_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                              cudaSharedmemCarveoutMaxShared));
            gconf[gdev] = true;
        }
        build_groups_kernel<true><<<1, 1024, gsm, stream>>>(lt, pl.probes, npairs, pl.ppi, pl.ppi_max, pl.info.grid, wl);
    } else {
        build_groups_kernel<false><<<1, 1024, 0, stream>>>(lt, pl.probes, npairs, pl.ppi, pl.ppi_max, pl.info.grid, wl);
    }
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t scan_enqueue_scan(const ScanPlan& pl, ScanWorkspace& ws, cudaStream_t stream) {
    ScanParams sp;
    sp.lt = pl.lt;
    sp.queries = pl.queries;
    sp.items = ws.items;
    sp.totals = ws.totals;
    sp.gpairs = ws.gpairs;
    sp.pair_slot = ws.pair_slot;
    sp.part_d = ws.part_d;
    sp.part_i = ws.part_i;
    sp.part_cnt = ws.part_cnt;
    sp.qthr = ws.qthr;
    sp.gtop_d = ws.gtop_d;
    sp.gtop_i = ws.gtop_i;
    sp.glock = ws.glock;
    sp.k = pl.k; sp.P = pl.info.P; sp.S = pl.info.S; sp.np = pl.np;
    sp.has_ids = pl.has_ids ? 1u : 0u;
    sp.dot_min_rows = pl.dot_min_rows;
    sp.stats = ws.stats;
    sp.dotform = (pl.has_norms && pl.has_ids && pl.metric == VDB_METRIC_L2) ? 1u : 0u;
    sp.qt = pl.info.QT;
    sp.stage_rows = pl.stage_rows;
    sp.work_counter = ws.totals + 2;
    sp.check_interval = pl.info.check_interval;
    sp.metric = pl.metric;
    const uint32_t grid = pl.info.grid, smem = pl.info.smem_bytes;
    if (pl.mirror) {
        screen::Params mp;
        mp.sp = sp;
        mp.sp.P = screen_pool(pl.k);
        mp.qimg = ws.qimg;
        mp.qconst = reinterpret_cast<const float4*>(ws.qconst);
        const bool i8 = pl.lt.mirror_kind == MIRROR_I8;
        mp.nkb = pl.lt.ld * mirror_elem_bytes(pl.lt.mirror_kind) / 128;
        mp.rescored = ws.stats + 2;
        mp.qt = std::min<uint32_t>(screen::NQ, screen::POOL_ENTRIES / mp.sp.P);
        uint32_t S = 8;
        while (S > 3 && screen::smem_bytes(pl.lt.ld, S) > SMEM_BUDGET) --S;
        mp.S = S;
        const uint32_t msmem = screen::smem_bytes(pl.lt.ld, S);
        return i8 ? launch_screen_width<true>(mp, pl.lt.ld, grid, msmem, stream)
                  : launch_screen_width<false>(mp, pl.lt.ld, grid, msmem, stream);
    }
    switch (pl.info.NJ) {
        case 1: VDB_TRY(launch_scan<1>(sp, grid, smem, stream)); break;
        case 2: VDB_TRY(launch_scan<2>(sp, grid, smem, stream)); break;
        case 4: VDB_TRY(launch_scan<4>(sp, grid, smem, stream)); break;
        case 6: VDB_TRY(launch_scan<6>(sp, grid, smem, stream)); break;
        case 8: VDB_TRY(launch_scan<8>(sp, grid, smem, stream)); break;
        case 12: VDB_TRY(launch_scan<12>(sp, grid, smem, stream)); break;
        default: VDB_TRY(launch_scan<16>(sp, grid, smem, stream)); break;
    }
    return VDB_OK;
}

static int32_t configure_merge() {
    static bool mconf[16] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && !mconf[dev]) {
        VDB_CUDA_TRY(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          4096 * 36 + (MERGE_THREADS / 32) * MERGE_W * 12));
        // every kernel of the search pipeline asks for the same (maximum) shared-memory carve-out, so the SMs
        // are not reconfigured between launches
        VDB_CUDA_TRY(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        mconf[dev] = true;
    }
    return VDB_OK;
}

int32_t scan_enqueue_merge(const ScanPlan& pl, ScanWorkspace& ws, float* out_d, uint64_t* out_i, uint32_t* out_u32,
                           const PublishTarget* pub, cudaStream_t stream) {
    MergeParams mp{};
    mp.part_d = ws.part_d; mp.part_i = ws.part_i; mp.pair_slot = ws.pair_slot; mp.part_cnt = ws.part_cnt;
    mp.qthr = ws.qthr;
    mp.nq = pl.nq; mp.np = pl.np; mp.k = pl.k; mp.P = merge_pool_size(pl.k);
    mp.out_d = out_d; mp.out_i = out_i; mp.out_u32 = out_u32;
    if (pub) mp.pub = *pub;
    VDB_TRY(configure_merge());
    const uint32_t msmem = mp.P * 36 + (MERGE_THREADS / 32) * MERGE_W * 12;
    merge_kernel<<<pl.nq, MERGE_THREADS, msmem, stream>>>(mp);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t scan_search(const ListTable& lt, const float* queries_dev, uint32_t nq, const uint32_t* probes_dev,
                    uint32_t np, uint32_t k, int metric, uint32_t ppi, uint64_t max_slots, ScanWorkspace& ws,
                    bool has_ids, float* out_d, uint64_t* out_i, uint32_t* out_u32, cudaStream_t stream,
                    ScanLaunchInfo* info) {
    ScanPlan pl;
    VDB_TRY(scan_plan(lt, queries_dev, nq, probes_dev, np, k, metric, ppi, max_slots, has_ids, 0, ws, &pl));
    VDB_TRY(scan_enqueue_groups(pl, ws, stream));
    VDB_TRY(scan_enqueue_scan(pl, ws, stream));
    VDB_TRY(scan_enqueue_merge(pl, ws, out_d, out_i, out_u32, nullptr, stream));
    if (info) *info = pl.info;
    return VDB_OK;
}

int32_t merge_parts(const float* dparts, const uint64_t* iparts, uint32_t parts, uint32_t nq, uint32_t k,
                    float* out_d, uint64_t* out_i, cudaStream_t stream) {
    VDB_REQUIRE(k >= 1 && k <= MAX_K && parts >= 1 && nq >= 1, "merge: bad shape");
    MergeParams mp{};
    mp.part_d = dparts; mp.part_i = iparts; mp.pair_slot = nullptr; mp.part_cnt = nullptr; mp.qthr = nullptr;
    mp.nq = nq; mp.np = parts; mp.k = k; mp.P = merge_pool_size(k);
    mp.out_d = out_d; mp.out_i = out_i; mp.out_u32 = nullptr;
    VDB_TRY(configure_merge());
    merge_kernel<<<nq, MERGE_THREADS, mp.P * 36 + (MERGE_THREADS / 32) * MERGE_W * 12, stream>>>(mp);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

}  // namespace vdb
