// Host-visible interface of the scan engine (scan.cu): probe grouping, the
// persistent list-scan kernel with fused top-k, and the (dist,id) merge.
#pragma once
#include "common.cuh"

namespace vdb {

// One unit of scan work: page range `range` of list `list`, scanned once per
// tile of the queries that probe the list.
struct alignas(16) ScanItem {
    uint32_t gbase, gcount;  // the (query, probe) pairs that name this list: gpairs[gbase .. gbase+gcount)
    uint32_t range;          // index of the page range inside the list (selects the partial-result slot)
    uint32_t pg0, npg;       // absolute first page in the page tables, pages in this item
    uint32_t row_base;       // list-relative row of the first page
    uint32_t rows_left;      // rows from row_base to the end of the list
    uint32_t list;
};

// Device scratch of one search call.  Grown on demand, reused across calls.
struct ScanWorkspace {
    int device = 0;
    // per-list arrays (sized nlist+1)
    uint32_t *gcount = nullptr, *gfill = nullptr, *goff = nullptr, *ioff = nullptr;
    uint32_t cap_lists = 0;
    // per-pair arrays
    uint32_t *gpairs = nullptr, *pair_slot = nullptr;
    uint32_t cap_pairs = 0;
    // items / partial results
    ScanItem* items = nullptr;
    uint32_t* part_cnt = nullptr;  // [cap_slots]
    uint32_t* qthr = nullptr;      // [cap_q]
    uint32_t cap_q = 0;
    float* gtop_d = nullptr;       // [nq][k] running top-k of every query (bound tightening)
    uint64_t* gtop_i = nullptr;
    uint32_t* glock = nullptr;     // [nq]
    uint64_t cap_gtop = 0;
    uint32_t cap_glock = 0;
    float* part_d = nullptr;
    uint64_t* part_i = nullptr;
    uint64_t cap_slots = 0, cap_part = 0;
    uint8_t* qimg = nullptr;                // bf16 image of the batch's queries (screen kernel), 64 slots
    void* qconst = nullptr;                 // [64] float4 {|q|^2, |q|, |q - bf16(q)|, 0}
    uint32_t* totals = nullptr;             // [0] items, [1] slots
    unsigned long long* stats = nullptr;    // [0] algorithmic rows, [1] unique rows
    uint64_t bytes = 0;

    int32_t reserve(uint32_t nlists, uint32_t npairs, uint64_t nslots, uint32_t k, uint32_t nq);
    void release();
};

struct ScanLaunchInfo {
    uint32_t QT, P, S, NJ, grid, smem_bytes, check_interval;
    uint32_t mirror;  // 1: this launch runs the bf16 tensor-core screen over the pages' shadow
};

struct PublishTarget;  // exchange.cuh

// One search's launch plan: the scan kernel's shared-memory layout for this shape plus the call's arguments.
struct ScanPlan {
    ListTable lt;
    const float* queries;
    const uint32_t* probes;
    uint32_t nq, np, k, ppi, stage_rows;
    uint32_t ppi_max;  // build_groups may lengthen the items up to this many pages (>= ppi), see scan.cu
    int metric;
    bool has_ids;
    unsigned long long* lifetime_rows = nullptr;  // optional device counter: += distinct probed rows of every search
    uint32_t dot_min_rows;  // dot-form screen only for launches with >= this many distinct rows per CTA (default 20000)
    bool has_norms;  // every page's id block is followed by [page_rows] f32 |v|^2 (index pages): L2 may screen by dot product
    bool mirror = false;  // the pages carry a bf16 shadow and the shape fits: the scan runs as the tensor-core screen kernel
    ScanLaunchInfo info;
};

// probes_dev: [nq][np] list ids (entries >= nlist or naming empty lists are
// skipped).  max_slots: caller's upper bound on sum over pairs of page ranges.
// ppi: pages per scan item.  has_ids: every page of `lt` carries an id block
// (false for flat views, whose ids are implicit or in lt.ids_flat).  max_ctas: cap on the persistent scan
// grid (0 = one CTA per SM) -- a pipelined index leaves a few SMs to the kernels of the neighbouring batches.
// The three phases may go to different streams (the caller orders them with events):
//   groups  memsets + build_groups_kernel (needs the probes)
//   scan    scan_kernel
//   merge   merge_kernel -> out_d/out_i [nq][k] (device; optional out_u32 = ids narrowed to 32 bits), and/or
//           straight into the peers' mailboxes (`pub`)
int32_t scan_plan(const ListTable& lt, const float* queries_dev, uint32_t nq, const uint32_t* probes_dev, uint32_t np,
                  uint32_t k, int metric, uint32_t ppi, uint64_t max_slots, bool has_ids, uint32_t max_ctas,
                  ScanWorkspace& ws, ScanPlan* out);
int32_t scan_enqueue_groups(const ScanPlan& pl, ScanWorkspace& ws, cudaStream_t stream);
int32_t scan_enqueue_scan(const ScanPlan& pl, ScanWorkspace& ws, cudaStream_t stream);
int32_t scan_enqueue_merge(const ScanPlan& pl, ScanWorkspace& ws, float* out_d, uint64_t* out_i, uint32_t* out_u32,
                           const PublishTarget* pub, cudaStream_t stream);
// all three on one stream
int32_t scan_search(const ListTable& lt, const float* queries_dev, uint32_t nq, const uint32_t* probes_dev,
                    uint32_t np, uint32_t k, int metric, uint32_t ppi, uint64_t max_slots, ScanWorkspace& ws,
                    bool has_ids, float* out_d, uint64_t* out_i, uint32_t* out_u32, cudaStream_t stream,
                    ScanLaunchInfo* info = nullptr);

// merge_results across `parts` blocks of [nq][k]
int32_t merge_parts(const float* dparts, const uint64_t* iparts, uint32_t parts, uint32_t nq, uint32_t k,
                    float* out_d, uint64_t* out_i, cudaStream_t stream);

int32_t scan_max_k();
// can an index of this row stride / page size / metric keep a low-precision shadow for the tensor-core screen?
bool screen_supported(uint32_t ld, uint32_t page_rows, int metric);
// the screen kernel takes batches of at most screen_max_batch() queries with k <= screen_max_k()
uint32_t screen_max_batch();
uint32_t screen_max_k();

}  // namespace vdb
