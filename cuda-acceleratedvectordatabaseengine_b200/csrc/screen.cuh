// The tensor-core screen of the inverted-list scan (included by scan.cu, inside its unnamed namespace).
//
// The list scan is HBM-bound: every probed fp32 row (3 KB at 768-D) is streamed once per batch.  An index that
// keeps a low-precision shadow of its pages (ListTable::mirror_off / mirror_kind: the rows rounded to bf16, or
// quantised to int8 with one scale per row, stored as ready-made 128-byte-swizzled [128 rows][128 bytes] operand
// tiles, plus |v - shadow(v)| and the scale per row) lets the scan stream a HALF / a QUARTER of the bytes: this kernel
// brings the shadow tiles in with 1-D bulk TMA copies, multiplies them on the tensor cores (tcgen05.mma kind::f16,
// bf16 x bf16 -> fp32, or kind::i8, int8 x int8 -> int32, accumulators in TMEM) against the image of the WHOLE query
// batch (<= 64 queries, resident in shared memory for the kernel's lifetime; int8 queries as two terms, the second
// the quantised residual of the first), and turns every (row, query) dot product into a LOWER BOUND of the exact
// distance.  With q~, v~ the shadows, e_q = |q - q~|, e_v = |v - v~| (exact differences, norms rounded up):
//     q.v - q~.v~ = (q - q~).v + q~.(v - v~)   =>   |q.v - q~.v~| <= e_q |v| + (|q| + e_q) e_v  (+ accumulation slack)
// (Cauchy-Schwarz twice, |q~| <= |q| + e_q).  A pair whose lower bound beats the query's running k-th distance --
// after the first rows of a query almost none does -- is re-scored EXACTLY from the fp32 row in the page with the
// arithmetic of scan_kernel (same per-lane FMA chain, same 16-8-4-2-1 summation tree), so the distances that reach
// the pools, the partial results and the merge are bit for bit those of the fp32 scan: the screen only decides
// which pairs are worth the exact arithmetic.  tests/test_screen_bound.py checks the inequality on the CPU,
// tests/test_scan_mirror.py the bit-identity on the GPU.
//
// Roles (192 threads, one CTA per SM, persistent over dynamically claimed items):
//   warp 4, one thread   producer: claims items, announces row tiles (descriptor ring; the tile's row constants are
//                        bulk-copied into the descriptor's slot), streams the shadow tiles of the item's pages
//                        through an S-stage ring of 16 KB (128 rows x 128 bytes) stages
//   warp 5, one thread   tcgen05.mma issuer + TMEM owner: per row tile (row bytes / 128) stages x 4 MMAs (M128, N64
//                        K16 bf16 / N128 K32 int8) into one of two accumulators; tcgen05.commit frees the stage /
//                        publishes the tile
//   warps 0-3            thread = row of the tile = TMEM lane.  Phase A: read the dot products of the queries
//                        that probe this list (tcgen05.ld), evaluate the lower bound, ballot the admitted pairs
//                        into shared memory, hand the accumulator back.  Phase B: warp w owns queries w, w+4, ...:
//                        exact re-score of their admitted rows (fp32 row from the page, query from global
//                        memory), push into the query's pool (owned by that warp: no atomics), compact when full.
//                        Last tile of an item: pools -> partial-result slots, global bound update (as scan_kernel).
#pragma once

namespace screen {

constexpr int NQ = 64;                    // query slots of a batch; UMMA N = 64 (bf16) or 128 (int8: two terms per query)
constexpr int THREADS = 192;
constexpr int CONSUMERS = 128;
constexpr int RT_RING = 3;                // row-tile descriptors (and row-constant blocks) in flight
constexpr uint32_t A_STAGE = MIRROR_TILE_BYTES;  // 16 KB
__host__ __device__ constexpr uint32_t ncol(bool i8) { return i8 ? 2u * NQ : (uint32_t)NQ; }  // UMMA N
__host__ __device__ constexpr uint32_t b_block(bool i8) { return ncol(i8) * 128u; }  // one K block of the query image
constexpr uint32_t POOL_ENTRIES = 2048;   // candidate-pool entries per CTA, split evenly over a pass's queries
constexpr uint32_t F_FIRST = 1u, F_LAST = 2u, F_END = 0x80000000u;
// tensor-core accumulation (fp32, K <= 1024 products that are exact in fp32) and the fp32 rounding of the exact
// inner-product chain, relative to |q||v|.  Measured through cuBLAS on this part (tools/tc_accumulate_check.py, K = 768
// and 1024, Gaussian / 16-binade / all-positive values): at most 1.3e-6; 2^-14 = 6.1e-5 leaves a factor of ~50.
constexpr float ACC_SLACK = 6.103515625e-5f;
// int8: the integer dot products are exact; what is left is the fp32 arithmetic that puts them back on the fp32 scale
// (three roundings) and the rounding of the exact inner-product chain (< 40 * 2^-24): 2^-17 |q||v| covers both 3x over
constexpr float ACC_SLACK_I8 = 7.62939453125e-6f;

// development aid (make EXTRA=-DVDB_SCREEN_PROF): where the three roles of the first CTAs spend their cycles
#ifdef VDB_SCREEN_PROF
#define SPROF_DECL(...) long long __VA_ARGS__
#define SPROF_T0(t) const long long t = clock64()
#define SPROF_ADD(acc, t) acc += clock64() - t
#else
#define SPROF_DECL(...)
#define SPROF_T0(t)
#define SPROF_ADD(acc, t)
#endif

struct Params {
    ScanParams sp;         // the scan's arguments; sp.P = pool entries per query of this kernel
    const uint8_t* qimg;   // [K blocks][64 or 128][128 B] image of the batch's queries (zero rows beyond nq)
    const float4* qconst;  // [64][2] {|q|^2, |q| rounded up, |q - shadow(q)| rounded up, 0}, {scale a, scale b, 0, 0}
    uint32_t S;            // ring stages
    uint32_t qt;           // queries per pass over an item = min(64, POOL_ENTRIES / P)
    uint32_t nkb;          // K blocks = row bytes of the shadow / 128
    unsigned long long* rescored;  // += (row, query) pairs re-scored exactly (statistics)
};

struct Smem {
    uint8_t* sa;        // [S][16 KB]
    uint8_t* sb;        // [nkb][8 KB]
    uint64_t* pool_i;   // [POOL_ENTRIES]
    float* pool_d;      // [POOL_ENTRIES]
    uint32_t* cnt;      // [64]
    float* thr;         // [64]
    uint32_t* spair;    // [64]
    uint32_t* sqidx;    // [64]
    float* qc;          // [64] |q|^2 (1 - DOT_SLACK)
    float* uc;          // [64] 2 e_q + 2 ACC_SLACK |q|
    float* wc;          // [64] 2 (|q| + e_q)
    float* qsa;         // [64] int8: scale of the query's first term
    float* qsb;         // [64] int8: scale of its second term (the quantised residual)
    uint32_t* adm;      // [64][4] admitted rows of the current tile, per query and consumer warp
    uint8_t* tq;        // [2][64] query index of each slot of a pass          (producer -> consumers)
    uint32_t* tslot;    // [2][64] partial-result slot of each (pair, range)
    uint32_t* rt;       // [RT_RING][8] row-tile descriptors
    float* meta;        // [RT_RING][3][128] |v|^2, |v - shadow(v)|, row scale of the tile's rows (bulk-copied with the tile)
    uint64_t* mfull;    // [RT_RING] the row constants of a descriptor's tile have landed
    uint64_t *full, *empty;        // [8] each
    uint64_t *rtfull, *rtempty;    // [RT_RING]
    uint64_t* pfree;               // [2]
    uint64_t *acc_full, *acc_empty;  // [2]
    uint64_t* bfull;               // [1]
    uint32_t* tmem_slot;
    uint32_t* redo;     // a warp left rows of a cold query out of round 0 of the current tile
};

__host__ __device__ inline uint32_t fixed_bytes() {
    return POOL_ENTRIES * 12 + 9 * NQ * 4 + NQ * 4 * 4 + 2 * NQ * 4 + 2 * NQ + RT_RING * 8 * 4 + RT_RING * 3 * 128 * 4 +
           (8 + 8 + 3 * RT_RING + 2 + 2 + 2 + 1) * 8 + 16;
}
__host__ __device__ inline uint32_t smem_bytes(uint32_t ld, uint32_t S) {
    return 1024 + S * A_STAGE + ld * 128u + fixed_bytes();  // query image: 64 x ld bf16 = 128 x ld int8 = 128 ld bytes
}

__device__ __forceinline__ Smem carve(uint8_t* raw, const Params& p) {
    uint8_t* q = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    Smem s;
    s.sa = q;                  q += p.S * A_STAGE;
    s.sb = q;                  q += p.sp.lt.ld * 128u;
    s.pool_i = (uint64_t*)q;   q += POOL_ENTRIES * 8;
    s.pool_d = (float*)q;      q += POOL_ENTRIES * 4;
    s.cnt = (uint32_t*)q;      q += NQ * 4;
    s.thr = (float*)q;         q += NQ * 4;
    s.spair = (uint32_t*)q;    q += NQ * 4;
    s.sqidx = (uint32_t*)q;    q += NQ * 4;
    s.qc = (float*)q;          q += NQ * 4;
    s.uc = (float*)q;          q += NQ * 4;
    s.wc = (float*)q;          q += NQ * 4;
    s.qsa = (float*)q;         q += NQ * 4;
    s.qsb = (float*)q;         q += NQ * 4;
    s.adm = (uint32_t*)q;      q += NQ * 4 * 4;
    s.meta = (float*)q;        q += RT_RING * 3 * 128 * 4;
    s.tslot = (uint32_t*)q;    q += 2 * NQ * 4;
    s.tq = (uint8_t*)q;        q += 2 * NQ;
    s.rt = (uint32_t*)q;       q += RT_RING * 8 * 4;
    s.full = (uint64_t*)q;     q += 8 * 8;
    s.empty = (uint64_t*)q;    q += 8 * 8;
    s.rtfull = (uint64_t*)q;   q += RT_RING * 8;
    s.rtempty = (uint64_t*)q;  q += RT_RING * 8;
    s.mfull = (uint64_t*)q;    q += RT_RING * 8;
    s.pfree = (uint64_t*)q;    q += 2 * 8;
    s.acc_full = (uint64_t*)q; q += 2 * 8;
    s.acc_empty = (uint64_t*)q; q += 2 * 8;
    s.bfull = (uint64_t*)q;    q += 8;
    s.tmem_slot = (uint32_t*)q;
    s.redo = s.tmem_slot + 1;
    return s;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// int8 x int8 -> int32 (exact)
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, rows 128 bytes apart, 128-byte swizzle, 8-row groups 1024 bytes apart (as tc_common.cuh)
__device__ __forceinline__ uint64_t desc_sw128(const void* tile) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(tile) & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// fp32 accumulator, bf16 x bf16, both K-major
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// int32 accumulator, signed 8-bit x signed 8-bit, both K-major
__host__ __device__ constexpr uint32_t idesc_i8(uint32_t M, uint32_t N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// four accumulator columns of this thread's TMEM lane (one tcgen05.ld each), complete on return
__device__ __forceinline__ void tmem_ld4(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3, float (&v)[4]) {
    uint32_t a, b, c, d;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%4];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%1}, [%5];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%2}, [%6];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%3}, [%7];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
        : "r"(t0), "r"(t1), "r"(t2), "r"(t3)
        : "memory");
    v[0] = __uint_as_float(a);
    v[1] = __uint_as_float(b);
    v[2] = __uint_as_float(c);
    v[3] = __uint_as_float(d);
}
// the same for two columns per query (int8: the query's two terms), as integers
__device__ __forceinline__ void tmem_ld8(const uint32_t (&t)[8], int (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%8];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%1}, [%9];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%2}, [%10];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%3}, [%11];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%4}, [%12];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%5}, [%13];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%6}, [%14];\n"
        "tcgen05.ld.sync.aligned.32x32b.x1.b32 {%7}, [%15];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7])
        : "memory");
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMERS) : "memory"); }

// this lane's share of the exact q.v of one page row: the arithmetic of the inner-product branch of score_batch
template <int NJ>
__device__ __forceinline__ float exact_ip_lane(const float4* __restrict__ g4, uint32_t ld4, uint32_t r, uint32_t lane,
                                               const float4 (&q)[NJ]) {
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
        const float4 v = __ldg(&g4[r * ld4 + lane + 32 * jj]);
        const float2 qlo = make_float2(q[jj].x, q[jj].y);
        acc = jj == 0 ? __fmul2_rn(qlo, make_float2(v.x, v.y)) : __ffma2_rn(qlo, make_float2(v.x, v.y), acc);
        acc = __ffma2_rn(make_float2(q[jj].z, q[jj].w), make_float2(v.z, v.w), acc);
    }
    return acc.x + acc.y;
}

template <bool I8>
__device__ __forceinline__ void producer(const Params& p, const Smem& s) {
    constexpr uint32_t B_BLOCK = b_block(I8);
    const ScanParams& sp = p.sp;
    const uint32_t total = *sp.totals;
    const uint32_t page_rows = sp.lt.page_rows;
    const uint64_t ids_off = (uint64_t)page_rows * sp.lt.ld * 4;  // page = rows | ids | norms | error norms | scales
    const uint64_t keep = l2_policy_evict_last(), stream = l2_policy_evict_first();
    // the batch's queries: the B operand of every MMA of this CTA
    mbar_expect_tx(s.bfull, p.nkb * B_BLOCK);
    for (uint32_t kb = 0; kb < p.nkb; ++kb)
        tma_bulk_g2s(s.sb + kb * B_BLOCK, p.qimg + (size_t)kb * B_BLOCK, B_BLOCK, s.bfull);
    uint32_t stage = 0, phase = 0, ri = 0, rphase = 0, pb = 0, pphase = 0;
    uint32_t ii = atomicAdd(sp.work_counter, 1u);
    ScanItem it = sp.items[min(ii, total ? total - 1 : 0)];
    // page addresses are fetched one page ahead (and the first page's with the item): a dependent global load in
    // front of every page would leave the ring -- one row tile deep -- half empty by the time it returns
    uint64_t pv0 = total ? sp.lt.page_vec[it.pg0] : 0;
    while (ii < total) {
        const uint32_t ii_next = atomicAdd(sp.work_counter, 1u);
        const ScanItem it_next = sp.items[min(ii_next, total - 1)];
        const uint64_t pv0_next = sp.lt.page_vec[it_next.pg0];
        for (uint32_t g0 = 0; g0 < it.gcount; g0 += p.qt) {
            const uint32_t qcount = min(p.qt, it.gcount - g0);
            const uint64_t policy = (g0 + p.qt < it.gcount) ? keep : stream;
            mbar_wait(&s.pfree[pb], pphase ^ 1);
#pragma unroll 4
            for (uint32_t j = 0; j < qcount; ++j) {
                const uint32_t pair = sp.gpairs[it.gbase + g0 + j];
                s.tq[pb * NQ + j] = (uint8_t)(pair / sp.np);
                s.tslot[pb * NQ + j] = sp.pair_slot[pair] + it.range;
            }
            bool first = true;
            uint64_t pv = pv0;
            for (uint32_t pgi = 0; pgi < it.npg; ++pgi) {
                const uint32_t pg = it.pg0 + pgi;
                const uint32_t rows_in_page = min(page_rows, it.rows_left - pgi * page_rows);
                const uint64_t page = pv;
                if (pgi + 1 < it.npg) pv = sp.lt.page_vec[pg + 1];  // used by the next iteration
                const uint8_t* mirror = reinterpret_cast<const uint8_t*>(page) + sp.lt.mirror_off;
                for (uint32_t r0 = 0; r0 < rows_in_page; r0 += MIRROR_TILE_ROWS) {
                    const bool last = pgi + 1 == it.npg && r0 + MIRROR_TILE_ROWS >= rows_in_page;
                    mbar_wait(&s.rtempty[ri], rphase ^ 1);
                    uint32_t* d = s.rt + ri * 8;
                    d[0] = (first ? F_FIRST : 0u) | (last ? F_LAST : 0u);
                    d[1] = qcount;
                    d[2] = pb;
                    d[3] = pg;
                    d[4] = r0;
                    d[5] = min((uint32_t)MIRROR_TILE_ROWS, rows_in_page - r0);
                    d[6] = (uint32_t)page;  // the page's address: consumers need no table lookup of their own
                    d[7] = (uint32_t)(page >> 32);
                    mbar_arrive(&s.rtfull[ri]);  // release: publishes the descriptor (and, on a first tile, the pass arrays)
                    {   // the tile's row constants ride along: a consumer-side load would queue behind ~13 MB of bulk copies
                        const uint32_t nb = ((d[5] + 3u) & ~3u) * 4u, narr = I8 ? 3u : 2u;
                        const float* nsrc = reinterpret_cast<const float*>(page + ids_off + (uint64_t)page_rows * 8) + r0;
                        mbar_expect_tx(&s.mfull[ri], nb * narr);
                        for (uint32_t a = 0; a < narr; ++a)
                            tma_bulk_g2s(s.meta + (ri * 3 + a) * 128, nsrc + (size_t)a * page_rows, nb, &s.mfull[ri]);
                    }
                    if (++ri == RT_RING) {
                        ri = 0;
                        rphase ^= 1;
                    }
                    first = false;
                    const uint8_t* src = mirror + (size_t)(r0 / MIRROR_TILE_ROWS) * p.nkb * A_STAGE;
                    for (uint32_t kb = 0; kb < p.nkb; ++kb) {
                        mbar_wait(&s.empty[stage], phase ^ 1);
                        mbar_expect_tx(&s.full[stage], A_STAGE);
                        tma_bulk_g2s_hint(s.sa + stage * A_STAGE, src + (size_t)kb * A_STAGE, A_STAGE, &s.full[stage],
                                          policy);
                        if (++stage == p.S) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
            if (++pb == 2) {
                pb = 0;
                pphase ^= 1;
            }
        }
        ii = ii_next;
        it = it_next;
        pv0 = pv0_next;
    }
    mbar_wait(&s.rtempty[ri], rphase ^ 1);
    s.rt[ri * 8] = F_END;
    mbar_arrive(&s.rtfull[ri]);
}

template <bool I8>
__device__ __forceinline__ void mma_issuer(const Params& p, const Smem& s, uint32_t tmem_base) {
    constexpr uint32_t NCOL = ncol(I8), B_BLOCK = b_block(I8);
    constexpr uint32_t idesc = I8 ? idesc_i8(MIRROR_TILE_ROWS, NCOL) : idesc_bf16(MIRROR_TILE_ROWS, NCOL);
    mbar_wait(s.bfull, 0);
    tc_fence_after();
    uint32_t stage = 0, phase = 0, ri = 0, rphase = 0, buf = 0, bph = 0;
    SPROF_DECL(w_rt = 0, w_acc = 0, w_full = 0, tiles = 0);
    SPROF_T0(t_begin);
    for (;;) {
        SPROF_T0(t_a);
        mbar_wait(&s.rtfull[ri], rphase);
        SPROF_ADD(w_rt, t_a);
        const uint32_t flags = s.rt[ri * 8];
        mbar_arrive(&s.rtempty[ri]);
        if (++ri == RT_RING) {
            ri = 0;
            rphase ^= 1;
        }
        if (flags & F_END) break;
        SPROF_T0(t_b);
        mbar_wait(&s.acc_empty[buf], bph ^ 1);  // the consumers have read this accumulator
        SPROF_ADD(w_acc, t_b);
#ifdef VDB_SCREEN_PROF
        ++tiles;
#endif
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * NCOL;
        for (uint32_t kb = 0; kb < p.nkb; ++kb) {
            SPROF_T0(t_c);
            mbar_wait(&s.full[stage], phase);
            SPROF_ADD(w_full, t_c);
            tc_fence_after();
            const uint64_t da = desc_sw128(s.sa + stage * A_STAGE), db = desc_sw128(s.sb + kb * B_BLOCK);
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {  // 16 bf16 / 32 int8 = 32 bytes per MMA inside the swizzle atom
                if (I8) umma_i8(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                else umma_bf16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            }
            umma_commit(&s.empty[stage]);
            if (++stage == p.S) {
                stage = 0;
                phase ^= 1;
            }
        }
        umma_commit(&s.acc_full[buf]);
        if (++buf == 2) {
            buf = 0;
            bph ^= 1;
        }
    }
#ifdef VDB_SCREEN_PROF
    if (blockIdx.x < 3 || blockIdx.x == gridDim.x - 1)
        printf("SPROF mma  cta %3u: total %8lld clk, tiles %5lld, wait descriptor %8lld, wait accumulator %8lld, wait stage data %8lld\n",
               blockIdx.x, clock64() - t_begin, tiles, w_rt, w_acc, w_full);
#endif
}

template <int NJ, bool I8>
__device__ __forceinline__ void consumers(const Params& p, const Smem& s, uint32_t tmem_base) {
    constexpr uint32_t NCOL = ncol(I8);
    const ScanParams& sp = p.sp;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ld4 = sp.lt.ld >> 2, page_rows = sp.lt.page_rows, P = sp.P;
    const bool l2 = sp.metric == VDB_METRIC_L2;
    // the helpers of scan_kernel (compact_pool, contribute_global) work on this view of the pools
    ScanSmem ss{};
    ss.pool_i = s.pool_i;
    ss.pool_d = s.pool_d;
    ss.cnt = s.cnt;
    ss.thr = s.thr;
    ss.spair = s.spair;
    ss.sqidx = s.sqidx;
    uint32_t ri = 0, rphase = 0, buf = 0, bph = 0, qcount = 0;
    const size_t ids_off = (size_t)page_rows * sp.lt.ld * 4;  // page = rows | ids | norms | error norms | scales
    uint32_t key_seen = KEY_INF;  // the query's global bound as last read (applied one tile later)
    SPROF_DECL(c_top = 0, c_rt = 0, c_acc = 0, c_a = 0, c_b = 0, c_b2 = 0, c_last = 0, c_first = 0, n_redo = 0, n_first = 0);
    SPROF_T0(t_begin);
    for (;;) {
        SPROF_T0(t_r);
        mbar_wait(&s.rtfull[ri], rphase);
        SPROF_ADD(c_rt, t_r);
        const uint32_t* d = s.rt + ri * 8;
        const uint32_t flags = d[0], qc_new = d[1], pbi = d[2], r0 = d[4], nrows = d[5];
        if (flags & F_END) break;
        SPROF_T0(t_f);
        const uint8_t* page = reinterpret_cast<const uint8_t*>((uint64_t)d[6] | ((uint64_t)d[7] << 32));
        // this thread's row of the tile: its constants came in with the tile's announcement
        const bool valid = tid < nrows;
        mbar_wait(&s.mfull[ri], rphase);
        const float* mt = s.meta + ri * 3 * 128;
        const float vn = valid ? mt[tid] : 0.f, ev = valid ? mt[128 + tid] : 0.f, sv = (I8 && valid) ? mt[256 + tid] : 0.f;
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.rtempty[ri]);
        if (++ri == RT_RING) {
            ri = 0;
            rphase ^= 1;
        }
        if (flags & F_FIRST) {  // a new pass over an item: per-query state of its queries
            qcount = qc_new;
            if (tid < qcount) {
                const uint32_t q = s.tq[pbi * NQ + tid];
                s.cnt[tid] = 0;
                s.sqidx[tid] = q;
                s.spair[tid] = s.tslot[pbi * NQ + tid];
                s.thr[tid] = key2f(__ldcg(&sp.qthr[q]));  // what earlier items already proved
                const float4 c = __ldg(&p.qconst[2 * q]), c2 = __ldg(&p.qconst[2 * q + 1]);
                s.qc[tid] = c.x * (1.f - DOT_SLACK);
                s.uc[tid] = 2.f * c.z + 2.f * (I8 ? ACC_SLACK_I8 : ACC_SLACK) * c.y;
                s.wc[tid] = 2.f * (c.y + c.z);
                s.qsa[tid] = c2.x;
                s.qsb[tid] = c2.y;
            }
            key_seen = KEY_INF;
            bar_consumers();
            if (lane == 0) mbar_arrive(&s.pfree[pbi]);
        } else if (tid < qcount) {
            // bounds other CTAs published meanwhile, read one tile late: the load was issued while the previous tile
            // was processed, so its round trip is not on this tile's critical path.  (A racing update by the query's
            // owner warp is another valid bound.)
            s.thr[tid] = fminf(s.thr[tid], key2f(key_seen));
        }
        if (tid < qcount) key_seen = __ldcg(&sp.qthr[s.sqidx[tid]]);  // raw: first touched by the next tile
        const uint64_t* ids = reinterpret_cast<const uint64_t*>(page + ids_off);
        const float nv = __fmul_ru(__fsqrt_ru(vn), 1.00001f);
        const float sr = l2 ? vn * (1.f - DOT_SLACK) : 0.f;

        // ---- phase A: lower bounds from the tensor-core dot products
#ifdef VDB_SCREEN_PROF
        if (flags & F_FIRST) {
            c_first += clock64() - t_f;
            ++n_first;
        }
#endif
        SPROF_ADD(c_top, t_f);
        SPROF_T0(t_w);
        mbar_wait(&s.acc_full[buf], bph);
        SPROF_ADD(c_acc, t_w);
        SPROF_T0(t_pa);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((warp * 32u) << 16) + buf * NCOL;
        // q~.v~ of this thread's row and four queries: the accumulator itself (bf16), or the two integer dot products
        // of the query's terms put back on the fp32 scale (int8; exact up to three roundings, inside ACC_SLACK_I8)
        auto load_dots = [&](uint32_t j0, const uint32_t (&col)[4], float (&dot)[4]) {
            if (I8) {
                const uint32_t t8[8] = {taddr + col[0], taddr + NQ + col[0], taddr + col[1], taddr + NQ + col[1],
                                        taddr + col[2], taddr + NQ + col[2], taddr + col[3], taddr + NQ + col[3]};
                int acc[8];
                tmem_ld8(t8, acc);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    dot[u] = sv * fmaf(s.qsa[(j0 + u) & (NQ - 1)], (float)acc[2 * u], s.qsb[(j0 + u) & (NQ - 1)] * (float)acc[2 * u + 1]);
            } else {
                tmem_ld4(taddr + col[0], taddr + col[1], taddr + col[2], taddr + col[3], dot);
            }
        };
        // A query that has no bound yet (a "cold" query: the first rows of its first items) would admit every row of
        // the tile, and every admitted pair costs an exact re-score.  Round 0 therefore admits, per warp, only the
        // rows with the smallest lower bounds for it; phase B turns those into a real bound; round 1 re-reads the
        // same accumulator and admits what still passes among the rows left out.  Any choice of the round-0 subset
        // is sound -- every row is tested against a bound that holds before it is dropped.
        const uint32_t cold_rows = sp.k <= 16 ? (sp.k + 1) / 2 + 1 : 32u;  // per warp: ~2k per tile
        uint64_t coldq = 0, coldtaken = 0;  // queries this warp treated as cold / rows (this thread's) taken for them
        bool want = false;                  // this thread's row was admitted for some query
        for (uint32_t j0 = 0; j0 < qcount; j0 += 4) {
            uint32_t col[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) col[u] = j0 + u < qcount ? s.sqidx[j0 + u] : 0u;
            float dot[4];
            load_dots(j0, col, dot);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = j0 + u;
                if (j < qcount) {  // warp-uniform
                    const float slack = fmaf(nv, s.uc[j], ev * s.wc[j]);
                    // L2: (|q|^2 + |v|^2)(1 - DOT_SLACK) - 2 q~.v~ - 2 E;  inner product: -q~.v~ - E
                    const float lb = l2 ? fmaf(-2.f, dot[u], sr + s.qc[j]) - slack : fmaf(-0.5f, slack, -dot[u]);
                    const float thr = s.thr[j];
                    bool pass = valid && lb <= thr;
                    if (thr == INFINITY && cold_rows < 32u) {  // warp-uniform: thr comes from shared memory
                        const uint32_t key = valid && !(lb != lb) ? f2key(lb) : 0xffffffffu;
                        bool taken = false;
                        for (uint32_t t = 0; t < cold_rows; ++t) {
                            const uint32_t mn = __reduce_min_sync(0xffffffffu, taken ? 0xffffffffu : key);
                            if (mn == 0xffffffffu) break;
                            taken |= key == mn;
                        }
                        pass = valid && taken;
                        coldq |= 1ull << j;
                        if (taken) coldtaken |= 1ull << j;
                    }
                    const uint32_t m = __ballot_sync(0xffffffffu, pass);
                    if (lane == 0) s.adm[j * 4 + warp] = m;
                    want |= pass;
                }
            }
        }
        // the fp32 copy of an admitted row is on its way to L2 while the other queries are still being tested
        if (want)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(page + (size_t)(r0 + tid) * sp.lt.ld * 4),
                         "r"(sp.lt.ld * 4)
                         : "memory");
        if (coldq != 0 && lane == 0) *s.redo = 1u;
        bar_consumers();
        SPROF_ADD(c_a, t_pa);
        const bool redo = *s.redo != 0u;  // some warp left rows out: the accumulator is needed again
        if (!redo) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.acc_empty[buf]);
        }

        const float4* g4 = reinterpret_cast<const float4*>(page);
        // ---- phase B: exact distances of the admitted pairs, by the warp that owns the query's pool
        auto phase_b = [&]() {
            for (uint32_t j = warp; j < qcount; j += 4) {
                const uint4 am = *reinterpret_cast<const uint4*>(&s.adm[j * 4]);
                if ((am.x | am.y | am.z | am.w) == 0u) continue;
                float4 q[NJ];
                {
                    const float4* q4 = reinterpret_cast<const float4*>(sp.queries) + (size_t)s.sqidx[j] * ld4;
#pragma unroll
                    for (int jj = 0; jj < NJ; ++jj) q[jj] = __ldg(&q4[lane + 32 * jj]);
                }
                float* pd = s.pool_d + (size_t)j * P;
                uint64_t* pi = s.pool_i + (size_t)j * P;
                auto push = [&](float e, uint32_t r) {
                    if (e <= s.thr[j]) {  // identical in every lane
                        const uint32_t c = s.cnt[j];
                        __syncwarp();
                        if (lane == 0) {
                            pd[c] = e;
                            pi[c] = __ldg(&ids[r]);
                            s.cnt[j] = c + 1;
                        }
                        __syncwarp();
                        if (c + 1 == P) compact_pool(ss, sp, j, lane);  // keeps the best k, tightens thr[j]
                    }
                };
                constexpr int RB = 2;  // admitted rows re-scored at a time: their row loads overlap (4 measured slower)
                unsigned long long mlo = (unsigned long long)am.x | ((unsigned long long)am.y << 32);
                unsigned long long mhi = (unsigned long long)am.z | ((unsigned long long)am.w << 32);
                uint32_t nres = 0;
                for (;;) {
                    uint32_t rows[RB];
                    int n = 0;
#pragma unroll
                    for (int t = 0; t < RB; ++t) {
                        if (mlo) {
                            rows[t] = r0 + (uint32_t)__ffsll((long long)mlo) - 1u;  // page row
                            mlo &= mlo - 1ull;
                            ++n;
                        } else if (mhi) {
                            rows[t] = r0 + 64u + (uint32_t)__ffsll((long long)mhi) - 1u;
                            mhi &= mhi - 1ull;
                            ++n;
                        } else {
                            rows[t] = t ? rows[0] : r0;
                        }
                    }
                    if (n == 0) break;
                    nres += (uint32_t)n;
                    float e[RB];
#pragma unroll
                    for (int t = 0; t < RB; ++t)
                        e[t] = l2 ? exact_l2_lane<NJ, true>(g4, ld4, rows[t], page_rows, lane, q)
                                  : exact_ip_lane<NJ>(g4, ld4, rows[t], lane, q);
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1)
#pragma unroll
                        for (int t = 0; t < RB; ++t) e[t] += __shfl_xor_sync(0xffffffffu, e[t], step);
#pragma unroll
                    for (int t = 0; t < RB; ++t)
                        if (t < n) push(l2 ? e[t] : -e[t], rows[t]);  // IP distance = -dot, kernels.cuh:59
                }
                if (lane == 0 && nres) atomicAdd(p.rescored, (unsigned long long)nres);
            }
        };
        SPROF_T0(t_pb);
        phase_b();
        SPROF_ADD(c_b, t_pb);
        SPROF_T0(t_pr);
#ifdef VDB_SCREEN_PROF
        n_redo += redo;
#endif
        if (redo) {
            // a bound for the cold queries from what round 0 found (a pool compacts by itself only when it is full)
            for (uint32_t j = warp; j < qcount; j += 4)
                if (s.thr[j] == INFINITY && s.cnt[j] >= sp.k) compact_pool(ss, sp, j, lane);
            bar_consumers();  // every pool has taken its round-0 rows: thr is what round 1 tests against
            if (tid == 0) *s.redo = 0u;
            for (uint32_t j0 = 0; j0 < qcount; j0 += 4) {
                uint32_t col[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) col[u] = j0 + u < qcount ? s.sqidx[j0 + u] : 0u;
                float dot[4] = {0.f, 0.f, 0.f, 0.f};
                if ((coldq >> j0) & 0xfull)  // warp-uniform
                    load_dots(j0, col, dot);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t j = j0 + u;
                    if (j < qcount) {
                        uint32_t m = 0;
                        if ((coldq >> j) & 1ull) {
                            const float slack = fmaf(nv, s.uc[j], ev * s.wc[j]);
                            const float lb = l2 ? fmaf(-2.f, dot[u], sr + s.qc[j]) - slack : fmaf(-0.5f, slack, -dot[u]);
                            m = __ballot_sync(0xffffffffu, valid && !((coldtaken >> j) & 1ull) && lb <= s.thr[j]);
                        }
                        if (lane == 0) s.adm[j * 4 + warp] = m;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.acc_empty[buf]);
            bar_consumers();
            phase_b();
        }
        SPROF_ADD(c_b2, t_pr);
        SPROF_T0(t_l);
        if (++buf == 2) {
            buf = 0;
            bph ^= 1;
        }

        if (flags & F_LAST) {  // item done: best k of every query of the pass -> its partial-result slot
            for (uint32_t j = warp; j < qcount; j += 4) {
                compact_pool(ss, sp, j, lane);
                const float bound = s.thr[j];
                const uint32_t kept = s.cnt[j];
                uint32_t nc = 0;
                for (uint32_t i0 = 0; i0 < kept; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    const bool ok = i < kept && s.pool_d[(size_t)j * P + i] <= bound;
                    nc += __popc(__ballot_sync(0xffffffffu, ok));  // sorted ascending: survivors are a prefix
                }
                const size_t slot = s.spair[j];
                for (uint32_t i = lane; i < nc; i += 32) {
                    sp.part_d[slot * sp.k + i] = s.pool_d[(size_t)j * P + i];
                    sp.part_i[slot * sp.k + i] = s.pool_i[(size_t)j * P + i];
                }
                if (lane == 0) sp.part_cnt[slot] = nc;
                if (nc > 0) contribute_global(ss, sp, j, nc, lane);
            }
        }
        bar_consumers();  // adm and the per-query state are rewritten by the next tile
        SPROF_ADD(c_last, t_l);
    }
#ifdef VDB_SCREEN_PROF
    if (tid == 0 && (blockIdx.x < 3 || blockIdx.x == gridDim.x - 1))
        printf("SPROF cons cta %3u: total %8lld clk, tile top %8lld, wait descriptor %8lld, item set-up %8lld (%lld items), wait accumulator %8lld, phase A %8lld, "
               "phase B %8lld, redo rounds %8lld (%lld tiles), item end + barrier %8lld\n",
               blockIdx.x, clock64() - t_begin, c_top, c_rt, c_first, n_first, c_acc, c_a, c_b, c_b2, n_redo, c_last);
#endif
}

template <int NJ, bool I8>
__global__ void __launch_bounds__(THREADS, 1) screen_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t screen_smem[];
    const Smem s = carve(screen_smem, p);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < 8; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        for (uint32_t i = 0; i < RT_RING; ++i) {
            mbar_init(&s.mfull[i], 1);
            mbar_init(&s.rtfull[i], 1);
            mbar_init(&s.rtempty[i], 5);  // four consumer warps + the MMA issuer
        }
        for (uint32_t i = 0; i < 2; ++i) {
            mbar_init(&s.pfree[i], 4);
            mbar_init(&s.acc_full[i], 1);
            mbar_init(&s.acc_empty[i], 4);
        }
        mbar_init(s.bfull, 1);
        *s.redo = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 5) {  // two accumulators of 64 (bf16) / 128 (int8) 32-bit columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s.tmem_slot)),
                     "n"(2 * ncol(I8))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s.tmem_slot;
    if (warp < 4) {
        consumers<NJ, I8>(p, s, tmem_base);
    } else if (warp == 4) {
        if (lane == 0) producer<I8>(p, s);
    } else if (lane == 0) {
        mma_issuer<I8>(p, s, tmem_base);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * ncol(I8)) : "memory");
    }
}

// Image + constants of the batch's queries: one warp per query slot (64 slots, zero rows beyond nq).
// bf16: the query rounded to bf16 (row `slot` of a 64-row operand).  int8: TWO terms, q ~ sa * a + sb * b with a =
// rint(q / sa), b = rint((q - sa a) / sb) (rows `slot` and 64 + `slot` of a 128-row operand), so that the query's own
// quantisation error is second order (~1e-4 |q|) and the bound is spent on the rows' error alone.
__global__ void __launch_bounds__(256) query_image_kernel(const float* __restrict__ queries, uint32_t nq, uint32_t ld,
                                                          uint32_t kind, uint8_t* __restrict__ qimg,
                                                          float4* __restrict__ qconst) {
    const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (slot >= (uint32_t)NQ) return;
    const float4* src = reinterpret_cast<const float4*>(queries + (size_t)(slot < nq ? slot : 0) * ld);
    const float live = slot < nq ? 1.f : 0.f;
    float nrm = 0.f, err = 0.f, sa = 0.f, sb = 0.f;
    if (kind == MIRROR_BF16) {
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            float4 t = src[c];
            t.x *= live; t.y *= live; t.z *= live; t.w *= live;
            const __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
            uint2 bits;
            bits.x = *reinterpret_cast<const uint32_t*>(&lo);
            bits.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(qimg + mirror_elem_off(slot, c * 4u, ld, 2, NQ)) = bits;
            nrm = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(t.z, t.z, fmaf(t.w, t.w, nrm))));
            const float dx = t.x - __low2float(lo), dy = t.y - __high2float(lo);
            const float dz = t.z - __low2float(hi), dw = t.w - __high2float(hi);
            err = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, err))));
        }
    } else {
        auto q8 = [](float v, float inv) { return fminf(fmaxf(rintf(v * inv), -127.f), 127.f); };
        auto pack = [](float a, float b, float c, float d) {
            return ((uint32_t)(int)a & 0xffu) | (((uint32_t)(int)b & 0xffu) << 8) | (((uint32_t)(int)c & 0xffu) << 16) |
                   (((uint32_t)(int)d & 0xffu) << 24);
        };
        float mx = 0.f;
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            const float4 t = src[c];
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(t.x), fabsf(t.y))), fmaxf(fabsf(t.z), fabsf(t.w)));
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sa = mx * live * (1.f / 127.f);
        const float ia = sa > 0.f ? 1.f / sa : 0.f;
        float mr = 0.f;  // largest residual of the first term
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            const float4 t = src[c];
            const float rx = fmaf(-sa, q8(t.x, ia), t.x), ry = fmaf(-sa, q8(t.y, ia), t.y);
            const float rz = fmaf(-sa, q8(t.z, ia), t.z), rw = fmaf(-sa, q8(t.w, ia), t.w);
            mr = fmaxf(fmaxf(mr, fmaxf(fabsf(rx), fabsf(ry))), fmaxf(fabsf(rz), fabsf(rw)));
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
        sb = mr * live * (1.f / 127.f);
        const float ib = sb > 0.f ? 1.f / sb : 0.f;
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            float4 t = src[c];
            t.x *= live; t.y *= live; t.z *= live; t.w *= live;
            const float ax = q8(t.x, ia), ay = q8(t.y, ia), az = q8(t.z, ia), aw = q8(t.w, ia);
            const float rx = fmaf(-sa, ax, t.x), ry = fmaf(-sa, ay, t.y), rz = fmaf(-sa, az, t.z), rw = fmaf(-sa, aw, t.w);
            const float bx = q8(rx, ib), by = q8(ry, ib), bz = q8(rz, ib), bw = q8(rw, ib);
            *reinterpret_cast<uint32_t*>(qimg + mirror_elem_off(slot, c * 4u, ld, 1, 2 * NQ)) = pack(ax, ay, az, aw);
            *reinterpret_cast<uint32_t*>(qimg + mirror_elem_off(NQ + slot, c * 4u, ld, 1, 2 * NQ)) = pack(bx, by, bz, bw);
            nrm = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(t.z, t.z, fmaf(t.w, t.w, nrm))));
            // what the two terms leave: each residual carries the rounding of the one before it (relative 2^-24 of a
            // value <= sa / 2), far inside the factor below
            const float dx = fmaf(-sb, bx, rx), dy = fmaf(-sb, by, ry), dz = fmaf(-sb, bz, rz), dw = fmaf(-sb, bw, rw);
            err = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, err))));
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        err += __shfl_xor_sync(0xffffffffu, err, o);
    }
    if (lane == 0) {
        // (int8) the residual chain's own roundings: |fl(r) - r| <= 2^-24 |r| per element, |r| <= sa / 2
        const float extra = kind == MIRROR_I8 ? sa * 1e-6f * sqrtf((float)ld) : 0.f;
        qconst[2 * slot] = make_float4(nrm, __fmul_ru(__fsqrt_ru(nrm), 1.0002f),
                                       __fmul_ru(__fsqrt_ru(err), 1.0002f) + extra, 0.f);
        qconst[2 * slot + 1] = make_float4(sa, sb, 0.f, 0.f);
    }
}

}  // namespace screen
