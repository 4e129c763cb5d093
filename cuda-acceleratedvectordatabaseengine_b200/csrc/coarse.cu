// Coarse selection on the 5th-generation tensor cores (sm_100a).
//
// Replaces select_nprobe_lists (ivf_flat_index.cpp:298-336) for a whole query
// batch:
//   1. score_gemm_kernel   S = Q * C^T as a dense TF32 contraction: tcgen05.mma
//      (cta_group::1, kind::tf32, M=128 x N=64 x K=8 per instruction) issued by
//      one thread, operands staged by 2-D TMA (cp.async.bulk.tensor, 128-byte
//      swizzle) through a 4-stage mbarrier ring, the fp32 accumulator in TMEM,
//      read back with tcgen05.ld for the epilogue.  fp32 rows are fed as they
//      are (the MMA reads the top 19 bits), so there is no conversion pass.
//   2. coarse_select_kernel  per query: approximate scores |c|^2 - 2 q.c (or
//      -q.c), radix-select of the nprobe-th smallest, every centroid within the
//      TF32 rounding bound of it becomes a candidate, candidates are re-scored
//      in exact fp32, and the best nprobe by (distance, list id) are returned.
//
// Error bound: TF32 keeps 10 mantissa bits and the unit truncates, so each
// input carries a relative error < 2^-10 and |q.c - tf32(q).tf32(c)| <
// 2^-9 |q||c| (Cauchy-Schwarz; the fp32 accumulation error is three orders of
// magnitude below).  With E_n = 2^-8 |q| |c_n| (covers the factor 2 of the L2
// score, and both metrics) the true score of centroid n lies in
// [s_n - E_n, s_n + E_n].  Let U be the nprobe-th smallest upper bound
// s_n + E_n: at least nprobe centroids truly score <= U, so a centroid whose
// lower bound s_n - E_n exceeds U cannot be among the nprobe best.  Everything
// else is a candidate and is re-scored in exact fp32, which makes the result
// independent of the tensor-core rounding.
#include "coarse.cuh"
#include "tc_common.cuh"
#include "topk.cuh"

namespace vdb {
namespace {

using namespace tc;

// ------------------------------------------------------------ 1. score GEMM

constexpr int GM = 128;      // rows of A per CTA (UMMA M)
constexpr int GSTAGES = 4;
constexpr int GEMM_THREADS = 192;  // warps 0-3 epilogue (TMEM lane quadrants), 4 TMA producer, 5 MMA issuer + TMEM owner

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
score_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  float* __restrict__ out, uint32_t M, uint32_t N, uint32_t ldo, uint32_t num_kb) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 128-byte-swizzled tiles must start on a 1024-byte boundary of the shared window
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr uint32_t A_BYTES = GM * GK * 4, B_BYTES = BN * GK * 4;
    uint8_t* sa = smem;
    uint8_t* sb = smem + GSTAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sb + GSTAGES * B_BYTES);
    uint64_t* empty = full + GSTAGES;
    uint64_t* acc_full = empty + GSTAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t m0 = blockIdx.y * GM, n0 = blockIdx.x * BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < GSTAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 5) {  // one warp owns the TMEM allocation: BN fp32 accumulator columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 4 && lane == 0) {
        // TMA producer: one A box [128 rows][32 k] and one B box [BN rows][32 k] per stage
        uint32_t s = 0, ph = 0;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
            tma_load_2d(sa + s * A_BYTES, &map_a, (int32_t)(kb * GK), (int32_t)m0, &full[s]);
            tma_load_2d(sb + s * B_BYTES, &map_b, (int32_t)(kb * GK), (int32_t)n0, &full[s]);
            if (++s == GSTAGES) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 5 && lane == 0) {
        // MMA issuer: 4 instructions of K = 8 per stage, advancing 32 bytes inside the swizzle atom
        constexpr uint32_t idesc = umma_idesc_tf32(GM, BN);
        uint32_t s = 0, ph = 0;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint64_t da = umma_desc_sw128(sa + s * A_BYTES), db = umma_desc_sw128(sb + s * B_BYTES);
#pragma unroll
            for (uint32_t k = 0; k < GK / 8; ++k)
                umma_tf32(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            umma_commit(&empty[s]);  // the stage is free once these MMAs have read it
            if (++s == GSTAGES) {
                s = 0;
                ph ^= 1;
            }
        }
        umma_commit(acc_full);  // accumulator complete
    } else if (warp < 4) {
        // epilogue: warp w reads TMEM lanes [32w, 32w+32) = rows m0+32w.. of the tile
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t row = m0 + warp * 32 + lane;
#pragma unroll 1
        for (uint32_t c0 = 0; c0 < (uint32_t)BN; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tmem_d + ((warp * 32u) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                  "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                  "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < M) {
                float* o = out + (size_t)row * ldo + n0 + c0;
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (n0 + c0 + i < N) o[i] = __uint_as_float(r[i]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(BN) : "memory");
    }
}

constexpr int BN_COARSE = 64;

// ------------------------------------------------------ 2. select + re-check

constexpr int SEL_THREADS = 256;
constexpr uint32_t SEL_CAND = 2048;  // sort capacity: running best np + one index chunk of candidates

__device__ __forceinline__ uint32_t f2key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void centroid_norms_kernel(const float* __restrict__ c, uint32_t n, uint32_t ld, float* __restrict__ norms,
                                      uint32_t* __restrict__ max_bits) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    float s = 0.f;
    for (uint32_t d = lane; d < ld; d += 32) {
        const float v = c[(size_t)w * ld + d];
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        norms[w] = s;
        atomicMax(max_bits, __float_as_uint(s));  // non-negative floats order like their bit patterns
    }
}

struct SelectParams {
    const float* dots;     // [nq][ldd] q.c from the tensor cores
    const float* queries;  // [nq][ld]
    const float* centroids;// [N][ld]
    const float* cnorm;    // [N] |c|^2
    const uint32_t* cmax_bits;
    uint32_t N, ld, ldd, np;
    int metric;
    uint32_t* probes;      // [nq][np]
    float* out_d;          // [nq][np] exact coarse distances (optional)
    uint32_t* cand_count;  // [nq] candidates re-checked (diagnostic, optional)
};

__global__ void __launch_bounds__(SEL_THREADS) coarse_select_kernel(const SelectParams p) {
    extern __shared__ __align__(16) uint8_t ssm[];
    float* score = reinterpret_cast<float*>(ssm);                       // [N]
    uint64_t* cid = reinterpret_cast<uint64_t*>(score + ((p.N + 1) & ~1u));  // [SEL_CAND]
    float* cd = reinterpret_cast<float*>(cid + SEL_CAND);               // [SEL_CAND]
    __shared__ uint32_t hist[256];
    __shared__ float s_red[SEL_THREADS / 32];
    __shared__ uint32_t s_prefix, s_rank, s_ncand, s_nbest;
    __shared__ float s_qnorm;
    const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* qv = p.queries + (size_t)q * p.ld;

    // |q|
    float qq = 0.f;
    for (uint32_t d = tid; d < p.ld; d += SEL_THREADS) qq = fmaf(qv[d], qv[d], qq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    if (lane == 0) s_red[warp] = qq;
    __syncthreads();
    if (tid == 0) {
        float t = 0.f;
        for (int i = 0; i < SEL_THREADS / 32; ++i) t += s_red[i];
        s_qnorm = sqrtf(t);
        s_prefix = 0;
        s_rank = p.np - 1;  // 0-based rank of the bound we are after
    }
    __syncthreads();
    // per-centroid rounding bound of the tensor-core score (see the file header): E_n = 2^-8 |q| |c_n|, with
    // 5% + 1e-6 relative slack for the fp32 rounding of the bound and of |c|^2 - 2 q.c themselves
    const float eq = 1.05f * 0.00390625f * s_qnorm;
    auto approx = [&](uint32_t n, float& e) {
        const float dot = p.dots[(size_t)q * p.ldd + n];
        const float cn = p.cnorm[n];
        const float sc = (p.metric == VDB_METRIC_L2) ? fmaf(-2.f, dot, cn) : -dot;
        e = eq * sqrtf(cn) + 1e-6f * fabsf(sc) + 1e-30f;
        return sc;
    };
    // upper bounds of the true scores
    for (uint32_t n = tid; n < p.N; n += SEL_THREADS) {
        float e;
        const float sc = approx(n, e);
        score[n] = sc + e;
    }
    __syncthreads();

    // radix select (4 x 8 bits, most significant first) of the np-th smallest upper bound U: at least np
    // centroids truly score <= U, so the true np-th score is <= U
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (uint32_t i = tid; i < 256; i += SEL_THREADS) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (uint32_t n = tid; n < p.N; n += SEL_THREADS) {
            const uint32_t key = f2key(score[n]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (warp == 0) {
            // bin holding the wanted rank: inclusive prefix over 256 bins, 8 per lane
            uint32_t h8[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                h8[i] = hist[lane * 8 + i];
                sum += h8[i];
            }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= (uint32_t)o) incl += y;
            }
            const uint32_t rank = s_rank;
            uint32_t before = incl - sum;  // keys in the bins of lower lanes
            const bool mine = rank >= before && rank < incl;
            if (mine) {
                uint32_t b = 0, r = rank - before;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (r >= h8[i] && b == (uint32_t)i) {
                        r -= h8[i];
                        b = i + 1;
                    }
                }
                s_rank = r;
                s_prefix = prefix | ((lane * 8 + b) << shift);
            }
        }
        __syncthreads();
    }
    const float U = key2f(s_prefix);
    // a centroid whose lower bound exceeds U cannot be among the np best: lower bound = upper bound - 2 E_n
    for (uint32_t n = tid; n < p.N; n += SEL_THREADS) {
        float e;
        const float sc = approx(n, e);
        score[n] = sc - e;
    }
    const float admit = U;
    __shared__ uint32_t s_total;
    if (tid == 0) s_total = 0;
    __syncthreads();
    {
        uint32_t mine = 0;
        for (uint32_t n = tid; n < p.N; n += SEL_THREADS) mine += (score[n] <= admit);
        if (mine) atomicAdd(&s_total, mine);
    }
    __syncthreads();

    // candidates, one index chunk at a time so that a chunk plus the running best always fit the sorter
    if (tid == 0) s_nbest = 0;
    uint32_t total_cand = 0;
    // all candidates at once when they fit the sorter (the usual case), else one index chunk at a time
    const uint32_t chunk = (s_total + p.np <= SEL_CAND) ? p.N : SEL_CAND - p.np;
    for (uint32_t n0 = 0; n0 < p.N; n0 += chunk) {
        __syncthreads();
        if (tid == 0) s_ncand = s_nbest;  // entries [0, nbest) hold the best so far
        __syncthreads();
        const uint32_t n1 = min(p.N, n0 + chunk);
        for (uint32_t n = n0 + tid; n < n1; n += SEL_THREADS)
            if (score[n] <= admit) {
                const uint32_t pos = atomicAdd(&s_ncand, 1u);
                cid[pos] = n;
            }
        __syncthreads();
        const uint32_t nb = s_nbest, nc = s_ncand;
        total_cand += nc - nb;
        // exact fp32 distance of every new candidate: one warp per candidate, 128-bit loads, two candidates
        // in flight per warp so the L2 round trips overlap
        {
            const uint32_t ld4 = p.ld >> 2;
            const float4* q4 = reinterpret_cast<const float4*>(qv);
            const float4* c4base = reinterpret_cast<const float4*>(p.centroids);
            constexpr uint32_t NW = SEL_THREADS / 32;
            for (uint32_t i = nb + warp; i < nc; i += 2 * NW) {
                const uint32_t i2 = i + NW;
                const bool two = i2 < nc;
                const float4* ca = c4base + (size_t)cid[i] * ld4;
                const float4* cb = c4base + (size_t)cid[two ? i2 : i] * ld4;
                float a = 0.f, b = 0.f;
                for (uint32_t d = lane; d < ld4; d += 32) {
                    const float4 qq4 = q4[d], va = ca[d], vb = cb[d];
                    if (p.metric == VDB_METRIC_L2) {
                        float t;
                        t = qq4.x - va.x; a = fmaf(t, t, a);
                        t = qq4.y - va.y; a = fmaf(t, t, a);
                        t = qq4.z - va.z; a = fmaf(t, t, a);
                        t = qq4.w - va.w; a = fmaf(t, t, a);
                        t = qq4.x - vb.x; b = fmaf(t, t, b);
                        t = qq4.y - vb.y; b = fmaf(t, t, b);
                        t = qq4.z - vb.z; b = fmaf(t, t, b);
                        t = qq4.w - vb.w; b = fmaf(t, t, b);
                    } else {
                        a = fmaf(qq4.x, va.x, a); a = fmaf(qq4.y, va.y, a); a = fmaf(qq4.z, va.z, a); a = fmaf(qq4.w, va.w, a);
                        b = fmaf(qq4.x, vb.x, b); b = fmaf(qq4.y, vb.y, b); b = fmaf(qq4.z, vb.z, b); b = fmaf(qq4.w, vb.w, b);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    b += __shfl_xor_sync(0xffffffffu, b, o);
                }
                if (lane == 0) {
                    cd[i] = (p.metric == VDB_METRIC_L2) ? a : -a;
                    if (two) cd[i2] = (p.metric == VDB_METRIC_L2) ? b : -b;
                }
            }
        }
        __syncthreads();
        const uint32_t n2 = dev_next_pow2(max(nc, 1u));
        for (uint32_t i = nc + tid; i < n2; i += SEL_THREADS) {
            cd[i] = FLT_MAX;
            cid[i] = ID_PAD;
        }
        __syncthreads();
        bitonic_sort_pairs(cd, cid, n2, tid, SEL_THREADS, [] { __syncthreads(); });
        if (tid == 0) s_nbest = min(nc, p.np);
    }
    __syncthreads();
    const uint32_t nb = s_nbest;
    for (uint32_t i = tid; i < p.np; i += SEL_THREADS) {
        p.probes[(size_t)q * p.np + i] = i < nb ? (uint32_t)cid[i] : 0xffffffffu;
        if (p.out_d) p.out_d[(size_t)q * p.np + i] = i < nb ? cd[i] : FLT_MAX;
    }
    if (p.cand_count && tid == 0) p.cand_count[q] = total_cand;
}

// nprobe beyond the select kernel's candidate pool (> 2047, e.g. exhaustive probing of 4096 lists): every centroid is
// scored in exact fp32 (warp per centroid) and the whole table is sorted by (distance, list id) in shared memory.
constexpr int WIDE_THREADS = 512;

__global__ void __launch_bounds__(WIDE_THREADS) coarse_wide_kernel(const float* __restrict__ queries,
                                                                   const float* __restrict__ centroids, uint32_t N,
                                                                   uint32_t ld, uint32_t np, int metric,
                                                                   uint32_t* __restrict__ probes,
                                                                   float* __restrict__ out_d) {
    extern __shared__ __align__(16) uint8_t wsm[];
    const uint32_t n2 = dev_next_pow2(N);
    float* sd = reinterpret_cast<float*>(wsm);          // [n2] distances
    uint32_t* si = reinterpret_cast<uint32_t*>(sd + n2);  // [n2] list ids
    float* sq = reinterpret_cast<float*>(si + n2);        // [ld] the query
    const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, ld4 = ld >> 2;
    for (uint32_t c = tid; c < ld4; c += WIDE_THREADS)
        reinterpret_cast<float4*>(sq)[c] = reinterpret_cast<const float4*>(queries + (size_t)q * ld)[c];
    for (uint32_t i = N + tid; i < n2; i += WIDE_THREADS) {
        sd[i] = INFINITY;
        si[i] = 0xffffffffu;
    }
    __syncthreads();
    const float4* q4 = reinterpret_cast<const float4*>(sq);
    for (uint32_t c = warp; c < N; c += WIDE_THREADS / 32) {
        const float4* c4 = reinterpret_cast<const float4*>(centroids + (size_t)c * ld);
        float a = 0.f;
        for (uint32_t j = lane; j < ld4; j += 32) {
            const float4 x = c4[j], y = q4[j];
            if (metric == VDB_METRIC_L2) {
                float u;
                u = y.x - x.x; a = fmaf(u, u, a);
                u = y.y - x.y; a = fmaf(u, u, a);
                u = y.z - x.z; a = fmaf(u, u, a);
                u = y.w - x.w; a = fmaf(u, u, a);
            } else {
                a = fmaf(y.x, x.x, a); a = fmaf(y.y, x.y, a); a = fmaf(y.z, x.z, a); a = fmaf(y.w, x.w, a);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            sd[c] = metric == VDB_METRIC_L2 ? a : -a;
            si[c] = c;
        }
    }
    __syncthreads();
    for (uint32_t size = 2; size <= n2; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (n2 >> 1); t += WIDE_THREADS) {
                const uint32_t i = ((t / stride) * (stride << 1)) + (t % stride), j = i + stride;
                const float di = sd[i], dj = sd[j];
                const uint32_t ii = si[i], ij = si[j];
                const bool i_first = di < dj || (di == dj && ii < ij);
                if (((i & size) == 0) ? !i_first : i_first) {
                    sd[i] = dj; sd[j] = di;
                    si[i] = ij; si[j] = ii;
                }
            }
            __syncthreads();
        }
    for (uint32_t i = tid; i < np; i += WIDE_THREADS) {
        probes[(size_t)q * np + i] = si[i];
        if (out_d) out_d[(size_t)q * np + i] = sd[i];
    }
}

}  // namespace

bool coarse_wide_supported(uint32_t N, uint32_t ld) {
    return (size_t)next_pow2(N) * 8 + (size_t)ld * 4 <= 200 * 1024 && ld % 4 == 0;
}

int32_t coarse_select_wide(const float* queries, uint32_t nq, const float* centroids, uint32_t N, uint32_t ld,
                           uint32_t np, int metric, uint32_t* probes, float* out_d, cudaStream_t stream) {
    static bool conf[8] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 8 && !conf[dev]) {
        VDB_CUDA_TRY(cudaFuncSetAttribute(coarse_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        conf[dev] = true;
    }
    coarse_wide_kernel<<<nq, WIDE_THREADS, (size_t)next_pow2(N) * 8 + (size_t)ld * 4, stream>>>(
        queries, centroids, N, ld, np, metric, probes, out_d);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

size_t coarse_select_smem(uint32_t N) { return (size_t)((N + 1) & ~1u) * 4 + (size_t)SEL_CAND * 12; }

bool coarse_tensor_supported(uint32_t N, uint32_t ld, uint32_t np) {
    return encode_tiled() != nullptr && coarse_select_smem(N) <= 200 * 1024 && np < SEL_CAND / 2 && ld % 4 == 0;
}

int32_t centroid_norms(const float* centroids, uint32_t n, uint32_t ld, float* norms, uint32_t* max_bits,
                       cudaStream_t stream) {
    VDB_CUDA_TRY(cudaMemsetAsync(max_bits, 0, 4, stream));
    centroid_norms_kernel<<<(n * 32 + 255) / 256, 256, 0, stream>>>(centroids, n, ld, norms, max_bits);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t score_gemm(const float* A, uint32_t M, uint32_t lda, const float* B, uint32_t N, uint32_t ldb, uint32_t K,
                   float* out, uint32_t ldo, cudaStream_t stream) {
    CUtensorMap ma, mb;
    VDB_TRY(make_map(&ma, A, M, K, lda, GM));
    VDB_TRY(make_map(&mb, B, N, K, ldb, BN_COARSE));
    constexpr uint32_t smem = GSTAGES * (GM * GK * 4 + BN_COARSE * GK * 4) + (2 * GSTAGES + 1) * 8 + 16 + 1024;
    static bool conf[8] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 8 && !conf[dev]) {
        VDB_CUDA_TRY(cudaFuncSetAttribute(score_gemm_kernel<BN_COARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        VDB_CUDA_TRY(cudaFuncSetAttribute(score_gemm_kernel<BN_COARSE>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        conf[dev] = true;
    }
    dim3 grid((N + BN_COARSE - 1) / BN_COARSE, (M + GM - 1) / GM);
    score_gemm_kernel<BN_COARSE><<<grid, GEMM_THREADS, smem, stream>>>(ma, mb, out, M, N, ldo, (K + GK - 1) / GK);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t coarse_select(const float* dots, uint32_t ldd, const float* queries, uint32_t nq, const float* centroids,
                      const float* cnorm, const uint32_t* cmax_bits, uint32_t N, uint32_t ld, uint32_t np, int metric,
                      uint32_t* probes, float* out_d, uint32_t* cand_count, cudaStream_t stream) {
    SelectParams p;
    p.dots = dots; p.queries = queries; p.centroids = centroids; p.cnorm = cnorm; p.cmax_bits = cmax_bits;
    p.N = N; p.ld = ld; p.ldd = ldd; p.np = np; p.metric = metric;
    p.probes = probes; p.out_d = out_d; p.cand_count = cand_count;
    const size_t smem = coarse_select_smem(N);
    static bool conf[8] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 8 && !conf[dev]) {
        VDB_CUDA_TRY(cudaFuncSetAttribute(coarse_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VDB_CUDA_TRY(cudaFuncSetAttribute(coarse_select_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
        conf[dev] = true;
    }
    coarse_select_kernel<<<nq, SEL_THREADS, smem, stream>>>(p);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

}  // namespace vdb
