// Nearest-centroid assignment on the tensor cores, exact by construction.
//
// Replaces the O(n * nlist * dim) scalar loop of assign_to_lists
// (ivf_flat_index.cpp:259-295) / kmeans_assign_kernel (kernels.cuh:315-354) for
// large centroid tables, in three steps:
//   1. row_norms_kernel      |x_v| of every row
//   2. rowtile_gemm_kernel<AssignEpi, ASSIGN_AN> (rowtile_gemm.cuh)  persistent tcgen05 GEMM: a CTA owns 128 rows and
//      walks ALL centroid tiles (128 columns each); operands arrive by 2-D TMA
//      through a 4-stage ring, the TF32 MMAs accumulate into one of two TMEM
//      buffers while the four epilogue warps drain the other (thread = row).
//      The epilogue never materialises the n x nlist score matrix: per row it
//      keeps the smallest UPPER bound of the true score seen so far and the
//      (at most 8) centroids whose LOWER bound does not exceed it.
//   3. assign_recheck_kernel the survivors (1-2 per row in practice) are scored
//      exactly like the reference -- fp32, ascending dimension, unfused multiply
//      and add, strict '<' in ascending centroid order -- so the result is
//      bit-identical to assign_exact_kernel; rows whose candidate list
//      overflowed fall back to that kernel.
// Bound: as in coarse.cu, |true - tf32 score| <= E = 2^-8 |x||c| (+5 % and a
// relative 1e-6 for the fp32 rounding of the bound arithmetic itself).
#include "kmeans.cuh"
#include "rowtile_gemm.cuh"

namespace vdb {
namespace {

using namespace tc;

constexpr int NCAND = 8;

__global__ void row_norms_kernel(const float* __restrict__ x, uint64_t n, uint32_t ld, float* __restrict__ out,
                                 bool take_sqrt) {
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (w >= n) return;
    float s = 0.f;
    for (uint32_t d = lane; d < ld; d += 32) {
        const float v = x[w * ld + d];
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[w] = take_sqrt ? sqrtf(s) : s;
}

// Epilogue of the assignment: per row, the smallest UPPER bound of a true score seen so far and the (at most
// NCAND) centroids whose LOWER bound does not exceed it.  With score = |c|^2 - 2 x.c (L2, the common |x|^2 left
// out) or -x.c (IP), and eps = (dim + 16) 2^-24 covering the fp32 rounding of this arithmetic and of the exact
// kernel's own summation:
//     t  = fma(dot, alpha, A_c)            A_c = |c|^2 (L2) / 0 (IP), alpha = -2 / -1
//     e  = fma(|x| c', |c|, E_c)           c' = 1.05 * 2^-8 + 2 eps,  E_c = eps |c|^2 (L2) / 0 (IP)
//     ub = t + e,  lb = t - e              a centroid survives while lb <= min ub + 2 eps |x|^2
// A_c, |c| and E_c come from column tables padded to the tile width (A = +inf there: never a survivor).
constexpr int ASSIGN_AN = 256;  // centroid columns per accumulator tile

__global__ void assign_columns_kernel(const float* __restrict__ c, uint32_t nc, uint32_t npad, uint32_t ld, int metric,
                                      float eps, float* __restrict__ colA, float* __restrict__ colB,
                                      float* __restrict__ colE) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= npad) return;
    if (w >= nc) {
        if (lane == 0) {
            colA[w] = INFINITY;
            colB[w] = 0.f;
            colE[w] = 0.f;
        }
        return;
    }
    float s = 0.f;
    for (uint32_t d = lane; d < ld; d += 32) {
        const float v = c[(size_t)w * ld + d];
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const bool l2 = metric == VDB_METRIC_L2;
        colA[w] = l2 ? s : 0.f;
        colB[w] = sqrtf(s) * (1.f + 1e-6f);
        colE[w] = l2 ? eps * s : 0.f;
    }
}

struct AssignEpi {
    const float* xnorm;  // [M] |x_v|
    const float* colA;   // [N padded to ASSIGN_AN]
    const float* colB;
    const float* colE;
    uint32_t M, N, num_kb, n_split;
    int metric;
    float eps;
    uint32_t* cand_idx;  // [M][NCAND]
    uint32_t* cand_cnt;  // [M]  (NCAND + 1 = overflow: fall back to the exact kernel for this row)

    static constexpr uint32_t WARP_SMEM = 0;
    struct State {
        float xe, U, slack, alpha;  // U = smallest ub so far; survivors need lb <= U + slack
        float clb[NCAND];
        uint32_t cix[NCAND];
        uint32_t cnt;
        float dropped;  // smallest lower bound among survivors that did not fit; harmless if it ends above the limit
    };
    __device__ __forceinline__ void begin(State& s, uint32_t row, uint8_t*) const {
        const float xn = xnorm[row];
        const bool l2 = metric == VDB_METRIC_L2;
        s.xe = xn * (1.05f * 0.00390625f + (l2 ? 2.f * eps : eps));
        s.slack = l2 ? 2.f * eps * xn * xn : 0.f;
        s.alpha = l2 ? -2.f : -1.f;
        s.U = INFINITY;
        s.cnt = 0;
        s.dropped = INFINITY;
#pragma unroll
        for (int i = 0; i < NCAND; ++i) {
            s.clb[i] = INFINITY;
            s.cix[i] = 0;
        }
    }
    // keep centroid n (lower bound lb); first drop survivors the tighter bound has ruled out meanwhile
    __device__ __forceinline__ void keep(State& s, uint32_t n, float lb) const {
        const float lim = s.U + s.slack;
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < NCAND; ++j)
            if ((uint32_t)j < s.cnt && s.clb[j] <= lim) {
                const float tl = s.clb[j];
                const uint32_t ti = s.cix[j];
#pragma unroll
                for (int t = 0; t < NCAND; ++t)
                    if ((uint32_t)t == w) {
                        s.clb[t] = tl;
                        s.cix[t] = ti;
                    }
                ++w;
            }
        s.cnt = w;
        if (s.cnt < NCAND) {
#pragma unroll
            for (int t = 0; t < NCAND; ++t)
                if ((uint32_t)t == s.cnt) {
                    s.clb[t] = lb;
                    s.cix[t] = n;
                }
            ++s.cnt;
        } else {
            // full: the entry with the largest lower bound (the new one included) makes room
            float worst = lb;
            int wi = -1;
#pragma unroll
            for (int t = 0; t < NCAND; ++t)
                if (s.clb[t] > worst) {
                    worst = s.clb[t];
                    wi = t;
                }
            s.dropped = fminf(s.dropped, worst);
#pragma unroll
            for (int t = 0; t < NCAND; ++t)
                if (t == wi) {
                    s.clb[t] = lb;
                    s.cix[t] = n;
                }
        }
    }
    __device__ __forceinline__ void consume_chunk(State& s, uint32_t, uint32_t n0, const uint32_t (&acc)[32]) const {
        const float4* A4 = reinterpret_cast<const float4*>(colA + n0);
        const float4* B4 = reinterpret_cast<const float4*>(colB + n0);
        const float4* E4 = reinterpret_cast<const float4*>(colE + n0);
        float lb[32];
        float U = s.U;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 a = __ldg(A4 + j), b = __ldg(B4 + j), ee = __ldg(E4 + j);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, ev[4] = {ee.x, ee.y, ee.z, ee.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float t = fmaf(__uint_as_float(acc[4 * j + i]), s.alpha, av[i]);
                const float e = fmaf(s.xe, bv[i], ev[i]);
                U = fminf(U, t + e);
                lb[4 * j + i] = t - e;
            }
        }
        s.U = U;
        const float lim = U + s.slack;
        uint32_t hit = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) hit |= (lb[i] <= lim ? 1u : 0u) << i;
        while (hit) {  // rare: a new running minimum, or a near tie with it
            const uint32_t i = __ffs(hit) - 1;
            hit &= hit - 1;
            float l = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if ((uint32_t)j == i) l = lb[j];
            if (n0 + i < N) keep(s, n0 + i, l);
        }
    }
    __device__ __forceinline__ void chunk_end(State&, uint8_t*) const {}
    __device__ __forceinline__ void end(State& s, uint32_t row, bool valid, uint8_t*) const {
        if (!valid) return;
        const float lim = s.U + s.slack;
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < NCAND; ++j)
            if ((uint32_t)j < s.cnt && s.clb[j] <= lim) cand_idx[(size_t)row * NCAND + w++] = s.cix[j];
        cand_cnt[row] = s.dropped <= lim ? NCAND + 1 : w;
    }
};

// exact fp32 re-scoring of the surviving centroids, in the reference's order (one thread per row; tiles are
// transposed through shared memory so the row reads stay coalesced)
__global__ void __launch_bounds__(128) assign_recheck_kernel(const float* __restrict__ x, uint64_t n, uint32_t ldx,
                                                             const float* __restrict__ c, uint32_t ldc, uint32_t dim,
                                                             int metric, const uint32_t* __restrict__ cand_idx,
                                                             const uint32_t* __restrict__ cand_cnt,
                                                             uint32_t* __restrict__ assign,
                                                             uint32_t* __restrict__ overflow_rows,
                                                             uint32_t* __restrict__ overflow_count) {
    __shared__ float tile[128][33];
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t v0 = (uint64_t)blockIdx.x * 128;
    const uint64_t v = v0 + tid;
    uint32_t cnt = v < n ? cand_cnt[v] : 0;
    const bool over = cnt > NCAND;
    if (over) cnt = 0;
    // candidates in ascending centroid order, so that strict '<' reproduces "lowest index wins ties"
    uint32_t ci[NCAND];
#pragma unroll
    for (int j = 0; j < NCAND; ++j) ci[j] = (uint32_t)j < cnt ? cand_idx[v * NCAND + j] : 0xffffffffu;
#pragma unroll
    for (int a = 0; a < NCAND; ++a)
#pragma unroll
        for (int b = 0; b + 1 < NCAND - a; ++b)
            if (ci[b] > ci[b + 1]) {
                const uint32_t t = ci[b];
                ci[b] = ci[b + 1];
                ci[b + 1] = t;
            }
    float acc[NCAND];
#pragma unroll
    for (int j = 0; j < NCAND; ++j) acc[j] = 0.f;
    for (uint32_t d0 = 0; d0 < dim; d0 += 32) {
        const bool dok = d0 + lane < dim;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            const uint64_t vv = v0 + w * 32 + i;
            tile[w * 32 + i][lane] = (vv < n && dok) ? x[vv * ldx + d0 + lane] : 0.f;
        }
        __syncthreads();
        const uint32_t dn = min(32u, dim - d0);
#pragma unroll
        for (int j = 0; j < NCAND; ++j) {
            if ((uint32_t)j >= cnt) continue;
            const float* cv = c + (size_t)ci[j] * ldc + d0;
            float a = acc[j];
            if (metric == VDB_METRIC_L2) {
                for (uint32_t dd = 0; dd < dn; ++dd) {
                    const float diff = __fsub_rn(tile[tid][dd], __ldg(cv + dd));
                    a = __fadd_rn(a, __fmul_rn(diff, diff));
                }
            } else {
                for (uint32_t dd = 0; dd < dn; ++dd) a = __fadd_rn(a, __fmul_rn(tile[tid][dd], __ldg(cv + dd)));
            }
            acc[j] = a;
        }
        __syncthreads();
    }
    if (v >= n) return;
    if (over) {
        overflow_rows[atomicAdd(overflow_count, 1u)] = (uint32_t)v;
        return;
    }
    float best = FLT_MAX;
    uint32_t bi = 0;
#pragma unroll
    for (int j = 0; j < NCAND; ++j)
        if ((uint32_t)j < cnt) {
            const float d = (metric == VDB_METRIC_L2) ? acc[j] : -acc[j];
            if (d < best) {
                best = d;
                bi = ci[j];
            }
        }
    assign[v] = bi;
}

}  // namespace

bool assign_tensor_supported(uint32_t nc, uint32_t ld) {
    return tc::encode_tiled() != nullptr && nc >= 256 && ld % 4 == 0;
}

int32_t AssignTcScratch::reserve(uint64_t n, uint32_t nc) {
    if (n > cap_n) {
        cudaFree(xnorm); cudaFree(cand_idx); cudaFree(cand_cnt); cudaFree(overflow_rows);
        cap_n = n + n / 8 + 128;
        VDB_CUDA_TRY(cudaMalloc(&xnorm, cap_n * 4));
        VDB_CUDA_TRY(cudaMalloc(&cand_idx, cap_n * NCAND * 4));
        VDB_CUDA_TRY(cudaMalloc(&cand_cnt, cap_n * 4));
        VDB_CUDA_TRY(cudaMalloc(&overflow_rows, cap_n * 4));
    }
    if (nc > cap_nc) {
        cudaFree(cnorm2);
        cap_nc = round_up(nc, (uint32_t)ASSIGN_AN);
        VDB_CUDA_TRY(cudaMalloc(&cnorm2, (size_t)cap_nc * 3 * 4));  // three column tables (assign_columns_kernel)
    }
    if (!overflow_count) {
        VDB_CUDA_TRY(cudaMalloc(&overflow_count, 4));
        VDB_CUDA_TRY(cudaMalloc(&fewrow_keys, FEWROWS_MAX * 8));
        VDB_CUDA_TRY(cudaMallocHost(&h_overflow, 4));
    }
    return VDB_OK;
}

void AssignTcScratch::release() {
    cudaFree(xnorm); cudaFree(cand_idx); cudaFree(cand_cnt); cudaFree(overflow_rows); cudaFree(cnorm2);
    cudaFree(overflow_count); cudaFree(fewrow_keys);
    if (h_overflow) cudaFreeHost(h_overflow);
    *this = AssignTcScratch();
}

// x [n][ldx] and c [nc][ldc] device arrays, 16-byte aligned rows.  Synchronises the stream once (to learn
// whether any row overflowed its candidate list).
int32_t kmeans_assign_tensor(const float* x, uint64_t n, uint32_t ldx, const float* c, uint32_t nc, uint32_t ldc,
                             uint32_t dim, int metric, uint32_t* assign, AssignTcScratch& sc, cudaStream_t stream) {
    if (n == 0) return VDB_OK;
    VDB_REQUIRE(n < (1ull << 31), "assign: too many rows for one call");
    VDB_TRY(sc.reserve(n, nc));
    row_norms_kernel<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, stream>>>(x, n, ldx, sc.xnorm, true);
    const uint32_t npad = round_up(nc, (uint32_t)ASSIGN_AN);
    const float eps = std::max(1e-6f, (float)(ldx + 16) * 5.9604645e-8f);
    float *colA = sc.cnorm2, *colB = sc.cnorm2 + npad, *colE = sc.cnorm2 + 2 * (size_t)npad;
    assign_columns_kernel<<<(uint32_t)(((uint64_t)npad * 32 + 255) / 256), 256, 0, stream>>>(c, nc, npad, ldc, metric, eps,
                                                                                          colA, colB, colE);
    CUtensorMap mx, mc;
    VDB_TRY(tc::make_map(&mx, x, n, ldx, ldx, AM));
    VDB_TRY(tc::make_map(&mc, c, nc, ldc, ldc, ASSIGN_AN));
    AssignEpi p;
    p.xnorm = sc.xnorm; p.colA = colA; p.colB = colB; p.colE = colE; p.eps = eps;
    p.M = (uint32_t)n; p.N = nc; p.num_kb = (ldx + GK - 1) / GK; p.n_split = 1; p.metric = metric;
    p.cand_idx = sc.cand_idx; p.cand_cnt = sc.cand_cnt;
    constexpr uint32_t smem = rowtile_smem<AssignEpi, ASSIGN_AN>();
    static bool conf[8] = {false};
    int dev = 0, sms = NUM_SMS_B200;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev < 8 && !conf[dev]) {
        VDB_CUDA_TRY(cudaFuncSetAttribute(rowtile_gemm_kernel<AssignEpi, ASSIGN_AN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf[dev] = true;
    }
    const uint32_t m_tiles = (uint32_t)((n + AM - 1) / AM);
    rowtile_gemm_kernel<AssignEpi, ASSIGN_AN><<<std::min<uint32_t>(m_tiles, (uint32_t)sms), ATHREADS, smem, stream>>>(mx, mc, p);
    VDB_CUDA_TRY(cudaGetLastError());
    VDB_CUDA_TRY(cudaMemsetAsync(sc.overflow_count, 0, 4, stream));
    assign_recheck_kernel<<<(uint32_t)((n + 127) / 128), 128, 0, stream>>>(x, n, ldx, c, ldc, dim, metric, sc.cand_idx,
                                                                         sc.cand_cnt, assign, sc.overflow_rows,
                                                                         sc.overflow_count);
    VDB_CUDA_TRY(cudaGetLastError());
    VDB_CUDA_TRY(cudaMemcpyAsync(sc.h_overflow, sc.overflow_count, 4, cudaMemcpyDeviceToHost, stream));
    VDB_CUDA_TRY(cudaStreamSynchronize(stream));
    sc.last_overflow = *sc.h_overflow;
    if (sc.last_overflow)  // rows with more than NCAND survivors: the scalar kernel, on just those rows
        VDB_TRY(kmeans_assign_exact_rows(x, sc.overflow_rows, sc.last_overflow, ldx, c, nc, ldc, dim, metric, assign,
                                         sc.fewrow_keys, stream));
    return VDB_OK;
}

}  // namespace vdb
