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
//      (at most 4) centroids whose LOWER bound does not exceed it.
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

constexpr int NCAND = 4;

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
// NCAND) centroids whose LOWER bound does not exceed it.
constexpr int ASSIGN_AN = 256;  // centroid columns per accumulator tile

struct AssignEpi {
    const float* xnorm;   // [M] |x_v|
    const float* cnorm2;  // [N] |c|^2
    uint32_t M, N, num_kb, n_split;
    int metric;
    uint32_t* cand_idx;   // [M][NCAND]
    uint32_t* cand_cnt;   // [M]  (NCAND + 1 = overflow: fall back to the exact kernel for this row)

    struct State {
        float xe, U;
        float clb[NCAND];
        uint32_t cix[NCAND];
        uint32_t cnt;
        bool overflow;
    };
    static constexpr uint32_t WARP_SMEM = 0;
    __device__ __forceinline__ void chunk_end(State&, uint8_t*) const {}
    __device__ __forceinline__ void begin(State& s, uint32_t row, uint8_t*) const {
        s.xe = 1.05f * 0.00390625f * xnorm[row];
        s.U = INFINITY;
        s.cnt = 0;
        s.overflow = false;
#pragma unroll
        for (int i = 0; i < NCAND; ++i) {
            s.clb[i] = INFINITY;
            s.cix[i] = 0;
        }
    }
    __device__ __forceinline__ void consume_chunk(State& s, uint32_t row, uint32_t n0, const uint32_t (&acc)[32]) const {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (n0 + i < N) consume(s, row, n0 + i, __uint_as_float(acc[i]));
    }
    __device__ __forceinline__ void consume(State& s, uint32_t, uint32_t n, float dot) const {
        const float cn2 = __ldg(cnorm2 + n);
        const float sc = (metric == VDB_METRIC_L2) ? fmaf(-2.f, dot, cn2) : -dot;
        const float e = s.xe * sqrtf(cn2) + 1e-6f * fabsf(sc) + 1e-30f;
        const float ub = sc + e, lb = sc - e;
        s.U = fminf(s.U, ub);
        if (lb <= s.U) {
            // keep it; first drop survivors the tighter bound has ruled out meanwhile
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < NCAND; ++j)
                if ((uint32_t)j < s.cnt && s.clb[j] <= s.U) {
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
                s.overflow = true;
            }
        }
    }
    __device__ __forceinline__ void end(State& s, uint32_t row, bool valid, uint8_t*) const {
        if (!valid) return;
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < NCAND; ++j)
            if ((uint32_t)j < s.cnt && s.clb[j] <= s.U) cand_idx[(size_t)row * NCAND + w++] = s.cix[j];
        cand_cnt[row] = s.overflow ? NCAND + 1 : w;
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
        cap_nc = nc;
        VDB_CUDA_TRY(cudaMalloc(&cnorm2, (size_t)cap_nc * 4));
    }
    if (!overflow_count) {
        VDB_CUDA_TRY(cudaMalloc(&overflow_count, 4));
        VDB_CUDA_TRY(cudaMallocHost(&h_overflow, 4));
    }
    return VDB_OK;
}

void AssignTcScratch::release() {
    cudaFree(xnorm); cudaFree(cand_idx); cudaFree(cand_cnt); cudaFree(overflow_rows); cudaFree(cnorm2);
    cudaFree(overflow_count);
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
    row_norms_kernel<<<(uint32_t)(((uint64_t)nc * 32 + 255) / 256), 256, 0, stream>>>(c, nc, ldc, sc.cnorm2, false);
    CUtensorMap mx, mc;
    VDB_TRY(tc::make_map(&mx, x, n, ldx, ldx, AM));
    VDB_TRY(tc::make_map(&mc, c, nc, ldc, ldc, ASSIGN_AN));
    AssignEpi p;
    p.xnorm = sc.xnorm; p.cnorm2 = sc.cnorm2;
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
                                         stream));
    return VDB_OK;
}

}  // namespace vdb
