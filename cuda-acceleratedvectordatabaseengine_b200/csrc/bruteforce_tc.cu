// Exact brute-force top-k on the tensor cores (BASELINE.json configs[1]: 1M x 768-D, 1024 queries, k = 100).
//
// Replaces launch_bruteforce_search (kernels.cu:13-43 / bruteforce_search_kernel kernels.cuh:84-185) for
// large inputs.  The scan kernel (scan.cu) computes every (row, query) distance in exact fp32 on the CUDA
// cores; here the n x nq contraction runs on tcgen05 and only a few candidates are re-scored exactly:
//   1. nested strided samples: level i holds every (8^(L-i))-th row, level L all of them.  Level 0 (a few hundred
//      rows) is searched exhaustively; its exact k-th distance T_q bounds the k-th distance of every larger level
//      from above, and in expectation only 8k rows of the next level fall below it;
//   2. per level, rowtile_gemm_kernel<BruteEpi>: TF32 scores of all (sample row, query) pairs, never materialised;
//      a pair whose LOWER bound (score - 2^-8 |v||q| rounding bound, coarse.cu) does not exceed T_q is appended to
//      the query's candidate list.  The strided sample is just a 2-D tensor map with a longer row pitch;
//   3. intermediate levels: the k-th smallest UPPER bound (score + rounding bound) among the admitted pairs is the
//      next level's T_q (kth_bound_kernel) -- no database row is touched;
//   4. last level: bf_select_kernel re-scores every candidate in exact fp32 (one warp per candidate, 128-bit
//      loads) and streams them through a block-wide pool that keeps the best k by (distance, id).
// The lower levels add 1/7 to the contraction work.  The candidate set is a superset of the exact answer whatever
// the tensor-core rounding, so ids and distances are those of the exact path.  If a candidate list overflows
// (heavy ties, or a sample pitch that resonates with the data order) the caller falls back to the scan kernel.
#include "bruteforce_tc.cuh"
#include "rowtile_gemm.cuh"
#include "topk.cuh"

#include <cmath>
#include <mutex>

namespace vdb {
namespace {

using namespace tc;

// |v| and |v|^2 of every row (warp per row, 128-bit loads)
__global__ void norms_kernel(const float* __restrict__ x, uint64_t n, uint32_t ld, float* __restrict__ norm,
                             float* __restrict__ norm2) {
    const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (w >= n) return;
    const float4* r4 = reinterpret_cast<const float4*>(x + w * ld);
    float s = 0.f;
    for (uint32_t c = lane; c < (ld >> 2); c += 32) {
        const float4 v = r4[c];
        s = fmaf(v.x, v.x, s);
        s = fmaf(v.y, v.y, s);
        s = fmaf(v.z, v.z, s);
        s = fmaf(v.w, v.w, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        norm2[w] = s;
        norm[w] = sqrtf(s);
    }
}

// Admission test of the contraction's epilogue.  A pair (v, q) must be kept when the LOWER bound of its true score
// does not exceed T_q (the exact k-th distance of the previous level):
//     L2:  |v|^2 + |q|^2 - 2 dot - E <= T_q        IP:  -dot - E <= T_q
// with E = c |v||q| bounding the TF32 rounding of dot (c = 1.05 * 2^-8, coarse.cu) plus 1e-6 of every magnitude
// involved (eps = (dim + 16) 2^-24, at least 1e-6) for the fp32 rounding of this arithmetic and of the exact kernel's
// own summation.  Split into a per-query part A_q, a per-row part R_v and
// a cross term, the test is two FMAs and a compare per pair:
//     fma(-|v| c', |q|, fma(dot, alpha, A_q)) <= -R_v
//     L2: alpha = -2, A_q = |q|^2 - T_q - eps (|q|^2 + |T_q|), R_v = |v|^2 (1 - eps), c' = c + 2 eps
//     IP: alpha = -1, A_q = -T_q - eps |T_q|,                   R_v = 0,              c' = c + eps
// threshold_kernel prepares A_q and |q| for the level about to run (T_q = +inf, i.e. A_q = -inf, admits everything;
// the padding columns up to a multiple of the tile get A = +inf and admit nothing) and clears the counters.
__global__ void threshold_kernel(const float* __restrict__ prev_d, uint32_t nq, uint32_t npad, uint32_t k, int metric,
                                 float eps,
                                 const float* __restrict__ qnorm, const float* __restrict__ qnorm2,
                                 float* __restrict__ colA, float* __restrict__ colB, uint32_t* __restrict__ ccount) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= npad) return;
    if (q >= nq) {
        colA[q] = INFINITY;
        colB[q] = 0.f;
        return;
    }
    const float d = prev_d[(size_t)q * k + k - 1];
    const float T = d >= FLT_MAX ? INFINITY : d;
    const float q2 = qnorm2[q];
    colA[q] = (metric == VDB_METRIC_L2) ? (q2 - T) - eps * (q2 + fabsf(T)) : -T - eps * fabsf(T);
    colB[q] = qnorm[q];
    ccount[q] = 0;
}

constexpr int BRUTE_AN = 256;    // queries per accumulator tile
constexpr uint32_t WQ_CAP = 254;  // admissions a warp queues in shared memory before it drains them

struct BruteEpi {
    const float* vnorm;   // [n] |v|
    const float* vnorm2;  // [n] |v|^2
    const float* colA;    // [N padded to BRUTE_AN] per-query part of the admission test
    const float* colB;    // [N padded] |q|
    uint32_t M, N, num_kb;
    uint32_t pitch;       // sample row r is database row r * pitch
    int metric;
    float eps;
    uint32_t* cand_rows;  // [N][cap]
    float* cand_ub;       // [N][cap] upper bound of the admitted pair's true score; null on the last level
    const float* qnorm2;  // [N] |q|^2 (admission path only)
    uint32_t* ccount;     // [N]
    uint32_t cap, n_split;

    // An admission needs a slot in the query's candidate list: a global atomic whose result is needed at once.
    // Done in place it costs the warp one L2 round trip per admitted pair; queued in shared memory and drained by
    // all 32 lanes together it costs one round trip per 32 pairs.
    static constexpr uint32_t WARP_SMEM = (WQ_CAP + 2) * 8 + WQ_CAP * 4;  // count, (query, row) pairs, bounds
    struct State {
        float nve, negR, alpha, v2;
        uint32_t* wq;
    };
    __device__ __forceinline__ void place(uint32_t n, uint32_t row, float ub) const {
        const uint32_t pos = atomicAdd(&ccount[n], 1u);
        if (pos < cap) {
            cand_rows[(size_t)n * cap + pos] = row;
            if (cand_ub) cand_ub[(size_t)n * cap + pos] = ub;
        }
    }
    __device__ __forceinline__ void drain(uint32_t* wq) const {
        const uint32_t cnt = min(wq[0], WQ_CAP);
        const uint2* e = reinterpret_cast<const uint2*>(wq + 2);
        const float* ub = reinterpret_cast<const float*>(wq + 2 + 2 * WQ_CAP);
        for (uint32_t i = threadIdx.x & 31; i < cnt; i += 32) place(e[i].x, e[i].y, ub[i]);
        __syncwarp();
        if ((threadIdx.x & 31) == 0) wq[0] = 0;
        __syncwarp();
    }
    __device__ __forceinline__ void begin(State& s, uint32_t row, uint8_t* wsm) const {
        const float vn = vnorm[(size_t)row * pitch], v2 = vnorm2[(size_t)row * pitch];
        const bool l2 = metric == VDB_METRIC_L2;
        s.nve = -vn * (1.05f * 0.00390625f + (l2 ? 2.f * eps : eps));
        s.negR = l2 ? -(v2 - eps * v2) : 0.f;
        s.alpha = l2 ? -2.f : -1.f;
        s.v2 = v2;
        s.wq = reinterpret_cast<uint32_t*>(wsm);
        if ((threadIdx.x & 31) == 0) s.wq[0] = 0;
        __syncwarp();
    }
    __device__ __forceinline__ void consume_chunk(State& s, uint32_t row, uint32_t n0, const uint32_t (&acc)[32]) const {
        const float4* A4 = reinterpret_cast<const float4*>(colA + n0);
        const float4* B4 = reinterpret_cast<const float4*>(colB + n0);
        uint32_t hit = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 a = __ldg(A4 + j), b = __ldg(B4 + j);
            const float t0 = fmaf(s.nve, b.x, fmaf(__uint_as_float(acc[4 * j + 0]), s.alpha, a.x));
            const float t1 = fmaf(s.nve, b.y, fmaf(__uint_as_float(acc[4 * j + 1]), s.alpha, a.y));
            const float t2 = fmaf(s.nve, b.z, fmaf(__uint_as_float(acc[4 * j + 2]), s.alpha, a.z));
            const float t3 = fmaf(s.nve, b.w, fmaf(__uint_as_float(acc[4 * j + 3]), s.alpha, a.w));
            hit |= (t0 <= s.negR ? 1u : 0u) << (4 * j);
            hit |= (t1 <= s.negR ? 1u : 0u) << (4 * j + 1);
            hit |= (t2 <= s.negR ? 1u : 0u) << (4 * j + 2);
            hit |= (t3 <= s.negR ? 1u : 0u) << (4 * j + 3);
        }
        while (hit) {  // rare: about ratio * k admissions per query and level
            const uint32_t i = __ffs(hit) - 1, n = n0 + i;
            hit &= hit - 1;
            float ub = 0.f;
            if (cand_ub) {  // score + E, with the same E as the admission test
                float dot = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if ((uint32_t)j == i) dot = __uint_as_float(acc[j]);
                const float q2 = __ldg(qnorm2 + n), cross = -s.nve * __ldg(colB + n);
                ub = (metric == VDB_METRIC_L2) ? fmaf(dot, -2.f, s.v2 + q2) + cross + eps * (s.v2 + q2)
                                               : -dot + cross;
            }
            const uint32_t pos = atomicAdd(&s.wq[0], 1u);
            if (pos < WQ_CAP) {
                reinterpret_cast<uint2*>(s.wq + 2)[pos] = make_uint2(n, row);
                reinterpret_cast<float*>(s.wq + 2 + 2 * WQ_CAP)[pos] = ub;
            } else {
                place(n, row, ub);  // queue full (a burst of admissions): the slow way
            }
        }
    }
    __device__ __forceinline__ void chunk_end(State& s, uint8_t*) const {
        if (s.wq[0] >= WQ_CAP / 2) drain(s.wq);  // warp-uniform: every lane reads the same word
    }
    __device__ __forceinline__ void end(State& s, uint32_t, bool, uint8_t*) const { drain(s.wq); }
};

// Intermediate level: T_q = k-th smallest upper bound among the query's admitted pairs.  At least k rows of the
// level have a true score <= that value, so it bounds the level's k-th score -- and hence every larger level's --
// from above.  Writes the next level's admission constants and clears the counter (threshold_kernel's job).
__global__ void __launch_bounds__(256)
kth_bound_kernel(const float* __restrict__ cand_ub, uint32_t* __restrict__ ccount, uint32_t cap, uint32_t k, int metric,
                 float eps, const float* __restrict__ qnorm, const float* __restrict__ qnorm2, float* __restrict__ colA,
                 float* __restrict__ colB) {
    extern __shared__ float sub[];
    const uint32_t q = blockIdx.x, tid = threadIdx.x;
    const uint32_t c = min(ccount[q], cap), n2 = dev_next_pow2(max(c, 1u));
    for (uint32_t i = tid; i < n2; i += 256) sub[i] = i < c ? cand_ub[(size_t)q * cap + i] : INFINITY;
    __syncthreads();
    for (uint32_t size = 2; size <= n2; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (n2 >> 1); t += 256) {
                const uint32_t i = ((t / stride) * (stride << 1)) + (t % stride), j = i + stride;
                const float a = sub[i], b = sub[j];
                if (((i & size) == 0) ? (b < a) : (a < b)) {
                    sub[i] = b;
                    sub[j] = a;
                }
            }
            __syncthreads();
        }
    if (tid == 0) {
        const float T = c >= k ? sub[k - 1] : INFINITY;
        const float q2 = qnorm2[q];
        colA[q] = (metric == VDB_METRIC_L2) ? (q2 - T) - eps * (q2 + fabsf(T)) : -T - eps * fabsf(T);
        colB[q] = qnorm[q];
        ccount[q] = 0;
    }
}

constexpr int SEL_ROUND = 8 * 16;  // candidates per round: 8 warps x 16

__global__ void __launch_bounds__(MERGE_THREADS)
bf_select_kernel(const float* __restrict__ db, uint32_t ld, const uint64_t* __restrict__ ids_flat,
                 const float* __restrict__ queries, uint32_t k, uint32_t P, int metric,
                 const uint32_t* __restrict__ cand_rows, const uint32_t* __restrict__ ccount, uint32_t cap,
                 uint32_t pitch, uint32_t implicit_rows,
                 float* __restrict__ out_d, uint64_t* __restrict__ out_i, uint32_t* __restrict__ overflow) {
    extern __shared__ __align__(16) uint8_t ssm[];
    uint64_t* pi = reinterpret_cast<uint64_t*>(ssm);  // [P] pool ids
    float* pd = reinterpret_cast<float*>(pi + P);     // [P] pool distances
    float* sq = pd + P;                               // [ld] the query
    __shared__ uint32_t cnt;
    __shared__ float thr;
    __shared__ uint32_t s_scan[MERGE_THREADS / 32 + 1];
    const MergePool pool{pd, pi, &cnt, &thr};
    const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ld4 = ld >> 2;
    for (uint32_t c = tid; c < ld4; c += MERGE_THREADS)
        reinterpret_cast<float4*>(sq)[c] = reinterpret_cast<const float4*>(queries + (size_t)q * ld)[c];
    if (tid == 0) {
        cnt = 0;
        thr = INFINITY;
        if (!implicit_rows && ccount[q] > cap) atomicExch(overflow, 1u);
    }
    __syncthreads();
    // level 0 has no threshold yet: every sample row is a candidate
    const uint32_t nc = implicit_rows ? implicit_rows : min(ccount[q], cap);
    const uint32_t* mine = cand_rows + (size_t)q * cap;
    const float4* q4 = reinterpret_cast<const float4*>(sq);
    for (uint32_t base = 0; base < nc; base += SEL_ROUND) {
        if (cnt + SEL_ROUND > P) pool_compact_block(pool, P, k, false, nullptr, nullptr, s_scan);
        const float t = thr;
        // warp w re-scores candidates base + w, base + w + 8, ... two at a time (overlapping their L2 round trips)
        for (uint32_t i = base + warp; i < min(nc, base + SEL_ROUND); i += 16) {
            const uint32_t i2 = i + 8;
            const bool two = i2 < min(nc, base + SEL_ROUND);
            const uint32_t ja = i, jb = two ? i2 : i;
            const uint32_t ra = (implicit_rows ? ja : mine[ja]) * pitch, rb = (implicit_rows ? jb : mine[jb]) * pitch;
            const float4* va = reinterpret_cast<const float4*>(db + (size_t)ra * ld);
            const float4* vb = reinterpret_cast<const float4*>(db + (size_t)rb * ld);
            float a = 0.f, b = 0.f;
            for (uint32_t c = lane; c < ld4; c += 32) {
                const float4 qq = q4[c], xa = va[c], xb = vb[c];
                if (metric == VDB_METRIC_L2) {
                    float u;
                    u = qq.x - xa.x; a = fmaf(u, u, a);
                    u = qq.y - xa.y; a = fmaf(u, u, a);
                    u = qq.z - xa.z; a = fmaf(u, u, a);
                    u = qq.w - xa.w; a = fmaf(u, u, a);
                    u = qq.x - xb.x; b = fmaf(u, u, b);
                    u = qq.y - xb.y; b = fmaf(u, u, b);
                    u = qq.z - xb.z; b = fmaf(u, u, b);
                    u = qq.w - xb.w; b = fmaf(u, u, b);
                } else {
                    a = fmaf(qq.x, xa.x, a); a = fmaf(qq.y, xa.y, a); a = fmaf(qq.z, xa.z, a); a = fmaf(qq.w, xa.w, a);
                    b = fmaf(qq.x, xb.x, b); b = fmaf(qq.y, xb.y, b); b = fmaf(qq.z, xb.z, b); b = fmaf(qq.w, xb.w, b);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (metric != VDB_METRIC_L2) {
                a = -a;
                b = -b;
            }
            if (lane == 0) {
                if (a <= t) {
                    const uint32_t pos = atomicAdd(&cnt, 1u);
                    if (pos < P) {
                        pd[pos] = a;
                        pi[pos] = ids_flat ? ids_flat[ra] : ra;
                    }
                }
                if (two && b <= t) {
                    const uint32_t pos = atomicAdd(&cnt, 1u);
                    if (pos < P) {
                        pd[pos] = b;
                        pi[pos] = ids_flat ? ids_flat[rb] : rb;
                    }
                }
            }
        }
        __syncthreads();
    }
    pool_compact_block(pool, P, k, false, nullptr, nullptr, s_scan);
    const uint32_t got = cnt;
    for (uint32_t i = tid; i < k; i += MERGE_THREADS) {
        out_d[(size_t)q * k + i] = i < got ? pd[i] : FLT_MAX;
        out_i[(size_t)q * k + i] = i < got ? pi[i] : ID_PAD;
    }
}

}  // namespace

bool bruteforce_tensor_supported(uint64_t n, uint32_t nq, uint32_t ld, uint32_t k) {
    return tc::encode_tiled() != nullptr && n >= 65536 && n < (1ull << 31) && nq >= 16 && k <= 1024 && ld % 4 == 0;
}

void BruteTcScratch::release() {
    if (block) cudaFreeAsync(block, stream);
    *this = BruteTcScratch();
}

// db [n][ld], queries [nq][ld] device arrays (16-byte aligned rows).  Synchronises the stream.  *overflowed = 1
// when a candidate list overflowed (results invalid, use the scan path).
int32_t bruteforce_tensor(const float* db, uint64_t n, uint32_t ld, const uint64_t* ids_flat, const float* queries,
                          uint32_t nq, uint32_t k, int metric, float* out_d, uint64_t* out_i, BruteTcScratch& sc,
                          int* overflowed, cudaStream_t stream) {
    // Level 0 must hold a k-th neighbour: at least max(2k, 256) rows.  The pitch shrinks by an integer factor of at
    // most BF_LEVEL_RATIO per level (so the samples nest) down to 1.
    const uint64_t floor0 = std::max<uint64_t>(2ull * k, 256);
    double rem = (double)n / (double)floor0;
    uint32_t levels = 0;
    for (double t = rem; t >= 2.0; t /= (double)BF_LEVEL_RATIO) ++levels;
    uint32_t ratio[32];
    uint64_t pitch0 = 1;
    for (uint32_t i = levels; i >= 1; --i) {
        uint32_t r = (uint32_t)std::floor(std::pow(rem, 1.0 / i) + 1e-9);
        r = std::min<uint32_t>(std::max<uint32_t>(r, 1), (uint32_t)BF_LEVEL_RATIO);
        ratio[i] = r;  // level i holds ratio[i] times the rows of level i - 1
        rem /= r;
        pitch0 *= r;
    }
    // a level admits about ratio * k rows per query: 4x head room
    const uint32_t cap = (uint32_t)(4ull * BF_LEVEL_RATIO * k);
    static std::once_flag pool_once[8];
    int dev = 0, sms = NUM_SMS_B200;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev < 8)
        std::call_once(pool_once[dev], [&] {  // keep the scratch block cached between calls
            cudaMemPool_t pool;
            uint64_t keep = 1ull << 30;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        });
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const uint32_t npad = round_up(nq, (uint32_t)BRUTE_AN);
    const size_t b_n = up(n * 4), b_q = up((size_t)npad * 4), b_c = up((size_t)nq * cap * 4);
    sc.stream = stream;
    VDB_CUDA_TRY(cudaMallocAsync(&sc.block, 2 * b_n + 5 * b_q + 2 * b_c + 256, stream));
    uint8_t* at = static_cast<uint8_t*>(sc.block);
    auto take = [&](size_t b) { uint8_t* r = at; at += b; return r; };
    sc.vnorm = (float*)take(b_n); sc.vnorm2 = (float*)take(b_n);
    sc.qnorm = (float*)take(b_q); sc.qnorm2 = (float*)take(b_q); sc.colA = (float*)take(b_q); sc.colB = (float*)take(b_q);
    sc.ccount = (uint32_t*)take(b_q); sc.cand_rows = (uint32_t*)take(b_c); sc.cand_ub = (float*)take(b_c); sc.overflow = (uint32_t*)take(256);
    VDB_CUDA_TRY(cudaMemsetAsync(sc.overflow, 0, 4, stream));
    norms_kernel<<<(uint32_t)((n * 32 + 255) / 256), 256, 0, stream>>>(db, n, ld, sc.vnorm, sc.vnorm2);
    norms_kernel<<<(uint32_t)(((uint64_t)nq * 32 + 255) / 256), 256, 0, stream>>>(queries, nq, ld, sc.qnorm, sc.qnorm2);
    VDB_CUDA_TRY(cudaGetLastError());

    static bool conf[8] = {false};
    constexpr uint32_t gemm_smem = rowtile_smem<BruteEpi, BRUTE_AN>();
    const uint32_t P = next_pow2(std::max<uint32_t>(2 * k, 1024));
    const uint32_t sel_smem = P * 12 + ld * 4;
    if (dev < 8 && !conf[dev]) {
        VDB_CUDA_TRY(cudaFuncSetAttribute(rowtile_gemm_kernel<BruteEpi, BRUTE_AN>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
        VDB_CUDA_TRY(cudaFuncSetAttribute(bf_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          2048 * 12 + 2048 * 4));
        VDB_CUDA_TRY(cudaFuncSetAttribute(kth_bound_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(4ull * BF_LEVEL_RATIO * 1024 * 4)));
        conf[dev] = true;
    }
    CUtensorMap mx, mq;
    VDB_TRY(tc::make_map(&mq, queries, nq, ld, ld, BRUTE_AN));
    BruteEpi p;
    p.vnorm = sc.vnorm; p.vnorm2 = sc.vnorm2; p.colA = sc.colA; p.colB = sc.colB;
    p.N = nq; p.num_kb = (ld + GK - 1) / GK; p.metric = metric;
    p.cand_rows = sc.cand_rows; p.ccount = sc.ccount; p.cap = cap;
    p.eps = std::max(1e-6f, (float)(ld + 16) * 5.9604645e-8f);
    p.qnorm2 = sc.qnorm2;
    const uint32_t n_tiles = npad / BRUTE_AN;
    uint64_t pitch = pitch0;
    for (uint32_t lv = 0; lv <= levels; ++lv) {
        if (lv) pitch /= ratio[lv];
        const uint64_t rows = (n + pitch - 1) / pitch;
        const bool last = lv == levels;
        if (lv) {
            VDB_TRY(tc::make_map(&mx, db, rows, ld, ld * pitch, AM));
            p.M = (uint32_t)rows;
            p.pitch = (uint32_t)pitch;
            p.cand_ub = last ? nullptr : sc.cand_ub;
            const uint32_t m_tiles = (uint32_t)((rows + AM - 1) / AM);
            // small levels: spread the query tiles of a row tile over several CTAs
            p.n_split = std::max<uint32_t>(1, std::min<uint32_t>(n_tiles, (uint32_t)sms / m_tiles));
            rowtile_gemm_kernel<BruteEpi, BRUTE_AN>
                <<<std::min<uint32_t>(m_tiles * p.n_split, (uint32_t)sms), ATHREADS, gemm_smem, stream>>>(mx, mq, p);
        }
        if (lv && !last) {
            const uint32_t c2 = next_pow2(cap);
            kth_bound_kernel<<<nq, 256, c2 * 4, stream>>>(sc.cand_ub, sc.ccount, cap, k, metric, p.eps, sc.qnorm,
                                                          sc.qnorm2, sc.colA, sc.colB);
        } else {
            // level 0 has no threshold yet: every one of its rows is a candidate
            bf_select_kernel<<<nq, MERGE_THREADS, sel_smem, stream>>>(db, ld, ids_flat, queries, k, P, metric,
                                                                      sc.cand_rows, sc.ccount, cap, (uint32_t)pitch,
                                                                      lv ? 0u : (uint32_t)rows, out_d, out_i,
                                                                      sc.overflow);
            if (!last)
                threshold_kernel<<<(npad + 255) / 256, 256, 0, stream>>>(out_d, nq, npad, k, metric, p.eps, sc.qnorm,
                                                                         sc.qnorm2, sc.colA, sc.colB, sc.ccount);
        }
        VDB_CUDA_TRY(cudaGetLastError());
    }
    uint32_t h = 0;
    VDB_CUDA_TRY(cudaMemcpyAsync(&h, sc.overflow, 4, cudaMemcpyDeviceToHost, stream));
    VDB_CUDA_TRY(cudaStreamSynchronize(stream));
    *overflowed = (int)h;
    return VDB_OK;
}

}  // namespace vdb
