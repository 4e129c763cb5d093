// k-means training and assignment kernels for sm_100a.
//
// Replaces IVFFlatIndex::train / assign_to_lists (ivf_flat_index.cpp:49-145,
// 259-295) and kmeans_assign_kernel (kernels.cuh:315-354).
//
// The kernels in this file are the ORDER-EXACT family: every fp32 sum is
// accumulated in the reference's order (ascending dimension for distances,
// ascending input row for centroid sums) with separate round-to-nearest
// multiply and add (__fmul_rn/__fadd_rn are never contracted into FMA), so
// centroids, assignments and list membership are bit-identical to the
// reference's CPU path, including its seeded k-means++ (std::mt19937(42) and
// libstdc++'s uniform distributions are restated on the device).
#include "kmeans.cuh"

#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "tc_common.cuh"

namespace vdb {
namespace {

// ---------------------------------------------------------------- assignment

constexpr int TV = 64, TC = 64, DK = 32;

// 64 vectors x 64 centroids per block step, 4x4 pairs per thread; the d loop
// runs in ascending order so each pair's sum is the reference's sequential sum.
__global__ void __launch_bounds__(256) assign_exact_kernel(const float* __restrict__ x, uint64_t n, uint32_t ldx,
                                                           const float* __restrict__ c, uint32_t nc, uint32_t ldc,
                                                           uint32_t dim, int metric, uint32_t* __restrict__ assign,
                                                           float* __restrict__ dist_out,
                                                           const uint32_t* __restrict__ row_index) {
    __shared__ float sv[TV][DK + 1];
    __shared__ float sc[TC][DK + 1];
    __shared__ float sbd[TV][17];
    __shared__ uint32_t sbi[TV][17];
    const uint32_t tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const uint64_t v0 = (uint64_t)blockIdx.x * TV;

    float bd[4];
    uint32_t bi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bd[i] = FLT_MAX;  // min_dist = numeric_limits<float>::max(), best_list = 0 (:266-267)
        bi[i] = 0;
    }
    for (uint32_t c0 = 0; c0 < nc; c0 += TC) {
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (uint32_t d0 = 0; d0 < dim; d0 += DK) {
            for (uint32_t e = tid; e < TV * DK; e += 256) {
                const uint32_t r = e / DK, dd = e % DK;
                const uint64_t v = v0 + r;
                const uint64_t src = (v < n && row_index) ? row_index[v] : v;  // optional gather of selected rows
                sv[r][dd] = (v < n && d0 + dd < dim) ? x[src * ldx + d0 + dd] : 0.f;
                const uint32_t cc = c0 + r;
                sc[r][dd] = (cc < nc && d0 + dd < dim) ? c[(size_t)cc * ldc + d0 + dd] : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int dd = 0; dd < DK; ++dd) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = sv[ty + 16 * i][dd];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = sc[tx + 16 * j][dd];
                if (metric == VDB_METRIC_L2) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float diff = __fsub_rn(a[i], b[j]);
                            acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(diff, diff));
                        }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = __fadd_rn(acc[i][j], __fmul_rn(a[i], b[j]));
                }
            }
            __syncthreads();
        }
        // strict '<' in ascending centroid order: the lowest index wins ties (:287)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t cc = c0 + tx + 16 * j;
            if (cc < nc) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float d = (metric == VDB_METRIC_L2) ? acc[i][j] : -acc[i][j];
                    if (d < bd[i]) {
                        bd[i] = d;
                        bi[i] = cc;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sbd[ty + 16 * i][tx] = bd[i];
        sbi[ty + 16 * i][tx] = bi[i];
    }
    __syncthreads();
    if (tid < TV) {
        float d = sbd[tid][0];
        uint32_t b = sbi[tid][0];
        for (int t = 1; t < 16; ++t) {
            const float dt = sbd[tid][t];
            const uint32_t bt = sbi[tid][t];
            if (dt < d || (dt == d && bt < b)) {
                d = dt;
                b = bt;
            }
        }
        const uint64_t v = v0 + tid;
        if (v < n) {
            const uint64_t dst = row_index ? row_index[v] : v;
            assign[dst] = b;
            if (dist_out) dist_out[dst] = d;
        }
    }
}

// ------------------------------------------------------------------ seeding

// std::mt19937 state kept in device memory
struct DevRng {
    uint32_t mt[624];
    uint32_t idx;
};

__device__ void rng_seed(DevRng* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

__device__ uint32_t rng_next(DevRng* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// libstdc++ uniform_int_distribution<uint64_t>(0, n-1) on a 32-bit engine (Lemire)
__device__ uint64_t rng_below(DevRng* g, uint32_t range) {
    uint64_t product = (uint64_t)rng_next(g) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        const uint32_t threshold = (0u - range) % range;
        while (low < threshold) {
            product = (uint64_t)rng_next(g) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return product >> 32;
}

// libstdc++ uniform_real_distribution<float>(0, b): generate_canonical<float,24>
__device__ float rng_real_0_b(DevRng* g, float b) {
    float ret = __fdiv_rn(__uint2float_rn(rng_next(g)), 4294967296.0f);
    if (ret >= 1.0f) ret = __uint_as_float(0x3f7fffffu);  // nextafter(1, 0)
    return __fadd_rn(__fmul_rn(ret, __fsub_rn(b, 0.0f)), 0.0f);
}

// first centroid = row uniform_int(0, n-1) (ivf_flat_index.cpp:53-60); min-distances start at FLT_MAX
__global__ void seed_init_kernel(DevRng* g, const float* __restrict__ x, uint32_t n, uint32_t ldx, uint32_t ld,
                                 float* __restrict__ centroids, uint32_t* picked) {
    __shared__ uint32_t s_first;
    if (threadIdx.x == 0) {
        rng_seed(g, 42);
        s_first = (uint32_t)rng_below(g, n);
        picked[0] = s_first;
    }
    __syncthreads();
    for (uint32_t d = threadIdx.x; d < ld; d += blockDim.x) centroids[d] = x[(size_t)s_first * ldx + d];
}

__global__ void fill_f32_kernel(float* p, uint64_t n, float v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// squared L2 distance of every training row to the newest centroid, summed in
// ascending d (one thread per row, tiles transposed through shared memory so
// the HBM reads stay coalesced), folded into the running minimum (:68-88).
__global__ void __launch_bounds__(128) seed_dist_kernel(const float* __restrict__ x, uint32_t n, uint32_t ldx,
                                                        uint32_t dim, const float* __restrict__ cnew,
                                                        const float* mind_in, const PeerF32 mind_out) {
    __shared__ float tile[128][33];
    __shared__ float sc[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint64_t v0 = (uint64_t)blockIdx.x * 128;
    float acc = 0.f;
    for (uint32_t d0 = 0; d0 < dim; d0 += 32) {
        const bool dok = d0 + lane < dim;
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
            const uint64_t v = v0 + w * 32 + i;
            tile[w * 32 + i][lane] = (v < n && dok) ? x[v * ldx + d0 + lane] : 0.f;
        }
        if (tid < 32) sc[tid] = dok ? cnew[d0 + tid] : 0.f;
        __syncthreads();
#pragma unroll
        for (int dd = 0; dd < 32; ++dd) {
            const float diff = __fsub_rn(tile[tid][dd], sc[dd]);
            acc = __fadd_rn(acc, __fmul_rn(diff, diff));
        }
        __syncthreads();
    }
    const uint64_t v = v0 + tid;
    if (v < n) {
        const float m = mind_in[v];
        const float r = acc < m ? acc : m;  // std::min(min_dist, dist)
        for (uint32_t pr = 0; pr < mind_out.n; ++pr) mind_out.p[pr][v] = r;
    }
}

// The same update with the rows arriving by TMA: a block owns 128 rows, one producer thread streams their 128-byte
// column blocks (2-D tensor map, 128-byte swizzle) through a 4-stage mbarrier ring, thread t adds up row t -- eight
// LDS.128 per stage, conflict-free under the swizzle -- in the same ascending order.  The thread-per-row kernel
// above stalls on its own load -> transpose -> add sequence (3.3 TB/s); this one keeps 64 KB per block in flight.
constexpr int SD_ROWS = 128, SD_STAGES = 4;
constexpr uint32_t SD_STAGE_BYTES = SD_ROWS * 128;
constexpr uint32_t SD_SMEM = SD_STAGES * SD_STAGE_BYTES + 2048 * 4 + 2 * SD_STAGES * 8 + 1024;

__global__ void __launch_bounds__(SD_ROWS + 32) seed_dist_tma_kernel(const __grid_constant__ CUtensorMap map_x,
                                                                      uint32_t n, uint32_t ld,
                                                                      const float* __restrict__ cnew,
                                                                      const float* mind_in, const PeerF32 mind_out) {
    using namespace tc;
    extern __shared__ __align__(1024) uint8_t sd_raw[];
    uint8_t* smem = sd_raw + ((1024u - (smem_u32(sd_raw) & 1023u)) & 1023u);
    float* sc = reinterpret_cast<float*>(smem + SD_STAGES * SD_STAGE_BYTES);  // the new centroid, zero padded
    uint64_t* full = reinterpret_cast<uint64_t*>(sc + 2048);
    uint64_t* empty = full + SD_STAGES;
    const uint32_t tid = threadIdx.x, num_kb = (ld + 31) / 32;
    const uint32_t row0 = blockIdx.x * SD_ROWS;
    for (uint32_t d = tid; d < num_kb * 32; d += SD_ROWS + 32) sc[d] = d < ld ? cnew[d] : 0.f;
    if (tid == 0) {
        for (int i = 0; i < SD_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], SD_ROWS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == SD_ROWS) {  // producer
        uint32_t s = 0, ph = 0;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], SD_STAGE_BYTES);
            tma_load_2d(smem + s * SD_STAGE_BYTES, &map_x, (int32_t)(kb * 32), (int32_t)row0, &full[s]);
            if (++s == SD_STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
    } else if (tid < SD_ROWS) {
        uint32_t s = 0, ph = 0;
        float acc = 0.f;
        const uint32_t swz = tid & 7u;
        for (uint32_t kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full[s], ph);
            const uint8_t* rowp = smem + s * SD_STAGE_BYTES + tid * 128;
            const float4* c4 = reinterpret_cast<const float4*>(sc + kb * 32);
#pragma unroll
            for (uint32_t c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(rowp + ((c ^ swz) << 4));
                const float4 q = c4[c];
                float diff;
                diff = __fsub_rn(v.x, q.x); acc = __fadd_rn(acc, __fmul_rn(diff, diff));
                diff = __fsub_rn(v.y, q.y); acc = __fadd_rn(acc, __fmul_rn(diff, diff));
                diff = __fsub_rn(v.z, q.z); acc = __fadd_rn(acc, __fmul_rn(diff, diff));
                diff = __fsub_rn(v.w, q.w); acc = __fadd_rn(acc, __fmul_rn(diff, diff));
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
            if (++s == SD_STAGES) {
                s = 0;
                ph ^= 1;
            }
        }
        const uint32_t v = row0 + tid;
        if (v < n) {
            const float m = mind_in[v];
            const float r = acc < m ? acc : m;  // std::min(min_dist, dist)
            // data-parallel training: this rank owns a slice of the rows and stores the slice's new minima
            // straight into every rank's copy (NVLink peer stores); single GPU: one copy, in place
            for (uint32_t pr = 0; pr < mind_out.n; ++pr) mind_out.p[pr][v] = r;
        }
    }
}

constexpr uint32_t SEQ_CHUNK = 1024;

// One warp: total = sequential fp32 sum of mind[] (ivf_flat_index.cpp:87),
// target = uniform_real(0,total) (:91-92), first row whose sequential running
// sum reaches the target becomes centroid `cidx` (:95-103).  Lane 0 carries the
// dependent add chain; the warp stages the data through shared memory.
__global__ void __launch_bounds__(32) seed_sample_kernel(DevRng* g, const float* __restrict__ x, uint32_t n,
                                                         uint32_t ldx, uint32_t ld, const float* __restrict__ mind,
                                                         float* __restrict__ ckpt, float* __restrict__ centroids,
                                                         uint32_t cidx, uint32_t* picked) {
    __shared__ __align__(16) float buf[SEQ_CHUNK];
    __shared__ uint32_t s_pick;
    __shared__ float s_start;
    const uint32_t lane = threadIdx.x;
    const uint32_t nchunks = (n + SEQ_CHUNK - 1) / SEQ_CHUNK;
    float run = 0.f;
    for (uint32_t ch = 0; ch < nchunks; ++ch) {
        const uint32_t base = ch * SEQ_CHUNK;
        for (uint32_t i = lane; i < SEQ_CHUNK; i += 32) buf[i] = (base + i < n) ? mind[base + i] : 0.f;
        __syncwarp();
        if (lane == 0) {
            // the dependent fp32 add chain is the critical path (4 cycles per element): keep the shared-memory
            // loads off it by fetching the next 32 values while the current 32 are being added
            const uint32_t m = min(SEQ_CHUNK, n - base);
            const float4* b4 = reinterpret_cast<const float4*>(buf);
            float4 cur[8], nxt[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[u] = b4[u];
            uint32_t i = 0;
            for (; i + 32 <= m; i += 32) {
                const uint32_t nb = (i + 32) >> 2;
#pragma unroll
                for (int u = 0; u < 8; ++u) nxt[u] = b4[min(nb + u, SEQ_CHUNK / 4 - 1)];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    run = __fadd_rn(run, cur[u].x);
                    run = __fadd_rn(run, cur[u].y);
                    run = __fadd_rn(run, cur[u].z);
                    run = __fadd_rn(run, cur[u].w);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) cur[u] = nxt[u];
            }
            for (; i < m; ++i) run = __fadd_rn(run, buf[i]);
            ckpt[ch] = run;  // running sum after this chunk
        }
        __syncwarp();
    }
    if (lane == 0) {
        const float total = run;
        const float target = rng_real_0_b(g, total);
        // running sums are non-decreasing (the terms are >= 0), so the first chunk whose
        // end-of-chunk sum reaches the target holds the first row that does
        uint32_t ch = 0;
        while (ch < nchunks && !(ckpt[ch] >= target)) ++ch;
        s_pick = 0xffffffffu;
        s_start = 0.f;
        if (ch < nchunks) {
            s_pick = ch;
            s_start = ch ? ckpt[ch - 1] : 0.f;
        }
        ckpt[nchunks] = target;
    }
    __syncwarp();
    const uint32_t ch = s_pick;
    if (ch == 0xffffffffu) {  // no row reaches the target: the reference leaves the slot as it was
        if (lane == 0) picked[cidx] = 0xffffffffu;
        return;
    }
    const uint32_t base = ch * SEQ_CHUNK;
    for (uint32_t i = lane; i < SEQ_CHUNK; i += 32) buf[i] = (base + i < n) ? mind[base + i] : 0.f;
    __syncwarp();
    if (lane == 0) {
        const float target = ckpt[nchunks];
        float cum = s_start;
        const uint32_t m = min(SEQ_CHUNK, n - base);
        uint32_t v = 0xffffffffu;
        for (uint32_t i = 0; i < m; ++i) {
            cum = __fadd_rn(cum, buf[i]);
            if (cum >= target) {
                v = base + i;
                break;
            }
        }
        s_pick = v;
        picked[cidx] = v;
    }
    __syncwarp();
    const uint32_t v = s_pick;
    if (v == 0xffffffffu) return;
    for (uint32_t d = lane; d < ld; d += 32) centroids[(size_t)cidx * ld + d] = x[(size_t)v * ldx + d];
}

// ---- the same sequential fp32 sums, evaluated in parallel, bit for bit ----------------------------------------
// s_{i+1} = fl(s_i + x_i) with x_i >= 0 looks inherently serial, but while the running sum stays inside one binade
// (s = S * 2^eu, S an integer below 2^24) round-to-nearest-even is integer arithmetic on S:
//     S' = S + y + c,   x / 2^eu = y + f,   c = [f > 1/2], or for a tie f = 1/2 the parity of S + y.
// An element is therefore a map S -> S + (increment that depends only on the parity of S), such maps compose into
// maps of the same kind (a pair of increments, one per parity of the start value), and composition is associative:
// a block-wide scan gives every thread the exact sum in front of its elements.  Leaving the binade (S' >= 2^24,
// about log2(sum / first term) ~ 20-40 times per pass) is detected at the first element where it happens; that one
// addition is done with a real __fadd_rn, a few dozen more follow sequentially (the sum grows fast at the start),
// and the scan resumes behind them with the new unit.  Results equal seed_sample_kernel's
// (tests/test_gpu_parity.py::test_parallel_exact_sampler_matches_sequential) at ~1/10 of its time.
constexpr uint32_t PS_THREADS = 1024, PS_E = 16;  // up to 16384 terms per round
constexpr uint32_t PS_SAT = 1u << 26;  // increments at or above 2^24 only ever mean "left the binade"
constexpr uint32_t PS_NONE = 0xffffffffu;

__device__ __forceinline__ uint32_t ps_sat(uint32_t v) { return v < PS_SAT ? v : PS_SAT; }

// x (>= 0) in units of 2^eu: integer part y and rounding class ct (0 down, 1 up, 2 tie)
__device__ __forceinline__ void ps_step(float x, int eu, uint32_t& y, uint32_t& ct) {
    const uint32_t b = __float_as_uint(x) & 0x7fffffffu;
    const uint32_t ef = b >> 23;
    const uint32_t m = ef ? ((b & 0x7fffffu) | 0x800000u) : b;
    y = 0;
    ct = 0;
    if (m == 0) return;
    const int d = eu - ((ef ? (int)ef : 1) - 150);  // x = m * 2^(ef - 150): shift that aligns it to the unit
    if (d <= 0) {
        y = (d < -2) ? PS_SAT : ps_sat(m << (-d));
    } else if (d <= 24) {
        const uint32_t half = 1u << (d - 1);
        const uint32_t r = m & ((half << 1) - 1u);
        y = (d == 24) ? 0u : (m >> d);
        ct = r > half ? 1u : (r == half ? 2u : 0u);
    }
}

__device__ __forceinline__ uint32_t ps_apply(uint32_t S, uint32_t y, uint32_t ct) {
    return ps_sat(S + y + (ct == 2u ? ((S + y) & 1u) : ct));
}

struct PsMap {
    uint32_t e, o;  // total increment for an even / odd start value
};
__device__ __forceinline__ PsMap ps_then(PsMap f, PsMap g) {  // g after f
    PsMap h;
    h.e = ps_sat(f.e + ((f.e & 1u) ? g.o : g.e));
    h.o = ps_sat(f.o + (((1u + f.o) & 1u) ? g.o : g.e));
    return h;
}

// Walks mind[from..n) with the sequential fp32 running sum (`from_sum` = the sum in front of `from`); returns the
// first index whose running sum is >= target (PS_NONE if none) and leaves the final sum in *total.  Runs on a
// cluster of PS_CTAS thread blocks (one round = up to PS_CTAS x 16384 terms): every block scans its slice, the block
// maps and the per-block outcomes are exchanged through distributed shared memory (two cluster barriers per round,
// buffers alternate by round parity), and every block then takes the same decision redundantly, so position and sum
// stay identical in all of them without a broadcast.  With `ck`, the (position, sum) at the start of every round is
// recorded so that a second walk can start next to its target instead of at 0.
constexpr uint32_t PS_CTAS = 8;
constexpr uint32_t PS_HEAD = 4096;
constexpr uint32_t PS_CKPTS = 512;
struct PsCkpt {
    uint32_t pos[PS_CKPTS];
    float sum[PS_CKPTS];
    uint32_t count;
};
struct PsShare {  // what a block shows to its cluster peers, per round parity
    PsMap map;    // composition of all the block's terms
    uint32_t found, cross, crossS, endS;
};

__device__ uint32_t ps_walk(const float* __restrict__ mind, uint32_t n, float target, float* total, uint32_t from,
                            float from_sum, PsCkpt* ck) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t crank = cluster.block_rank();
    __shared__ float s_sum;
    __shared__ uint32_t s_pos, s_found, s_cross, s_crossS, s_endS, s_round;
    __shared__ PsMap s_warp[32];
    __shared__ PsMap s_ctapre;
    __shared__ PsShare s_share[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) {
        s_sum = from_sum;
        s_pos = from;
        s_round = 0;
        if (ck) {  // the walk's own start is the first checkpoint
            ck->count = 1;
            ck->pos[0] = from;
            ck->sum[0] = from_sum;
        }
    }
    if (from == 0) {
        // The sum leaves a binade every few terms while it is small: the first PS_HEAD terms go through the plain
        // one-lane chain (staged in shared memory, 4 cycles per add), every block redundantly.
        __shared__ __align__(16) float s_head[PS_HEAD];
        const uint32_t m = min(n, PS_HEAD);
        for (uint32_t i = tid; i < PS_HEAD; i += PS_THREADS) s_head[i] = i < m ? mind[i] : 0.f;
        __syncthreads();
        if (tid == 0) {
            float run = from_sum;
            uint32_t hit = PS_NONE;
            const float4* h4 = reinterpret_cast<const float4*>(s_head);
            for (uint32_t i = 0; i < PS_HEAD / 4 && hit == PS_NONE; ++i) {
                const float4 v = h4[i];
                const float r0 = __fadd_rn(run, v.x), r1 = __fadd_rn(r0, v.y), r2 = __fadd_rn(r1, v.z),
                            r3 = __fadd_rn(r2, v.w);
                run = r3;
                if (r3 >= target) {  // padding terms are zeros: a hit can only be at a real row or repeat one
                    hit = 4 * i + (r0 >= target ? 0u : r1 >= target ? 1u : r2 >= target ? 2u : 3u);
                }
            }
            if (hit != PS_NONE && hit < m) {
                s_pos = 0xfffffffeu;
                s_endS = hit;
            } else {
                s_sum = run;
                s_pos = m;
            }
        }
        __syncthreads();
        if (s_pos == 0xfffffffeu) {
            const uint32_t r = s_endS;
            cluster.sync();
            return r;
        }
    }
    __syncthreads();
    for (;;) {
        const uint32_t pos = s_pos;
        if (pos >= n) break;
        const float sum = s_sum;
        const uint32_t par = s_round & 1u;
        if (ck && tid == 0 && ck->count < PS_CKPTS) {
            ck->pos[ck->count] = pos;
            ck->sum[ck->count] = sum;
            ++ck->count;
        }
        // terms per thread this round: right after leaving a binade the next exit is about `pos` terms away (the
        // sum has to double), so early rounds do not need -- and would mostly waste -- the full span
        const uint32_t E = min(PS_E, max(1u, (pos + PS_CTAS * PS_THREADS - 1) / (PS_CTAS * PS_THREADS)));
        const uint32_t sb = __float_as_uint(sum), sef = (sb >> 23) & 0xffu;
        const int eu = (sef ? (int)sef : 1) - 150;
        const uint32_t S0 = sef ? ((sb & 0x7fffffu) | 0x800000u) : (sb & 0x7fffffu);
        if (tid == 0) {
            s_found = PS_NONE;
            s_cross = PS_NONE;
        }
        uint32_t y[PS_E], ct[PS_E];
        PsMap f{0u, 0u};
        // pos + 8 * 16384 cannot wrap: n < 2^31
        const uint32_t cta_base = pos + crank * PS_THREADS * E, base = cta_base + tid * E;
#pragma unroll
        for (uint32_t j = 0; j < PS_E; ++j) {
            if (j >= E) break;
            const float x = (base + j < n) ? mind[base + j] : 0.f;
            ps_step(x, eu, y[j], ct[j]);
            const uint32_t ie = y[j] + (ct[j] == 2u ? ((f.e + y[j]) & 1u) : ct[j]);
            const uint32_t io = y[j] + (ct[j] == 2u ? ((1u + f.o + y[j]) & 1u) : ct[j]);
            f.e = ps_sat(f.e + ie);
            f.o = ps_sat(f.o + io);
        }
        // block-wide scan of the maps
        PsMap inc = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            PsMap other;
            other.e = __shfl_up_sync(0xffffffffu, inc.e, o);
            other.o = __shfl_up_sync(0xffffffffu, inc.o, o);
            if ((int)lane >= o) inc = ps_then(other, inc);
        }
        if (lane == 31) s_warp[w] = inc;
        PsMap ex;
        ex.e = __shfl_up_sync(0xffffffffu, inc.e, 1);
        ex.o = __shfl_up_sync(0xffffffffu, inc.o, 1);
        if (lane == 0) ex = PsMap{0u, 0u};
        __syncthreads();
        if (w == 0) {
            PsMap v = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                PsMap other;
                other.e = __shfl_up_sync(0xffffffffu, v.e, o);
                other.o = __shfl_up_sync(0xffffffffu, v.o, o);
                if ((int)lane >= o) v = ps_then(other, v);
            }
            if (lane == 31) s_share[par].map = v;  // the whole block
            PsMap pv;
            pv.e = __shfl_up_sync(0xffffffffu, v.e, 1);
            pv.o = __shfl_up_sync(0xffffffffu, v.o, 1);
            if (lane == 0) pv = PsMap{0u, 0u};
            s_warp[lane] = pv;  // exclusive prefix over the warps
        }
        cluster.sync();  // every block's map is visible
        if (tid == 0) {
            PsMap pre{0u, 0u};
            for (uint32_t r = 0; r < crank; ++r) pre = ps_then(pre, cluster.map_shared_rank(&s_share[par], r)->map);
            s_ctapre = pre;
        }
        __syncthreads();
        const PsMap pre = ps_then(s_ctapre, ps_then(s_warp[w], ex));
        uint32_t S = ps_sat(S0 + ((S0 & 1u) ? pre.o : pre.e));
        // walk my terms with the real start value: first target hit, first binade exit
        uint32_t found = PS_NONE, cross = PS_NONE, crossS = 0;
        if (S < (1u << 24)) {
#pragma unroll
            for (uint32_t j = 0; j < PS_E; ++j) {
                if (j >= E) break;
                if (found == PS_NONE && cross == PS_NONE && base + j < n) {
                    const uint32_t Sn = ps_apply(S, y[j], ct[j]);
                    if (Sn >= (1u << 24)) {
                        cross = base + j;
                        crossS = S;
                    } else {
                        S = Sn;
                        if (ldexpf((float)Sn, eu) >= target) found = base + j;
                    }
                }
            }
        } else {
            cross = base;  // an earlier thread has left the binade already: never the minimum
        }
        if (found != PS_NONE) atomicMin(&s_found, found);
        if (cross != PS_NONE && base < n) atomicMin(&s_cross, cross);
        __syncthreads();
        if (cross != PS_NONE && cross == s_cross && S < (1u << 24)) s_crossS = crossS;
        if (tid == PS_THREADS - 1) s_endS = S;
        __syncthreads();
        if (tid == 0) {
            s_share[par].found = s_found;
            s_share[par].cross = s_cross;
            s_share[par].crossS = s_crossS;
            s_share[par].endS = s_endS;
        }
        cluster.sync();  // every block's outcome is visible
        if (w == 0) {
            // lane r reads block r's outcome through distributed shared memory, a few shuffles combine them
            uint32_t fnd = PS_NONE, crs = PS_NONE, crsS = 0, endS = 0;
            if (lane < PS_CTAS) {
                const PsShare* o = cluster.map_shared_rank(&s_share[par], lane);
                fnd = o->found;
                crs = o->cross;
                crsS = o->crossS;
                endS = o->endS;
            }
            endS = __shfl_sync(0xffffffffu, endS, PS_CTAS - 1);
#pragma unroll
            for (int o = 1; o < (int)PS_CTAS; o <<= 1) {
                fnd = min(fnd, __shfl_xor_sync(0xffffffffu, fnd, o));
                const uint32_t oc = __shfl_xor_sync(0xffffffffu, crs, o), os = __shfl_xor_sync(0xffffffffu, crsS, o);
                if (oc < crs) {  // crossing indices are distinct rows (or NONE): no tie to break
                    crs = oc;
                    crsS = os;
                }
            }
            fnd = __shfl_sync(0xffffffffu, fnd, 0);
            crs = __shfl_sync(0xffffffffu, crs, 0);
            crsS = __shfl_sync(0xffffffffu, crsS, 0);
            // the term that leaves the binade and the stretch behind it, fetched in one go by the warp
            float xv = 0.f;
            if (crs != PS_NONE && !(fnd < crs)) xv = (crs + lane < n) ? mind[crs + lane] : 0.f;
            float nx[32];
#pragma unroll
            for (int t = 0; t < 32; ++t) nx[t] = __shfl_sync(0xffffffffu, xv, t);
          if (lane == 0) {
            if (fnd < crs) {  // reached the target before leaving the binade (fnd != NONE)
                s_pos = 0xfffffffeu;
                s_endS = fnd;
            } else if (crs != PS_NONE) {
                // the addition that leaves the binade, then a short sequential stretch, both with real fp32 adds
                float sacc = __fadd_rn(ldexpf((float)crsS, eu), nx[0]);
                uint32_t at = crs, hit = PS_NONE;
                if (sacc >= target) hit = at;
#pragma unroll
                for (int t = 1; t < 32; ++t) {
                    if (hit == PS_NONE && crs + t < n) {
                        sacc = __fadd_rn(sacc, nx[t]);
                        at = crs + t;
                        if (sacc >= target) hit = at;
                    }
                }
                if (hit != PS_NONE) {
                    s_pos = 0xfffffffeu;
                    s_endS = hit;
                } else {
                    s_sum = sacc;
                    s_pos = at + 1;
                }
            } else {
                s_sum = ldexpf((float)endS, eu);
                s_pos = pos + PS_CTAS * PS_THREADS * E;
            }
            ++s_round;
          }
        }
        __syncthreads();
        if (s_pos == 0xfffffffeu) {
            const uint32_t r = s_endS;
            cluster.sync();  // no block leaves (or starts another walk) while a peer may still read its outcome
            return r;
        }
    }
    if (tid == 0) *total = s_sum;
    cluster.sync();
    return PS_NONE;
}

__global__ void __cluster_dims__(PS_CTAS, 1, 1) __launch_bounds__(PS_THREADS)
seed_sample_par_kernel(DevRng* g, const float* __restrict__ x, uint32_t n, uint32_t ldx, uint32_t ld,
                       const float* __restrict__ mind, float* __restrict__ centroids, uint32_t cidx, uint32_t* picked) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float s_total, s_target, s_from_sum;
    __shared__ uint32_t s_from;
    __shared__ PsCkpt ck;
    ps_walk(mind, n, INFINITY, &s_total, 0, 0.f, &ck);  // total_dist (ivf_flat_index.cpp:87)
    // one draw from the generator (block 0), shown to the other blocks through distributed shared memory
    if (cluster.block_rank() == 0 && threadIdx.x == 0) s_target = rng_real_0_b(g, s_total);
    cluster.sync();
    if (threadIdx.x == 0) {
        const float target = *cluster.map_shared_rank(&s_target, 0);
        // running sums never decrease: the first row reaching the target lies at or behind the last recorded
        // round start whose sum is still below it
        uint32_t lo = 0;
        for (uint32_t i = 1; i < ck.count; ++i)
            if (ck.sum[i] < target) lo = i;
        s_from = ck.pos[lo];
        s_from_sum = ck.sum[lo];
        s_total = target;  // (reused as this block's copy of the target)
    }
    cluster.sync();  // block 0's s_target has been read by everyone before anything can overwrite it
    float unused;
    const uint32_t v = ps_walk(mind, n, s_total, &unused, s_from, s_from_sum, nullptr);
    if (cluster.block_rank() != 0) return;
    if (threadIdx.x == 0) picked[cidx] = v;
    if (v == PS_NONE) return;  // no row reaches the target: the reference leaves the slot as it was
    for (uint32_t d = threadIdx.x; d < ld; d += PS_THREADS) centroids[(size_t)cidx * ld + d] = x[(size_t)v * ldx + d];
}

// FAST-mode sampler: the same D^2 sampling (same RNG stream), but the total and the running sums are
// accumulated in parallel (double precision, 1024 chunks) instead of one sequential fp32 chain, so the row
// picked can differ from the reference's when the target falls within rounding distance of a boundary.
__global__ void __launch_bounds__(1024) seed_sample_fast_kernel(DevRng* g, const float* __restrict__ x, uint32_t n,
                                                                uint32_t ldx, uint32_t ld,
                                                                const float* __restrict__ mind,
                                                                float* __restrict__ centroids, uint32_t cidx,
                                                                uint32_t* picked) {
    __shared__ double s_part[1024];
    __shared__ double s_target;
    __shared__ uint32_t s_pick;
    const uint32_t tid = threadIdx.x;
    const uint32_t chunk = (n + 1023) / 1024;
    const uint32_t lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    double s = 0.0;
    for (uint32_t i = lo; i < hi; ++i) s += (double)mind[i];
    s_part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        double run = 0.0;
        for (int i = 0; i < 1024; ++i) {
            const double t = s_part[i];
            s_part[i] = run;  // exclusive prefix
            run += t;
        }
        s_target = (double)rng_real_0_b(g, (float)run);
        s_pick = 0xffffffffu;
    }
    __syncthreads();
    const double target = s_target;
    const double before = s_part[tid];
    if (before < target || (tid == 0 && target <= 0.0)) {
        double run = before;
        for (uint32_t i = lo; i < hi; ++i) {
            run += (double)mind[i];
            if (run >= target) {
                atomicMin(&s_pick, i);  // first row whose running sum reaches the target
                break;
            }
        }
    }
    __syncthreads();
    const uint32_t v = s_pick;
    if (tid == 0) picked[cidx] = v;
    if (v == 0xffffffffu) return;
    for (uint32_t d = tid; d < ld; d += 1024) centroids[(size_t)cidx * ld + d] = x[(size_t)v * ldx + d];
}

// ------------------------------------------------------------- Lloyd update

// Stable bucketing of row numbers by cluster (members of a cluster in input
// order) = a counting sort whose chunks are walked by one warp each.
__global__ void __launch_bounds__(32) member_hist_kernel(const uint32_t* __restrict__ assign, uint32_t n,
                                                         uint32_t chunk, uint32_t nc, uint32_t* __restrict__ M) {
    const uint32_t ch = blockIdx.x, lane = threadIdx.x;
    const uint64_t lo = (uint64_t)ch * chunk, hi = min((uint64_t)n, lo + chunk);
    uint32_t* row = M + (size_t)ch * nc;
    for (uint64_t v = lo + lane; v < hi; v += 32) atomicAdd(&row[assign[v]], 1u);
}

// per cluster: exclusive prefix of the chunk counts down the column, cluster total
__global__ void member_colscan_kernel(uint32_t* __restrict__ M, uint32_t nchunks, uint32_t nc,
                                      uint32_t* __restrict__ counts) {
    const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= nc) return;
    uint32_t run = 0;
    for (uint32_t ch = 0; ch < nchunks; ++ch) {
        const uint32_t t = M[(size_t)ch * nc + key];
        M[(size_t)ch * nc + key] = run;
        run += t;
    }
    counts[key] = run;
}

// exclusive scan of counts -> coff[nc+1] (single block)
__global__ void __launch_bounds__(1024) member_keyscan_kernel(const uint32_t* __restrict__ counts, uint32_t nc,
                                                              uint32_t* __restrict__ coff) {
    __shared__ uint32_t s_part[1024];
    const uint32_t tid = threadIdx.x;
    const uint32_t chunk = (nc + 1023) / 1024;
    const uint32_t lo = min(tid * chunk, nc), hi = min(lo + chunk, nc);
    uint32_t s = 0;
    for (uint32_t i = lo; i < hi; ++i) s += counts[i];
    s_part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; ++i) {
            const uint32_t t = s_part[i];
            s_part[i] = run;
            run += t;
        }
        coff[nc] = run;
    }
    __syncthreads();
    uint32_t run = s_part[tid];
    for (uint32_t i = lo; i < hi; ++i) {
        coff[i] = run;
        run += counts[i];
    }
}

__global__ void __launch_bounds__(32) member_rank_kernel(const uint32_t* __restrict__ assign, uint32_t n,
                                                         uint32_t chunk, uint32_t nc, uint32_t* __restrict__ M,
                                                         const uint32_t* __restrict__ coff,
                                                         uint32_t* __restrict__ members) {
    const uint32_t ch = blockIdx.x, lane = threadIdx.x;
    const uint64_t lo = (uint64_t)ch * chunk, hi = min((uint64_t)n, lo + chunk);
    uint32_t* row = M + (size_t)ch * nc;
    for (uint64_t v0 = lo; v0 < hi; v0 += 32) {
        const uint64_t v = v0 + lane;
        const bool ok = v < hi;
        const unsigned active = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            const uint32_t key = assign[v];
            const unsigned same = __match_any_sync(active, key);
            const uint32_t rank = __popc(same & ((1u << lane) - 1u));
            const uint32_t before = row[key];
            members[coff[key] + before + rank] = (uint32_t)v;
            __syncwarp(active);
            if (rank == 0) row[key] = before + __popc(same);
        }
        __syncwarp();
    }
}

// sums of each cluster's rows, added in ascending input row (ivf_flat_index.cpp:123-131)
__global__ void __launch_bounds__(128) cluster_sum_kernel(const float* __restrict__ x, uint32_t ldx,
                                                          const uint32_t* __restrict__ members,
                                                          const uint32_t* __restrict__ coff, uint32_t ld,
                                                          float* __restrict__ sums) {
    const uint32_t c = blockIdx.x;
    const uint32_t d4 = blockIdx.y * blockDim.x + threadIdx.x;
    if (d4 >= (ld >> 2)) return;
    const uint32_t lo = coff[c], hi = coff[c + 1];
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const uint32_t ldx4 = ldx >> 2;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t m = lo;
    for (; m + 4 <= hi; m += 4) {
        const float4 a = x4[(size_t)members[m] * ldx4 + d4];
        const float4 b = x4[(size_t)members[m + 1] * ldx4 + d4];
        const float4 cc = x4[(size_t)members[m + 2] * ldx4 + d4];
        const float4 d = x4[(size_t)members[m + 3] * ldx4 + d4];
        s.x = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s.x, a.x), b.x), cc.x), d.x);
        s.y = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s.y, a.y), b.y), cc.y), d.y);
        s.z = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s.z, a.z), b.z), cc.z), d.z);
        s.w = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s.w, a.w), b.w), cc.w), d.w);
    }
    for (; m < hi; ++m) {
        const float4 a = x4[(size_t)members[m] * ldx4 + d4];
        s.x = __fadd_rn(s.x, a.x);
        s.y = __fadd_rn(s.y, a.y);
        s.z = __fadd_rn(s.z, a.z);
        s.w = __fadd_rn(s.w, a.w);
    }
    reinterpret_cast<float4*>(sums)[(size_t)c * (ld >> 2) + d4] = s;
}

// centroid = sum / count where count > 0; an empty cluster keeps its centroid (:134-141)
__global__ void centroid_divide_kernel(const float* __restrict__ sums, const uint32_t* __restrict__ counts,
                                       uint32_t nc, uint32_t ld, float* __restrict__ centroids) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (uint64_t)nc * ld) return;
    const uint32_t c = (uint32_t)(i / ld);
    const uint32_t cnt = counts[c];
    if (cnt > 0) centroids[i] = __fdiv_rn(sums[i], __uint2float_rn(cnt));
}

// ----------------------------------------------------------------------- add

// owner == nullptr: every list counts; else only the lists this shard owns.  hist has nlist + 1 entries: the last
// one counts assignments that name no list (caller-supplied assignments, vdb_index_add_assigned)
__global__ void hist_kernel(const uint32_t* __restrict__ assign, uint64_t n, uint32_t nlist, uint32_t shard_rank,
                            const uint8_t* __restrict__ owner, uint32_t* __restrict__ hist) {
    const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t l = assign[v];
    if (l >= nlist) {
        atomicAdd(&hist[nlist], 1u);
        return;
    }
    if (!owner || owner[l] == shard_rank) atomicAdd(&hist[l], 1u);
}

// Shadow of one page row for the tensor-core screen of the list scan (ListTable::mirror_off / mirror_kind), written
// by the warp that just stored the fp32 row (src = that row, 16-byte aligned, ld floats): the values rounded to bf16,
// or quantised to int8 with the row's own scale (max |v_i| / 127), go to their place in the operand-tile image;
// err_out[0] = |v - shadow(v)| rounded up, err_out[page_rows] = the row scale (int8).  The differences are exact in
// fp32 (bf16) or one fused rounding each (int8: v_i - scale * n_i), the sum's own rounding (< 2^-14 relative) is
// inside the factor 1.0002.
__device__ __forceinline__ void mirror_write_row(uint8_t* mirror, uint32_t kind, uint32_t r, uint32_t ld,
                                                 const float4* __restrict__ src, float* err_out, uint32_t page_rows,
                                                 uint32_t lane) {
    float err = 0.f;
    if (kind == MIRROR_BF16) {
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            const float4 t = src[c];
            const __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
            uint2 bits;
            bits.x = *reinterpret_cast<const uint32_t*>(&lo);
            bits.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(mirror + mirror_elem_off(r, c * 4u, ld, 2)) = bits;
            const float dx = t.x - __low2float(lo), dy = t.y - __high2float(lo);
            const float dz = t.z - __low2float(hi), dw = t.w - __high2float(hi);
            err += fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
        }
    } else {
        float mx = 0.f;
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            const float4 t = src[c];
            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(t.x), fabsf(t.y))), fmaxf(fabsf(t.z), fabsf(t.w)));
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        const float scale = mx * (1.f / 127.f);
        const float inv = scale > 0.f ? 1.f / scale : 0.f;
        for (uint32_t c = lane; c < (ld >> 2); c += 32) {
            const float4 t = src[c];
            const float nx = fminf(fmaxf(rintf(t.x * inv), -127.f), 127.f), ny = fminf(fmaxf(rintf(t.y * inv), -127.f), 127.f);
            const float nz = fminf(fmaxf(rintf(t.z * inv), -127.f), 127.f), nw = fminf(fmaxf(rintf(t.w * inv), -127.f), 127.f);
            const uint32_t bits = ((uint32_t)(int)nx & 0xffu) | (((uint32_t)(int)ny & 0xffu) << 8) |
                                  (((uint32_t)(int)nz & 0xffu) << 16) | (((uint32_t)(int)nw & 0xffu) << 24);
            *reinterpret_cast<uint32_t*>(mirror + mirror_elem_off(r, c * 4u, ld, 1)) = bits;
            const float dx = fmaf(-scale, nx, t.x), dy = fmaf(-scale, ny, t.y);
            const float dz = fmaf(-scale, nz, t.z), dw = fmaf(-scale, nw, t.w);
            err += fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)));
        }
        if (lane == 0) err_out[page_rows] = scale;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) err += __shfl_xor_sync(0xffffffffu, err, o);
    if (lane == 0) err_out[0] = __fmul_ru(__fsqrt_ru(err), 1.0002f);
}

// one warp per new row: claim the next slot of its list, copy row + id into the page
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ x, uint32_t ldx,
                                                           const uint64_t* __restrict__ ids, uint64_t id_base,
                                                           uint64_t n, const uint32_t* __restrict__ assign,
                                                           const uint32_t* __restrict__ old_rows,
                                                           uint32_t* __restrict__ fill,
                                                           const uint32_t* __restrict__ page_off,
                                                           const uint64_t* __restrict__ page_vec,
                                                           const uint64_t* __restrict__ page_ids, uint32_t page_rows,
                                                           uint32_t ld, uint32_t nlist, uint32_t shard_rank,
                                                           const uint8_t* __restrict__ owner, uint32_t mirror_off,
                                                           uint32_t mirror_kind) {
    const uint64_t v = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (v >= n) return;
    const uint32_t l = assign[v];
    if (l >= nlist) return;  // no page was allocated for it (hist_kernel counted it as invalid)
    if (owner && owner[l] != shard_rank) return;
    uint32_t pos = 0;
    if (lane == 0) pos = old_rows[l] + atomicAdd(&fill[l], 1u);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    const uint32_t pg = page_off[l] + pos / page_rows, r = pos % page_rows;
    float4* dst = reinterpret_cast<float4*>(page_vec[pg]) + (size_t)r * (ld >> 2);
    const float4* src = reinterpret_cast<const float4*>(x + v * ldx);
    float nrm = 0.f;  // |row|^2 for the scan's dot-form screen (any fixed order: its rounding is inside the slack)
    for (uint32_t c = lane; c < (ld >> 2); c += 32) {
        const float4 t = src[c];
        dst[c] = t;
        nrm = fmaf(t.x, t.x, nrm);
        nrm = fmaf(t.y, t.y, nrm);
        nrm = fmaf(t.z, t.z, nrm);
        nrm = fmaf(t.w, t.w, nrm);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    uint64_t* idp = reinterpret_cast<uint64_t*>(page_ids[pg]);
    float* nb = reinterpret_cast<float*>(idp + page_rows);  // the norms follow the page's id block
    if (lane == 0) {
        idp[r] = ids ? ids[v] : id_base + v;
        nb[r] = nrm;
    }
    if (mirror_off)  // then |v - shadow(v)| and the row scales
        mirror_write_row(reinterpret_cast<uint8_t*>(page_vec[pg]) + mirror_off, mirror_kind, r, ld, src,
                         nb + page_rows + r, page_rows, lane);
}

// |row|^2 of rows [r0, r0 + count) of one page (rows copied in without the scatter kernel: vdb_index_append_list)
// (+ the low-precision shadow, its error norms and row scales when the index keeps one: mirror_off != 0)
__global__ void __launch_bounds__(256) page_norms_kernel(const float* __restrict__ rows, uint32_t ld,
                                                         float* __restrict__ norms, uint32_t r0, uint32_t count,
                                                         uint32_t mirror_off, uint32_t mirror_kind,
                                                         uint32_t page_rows) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= count) return;
    const float4* src = reinterpret_cast<const float4*>(rows + (size_t)(r0 + w) * ld);
    float nrm = 0.f;
    for (uint32_t c = lane; c < (ld >> 2); c += 32) {
        const float4 t = src[c];
        nrm = fmaf(t.x, t.x, nrm);
        nrm = fmaf(t.y, t.y, nrm);
        nrm = fmaf(t.z, t.z, nrm);
        nrm = fmaf(t.w, t.w, nrm);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) norms[r0 + w] = nrm;
    if (mirror_off)
        mirror_write_row(const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(rows)) + mirror_off, mirror_kind, r0 + w,
                         ld, src, norms + page_rows + r0 + w, page_rows, lane);
}

// [n][dim] (any stride) -> [n][ld] zero-padded
__global__ void pad_rows_kernel(const float* __restrict__ src, uint32_t lds, uint32_t dim, float* __restrict__ dst,
                                uint32_t ld, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * ld) return;
    const uint64_t r = i / ld;
    const uint32_t d = (uint32_t)(i - r * ld);
    dst[i] = d < dim ? src[r * lds + d] : 0.f;
}

}  // namespace

size_t rng_state_bytes() { return sizeof(DevRng); }

int32_t kmeans_assign_exact(const float* x, uint64_t n, uint32_t ldx, const float* c, uint32_t nc, uint32_t ldc,
                            uint32_t dim, int metric, uint32_t* assign, float* dist_out, cudaStream_t stream) {
    if (n == 0) return VDB_OK;
    VDB_REQUIRE(nc >= 1 && dim >= 1, "assign: empty centroid table");
    const uint64_t blocks = (n + TV - 1) / TV;
    VDB_REQUIRE(blocks < (1ull << 31), "assign: too many rows for one call");
    assign_exact_kernel<<<(uint32_t)blocks, 256, 0, stream>>>(x, n, ldx, c, nc, ldc, dim, metric, assign, dist_out,
                                                              nullptr);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

// A handful of rows (the tensor path's overflow list): the tiled kernel above would leave the GPU idle, so one row
// is spread over FEW_SPLIT blocks, one thread per centroid with the reference's sequential sum, and the block
// results meet in an atomicMin on (distance, centroid) keys -- strict '<' with the lowest index winning ties.
constexpr uint32_t FEW_SPLIT = 16;

__device__ __forceinline__ uint32_t ordered_bits(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256) assign_fewrows_kernel(const float* __restrict__ x, uint32_t ldx,
                                                             const uint32_t* __restrict__ row_index,
                                                             const float* __restrict__ c, uint32_t nc, uint32_t ldc,
                                                             uint32_t dim, int metric,
                                                             unsigned long long* __restrict__ keys) {
    extern __shared__ float srow[];
    const uint32_t r = blockIdx.x, tid = threadIdx.x;
    const float* xr = x + (size_t)row_index[r] * ldx;
    for (uint32_t d = tid; d < dim; d += 256) srow[d] = xr[d];
    __syncthreads();
    const uint32_t per = (nc + FEW_SPLIT - 1) / FEW_SPLIT;
    const uint32_t lo = blockIdx.y * per, hi = min(nc, lo + per);
    unsigned long long best = ~0ull;
    for (uint32_t cc = lo + tid; cc < hi; cc += 256) {
        const float* cv = c + (size_t)cc * ldc;
        float a = 0.f;
        if (metric == VDB_METRIC_L2) {
            for (uint32_t d = 0; d < dim; ++d) {
                const float diff = __fsub_rn(srow[d], __ldg(cv + d));
                a = __fadd_rn(a, __fmul_rn(diff, diff));
            }
        } else {
            for (uint32_t d = 0; d < dim; ++d) a = __fadd_rn(a, __fmul_rn(srow[d], __ldg(cv + d)));
            a = -a;
        }
        if (a == 0.f) a = 0.f;  // -0 and +0 compare equal in the reference; keep one encoding
        if (a < FLT_MAX) {
            const unsigned long long key = ((unsigned long long)ordered_bits(a) << 32) | cc;
            best = key < best ? key : best;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if ((tid & 31) == 0 && best != ~0ull) atomicMin(&keys[r], best);
}

__global__ void assign_fewrows_finish_kernel(const unsigned long long* __restrict__ keys,
                                             const uint32_t* __restrict__ row_index, uint32_t m,
                                             uint32_t* __restrict__ assign) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    // nothing below FLT_MAX: min_dist stays at its initial value and best_list at 0 (:266-267)
    assign[row_index[r]] = keys[r] == ~0ull ? 0u : (uint32_t)(keys[r] & 0xffffffffu);
}

// the same for the rows listed in row_index[0..m): results go to assign[row_index[i]]
int32_t kmeans_assign_exact_rows(const float* x, const uint32_t* row_index, uint64_t m, uint32_t ldx, const float* c,
                                 uint32_t nc, uint32_t ldc, uint32_t dim, int metric, uint32_t* assign,
                                 unsigned long long* keys, cudaStream_t stream) {
    if (m == 0) return VDB_OK;
    if (keys && m <= FEWROWS_MAX && nc >= 4096) {  // keys: FEWROWS_MAX words of scratch
        VDB_CUDA_TRY(cudaMemsetAsync(keys, 0xff, m * 8, stream));
        assign_fewrows_kernel<<<dim3((uint32_t)m, FEW_SPLIT), 256, dim * 4, stream>>>(x, ldx, row_index, c, nc, ldc, dim,
                                                                                      metric, keys);
        assign_fewrows_finish_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, stream>>>(keys, row_index, (uint32_t)m,
                                                                                      assign);
        VDB_CUDA_TRY(cudaGetLastError());
        return VDB_OK;
    }
    assign_exact_kernel<<<(uint32_t)((m + TV - 1) / TV), 256, 0, stream>>>(x, m, ldx, c, nc, ldc, dim, metric, assign,
                                                                            nullptr, row_index);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t kmeanspp_init(const float* x, uint32_t n, uint32_t ldx, uint32_t ld, float* centroids, KMeansScratch& sc,
                      float* mind, uint64_t n_fill, cudaStream_t stream) {
    seed_init_kernel<<<1, 256, 0, stream>>>((DevRng*)sc.rng, x, n, ldx, ld, centroids, sc.picked);
    fill_f32_kernel<<<(uint32_t)((n_fill + 255) / 256), 256, 0, stream>>>(mind, n_fill, FLT_MAX);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t kmeanspp_dist_plan(const float* x, uint32_t n, uint32_t ldx, uint32_t dim, uint32_t ld, SeedDistPlan* plan) {
    plan->x = x; plan->n = n; plan->ldx = ldx; plan->dim = dim;
    // rows by TMA when the driver offers tensor maps and the rows are 16-byte aligned (always, for staged rows)
    plan->tma = n > 0 && tc::encode_tiled() != nullptr && ldx % 4 == 0 && ((uintptr_t)x & 15) == 0 && ldx <= 2048 && ld == ldx;
    if (plan->tma) {
        static_assert(sizeof(plan->map) >= sizeof(CUtensorMap), "tensor map storage");
        VDB_TRY(tc::make_map(reinterpret_cast<CUtensorMap*>(plan->map), x, n, ldx, ldx, SD_ROWS));
        static bool conf[16] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 16 && !conf[dev]) {
            VDB_CUDA_TRY(cudaFuncSetAttribute(seed_dist_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SD_SMEM));
            conf[dev] = true;
        }
    }
    return VDB_OK;
}

int32_t kmeanspp_dist(const SeedDistPlan& plan, const float* cnew, const float* mind_in, const PeerF32& mind_out,
                      cudaStream_t stream) {
    if (plan.n == 0) return VDB_OK;
    if (plan.tma)
        seed_dist_tma_kernel<<<(plan.n + SD_ROWS - 1) / SD_ROWS, SD_ROWS + 32, SD_SMEM, stream>>>(
            *reinterpret_cast<const CUtensorMap*>(plan.map), plan.n, plan.ldx, cnew, mind_in, mind_out);
    else
        seed_dist_kernel<<<(plan.n + 127) / 128, 128, 0, stream>>>(plan.x, plan.n, plan.ldx, plan.dim, cnew, mind_in,
                                                                     mind_out);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t kmeanspp_sample(const float* x, uint32_t n, uint32_t ldx, uint32_t ld, const float* mind, float* centroids,
                        uint32_t c, KMeansScratch& sc, SeedSampler sampler, cudaStream_t stream) {
    if (sampler == SeedSampler::Sequential)
        seed_sample_kernel<<<1, 32, 0, stream>>>((DevRng*)sc.rng, x, n, ldx, ld, mind, sc.ckpt, centroids, c, sc.picked);
    else if (sampler == SeedSampler::ExactParallel)
        seed_sample_par_kernel<<<PS_CTAS, PS_THREADS, 0, stream>>>((DevRng*)sc.rng, x, n, ldx, ld, mind, centroids, c,
                                                                     sc.picked);
    else
        seed_sample_fast_kernel<<<1, 1024, 0, stream>>>((DevRng*)sc.rng, x, n, ldx, ld, mind, centroids, c, sc.picked);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t kmeanspp_seed(const float* x, uint32_t n, uint32_t ldx, uint32_t dim, uint32_t ld, uint32_t nlist,
                      float* centroids, KMeansScratch& sc, SeedSampler sampler, cudaStream_t stream) {
    VDB_TRY(kmeanspp_init(x, n, ldx, ld, centroids, sc, sc.mind, n, stream));
    SeedDistPlan plan;
    VDB_TRY(kmeanspp_dist_plan(x, n, ldx, dim, ld, &plan));
    PeerF32 out{};
    out.p[0] = sc.mind;
    out.n = 1;
    for (uint32_t c = 1; c < nlist; ++c) {
        VDB_TRY(kmeanspp_dist(plan, centroids + (size_t)(c - 1) * ld, sc.mind, out, stream));
        VDB_TRY(kmeanspp_sample(x, n, ldx, ld, sc.mind, centroids, c, sc, sampler, stream));
    }
    return VDB_OK;
}

int32_t kmeans_members(const uint32_t* assign, uint32_t n, uint32_t nc, KMeansScratch& sc, cudaStream_t stream) {
    const uint32_t nchunks = sc.nchunks, chunk = sc.chunk;
    VDB_CUDA_TRY(cudaMemsetAsync(sc.M, 0, (size_t)nchunks * nc * 4, stream));
    member_hist_kernel<<<nchunks, 32, 0, stream>>>(assign, n, chunk, nc, sc.M);
    member_colscan_kernel<<<(nc + 255) / 256, 256, 0, stream>>>(sc.M, nchunks, nc, sc.counts);
    member_keyscan_kernel<<<1, 1024, 0, stream>>>(sc.counts, nc, sc.coff);
    member_rank_kernel<<<nchunks, 32, 0, stream>>>(assign, n, chunk, nc, sc.M, sc.coff, sc.members);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t kmeans_update_exact(const float* x, uint32_t n, uint32_t ldx, const uint32_t* assign, uint32_t nc,
                            uint32_t ld, float* centroids, KMeansScratch& sc, cudaStream_t stream) {
    VDB_TRY(kmeans_members(assign, n, nc, sc, stream));
    dim3 grid(nc, ((ld >> 2) + 127) / 128);
    cluster_sum_kernel<<<grid, 128, 0, stream>>>(x, ldx, sc.members, sc.coff, ld, sc.sums);
    const uint64_t tot = (uint64_t)nc * ld;
    centroid_divide_kernel<<<(uint32_t)((tot + 255) / 256), 256, 0, stream>>>(sc.sums, sc.counts, nc, ld, centroids);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

// data-parallel training: this rank sums and divides only clusters [c_lo, c_hi) -- each cluster's rows are still
// added in input order by one rank, so the centroids equal the single-GPU ones bit for bit
int32_t kmeans_update_exact_range(const float* x, uint32_t n, uint32_t ldx, const uint32_t* assign, uint32_t nc,
                                  uint32_t ld, float* centroids, KMeansScratch& sc, uint32_t c_lo, uint32_t c_hi,
                                  cudaStream_t stream) {
    VDB_TRY(kmeans_members(assign, n, nc, sc, stream));
    if (c_hi <= c_lo) return VDB_OK;
    const uint32_t m = c_hi - c_lo;
    dim3 grid(m, ((ld >> 2) + 127) / 128);
    cluster_sum_kernel<<<grid, 128, 0, stream>>>(x, ldx, sc.members, sc.coff + c_lo, ld, sc.sums + (size_t)c_lo * ld);
    const uint64_t tot = (uint64_t)m * ld;
    centroid_divide_kernel<<<(uint32_t)((tot + 255) / 256), 256, 0, stream>>>(
        sc.sums + (size_t)c_lo * ld, sc.counts + c_lo, m, ld, centroids + (size_t)c_lo * ld);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t kmeans_cluster_sums(const float* x, uint32_t n, uint32_t ldx, const uint32_t* assign, uint32_t nc,
                            uint32_t ld, float* sums, uint32_t* counts, KMeansScratch& sc, cudaStream_t stream) {
    VDB_TRY(kmeans_members(assign, n, nc, sc, stream));
    dim3 grid(nc, ((ld >> 2) + 127) / 128);
    cluster_sum_kernel<<<grid, 128, 0, stream>>>(x, ldx, sc.members, sc.coff, ld, sums);
    VDB_CUDA_TRY(cudaMemcpyAsync(counts, sc.counts, (size_t)nc * 4, cudaMemcpyDeviceToDevice, stream));
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t kmeans_divide(const float* sums, const uint32_t* counts, uint32_t nc, uint32_t ld, float* centroids,
                      cudaStream_t stream) {
    const uint64_t tot = (uint64_t)nc * ld;
    centroid_divide_kernel<<<(uint32_t)((tot + 255) / 256), 256, 0, stream>>>(sums, counts, nc, ld, centroids);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t KMeansScratch::reserve(uint32_t n, uint32_t nc, uint32_t ld) {
    release();
    nchunks = std::min<uint32_t>(1024, (n + 1023) / 1024);
    if (nchunks == 0) nchunks = 1;
    chunk = ((n + nchunks - 1) / nchunks + 31) / 32 * 32;
    if (chunk == 0) chunk = 32;
    VDB_CUDA_TRY(cudaMalloc(&rng, rng_state_bytes()));
    VDB_CUDA_TRY(cudaMalloc(&mind, (size_t)std::max(n, 1u) * 4 * mind_copies));
    VDB_CUDA_TRY(cudaMalloc(&ckpt, ((size_t)(n + SEQ_CHUNK - 1) / SEQ_CHUNK + 2) * 4));
    VDB_CUDA_TRY(cudaMalloc(&picked, (size_t)nc * 4));
    VDB_CUDA_TRY(cudaMalloc(&M, (size_t)nchunks * nc * 4));
    VDB_CUDA_TRY(cudaMalloc(&counts, (size_t)nc * 4));
    VDB_CUDA_TRY(cudaMalloc(&coff, (size_t)(nc + 1) * 4));
    VDB_CUDA_TRY(cudaMalloc(&members, (size_t)std::max(n, 1u) * 4));
    VDB_CUDA_TRY(cudaMalloc(&sums, (size_t)nc * ld * 4));
    VDB_CUDA_TRY(cudaMalloc(&assign, (size_t)std::max(n, 1u) * 4));
    return VDB_OK;
}

void KMeansScratch::release() {
    cudaFree(rng); cudaFree(mind); cudaFree(ckpt); cudaFree(picked); cudaFree(M);
    cudaFree(counts); cudaFree(coff); cudaFree(members); cudaFree(sums); cudaFree(assign);
    rng = nullptr; mind = nullptr; ckpt = nullptr; picked = nullptr; M = nullptr; counts = nullptr;
    coff = nullptr; members = nullptr; sums = nullptr; assign = nullptr;
}

int32_t launch_hist(const uint32_t* assign, uint64_t n, uint32_t nlist, uint32_t shard_rank, const uint8_t* owner,
                    uint32_t* hist, cudaStream_t stream) {
    if (n == 0) return VDB_OK;
    hist_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, stream>>>(assign, n, nlist, shard_rank, owner, hist);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t launch_scatter_rows(const float* x, uint32_t ldx, const uint64_t* ids, uint64_t id_base, uint64_t n,
                            const uint32_t* assign, const uint32_t* old_rows, uint32_t* fill,
                            const uint32_t* page_off, const uint64_t* page_vec, const uint64_t* page_ids,
                            uint32_t page_rows, uint32_t ld, uint32_t nlist, uint32_t shard_rank,
                            const uint8_t* owner, uint32_t mirror_off, uint32_t mirror_kind, cudaStream_t stream) {
    if (n == 0) return VDB_OK;
    const uint64_t blocks = (n * 32 + 255) / 256;
    scatter_rows_kernel<<<(uint32_t)blocks, 256, 0, stream>>>(x, ldx, ids, id_base, n, assign, old_rows, fill,
                                                              page_off, page_vec, page_ids, page_rows, ld, nlist,
                                                              shard_rank, owner, mirror_off, mirror_kind);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t launch_page_norms(const float* rows, uint32_t ld, float* norms, uint32_t r0, uint32_t count,
                          uint32_t mirror_off, uint32_t mirror_kind, uint32_t page_rows, cudaStream_t stream) {
    if (count == 0) return VDB_OK;
    page_norms_kernel<<<(count * 32 + 255) / 256, 256, 0, stream>>>(rows, ld, norms, r0, count, mirror_off, mirror_kind,
                                                                    page_rows);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

int32_t launch_pad_rows(const float* src, uint32_t lds, uint32_t dim, float* dst, uint32_t ld, uint64_t n,
                        cudaStream_t stream) {
    if (n == 0) return VDB_OK;
    const uint64_t tot = n * ld;
    pad_rows_kernel<<<(uint32_t)((tot + 255) / 256), 256, 0, stream>>>(src, lds, dim, dst, ld, n);
    VDB_CUDA_TRY(cudaGetLastError());
    return VDB_OK;
}

}  // namespace vdb
