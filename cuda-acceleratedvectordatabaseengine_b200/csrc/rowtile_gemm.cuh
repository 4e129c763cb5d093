// Persistent row-tile GEMM on the 5th-generation tensor cores, shared by the assignment (assign_tc.cu) and the
// brute-force search (bruteforce_tc.cu): D = A * B^T with TF32 inputs and fp32 accumulation, where a CTA owns
// 128 rows of A and walks ALL 128-column tiles of B.  Operands arrive by 2-D TMA (128-byte swizzle) through a
// 4-stage mbarrier ring; tcgen05.mma accumulates into one of two TMEM buffers while the four epilogue warps drain
// the other (thread = row).  The n x m product is never materialised: the Epilogue functor sees every (row,
// column, dot) once and keeps whatever per-row state it needs in registers.
//
// When there are fewer row tiles than SMs the column tiles of one row tile can be split over n_split CTAs (the
// functor then sees several begin/end pairs per row, one per column range).
//
//   struct Epi { uint32_t M, N, num_kb, n_split; static constexpr uint32_t WARP_SMEM;  struct State;
//                begin(State&, row, warp_smem)          all lanes (row clamped to 0 when out of range)
//                consume_chunk(State&, row, col0, acc)  rows < M only: 32 accumulators of columns col0 .. col0+31
//                                                       (columns >= N hold zeros -- the functor masks them)
//                chunk_end(State&, warp_smem)           all 32 lanes converged, after every 32 columns
//                end(State&, row, valid, warp_smem) }   all 32 lanes converged, after the last column tile
// WARP_SMEM bytes of shared memory are private to each epilogue warp (e.g. to batch global atomics).
#pragma once
#include "tc_common.cuh"

namespace vdb {
namespace tc {

constexpr int AM = 128;        // rows per CTA tile (UMMA M)
constexpr int ASTAGES = 4;
constexpr int ATHREADS = 192;  // warps 0-3 epilogue, 4 TMA producer, 5 MMA issuer + TMEM owner

// AN = columns per tile (UMMA N): 128, or 256 (all 512 TMEM columns, 25 % less L2->SMEM traffic per MAC)
template <typename Epi, int AN>
constexpr uint32_t rowtile_smem() {
    return ASTAGES * (AM * GK * 4 + AN * GK * 4) + (2 * ASTAGES + 4) * 8 + 16 + 4 * Epi::WARP_SMEM + 1024;
}

template <typename Epi, int AN>
__global__ void __launch_bounds__(ATHREADS, 1)
rowtile_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_c, const Epi p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr uint32_t A_BYTES = AM * GK * 4, B_BYTES = AN * GK * 4;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ASTAGES * A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sb + ASTAGES * B_BYTES);
    uint64_t* empty = full + ASTAGES;
    uint64_t* acc_full = empty + ASTAGES;  // [2]
    uint64_t* acc_empty = acc_full + 2;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint8_t* epi_smem = reinterpret_cast<uint8_t*>(tmem_slot + 4);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t m_tiles = (p.M + AM - 1) / AM, n_tiles = (p.N + AN - 1) / AN;
    const uint32_t items = m_tiles * p.n_split;  // host: 1 <= n_split <= n_tiles

    if (threadIdx.x == 0) {
        for (int i = 0; i < ASTAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_empty[i], 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 5) {  // two accumulator buffers of AN fp32 columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(2 * AN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4 && lane == 0) {
        uint32_t s = 0, ph = 0;
        for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
            const uint32_t mt = item / p.n_split, ns = item % p.n_split;
            for (uint32_t nt = ns * n_tiles / p.n_split; nt < (ns + 1) * n_tiles / p.n_split; ++nt)
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], A_BYTES + B_BYTES);
                    tma_load_2d(sa + s * A_BYTES, &map_x, (int32_t)(kb * GK), (int32_t)(mt * AM), &full[s]);
                    tma_load_2d(sb + s * B_BYTES, &map_c, (int32_t)(kb * GK), (int32_t)(nt * AN), &full[s]);
                    if (++s == ASTAGES) {
                        s = 0;
                        ph ^= 1;
                    }
                }
        }
    } else if (warp == 5 && lane == 0) {
        constexpr uint32_t idesc = umma_idesc_tf32(AM, AN);
        uint32_t s = 0, ph = 0, buf = 0, bph = 0;
        for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
            const uint32_t ns = item % p.n_split;
            for (uint32_t nt = ns * n_tiles / p.n_split; nt < (ns + 1) * n_tiles / p.n_split; ++nt) {
                mbar_wait(&acc_empty[buf], bph ^ 1);  // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * AN;
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint64_t da = umma_desc_sw128(sa + s * A_BYTES), db = umma_desc_sw128(sb + s * B_BYTES);
#pragma unroll
                    for (uint32_t k = 0; k < GK / 8; ++k)
                        umma_tf32(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit(&empty[s]);
                    if (++s == ASTAGES) {
                        s = 0;
                        ph ^= 1;
                    }
                }
                umma_commit(&acc_full[buf]);
                if (++buf == 2) {
                    buf = 0;
                    bph ^= 1;
                }
            }
        }
    } else if (warp < 4) {
        uint32_t buf = 0, bph = 0;
        for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
            const uint32_t mt = item / p.n_split, ns = item % p.n_split;
            const uint32_t row = mt * AM + warp * 32 + lane;
            const bool valid = row < p.M;
            uint8_t* wsm = epi_smem + warp * Epi::WARP_SMEM;
            typename Epi::State st;
            p.begin(st, valid ? row : 0, wsm);
            for (uint32_t nt = ns * n_tiles / p.n_split; nt < (ns + 1) * n_tiles / p.n_split; ++nt) {
                mbar_wait(&acc_full[buf], bph);
                tc_fence_after();
#pragma unroll 1
                for (uint32_t c0 = 0; c0 < (uint32_t)AN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem_base + ((warp * 32u) << 16) + buf * AN + c0, r);
                    const uint32_t nb = nt * AN + c0;
                    if (valid) p.consume_chunk(st, row, nb, r);
                    __syncwarp();
                    p.chunk_end(st, wsm);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
                if (++buf == 2) {
                    buf = 0;
                    bph ^= 1;
                }
            }
            __syncwarp();
            p.end(st, row, valid, wsm);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * AN) : "memory");
    }
}


}  // namespace tc
}  // namespace vdb
