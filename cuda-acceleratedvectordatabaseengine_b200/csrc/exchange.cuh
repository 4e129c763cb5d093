// Peer mailboxes of the cross-GPU top-k exchange (exchange.cu), shared with the merge kernel (scan.cu) so that a
// shard's merge can store its result straight into the peers' HBM instead of handing it to a separate launch.
#pragma once
#include "common.cuh"

namespace vdb {

constexpr uint32_t EX_MAX_WORLD = 16;

struct Mailbox {  // device pointers into ONE rank's mailbox allocation
    uint32_t* flags;  // [2][world][max_nq]
    float* dist;      // [2][world][max_nq * max_k]
    uint64_t* ids;    // [2][world][max_nq * max_k]
};

// Where a rank's local [nq][k] block goes: slot [epoch & 1][rank] of the mailbox of every rank in [dst_lo, dst_hi).
struct PublishTarget {
    Mailbox box[EX_MAX_WORLD];  // box[r] = rank r's mailbox as mapped into this process
    uint32_t rank, world, max_nq, max_k, epoch;
    uint32_t dst_lo, dst_hi;    // all ranks [0, world), or just the root of a single-process sharded index
    uint32_t enabled;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All threads of the CTA: query q's k entries (src_d/src_i hold `got` valid ones, the rest is padding) -> the
// targets' mailboxes, then one release flag per target.  `k` is the row stride inside the mailbox slot.
__device__ __forceinline__ void publish_query(const PublishTarget& t, uint32_t q, uint32_t k, const float* src_d,
                                              const uint64_t* src_i, uint32_t got, uint32_t nthreads) {
    const uint32_t tid = threadIdx.x, half = t.epoch & 1u, nd = t.dst_hi - t.dst_lo;
    const size_t slot_stride = (size_t)t.max_nq * t.max_k;
    const size_t base = ((size_t)half * t.world + t.rank) * slot_stride + (size_t)q * k;
    for (uint32_t e = tid; e < nd * k; e += nthreads) {
        const uint32_t r = t.dst_lo + e / k, j = e % k;
        t.box[r].dist[base + j] = j < got ? src_d[j] : FLT_MAX;
        t.box[r].ids[base + j] = j < got ? src_i[j] : ID_PAD;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < nd)
        st_release_sys(&t.box[t.dst_lo + tid].flags[((size_t)half * t.world + t.rank) * t.max_nq + q], t.epoch);
}

// host side (exchange.cu): fill `t` for the next collective step of `ex` (bumps its call counter)
}  // namespace vdb

struct vdb_exchange;
namespace vdb {
int32_t exchange_begin_publish(vdb_exchange* ex, uint32_t nq, uint32_t k, PublishTarget* t);
// wait for the peers' blocks of the step begun last and merge them into out_d/out_i (device), on `s`
int32_t exchange_collect(vdb_exchange* ex, float* out_d, uint64_t* out_i, cudaStream_t s);
}  // namespace vdb
