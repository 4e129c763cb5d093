// (distance, id) selection primitives shared by the scan and merge kernels.
// Ordering everywhere is std::pair<float,uint64_t>::operator< -- ascending
// distance, ties by ascending id (ivf_flat_index.cpp:368-370,493).
#pragma once
#include "common.cuh"

namespace vdb {

__device__ __forceinline__ bool pair_less(float da, uint64_t ia, float db, uint64_t ib) {
    return (da < db) || (da == db && ia < ib);
}

// In-place ascending bitonic sort of n (power of two) pairs held in shared
// memory by `nthreads` cooperating threads; `sync()` separates the steps
// (__syncwarp for one warp, a named barrier or __syncthreads for more).
template <typename Sync>
__device__ __forceinline__ void bitonic_sort_pairs(float* d, uint64_t* id, uint32_t n, uint32_t tid,
                                                   uint32_t nthreads, Sync sync) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < (n >> 1); t += nthreads) {
                uint32_t i = ((t / stride) * (stride << 1)) + (t % stride);
                uint32_t j = i + stride;
                bool up = ((i & size) == 0);
                float di = d[i], dj = d[j];
                uint64_t ii = id[i], ij = id[j];
                bool swap = up ? pair_less(dj, ij, di, ii) : pair_less(di, ii, dj, ij);
                if (swap) {
                    d[i] = dj; d[j] = di;
                    id[i] = ij; id[j] = ii;
                }
            }
            sync();
        }
    }
}

__device__ __forceinline__ uint32_t dev_next_pow2(uint32_t v) {
    return v <= 1 ? 1u : (1u << (32 - __clz(v - 1)));
}


// ---- block-wide (256 threads) candidate pool: push with an admission threshold, compact to the best k ----

constexpr int MERGE_THREADS = 256;

struct MergePool {
    float* d;
    uint64_t* id;
    uint32_t* cnt;
    float* thr;
};

__device__ __forceinline__ void pool_push_block(const MergePool& pl, uint32_t P, const float* src_d,
                                                const uint64_t* src_i, uint32_t n) {
    // all threads; caller guarantees cnt + n <= P
    const float thr = *pl.thr;
    for (uint32_t i = threadIdx.x; i < n; i += MERGE_THREADS) {
        float d = src_d[i];
        uint64_t id = src_i[i];
        if (id == ID_PAD && d == FLT_MAX) continue;  // padding of a short partial
        if (d <= thr) {
            uint32_t pos = atomicAdd(pl.cnt, 1u);
            if (pos < P) {
                pl.d[pos] = d;
                pl.id[pos] = id;
            }
        }
    }
    __syncthreads();
}

// sort the pool; optionally drop later occurrences of an id (merge_results
// keeps the first = best); keep k; refresh the admission threshold.
__device__ __forceinline__ void pool_compact_block(const MergePool& pl, uint32_t P, uint32_t k, bool dedup,
                                                   float* tmp_d, uint64_t* tmp_i, uint32_t* s_scan) {
    const uint32_t tid = threadIdx.x;
    uint32_t c = min(*pl.cnt, P);
    const uint32_t n2 = dev_next_pow2(max(c, 1u));
    for (uint32_t i = c + tid; i < n2; i += MERGE_THREADS) {
        pl.d[i] = FLT_MAX;
        pl.id[i] = ID_PAD;
    }
    __syncthreads();
    bitonic_sort_pairs(pl.d, pl.id, n2, tid, MERGE_THREADS, [] { __syncthreads(); });
    bool may_dup = dedup && c > 1;
    if (may_dup && c <= P / 2) {
        // cheap screen: insert the ids into an open-addressing table (the scratch id array); only if some id
        // really occurs twice is the quadratic stable de-duplication below needed
        __shared__ uint32_t s_dup;
        if (tid == 0) s_dup = 0;
        for (uint32_t i = tid; i < P; i += MERGE_THREADS) tmp_i[i] = ID_PAD;
        __syncthreads();
        for (uint32_t i = tid; i < c; i += MERGE_THREADS) {
            const uint64_t id = pl.id[i];
            uint32_t h = (uint32_t)((id * 0x9E3779B97F4A7C15ull) >> 40) & (P - 1);
            for (uint32_t probe = 0; probe < P; ++probe) {
                const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(&tmp_i[h]),
                                                         (unsigned long long)ID_PAD, (unsigned long long)id);
                if (old == ID_PAD) break;
                if (old == id) {
                    s_dup = 1;
                    break;
                }
                h = (h + 1) & (P - 1);
            }
        }
        __syncthreads();
        may_dup = s_dup != 0;
        __syncthreads();
    }
    if (may_dup) {
        // keep[i] = no earlier entry carries the same id; stable compaction through tmp
        const uint32_t per = (c + MERGE_THREADS - 1) / MERGE_THREADS;
        const uint32_t lo = min(tid * per, c), hi = min(lo + per, c);
        uint32_t kept = 0;
        for (uint32_t i = lo; i < hi; ++i) {
            const uint64_t id = pl.id[i];
            bool dup = false;
            for (uint32_t j = 0; j < i; ++j)
                if (pl.id[j] == id) {
                    dup = true;
                    break;
                }
            if (!dup) ++kept;
        }
        // block exclusive scan of kept (MERGE_THREADS = 256 threads)
        const uint32_t lane = tid & 31, w = tid >> 5;
        uint32_t x = kept;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_scan[w] = x;
        __syncthreads();
        if (tid == 0) {
            uint32_t run = 0;
            for (int i = 0; i < MERGE_THREADS / 32; ++i) {
                uint32_t t = s_scan[i];
                s_scan[i] = run;
                run += t;
            }
            s_scan[MERGE_THREADS / 32] = run;
        }
        __syncthreads();
        uint32_t pos = s_scan[w] + x - kept;
        const uint32_t total = s_scan[MERGE_THREADS / 32];
        for (uint32_t i = lo; i < hi; ++i) {
            const uint64_t id = pl.id[i];
            bool dup = false;
            for (uint32_t j = 0; j < i; ++j)
                if (pl.id[j] == id) {
                    dup = true;
                    break;
                }
            if (!dup) {
                tmp_d[pos] = pl.d[i];
                tmp_i[pos] = id;
                ++pos;
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < total; i += MERGE_THREADS) {
            pl.d[i] = tmp_d[i];
            pl.id[i] = tmp_i[i];
        }
        c = total;
        __syncthreads();
    }
    const uint32_t nc = min(c, k);
    if (tid == 0) {
        *pl.cnt = nc;
        *pl.thr = (nc >= k) ? pl.d[k - 1] : INFINITY;
    }
    __syncthreads();
}


}  // namespace vdb
