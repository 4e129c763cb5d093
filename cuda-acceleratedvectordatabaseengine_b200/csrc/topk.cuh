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

}  // namespace vdb
