// Tensor-core brute-force search (bruteforce_tc.cu) -- host entry points.
#pragma once
#include "common.cuh"

namespace vdb {

constexpr uint64_t BF_LEVEL_RATIO = 8;  // a nested sample level holds at most 8x the rows of the one before

struct BruteTcScratch {
    float *vnorm = nullptr, *vnorm2 = nullptr, *qnorm = nullptr, *qnorm2 = nullptr, *colA = nullptr, *colB = nullptr;
    uint32_t *cand_rows = nullptr, *ccount = nullptr, *overflow = nullptr;
    float* cand_ub = nullptr;
    void* block = nullptr;  // one stream-ordered allocation carved into the arrays above
    cudaStream_t stream = nullptr;
    void release();
};

bool bruteforce_tensor_supported(uint64_t n, uint32_t nq, uint32_t ld, uint32_t k);

int32_t bruteforce_tensor(const float* db, uint64_t n, uint32_t ld, const uint64_t* ids_flat, const float* queries,
                          uint32_t nq, uint32_t k, int metric, float* out_d, uint64_t* out_i, BruteTcScratch& sc,
                          int* overflowed, cudaStream_t stream);

}  // namespace vdb
