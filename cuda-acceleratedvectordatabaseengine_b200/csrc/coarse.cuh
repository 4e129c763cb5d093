// Host-visible interface of the tensor-core coarse selection (coarse.cu).
#pragma once
#include "common.cuh"

namespace vdb {

// true when the driver exposes cuTensorMapEncodeTiled and the shapes fit the select kernel
bool coarse_tensor_supported(uint32_t N, uint32_t ld, uint32_t np);
size_t coarse_select_smem(uint32_t N);
// |c|^2 of every centroid and the bit pattern of their maximum
int32_t centroid_norms(const float* centroids, uint32_t n, uint32_t ld, float* norms, uint32_t* max_bits,
                       cudaStream_t stream);
// out[m][n] = sum_k A[m][k] * B[n][k] with TF32 tensor-core inputs, fp32 accumulation (tcgen05 + TMA + TMEM)
int32_t score_gemm(const float* A, uint32_t M, uint32_t lda, const float* B, uint32_t N, uint32_t ldb, uint32_t K,
                   float* out, uint32_t ldo, cudaStream_t stream);
// per query: candidates within the TF32 error bound of the np-th score, exact fp32 re-check, best np by (dist, id)
int32_t coarse_select(const float* dots, uint32_t ldd, const float* queries, uint32_t nq, const float* centroids,
                      const float* cnorm, const uint32_t* cmax_bits, uint32_t N, uint32_t ld, uint32_t np, int metric,
                      uint32_t* probes, float* out_d, uint32_t* cand_count, cudaStream_t stream);

// nprobe beyond the select kernel's pool (> 2047): exact fp32 scores of every centroid + a full sort per query
bool coarse_wide_supported(uint32_t N, uint32_t ld);
int32_t coarse_select_wide(const float* queries, uint32_t nq, const float* centroids, uint32_t N, uint32_t ld,
                           uint32_t np, int metric, uint32_t* probes, float* out_d, cudaStream_t stream);

}  // namespace vdb
