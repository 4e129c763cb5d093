// Host-visible interface of the training / assignment kernels (kmeans.cu).
#pragma once
#include <algorithm>
#include "common.cuh"

namespace vdb {

// the same array on every rank of a data-parallel run (peer-mapped pointers; n = 1 on a single GPU)
struct PeerF32 {
    float* p[16];
    uint32_t n;
};

// rows [n][ldx] whose distance to the newest seed one rank folds into the running minimum
struct SeedDistPlan {
    const float* x = nullptr;
    uint32_t n = 0, ldx = 0, dim = 0;
    bool tma = false;
    alignas(64) unsigned char map[128];  // CUtensorMap
};

struct KMeansScratch {
    void* rng = nullptr;        // device std::mt19937 state
    uint32_t mind_copies = 1;   // 2 for data-parallel training (the ranks write the next copy while the current is read)
    float* mind = nullptr;      // [mind_copies][n] running min squared distance to the chosen seeds
    float* ckpt = nullptr;      // sequential-sum checkpoints
    uint32_t* picked = nullptr; // [nlist] training row chosen for each seed
    uint32_t* M = nullptr;      // [nchunks][nlist] counting-sort matrix
    uint32_t* counts = nullptr; // [nlist]
    uint32_t* coff = nullptr;   // [nlist+1]
    uint32_t* members = nullptr;// [n] rows bucketed by cluster, input order inside a cluster
    float* sums = nullptr;      // [nlist][ld]
    uint32_t* assign = nullptr; // [n]
    uint32_t nchunks = 0, chunk = 0;
    int32_t reserve(uint32_t n, uint32_t nc, uint32_t ld);
    void release();
};

int32_t kmeans_assign_exact(const float* x, uint64_t n, uint32_t ldx, const float* c, uint32_t nc, uint32_t ldc,
                            uint32_t dim, int metric, uint32_t* assign, float* dist_out, cudaStream_t stream);
constexpr uint64_t FEWROWS_MAX = 2048;  // up to this many listed rows take the row-parallel kernel
int32_t kmeans_assign_exact_rows(const float* x, const uint32_t* row_index, uint64_t m, uint32_t ldx, const float* c,
                                 uint32_t nc, uint32_t ldc, uint32_t dim, int metric, uint32_t* assign,
                                 unsigned long long* keys, cudaStream_t stream);

// tensor-core assignment (assign_tc.cu): bit-identical to kmeans_assign_exact, O(n nlist dim) on tcgen05
struct AssignTcScratch {
    float* xnorm = nullptr;
    float* cnorm2 = nullptr;
    uint32_t* cand_idx = nullptr;
    uint32_t* cand_cnt = nullptr;
    uint32_t* overflow_rows = nullptr;
    uint32_t* overflow_count = nullptr;
    unsigned long long* fewrow_keys = nullptr;  // [FEWROWS_MAX]
    uint32_t* h_overflow = nullptr;
    uint64_t cap_n = 0;
    uint32_t cap_nc = 0;
    uint32_t last_overflow = 0;
    int32_t reserve(uint64_t n, uint32_t nc);
    void release();
};
bool assign_tensor_supported(uint32_t nc, uint32_t ld);
int32_t kmeans_assign_tensor(const float* x, uint64_t n, uint32_t ldx, const float* c, uint32_t nc, uint32_t ldc,
                             uint32_t dim, int metric, uint32_t* assign, AssignTcScratch& sc, cudaStream_t stream);

// D^2 sampling sums: Sequential = the reference's fp32 chain on one lane (the literal restatement);
// ExactParallel = the same sums bit for bit from a block-wide scan of parity-dependent integer maps (default);
// Fast = double-precision partial sums (same RNG stream, a pick can differ within rounding distance)
enum class SeedSampler { Fast = 0, ExactParallel = 1, Sequential = 2 };
int32_t kmeanspp_seed(const float* x, uint32_t n, uint32_t ldx, uint32_t dim, uint32_t ld, uint32_t nlist,
                      float* centroids, KMeansScratch& sc, SeedSampler sampler, cudaStream_t stream);
// the steps of kmeanspp_seed, for the data-parallel driver (sharded_index.cu)
int32_t kmeanspp_init(const float* x, uint32_t n, uint32_t ldx, uint32_t ld, float* centroids, KMeansScratch& sc,
                      float* mind, uint64_t n_fill, cudaStream_t stream);
int32_t kmeanspp_dist_plan(const float* x, uint32_t n, uint32_t ldx, uint32_t dim, uint32_t ld, SeedDistPlan* plan);
int32_t kmeanspp_dist(const SeedDistPlan& plan, const float* cnew, const float* mind_in, const PeerF32& mind_out,
                      cudaStream_t stream);
int32_t kmeanspp_sample(const float* x, uint32_t n, uint32_t ldx, uint32_t ld, const float* mind, float* centroids,
                        uint32_t c, KMeansScratch& sc, SeedSampler sampler, cudaStream_t stream);
int32_t kmeans_update_exact_range(const float* x, uint32_t n, uint32_t ldx, const uint32_t* assign, uint32_t nc,
                                  uint32_t ld, float* centroids, KMeansScratch& sc, uint32_t c_lo, uint32_t c_hi,
                                  cudaStream_t stream);
int32_t kmeans_update_exact(const float* x, uint32_t n, uint32_t ldx, const uint32_t* assign, uint32_t nc,
                            uint32_t ld, float* centroids, KMeansScratch& sc, cudaStream_t stream);
int32_t kmeans_cluster_sums(const float* x, uint32_t n, uint32_t ldx, const uint32_t* assign, uint32_t nc,
                            uint32_t ld, float* sums, uint32_t* counts, KMeansScratch& sc, cudaStream_t stream);
int32_t kmeans_divide(const float* sums, const uint32_t* counts, uint32_t nc, uint32_t ld, float* centroids,
                      cudaStream_t stream);
// hist: [nlist + 1]; hist[nlist] counts assignments >= nlist
int32_t launch_hist(const uint32_t* assign, uint64_t n, uint32_t nlist, uint32_t shard_rank, const uint8_t* owner,
                    uint32_t* hist, cudaStream_t stream);
int32_t launch_scatter_rows(const float* x, uint32_t ldx, const uint64_t* ids, uint64_t id_base, uint64_t n,
                            const uint32_t* assign, const uint32_t* old_rows, uint32_t* fill,
                            const uint32_t* page_off, const uint64_t* page_vec, const uint64_t* page_ids,
                            uint32_t page_rows, uint32_t ld, uint32_t nlist, uint32_t shard_rank,
                            const uint8_t* owner, uint32_t mirror_off, uint32_t mirror_kind, cudaStream_t stream);
int32_t launch_page_norms(const float* rows, uint32_t ld, float* norms, uint32_t r0, uint32_t count,
                          uint32_t mirror_off, uint32_t mirror_kind, uint32_t page_rows, cudaStream_t stream);
int32_t launch_pad_rows(const float* src, uint32_t lds, uint32_t dim, float* dst, uint32_t ld, uint64_t n,
                        cudaStream_t stream);

}  // namespace vdb
