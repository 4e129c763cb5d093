// C++ host mirror of the reference's engine interface, implemented on the C
// ABI (include/vdb_b200.h) and nothing else -- no CUDA headers, no torch.
//
// A caller of the reference (server/query_service.cpp:134, bench/benchmark.cpp:
// 63-96, test/simple_test.cpp:111-196) compiles against this header unchanged:
//   vdb::IVFFlatIndex            engine/ivf_flat_index.h:14-67
//   vdb::IVFFlatIndex::Config    :16-22   (same field names and defaults)
//   vdb::IVFFlatIndex::SearchParams :38-42
//   vdb::kernels::Metric         engine/kernels.cuh:24-28
//   vdb::TransferManager         engine/transfer_manager.h:22-88 (subset the index and server use)
// plus the methods the reference's server calls but its engine never defined:
//   get_dimension()  query_service.cpp:112,   warmup_lists()/warmup_all() :191,195.
//
// Error behaviour follows the reference: the constructor throws
// std::invalid_argument on dimension == 0 || nlist == 0 (ivf_flat_index.cpp:17-19);
// every other failure surfaces as std::runtime_error, which the server's
// catch (std::exception&) turns into grpc INTERNAL (query_service.cpp:164-167).
// There is no CPU path.  Config::use_gpu = false -- which the reference's own smoke test sets
// (test/simple_test.cpp:114) -- is ACCEPTED and has no effect: the search still runs on the GPU, and returns what
// the reference's CPU path returns (that equality is the parity contract of this repository).  A missing device
// or library is an exception, never a silent fallback.
//
// The reference's call sites compile against this header unchanged: host/dropin/engine/{ivf_flat_index,
// transfer_manager}.h forward to it, and tests/test_dropin_compile.py builds /root/reference/test/simple_test.cpp,
// test/gpu_vs_cpu_test.cpp and bench/benchmark.cpp, unmodified, on top of it.
#pragma once
// the standard headers the reference's engine headers bring in (its call sites rely on them transitively)
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <mutex>
#include <queue>
#include <shared_mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "vdb_b200.h"
#include "vdb_b200_storage.h"

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
#endif

namespace vdb {
namespace kernels {
enum class Metric { L2, InnerProduct, Cosine };
}  // namespace kernels

namespace detail {
inline void check(int32_t st, const char* what) {
    if (st == VDB_OK) return;
    std::string msg = std::string(what) + ": " + vdb_last_error_string() + " [" + vdb_status_string(st) + "]";
    if (st == VDB_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
inline void check_storage(int32_t st, const char* what) {
    if (st == VDB_OK) return;
    std::string msg = std::string(what) + ": " + vdb_storage_last_error() + " [" + vdb_status_string(st) + "]";
    if (st == VDB_INVALID_ARGUMENT) throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
}  // namespace detail

namespace kernels {
// The two launchers the reference's host object links against (engine/kernels.cu:13-43 and :80-92; the kernel seam of
// SURVEY.md 8b), same names and argument order.  Host or device pointers; the reference's launchers return void and
// leave errors to cudaGetLastError(), here a failure throws like every other call of this header.
template <typename T>
void launch_bruteforce_search(const T* database, const T* queries, const uint64_t* ids, uint64_t n_vectors,
                              uint32_t n_queries, uint32_t dim, uint32_t k, float* out_distances,
                              uint64_t* out_indices, Metric metric, cudaStream_t stream);
template <>
inline void launch_bruteforce_search<float>(const float* database, const float* queries, const uint64_t* ids,
                                            uint64_t n_vectors, uint32_t n_queries, uint32_t dim, uint32_t k,
                                            float* out_distances, uint64_t* out_indices, Metric metric,
                                            cudaStream_t stream) {
    detail::check(vdb_bruteforce_search(database, queries, ids, n_vectors, n_queries, dim, k, out_distances,
                                        out_indices, static_cast<int32_t>(metric), stream),
                  "launch_bruteforce_search");
}
// kernels.cuh:315-354 is L2-only; so is this launcher (IVFFlatIndex::add/train honour the index metric)
template <typename T>
void launch_kmeans_assign(const T* vectors, const T* centroids, uint32_t* assignments, float* distances,
                          uint64_t n_vectors, uint32_t n_centroids, uint32_t dim, cudaStream_t stream);
template <>
inline void launch_kmeans_assign<float>(const float* vectors, const float* centroids, uint32_t* assignments,
                                        float* distances, uint64_t n_vectors, uint32_t n_centroids, uint32_t dim,
                                        cudaStream_t stream) {
    detail::check(vdb_kmeans_assign(vectors, centroids, assignments, distances, n_vectors, n_centroids, dim,
                                    VDB_METRIC_L2, stream),
                  "launch_kmeans_assign");
}
}  // namespace kernels

class TransferManager {
public:
    struct Config {
        size_t pinned_pool_size = 1ULL << 30;
        size_t device_pool_size = 4ULL << 30;
        int num_streams = 4;
        bool use_async = true;
        int device = 0;
    };
    struct Transfer {
        void* src;
        void* dst;
        size_t size;
        cudaMemcpyKind kind;
        cudaStream_t stream;
        std::function<void()> callback;  // runs on a driver thread once the copy has completed (cudaLaunchHostFunc)
    };
    explicit TransferManager(const Config& config) : config_(config) {
        detail::check(vdb_arena_create(config.device, config.device_pool_size, config.pinned_pool_size,
                                       config.num_streams, &arena_), "TransferManager");
    }
    ~TransferManager() { vdb_arena_destroy(arena_); }
    TransferManager(const TransferManager&) = delete;
    TransferManager& operator=(const TransferManager&) = delete;

    void* allocate_pinned(size_t size) { return vdb_arena_allocate_pinned(arena_, size); }
    void free_pinned(void* p) { vdb_arena_free_pinned(arena_, p); }
    void* allocate_device(size_t size) { return vdb_arena_allocate_device(arena_, size); }
    void free_device(void* p) { vdb_arena_free_device(arena_, p); }
    cudaStream_t get_stream() { return static_cast<cudaStream_t>(vdb_arena_get_stream(arena_)); }
    void return_stream(cudaStream_t s) { vdb_arena_return_stream(arena_, s); }
    void enqueue_transfer(const Transfer& t) {
        if (!t.callback) {
            detail::check(vdb_arena_enqueue_transfer(arena_, t.dst, t.src, t.size, static_cast<int32_t>(t.kind), t.stream),
                          "enqueue_transfer");
            return;
        }
        // the callback travels with the copy and runs behind it on the stream, as transfer_manager.cpp:250-257 does
        auto* fn = new std::function<void()>(t.callback);
        const int32_t st = vdb_arena_enqueue_transfer_cb(
            arena_, t.dst, t.src, t.size, static_cast<int32_t>(t.kind), t.stream,
            [](void* u) {
                auto* f = static_cast<std::function<void()>*>(u);
                (*f)();
                delete f;
            },
            fn);
        if (st != VDB_OK) delete fn;
        detail::check(st, "enqueue_transfer");
    }
    void enqueue_batch(const std::vector<Transfer>& ts) { for (const auto& t : ts) enqueue_transfer(t); }
    void synchronize() { detail::check(vdb_arena_synchronize(arena_), "synchronize"); }
    void synchronize_stream(cudaStream_t s) { detail::check(vdb_arena_synchronize_stream(arena_, s), "synchronize_stream"); }
    struct MemoryStats {
        size_t total_device_allocated = 0, total_pinned_allocated = 0, active_allocations = 0,
               peak_device_usage = 0, peak_pinned_usage = 0;
    };
    MemoryStats get_memory_stats() const {
        uint64_t s[4] = {0, 0, 0, 0};
        vdb_arena_stats(arena_, s);
        MemoryStats m;
        m.total_device_allocated = s[0]; m.peak_device_usage = s[1]; m.total_pinned_allocated = s[2];
        m.active_allocations = s[3];
        return m;
    }

    vdb_arena* handle() const { return arena_; }
    int device() const { return config_.device; }

private:
    Config config_;
    vdb_arena* arena_ = nullptr;
};

class IVFFlatIndex {
public:
    struct Config {
        uint32_t dimension;
        uint32_t nlist;
        kernels::Metric metric;
        bool use_gpu = true;        // accepted either way, see the header comment
        size_t max_gpu_memory = 0;  // 0 = uncapped (the reference's 8 GiB default cannot hold the headline index)
        int device = 0;
        uint32_t shard_rank = 0, shard_count = 1;  // one process per GPU (this process holds shard_rank's lists)
        std::vector<int> devices;   // more than one entry: ONE process, one list shard per device
        uint32_t pipeline_depth = 0;  // searches in flight, 0 = 4
        uint32_t scan_mirror = 0;     // low-precision shadow for the tensor-core screen: 0 auto, 1 off, 2 bf16, 3 int8
    };
    struct SearchParams {
        uint32_t nprobe = 10;
        uint32_t k = 10;
        bool use_exact_rerank = false;  // unused by the reference as well
    };

    IVFFlatIndex(const Config& config, TransferManager* tm) : config_(config), tm_(tm) {
        if (config.dimension == 0 || config.nlist == 0)
            throw std::invalid_argument("Invalid configuration: dimension and nlist must be > 0");
        vdb_config c;
        vdb_config_default(&c);
        c.dimension = config.dimension;
        c.nlist = config.nlist;
        c.metric = static_cast<int32_t>(config.metric);
        c.device = config.device;
        c.max_gpu_memory = config.max_gpu_memory;
        c.shard_rank = config.shard_rank;
        c.shard_count = config.shard_count;
        c.pipeline_depth = config.pipeline_depth;
        c.scan_mirror = config.scan_mirror;
        if (config.devices.size() > 1) {
            std::vector<int32_t> devs(config.devices.begin(), config.devices.end());
            detail::check(vdb_index_create_sharded(&c, devs.data(), static_cast<int32_t>(devs.size()), &ix_),
                          "IVFFlatIndex");
        } else {
            if (config.devices.size() == 1) c.device = config.devices[0];
            detail::check(vdb_index_create(&c, &ix_), "IVFFlatIndex");
        }
        // device memory for the lists comes from the TransferManager's pool, as in the reference (:424-433)
        if (tm_) detail::check(vdb_index_set_arena(ix_, tm_->handle(), tm_->device()), "IVFFlatIndex");
    }
    ~IVFFlatIndex() { vdb_index_destroy(ix_); }
    IVFFlatIndex(const IVFFlatIndex&) = delete;
    IVFFlatIndex& operator=(const IVFFlatIndex&) = delete;

    // train/add are exclusive with search (SURVEY.md 8b threading): writers take the lock exclusively
    void train(const float* vectors, uint64_t n_vectors) {
        std::unique_lock<std::shared_mutex> l(mu_);
        detail::check(vdb_index_train(ix_, vectors, n_vectors), "train");
    }
    void add(const float* vectors, const uint64_t* ids, uint64_t n_vectors) {
        std::unique_lock<std::shared_mutex> l(mu_);
        detail::check(vdb_index_add(ix_, vectors, ids, n_vectors), "add");
    }
    void search(const float* queries, uint32_t n_queries, const SearchParams& params, float* distances,
                uint64_t* indices) {
        std::shared_lock<std::shared_mutex> l(mu_);
        detail::check(vdb_index_search(ix_, queries, n_queries, params.nprobe, params.k, distances, indices), "search");
    }
    // the pipelined form (the reference server's intended batches in flight, query_service.h:26-27)
    uint64_t search_submit(const float* queries, uint32_t n_queries, const SearchParams& params, float* distances,
                           uint64_t* indices) {
        std::shared_lock<std::shared_mutex> l(mu_);
        uint64_t ticket = 0;
        detail::check(vdb_index_search_submit(ix_, queries, n_queries, params.nprobe, params.k, distances, indices,
                                              &ticket), "search_submit");
        return ticket;
    }
    void search_wait(uint64_t ticket) { detail::check(vdb_index_search_wait(ix_, ticket), "search_wait"); }
    // ivf_flat_index.h:66-67 (declared by the reference, never defined): the epoch directory of format/storage.cpp
    // -- manifest.json + centroids + one Arrow IPC vector file per non-empty list (libvdb_b200_storage.so)
    void save(const std::string& path) const {
        std::shared_lock<std::shared_mutex> l(mu_);
        detail::check_storage(vdb_index_save(ix_, path.c_str()), "save");
    }
    void load(const std::string& path) {
        std::unique_lock<std::shared_mutex> l(mu_);
        detail::check_storage(vdb_index_load(ix_, path.c_str()), "load");
    }
    // server/query_service.cpp:245 calls index->load_from_epoch(epoch_ptr) with a storage::EpochManager::Epoch;
    // any type with a std::string base_path member (format/storage.h:175-181) works
    template <typename EpochPtr>
    void load_from_epoch(const EpochPtr& epoch) { load(epoch->base_path); }
    void warmup_lists(const std::vector<uint32_t>& list_ids) {
        detail::check(vdb_index_warmup(ix_, list_ids.data(), static_cast<uint32_t>(list_ids.size())), "warmup_lists");
    }
    void warmup_all() {}
    void evict_list(uint32_t) {}  // lists are HBM-resident by construction
    size_t get_gpu_memory_usage() const { return stats().gpu_memory_bytes; }
    size_t get_total_vectors() const { return stats().total_vectors; }
    uint32_t get_dimension() const { return config_.dimension; }
    const Config& config() const { return config_; }
    vdb_index* handle() { return ix_; }

private:
    vdb_stats stats() const {
        vdb_stats s;
        detail::check(vdb_index_stats(ix_, &s), "stats");
        return s;
    }
    Config config_;
    TransferManager* tm_;  // borrowed, must outlive the index (as in the reference); not needed by the kernels
    vdb_index* ix_ = nullptr;
    mutable std::shared_mutex mu_;
};

}  // namespace vdb
