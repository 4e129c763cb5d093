// vdb_arena: the TransferManager rewrite (transfer_manager.h:42-88,
// transfer_manager.cpp:12-271).  One HBM slab and one pinned-host slab, each
// carved by a best-fit allocator with 256-byte granularity and eager
// coalescing (the reference sorts and rescans its block vector on every free,
// transfer_manager.cpp:103-162); a pool of non-blocking streams; async copies.
// The index itself places inverted-list pages in its own slabs (index.cu); the
// arena serves callers that used TransferManager for staging buffers.
#include <condition_variable>
#include <map>
#include <mutex>
#include <queue>
#include <vector>

#include "common.cuh"

namespace vdb {
namespace {

class SlabAllocator {
public:
    void init(uint8_t* base, uint64_t size) {
        base_ = base;
        size_ = size;
        free_.clear();
        used_.clear();
        if (size) free_[0] = size;
        in_use_ = peak_ = 0;
    }
    void* allocate(uint64_t bytes) {
        if (!bytes) return nullptr;
        bytes = (bytes + 255) / 256 * 256;
        std::lock_guard<std::mutex> l(mu_);
        auto best = free_.end();
        for (auto it = free_.begin(); it != free_.end(); ++it)
            if (it->second >= bytes && (best == free_.end() || it->second < best->second)) best = it;
        if (best == free_.end()) return nullptr;  // pool exhausted -> nullptr (transfer_manager.cpp:125)
        const uint64_t off = best->first, sz = best->second;
        free_.erase(best);
        if (sz > bytes) free_[off + bytes] = sz - bytes;
        used_[off] = bytes;
        in_use_ += bytes;
        peak_ = std::max(peak_, in_use_);
        return base_ + off;
    }
    bool release(void* p) {
        std::lock_guard<std::mutex> l(mu_);
        const uint64_t off = (uint64_t)((uint8_t*)p - base_);
        auto it = used_.find(off);
        if (it == used_.end()) return false;
        uint64_t sz = it->second;
        used_.erase(it);
        in_use_ -= sz;
        uint64_t o = off;
        auto nx = free_.lower_bound(o);
        if (nx != free_.end() && nx->first == o + sz) {  // merge with the right neighbour
            sz += nx->second;
            nx = free_.erase(nx);
        }
        if (nx != free_.begin()) {  // merge with the left neighbour
            auto pv = std::prev(nx);
            if (pv->first + pv->second == o) {
                o = pv->first;
                sz += pv->second;
                free_.erase(pv);
            }
        }
        free_[o] = sz;
        return true;
    }
    bool owns(const void* p) const { return (const uint8_t*)p >= base_ && (const uint8_t*)p < base_ + size_; }
    uint64_t in_use() const { return in_use_; }
    uint64_t peak() const { return peak_; }
    uint64_t live() const { return used_.size(); }

private:
    uint8_t* base_ = nullptr;
    uint64_t size_ = 0, in_use_ = 0, peak_ = 0;
    std::map<uint64_t, uint64_t> free_, used_;
    std::mutex mu_;
};

}  // namespace
}  // namespace vdb

using namespace vdb;

struct vdb_arena {
    int device = 0;
    void* dev_base = nullptr;
    void* pin_base = nullptr;
    SlabAllocator dev, pin;
    std::vector<cudaStream_t> streams;
    std::queue<cudaStream_t> available;
    std::mutex smu;
    std::condition_variable scv;
};

extern "C" {

int32_t vdb_arena_create(int32_t device, uint64_t device_bytes, uint64_t pinned_bytes, int32_t num_streams,
                         vdb_arena** out) {
    VDB_REQUIRE(out && num_streams >= 1 && num_streams <= 64, "arena: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        set_last_error("arena: no such CUDA device");
        return VDB_CUDA_ERROR;
    }
    vdb_arena* a = new vdb_arena();
    a->device = device;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    auto fail = [&](cudaError_t e) {
        set_last_error(std::string("arena: ") + cudaGetErrorString(e));
        if (a->dev_base) cudaFree(a->dev_base);
        if (a->pin_base) cudaFreeHost(a->pin_base);
        for (auto s : a->streams) cudaStreamDestroy(s);
        delete a;
        cudaSetDevice(prev);
        return e == cudaErrorMemoryAllocation ? VDB_OUT_OF_MEMORY : VDB_CUDA_ERROR;
    };
    cudaError_t e;
    if (device_bytes && (e = cudaMalloc(&a->dev_base, device_bytes)) != cudaSuccess) return fail(e);
    if (pinned_bytes && (e = cudaMallocHost(&a->pin_base, pinned_bytes)) != cudaSuccess) return fail(e);
    a->dev.init((uint8_t*)a->dev_base, device_bytes);
    a->pin.init((uint8_t*)a->pin_base, pinned_bytes);
    for (int i = 0; i < num_streams; ++i) {
        cudaStream_t s;
        if ((e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
        a->streams.push_back(s);
        a->available.push(s);
    }
    cudaSetDevice(prev);
    *out = a;
    return VDB_OK;
}

int32_t vdb_arena_destroy(vdb_arena* a) {
    if (!a) return VDB_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(a->device);
    for (auto s : a->streams) {
        cudaStreamSynchronize(s);
        cudaStreamDestroy(s);
    }
    if (a->dev_base) cudaFree(a->dev_base);
    if (a->pin_base) cudaFreeHost(a->pin_base);
    cudaSetDevice(prev);
    delete a;
    return VDB_OK;
}

void* vdb_arena_allocate_device(vdb_arena* a, uint64_t bytes) { return a ? a->dev.allocate(bytes) : nullptr; }
void* vdb_arena_allocate_pinned(vdb_arena* a, uint64_t bytes) { return a ? a->pin.allocate(bytes) : nullptr; }

int32_t vdb_arena_free_device(vdb_arena* a, void* p) {
    VDB_REQUIRE(a, "arena: null handle");
    if (!p) return VDB_OK;
    VDB_REQUIRE(a->dev.owns(p) && a->dev.release(p), "arena: pointer was not allocated from the device pool");
    return VDB_OK;
}

int32_t vdb_arena_free_pinned(vdb_arena* a, void* p) {
    VDB_REQUIRE(a, "arena: null handle");
    if (!p) return VDB_OK;
    VDB_REQUIRE(a->pin.owns(p) && a->pin.release(p), "arena: pointer was not allocated from the pinned pool");
    return VDB_OK;
}

void* vdb_arena_get_stream(vdb_arena* a) {
    if (!a) return nullptr;
    std::unique_lock<std::mutex> l(a->smu);
    a->scv.wait(l, [&] { return !a->available.empty(); });  // blocks like transfer_manager.cpp:202-209
    cudaStream_t s = a->available.front();
    a->available.pop();
    return s;
}

int32_t vdb_arena_return_stream(vdb_arena* a, void* stream) {
    VDB_REQUIRE(a && stream, "arena: null handle or stream");
    {
        std::lock_guard<std::mutex> l(a->smu);
        a->available.push((cudaStream_t)stream);
    }
    a->scv.notify_one();
    return VDB_OK;
}

int32_t vdb_arena_enqueue_transfer(vdb_arena* a, void* dst, const void* src, uint64_t bytes, int32_t kind,
                                   void* stream) {
    VDB_REQUIRE(a && dst && src, "arena: null transfer endpoint");
    VDB_REQUIRE(kind >= 1 && kind <= 4, "arena: kind must be a cudaMemcpyKind in [1,4]");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(a->device);
    cudaStream_t s = (cudaStream_t)stream;
    bool borrowed = false;
    if (!s) {  // "stream (optional)", transfer_manager.h:37: pick one from the pool
        s = (cudaStream_t)vdb_arena_get_stream(a);
        borrowed = true;
    }
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, (cudaMemcpyKind)kind, s);
    if (borrowed) vdb_arena_return_stream(a, s);
    cudaSetDevice(prev);
    VDB_CUDA_TRY(e);
    return VDB_OK;
}

// enqueue_transfer with Transfer::callback (transfer_manager.cpp:218-261 passes it to cudaLaunchHostFunc): the copy
// and then the host function are enqueued on the stream; the caller is not blocked
int32_t vdb_arena_enqueue_transfer_cb(vdb_arena* a, void* dst, const void* src, uint64_t bytes, int32_t kind,
                                      void* stream, void (*callback)(void*), void* user) {
    VDB_REQUIRE(a && dst && src, "arena: null transfer endpoint");
    VDB_REQUIRE(kind >= 1 && kind <= 4, "arena: kind must be a cudaMemcpyKind in [1,4]");
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(a->device);
    cudaStream_t s = (cudaStream_t)stream;
    bool borrowed = false;
    if (!s) {
        s = (cudaStream_t)vdb_arena_get_stream(a);
        borrowed = true;
    }
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, (cudaMemcpyKind)kind, s);
    if (e == cudaSuccess && callback) e = cudaLaunchHostFunc(s, callback, user);
    if (borrowed) vdb_arena_return_stream(a, s);
    cudaSetDevice(prev);
    VDB_CUDA_TRY(e);
    return VDB_OK;
}

int32_t vdb_arena_synchronize(vdb_arena* a) {
    VDB_REQUIRE(a, "arena: null handle");
    for (auto s : a->streams) VDB_CUDA_TRY(cudaStreamSynchronize(s));
    return VDB_OK;
}

int32_t vdb_arena_synchronize_stream(vdb_arena* a, void* stream) {
    VDB_REQUIRE(a, "arena: null handle");
    VDB_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return VDB_OK;
}

int32_t vdb_arena_stats(vdb_arena* a, uint64_t* out) {
    VDB_REQUIRE(a && out, "arena: null handle or buffer");
    out[0] = a->dev.in_use();
    out[1] = a->dev.peak();
    out[2] = a->pin.in_use();
    out[3] = a->dev.live() + a->pin.live();
    return VDB_OK;
}

}  // extern "C"
