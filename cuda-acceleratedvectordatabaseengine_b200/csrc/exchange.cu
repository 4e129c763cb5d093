// Cross-GPU top-k exchange fused with the merge, over NVLink peer memory.
//
// The sharded search ends with merge_results across shards (ivf_flat_index.cpp:474-518 applied to the per-rank
// results): every rank holds a local [nq][k] (distance, id) block and every rank needs the merged [nq][k].  The
// portable form is two NCCL all-gathers followed by vdb_merge_topk (three launches, each latency-bound: the payload
// is 7.7 KB per rank at nq = 64, k = 10).  Here it is ONE kernel per rank:
//   publish   each CTA stores its queries' local top-k straight into the mailbox of every peer (peer-to-peer
//             stores through NVSwitch), fences system-wide and raises one flag per (source rank, query) there;
//   collect   the same CTA then waits on the flags its peers raise in ITS mailbox (acquire loads from local HBM),
//             pools the world x k pairs of a query, sorts by (distance, id), drops later occurrences of an id and
//             the padding, and writes the k survivors.
// Mailboxes are cudaMalloc'ed per rank and mapped into the peers with CUDA IPC handles (exchanged by the host
// through torch.distributed).  Flags carry a call counter that only grows, so nothing is ever reset; two mailbox
// halves alternate by call parity, which is enough: a rank cannot publish call e + 2 before every peer has finished
// reading call e, because its own call e + 1 had to wait for their publish of e + 1.
// publish and collect can also be issued as two launches (vdb_exchange_publish / vdb_exchange_collect) with other
// work of the same stream in between -- e.g. the next batch's scan -- so that by the time a rank collects, its peers
// have long published and nobody waits for the slowest rank of a batch.  One batch may be in flight: collect(i) comes
// before publish(i + 1) on every rank, which is what keeps two mailbox halves sufficient (a peer's publish(i + 1)
// follows its collect(i), which needed my publish(i), which followed my collect(i - 1)).
// The grid never exceeds the number of SMs and a CTA publishes ALL its queries before it waits for any, so every
// rank's publishes are issued whatever the scheduling order -- no rank can wait on a CTA that is not resident.
#include "common.cuh"
#include "topk.cuh"

#include <cstring>
#include <string>
#include <vector>

namespace vdb {
namespace {

constexpr uint32_t EX_MAX_WORLD = 16;
constexpr unsigned long long EX_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;  // a dead peer must not hang the GPU

struct Mailbox {  // device pointers into ONE rank's mailbox allocation
    uint32_t* flags;  // [2][world][max_nq]
    float* dist;      // [2][world][max_nq * max_k]
    uint64_t* ids;    // [2][world][max_nq * max_k]
};

struct ExchangeParams {
    Mailbox box[EX_MAX_WORLD];  // box[r] = rank r's mailbox as mapped into this process
    uint32_t rank, world, max_nq, max_k;
    uint32_t nq, k, P, epoch;
    uint32_t mode;  // 0 = publish + collect (one call), 1 = publish only, 2 = collect only
    const float* local_d;
    const uint64_t* local_i;
    float* out_d;
    uint64_t* out_i;
    uint32_t* error;  // set to 1 when a wait timed out
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(MERGE_THREADS) exchange_merge_kernel(const ExchangeParams p) {
    extern __shared__ __align__(16) uint8_t esm[];
    uint64_t* pi = reinterpret_cast<uint64_t*>(esm);
    uint64_t* ti = pi + p.P;
    float* pd = reinterpret_cast<float*>(ti + p.P);
    float* td = pd + p.P;
    __shared__ uint32_t cnt;
    __shared__ float thr;
    __shared__ uint32_t s_scan[MERGE_THREADS / 32 + 1];
    __shared__ uint32_t s_fail;
    const MergePool pool{pd, pi, &cnt, &thr};
    const uint32_t tid = threadIdx.x, half = p.epoch & 1u;
    const size_t slot_stride = (size_t)p.max_nq * p.max_k;

    // ---- publish: my rows of every query this CTA owns -> slot [half][my rank] of every mailbox
    for (uint32_t q = blockIdx.x; q < p.nq && p.mode != 2u; q += gridDim.x) {
        for (uint32_t e = tid; e < p.world * p.k; e += MERGE_THREADS) {
            const uint32_t r = e / p.k, j = e % p.k;
            const size_t dst = ((size_t)half * p.world + p.rank) * slot_stride + (size_t)q * p.k + j;
            p.box[r].dist[dst] = p.local_d[(size_t)q * p.k + j];
            p.box[r].ids[dst] = p.local_i[(size_t)q * p.k + j];
        }
        __threadfence_system();
        __syncthreads();
        if (tid < p.world)
            st_release_sys(&p.box[tid].flags[((size_t)half * p.world + p.rank) * p.max_nq + q], p.epoch);
    }
    // ---- collect + merge, from my own mailbox
    const Mailbox mine = p.box[p.rank];
    for (uint32_t q = blockIdx.x; q < p.nq && p.mode != 1u; q += gridDim.x) {
        if (tid == 0) {
            cnt = 0;
            thr = INFINITY;
            s_fail = 0;
        }
        __syncthreads();
        if (tid < p.world) {
            const uint32_t* f = &mine.flags[((size_t)half * p.world + tid) * p.max_nq + q];
            const unsigned long long t0 = global_ns();
            while (ld_acquire_sys(f) != p.epoch) {
                if (global_ns() - t0 > EX_TIMEOUT_NS) {
                    s_fail = 1;
                    break;
                }
                __nanosleep(64);
            }
        }
        __syncthreads();
        if (s_fail) {
            if (tid == 0) *p.error = 1;
            for (uint32_t j = tid; j < p.k; j += MERGE_THREADS) {
                p.out_d[(size_t)q * p.k + j] = FLT_MAX;
                p.out_i[(size_t)q * p.k + j] = ID_PAD;
            }
            __syncthreads();
            continue;
        }
        for (uint32_t r = 0; r < p.world; ++r) {
            const size_t src = ((size_t)half * p.world + r) * slot_stride + (size_t)q * p.k;
            pool_push_block(pool, p.P, mine.dist + src, mine.ids + src, p.k);
        }
        pool_compact_block(pool, p.P, p.k, true, td, ti, s_scan);
        const uint32_t got = cnt;
        for (uint32_t j = tid; j < p.k; j += MERGE_THREADS) {
            p.out_d[(size_t)q * p.k + j] = j < got ? pd[j] : FLT_MAX;
            p.out_i[(size_t)q * p.k + j] = j < got ? pi[j] : ID_PAD;
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace vdb

using namespace vdb;

struct vdb_exchange {
    int device = 0;
    uint32_t rank = 0, world = 1, max_nq = 0, max_k = 0, epoch = 0;
    void* local = nullptr;  // this rank's mailbox allocation
    size_t bytes = 0;
    std::vector<void*> peer_base;  // [world] (own entry = local)
    std::vector<bool> opened;
    uint32_t* d_error = nullptr;
    uint32_t* h_error = nullptr;
    bool connected = false;
    bool pending = false;  // a batch was published and not yet collected
    uint32_t pending_nq = 0, pending_k = 0;
};

namespace {

size_t flags_bytes(const vdb_exchange* ex) { return (size_t)2 * ex->world * ex->max_nq * 4; }
size_t dist_off(const vdb_exchange* ex) { return (flags_bytes(ex) + 255) & ~(size_t)255; }
size_t dist_bytes(const vdb_exchange* ex) { return (size_t)2 * ex->world * ex->max_nq * ex->max_k * 4; }
size_t ids_off(const vdb_exchange* ex) { return (dist_off(ex) + dist_bytes(ex) + 255) & ~(size_t)255; }
size_t total_bytes(const vdb_exchange* ex) { return ids_off(ex) + (size_t)2 * ex->world * ex->max_nq * ex->max_k * 8; }

struct DevGuard {
    int prev = 0;
    explicit DevGuard(int d) {
        cudaGetDevice(&prev);
        if (prev != d) cudaSetDevice(d);
    }
    ~DevGuard() { cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

int32_t vdb_exchange_create(int32_t device, uint32_t rank, uint32_t world, uint32_t max_nq, uint32_t max_k,
                            vdb_exchange** out) {
    VDB_REQUIRE(out, "exchange_create: null output handle");
    VDB_REQUIRE(world >= 1 && world <= EX_MAX_WORLD && rank < world, "exchange_create: bad rank/world (<= 16 ranks)");
    VDB_REQUIRE(max_nq >= 1 && max_k >= 1 && (uint64_t)world * max_k <= 4096,
                "exchange_create: world * max_k must be <= 4096 (use the all-gather path beyond that)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        set_last_error("exchange_create: no such CUDA device");
        return VDB_CUDA_ERROR;
    }
    vdb_exchange* ex = new vdb_exchange();
    ex->device = device; ex->rank = rank; ex->world = world; ex->max_nq = max_nq; ex->max_k = max_k;
    DevGuard g(device);
    ex->bytes = total_bytes(ex);
    auto fail = [&](cudaError_t e) {
        set_last_error(std::string("exchange_create: ") + cudaGetErrorString(e));
        cudaFree(ex->local); cudaFree(ex->d_error);
        if (ex->h_error) cudaFreeHost(ex->h_error);
        delete ex;
        return VDB_CUDA_ERROR;
    };
    cudaError_t e;
    if ((e = cudaMalloc(&ex->local, ex->bytes)) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(ex->local, 0, ex->bytes)) != cudaSuccess) return fail(e);  // flags = 0 < every epoch
    if ((e = cudaMalloc(&ex->d_error, 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMemset(ex->d_error, 0, 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMallocHost(&ex->h_error, 4)) != cudaSuccess) return fail(e);
    ex->peer_base.assign(world, nullptr);
    ex->opened.assign(world, false);
    ex->peer_base[rank] = ex->local;
    ex->connected = world == 1;
    *out = ex;
    return VDB_OK;
}

int32_t vdb_exchange_handle(vdb_exchange* ex, uint8_t* out64) {
    VDB_REQUIRE(ex && out64, "exchange_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DevGuard g(ex->device);
    cudaIpcMemHandle_t h;
    VDB_CUDA_TRY(cudaIpcGetMemHandle(&h, ex->local));
    std::memcpy(out64, &h, 64);
    return VDB_OK;
}

int32_t vdb_exchange_connect(vdb_exchange* ex, const uint8_t* handles /* [world][64], own entry ignored */) {
    VDB_REQUIRE(ex && handles, "exchange_connect: null argument");
    DevGuard g(ex->device);
    for (uint32_t r = 0; r < ex->world; ++r) {
        if (r == ex->rank || ex->opened[r]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        VDB_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ex->peer_base[r] = p;
        ex->opened[r] = true;
    }
    ex->connected = true;
    return VDB_OK;
}

static int32_t exchange_launch(vdb_exchange* ex, uint32_t mode, uint32_t epoch, const float* local_d,
                               const uint64_t* local_i, uint32_t nq, uint32_t k, float* out_d, uint64_t* out_i,
                               cudaStream_t s) {
    DevGuard g(ex->device);
    if (*ex->h_error) {
        set_last_error("exchange: an earlier call timed out waiting for a peer");
        return VDB_NCCL_ERROR;
    }
    ExchangeParams p{};
    for (uint32_t r = 0; r < ex->world; ++r) {
        uint8_t* base = static_cast<uint8_t*>(ex->peer_base[r]);
        p.box[r].flags = reinterpret_cast<uint32_t*>(base);
        p.box[r].dist = reinterpret_cast<float*>(base + dist_off(ex));
        p.box[r].ids = reinterpret_cast<uint64_t*>(base + ids_off(ex));
    }
    p.rank = ex->rank; p.world = ex->world; p.max_nq = ex->max_nq; p.max_k = ex->max_k;
    p.nq = nq; p.k = k;
    p.P = next_pow2(2 * ex->world * k);  // pool never more than half full: the hashed duplicate screen applies
    p.epoch = epoch;
    p.mode = mode;
    p.local_d = local_d; p.local_i = local_i; p.out_d = out_d; p.out_i = out_i;
    p.error = ex->d_error;
    const uint32_t smem = p.P * 24;
    static bool conf[16] = {false};
    if (ex->device < 16 && !conf[ex->device]) {
        VDB_CUDA_TRY(cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 24));
        conf[ex->device] = true;
    }
    int sms = NUM_SMS_B200;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ex->device);
    exchange_merge_kernel<<<std::min<uint32_t>(nq, (uint32_t)sms), MERGE_THREADS, smem, s>>>(p);
    VDB_CUDA_TRY(cudaGetLastError());
    if (mode != 1u)
        VDB_CUDA_TRY(cudaMemcpyAsync(ex->h_error, ex->d_error, 4, cudaMemcpyDeviceToHost, s));  // seen by the NEXT call
    return VDB_OK;
}

int32_t vdb_exchange_merge_topk(vdb_exchange* ex, const float* local_dist_dev, const uint64_t* local_ids_dev,
                                uint32_t nq, uint32_t k, float* distances_dev, uint64_t* indices_dev, void* stream) {
    VDB_REQUIRE(ex && local_dist_dev && local_ids_dev && distances_dev && indices_dev, "exchange_merge: null buffer");
    VDB_REQUIRE(ex->connected, "exchange_merge: vdb_exchange_connect has not been called");
    VDB_REQUIRE(!ex->pending, "exchange_merge: a published batch is waiting for vdb_exchange_collect");
    VDB_REQUIRE(nq >= 1 && nq <= ex->max_nq && k >= 1 && k <= ex->max_k, "exchange_merge: nq or k above the mailbox size");
    return exchange_launch(ex, 0, ++ex->epoch, local_dist_dev, local_ids_dev, nq, k, distances_dev, indices_dev,
                           (cudaStream_t)stream);
}

int32_t vdb_exchange_publish(vdb_exchange* ex, const float* local_dist_dev, const uint64_t* local_ids_dev, uint32_t nq,
                             uint32_t k, void* stream) {
    VDB_REQUIRE(ex && local_dist_dev && local_ids_dev, "exchange_publish: null buffer");
    VDB_REQUIRE(ex->connected, "exchange_publish: vdb_exchange_connect has not been called");
    VDB_REQUIRE(!ex->pending, "exchange_publish: collect the previous batch first (one batch in flight)");
    VDB_REQUIRE(nq >= 1 && nq <= ex->max_nq && k >= 1 && k <= ex->max_k, "exchange_publish: nq or k above the mailbox size");
    VDB_TRY(exchange_launch(ex, 1, ++ex->epoch, local_dist_dev, local_ids_dev, nq, k, nullptr, nullptr,
                            (cudaStream_t)stream));
    ex->pending = true;
    ex->pending_nq = nq;
    ex->pending_k = k;
    return VDB_OK;
}

int32_t vdb_exchange_collect(vdb_exchange* ex, float* distances_dev, uint64_t* indices_dev, void* stream) {
    VDB_REQUIRE(ex && distances_dev && indices_dev, "exchange_collect: null buffer");
    VDB_REQUIRE(ex->pending, "exchange_collect: nothing was published");
    VDB_TRY(exchange_launch(ex, 2, ex->epoch, nullptr, nullptr, ex->pending_nq, ex->pending_k, distances_dev,
                            indices_dev, (cudaStream_t)stream));
    ex->pending = false;
    return VDB_OK;
}

int32_t vdb_exchange_destroy(vdb_exchange* ex) {
    if (!ex) return VDB_OK;
    {
        DevGuard g(ex->device);
        cudaDeviceSynchronize();
        for (uint32_t r = 0; r < ex->world; ++r)
            if (ex->opened[r]) cudaIpcCloseMemHandle(ex->peer_base[r]);
        cudaFree(ex->local);
        cudaFree(ex->d_error);
        if (ex->h_error) cudaFreeHost(ex->h_error);
    }
    delete ex;
    return VDB_OK;
}

}  // extern "C"
