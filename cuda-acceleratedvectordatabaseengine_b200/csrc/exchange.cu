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
#include "exchange.cuh"
#include "topk.cuh"

#include <cstring>
#include <string>
#include <vector>

namespace vdb {
namespace {

constexpr unsigned long long EX_TIMEOUT_NS = 20ull * 1000 * 1000 * 1000;  // a dead peer must not hang the GPU

struct ExchangeParams {
    PublishTarget pub;  // mailboxes, rank/world, epoch
    uint32_t nq, k, P;
    uint32_t mode;  // 0 = publish + collect (one call), 1 = publish only, 2 = collect only
    const float* local_d;
    const uint64_t* local_i;
    float* out_d;
    uint64_t* out_i;
    uint32_t* error;  // set to 1 when a wait timed out
};

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
    const uint32_t tid = threadIdx.x, half = p.pub.epoch & 1u, world = p.pub.world;
    const size_t slot_stride = (size_t)p.pub.max_nq * p.pub.max_k;

    // ---- publish: my rows of every query this CTA owns -> slot [half][my rank] of every target mailbox
    for (uint32_t q = blockIdx.x; q < p.nq && p.mode != 2u; q += gridDim.x) {
        publish_query(p.pub, q, p.k, p.local_d + (size_t)q * p.k, p.local_i + (size_t)q * p.k, p.k, MERGE_THREADS);
        __syncthreads();
    }
    // ---- collect + merge, from my own mailbox
    const Mailbox mine = p.pub.box[p.pub.rank];
    for (uint32_t q = blockIdx.x; q < p.nq && p.mode != 1u; q += gridDim.x) {
        if (tid == 0) {
            cnt = 0;
            thr = INFINITY;
            s_fail = 0;
        }
        __syncthreads();
        if (tid < world) {
            const uint32_t* f = &mine.flags[((size_t)half * world + tid) * p.pub.max_nq + q];
            const unsigned long long t0 = global_ns();
            while (ld_acquire_sys(f) != p.pub.epoch) {
                if (global_ns() - t0 > EX_TIMEOUT_NS) {
                    s_fail = 1;
                    break;
                }
                __nanosleep(256);  // the flags live in local HBM: keep the pollers off the L2 the scan streams through
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
        for (uint32_t r = 0; r < world; ++r) {
            const size_t src = ((size_t)half * world + r) * slot_stride + (size_t)q * p.k;
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
    std::vector<bool> opened;      // mapped through a CUDA IPC handle (to be closed)
    uint32_t* d_error = nullptr;
    uint32_t* h_error = nullptr;
    bool connected = false;
    bool pending = false;  // a batch was published and not yet collected
    uint32_t pending_nq = 0, pending_k = 0;
    // publish only into rank `root`'s mailbox (single-process sharded index: the root device alone collects)
    bool root_only = false;
    uint32_t root = 0;
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

void fill_target(const vdb_exchange* ex, uint32_t epoch, PublishTarget* t) {
    std::memset(t, 0, sizeof(*t));
    for (uint32_t r = 0; r < ex->world; ++r) {
        uint8_t* base = static_cast<uint8_t*>(ex->peer_base[r]);
        if (!base) continue;  // root_only: peers other than the root are never addressed
        t->box[r].flags = reinterpret_cast<uint32_t*>(base);
        t->box[r].dist = reinterpret_cast<float*>(base + dist_off(ex));
        t->box[r].ids = reinterpret_cast<uint64_t*>(base + ids_off(ex));
    }
    t->rank = ex->rank; t->world = ex->world; t->max_nq = ex->max_nq; t->max_k = ex->max_k;
    t->epoch = epoch;
    t->dst_lo = ex->root_only ? ex->root : 0;
    t->dst_hi = ex->root_only ? ex->root + 1 : ex->world;
    t->enabled = 1;
}

int32_t check_healthy(const vdb_exchange* ex) {
    if (*ex->h_error) {
        set_last_error("exchange: a call timed out waiting for a peer (vdb_exchange_reset clears the state)");
        return VDB_NCCL_ERROR;
    }
    return VDB_OK;
}

int32_t exchange_launch(vdb_exchange* ex, uint32_t mode, uint32_t epoch, const float* local_d, const uint64_t* local_i,
                        uint32_t nq, uint32_t k, float* out_d, uint64_t* out_i, cudaStream_t s) {
    DevGuard g(ex->device);
    ExchangeParams p{};
    fill_target(ex, epoch, &p.pub);
    p.nq = nq; p.k = k;
    p.P = next_pow2(2 * ex->world * k);  // pool never more than half full: the hashed duplicate screen applies
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
    // a collect-only launch just waits and merges 7.7 KB per rank: a few CTAs looping over the queries leave the
    // SMs to the scans it overlaps with (and a waiting CTA never blocks a publish: those are separate launches)
    const uint32_t grid = mode == 2u ? std::min<uint32_t>(nq, 16u) : std::min<uint32_t>(nq, (uint32_t)sms);
    exchange_merge_kernel<<<grid, MERGE_THREADS, smem, s>>>(p);
    VDB_CUDA_TRY(cudaGetLastError());
    if (mode != 1u)  // the host sees a timeout after synchronising with `s` (vdb_exchange_status)
        VDB_CUDA_TRY(cudaMemcpyAsync(ex->h_error, ex->d_error, 4, cudaMemcpyDeviceToHost, s));
    return VDB_OK;
}

}  // namespace

namespace vdb {

int32_t exchange_begin_publish(vdb_exchange* ex, uint32_t nq, uint32_t k, PublishTarget* t) {
    VDB_REQUIRE(ex && t, "exchange: null argument");
    VDB_REQUIRE(ex->connected, "exchange: vdb_exchange_connect has not been called");
    VDB_REQUIRE(!ex->pending, "exchange: collect the previous batch first (one batch in flight)");
    VDB_REQUIRE(nq >= 1 && nq <= ex->max_nq && k >= 1 && k <= ex->max_k, "exchange: nq or k above the mailbox size");
    VDB_TRY(check_healthy(ex));
    fill_target(ex, ex->epoch + 1, t);
    ++ex->epoch;  // the caller's launch is what publishes; a failed launch there is fatal for the index anyway
    ex->pending = !(ex->root_only && ex->rank != ex->root);  // only the root of a root-only exchange collects
    ex->pending_nq = nq;
    ex->pending_k = k;
    return VDB_OK;
}

int32_t exchange_collect(vdb_exchange* ex, float* out_d, uint64_t* out_i, cudaStream_t s) {
    VDB_REQUIRE(ex && out_d && out_i, "exchange_collect: null buffer");
    VDB_REQUIRE(ex->pending, "exchange_collect: nothing was published");
    VDB_TRY(exchange_launch(ex, 2, ex->epoch, nullptr, nullptr, ex->pending_nq, ex->pending_k, out_d, out_i, s));
    ex->pending = false;
    return VDB_OK;
}

}  // namespace vdb

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
    *ex->h_error = 0;
    // the memsets ran on the legacy default stream, which the callers' non-blocking streams do not wait for: a
    // publish launched right after create could otherwise have its flags zeroed behind it
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return fail(e);
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

int32_t vdb_exchange_connect_local(vdb_exchange** all, uint32_t world, uint32_t root) {
    VDB_REQUIRE(all && world >= 1 && world <= EX_MAX_WORLD && root < world, "exchange_connect_local: bad arguments");
    for (uint32_t r = 0; r < world; ++r)
        VDB_REQUIRE(all[r] && all[r]->rank == r && all[r]->world == world && all[r]->max_nq == all[0]->max_nq &&
                        all[r]->max_k == all[0]->max_k,
                    "exchange_connect_local: exchanges must be given in rank order with equal shapes");
    // same process: the root's mailbox is addressed directly once peer access is on (no IPC mapping); only the
    // root collects, so every rank publishes into the root's mailbox alone
    for (uint32_t r = 0; r < world; ++r) {
        vdb_exchange* ex = all[r];
        if (r != root && ex->device != all[root]->device) {
            DevGuard g(ex->device);
            int can = 0;
            VDB_CUDA_TRY(cudaDeviceCanAccessPeer(&can, ex->device, all[root]->device));
            if (!can) {
                set_last_error("exchange_connect_local: no peer access between the shard devices");
                return VDB_CUDA_ERROR;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(all[root]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) VDB_CUDA_TRY(e);
            cudaGetLastError();
        }
        ex->peer_base.assign(world, nullptr);
        ex->peer_base[r] = ex->local;
        ex->peer_base[root] = all[root]->local;
        ex->root_only = true;
        ex->root = root;
        ex->connected = true;
    }
    return VDB_OK;
}

int32_t vdb_exchange_merge_topk(vdb_exchange* ex, const float* local_dist_dev, const uint64_t* local_ids_dev,
                                uint32_t nq, uint32_t k, float* distances_dev, uint64_t* indices_dev, void* stream) {
    VDB_REQUIRE(ex && local_dist_dev && local_ids_dev && distances_dev && indices_dev, "exchange_merge: null buffer");
    VDB_REQUIRE(ex->connected, "exchange_merge: vdb_exchange_connect has not been called");
    VDB_REQUIRE(!ex->root_only, "exchange_merge: a root-only exchange publishes and collects separately");
    VDB_REQUIRE(!ex->pending, "exchange_merge: a published batch is waiting for vdb_exchange_collect");
    VDB_REQUIRE(nq >= 1 && nq <= ex->max_nq && k >= 1 && k <= ex->max_k, "exchange_merge: nq or k above the mailbox size");
    VDB_TRY(check_healthy(ex));
    VDB_TRY(exchange_launch(ex, 0, ex->epoch + 1, local_dist_dev, local_ids_dev, nq, k, distances_dev, indices_dev,
                            (cudaStream_t)stream));
    ++ex->epoch;  // only a launch that went out advances the call counter (the ranks' epochs must stay in step)
    return VDB_OK;
}

int32_t vdb_exchange_publish(vdb_exchange* ex, const float* local_dist_dev, const uint64_t* local_ids_dev, uint32_t nq,
                             uint32_t k, void* stream) {
    VDB_REQUIRE(ex && local_dist_dev && local_ids_dev, "exchange_publish: null buffer");
    VDB_REQUIRE(ex->connected, "exchange_publish: vdb_exchange_connect has not been called");
    VDB_REQUIRE(!ex->pending, "exchange_publish: collect the previous batch first (one batch in flight)");
    VDB_REQUIRE(nq >= 1 && nq <= ex->max_nq && k >= 1 && k <= ex->max_k, "exchange_publish: nq or k above the mailbox size");
    VDB_TRY(check_healthy(ex));
    VDB_TRY(exchange_launch(ex, 1, ex->epoch + 1, local_dist_dev, local_ids_dev, nq, k, nullptr, nullptr,
                            (cudaStream_t)stream));
    ++ex->epoch;
    ex->pending = true;
    ex->pending_nq = nq;
    ex->pending_k = k;
    return VDB_OK;
}

int32_t vdb_exchange_collect(vdb_exchange* ex, float* distances_dev, uint64_t* indices_dev, void* stream) {
    return exchange_collect(ex, distances_dev, indices_dev, (cudaStream_t)stream);
}

int32_t vdb_exchange_status(vdb_exchange* ex) {
    VDB_REQUIRE(ex, "exchange_status: null handle");
    return check_healthy(ex);
}

int32_t vdb_exchange_reset(vdb_exchange* ex) {
    VDB_REQUIRE(ex, "exchange_reset: null handle");
    DevGuard g(ex->device);
    VDB_CUDA_TRY(cudaDeviceSynchronize());
    VDB_CUDA_TRY(cudaMemset(ex->d_error, 0, 4));
    *ex->h_error = 0;
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
