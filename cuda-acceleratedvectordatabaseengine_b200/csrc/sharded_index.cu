// vdb_index_create_sharded: ONE process, one IVF-Flat shard per device behind a single handle.
//
// The reference server is one process that constructs one IVFFlatIndex (server/query_service.cpp:232-245,
// server/main.cpp:88-94); this is that object over the GPUs of the box.  Lists are partitioned by list
// (merge_results, ivf_flat_index.cpp:474-518, makes them independent units): shard r is a full vdb_index on
// devices[r] holding the lists whose owner is r, every shard keeps the same centroids and runs the same coarse
// selection, and the only exchange is the final merge -- each shard's merge kernel stores its [nq][k] block
// straight into the ROOT shard's mailbox through NVLink peer access (no IPC: same process), where a collect kernel
// merges the blocks (exchange.cu, vdb_exchange_connect_local).  One host thread drives all devices: a search is
// enqueued shard by shard on the shards' own pipelines (index.cu), so batches in flight overlap on every device.
// A device may be listed more than once (shards then share it): that is how the single-GPU tests cover this path.
#include <algorithm>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "exchange.cuh"
#include "index_internal.cuh"
#include "kmeans.cuh"

namespace vdb {

struct Composite {
    std::vector<vdb_index*> shards;
    std::vector<vdb_exchange*> exchanges;
    uint32_t root = 0;
    uint32_t max_nq = 0, max_k = 0;  // mailbox shape (grown on demand between searches)
    uint64_t next_ticket = 0;
    uint64_t total_vectors = 0;
    uint32_t seen_nq = 0, seen_np = 0, seen_k = 0;  // largest search shape so far (scratch is reserved for it)
    struct Stage {  // one shard's slice of the batch being added: rows, ids, list assignments (its own HBM)
        DevBuf<float> x;
        DevBuf<uint64_t> ids;
        DevBuf<uint32_t> asg;
        uint64_t n = 0;
    };
    std::vector<Stage> stage;
    bool peer_all = false;  // every pair of distinct shard devices has peer access (data-parallel training)
};

namespace {

int32_t connect_mailboxes(Composite* c, uint32_t max_nq, uint32_t max_k) {
    const uint32_t world = (uint32_t)c->shards.size();
    // searches in flight reference the old mailboxes: drain them first
    for (vdb_index* s : c->shards)
        for (uint32_t i = 0; i < s->depth; ++i) VDB_TRY(index_finish_slot(s, s->slots[i]));
    for (uint32_t r = 0; r < world; ++r) c->shards[r]->exchange = nullptr;
    for (vdb_exchange* e : c->exchanges) vdb_exchange_destroy(e);
    c->exchanges.assign(world, nullptr);
    for (uint32_t r = 0; r < world; ++r)
        VDB_TRY(vdb_exchange_create(c->shards[r]->device, r, world, max_nq, max_k, &c->exchanges[r]));
    VDB_TRY(vdb_exchange_connect_local(c->exchanges.data(), world, c->root));
    for (uint32_t r = 0; r < world; ++r) c->shards[r]->exchange = c->exchanges[r];
    c->max_nq = max_nq;
    c->max_k = max_k;
    return VDB_OK;
}

int device_of(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

// run f(r) for every shard on its own host thread (the shards' train / add calls synchronise internally)
template <typename F>
int32_t for_each_shard_parallel(Composite* c, F f) {
    const uint32_t world = (uint32_t)c->shards.size();
    std::vector<int32_t> st(world, VDB_OK);
    std::vector<std::string> msg(world);
    std::vector<std::thread> th;
    for (uint32_t r = 0; r < world; ++r)
        th.emplace_back([&, r] {
            st[r] = f(r);
            if (st[r] != VDB_OK) msg[r] = vdb_last_error_string();  // the error string is thread-local
        });
    for (auto& t : th) t.join();
    for (uint32_t r = 0; r < world; ++r)
        if (st[r] != VDB_OK) {
            set_last_error("shard " + std::to_string(r) + ": " + msg[r]);
            return st[r];
        }
    return VDB_OK;
}

}  // namespace

vdb_index* composite_root(vdb_index* ix) { return ix->composite->shards[ix->composite->root]; }
uint32_t composite_size(vdb_index* ix) { return (uint32_t)ix->composite->shards.size(); }
vdb_index* composite_shard(vdb_index* ix, uint32_t r) { return ix->composite->shards[r]; }

int32_t composite_destroy(vdb_index* ix) {
    Composite* c = ix->composite;
    for (vdb_index* s : c->shards)
        if (s) {
            DeviceGuard g(s->device);
            cudaDeviceSynchronize();
            s->exchange = nullptr;
        }
    for (size_t r = 0; r < c->stage.size(); ++r) {
        DeviceGuard g(c->shards[r]->device);
        c->stage[r].x.release(); c->stage[r].ids.release(); c->stage[r].asg.release();
    }
    for (vdb_exchange* e : c->exchanges) vdb_exchange_destroy(e);
    for (vdb_index* s : c->shards) vdb_index_destroy(s);
    delete c;
    ix->composite = nullptr;
    delete ix;
    return VDB_OK;
}

// Data-parallel training, bit-identical to the single-GPU trainer (and so to ivf_flat_index.cpp:49-145).
// Every device holds the whole training sample (3 GB at the largest config) but WORKS on a slice of it:
//   k-means++   per seed: each rank folds the newest seed into the running minimum of ITS rows
//               (seed_dist_tma_kernel, HBM-bound) and stores the new minima straight into every rank's copy over
//               NVLink; then every rank runs the exact D^2 sampler on the full array -- redundantly, since the
//               reference's sequential fp32 sum does not split -- and so picks the same row without any exchange.
//               Two copies of the minima alternate by seed, so a rank may write seed c+1 while a peer still samples c.
//   Lloyd x 10  assignment of the rank's rows (tensor cores), slices copied to all ranks; then the CLUSTERS are
//               partitioned: a rank sums and divides only its clusters, over all rows in input order -- one rank
//               per cluster keeps the reference's summation order -- and copies its centroid rows to all ranks.
// Ranks meet through events recorded after each step and waited for by the other ranks' streams (no host sync
// inside the loops).  The last assignment also yields the list sizes that balance the list ownership by bytes.
int32_t composite_train(vdb_index* ix, const float* vectors, uint64_t n64) {
    std::lock_guard<std::mutex> lock(ix->mu);
    Composite* c = ix->composite;
    const uint32_t R = (uint32_t)c->shards.size();
    VDB_REQUIRE(n64 < 0xffffffffull, "train: at most 2^32-2 training vectors");
    const uint32_t n = (uint32_t)n64, dim = ix->cfg.dimension, nlist = ix->cfg.nlist;
    const uint32_t ld = c->shards[0]->ld;
    const int vdev = device_of(vectors);
    if (R == 1 || n < 4096 || !c->peer_all) {  // nothing to split (or no peer access between all devices)
        vdb_index* root = c->shards[c->root];
        VDB_TRY(vdb_index_train(root, vectors, n));
        std::vector<float> cent((size_t)nlist * dim);
        std::vector<uint8_t> owners(nlist);
        VDB_TRY(vdb_index_get_centroids(root, cent.data()));
        VDB_TRY(vdb_index_get_owners(root, owners.data()));
        for (uint32_t r = 0; r < R; ++r) {
            if (r == c->root) continue;
            VDB_TRY(vdb_index_set_centroids(c->shards[r], cent.data()));
            if (c->total_vectors == 0) VDB_TRY(vdb_index_set_owners(c->shards[r], owners.data()));
        }
        return VDB_OK;
    }

    struct Rank {
        vdb_index* ix = nullptr;
        DevBuf<float> xbuf, xraw;
        const float* x = nullptr;
        KMeansScratch sc;
        SeedDistPlan plan;
        cudaEvent_t ev = nullptr;
        uint32_t lo = 0, hi = 0, c_lo = 0, c_hi = 0;
    };
    std::vector<Rank> rk(R);
    auto cleanup = [&] {
        for (Rank& r : rk) {
            if (!r.ix) continue;
            DeviceGuard g(r.ix->device);
            cudaStreamSynchronize(r.ix->stream);
            r.sc.release();
            r.xbuf.release();
            r.xraw.release();
            if (r.ev) cudaEventDestroy(r.ev);
        }
    };
    auto run = [&]() -> int32_t {
        const uint32_t per = ((n + R - 1) / R + 127) / 128 * 128;  // slices start on a TMA box boundary
        const uint32_t cper = (nlist + R - 1) / R;
        // ---- the sample on every device, [n][ld] zero padded
        for (uint32_t r = 0; r < R; ++r) {
            Rank& k = rk[r];
            k.ix = c->shards[r];
            DeviceGuard g(k.ix->device);
            std::lock_guard<std::mutex> l2(k.ix->mu);
            cudaStream_t st = k.ix->stream;
            k.lo = std::min(n, r * per);
            k.hi = std::min(n, (r + 1) * per);
            k.c_lo = std::min(nlist, r * cper);
            k.c_hi = std::min(nlist, (r + 1) * cper);
            VDB_CUDA_TRY(cudaEventCreateWithFlags(&k.ev, cudaEventDisableTiming));
            if (vdev == k.ix->device && dim == ld && !((uintptr_t)vectors & 15)) {
                k.x = vectors;
            } else {
                VDB_TRY(k.xbuf.reserve((size_t)n * ld));
                const float* src = vectors;
                if (vdev >= 0 && vdev != k.ix->device) {  // another GPU's rows: over NVLink first
                    VDB_TRY(k.xraw.reserve((size_t)n * dim));
                    VDB_CUDA_TRY(cudaMemcpyPeerAsync(k.xraw.p, k.ix->device, vectors, vdev, (size_t)n * dim * 4, st));
                    src = k.xraw.p;
                }
                if (vdev >= 0) {
                    VDB_TRY(launch_pad_rows(src, dim, dim, k.xbuf.p, ld, n, st));
                } else if (dim == ld) {
                    VDB_CUDA_TRY(cudaMemcpyAsync(k.xbuf.p, src, (size_t)n * ld * 4, cudaMemcpyHostToDevice, st));
                } else {
                    VDB_CUDA_TRY(cudaMemsetAsync(k.xbuf.p, 0, (size_t)n * ld * 4, st));
                    VDB_CUDA_TRY(cudaMemcpy2DAsync(k.xbuf.p, (size_t)ld * 4, src, (size_t)dim * 4, (size_t)dim * 4, n,
                                                   cudaMemcpyHostToDevice, st));
                }
                k.x = k.xbuf.p;
            }
            k.sc.mind_copies = 2;
            VDB_TRY(k.sc.reserve(n, nlist, ld));
            VDB_TRY(kmeanspp_dist_plan(k.x + (size_t)k.lo * ld, k.hi - k.lo, ld, dim, ld, &k.plan));
            // first seed (same generator state on every rank); both copies of the minima start at FLT_MAX
            VDB_TRY(kmeanspp_init(k.x, n, ld, ld, k.ix->centroids.p, k.sc, k.sc.mind, (uint64_t)2 * n, st));
        }
        // every rank's buffers exist and are initialised before any peer writes into them
        for (Rank& k : rk) {
            DeviceGuard g(k.ix->device);
            VDB_CUDA_TRY(cudaStreamSynchronize(k.ix->stream));
        }
        auto meet = [&]() -> int32_t {  // every stream continues only when every rank reached this point
            for (Rank& k : rk) {
                DeviceGuard g(k.ix->device);
                VDB_CUDA_TRY(cudaEventRecord(k.ev, k.ix->stream));
            }
            for (uint32_t r = 0; r < R; ++r) {
                DeviceGuard g(rk[r].ix->device);
                for (uint32_t p = 0; p < R; ++p)
                    if (p != r) VDB_CUDA_TRY(cudaStreamWaitEvent(rk[r].ix->stream, rk[p].ev, 0));
            }
            return VDB_OK;
        };
        const vdb_config& cfg = c->shards[0]->cfg;
        const SeedSampler sampler = cfg.train_mode == VDB_TRAIN_EXACT  ? SeedSampler::Sequential
                                    : cfg.train_mode == VDB_TRAIN_FAST ? SeedSampler::Fast
                                                                       : SeedSampler::ExactParallel;
        // ---- k-means++ (always L2, ivf_flat_index.cpp:63-104)
        for (uint32_t s = 1; s < nlist; ++s) {
            const uint32_t cur = s & 1u, prev = cur ^ 1u;
            for (Rank& k : rk) {
                DeviceGuard g(k.ix->device);
                PeerF32 out{};
                out.n = R;
                for (uint32_t p = 0; p < R; ++p) out.p[p] = rk[p].sc.mind + (size_t)cur * n + k.lo;
                VDB_TRY(kmeanspp_dist(k.plan, k.ix->centroids.p + (size_t)(s - 1) * ld,
                                      k.sc.mind + (size_t)prev * n + k.lo, out, k.ix->stream));
            }
            VDB_TRY(meet());
            for (Rank& k : rk) {
                DeviceGuard g(k.ix->device);
                VDB_TRY(kmeanspp_sample(k.x, n, ld, ld, k.sc.mind + (size_t)cur * n, k.ix->centroids.p, s, k.sc, sampler,
                                        k.ix->stream));
            }
        }
        // ---- slice assignment copied to every rank
        auto assign_all = [&]() -> int32_t {
            // one host thread per rank: the tensor-core assignment synchronises its stream once (overflow count),
            // which would otherwise serialise the ranks
            VDB_TRY(for_each_shard_parallel(c, [&](uint32_t r) -> int32_t {
                Rank& k = rk[r];
                DeviceGuard g(k.ix->device);
                if (k.hi > k.lo)
                    VDB_TRY(index_assign_rows(k.ix, k.x + (size_t)k.lo * ld, k.hi - k.lo, k.sc.assign + k.lo, k.ix->stream));
                for (Rank& p : rk)
                    if (&p != &k && k.hi > k.lo)
                        VDB_CUDA_TRY(cudaMemcpyPeerAsync(p.sc.assign + k.lo, p.ix->device, k.sc.assign + k.lo, k.ix->device,
                                                         (size_t)(k.hi - k.lo) * 4, k.ix->stream));
                return VDB_OK;
            }));
            return meet();
        };
        // ---- exactly 10 Lloyd iterations; assignment honours the index metric (:109-142, :275-285)
        for (int iter = 0; iter < 10; ++iter) {
            VDB_TRY(assign_all());
            for (Rank& k : rk) {
                DeviceGuard g(k.ix->device);
                VDB_TRY(kmeans_update_exact_range(k.x, n, ld, k.sc.assign, nlist, ld, k.ix->centroids.p, k.sc, k.c_lo, k.c_hi,
                                                  k.ix->stream));
                for (Rank& p : rk)
                    if (&p != &k && k.c_hi > k.c_lo)
                        VDB_CUDA_TRY(cudaMemcpyPeerAsync(p.ix->centroids.p + (size_t)k.c_lo * ld, p.ix->device,
                                                         k.ix->centroids.p + (size_t)k.c_lo * ld, k.ix->device,
                                                         (size_t)(k.c_hi - k.c_lo) * ld * 4, k.ix->stream));
            }
            VDB_TRY(meet());
        }
        // ---- list sizes of the sample under the final centroids -> byte-balanced ownership (as vdb_index_train)
        std::vector<uint32_t> counts(nlist + 1, 0);
        if (c->total_vectors == 0) {
            VDB_TRY(assign_all());
            Rank& k = rk[c->root];
            DeviceGuard g(k.ix->device);
            VDB_TRY(k.ix->hist_buf.reserve(nlist + 1));
            VDB_CUDA_TRY(cudaMemsetAsync(k.ix->hist_buf.p, 0, (size_t)(nlist + 1) * 4, k.ix->stream));
            VDB_TRY(launch_hist(k.sc.assign, n, nlist, 0, nullptr, k.ix->hist_buf.p, k.ix->stream));
            VDB_CUDA_TRY(cudaMemcpyAsync(counts.data(), k.ix->hist_buf.p, (size_t)(nlist + 1) * 4, cudaMemcpyDeviceToHost,
                                         k.ix->stream));
        }
        for (Rank& k : rk) {
            DeviceGuard g(k.ix->device);
            VDB_CUDA_TRY(cudaStreamSynchronize(k.ix->stream));
        }
        counts.resize(nlist);
        for (Rank& k : rk) {
            DeviceGuard g(k.ix->device);
            std::lock_guard<std::mutex> l2(k.ix->mu);
            VDB_TRY(index_refresh_centroids(k.ix));
            if (c->total_vectors == 0) {
                index_balance_owners(k.ix, counts);
                VDB_TRY(index_upload_owners(k.ix));
            }
            VDB_CUDA_TRY(cudaStreamSynchronize(k.ix->stream));
            k.ix->trained = true;
        }
        return VDB_OK;
    };
    const int32_t st = run();
    std::string msg = st == VDB_OK ? std::string() : std::string(vdb_last_error_string());
    cleanup();
    if (st != VDB_OK) set_last_error(msg);
    return st;
}

int32_t composite_set_centroids(vdb_index* ix, const float* in) {
    std::lock_guard<std::mutex> lock(ix->mu);
    std::vector<float> host;
    if (is_device_ptr(in)) {  // may live on any device: go through the host once
        host.resize((size_t)ix->cfg.nlist * ix->cfg.dimension);
        VDB_CUDA_TRY(cudaMemcpy(host.data(), in, host.size() * 4, cudaMemcpyDefault));
        in = host.data();
    }
    for (vdb_index* s : ix->composite->shards) VDB_TRY(vdb_index_set_centroids(s, in));
    return VDB_OK;
}

int32_t composite_set_owners(vdb_index* ix, const uint8_t* in) {
    std::lock_guard<std::mutex> lock(ix->mu);
    for (vdb_index* s : ix->composite->shards) VDB_TRY(vdb_index_set_owners(s, in));
    return VDB_OK;
}

__global__ void iota_u64_kernel(uint64_t* out, uint64_t base, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = base + i;
}

// add, data-parallel (the default when every pair of shard devices has peer access):
//   1. the batch is cut into one slice per shard; shard r stages ITS slice (H2D from the caller's host memory, or a
//      peer copy from the GPU that holds the rows) and assigns it on its tensor cores -- 1/R of the O(n nlist dim) work;
//   2. every shard then appends the rows of the lists it OWNS, slice by slice, reading rows, ids and assignments
//      straight out of the staging buffers of the shard that assigned them (NVLink peer loads inside the histogram
//      and scatter kernels): each row crosses NVLink once, to the GPU whose list it joins.
// Both phases run on all shards concurrently, one host thread per shard.
static int32_t composite_add_distributed(vdb_index* ix, const float* vectors, const uint64_t* ids, uint64_t n) {
    Composite* c = ix->composite;
    const uint32_t R = (uint32_t)c->shards.size(), dim = ix->cfg.dimension;
    const int vdev = device_of(vectors), idev = ids ? device_of(ids) : -1;
    const uint64_t per = (n + R - 1) / R;
    const uint64_t id_base = c->total_vectors;
    if (c->stage.size() != R) c->stage.resize(R);
    VDB_TRY(for_each_shard_parallel(c, [&](uint32_t r) -> int32_t {
        vdb_index* s = c->shards[r];
        Composite::Stage& st = c->stage[r];
        DeviceGuard g(s->device);
        const uint64_t lo = std::min(n, r * per), m = std::min(n, (r + 1) * per) - lo;
        st.n = m;
        if (m == 0) return VDB_OK;
        VDB_TRY(st.x.reserve(m * dim));
        VDB_TRY(st.ids.reserve(m));
        VDB_TRY(st.asg.reserve(m));
        const float* src = vectors + lo * dim;
        if (vdev < 0) VDB_CUDA_TRY(cudaMemcpyAsync(st.x.p, src, m * dim * 4, cudaMemcpyHostToDevice, s->stream));
        else if (vdev == s->device) VDB_CUDA_TRY(cudaMemcpyAsync(st.x.p, src, m * dim * 4, cudaMemcpyDeviceToDevice, s->stream));
        else VDB_CUDA_TRY(cudaMemcpyPeerAsync(st.x.p, s->device, src, vdev, m * dim * 4, s->stream));
        if (!ids) iota_u64_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, s->stream>>>(st.ids.p, id_base + lo, m);
        else if (idev < 0) VDB_CUDA_TRY(cudaMemcpyAsync(st.ids.p, ids + lo, m * 8, cudaMemcpyHostToDevice, s->stream));
        else if (idev == s->device) VDB_CUDA_TRY(cudaMemcpyAsync(st.ids.p, ids + lo, m * 8, cudaMemcpyDeviceToDevice, s->stream));
        else VDB_CUDA_TRY(cudaMemcpyPeerAsync(st.ids.p, s->device, ids + lo, idev, m * 8, s->stream));
        VDB_CUDA_TRY(cudaStreamSynchronize(s->stream));  // vdb_index_assign works on the same stream; be explicit
        return vdb_index_assign(s, st.x.p, m, st.asg.p);  // synchronises: the slice's assignments are complete
    }));
    VDB_TRY(for_each_shard_parallel(c, [&](uint32_t o) -> int32_t {
        vdb_index* s = c->shards[o];
        DeviceGuard g(s->device);
        for (uint32_t i = 0; i < R; ++i) {
            const Composite::Stage& st = c->stage[(o + i) % R];  // start with the own slice: spreads the peer reads
            if (st.n) VDB_TRY(vdb_index_add_assigned(s, st.x.p, st.ids.p, st.asg.p, st.n, st.n));
        }
        return VDB_OK;
    }));
    c->total_vectors += n;
    return VDB_OK;
}

// add, replicated (fallback): every shard assigns the whole batch and keeps the rows of the lists it owns; the shards
// work concurrently, one host thread each.  Rows that live on another device are staged chunk by chunk over NVLink.
int32_t composite_add(vdb_index* ix, const float* vectors, const uint64_t* ids, uint64_t n) {
    std::lock_guard<std::mutex> lock(ix->mu);
    Composite* c = ix->composite;
    if (c->shards.size() > 1 && c->peer_all && n >= 4096) return composite_add_distributed(ix, vectors, ids, n);
    const uint32_t dim = ix->cfg.dimension;
    const int vdev = device_of(vectors), idev = ids ? device_of(ids) : -1;
    const uint64_t chunk = std::max<uint64_t>(1024, (256ull << 20) / (dim * 4ull));
    VDB_TRY(for_each_shard_parallel(c, [&](uint32_t r) -> int32_t {
        vdb_index* s = c->shards[r];
        DeviceGuard g(s->device);
        const bool v_remote = vdev >= 0 && vdev != s->device, i_remote = idev >= 0 && idev != s->device;
        if (!v_remote && !i_remote) return vdb_index_add(s, vectors, ids, n);
        // ids must accompany staged rows: implicit ids are offsets from the index's running total, which a
        // chunked call would keep valid -- but explicit is simpler to reason about here
        DevBuf<float> vb;
        DevBuf<uint64_t> ib;
        int32_t st = VDB_OK;
        const uint64_t base = s->total_vectors;
        for (uint64_t lo = 0; lo < n && st == VDB_OK; lo += chunk) {
            const uint64_t m = std::min(chunk, n - lo);
            const float* v = vectors + lo * dim;
            if (v_remote) {
                if ((st = vb.reserve(m * dim)) != VDB_OK) break;
                // on the shard's own (non-blocking) stream, so that add()'s kernels are ordered behind the copy
                if (cudaMemcpyPeerAsync(vb.p, s->device, v, vdev, m * dim * 4, s->stream) != cudaSuccess) { st = VDB_CUDA_ERROR; break; }
                v = vb.p;
            }
            const uint64_t* id = ids ? ids + lo : nullptr;
            std::vector<uint64_t> seq;
            if (ids && i_remote) {
                if ((st = ib.reserve(m)) != VDB_OK) break;
                if (cudaMemcpyPeerAsync(ib.p, s->device, id, idev, m * 8, s->stream) != cudaSuccess) { st = VDB_CUDA_ERROR; break; }
                id = ib.p;
            } else if (!ids) {
                seq.resize(m);
                for (uint64_t i = 0; i < m; ++i) seq[i] = base + lo + i;
                id = seq.data();
            }
            st = vdb_index_add(s, v, id, m);
        }
        vb.release();
        ib.release();
        return st;
    }));
    c->total_vectors += n;
    return VDB_OK;
}

int32_t composite_note_added(vdb_index* ix, uint64_t n) {
    std::lock_guard<std::mutex> lock(ix->mu);
    ix->composite->total_vectors += n;
    return VDB_OK;
}

int32_t composite_reserve_search(vdb_index* ix, uint32_t nq, uint32_t np, uint32_t k) {
    std::lock_guard<std::mutex> lock(ix->mu);
    Composite* c = ix->composite;
    if (nq > c->max_nq || k > c->max_k) VDB_TRY(connect_mailboxes(c, std::max(nq, c->max_nq), std::max(k, c->max_k)));
    c->seen_nq = std::max(c->seen_nq, nq);
    c->seen_np = std::max(c->seen_np, np);
    c->seen_k = std::max(c->seen_k, k);
    for (vdb_index* s : c->shards) VDB_TRY(vdb_index_reserve_search(s, c->seen_nq, c->seen_np, c->seen_k));
    return VDB_OK;
}

// One search over all shards.  The shards' tickets advance in lockstep with the composite's (nobody else submits
// to them), so ticket t uses slot t % depth on every shard.
int32_t composite_submit(vdb_index* ix, const float* queries, uint32_t nq, uint32_t nprobe, uint32_t k,
                         float* distances, uint64_t* indices, uint64_t* ticket) {
    std::unique_lock<std::mutex> lock(ix->mu);
    Composite* c = ix->composite;
    const uint32_t world = (uint32_t)c->shards.size();
    VDB_REQUIRE((uint64_t)world * k <= 4096, "search: shards * k must be <= 4096");
    if (nq > c->max_nq || k > c->max_k) VDB_TRY(connect_mailboxes(c, std::max(nq, c->max_nq), std::max(k, c->max_k)));
    // a shape larger than any seen so far: every shard allocates its search scratch NOW, before anything of this
    // search is enqueued (see below why no allocation may happen in between)
    if (nq > c->seen_nq || nprobe > c->seen_np || k > c->seen_k) {
        c->seen_nq = std::max(c->seen_nq, nq);
        c->seen_np = std::max(c->seen_np, nprobe);
        c->seen_k = std::max(c->seen_k, k);
        for (vdb_index* s : c->shards) VDB_TRY(vdb_index_reserve_search(s, c->seen_nq, c->seen_np, c->seen_k));
    }
    // a batch whose partial results do not fit one pass on some shard is split evenly on all of them
    uint32_t chunk = nq;
    for (vdb_index* s : c->shards) {
        uint32_t ppi = 1, fit = nq;
        index_choose_ppi(s, nq, std::min(nprobe, s->nlist), k, &ppi, &fit);
        chunk = std::min(chunk, fit);
    }
    const int qdev = device_of(queries);
    vdb_index* root = c->shards[c->root];
    uint64_t last = 0;
    for (uint32_t lo = 0; lo < nq; lo += chunk) {
        const uint32_t m = std::min(chunk, nq - lo);
        const float* q = queries + (size_t)lo * ix->cfg.dimension;
        const uint64_t t = ++c->next_ticket;
        // The root goes LAST: its collect kernel spins until every shard has published, so nothing that can block
        // the host (a first-seen shape allocating scratch: cudaFree synchronises the device and, through the
        // peer mappings, its peers) may sit between enqueuing the collect and enqueuing the publishes it waits for.
        std::vector<SearchSlot*> slot(world, nullptr);
        for (uint32_t i = 0; i < world; ++i) {
            const uint32_t r = (c->root + 1 + i) % world;
            vdb_index* s = c->shards[r];
            DeviceGuard g(s->device);
            uint64_t ts = 0;
            VDB_TRY(index_acquire_slot(s, &slot[r], &ts));
            VDB_REQUIRE(ts == t, "sharded search: a shard was searched behind the composite's back");
        }
        // the two mailbox halves alternate by ticket: a non-root shard may overwrite half (t & 1) only after the
        // root has collected ticket t - 2 from it (with depth <= 2 recycling the root's slot has waited for that)
        cudaEvent_t collected = nullptr;
        if (t > 2 && root->depth > 2) {
            SearchSlot& prev = root->slots[(t - 2) % root->depth];
            if (prev.busy) collected = prev.ev_done;
        }
        for (uint32_t i = 0; i < world; ++i) {
            const uint32_t r = (c->root + 1 + i) % world;
            vdb_index* s = c->shards[r];
            DeviceGuard g(s->device);
            SearchStreams st = index_pipeline_streams(s, t);
            const float* qs = q;
            if (qdev >= 0 && qdev != s->device) {  // device queries of another GPU: peer copy on the front stream
                SearchSlot& sl = *slot[r];
                VDB_TRY(sl.q_raw.reserve((size_t)m * ix->cfg.dimension));
                VDB_CUDA_TRY(cudaMemcpyPeerAsync(sl.q_raw.p, s->device, q, qdev, (size_t)m * ix->cfg.dimension * 4,
                                                 st.front));
                qs = sl.q_raw.p;
            }
            const bool is_root = r == c->root;
            if (!is_root) st.back_wait = collected;
            VDB_TRY(index_enqueue_search(s, *slot[r], qs, m, nprobe, k, is_root ? distances + (size_t)lo * k : nullptr,
                                         is_root ? indices + (size_t)lo * k : nullptr, st, is_root));
        }
        last = t;
        if (chunk < nq) {  // chunked: one pass at a time
            lock.unlock();
            VDB_TRY(composite_wait(ix, t));
            lock.lock();
        }
    }
    *ticket = last;
    return VDB_OK;
}

int32_t composite_wait(vdb_index* ix, uint64_t ticket) {
    Composite* c = ix->composite;
    vdb_index* root = c->shards[c->root];
    SearchSlot* rs = nullptr;
    {
        std::lock_guard<std::mutex> lock(ix->mu);
        VDB_REQUIRE(ticket >= 1 && ticket <= c->next_ticket, "search_wait: unknown ticket");
        rs = &root->slots[ticket % root->depth];
        if (rs->ticket != ticket || !rs->busy) rs = nullptr;  // finished (and delivered) when its slot was recycled
        else VDB_TRY(index_flush_deferred(root, *rs));        // nobody submitted after it: its collect goes out now
    }
    if (rs) {
        DeviceGuard g(root->device);
        VDB_CUDA_TRY(cudaEventSynchronize(rs->ev_done));  // without the lock: other threads keep submitting
    }
    std::lock_guard<std::mutex> lock(ix->mu);
    int32_t st = VDB_OK;
    for (uint32_t i = 0; i < c->shards.size(); ++i) {
        vdb_index* s = c->shards[(c->root + i) % c->shards.size()];
        SearchSlot& sl = s->slots[ticket % s->depth];
        if (sl.ticket != ticket) continue;
        const int32_t r = index_finish_slot(s, sl);
        if (r != VDB_OK && st == VDB_OK) st = r;
    }
    return st;
}

int32_t composite_wait_stream(vdb_index* ix, uint64_t ticket, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(ix->mu);
    Composite* c = ix->composite;
    VDB_REQUIRE(ticket >= 1 && ticket <= c->next_ticket, "search_wait_stream: unknown ticket");
    vdb_index* root = c->shards[c->root];
    SearchSlot& sl = root->slots[ticket % root->depth];
    if (sl.ticket != ticket || !sl.busy) return VDB_OK;
    VDB_TRY(index_flush_deferred(root, sl));
    VDB_CUDA_TRY(cudaStreamWaitEvent(stream, sl.ev_done, 0));
    return VDB_OK;
}

int32_t composite_list_sizes(vdb_index* ix, uint64_t* out) {
    const uint32_t nlist = ix->cfg.nlist;
    std::vector<uint64_t> tmp(nlist);
    std::fill(out, out + nlist, 0);
    for (vdb_index* s : ix->composite->shards) {
        VDB_TRY(vdb_index_list_sizes(s, tmp.data()));
        for (uint32_t l = 0; l < nlist; ++l) out[l] += tmp[l];
    }
    return VDB_OK;
}

int32_t composite_list_ids(vdb_index* ix, uint32_t list, uint64_t* out) {
    VDB_REQUIRE(list < ix->cfg.nlist, "list_ids: bad arguments");
    std::vector<uint8_t> owners(ix->cfg.nlist);
    VDB_TRY(vdb_index_get_owners(composite_root(ix), owners.data()));
    return vdb_index_list_ids(ix->composite->shards[owners[list]], list, out);
}

int32_t composite_stats(vdb_index* ix, vdb_stats* out) {
    std::memset(out, 0, sizeof(*out));
    for (vdb_index* s : ix->composite->shards) {
        vdb_stats st;
        VDB_TRY(vdb_index_stats(s, &st));
        out->local_vectors += st.local_vectors;
        out->gpu_memory_bytes += st.gpu_memory_bytes;
        out->pages += st.pages;
        out->scanned_bytes += st.scanned_bytes;
        out->dimension = st.dimension; out->nlist = st.nlist; out->row_stride = st.row_stride;
        out->page_rows = st.page_rows;
        out->trained = st.trained;
        out->metric = st.metric;
    }
    out->total_vectors = ix->composite->total_vectors;
    return VDB_OK;
}

int32_t composite_last_search_stats(vdb_index* ix, vdb_search_stats* out) {
    std::memset(out, 0, sizeof(*out));
    for (vdb_index* s : ix->composite->shards) {
        vdb_search_stats st;
        VDB_TRY(vdb_index_last_search_stats(s, &st));
        out->algorithmic_rows += st.algorithmic_rows;
        out->unique_rows += st.unique_rows;
        out->scan_items += st.scan_items;
        out->bytes_per_row = st.bytes_per_row;
        out->streamed_bytes_per_row = st.streamed_bytes_per_row;
        out->rescored_pairs += st.rescored_pairs;
        out->scan_ctas = st.scan_ctas;
    }
    return VDB_OK;
}

int32_t composite_set_profiling(vdb_index* ix, int32_t enable) {
    for (vdb_index* s : ix->composite->shards) VDB_TRY(vdb_index_set_profiling(s, enable));
    return VDB_OK;
}

// per phase: the slowest shard's sum (the shards run concurrently); collect runs on the root only
int32_t composite_read_profile(vdb_index* ix, float* out_ms, uint32_t* searches) {
    for (int i = 0; i < 8; ++i) out_ms[i] = 0.f;
    *searches = 0;
    for (vdb_index* s : ix->composite->shards) {
        float ms[8];
        uint32_t n = 0;
        VDB_TRY(vdb_index_read_profile(s, ms, &n));
        for (int i = 0; i < 8; ++i) out_ms[i] = std::max(out_ms[i], ms[i]);
        *searches = std::max(*searches, n);
    }
    return VDB_OK;
}

}  // namespace vdb

using namespace vdb;

extern "C" int32_t vdb_index_create_sharded(const vdb_config* cfg, const int32_t* devices, int32_t ndev,
                                            vdb_index** out) {
    VDB_REQUIRE(cfg && devices && out, "create_sharded: null argument");
    VDB_REQUIRE(ndev >= 1 && ndev <= (int32_t)EX_MAX_WORLD, "create_sharded: 1 to 16 shards");
    std::unique_ptr<vdb_index> ix(new vdb_index());
    std::unique_ptr<Composite> c(new Composite());
    ix->cfg = *cfg;
    ix->cfg.shard_rank = 0;
    ix->cfg.shard_count = (uint32_t)ndev;
    ix->dim = cfg->dimension;
    ix->nlist = cfg->nlist;
    auto fail = [&](int32_t st) {
        const std::string msg = vdb_last_error_string();
        for (vdb_index* s : c->shards) vdb_index_destroy(s);
        set_last_error(msg);
        return st;
    };
    for (int32_t r = 0; r < ndev; ++r) {
        vdb_config sc = *cfg;
        sc.device = devices[r];
        sc.shard_rank = (uint32_t)r;
        sc.shard_count = (uint32_t)ndev;
        vdb_index* s = nullptr;
        const int32_t st = vdb_index_create(&sc, &s);
        if (st != VDB_OK) return fail(st);
        c->shards.push_back(s);
    }
    c->peer_all = true;
    for (int32_t a = 0; a < ndev; ++a)
        for (int32_t b = 0; b < ndev; ++b) {
            if (devices[a] == devices[b]) continue;
            DeviceGuard g(devices[a]);
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[a], devices[b]) != cudaSuccess || !can) {
                cudaGetLastError();
                c->peer_all = false;
                continue;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) c->peer_all = false;
            cudaGetLastError();
        }
    const int32_t st = connect_mailboxes(c.get(), 256, 64);
    if (st != VDB_OK) {
        for (vdb_exchange* e : c->exchanges) vdb_exchange_destroy(e);
        for (vdb_index* s : c->shards) s->exchange = nullptr;
        return fail(st);
    }
    ix->composite = c.release();
    *out = ix.release();
    return VDB_OK;
}
