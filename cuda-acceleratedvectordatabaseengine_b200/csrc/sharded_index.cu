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

namespace vdb {

struct Composite {
    std::vector<vdb_index*> shards;
    std::vector<vdb_exchange*> exchanges;
    uint32_t root = 0;
    uint32_t max_nq = 0, max_k = 0;  // mailbox shape (grown on demand between searches)
    uint64_t next_ticket = 0;
    uint64_t total_vectors = 0;
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
    for (vdb_exchange* e : c->exchanges) vdb_exchange_destroy(e);
    for (vdb_index* s : c->shards) vdb_index_destroy(s);
    delete c;
    ix->composite = nullptr;
    delete ix;
    return VDB_OK;
}

// train: k-means on the root (bit-exact, ivf_flat_index.cpp:49-145), centroids and the byte-balanced owner table
// copied to the other shards
int32_t composite_train(vdb_index* ix, const float* vectors, uint64_t n) {
    std::lock_guard<std::mutex> lock(ix->mu);
    Composite* c = ix->composite;
    vdb_index* root = c->shards[c->root];
    VDB_TRY(vdb_index_train(root, vectors, n));
    std::vector<float> cent((size_t)ix->cfg.nlist * ix->cfg.dimension);
    std::vector<uint8_t> owners(ix->cfg.nlist);
    VDB_TRY(vdb_index_get_centroids(root, cent.data()));
    VDB_TRY(vdb_index_get_owners(root, owners.data()));
    for (uint32_t r = 0; r < c->shards.size(); ++r) {
        if (r == c->root) continue;
        VDB_TRY(vdb_index_set_centroids(c->shards[r], cent.data()));
        if (c->total_vectors == 0) VDB_TRY(vdb_index_set_owners(c->shards[r], owners.data()));
    }
    return VDB_OK;
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

// add: every shard assigns the batch (tensor cores) and keeps the rows of the lists it owns; the shards work
// concurrently, one host thread each.  Rows that live on another device are staged chunk by chunk over NVLink.
int32_t composite_add(vdb_index* ix, const float* vectors, const uint64_t* ids, uint64_t n) {
    std::lock_guard<std::mutex> lock(ix->mu);
    Composite* c = ix->composite;
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
                if (cudaMemcpyPeer(vb.p, s->device, v, vdev, m * dim * 4) != cudaSuccess) { st = VDB_CUDA_ERROR; break; }
                v = vb.p;
            }
            const uint64_t* id = ids ? ids + lo : nullptr;
            std::vector<uint64_t> seq;
            if (ids && i_remote) {
                if ((st = ib.reserve(m)) != VDB_OK) break;
                if (cudaMemcpyPeer(ib.p, s->device, id, idev, m * 8) != cudaSuccess) { st = VDB_CUDA_ERROR; break; }
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
    for (vdb_index* s : c->shards) VDB_TRY(vdb_index_reserve_search(s, nq, np, k));
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
        // root first: recycling its slot waits (on the host) for the collect that used it `depth` tickets ago
        std::vector<SearchSlot*> slot(world, nullptr);
        for (uint32_t i = 0; i < world; ++i) {
            const uint32_t r = (c->root + i) % world;
            vdb_index* s = c->shards[r];
            DeviceGuard g(s->device);
            uint64_t ts = 0;
            VDB_TRY(index_acquire_slot(s, &slot[r], &ts));
            VDB_REQUIRE(ts == t, "sharded search: a shard was searched behind the composite's back");
        }
        // the two mailbox halves alternate by ticket: a non-root shard may overwrite half (t & 1) only after the
        // root has collected ticket t - 2 from it
        cudaEvent_t collected = nullptr;
        if (t > 2 && root->depth > 2) {
            SearchSlot& prev = root->slots[(t - 2) % root->depth];
            if (prev.busy) collected = prev.ev_done;
        }
        for (uint32_t i = 0; i < world; ++i) {
            const uint32_t r = (c->root + i) % world;
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
