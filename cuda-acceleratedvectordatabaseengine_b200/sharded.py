"""Multi-GPU host side: inverted lists partitioned by list over the ranks of one
torch.distributed process group (one process per GPU), per-rank top-k merged
after an all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference has no multi-GPU path at all ("Parallelism"); the
partitioning follows from merge_results (ivf_flat_index.cpp:474-518): a query's
answer is the top-k of the union of per-list top-k's, so lists are independent
units and the only exchange step is the final merge.

    owner[l] = l % world until train(), which re-balances the lists over the ranks by bytes
    (greedy, largest first, identical table on every rank; IVFFlatIndex.owners())
    every rank: same centroids, same coarse selection on the full query batch,
                scan of the probed lists it owns -> local [nq][k] padded FLT_MAX/UINT64_MAX
    exchange + merge by (distance, id), duplicates dropped:
      exchange="p2p"   one kernel per rank over NVLink peer memory (PeerExchange / csrc/exchange.cu) -- the default
                       on a CUDA group with peer access
      exchange="nccl"  all_gather of the local (distances, ids) (gather_topk) + vdb_merge_topk kernel

The result is independent of the number of ranks by construction (same
tie-break everywhere); tests/test_sharded_gloo.py checks exactly that.
"""
import torch
import torch.distributed as dist


def owner_of(list_id, world):
    """Default owner of inverted list `list_id` (the table an untrained sharded index starts from)."""
    return int(list_id) % int(world)


def gather_topk(D, I, group=None):
    """all_gather of every rank's local top-k: [nq][k] -> [world][nq][k] (same order on every rank)."""
    world = dist.get_world_size(group)
    Dg = torch.empty((world,) + tuple(D.shape), dtype=D.dtype, device=D.device)
    Ig = torch.empty((world,) + tuple(I.shape), dtype=I.dtype, device=I.device)
    if D.is_cuda:  # NCCL: one fused collective per tensor, straight into the [world][nq][k] buffer
        dist.all_gather_into_tensor(Dg, D.contiguous(), group=group)
        dist.all_gather_into_tensor(Ig, I.contiguous(), group=group)
    else:  # gloo (CPU tests): list form, views of the same buffer
        dist.all_gather(list(Dg.unbind(0)), D.contiguous(), group=group)
        dist.all_gather(list(Ig.unbind(0)), I.contiguous(), group=group)
    return Dg, Ig


class PeerExchange:
    """vdb_exchange_*: the per-rank mailboxes are mapped into every peer with CUDA IPC handles gathered over the
    process group; merge_topk() is then a single kernel launch per rank (publish -> wait -> merge)."""

    def __init__(self, pkg, device, group=None, max_nq=1024, max_k=64):
        import ctypes as C
        self.pkg, self.group = pkg, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.max_nq, self.max_k = max_nq, max_k
        # every rank runs every collective below whatever happens locally, so that a rank that cannot create or map
        # a mailbox makes ALL ranks raise instead of leaving the others stuck in a collective
        self._h, err, mine = None, None, None
        try:
            h = C.c_void_p()
            pkg._check(pkg.lib().vdb_exchange_create(device, self.rank, self.world, max_nq, max_k, C.byref(h)))
            self._h = h
            blob = (C.c_uint8 * 64)()
            pkg._check(pkg.lib().vdb_exchange_handle(self._h, blob))
            mine = bytes(blob)
        except Exception as e:  # noqa: BLE001
            err = str(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        if err is None and all(x is not None for x in handles):
            try:
                buf = (C.c_uint8 * (64 * self.world)).from_buffer_copy(b"".join(handles))
                pkg._check(pkg.lib().vdb_exchange_connect(self._h, buf))
            except Exception as e:  # noqa: BLE001
                err = str(e)
        elif err is None:
            err = "a peer could not create its mailbox"
        oks = [None] * self.world
        dist.all_gather_object(oks, err is None, group=group)  # also: every peer has mapped every mailbox
        if not all(oks):
            self.close()
            raise RuntimeError("peer exchange unavailable: " + (err or "a peer could not map the mailboxes"))

    def merge_topk(self, D, I, stream=0):
        """local [nq][k] CUDA tensors -> merged ([nq][k] f32, [nq][k] i64); collective, same order on all ranks"""
        nq, k = D.shape
        Do = torch.empty((nq, k), dtype=torch.float32, device=D.device)
        Io = torch.empty((nq, k), dtype=torch.int64, device=D.device)
        self.merge_topk_into(D, I, Do, Io, stream)
        return Do, Io

    def merge_topk_into(self, D, I, Do, Io, stream=0):
        nq, k = D.shape
        self.pkg._check(self.pkg.lib().vdb_exchange_merge_topk(self._h, D.data_ptr(), I.data_ptr(), nq, k,
                                                               Do.data_ptr(), Io.data_ptr(), stream))

    def publish(self, D, I, stream=0):
        """first half of merge_topk: store this rank's local [nq][k] block into every peer's mailbox"""
        nq, k = D.shape
        self.pkg._check(self.pkg.lib().vdb_exchange_publish(self._h, D.data_ptr(), I.data_ptr(), nq, k, stream))

    def collect_into(self, Do, Io, stream=0):
        """second half: wait for the peers' blocks of the published batch and merge them into Do, Io"""
        self.pkg._check(self.pkg.lib().vdb_exchange_collect(self._h, Do.data_ptr(), Io.data_ptr(), stream))

    def check(self):
        """raise if a call timed out waiting for a peer (valid after synchronising with the call's stream)"""
        self.pkg._check(self.pkg.lib().vdb_exchange_status(self._h))

    def reset(self):
        self.pkg._check(self.pkg.lib().vdb_exchange_reset(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.pkg.lib().vdb_exchange_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedIVFFlatIndex:
    """vdb::IVFFlatIndex surface over `world` list shards, one per rank.

    train(): every rank trains on the same rows (deterministic -> identical centroids and owner table);
    add():   every rank sees the batch, assigns it, and keeps only the rows of the lists it owns;
    add_distributed(): each rank assigns its own slice, one all-to-all routes rows to the owning ranks;
    search(): collective; every rank returns the full answer.  With the peer-memory exchange the index itself is
              collective (vdb_index_attach_exchange): the shard's merge kernel stores its top-k into the peers'
              mailboxes and a collect kernel merges the world's blocks, pipelined over batches by search_submit /
              search_wait.  Otherwise: local search + NCCL all-gather + merge kernel."""

    def __init__(self, pkg, config, group=None, exchange="auto", max_nq=1024, max_k=64):
        self.pkg, self.group = pkg, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        config.shard_rank, config.shard_count = self.rank, self.world
        self.local = pkg.IVFFlatIndex(config)
        self.config = config
        # exchange: "p2p" (peer-memory kernel), "nccl" (all-gather + merge kernel), "auto" = p2p on an NCCL group
        # of more than one rank whose (world * max_k) fits the mailbox merge
        auto = exchange == "auto"
        if auto:
            exchange = "p2p" if (self.world > 1 and torch.cuda.is_available() and
                                 dist.get_backend(group) == "nccl" and self.world * max_k <= 4096) else "nccl"
        self.exchange = None
        if exchange == "p2p":
            try:
                self.exchange = PeerExchange(pkg, config.device, group, max_nq, max_k)  # raises on all ranks or none
            except RuntimeError:
                if not auto:
                    raise  # asked for explicitly: report it; "auto" falls back to the all-gather path
        self._attached = False
        self._synced = False
        self._set_attached(self.exchange is not None)

    def _set_attached(self, on):
        if on != self._attached:
            self.local.attach_exchange(self.exchange._h if on else None)
            self._attached = on

    def _sync_once(self):
        # ranks finish building at different times and the exchange kernel gives a peer 20 s to show up: meet
        # once after every build step before the first collective search
        if not self._synced and self.world > 1:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            dist.barrier(group=self.group)
            self._synced = True

    def train(self, vectors):
        self.local.train(vectors)
        self._synced = False

    def add(self, vectors, ids=None):
        self.local.add(vectors, ids)
        self._synced = False

    def add_distributed(self, vectors, ids):
        """Data-parallel add: every rank passes ITS OWN slice of the batch (CUDA tensors [n_r][dim] f32 and
        [n_r] int64 ids).  Each rank assigns only its slice (tensor cores), rows are routed to the ranks that
        own their lists with one all-to-all over NVLink, and land in the owners' HBM pages."""
        world = self.world
        self._synced = False
        a = self.local.assign_device(vectors)
        if world == 1:
            self.local.add_assigned(vectors.contiguous(), ids.contiguous(), a, vectors.shape[0])
            return
        if not hasattr(self, "_owners_dev"):
            self._owners_dev = torch.from_numpy(self.local.owners().astype("int64")).to(vectors.device)
        dest = self._owners_dev[a.long()]
        order = torch.argsort(dest, stable=True)
        send_counts = torch.bincount(dest, minlength=world)
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        sc, rc = send_counts.tolist(), recv_counts.tolist()
        nrecv, dim = int(sum(rc)), vectors.shape[1]
        rv = torch.empty((nrecv, dim), dtype=torch.float32, device=vectors.device)
        ri = torch.empty(nrecv, dtype=torch.int64, device=vectors.device)
        ra = torch.empty(nrecv, dtype=torch.int32, device=vectors.device)
        dist.all_to_all_single(rv, vectors[order].contiguous(), rc, sc, group=self.group)
        dist.all_to_all_single(ri, ids[order].contiguous(), rc, sc, group=self.group)
        dist.all_to_all_single(ra, a[order].contiguous(), rc, sc, group=self.group)
        total = torch.tensor([vectors.shape[0]], dtype=torch.int64, device=vectors.device)
        dist.all_reduce(total, group=self.group)
        self.local.add_assigned(rv, ri, ra, int(total.item()))

    def _fits_mailbox(self, nq, k):
        ex = self.exchange
        return ex is not None and nq <= ex.max_nq and k <= ex.max_k

    def search_device(self, queries, nprobe, k, stream=None):
        """queries: float32 CUDA tensor [nq][dim]; returns CUDA tensors ([nq][k] f32, [nq][k] i64 view of u64 ids)."""
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        queries = self.local._rows(queries)
        nq = queries.shape[0]
        D = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
        self._sync_once()
        if self.world == 1 or self._fits_mailbox(nq, k):
            self._set_attached(self.exchange is not None)
            self.local.search_async(queries, nprobe, k, D, I, s)  # collective inside when the exchange is attached
            return D, I
        self._set_attached(False)
        self.local.search_async(queries, nprobe, k, D, I, s)
        Dg, Ig = gather_topk(D, I, self.group)
        return self.pkg.merge_topk(Dg, Ig, s)

    def search_submit(self, queries, nprobe, k, distances, indices):
        """pipelined collective search (same order on every rank); see IVFFlatIndex.search_submit"""
        nq = queries.shape[0]
        if self.world > 1 and not self._fits_mailbox(nq, k):
            raise ValueError("search_submit on a sharded index needs the peer-memory exchange and nq, k within its mailbox")
        self._sync_once()
        self._set_attached(self.exchange is not None)
        return self.local.search_submit(queries, nprobe, k, distances, indices)

    def search_wait(self, ticket):
        self.local.search_wait(ticket)

    def search(self, queries, nprobe, k):
        """host (numpy, any float dtype / layout) or device queries in, numpy out on every rank"""
        if hasattr(queries, "data_ptr"):
            q = self.local._rows(queries.float())
        else:
            q = torch.from_numpy(self.local._rows(queries))  # float32, contiguous, [nq][dim]
        D, I = self.search_device(q.cuda(non_blocking=True), nprobe, k)
        torch.cuda.current_stream().synchronize()
        if self.exchange is not None:
            self.exchange.check()  # a peer that never showed up: raise instead of returning padding
        return D.cpu().numpy(), I.cpu().numpy().view("uint64")

    def get_total_vectors(self):
        return self.local.get_total_vectors()

    def get_gpu_memory_usage(self):
        t = torch.tensor([self.local.get_gpu_memory_usage()], dtype=torch.int64,
                         device="cuda" if torch.cuda.is_available() else "cpu")
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def close(self):
        if getattr(self, "local", None) is not None:
            if self._attached:
                self._set_attached(False)
            self.local.close()
        if self.exchange is not None:
            self.exchange.close()
            self.exchange = None
