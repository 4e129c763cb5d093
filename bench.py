#!/usr/bin/env python
"""Benchmark of the IVF-Flat search hot path (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    torchrun ... bench.py --gpus N --steps K --warmup W      # lists sharded over N GPUs, NCCL all-gather merge
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path, host cores

A step = one search of one batch of 64 fresh synthetic queries against the
10M x 768D L2 index (nlist 4096, nprobe 32, k 10) -- BASELINE.json configs[2],
the configuration the metric is quoted on; it fits one GPU (31 GB of 180).
Every batch streams far more list data than the 126 MB L2 holds, so no L2
flush is needed between iterations.

value   device-resident: queries and results in HBM, CUDA events, max over ranks
e2e     host buffers through the C ABI call a user makes (vdb_index_search):
        pinned H2D of the queries and D2H of ids+distances inside the timed region
roofline  list-scan kernel: algorithmic bytes (sum over queries of probed rows x (4D+8),
        SURVEY.md 8d) / its CUDA-event time, against the measured HBM copy bandwidth
cpu_baseline  the reference's unmodified CPU search (oracle/_ref) on a bounded sample
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
PKG = "cuda-acceleratedvectordatabaseengine_b200"
METRIC = "IVF-Flat QPS @10M x 768D nprobe=32 k=10"


def metric_name(a):
    if (a.n, a.dim, a.nprobe, a.k) == (10_000_000, 768, 32, 10):
        return METRIC
    return f"IVF-Flat QPS @{a.n / 1e6:g}M x {a.dim}D nprobe={a.nprobe} k={a.k}"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--rows", dest="n", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--metric", default="l2", choices=["l2", "ip"])
    ap.add_argument("--nlist", type=int, default=4096)
    ap.add_argument("--nprobe", type=int, default=32)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ntrain", type=int, default=262144)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: peer-memory mailboxes written by the merge kernel + collect kernel, pipelined over "
                         "batches (default), or the portable form: NCCL all-gathers + merge kernel per batch")
    ap.add_argument("--depth", type=int, default=4, help="batches in flight (vdb_index_search_submit)")
    ap.add_argument("--emulate-shards", type=int, default=1,
                    help="tuning aid, 1 GPU: hold only shard 0 of W (what one rank of a W-GPU run scans), no exchange")
    ap.add_argument("--emulate-rank", type=int, default=0, help="with --emulate-shards: which shard this GPU holds")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--parity-mode", default="auto", choices=["auto", "unsharded", "truth"],
                    help="N > 1, untimed: compare the sharded answer with an unsharded index on rank 0 (auto: when it "
                         "fits) or with a torch truth kept on rank 0 (auto: otherwise)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-n", type=int, default=1_000_000)
    ap.add_argument("--dump-groups", action="store_true", help="stderr: probed-row work by queries-per-list")
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json configs[0..4]: c1 IVF 100Kx128 (reference's CPU-runnable case), c2 brute force "
                         "1Mx768 k=100, c3 the headline (default), c4 100Mx768 IP nlist 16384 (8 GPUs), c5 k-means train "
                         "+ add 10Mx768 nlist 16384 (single process, --devices)")
    ap.add_argument("--devices", default="", help="c5: comma-separated devices of the single-process sharded index")
    a = ap.parse_args()
    presets = {"c1": dict(n=100_000, dim=128, nlist=128, nprobe=16, ntrain=10_000, metric="l2"),
               "c4": dict(n=100_000_000, dim=768, nlist=16384, nprobe=64, ntrain=1_000_000, metric="ip")}
    if a.config in presets:
        for k_, v in presets[a.config].items():
            setattr(a, k_, v)
    return a


def workload_name(a):
    shape = (a.n, a.dim, a.metric, a.nlist, a.nprobe, a.k, a.batch)
    tag = {(10_000_000, 768, "l2", 4096, 32, 10, 64): " (BASELINE.json configs[2])",
           (100_000_000, 768, "ip", 16384, 64, 10, 64): " (BASELINE.json configs[3])",
           (100_000, 128, "l2", 128, 16, 10, 64): " (BASELINE.json configs[0])"}.get(shape, "")
    return (f"IVF-Flat {a.n / 1e6:g}M x {a.dim}D {'L2' if a.metric == 'l2' else 'inner product'} nlist={a.nlist} "
            f"nprobe={a.nprobe} k={a.k} batch={a.batch}{tag}")


# ---------------------------------------------------------------- clocks

class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[5 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def profiled_traffic(mirror=False):
    """dram__bytes_read.sum + dram__bytes_write.sum of one scan launch from the committed ncu --set full capture of
    this same workload (profiles/scan_traffic.json: the fp32 scan_kernel; profiles/screen_traffic.json: the bf16
    screen_kernel; written from the .ncu-rep by tools/ncu_summary.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "screen_traffic.json" if mirror else "scan_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------- reference CPU arm

def reference_sample(a, steps, warmup, max_seconds=150.0):
    """The reference's CPU IVFFlatIndex::search (unmodified, oracle/_ref) on a bounded sample of the
    workload: cpu_sample_n x dim rows, nlist scaled to keep the mean list length, same nprobe/k, every
    host thread searching its own queries.  Index construction is NOT timed and uses numpy only."""
    import numpy as np
    import oracle_lib as O

    ncores = os.cpu_count() or 1
    ns = min(a.cpu_sample_n, a.n)
    nlist_s = max(a.nprobe, int(round(a.nlist * ns / a.n)))
    rng = np.random.default_rng(12345)
    x = np.empty((ns, a.dim), np.float32)
    for lo in range(0, ns, 100_000):
        x[lo:lo + 100_000] = rng.standard_normal((min(100_000, ns - lo), a.dim), dtype=np.float32)
    # centroids: sampled rows refined by two Lloyd passes on a subsample (BLAS), then one assignment of all rows
    cent = x[rng.choice(ns, nlist_s, replace=False)].copy()
    sub = x[: min(ns, 64 * nlist_s)]

    def assign(v, c):
        out = np.empty(v.shape[0], np.uint32)
        cn = (c * c).sum(1)
        for lo in range(0, v.shape[0], 50_000):
            out[lo:lo + 50_000] = np.argmin(cn[None, :] - 2.0 * (v[lo:lo + 50_000] @ c.T), axis=1)
        return out

    for _ in range(2):
        asg = assign(sub, cent)
        cnt = np.bincount(asg, minlength=nlist_s)
        sums = np.zeros_like(cent)
        np.add.at(sums, asg, sub)
        m = cnt > 0
        cent[m] = sums[m] / cnt[m, None]
    asg = assign(x, cent)
    kind = "reference" if O.ref_lib() is not None else "port"
    ix = (O.RefIndex if kind == "reference" else O.OracleIndex)(a.dim, nlist_s, O.METRIC_L2)
    ix.centroids = cent
    ix.load_assigned(x, np.arange(ns, dtype=np.uint64), asg)
    sizes = ix.list_sizes().astype(np.int64)
    nq_pool = a.batch * 4
    q = rng.standard_normal((nq_pool, a.dim), dtype=np.float32)
    # calibrate one query on one thread, then size the per-step sample
    t = time.perf_counter()
    ix.search(q[:1], a.nprobe, a.k, 1)
    t_q = max(time.perf_counter() - t, 1e-4)
    per_step = int(max_seconds * ncores / ((steps + warmup) * t_q))
    per_step = max(min(per_step, a.batch), min(ncores, a.batch), 1)
    probes = np.stack([ix.select_nprobe(q[i], a.nprobe) for i in range(min(nq_pool, 32))])
    rows_per_query = float(sizes[probes].sum(1).mean())

    def step(s, threads):
        lo = (s * per_step) % (nq_pool - per_step + 1)
        ix.search(q[lo:lo + per_step], a.nprobe, a.k, threads)

    for s in range(warmup):
        step(s, ncores)
    t = time.perf_counter()
    for s in range(steps):
        step(warmup + s, ncores)
    dt = time.perf_counter() - t
    qps_all = per_step * steps / dt
    # single-thread figure: the reference path as shipped (it has no threading)
    n1 = max(1, min(per_step, int(10.0 / t_q)))
    t = time.perf_counter()
    ix.search(q[:n1], a.nprobe, a.k, 1)
    qps_1 = n1 / (time.perf_counter() - t)
    sample = (f"{ns} x {a.dim}D Gaussian, nlist={nlist_s} (mean list {ns / nlist_s:.0f} rows as in the full workload), "
              f"nprobe={a.nprobe}, k={a.k}, {per_step} queries/step, {rows_per_query:.0f} probed rows/query; "
              f"single-thread {qps_1:.2f} QPS")
    return {"value": qps_all, "unit": "queries/s", "cores": ncores, "kind": kind, "sample": sample,
            "single_thread_qps": qps_1, "rows_per_query": rows_per_query, "ms_per_step": dt / steps * 1e3,
            "queries_per_step": per_step}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = reference_sample(a, a.steps, a.warmup)
    line = {"impl": "reference", "metric": metric_name(a), "value": r["value"], "unit": "queries/s", "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a)},
            "cpu_baseline": {"value": r["value"], "unit": "queries/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------- B200 arm

def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this implementation has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL's init lines (ranks, transports) are wanted -- on stderr: stdout carries exactly one JSON line (fd 1 is
        # diverted to stderr for the whole run, see _divert_stdout).  A weaker inherited level (the image exports
        # NCCL_DEBUG=VERSION) is raised to INFO / INIT; an inherited INFO or TRACE is left alone.
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        dist.init_process_group("nccl", device_id=dev)
    pkg = importlib.import_module(PKG)
    pkg.lib()

    # ---- build: synthetic N(0,1) rows generated on the device, trained + added by the CUDA path
    t_build = time.perf_counter()
    shard_rank, shard_count = (a.emulate_rank, a.emulate_shards) if a.emulate_shards > 1 and world == 1 else (rank, world)
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=a.dim, nlist=a.nlist, metric=pkg.Metric.L2 if a.metric == "l2" else pkg.Metric.InnerProduct, device=local,
                                     shard_rank=shard_rank, shard_count=shard_count, pipeline_depth=a.depth))
    gen = torch.Generator(device=dev).manual_seed(12345)
    chunk = 1_000_000
    t_train = t_add = 0.0
    # N > 1: rank 0 also holds an UNSHARDED index of the same rows (if it fits next to its shard) -- the untimed
    # parity check below compares the sharded answer with it
    ref = None
    fits = a.n * (a.dim * 4 + 8) * 1.2 < 120e9
    use_truth = world > 1 and not a.no_parity_check and (a.parity_mode == "truth" or (a.parity_mode == "auto" and not fits))
    if world > 1 and rank == 0 and not a.no_parity_check and not use_truth:
        ref = pkg.IVFFlatIndex(pkg.Config(dimension=a.dim, nlist=a.nlist, device=local, metric=pkg.Metric.L2 if a.metric == "l2" else pkg.Metric.InnerProduct))
    # ... and where that copy does not fit (configs[3]: 307 GB), rank 0 keeps an independent TRUTH for a few check
    # queries instead: their distance to every row, evaluated by torch while the rows stream past, and every row's
    # list (4 bytes per row each) -- enough to state exactly what an nprobe search and an exhaustive search must return
    truth = None
    nchk = 8
    qchk = torch.randn(nchk, a.dim, generator=torch.Generator(device=dev).manual_seed(999), device=dev)
    if use_truth and rank == 0:
        truth = {"dist": torch.empty((nchk, a.n), dtype=torch.float32, device=dev),
                 "asg": torch.empty(a.n, dtype=torch.int32, device=dev), "vmax": 0.0}
    for lo in range(0, a.n, chunk):
        x = torch.randn(min(chunk, a.n - lo), a.dim, generator=gen, device=dev)
        torch.cuda.current_stream().synchronize()  # the library works on its own streams: hand it finished rows
        if lo == 0:
            t = time.perf_counter()
            ix.train(x[: min(a.ntrain, x.shape[0])])  # every rank trains on the same rows: identical centroids
            t_train = time.perf_counter() - t
            if ref is not None:
                ref.centroids = ix.centroids
        t = time.perf_counter()
        ix.add(x)  # ids = row numbers; a sharded index keeps only the lists it owns
        t_add += time.perf_counter() - t
        if ref is not None:
            ref.add(x)
        if truth is not None:
            m = x.shape[0]
            if a.metric == "ip":
                truth["dist"][:, lo:lo + m] = -(qchk @ x.T)
            else:
                for c0 in range(0, m, 131072):
                    xc = x[c0:c0 + 131072]
                    truth["dist"][:, lo + c0:lo + c0 + xc.shape[0]] = ((xc[None, :, :] - qchk[:, None, :]) ** 2).sum(-1)
            truth["asg"][lo:lo + m] = ix.assign_device(x)
            truth["vmax"] = max(truth["vmax"], float(x.norm(dim=1).max()))
        del x
    nb = a.warmup + a.steps
    q_all = torch.randn(nb, a.batch, a.dim, generator=gen, device=dev)
    t_build = time.perf_counter() - t_build
    st = ix.stats()

    if a.dump_groups and rank == 0:
        pr = ix.select_nprobe(q_all[0], a.nprobe)
        sizes = ix.list_sizes().astype(np.int64)
        g = np.bincount(pr.ravel(), minlength=a.nlist)
        edges = [1, 2, 3, 5, 9, 17, 33, 65]
        tot = float((sizes * g).sum())
        print("list sizes: min %d med %d mean %.0f max %d" % (sizes.min(), np.median(sizes), sizes.mean(), sizes.max()),
              file=sys.stderr)
        for lo, hi in zip(edges[:-1], edges[1:]):
            m = (g >= lo) & (g < hi)
            print(f"queries/list in [{lo},{hi}): lists {int(m.sum())}, unique rows {int(sizes[m].sum())}, "
                  f"pair-rows share {float((sizes[m] * g[m]).sum()) / tot:.3f}", file=sys.stderr)

    stream = torch.cuda.current_stream().cuda_stream
    depth = a.depth
    Dd = [torch.empty((a.batch, a.k), dtype=torch.float32, device=dev) for _ in range(depth)]
    Id = [torch.empty((a.batch, a.k), dtype=torch.int64, device=dev) for _ in range(depth)]
    exch = None
    if world > 1:
        sharded = importlib.import_module(PKG + ".sharded")
        Dg = torch.empty((world, a.batch, a.k), dtype=torch.float32, device=dev)
        Ig = torch.empty((world, a.batch, a.k), dtype=torch.int64, device=dev)
        if a.exchange == "p2p":
            # peer mailboxes need CUDA IPC + peer access between all GPUs of the box; if any rank cannot map them,
            # every rank takes the NCCL all-gather path (still all on the GPUs) and the config line says so
            try:
                exch = sharded.PeerExchange(pkg, local, None, a.batch, a.k)  # raises on every rank or on none
            except RuntimeError as e:
                print(f"rank {rank}: {e}; using NCCL all-gather", file=sys.stderr)
                exch, a.exchange = None, "nccl"
    ix.reserve_search(a.batch, a.nprobe, a.k)
    pipelined = world == 1 or exch is not None

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def nccl_step(q, D, I):
        # the portable form: local search, two all-gathers, merge kernel -- all on the current stream
        ix.search_async(q, a.nprobe, a.k, D, I, stream)
        dist.all_gather_into_tensor(Dg, D)
        dist.all_gather_into_tensor(Ig, I)
        return pkg.merge_topk(Dg, Ig, stream)

    def run_device(first, count):
        """`count` batches, device-resident queries and results, as many in flight as the index pipelines;
        returns when the last one is ordered before the current stream"""
        if not pipelined:
            for s in range(first, first + count):
                nccl_step(q_all[s], Dd[0], Id[0])
            return
        t = 0
        for s in range(first, first + count):
            t = ix.search_submit(q_all[s], a.nprobe, a.k, Dd[s % depth], Id[s % depth])
        ix.search_wait_stream(t, stream)  # the back stream is in order: the last ticket covers them all

    # ---- untimed checks: (1) N > 1: the sharded answer equals the answer of an UNSHARDED index of the same rows
    # built on rank 0; (2) the peer-memory exchange reproduces the all-gather + merge-kernel result bit for bit
    sync_all()
    parity = None
    if world > 1:
        if exch is not None:
            ix.attach_exchange(exch._h)
        run_device(0, 1) if pipelined else None
        Ds, Is = (Dd[0].clone(), Id[0].clone()) if pipelined else nccl_step(q_all[0], Dd[0], Id[0])
        torch.cuda.synchronize()
        if exch is not None:
            exch.check()
            ix.attach_exchange(None)
            Dn, In = nccl_step(q_all[0], Dd[1 % depth], Id[1 % depth])
            torch.cuda.synchronize()
            ix.attach_exchange(exch._h)
            if not (torch.equal(Ds, Dn) and torch.equal(Is, In)):
                raise SystemExit(f"rank {rank}: peer-memory exchange and NCCL all-gather merge disagree")
        if use_truth:
            # (all ranks search; rank 0 judges) the check queries at the run's nprobe and exhaustively, against the
            # torch truth: top-k by (distance, id) over the rows of the probed lists / over all rows
            from parity import check_search
            Dc = torch.empty((nchk, a.k), dtype=torch.float32, device=dev)
            Ic = torch.empty((nchk, a.k), dtype=torch.int64, device=dev)
            verdict = torch.zeros(1, dtype=torch.int32, device=dev)
            checked = []
            for npr in (a.nprobe, a.nlist):
                ix.search_wait(ix.search_submit(qchk, npr, a.k, Dc, Ic))
                if rank == 0:
                    probes = torch.from_numpy(ix.select_nprobe(qchk, npr).astype(np.int64)).to(dev)
                    Dt = np.empty((nchk, a.k), np.float32)
                    It = np.empty((nchk, a.k), np.uint64)
                    for qi in range(nchk):
                        rows = torch.isin(truth["asg"], probes[qi].to(torch.int32)).nonzero().squeeze(1) if npr < a.nlist \
                            else torch.arange(a.n, device=dev)
                        d = truth["dist"][qi][rows]
                        order = torch.argsort(d, stable=True)[:a.k]
                        Dt[qi] = d[order].cpu().numpy()
                        It[qi] = rows[order].cpu().numpy().astype(np.uint64)
                    scale = (qchk.norm(dim=1) * truth["vmax"]).cpu().numpy() if a.metric == "ip" else None
                    try:
                        ties = check_search(Dc.cpu().numpy(), Ic.cpu().numpy().view(np.uint64), Dt, It, scale)
                        checked.append(f"nprobe={npr}: {nchk} queries equal the torch truth ({ties} tie-explained swaps)")
                    except AssertionError as e:
                        print(f"rank 0: sharded search at nprobe={npr} differs from the torch truth: {e}", file=sys.stderr)
                        verdict += 1
            dist.broadcast(verdict, 0)
            if int(verdict.item()):
                raise SystemExit("sharded search differs from the independent truth")
            if rank == 0:
                parity = {"sharded_equals_truth": True, "checked_on": "; ".join(checked),
                          "truth": f"torch distances of {nchk} check queries to all {a.n} rows + every row's list, kept on rank 0"}
                truth = None
        if not use_truth and not a.no_parity_check:
            verdict = torch.zeros(1, dtype=torch.int32, device=dev)
            if rank == 0 and ref is not None:
                Dr = torch.empty_like(Ds)
                Ir = torch.empty_like(Is)
                ref.search_async(q_all[0], a.nprobe, a.k, Dr, Ir, stream)
                torch.cuda.synchronize()
                same = bool(torch.equal(Is, Ir)) and bool(torch.equal(Ds, Dr))
                parity = {"sharded_equals_unsharded": same, "queries": a.batch,
                          "checked_on": f"batch 0 against an unsharded {a.n}-row index on rank 0"}
                if not same:
                    bad = int(((Is != Ir) | (Ds != Dr)).any(dim=1).sum())
                    print(f"rank 0: sharded search differs from the unsharded index on {bad} of {a.batch} queries",
                          file=sys.stderr)
                    verdict += 1
            dist.broadcast(verdict, 0)  # every rank leaves together: a parity failure must fail the run, not hang it
            if int(verdict.item()):
                raise SystemExit("sharded search differs from the unsharded index")
        if rank == 0 and ref is not None:
            ref.close()
        sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    run_device(0, a.warmup)
    sync_all()
    ix.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    run_device(a.warmup, a.steps)
    e1.record()
    sync_all()
    t1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    prof = ix.read_profile()
    ix.set_profiling(False)
    if exch is not None:
        exch.check()
    qps = a.batch * a.steps / (ms / 1e3)

    # ---- e2e: HOST buffers through the public call a serving thread makes (search_submit / search_wait on pinned
    # memory: H2D of the queries and D2H of ids + distances inside the timed region, `depth` batches in flight)
    q_host = q_all.cpu().pin_memory()
    Dh = [torch.empty((a.batch, a.k), dtype=torch.float32).pin_memory() for _ in range(depth)]
    Ih = [torch.empty((a.batch, a.k), dtype=torch.int64).pin_memory() for _ in range(depth)]

    def run_e2e(first, count):
        if not pipelined:
            for s in range(first, first + count):
                qd = q_host[s].to(dev, non_blocking=True)
                Do, Io = nccl_step(qd, Dd[0], Id[0])
                Dh[0].copy_(Do, non_blocking=True)
                Ih[0].copy_(Io, non_blocking=True)
                torch.cuda.synchronize()
            return
        tickets = []
        for s in range(first, first + count):
            if len(tickets) >= depth:
                ix.search_wait(tickets[len(tickets) - depth])  # the buffers of that batch are about to be reused
            tickets.append(ix.search_submit(q_host[s], a.nprobe, a.k, Dh[s % depth], Ih[s % depth]))
        for t in tickets[-depth:]:
            ix.search_wait(t)

    run_e2e(0, min(a.warmup, 5))
    sync_all()
    te = time.perf_counter()
    run_e2e(a.warmup, a.steps)
    sync_all()
    te = torch.tensor([time.perf_counter() - te], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_qps = a.batch * a.steps / float(te.item())
    clocks = sampler.stop(t0, time.time()) if sampler else None

    # ---- byte accounting of the timed batches (untimed replay; stats come from the grouping kernel)
    if exch is not None:
        ix.attach_exchange(None)  # the replay is local: no collective needed for the byte counts
    alg_rows = uniq_rows = items = rescored = 0
    scan_ctas = 0
    for s in range(a.steps):
        ix.search_async(q_all[a.warmup + s], a.nprobe, a.k, Dd[0], Id[0], stream)
        ss = ix.last_search_stats()
        alg_rows += ss.algorithmic_rows
        uniq_rows += ss.unique_rows
        items += ss.scan_items
        rescored += ss.rescored_pairs
    t = ix.search_submit(q_all[0], a.nprobe, a.k, Dd[0], Id[0])
    ix.search_wait(t)
    ss = ix.last_search_stats()
    scan_ctas = int(ss.scan_ctas)
    bpr = 4 * a.dim + 8
    bpr_streamed = int(ss.streamed_bytes_per_row) or bpr  # 2*ld + 8 when the bf16 tensor-core screen ran
    mirror = bpr_streamed != bpr
    peak, peak_src = measured_peak()
    nsearch = max(prof["searches"], 1)
    # kernel time: the scan streams' busy time per launch (first scan start .. last scan end over the timed
    # region).  With batches in flight consecutive scan launches overlap (batch i+1's CTAs start on the SMs batch
    # i's tail frees), so the per-launch event brackets (`kernel_ms_bracketed`) count the overlap twice.
    scan_ms = (prof["scan_span_ms"] / nsearch) if (pipelined and prof["scan_span_ms"] > 0) else prof["scan_ms"] / nsearch
    alg_bytes = alg_rows * bpr / a.steps
    uniq_bytes = uniq_rows * bpr / a.steps
    streamed_bytes = uniq_rows * bpr_streamed / a.steps
    achieved = streamed_bytes / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0

    rank_scan_ms = None
    if world > 1:  # per-rank scan time and distinct bytes: shows how well the list ownership balances
        mine = torch.tensor([scan_ms, streamed_bytes / 1e9], device=dev, dtype=torch.float64)
        allr = torch.empty((world, 2), device=dev, dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine)
        rank_scan_ms = [[round(float(a), 3), round(float(b_), 2)] for a, b_ in allr.cpu().tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        r = reference_sample(a, steps=3, warmup=1, max_seconds=20.0)
        cpu = {"value": r["value"], "unit": "queries/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    headline = (a.n, a.dim, a.metric, a.nlist, a.nprobe, a.k, a.batch) == (10_000_000, 768, "l2", 4096, 32, 10, 64)
    launches_per_step = 5 + (1 if mirror else 0) + (1 if world > 1 else 0)  # + query_image_kernel with a shadow
    line = {
        "metric": metric_name(a), "value": qps, "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": ("f32" if not mirror else
                  f"f32 ({'int8' if bpr_streamed == a.dim + 12 else 'bf16'} tensor-core screen, fp32 exact re-score of the admitted pairs)"),
        "data": "synthetic",
        "config": {"workload": workload_name(a) + (f" [shard {a.emulate_rank} of {a.emulate_shards} only: tuning run]" if shard_count != world else ""),
                   "parallelism": f"lists sharded over {world} GPU(s), byte-balanced ownership, " +
                   ("single GPU: no exchange" if world == 1 else
                    "merge kernel stores into the peers' NVLink mailboxes + collect kernel, pipelined over batches"
                    if exch is not None else "NCCL all-gather + merge kernel per batch"),
                   "pipeline": (f"{depth} batches in flight: coarse/grouping of batch i+1 and merge/exchange of batch i-1 "
                                f"overlap the scan of batch i ({scan_ctas} scan CTAs of 148 SMs)") if pipelined else "none",
                   "cache": f"inputs larger than L2: each batch streams {streamed_bytes / 1e9:.2f} GB of distinct list data" +
                            (f" (bf16 shadow of {uniq_bytes / 1e9:.2f} GB of fp32 rows)" if mirror else ""),
                   "scan": (f"tensor-core screen over the lists' {'int8' if bpr_streamed == a.dim + 12 else 'bf16'} shadow "
                            "(tcgen05), exact fp32 re-score of the admitted pairs: results bit-identical to the fp32 scan")
                           if mirror else "fp32 list scan",
                   "ntrain": min(a.ntrain, a.n), "page_rows": st.page_rows,
                   "build_s": round(t_build, 1), "train_s": round(t_train, 1), "add_s": round(t_add, 1),
                   "index_gb": round(st.gpu_memory_bytes / 1e9, 2),
                   **({"rank_scan_ms_and_unique_gb": rank_scan_ms} if rank_scan_ms else {}),
                   **({"parity": parity} if parity else {})},
        # frac = bytes the scan streams (distinct probed rows x bytes per row as stored for the scan) / scan time / peak:
        # a list probed by several queries of the batch streams from HBM once, and the bf16 screen streams the rows'
        # bf16 shadow (2*ld + 8 bytes per row) instead of the fp32 rows.  `unique_bytes_per_launch` is the fp32 size of
        # the same rows (4*dim + 8 each: what SURVEY 8d counts), `fp32_equivalent_gbs` that size over the scan time;
        # `reuse` = algorithmic bytes (every query's probed rows) / distinct bytes.
        "roofline": {"bound": "hbm", "kernel": "screen_kernel" if mirror else "scan_kernel", "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": profiled_traffic(mirror) if (headline and world == 1 and shard_count == 1) else None,
                     "peak_source": peak_src,
                     "streamed_bytes_per_launch": streamed_bytes,
                     # pairs the screen admitted: each re-reads one fp32 row (4*ld bytes) for the exact distance
                     "rescored_pairs_per_launch": rescored / a.steps,
                     "rescore_bytes_per_launch": rescored / a.steps * 4 * a.dim,
                     "fp32_equivalent_gbs": uniq_bytes / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0,
                     "algorithmic_bytes_per_launch": alg_bytes, "unique_bytes_per_launch": uniq_bytes,
                     "reuse": alg_bytes / uniq_bytes if uniq_bytes else None,
                     "algorithmic_gbs": alg_bytes / (scan_ms / 1e3) / 1e9 if scan_ms > 0 else 0.0,
                     "kernel_ms": scan_ms, "kernel_ms_bracketed": prof["scan_ms"] / nsearch,
                     "scan_items_per_launch": items / a.steps, "scan_ctas": scan_ctas,
                     "step_breakdown_ms": {k_: prof[k_] / nsearch
                                           for k_ in ("coarse_ms", "group_ms", "scan_ms", "merge_ms", "collect_ms")}},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": a.batch * a.dim * 4,
                "d2h_bytes_per_step": a.batch * a.k * 12},
        # per step: score_gemm (tcgen05) + coarse_select + (query_image) + build_groups + scan / screen + merge (+ collect,
        # or merge of the all-gathered parts)
        "gpu_launches": a.steps * launches_per_step,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------- configs[1]: brute force

def tensor_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"]) / 2.0, "measured bf16 burst / 2 (TF32 operands, MEASURED_PEAKS.json)"
    except Exception:
        return 1100.0, "fallback: 2250 / 2 nominal"


def run_c2(a):
    """BASELINE configs[1]: exact top-100 of 1024 queries against a flat 1M x 768 database (vdb_bruteforce_search:
    TF32 tcgen05 contraction over nested samples + exact fp32 re-scoring).  A step = one batch of 1024 fresh queries."""
    import numpy as np
    import torch
    import oracle_lib as O
    pkg = importlib.import_module(PKG)
    pkg.lib()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n, dim, nq, k = 1_000_000, 768, 1024, 100
    gen = torch.Generator(device=dev).manual_seed(12345)
    db = torch.randn(n, dim, generator=gen, device=dev)
    nb = a.warmup + a.steps
    q_all = torch.randn(nb, nq, dim, generator=gen, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    sampler = ClockSampler(0)
    for s in range(a.warmup):
        pkg.bruteforce_search(db, q_all[s], k, stream=stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for s in range(a.steps):
        D, I = pkg.bruteforce_search(db, q_all[a.warmup + s], k, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    q_host = q_all.cpu().pin_memory()
    Dh = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    Ih = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    lib = pkg.lib()

    def e2e_step(s):
        pkg._check(lib.vdb_bruteforce_search(db.data_ptr(), q_host[s].data_ptr(), None, n, nq, dim, k, Dh.data_ptr(),
                                             Ih.data_ptr(), 0, stream))
    for s in range(min(3, a.warmup)):
        e2e_step(s)
    te = time.perf_counter()
    for s in range(a.steps):
        e2e_step(a.warmup + s)
    te = time.perf_counter() - te
    clocks = sampler.stop(t0, time.time())
    # untimed check of the last batch's first queries against the oracle's flat search (all host cores)
    ncores = os.cpu_count() or 1
    cpu = None
    if not a.no_cpu_baseline:
        ora = O.OracleIndex(dim, 1)
        ora.centroids = np.zeros((1, dim), np.float32)
        ora.load_assigned(db.cpu().numpy(), np.arange(n, dtype=np.uint64), np.zeros(n, np.uint32))
        m = min(nq, 2 * ncores)
        qs = q_all[nb - 1, :m].cpu().numpy()
        t = time.perf_counter()
        Dr, Ir = ora.search(qs, 1, k, ncores)
        dt = time.perf_counter() - t
        from parity import check_search
        check_search(D[:m].cpu().numpy(), I[:m].cpu().numpy().view(np.uint64), Dr, Ir)
        cpu = {"value": m / dt, "unit": "queries/s", "cores": ncores, "kind": "port",
               "sample": f"{m} of the 1024 queries of one batch against the full 1M x 768 database (oracle search_list_cpu "
                         f"over one list), results equal to the GPU's"}
    flops = 2.0 * nq * n * dim
    peak, src = tensor_peak()
    ach = flops / (ms / a.steps / 1e3) / 1e12
    print(json.dumps({
        "metric": "brute-force flat L2 QPS @1M x 768D, 1024 queries, k=100", "value": nq * a.steps / (ms / 1e3),
        "unit": "queries/s", "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (tf32 tensor-core screen, fp32 exact re-score)",
        "data": "synthetic",
        "config": {"workload": "brute-force flat L2 1M x 768D, query batch 1024, k=100 (BASELINE.json configs[1])",
                   "cache": "the 3.07 GB database exceeds L2; fresh queries every step"},
        "roofline": {"bound": "tensor", "kernel": "vdb_bruteforce_search (whole call: row norms, 4 sampled GEMM levels, "
                     "full GEMM, threshold and select kernels; the full-level rowtile_gemm_kernel is ~half of it)",
                     "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                     "peak_source": src, "algorithmic_flops_per_launch": flops},
        "cpu_baseline": cpu,
        "e2e": {"value": nq * a.steps / te, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4,
                "d2h_bytes_per_step": nq * k * 12},
        "gpu_launches": a.steps * 14, "clocks": clocks}), flush=True)


# ------------------------------------------- configs[4]: k-means train + add

def run_c5(a):
    """BASELINE configs[4]: k-means training on the first 1M rows + add() of 10M x 768 rows at nlist 16384, one
    process over --devices (vdb_index_create_sharded).  A step = one add() batch of 1M rows; training is timed
    once and reported in config."""
    import hashlib
    import numpy as np
    import torch
    pkg = importlib.import_module(PKG)
    pkg.lib()
    devs = tuple(int(d) for d in a.devices.split(",")) if a.devices else (0,)
    dev = torch.device("cuda", devs[0])
    torch.cuda.set_device(devs[0])
    dim, nlist, ntrain, rows, bs = 768, 16384, 1_000_000, 10_000_000, 1_000_000
    ix = pkg.IVFFlatIndex(pkg.Config(dimension=dim, nlist=nlist, devices=devs if len(devs) > 1 else (), device=devs[0]))
    gen = torch.Generator(device=dev).manual_seed(12345)
    xt = torch.randn(ntrain, dim, generator=gen, device=dev)
    torch.cuda.synchronize()
    sampler = ClockSampler(devs[0])
    t0 = time.time()
    t = time.perf_counter()
    ix.train(xt)
    train_s = time.perf_counter() - t
    del xt
    sha = hashlib.sha256(np.ascontiguousarray(ix.centroids).tobytes()).hexdigest()
    nb = rows // bs
    t_add = 0.0
    for b in range(nb):
        x = torch.randn(bs, dim, generator=gen, device=dev)
        ids = torch.arange(b * bs, (b + 1) * bs, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        t = time.perf_counter()
        ix.add(x, ids)
        t_add += time.perf_counter() - t
    # e2e: the same add() from HOST rows (pinned), two more batches
    xh = torch.randn(bs, dim, generator=gen, device=dev).cpu().pin_memory()
    idh = torch.arange(rows, rows + bs, dtype=torch.int64).pin_memory()
    t = time.perf_counter()
    for b in range(2):
        ix.add(xh, idh + b * bs)
    e2e_s = (time.perf_counter() - t) / 2
    clocks = sampler.stop(t0, time.time())
    flops = 2.0 * bs * nlist * dim
    peak, src = tensor_peak()
    ach = flops / (t_add / nb) / 1e12 / len(devs)  # every shard assigns the whole batch: per-GPU rate
    print(json.dumps({
        "metric": "add() rows/s @10M x 768D nlist=16384 (after k-means training on 1M rows)", "value": rows / t_add,
        "unit": "rows/s", "n_gpus": len(devs), "steps": nb, "warmup": 0, "ms_per_step": t_add / nb * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (tf32 tensor-core screen, fp32 exact re-check)",
        "data": "synthetic",
        "config": {"workload": "k-means IVF training + batched add() 10M x 768D nlist=16384 (BASELINE.json configs[4])",
                   "parallelism": f"one process, {len(devs)} shard(s): data-parallel bit-exact training; every shard assigns "
                                  f"each batch and keeps the lists it owns",
                   "train_s": round(train_s, 3), "ntrain": ntrain, "centroid_sha256": sha,
                   "rows_held_all_shards": int(ix.list_sizes().sum()), "index_gb": round(ix.stats().gpu_memory_bytes / 1e9, 2)},
        "roofline": {"bound": "tensor", "kernel": "rowtile_gemm_kernel<AssignEpi> inside add() (assignment + staging + scatter timed together)",
                     "achieved": ach * len(devs) if len(devs) == 1 else ach, "peak": peak, "unit": "TFLOP/s",
                     "frac": ach / peak, "traffic": None, "peak_source": src, "algorithmic_flops_per_launch": flops},
        "cpu_baseline": None,
        "e2e": {"value": bs / e2e_s, "unit": "rows/s", "h2d_bytes_per_step": bs * (dim * 4 + 8) * len(devs), "d2h_bytes_per_step": 0},
        "gpu_launches": nb * 8 * len(devs), "clocks": clocks}), flush=True)


def _divert_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on stdout.
    fd 1 is pointed at stderr for the whole run and print() is given the real stdout back."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    args = parse()
    _divert_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c2":
        run_c2(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_b200(args)
