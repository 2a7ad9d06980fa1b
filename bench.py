#!/usr/bin/env python
"""
bench.py -- QPS at recall@10 >= 0.95 on 1M x 128 sq_euclid (BASELINE.json metric, config C2).

    python bench.py --gpus N --steps K --warmup W            # the B200 engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A "step" is one pass of the hot path (BatchKnnQuery) over one 10k-query batch.
  value       whole-job QPS, queries and outputs resident in HBM, CUDA events on the launching stream
  e2e         the same through the reference-facing C-ABI call hnsw_knn_query with pinned HOST buffers
              (H2D of the queries and D2H of ids+dists inside the timed region)
  roofline    HBM gather roofline of the search kernel: algorithmic bytes (oracle-counted distance
              evaluations x dim x 4 + adjacency words x 4) / CUDA-event kernel time / measured HBM peak
  cpu_baseline the oracle (C++ restatement of the reference, Parallel.For stand-in on all host cores) on the
              same graph and queries, bounded sample
Multi-GPU: one process per GPU (torchrun), graph replicated, each rank searches its own 10k-query batches,
no data-path collective (SURVEY.md 8e mode 1) -> weak scaling; time = max over ranks.
"""
import argparse
import ctypes as ct
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))

EF_GRID = [10, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096]
EF_GRID_WIDE = EF_GRID + [6144, 8192, 12288, 16384, 20480, 24576]   # the reference's heaps have no ceiling (BinaryHeap.cs:109-113)
DIM, RANK, K = 128, 16, 10
M, EFC = 16, 200
# the BASELINE configs that fit one GPU.  c2 is the one the metric is quoted on (the default).
WORKLOADS = {
    "c2": dict(tag="C2", dim=128, rank=16, metric="sq_euclid", M=16, efc=200, k=10, dist_rate=None,
               metric_name="QPS at recall@10>=0.95, 1M x 128 sq_euclid", shape="SIFT-shaped (rank-16 + 0.25 noise)"),
    "c2u": dict(tag="C2-uniform", dim=128, rank=0, metric="sq_euclid", M=16, efc=200, k=10, dist_rate=None, extra_nq=500, extra_steps=3,
                metric_name="QPS at recall@10>=0.95, 1M x 128 sq_euclid, U[0,1) data (hard case of SURVEY 8d)",
                shape="uniform U[0,1)^128 (the reference tests' distribution; intrinsic dimension 128)"),
    "c3": dict(tag="C3", dim=960, rank=32, metric="cosine", M=32, efc=200, k=100, dist_rate=1.0 / math.log(32.0), extra_steps=5,
               metric_name="QPS at recall@100>=0.95, 1M x 960 cosine (C3, GIST-shaped)", shape="GIST-shaped (rank-32 + 0.25 noise, not normalised)"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_data(n, nq_total, seed_data=1234, seed_q=4321, dim=DIM, rank=RANK):
    """x = zA + 0.25 eps, z~N(0,I_rank) (low intrinsic dimension); queries from the same
    distribution, held out."""
    rng = np.random.default_rng(seed_data)
    if rank <= 0:  # uniform hard case
        data = rng.random((n, dim), dtype=np.float32)
        return data, np.random.default_rng(seed_q).random((nq_total, dim), dtype=np.float32)
    A = rng.standard_normal((rank, dim)).astype(np.float32)
    data = np.empty((n, dim), dtype=np.float32)
    for s in range(0, n, 131072):
        e = min(n, s + 131072)
        z = rng.standard_normal((e - s, rank)).astype(np.float32)
        data[s:e] = z @ A + 0.25 * rng.standard_normal((e - s, dim)).astype(np.float32)
    rq = np.random.default_rng(seed_q)
    q = (rq.standard_normal((nq_total, rank)).astype(np.float32) @ A + 0.25 * rq.standard_normal((nq_total, dim)).astype(np.float32))
    return data, np.ascontiguousarray(q.astype(np.float32))


def recall_at_k(found, truth):
    return float(np.mean([len(set(found[i]) & set(truth[i])) / truth.shape[1] for i in range(len(truth))]))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


_REF_BINDINGS = None


def reference_bindings():
    """The reference's OWN bindings.py (installed byte for byte into baseline/_ref by __graft_entry__.build()), staged in a temp
    dir with HNSWIndex.Native.so where its loader looks (bindings.py:15-41) and imported under another module name, so that it
    does not collide with this repo's mirror package.  None when baseline/_ref is absent."""
    global _REF_BINDINGS
    if _REF_BINDINGS is not None:
        return _REF_BINDINGS or None
    import importlib.util
    import shutil
    src = ROOT / "baseline" / "_ref" / "hnswindex" / "bindings.py"
    if not src.exists():
        _REF_BINDINGS = False
        return None
    import hnswindex
    td = Path(tempfile.mkdtemp(prefix="hnsw_ref_pkg_"))
    native = td / "artifacts" / "native" / "linux-x64"
    native.mkdir(parents=True)
    shutil.copyfile(src, td / "bindings.py")
    os.symlink(hnswindex.LIB_PATH, native / "HNSWIndex.Native.so")
    spec = importlib.util.spec_from_file_location("hnswindex_reference_bindings", td / "bindings.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _REF_BINDINGS = mod
    return mod


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def pick_ef(search, truth, grid=EF_GRID, target=0.95, k=K):
    grid = [k] * (k not in grid) + [e for e in grid if e >= k]  # ef = max(MinNN, k): values below k are all k
    grid.sort()
    for ef in grid:
        r = recall_at_k(search(ef), truth)
        log(f"  ef={ef}: recall@{k}={r:.4f}")
        if r >= target:
            return ef, r
    return grid[-1], r


# ------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from hnswindex import bindings as hb
    hb.set_device(local)

    line = replicated_pass(args, args.workload, args.steps, args.warmup, headline=True)

    # ---- N > 1: the partitioned mode (BASELINE config C4) on the same ranks, as a sub-record of this line
    if world > 1 and not args.no_partitioned:
        try:
            part = partitioned_pass(args, steps=min(args.steps, 20), warmup=3)
        except Exception as e:   # the headline line must still be printed; the failure is reported in its place
            part = {"error": f"{type(e).__name__}: {e}"[:500]}
            log(f"[bench] rank {rank}: partitioned pass failed: {e}")
        if rank == 0:
            line["partitioned"] = part

    # ---- N = 1: the other named shapes of BASELINE.json as sub-records (each measured like the headline, fewer steps)
    if world == 1 and args.workload == "c2" and args.extras:
        others = {}
        t_ex = time.time()
        for name in [x for x in args.extras.split(",") if x]:
            if time.time() - t_ex > args.extras_budget_s:
                others[name] = {"skipped": f"extras budget of {args.extras_budget_s:.0f} s used up"}
                continue
            try:
                ts = time.time()
                if name == "c5":
                    others["c5_build"] = build_pass(args, "c2", steps=2)
                elif name in WORKLOADS and name != "c2":
                    sub = replicated_pass(args, name, min(args.steps, WORKLOADS[name].get("extra_steps", 10)), 3, headline=False)
                    others[name] = compact_record(sub)
                log(f"[bench] extra workload {name}: {time.time() - ts:.1f}s")
            except Exception as e:
                others[name] = {"error": f"{type(e).__name__}: {e}"[:500]}
                log(f"[bench] extra workload {name} failed: {e}")
        line["other_workloads"] = others

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def compact_record(line):
    """Sub-record of an extra workload: the numbers of its own JSON line without the boilerplate."""
    cfg = line["config"]
    keep = {k: v for k, v in cfg.items() if k.startswith("recall_at_") or k in ("ef", "workload", "index_nodes", "inserts_per_s",
                                                                                "build_kernel_inserts_per_s", "build_gathered_GBps")}
    return {"metric": line["metric"], "value": line["value"], "unit": line["unit"], "ms_per_step": line["ms_per_step"], "steps": line["steps"],
            "e2e": line["e2e"], "roofline": line["roofline"], "cpu_baseline": line["cpu_baseline"], "clocks": line["clocks"],
            "gpu_launches": line["gpu_launches"], **keep}


def replicated_pass(args, workload, steps, warmup, headline):
    """One workload of SURVEY 8(d) on a replicated graph (mode 1): build (untimed), pick ef, time `steps` query batches three
    ways (queued device-resident = value, one by one = roofline, host buffers through the C ABI = e2e), oracle on the same
    graph for the algorithmic bytes and the CPU baseline.  Returns the JSON line as a dict on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    import hnswindex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    class _A:   # the pass reads steps / warmup / workload from here
        pass
    a = _A()
    a.__dict__.update(vars(args))
    a.steps, a.warmup, a.workload = steps, warmup, workload
    args = a

    W = WORKLOADS[args.workload]
    DIM, RANK, K, M, EFC, METRIC = W["dim"], W["rank"], W["k"], W["M"], W["efc"], W["metric"]
    n, nq, nb = args.n, args.nq, args.query_batches
    if not headline:
        nq = min(nq, W.get("extra_nq", nq))
    data, queries = make_data(n, nq * nb, seed_q=4321 + 1000 * rank, dim=DIM, rank=RANK)
    snap = Path(args.snapshot_dir) / f"bench_{args.workload}_{n}.hnsw"

    def new_index():
        ix = hnswindex.Index(DIM, METRIC)
        ix.set_collection_size(n); ix.set_max_edges(M); ix.set_max_candidates(EFC); ix.set_allow_removals(False)
        if W["dist_rate"]:
            ix.set_distribution_rate(W["dist_rate"])
        return ix

    # ---- setup (untimed): build once on rank 0, replicate through the serialized snapshot
    t0 = time.time()
    build_info = {}
    if rank == 0:
        ix = new_index()
        tb = time.time()
        ix.add(data)
        build_s = time.time() - tb
        bp = ix.build_profile()
        kms_build = ix.stats()["last_build_ms"]
        nd = bp["ndist_search"] + bp["ndist_select"] + bp["ndist_prune"]
        build_info = {"inserts_per_s": round(n / build_s), "build_s": round(build_s, 3), "build_kernel_ms": round(kms_build, 1),
                      "build_kernel_inserts_per_s": round(n / (kms_build / 1e3)),
                      "build_ndist_per_insert": round(nd / max(1, bp["inserts"]), 1),
                      "build_gathered_GBps": round(nd * DIM * 4 / (kms_build / 1e3) / 1e9, 1)}
        log(f"[bench] GPU build: {n} inserts in {build_s:.2f}s ({n / build_s:.0f}/s), kernels {kms_build:.0f} ms")
        if world > 1 or not args.no_cpu_baseline:
            ts = time.time(); ix.serialize(snap, packed=True); log(f"[bench] serialize {time.time() - ts:.1f}s")
    if world > 1:
        dist.barrier()
        if rank != 0:
            ix = hnswindex.Index.deserialize(METRIC, snap)
    del data

    # ---- ef = smallest grid value with recall@10 >= 0.95 (GPU search == oracle search on the same graph: tests/)
    ns = min(2000, nq)
    truth, _ = ix.exact_knn(queries[:ns], K)

    def search_ef(ef):
        ix.set_min_nn(ef)
        return ix.knn_query(queries[:ns], K)[0]
    if rank == 0:
        ef, rec = pick_ef(search_ef, truth, grid=EF_GRID_WIDE if RANK <= 0 else EF_GRID, k=K)
        sel = torch.tensor([ef], dtype=torch.int32, device="cuda")
    else:
        sel = torch.zeros(1, dtype=torch.int32, device="cuda")
    if world > 1:
        dist.broadcast(sel, 0)
    ef = int(sel.item())
    ix.set_min_nn(ef)
    recall = recall_at_k(ix.knn_query(queries[:ns], K)[0], truth)
    log(f"[bench] rank {rank}: ef={ef} recall@{K}={recall:.4f} setup {time.time() - t0:.1f}s")

    # ---- device-resident buffers, engine runs on torch's current stream so torch events bracket it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ix.set_stream(stream.cuda_stream)
    q_dev = [torch.from_numpy(queries[b * nq:(b + 1) * nq]).cuda() for b in range(nb)]
    ids_dev = torch.empty((nq, K), dtype=torch.int32, device="cuda")
    d_dev = torch.empty((nq, K), dtype=torch.float32, device="cuda")
    ids_dev2, d_dev2 = torch.empty_like(ids_dev), torch.empty_like(d_dev)   # consecutive queued batches overlap: own outputs
    # pinned host buffers for the end-to-end (C-ABI, host pointers) path
    q_pin = [torch.from_numpy(queries[b * nq:(b + 1) * nq]).pin_memory() for b in range(nb)]
    ids_pin = torch.empty((nq, K), dtype=torch.int32).pin_memory()
    d_pin = torch.empty((nq, K), dtype=torch.float32).pin_memory()
    f32p, i32p = ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)

    def step_resident(i):   # one batch, host waits for it (per-launch CUDA-event time -> roofline)
        ix.knn_query_device(q_dev[i % nb].data_ptr(), nq, K, ids_dev.data_ptr(), d_dev.data_ptr())

    def step_queued(i):     # one batch, only queued: the serving loop (host jitter does not idle the GPU, tails overlap)
        oi, od = (ids_dev, d_dev) if i % 2 == 0 else (ids_dev2, d_dev2)
        ix.knn_query_device_async(q_dev[i % nb].data_ptr(), nq, K, oi.data_ptr(), od.data_ptr())

    def step_e2e(i):
        rc = hnswindex.lib.hnsw_knn_query(ix._h, ct.cast(q_pin[i % nb].data_ptr(), f32p), nq, DIM, K, ct.cast(ids_pin.data_ptr(), i32p),
                                          ct.cast(d_pin.data_ptr(), f32p))
        assert rc == 0, hnswindex.last_error()

    def timed(step_fn, steps, warmup, collect_kernel_ms=False):
        for i in range(warmup):
            step_fn(i)
        ix.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kms = []
        t_wall = time.perf_counter()
        e0.record(stream)
        for i in range(steps):
            step_fn(warmup + i)
            if collect_kernel_ms:
                kms.append(ix.stats()["last_search_ms"])
        ix.async_join()   # queued batches on the engine's auxiliary stream end before e1 (stream-ordered, no host wait)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t_wall
        if world > 1:
            dist.barrier()
        dev_ms = e0.elapsed_time(e1)
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item()), kms

    sampler = ClockSampler(local)
    launches0 = ix.stats()["search_launches"]
    if rank == 0:
        sampler.start()
    dev_ms, wall_ms, _ = timed(step_queued, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    launches = ix.stats()["search_launches"] - launches0 - args.warmup
    _, _, kms = timed(step_resident, args.steps, args.warmup, collect_kernel_ms=True)   # the same launches, timed one by one
    e2e_dev_ms, e2e_wall_ms, _ = timed(step_e2e, args.steps, args.warmup)

    # ---- the same through the reference's OWN Python package: pageable numpy in, numpy out (what a user of the reference
    #      switching libraries actually calls); the handle is the index built above
    refb = reference_bindings()
    q_np = [np.ascontiguousarray(queries[b * nq:(b + 1) * nq]) for b in range(nb)]
    if refb is not None:
        rix = refb.Index(DIM, metric=METRIC)
        rix._h, rix._initialized = ix._h, True
        pageable_api = "reference bindings.py Index.knn_query (unmodified package from baseline/_ref; pageable numpy in, numpy out)"
    else:
        rix = ix
        pageable_api = "hnswindex mirror Index.knn_query (baseline/_ref absent; pageable numpy in, numpy out)"
    page_out = {}

    def step_pageable(i):
        page_out["r"] = rix.knn_query(q_np[i % nb], K)
    _, page_wall_ms, _ = timed(step_pageable, args.steps, args.warmup)
    if refb is not None:
        rix._h = None   # the handle belongs to `ix`
    # ---- and from several host threads at once ("thread-safe per operation type", README.md:64-65 of the reference):
    #      hnsw_knn_query holds the index's lock shared and runs on its own stream + scratch, so the calls overlap on the GPU
    nthreads = 2
    tq = [torch.from_numpy(queries[(t % nb) * nq:((t % nb) + 1) * nq]).pin_memory() for t in range(nthreads)]
    ti = [torch.empty((nq, K), dtype=torch.int32).pin_memory() for _ in range(nthreads)]
    td = [torch.empty((nq, K), dtype=torch.float32).pin_memory() for _ in range(nthreads)]
    per_thread = max(10, args.steps)
    go = threading.Barrier(nthreads + 1)

    def thread_body(t):
        pq, pi, pd = ct.cast(tq[t].data_ptr(), f32p), ct.cast(ti[t].data_ptr(), i32p), ct.cast(td[t].data_ptr(), f32p)
        fn, h = hnswindex.lib.hnsw_knn_query, ix._h

        def call():   # (pointers prepared outside: the GIL is only held for the ctypes call set-up)
            rc = fn(h, pq, nq, DIM, K, pi, pd)
            assert rc == 0, hnswindex.last_error()
        for _ in range(2):
            call()
        go.wait()
        for _ in range(per_thread):
            call()
    ths = [threading.Thread(target=thread_body, args=(t,)) for t in range(nthreads)]
    for th in ths:
        th.start()
    go.wait()
    t_c = time.perf_counter()
    for th in ths:
        th.join()
    conc_ms = (time.perf_counter() - t_c) * 1e3
    ct_t = torch.tensor([conc_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ct_t, op=dist.ReduceOp.MAX)
    conc_ms = float(ct_t.item())
    pid, pdist = page_out["r"]
    assert pid.shape == (nq, K) and pid.dtype == np.int32 and pdist.dtype == np.float32

    if rank != 0:
        return None

    total_q = world * nq * args.steps
    value = total_q / (dev_ms / 1e3)
    e2e_value = total_q / (e2e_wall_ms / 1e3)   # the host call returns after the D2H: wall clock is the honest end-to-end time
    kernel_ms = float(np.mean(kms))

    # ---- oracle on the same graph: algorithmic bytes for the roofline + CPU baseline (bounded sample)
    roof = {"bound": "hbm", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None}
    cpu = None
    peak, peak_src = measured_peak()
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        ts = time.time()
        ref = O.OracleIndex.deserialize(METRIC, snap)
        ref.set_min_nn(ef)
        log(f"[bench] oracle loaded the GPU-built snapshot in {time.time() - ts:.1f}s")
        cores = os.cpu_count() or 1
        ncheck = nq if args.workload == "c2" else min(nq, 2000)   # wide rows: count the work on a sample, scale to the batch
        o_ids, o_d, cnt = ref.knn_query(queries[:ncheck], K, threads=cores, counters=True)
        ix.set_stream(0)
        g_ids, g_d = ix.knn_query(queries[:ncheck], K)
        same = float((g_ids == o_ids).all(axis=1).mean())
        cnt = [int(c) * nq // ncheck for c in cnt]
        bytes_vec = int(cnt[0]) * DIM * 4
        bytes_adj = int(cnt[2]) * 4
        alg = bytes_vec + bytes_adj            # per launch (one batch; batches are i.i.d.)
        ach = alg / (kernel_ms / 1e3) / 1e9
        roof.update({"achieved": round(ach, 1), "peak": peak, "frac": round(ach / peak, 4), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg, "gathered_vector_GBps": round(bytes_vec / (kernel_ms / 1e3) / 1e9, 1),
                     "ndist_per_query": round(int(cnt[0]) / nq, 1), "nexp_per_query": round(int(cnt[1]) / nq, 1),
                     "kernel": (f"knn_search_kernel<{METRIC}>" if ef < 1024 else f"knn_wide_kernel<{METRIC}> (one CTA per query)"),
                     "kernel_ms_avg": round(kernel_ms, 4),
                     # the queued batches of the timed region overlap in their tails: whole-region figure, for the record
                     "pipelined_algorithmic_GBps": round(alg * args.steps / (dev_ms / 1e3) / 1e9, 1)})
        prof = ROOT / "profiles" / {"c2": "search_kernel_traffic.json", "c3": "search_kernel_traffic_c3.json"}.get(args.workload, "none")
        if prof.exists() and nq == 10000:
            # ncu's DRAM bytes are only quoted while the kernel's machine code is the one that was captured
            tj = json.load(open(prof))
            have = ge._load_build_module().kernel_sass_stamp(tj.get("kernel_mangled", ""))
            if have is not None and have == tj.get("kernel_sass_sha256") and ef == (48 if args.workload == "c2" else 192):
                roof["traffic"] = tj.get("dram_bytes_per_launch")
                roof["traffic_source"] = f"{tj.get('source')}; kernel SASS sha256 {have[:16]} matches the library of this run"
            else:
                roof["traffic_stale"] = ("profiles/search_kernel_traffic.json was captured from different machine code (or at another "
                                         "search width) than this run's kernel: no traffic figure")
        # CPU baseline (N=1 only): repeat the 10k batch until ~10 s of work (max 5 passes)
        t_cpu, passes = 0.0, 0
        while world == 1 and t_cpu < (10.0 if headline else 3.0) and passes < 5:
            tc = time.perf_counter(); ref.knn_query(queries[:ncheck], K, threads=cores); t_cpu += time.perf_counter() - tc; passes += 1
        cpu = None if world > 1 else {"value": round(ncheck * passes / t_cpu, 1), "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"{passes} x {ncheck} queries of batch 0 on the same GPU-built {n}-node graph, ef={ef}, all host threads "
                         f"(C++ restatement of the reference's Parallel.For BatchKnnQuery, not the .NET runtime)",
               "id_lists_identical_to_gpu": round(same, 5)}
        del ref
    try:
        snap.unlink()
    except OSError:
        pass

    line = {
        "metric": W["metric_name"], "value": round(value, 1), "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{W['tag']}: {n} x {DIM} fp32 {METRIC} {W['shape']}, M={M}, efConstruction={EFC}, "
                               f"{nq}-query batch per GPU per step, k={K}, MinNN(ef)={ef} = smallest grid value with recall@{K}>=0.95",
                   f"recall_at_{K}": round(recall, 4), "ef": ef, "index_nodes": n, "queries_per_step_per_gpu": nq,
                   "parallelism": f"graph replicated x{world}, query batches sharded, no collective",
                   "timed_region": f"{args.steps} batches queued back to back (hnsw_knn_query_device_async: consecutive batches alternate two "
                                   "streams and overlap in their tails), CUDA events around them, one host wait at the end; "
                                   "roofline.kernel_ms_avg = the same launches timed one by one",
                   "l2_policy": f"inputs larger than L2: {n * (DIM * 4 + 2 * M * 4) / 1e6:.0f} MB graph vs 126 MB L2, "
                                f"{nb} distinct query batches rotated", "graph_built_by": "GPU engine (hnsw_add)", **build_info},
        "e2e": {"value": round(e2e_value, 1), "unit": "queries/s", "h2d_bytes_per_step": nq * DIM * 4, "d2h_bytes_per_step": nq * K * 8,
                "ms_per_step": round(e2e_wall_ms / args.steps, 4), "api": "hnsw_knn_query (C ABI, pinned host buffers)",
                "pageable": {"value": round(total_q / (page_wall_ms / 1e3), 1), "unit": "queries/s",
                             "ms_per_step": round(page_wall_ms / args.steps, 4), "api": pageable_api},
                "concurrent": {"value": round(world * nq * per_thread * nthreads / (conc_ms / 1e3), 1), "unit": "queries/s", "host_threads": nthreads,
                               "calls_per_thread": per_thread, "api": "hnsw_knn_query (C ABI, pinned host buffers) called from "
                               f"{nthreads} host threads at once on one handle"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "wall_ms_per_step": round(wall_ms / args.steps, 4),
    }
    del ix
    return line


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU path (oracle port: C++ restatement, all host threads).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    cores = os.cpu_count() or 1
    W = WORKLOADS[args.workload]
    DIM, RANK, K, M, EFC, METRIC = W["dim"], W["rank"], W["k"], W["M"], W["efc"], W["metric"]
    n, nq = args.n, args.nq
    data, queries = make_data(n, nq, dim=DIM, rank=RANK)
    t0 = time.time()
    okw = {"distribution_rate": W["dist_rate"]} if W["dist_rate"] else {}
    ref = O.OracleIndex(METRIC, max_edges=M, max_candidates=EFC, collection_size=n, allow_removals=False, **okw)
    # probe the insert rate, then decide who builds the (untimed) graph
    probe = min(n, 20000)
    ref.add(data[:probe], threads=cores)
    rate = probe / (time.time() - t0)
    built_by = "oracle (threaded Add on all host cores)"
    if n > probe and (n - probe) / max(rate * 0.6, 1.0) > args.reference_build_budget_s:
        # too slow on this host: the graph (setup, untimed) comes from the GPU engine's snapshot instead
        log(f"[bench-ref] oracle insert rate {rate:.0f}/s: building {n} on CPU would exceed the budget; using the GPU engine for SETUP only")
        del ref
        import __graft_entry__ as ge
        ge.build()
        import hnswindex
        ix = hnswindex.Index(DIM, METRIC)
        ix.set_collection_size(n); ix.set_max_edges(M); ix.set_max_candidates(EFC); ix.set_allow_removals(False)
        if W["dist_rate"]:
            ix.set_distribution_rate(W["dist_rate"])
        ix.add(data)
        snap = Path(args.snapshot_dir) / f"bench_ref_{n}.hnsw"
        ix.serialize(snap, packed=True)
        del ix
        ref = O.OracleIndex.deserialize(METRIC, snap)
        snap.unlink()
        built_by = "GPU engine snapshot (setup only; timed path is the oracle)"
    elif n > probe:
        ref.add(data[probe:], threads=cores)
    log(f"[bench-ref] graph ready in {time.time() - t0:.1f}s ({built_by})")
    ns = min(1000, nq)
    truth, _ = ref.exact_knn(queries[:ns], K, threads=cores)
    # (threaded Add hands out ids in completion order; exact_knn ran on the index's own ids, so they agree)

    def search_ef(ef):
        ref.set_min_nn(ef)
        return ref.knn_query(queries[:ns], K, threads=cores)[0]
    ef, rec = pick_ef(search_ef, truth, k=K)
    ref.set_min_nn(ef)
    sample = min(nq, args.reference_sample)
    for _ in range(args.warmup):
        ref.knn_query(queries[:sample], K, threads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        ref.knn_query(queries[:sample], K, threads=cores)
    el = time.perf_counter() - t
    value = sample * args.steps / el
    line = {
        "impl": "reference",
        "metric": W["metric_name"], "value": round(value, 1), "unit": "queries/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(el / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload string as the B200 arm prints (the arms differ in who built the graph and in the sample size,
        # which have their own keys)
        "config": {"workload": f"{W['tag']}: {n} x {DIM} fp32 {METRIC} {W['shape']}, M={M}, efConstruction={EFC}, "
                               f"{nq}-query batch per GPU per step, k={K}, MinNN(ef)={ef} = smallest grid value with recall@{K}>=0.95",
                   f"recall_at_{K}": round(rec, 4), "ef": ef, "index_nodes": n, "queries_per_step_per_gpu": sample, "graph_built_by": built_by},
        "cpu_baseline": {"value": round(value, 1), "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} queries per step on the {n}-node graph; C++ restatement of the reference "
                                   f"(the .NET runtime is not available), Parallel.For stand-in on all host threads"},
        "e2e": {"value": round(value, 1), "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- partitioned mode (C4)
def partitioned_pass(args, steps, warmup, per_rank=None, nq=None):
    """BASELINE config C4 (10M x 96 ucosine over 8 GPUs = 1.25M per GPU): every rank owns one sub-graph and searches ALL
    queries on it; the per-rank top-k lists are exchanged and merged.  Two exchanges are timed on the same sub-graphs:
      fused      the search kernel's epilogue stores every query's top-k into every peer's gather buffer (CUDA IPC pointers,
                 NVLink stores) + flags; ONE wait+merge kernel follows in stream order -- no collective
      allgather  search -> NCCL all_gather of ids and distances -> merge kernel
    All times are CUDA-event times on the stream the kernels run on, max over ranks.  The process group must be initialised
    (world > 1) and the device set.  Returns the record on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    import hnswindex
    from hnswindex import partitioned as P
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dim, rank_lr, k = 96, 16, 10
    per_rank = per_rank or args.n_per_rank
    nq = nq or args.nq
    total = per_rank * world
    rng = np.random.default_rng(99)
    A = rng.standard_normal((rank_lr, dim)).astype(np.float32)

    def gen(n, seed):
        r = np.random.default_rng(seed)
        out = np.empty((n, dim), dtype=np.float32)
        for s0 in range(0, n, 131072):
            e = min(n, s0 + 131072)
            out[s0:e] = r.standard_normal((e - s0, rank_lr)).astype(np.float32) @ A + 0.25 * r.standard_normal((e - s0, dim)).astype(np.float32)
        out /= np.linalg.norm(out, axis=1, keepdims=True)
        return out
    s0, e0 = P.shard_range(total, rank, world)
    shard = gen(e0 - s0, 1234 + rank)          # shard r of the data set (independent seeds = disjoint ranges)
    queries = gen(nq, 4321)                    # identical on every rank
    ix = hnswindex.Index(dim, "ucosine")
    ix.set_collection_size(e0 - s0); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
    tb = time.time(); ix.add(shard); build_s = time.time() - tb
    build_kernel_ms = ix.stats()["last_build_ms"]
    del shard
    stream = torch.cuda.current_stream()
    ix.set_stream(stream.cuda_stream)
    pix = P.PartitionedIndex(rank, world, total, index=ix)
    # ground truth: exact top-k per shard, merged the same way
    ns = min(1000, nq)
    e_ids, e_d = ix.exact_knn(queries[:ns], k)
    e_ids = np.where(e_ids >= 0, e_ids + s0, -1).astype(np.int32)
    g_ids, g_d = pix._gather(torch.from_numpy(e_ids).cuda(), torch.from_numpy(e_d).cuda())
    truth, _ = P.merge_topk_numpy(g_ids.cpu().numpy(), g_d.cpu().numpy(), k)
    ef = None
    for cand in EF_GRID:
        ix.set_min_nn(cand)
        r = recall_at_k(pix.knn_query(queries[:ns], k)[0], truth)
        if rank == 0:
            log(f"  [partitioned] ef={cand}: recall@10={r:.4f}")
        if r >= 0.95:
            ef = cand
            break
    ix.set_min_nn(ef or EF_GRID[-1])
    q_dev = torch.from_numpy(queries).cuda()

    def timed(fn, end=None):
        for i in range(warmup):
            fn()
        if end:
            end()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        e0_.record(stream)
        for i in range(steps):
            fn()
        e1_.record(stream)
        if end:
            end()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t
        tt = torch.tensor([e0_.elapsed_time(e1_), wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt[0].item()) / steps, float(tt[1].item()) / steps      # ms per batch: device, wall
    # (0) the local search alone (no exchange, no merge): what the exchange is measured against
    ids_loc = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    d_loc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    #     queued back to back on ONE stream exactly like the fused batches below (no host wait between batches, no two-stream
    #     tail overlap), so that the difference is the exchange + merge and nothing else
    ix.set_tuning("async_overlap", 0)
    ms_search, _ = timed(lambda: ix.knn_query_device_async(q_dev.data_ptr(), nq, k, ids_loc.data_ptr(), d_loc.data_ptr()), end=ix.synchronize)
    ix.set_tuning("async_overlap", 1)
    # (a) NCCL path: search, all-gather of the per-rank top-k, merge kernel -- queries resident on the device
    ms_gather, _ = timed(lambda: pix.knn_query_device(q_dev, k))
    ref_ids, ref_d = [x.cpu().numpy() for x in pix.knn_query_device(q_dev, k)]
    # (b) fused path, batches queued back to back on the stream (results stay on the device; one host wait at the end)
    pix.enable_fused(nq, k)
    ms_fused, _ = timed(lambda: pix.knn_query_device(q_dev, k, sync=False), end=pix.sync)
    ms_fused_sync, _ = timed(lambda: pix.knn_query_device(q_dev, k))          # a host wait per batch
    f_ids, f_d = [x.cpu().numpy() for x in pix.knn_query_device(q_dev, k)]
    same = bool(np.array_equal(f_ids, ref_ids) and np.array_equal(f_d.view(np.uint32), ref_d.view(np.uint32)))
    q_pin = torch.from_numpy(queries).pin_memory()

    def step_e2e():                                       # pinned host queries in, host results out
        ids, d = pix.knn_query_device(q_pin.to("cuda", non_blocking=True), k)
        return ids.cpu(), d.cpu()
    _, ms_e2e_wall = timed(step_e2e)
    r = recall_at_k(f_ids[:ns], truth)
    same_all = torch.tensor([1 if same else 0], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(same_all, op=dist.ReduceOp.MIN)
    pix.disable_fused()
    ix.set_stream(0)
    del pix, ix
    if rank != 0:
        return None
    return {
        "workload": f"C4: {total} x {dim} fp32 ucosine (rank-16 + 0.25 noise, L2-normalised) in {world} sub-graphs of {per_rank}, "
                    f"{nq} queries seen by every rank, k={k}, ef={ef}; every rank ends with the merged global top-k",
        "qps_fused": round(nq / (ms_fused / 1e3), 1), "qps_allgather": round(nq / (ms_gather / 1e3), 1),
        "qps_fused_host_wait_per_batch": round(nq / (ms_fused_sync / 1e3), 1),
        "local_search_ms": round(ms_search, 4), "fused_ms": round(ms_fused, 4), "allgather_ms": round(ms_gather, 4),
        "exchange_plus_merge_ms": round(ms_fused - ms_search, 4), "allgather_exchange_plus_merge_ms": round(ms_gather - ms_search, 4),
        "fused_equals_allgather": bool(int(same_all.item()) == 1), "recall_at_10": round(r, 4), "ef": ef,
        "e2e_qps": round(nq / (ms_e2e_wall / 1e3), 1), "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * k * 8,
        "build_s_per_shard": round(build_s, 2), "build_kernel_inserts_per_s_per_gpu": round((e0 - s0) / (build_kernel_ms / 1e3)),
        "steps": steps, "warmup": warmup, "timing": "CUDA events on the stream of the kernels, max over ranks",
        "exchange": "search-kernel epilogue stores into every peer's gather buffer (CUDA IPC pointers over NVLink) + flag, "
                    "one wait+merge kernel; no collective",
        "gpu_launches_per_step": 2,
        "l2_policy": f"inputs larger than L2: {per_rank * (dim * 4 + 128) / 1e6:.0f} MB sub-graph per GPU vs 126 MB L2"}


def run_partitioned(args):
    """`--mode partitioned`: the C4 pass alone, printed as its own line (secondary workload, own metric name)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from hnswindex import bindings as hb
    hb.set_device(local)
    rec = partitioned_pass(args, args.steps, args.warmup)
    if rank == 0:
        nq = args.nq
        print(json.dumps({
            "metric": "QPS at recall@10>=0.95, partitioned index (C4: 1.25M x 96 ucosine per GPU, search fused with the peer-store exchange of top-k)",
            "value": rec["qps_fused"], "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": rec["fused_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": rec,
            "e2e": {"value": rec["e2e_qps"], "unit": "queries/s", "h2d_bytes_per_step": rec["h2d_bytes_per_step"],
                    "d2h_bytes_per_step": rec["d2h_bytes_per_step"]},
            "gpu_launches": int(args.steps * 2)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------- construction (C5)
def run_build(args):
    """`--mode build`: the C5 pass alone, printed as its own line (secondary workload, own metric name)."""
    import torch
    import __graft_entry__ as ge
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    ge.build()
    print(json.dumps(build_pass(args, args.workload, max(1, min(args.steps, 5)))), flush=True)


def build_pass(args, workload, steps):
    """BASELINE config C5: batched GPU construction (hnsw_add) of the C2 data set vs the reference's Parallel.For Add
    (oracle port on all host threads, bounded sample), with a recall-parity check.
    A step = one complete build of the n-node index.  Returns the JSON line as a dict."""
    import torch
    import hnswindex
    W = WORKLOADS[workload]
    DIM, RANK, K, M, EFC, METRIC = W["dim"], W["rank"], W["k"], W["M"], W["efc"], W["metric"]
    n = args.n
    data, queries = make_data(n, 2000, dim=DIM, rank=RANK)
    pinned = torch.from_numpy(data).pin_memory()
    f32p, i32p = ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)
    out_ids = np.empty(n, dtype=np.int32)

    def new_index(size):
        ix = hnswindex.Index(DIM, METRIC)
        ix.set_collection_size(size); ix.set_max_edges(M); ix.set_max_candidates(EFC); ix.set_allow_removals(False)
        if W["dist_rate"]:
            ix.set_distribution_rate(W["dist_rate"])
        ix._create()
        return ix

    walls, kms, prof = [], [], None
    sampler = ClockSampler(0)
    for i in range(1 + steps):           # first build = warm-up (context, allocator)
        ix = new_index(n)
        if i == 1:
            sampler.start()
        torch.cuda.synchronize()
        t = time.perf_counter()
        rc = hnswindex.lib.hnsw_add(ix._h, ct.cast(pinned.data_ptr(), f32p), n, DIM, out_ids.ctypes.data_as(i32p))
        el = time.perf_counter() - t
        assert rc == n, hnswindex.last_error()
        if i > 0:
            walls.append(el); kms.append(ix.stats()["last_build_ms"])
            prof = ix.build_profile()
        if i < steps:
            del ix
    clocks = sampler.stop()
    wall, kernel_ms = float(np.mean(walls)), float(np.mean(kms))
    nd = prof["ndist_search"] + prof["ndist_select"] + prof["ndist_prune"]
    peak, peak_src = measured_peak()
    ach = nd * DIM * 4 / (kernel_ms / 1e3) / 1e9
    truth, _ = ix.exact_knn(queries, K)
    ef, rec = pick_ef(lambda e: (ix.set_min_nn(e), ix.knn_query(queries, K)[0])[1], truth, k=K)
    # CPU baseline + recall parity on a bounded sample: both engines build the first `sample` vectors
    cpu, parity = None, None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        cores = os.cpu_count() or 1
        sample = min(n, args.build_sample)
        okw = {"distribution_rate": W["dist_rate"]} if W["dist_rate"] else {}
        ref = O.OracleIndex(METRIC, max_edges=M, max_candidates=EFC, collection_size=sample, allow_removals=False, **okw)
        t = time.perf_counter(); oids = ref.add(data[:sample], threads=cores); t_cpu = time.perf_counter() - t
        small = new_index(sample)
        gids = small.add(data[:sample])
        st, _ = small.exact_knn(queries, K)
        small.set_min_nn(ef); ref.set_min_nn(ef)
        r_gpu = recall_at_k(small.knn_query(queries, K)[0], st)
        inv = np.empty(sample, dtype=np.int64); inv[oids] = np.arange(sample)   # threaded Add hands out ids in completion order
        o_res = ref.knn_query(queries, K, threads=cores)[0]
        r_cpu = recall_at_k(np.where(o_res >= 0, inv[np.maximum(o_res, 0)], -1), st)
        cpu = {"value": round(sample / t_cpu, 1), "unit": "inserts/s", "cores": cores, "kind": "port",
               "sample": f"the first {sample} vectors of the same data set built by the oracle's threaded Add (Parallel.For stand-in, "
                         f"all host threads); insert cost grows ~log n, so the full-size CPU rate is lower still"}
        parity = {"nodes": sample, "ef": ef, f"recall_at_{K}_gpu_built": round(r_gpu, 4), f"recall_at_{K}_oracle_built": round(r_cpu, 4)}
    return ({
        "metric": f"build inserts/s, {n} x {DIM} {METRIC}, M={M}, efConstruction={EFC} (C5)", "value": round(n / (kernel_ms / 1e3), 1),
        "unit": "inserts/s", "n_gpus": 1, "steps": steps, "warmup": 1, "ms_per_step": round(kernel_ms, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C5: batched GPU construction of {n} x {DIM} fp32 {METRIC} {W['shape']}, M={M}, efConstruction={EFC}; "
                               f"value = device time of all insertion batches (vectors resident), e2e = hnsw_add from pinned host memory",
                   f"recall_at_{K}": round(rec, 4), "ef": ef, "recall_parity": parity,
                   "l2_policy": "every step builds a fresh index: 640 MB of vectors + adjacency vs 126 MB L2"},
        "e2e": {"value": round(n / wall, 1), "unit": "inserts/s", "h2d_bytes_per_step": n * DIM * 4, "d2h_bytes_per_step": n * 4,
                "ms_per_step": round(wall * 1e3, 1), "api": "hnsw_add (C ABI, pinned host buffer)"},
        "gpu_launches": int(ix.stats()["build_launches"]), "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None,
                     "peak_source": peak_src, "kernel": "build_select_kernel + build_link_kernel (all batches)",
                     "bytes_model": "GPU-counted distance evaluations (search + heuristic + overflow prune) x dim x 4",
                     "ndist_per_insert": {"search": round(prof["ndist_search"] / n, 1), "heuristic": round(prof["ndist_select"] / n, 1),
                                          "overflow_prune": round(prof["ndist_prune"] / n, 1)}},
        "cpu_baseline": cpu})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="index size (C2 = 1M)")
    ap.add_argument("--nq", type=int, default=10_000, help="queries per step per GPU (C2 = 10k)")
    ap.add_argument("--query-batches", type=int, default=4, help="distinct query batches rotated across steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--snapshot-dir", default="/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    ap.add_argument("--reference-sample", type=int, default=10_000)
    ap.add_argument("--reference-build-budget-s", type=float, default=150.0)
    ap.add_argument("--mode", default="replicated", choices=["replicated", "partitioned", "build"],
                    help="partitioned = BASELINE config C4, build = C5 (secondary workloads, own metric names)")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="replicated mode: c2 (the metric's config) or c3")
    ap.add_argument("--build-sample", type=int, default=400_000, help="build mode: nodes the CPU oracle builds (bounded sample, ~10-15 s on 16 cores)")
    ap.add_argument("--n-per-rank", type=int, default=1_250_000)
    ap.add_argument("--no-partitioned", action="store_true", help="N > 1: skip the partitioned-index (C4) sub-record")
    ap.add_argument("--extras", default="c5,c2u,c3", help="N = 1, workload c2: the other named shapes measured as sub-records "
                    "(c5 = construction, c2u = uniform hard case, c3 = 1M x 960 cosine); '' = none")
    ap.add_argument("--extras-budget-s", type=float, default=400.0, help="no further extra workload is started after this many seconds")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3 if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "partitioned":
        run_partitioned(args)
    elif args.mode == "build":
        run_build(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
