"""Latency of one- and four-query calls through hnsw_knn_query (pageable numpy buffers) per number of warps of the CTA-per-query kernel,
and of one-vector hnsw_add calls, on the 1M x 128 C2 index.  Usage: python tools/small_call_latency.py"""
import json, sys, time, ctypes as ct
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/hnswindex.net_b200')
import numpy as np, bench, hnswindex
N = 1000000
data, q = bench.make_data(N, 20000)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N + 1000); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data); ix.set_min_nn(48)
f32p, i32p = ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)
for ww in (0, 2, 3, 4, 0):
    ix.set_tuning("wide_warps", ww)
    row = {"wide_warps": ww}
    for nb in (1, 4):
        ids = np.empty((nb, 10), np.int32); dd = np.empty((nb, 10), np.float32)
        ts, ks = [], []
        for r in range(60):
            qq = np.ascontiguousarray(q[(r * nb) % 10000:(r * nb) % 10000 + nb])
            t = time.perf_counter()
            rc = hnswindex.lib.hnsw_knn_query(ix._h, qq.ctypes.data_as(f32p), nb, 128, 10, ids.ctypes.data_as(i32p), dd.ctypes.data_as(f32p))
            ts.append(time.perf_counter() - t); ks.append(ix.stats()["last_search_ms"] * 1e3)
        row[f"b{nb}"] = {"call_us": round(float(np.median(ts[10:]) * 1e6), 1), "kernel_us": round(float(np.median(ks[10:])), 1)}
    print(json.dumps(row), flush=True)
ix.set_tuning("wide_warps", 0)
j = 0
for cfg in (dict(), dict(build_tail_fill=0), dict(dim128=0), dict(build_tail_fill=0, dim128=0), dict()):
    for k_, v_ in dict(build_tail_fill=1, dim128=1).items(): ix.set_tuning(k_, v_)
    for k_, v_ in cfg.items(): ix.set_tuning(k_, v_)
    ts, ks = [], []
    for i in range(30):
        v = np.ascontiguousarray(q[10000 + j:10001 + j]); j += 1; out = np.empty(1, np.int32)
        t = time.perf_counter(); hnswindex.lib.hnsw_add(ix._h, v.ctypes.data_as(f32p), 1, 128, out.ctypes.data_as(i32p)); ts.append(time.perf_counter() - t)
        ks.append(ix.stats()["last_build_ms"] * 1e3)
    print(json.dumps({"cfg": cfg, "single_add_us": round(float(np.median(ts[5:]) * 1e6), 1), "kernel_us": round(float(np.median(ks[5:])), 1)}), flush=True)
