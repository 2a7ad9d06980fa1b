"""Latency of small calls through the host API (hnsw_knn_query with numpy arrays) on a C2-shaped index: the one-warp-per-query
kernel against the CTA-per-query kernel (leader + helpers evaluating the next expansion ahead of time).
Usage: python tools/latency_perf.py [n]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np  # noqa: E402
import bench  # noqa: E402
import hnswindex  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
data, q = bench.make_data(N, 20000)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data)
for ef in (5, 48):
    ix.set_min_nn(ef)
    for nb in (1, 2, 4, 8, 16, 64, 256, 1024):
        row = {"ef": ef, "batch": nb}
        for name, tn in (("warp", dict(wide_max_small_batch=0)), ("cta2", dict(wide_max_small_batch=100000, wide_warps=2)),
                         ("cta3", dict(wide_max_small_batch=100000, wide_warps=3)), ("cta4", dict(wide_max_small_batch=100000, wide_warps=4))):
            for k, v in tn.items():
                ix.set_tuning(k, v)
            ts, ks = [], []
            for r in range(40):
                qq = q[(r * nb) % 10000:(r * nb) % 10000 + nb]
                t = time.perf_counter(); ix.knn_query(qq, 10); ts.append(time.perf_counter() - t)
                ks.append(ix.stats()["last_search_ms"] * 1e3)
            row[name] = {"median_us": round(float(np.median(ts[5:]) * 1e6), 1), "kernel_us": round(float(np.median(ks[5:])), 1)}
        print(json.dumps(row), flush=True)
ix.set_tuning("wide_max_small_batch", 0)
