"""Exploration: latency of small batches through the host API (hnsw_knn_query with numpy arrays) on a C2-shaped index."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np
import bench, hnswindex
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
data, q = bench.make_data(N, 20000)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data); ix.set_min_nn(48)
for nb in (1, 4, 16, 64, 256, 1024, 4096, 10000):
    ts = []
    for r in range(30):
        qq = q[(r * nb) % 10000:(r * nb) % 10000 + nb]
        t = time.perf_counter(); ix.knn_query(qq, 10); ts.append(time.perf_counter() - t)
    ts = np.array(ts[5:]) * 1e6
    print(json.dumps({"batch": nb, "median_us": round(float(np.median(ts)), 1), "p90_us": round(float(np.percentile(ts, 90)), 1),
                      "kernel_us": round(ix.stats()["last_search_ms"] * 1e3, 1), "qps": round(nb / np.median(ts) * 1e6)}), flush=True)
