"""Exploration: build C2 once on the GPU, then sweep tuning knobs of the search kernel (kernel ms via CUDA events)."""
import sys, time, json, itertools
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np
import torch
import bench
import hnswindex

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 48
configs = [dict(kv.split("=") for kv in c.split(",")) for c in sys.argv[3:]] or [{}]
data, q = bench.make_data(N, 40000)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N); ix.set_max_candidates(200); ix.set_allow_removals(False)
t = time.time(); ix.add(data); print("build_s", round(time.time() - t, 2), "kernel_ms", round(ix.stats()["last_build_ms"], 1), flush=True)
ix.set_min_nn(ef)
qd = [torch.from_numpy(q[i * 10000:(i + 1) * 10000]).cuda() for i in range(4)]
ids = torch.empty((10000, 10), dtype=torch.int32, device="cuda"); dd = torch.empty((10000, 10), dtype=torch.float32, device="cuda")
ref_ids = None
for cfg in configs:
    for k, v in cfg.items():
        ix.set_tuning(k, int(v))
    ix.enable_stats(True)
    ix.knn_query_device(qd[0].data_ptr(), 10000, 10, ids.data_ptr(), dd.data_ptr())
    st = ix.stats()
    cur = ids.cpu().numpy().copy()
    if ref_ids is None: ref_ids = cur
    same = float((cur == ref_ids).all(axis=1).mean())
    ix.enable_stats(False)
    for i in range(3): ix.knn_query_device(qd[i % 4].data_ptr(), 10000, 10, ids.data_ptr(), dd.data_ptr())
    ms = []
    for i in range(12):
        ix.knn_query_device(qd[i % 4].data_ptr(), 10000, 10, ids.data_ptr(), dd.data_ptr()); ms.append(ix.stats()["last_search_ms"])
    print(json.dumps({"cfg": cfg, "kernel_ms": round(float(np.mean(ms)), 4), "min_ms": round(min(ms), 4), "qps": round(10000 / np.mean(ms) * 1e3),
                      "ndist_per_q": st["ndist"] / 10000, "nexp_per_q": st["nexp"] / 10000, "same_as_first": same}), flush=True)
