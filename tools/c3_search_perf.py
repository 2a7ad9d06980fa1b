"""Exploration: C3 at full size (1M x 960 cosine, M=32): search kernel time with / without the exact visited bitmap."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np, torch
import bench, hnswindex
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
data, q = bench.make_data(N, 20000, dim=960, rank=32)
ix = hnswindex.Index(960, "cosine")
ix.set_collection_size(N); ix.set_max_edges(32); ix.set_max_candidates(200); ix.set_allow_removals(False); ix.set_distribution_rate(1 / np.log(32))
t = time.time(); ix.add(data); print("build_s", round(time.time() - t, 1), flush=True)
del data
ix.set_min_nn(192)
qd = [torch.from_numpy(q[i * 10000:(i + 1) * 10000]).cuda() for i in range(2)]
ids = torch.empty((10000, 100), dtype=torch.int32, device="cuda"); dd = torch.empty((10000, 100), dtype=torch.float32, device="cuda")
base = None
cfgs = [{"vis_global_search": 0}, {"vis_global_search": 1}, {"vis_global_search": 0}, {"vis_global_search": 1},
        {"vis_global_search": 1, "vis_size_search": 256}, {"vis_global_search": 1, "vis_size_search": 128}, {"vis_global_search": 1, "vis_size_search": 1024}]
for cfg in cfgs:
    for k, v in cfg.items(): ix.set_tuning(k, v)
    ix.enable_stats(True); ix.knn_query_device(qd[0].data_ptr(), 10000, 100, ids.data_ptr(), dd.data_ptr()); st = ix.stats(); ix.enable_stats(False)
    cur = ids.cpu().numpy().copy()
    if base is None: base = cur
    ms = []
    for i in range(5):
        ix.knn_query_device(qd[i % 2].data_ptr(), 10000, 100, ids.data_ptr(), dd.data_ptr()); ms.append(ix.stats()["last_search_ms"])
    print(json.dumps({"cfg": cfg, "kernel_ms": round(float(np.mean(ms[1:])), 2), "qps": round(10000 / (float(np.mean(ms[1:])) / 1e3)), "ndist_per_q": st["ndist"] / 10000,
                      "same": float((cur == base).all(axis=1).mean())}), flush=True)
