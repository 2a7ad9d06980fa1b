"""Construction schedule sweep on the C5 workload: kernel time / recall of GPU-built graphs per batch schedule.
Usage: python tools/build_sweep.py [n]"""
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
data, q = bench.make_data(N, 2000)
truth = None
cfgs = [dict(), dict(dim128=0), dict(build_tail_fill=0), dict()]
for cfg in cfgs:
    ix = hnswindex.Index(128, "sq_euclid")
    ix.set_collection_size(N); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
    ix._create()
    for k, v in cfg.items():
        ix.set_tuning(k, v)
    t = time.perf_counter(); ix.add(data); wall = time.perf_counter() - t
    st = ix.stats()
    if truth is None:
        truth, _ = ix.exact_knn(q, 10, tensor_cores=True)
    rec = {}
    for ef in (32, 48, 96):
        ix.set_min_nn(ef)
        rec[ef] = round(bench.recall_at_k(ix.knn_query(q, 10)[0], truth), 4)
    print(json.dumps({"cfg": cfg, "kernel_ms": round(st["last_build_ms"], 1), "wall_s": round(wall, 3), "launches": st["build_launches"],
                      "recall": rec}), flush=True)
    del ix
