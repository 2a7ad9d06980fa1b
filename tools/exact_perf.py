"""Exact top-k at the C2 size: FFMA scan (hnsw_exact_knn) against the tensor-core variant (hnsw_exact_knn_tf32).
Usage: python tools/exact_perf.py [n_items] [n_queries] [repetitions]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import hnswindex  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
rng = np.random.default_rng(1234)
A = rng.standard_normal((16, 128)).astype(np.float32)
x = (rng.standard_normal((n + nq, 16)).astype(np.float32) @ A + 0.25 * rng.standard_normal((n + nq, 128)).astype(np.float32)).astype(np.float32)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_max_edges(4)
ix.set_max_candidates(8)
ix.set_collection_size(n)
ix.set_allow_removals(False)
t = time.time()
ix.add(x[:n])
print("build (M=4, efC=8): %.2f s" % (time.time() - t), flush=True)
q = x[n:]
out = {"n": n, "nq": nq, "dim": 128}
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
for k in (10,) if reps == 1 else (10, 100):
    ref = None
    for name, tc, t5 in (("scan", False, 0), ("mma_sync", True, 0), ("tcgen05", True, 1), ("tcgen05_ws", True, 2)):
        ix.set_tuning("exact_tcgen05", t5)
        ix.exact_knn(q[:64], k, tensor_cores=tc)
        best = 1e9
        for _ in range(reps):
            t = time.time()
            r = ix.exact_knn(q, k, tensor_cores=tc)
            best = min(best, time.time() - t)
        out[f"{name}_k{k}_s"] = round(best, 4)
        out[f"{name}_k{k}_tflops_fp32_equiv"] = round(2.0 * n * nq * 128 / best / 1e12, 2)
        if ref is None:
            ref = r
        else:
            out[f"{name}_k{k}_fallbacks"] = ix.last_exact_fallbacks
            out[f"{name}_k{k}_identical"] = bool(np.array_equal(r[0], ref[0]) and np.array_equal(r[1].view(np.uint32), ref[1].view(np.uint32)))
print(json.dumps(out))
