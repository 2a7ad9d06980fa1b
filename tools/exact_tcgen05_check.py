"""One-shot check of the tcgen05 / TMEM pass A of hnsw_exact_knn_tf32 (tuning key "exact_tcgen05") against the FFMA scan:
ids and distance bits must be identical.  Run under `timeout` on a B200.  Prints two JSON lines.
Usage: python tools/exact_tcgen05_check.py [mode]   (mode 1 = the default tcgen05 kernel, 2 = the experimental
warp-specialised one, see csrc/exact_tf32_kernels.cu)"""
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
from hnswindex.bindings import lib  # noqa: E402

MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 1
out = {"mode": MODE, "cases": [], "ok": True}
rng = np.random.default_rng(21)
for metric, dim, n in (("sq_euclid", 128, 5000), ("cosine", 100, 5000), ("ucosine", 96, 5000), ("euclid", 64, 5000), ("sq_euclid", 3, 2000),
                       ("sq_euclid", 128, 37)):
    data = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((150, dim)).astype(np.float32)
    if metric == "ucosine":
        data /= np.linalg.norm(data, axis=1, keepdims=True)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    ix = hnswindex.Index(dim, metric)
    ix.set_max_edges(4)
    ix.set_max_candidates(8)
    ix.set_collection_size(n)
    ix.add(data)
    if n >= 2000:
        ix.remove(np.arange(0, n, 7, dtype=np.int32))
    ix.set_tuning("exact_tcgen05", MODE)
    for k in (1, 10):
        sup = lib.hnsw_exact_tcgen05_supported(dim, k)
        a_ids, a_d = ix.exact_knn(q, k)
        b_ids, b_d = ix.exact_knn(q, k, tensor_cores=True)
        same = bool(np.array_equal(a_ids, b_ids) and np.array_equal(a_d.view(np.uint32), b_d.view(np.uint32)))
        out["cases"].append({"metric": metric, "dim": dim, "n": n, "k": k, "tcgen05": sup, "identical": same,
                             "rows_differ": int((a_ids != b_ids).any(axis=1).sum()), "fallbacks": ix.last_exact_fallbacks})
        out["ok"] = out["ok"] and same and sup == 1
print(json.dumps(out), flush=True)

# timing at a larger size (the same recipe as the bench data)
n, nq = 200_000, 2000
A = rng.standard_normal((16, 128)).astype(np.float32)
x = (rng.standard_normal((n + nq, 16)).astype(np.float32) @ A + 0.25 * rng.standard_normal((n + nq, 128)).astype(np.float32)).astype(np.float32)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_max_edges(4)
ix.set_max_candidates(8)
ix.set_collection_size(n)
ix.set_allow_removals(False)
ix.add(x[:n])
q = x[n:]
res = {}
ref = None
for name, tc, t5 in (("scan", False, 0), ("mma_sync", True, 0), ("tcgen05", True, MODE)):
    ix.set_tuning("exact_tcgen05", t5)
    ix.exact_knn(q[:128], 10, tensor_cores=tc)
    best = 1e9
    for _ in range(3):
        t = time.time()
        r = ix.exact_knn(q, 10, tensor_cores=tc)
        best = min(best, time.time() - t)
    res[name + "_s"] = round(best, 4)
    if ref is None:
        ref = r
    else:
        res[name + "_identical"] = bool(np.array_equal(r[0], ref[0]) and np.array_equal(r[1].view(np.uint32), ref[1].view(np.uint32)))
        res[name + "_fallbacks"] = ix.last_exact_fallbacks
print(json.dumps(res), flush=True)
