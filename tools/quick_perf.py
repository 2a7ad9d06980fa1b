"""Exploration script (not the bench): build N sift-like vectors on the GPU, sweep ef, print QPS / recall / work."""
import sys, time, json, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np
import __graft_entry__ as ge
ge.build()
import hnswindex

def sift_like(n, dim=128, rank=16, seed=1234, noise=0.25):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((rank, dim)).astype(np.float32)
    out = np.empty((n, dim), dtype=np.float32)
    for s in range(0, n, 100000):
        e = min(n, s + 100000)
        z = rng.standard_normal((e - s, rank)).astype(np.float32)
        out[s:e] = z @ A + noise * rng.standard_normal((e - s, dim)).astype(np.float32)
    return out, A

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
efs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [16, 24, 32, 48, 64, 96]
data, A = sift_like(N)
rng = np.random.default_rng(4321)
q = (rng.standard_normal((NQ, 16)).astype(np.float32) @ A + 0.25 * rng.standard_normal((NQ, 128)).astype(np.float32)).astype(np.float32)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N); ix.set_max_candidates(200); ix.set_allow_removals(False)
for key, val in [a.split("=") for a in sys.argv[4:]]:
    pass
t = time.time(); ix.add(data); tb = time.time() - t
print(json.dumps({"N": N, "build_s": tb, "inserts_per_s": N / tb, "build_kernel_ms": ix.stats()["last_build_ms"]}), flush=True)
t = time.time(); gt, _ = ix.exact_knn(q[:2000], 10); print("exact_knn 2000 q: %.3fs" % (time.time() - t), flush=True)
ix.enable_stats(True)
for ef in efs:
    ix.set_min_nn(ef)
    ids, d = ix.knn_query(q, 10)
    st = ix.stats()
    rec = float(np.mean([len(set(ids[i]) & set(gt[i])) / 10 for i in range(2000)]))
    ms = st["last_search_ms"]
    print(json.dumps({"ef": ef, "recall@10": round(rec, 4), "kernel_ms": round(ms, 3), "qps_kernel": round(NQ / ms * 1e3), "ndist_per_q": st["ndist"] / NQ,
                      "nexp_per_q": st["nexp"] / NQ, "gathered_GBps": round(st["ndist"] * 512 / ms / 1e6, 1)}), flush=True)
ix.enable_stats(False)
for ef in efs:
    ix.set_min_nn(ef)
    ts = []
    for _ in range(5):
        t = time.time(); ix.knn_query(q, 10); ts.append(time.time() - t)
    print(json.dumps({"ef": ef, "e2e_ms_best": round(min(ts) * 1e3, 3), "kernel_ms": round(ix.stats()["last_search_ms"], 3)}), flush=True)
