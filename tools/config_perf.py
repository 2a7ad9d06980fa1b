"""Exploration: the other BASELINE configs at (scaled) size -- C3-like 960-d cosine M=32 k=100, C4-like 96-d ucosine shard."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np
import hnswindex

def lowrank(n, dim, rank, seed, noise=0.25, normalize=False):
    rng = np.random.default_rng(seed)
    A = np.random.default_rng(99).standard_normal((rank, dim)).astype(np.float32)
    out = np.empty((n, dim), dtype=np.float32)
    for s in range(0, n, 50000):
        e = min(n, s + 50000)
        out[s:e] = rng.standard_normal((e - s, rank)).astype(np.float32) @ A + noise * rng.standard_normal((e - s, dim)).astype(np.float32)
    if normalize: out /= np.linalg.norm(out, axis=1, keepdims=True)
    return out

def run(name, metric, n, dim, rank, M, efc, k, efs, nq=10000, rate=None, normalize=False):
    data = lowrank(n, dim, rank, 1234, normalize=normalize); q = lowrank(nq, dim, rank, 4321, normalize=normalize)
    ix = hnswindex.Index(dim, metric)
    ix.set_collection_size(n); ix.set_max_edges(M); ix.set_max_candidates(efc); ix.set_allow_removals(False)
    if rate: ix.set_distribution_rate(rate)
    t = time.time(); ix.add(data); tb = time.time() - t
    print(json.dumps({"cfg": name, "n": n, "dim": dim, "build_s": round(tb, 2), "inserts_per_s": round(n / tb), "build_kernel_ms": round(ix.stats()["last_build_ms"])}), flush=True)
    for kv in sys.argv[2:]:
        kk, vv = kv.split("="); ix.set_tuning(kk, int(vv))
    t = time.time(); gt, _ = ix.exact_knn(q[:1000], k); te = time.time() - t
    for ef in efs:
        ix.set_min_nn(ef); ix.enable_stats(True)
        ids, d = ix.knn_query(q, k); st = ix.stats(); ix.enable_stats(False)
        ms = []
        for _ in range(5): ix.knn_query(q, k); ms.append(ix.stats()["last_search_ms"])
        rec = float(np.mean([len(set(ids[i]) & set(gt[i])) / k for i in range(1000)]))
        print(json.dumps({"cfg": name, "ef": max(ef, k), "recall@k": round(rec, 4), "kernel_ms": round(min(ms), 3), "qps": round(nq / min(ms) * 1e3),
                          "ndist_per_q": round(st["ndist"] / nq, 1), "gathered_GBps": round(st["ndist"] * dim * 4 / min(ms) / 1e6, 1), "exact_1000q_s": round(te, 2)}), flush=True)

which = (sys.argv[1] if len(sys.argv) > 1 else "all") if __name__ == "__main__" else "none"
if which in ("c3", "all"): run("C3-like 200k x 960 cosine M=32 k=100", "cosine", 200000, 960, 32, 32, 200, 100, [100, 200], rate=1 / np.log(32))
if which in ("c4", "all"): run("C4-shard 1.25M x 96 ucosine M=16 k=10", "ucosine", 1250000, 96, 16, 16, 200, 10, [32, 48, 64], normalize=True)
