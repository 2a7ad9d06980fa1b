"""Randomised cross-check of the search kernels: for random index shapes / widths / table sizes the CTA-per-query kernel (with and
without the exact bitmap, 2-4 warps) and the one-warp kernel on the non-shifting list must return exactly what the one-warp
kernel on the shifting list returns (ids, distance bits, expansion counters).  Usage: python tools/wide_stress.py [cases] [seed]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import hnswindex  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for case in range(cases):
    metric = ["sq_euclid", "cosine", "ucosine", "euclid"][int(rng.integers(4))]
    dim = int(rng.choice([3, 17, 64, 96, 128, 129, 200, 384]))
    n = int(rng.choice([50, 700, 5000, 40000]))
    M = int(rng.choice([2, 8, 16, 24, 40]))
    nq = int(rng.choice([1, 3, 33, 200]))
    k = int(rng.choice([1, 10, 50]))
    data = rng.standard_normal((n, dim)).astype(np.float32)
    if rng.random() < 0.3:
        data = np.round(data * 4) / 4          # coarse grid: many exact distance ties
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    if metric == "ucosine":
        data /= np.maximum(np.linalg.norm(data, axis=1, keepdims=True), 1e-6); q /= np.maximum(np.linalg.norm(q, axis=1, keepdims=True), 1e-6)
    ix = hnswindex.Index(dim, metric)
    ix.set_max_edges(M); ix.set_max_candidates(int(rng.choice([20, 100]))); ix.set_collection_size(n); ix.set_allow_removals(False)
    ix.add(data)
    ix.enable_stats(True)
    for ef in sorted(set(int(x) for x in rng.choice([1, 7, 64, 65, 100, 255, 300, 777, 1300, 2500, 6000], size=4))):
        ix.set_min_nn(ef)
        vis = int(rng.choice([0, 0, 64, 512]))
        ix.set_tuning("vis_size_search", vis)

        def run(**t):
            for kk, v in t.items():
                ix.set_tuning(kk, v)
            ids, d = ix.knn_query(q, k)
            st = ix.stats()
            return ids, d.view(np.uint32), (st["nexp"], st["row_words"])
        base = run(wide_min_ef=0, wl_min_ef=0, wide_max_small_batch=0)
        variants = {"wl": dict(wide_min_ef=0, wl_min_ef=1), "cta2": dict(wide_min_ef=1, wide_warps=2, wide_prefetch=1, wide_bitmap_min_ef=0),
                    "cta3_nospec": dict(wide_min_ef=1, wide_warps=3, wide_prefetch=0, wide_bitmap_min_ef=0),
                    "bitmap2": dict(wide_min_ef=1, wide_warps=2, wide_prefetch=1, wide_bitmap_min_ef=1),
                    "bitmap4": dict(wide_min_ef=1, wide_warps=4, wide_prefetch=1, wide_bitmap_min_ef=1)}
        for name, t in variants.items():
            got = run(**t)
            ok = np.array_equal(got[0], base[0]) and np.array_equal(got[1], base[1]) and got[2] == base[2]
            if not ok:
                bad += 1
                print(json.dumps({"MISMATCH": name, "metric": metric, "dim": dim, "n": n, "M": M, "nq": nq, "k": k, "ef": ef, "vis": vis,
                                  "rows_differ": int((got[0] != base[0]).any(axis=1).sum()), "counters": [got[2], base[2]]}), flush=True)
    del ix
print(json.dumps({"cases": cases, "mismatches": bad}), flush=True)
