"""Exploration: GPU build of N C2-shaped vectors; prints time.  Run under `ncu --metrics gpu__time_duration.sum -k regex:build_` for the kernel split."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np
import bench, hnswindex
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
data, q = bench.make_data(N, 2000)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N); ix.set_max_candidates(200); ix.set_allow_removals(False)
for kv in sys.argv[2:]:
    k, v = kv.split("="); ix_t = (k, int(v))
t = time.time(); ix.add(data[:1000]); 
for kv in sys.argv[2:]:
    k, v = kv.split("="); ix.set_tuning(k, int(v))
ix.add(data[1000:]); tb = time.time() - t
gt, _ = ix.exact_knn(q, 10)
out = {"N": N, "build_s": round(tb, 3), "inserts_per_s": round(N / tb), "last_add_kernel_ms": round(ix.stats()["last_build_ms"], 1)}
for ef in (32, 48, 64):
    ix.set_min_nn(ef); ids, _ = ix.knn_query(q, 10)
    out[f"recall@10_ef{ef}"] = round(float(np.mean([len(set(ids[i]) & set(gt[i])) / 10 for i in range(len(q))])), 4)
bp = ix.build_profile()
cyc = {k: v for k, v in bp.items() if k.startswith("cyc_")}
sel_total = max(1, bp["cyc_select_kernel"]); link_total = max(1, bp["cyc_link_kernel"])
out["profile"] = {"select_share": {k: round(bp[k] / sel_total, 3) for k in ("cyc_descent", "cyc_search", "cyc_select")},
                  "link_share": {k: round(bp[k] / link_total, 3) for k in ("cyc_prune", "cyc_lock")},
                  "select_vs_link_cycles": round(sel_total / link_total, 2),
                  "per_insert": {k: round(bp[k] / max(1, bp["inserts"]), 1) for k in ("ndist_search", "ndist_select", "ndist_prune", "prunes", "links")}}
print(json.dumps(out), flush=True)
