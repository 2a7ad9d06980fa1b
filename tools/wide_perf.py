"""Wide-beam search on the C2 data (1M x 128 sift-like, M=16, efC=200): QPS of the one-warp kernel vs the CTA-per-query kernel
per ef, device-resident 10k-query batches, CUDA-event kernel time.  Usage: python tools/wide_perf.py [n] [nq] [uniform]"""
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
import torch  # noqa: E402
import hnswindex  # noqa: E402
from bench import make_data, recall_at_k  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
uniform = len(sys.argv) > 3 and sys.argv[3] == "uniform"
data, queries = make_data(n, nq, rank=0 if uniform else 16)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(n); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
t = time.time(); ix.add(data); print(f"build {time.time() - t:.1f}s", flush=True)
truth, _ = ix.exact_knn(queries[:1000], 10, tensor_cores=True)
q_dev = torch.from_numpy(queries).cuda()
ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda"); d = torch.empty((nq, 10), dtype=torch.float32, device="cuda")


def run(ef, reps=3, **tuning):
    for k, v in tuning.items():
        ix.set_tuning(k, v)
    ix.set_min_nn(ef)
    best = 1e9
    m = nq if ef <= 2048 else min(nq, max(3000, int(nq * 256 / ef)))    # keep a run short at huge ef
    for _ in range(reps):
        ix.knn_query_device(q_dev.data_ptr(), m, 10, ids.data_ptr(), d.data_ptr())
        best = min(best, ix.stats()["last_search_ms"])
    rec = recall_at_k(ids[:min(m, 1000)].cpu().numpy(), truth[:min(m, 1000)])
    return m / (best / 1e3), rec, ids[:m].cpu().numpy().copy()


efs = [int(x) for x in (sys.argv[4].split(",") if len(sys.argv) > 4 else "48,64,96,128,192,256,384,512,1024,2048,4096".split(","))]
for ef in efs:
    row = {"ef": ef}
    base_ids = None
    for name, tn in (("warp", dict(wide_min_ef=0, wl_min_ef=0)), ("wl", dict(wide_min_ef=0, wl_min_ef=1)),
                     ("wide2s", dict(wide_min_ef=1, wide_warps=2, wide_prefetch=1, wide_bitmap_min_ef=0)),
                     ("wide3s", dict(wide_min_ef=1, wide_warps=3, wide_prefetch=1, wide_bitmap_min_ef=0)),
                     ("bm2", dict(wide_min_ef=1, wide_warps=2, wide_prefetch=1, wide_bitmap_min_ef=1)),
                     ("bm3", dict(wide_min_ef=1, wide_warps=3, wide_prefetch=1, wide_bitmap_min_ef=1)),
                     ("bm4", dict(wide_min_ef=1, wide_warps=4, wide_prefetch=1, wide_bitmap_min_ef=1))):
        if name == "warp" and ef > 4096:
            continue
        try:
            qps, rec, got = run(ef, **tn)
        except RuntimeError as e:
            row[name] = str(e)[:60]
            continue
        row[name] = round(qps)
        if name.startswith("wide") or name.startswith("bm"):
            pr = ix.wide_profile(reset=True)
            ex = max(1, pr["expansions"])
            row[name + "_prof"] = {"hit": round(pr["spec_used"] / max(1, pr["spec_issued"]), 3), "issued": round(pr["spec_issued"] / ex, 3),
                                   "cyc/exp": {k[4:]: round(pr[k] / ex) for k in pr if k.startswith("cyc_") and k != "cyc_descent"},
                                   "descent": round(pr["cyc_descent"] / max(1, pr["queries"]))}
        row["recall"] = round(rec, 4)
        if base_ids is None:
            base_ids = got
        else:
            row[name + "_same"] = bool(np.array_equal(got, base_ids))
    print(json.dumps(row), flush=True)
