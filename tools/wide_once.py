"""One search launch of a chosen kernel variant on the C2 index, for ncu.  Usage: python tools/wide_once.py <ef> <warp|wl|wide1|wide2p|wide4p> [nq]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import __graft_entry__ as ge  # noqa: E402

ge.build()
import torch  # noqa: E402
import hnswindex  # noqa: E402
from bench import make_data  # noqa: E402

ef, variant = int(sys.argv[1]), sys.argv[2]
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
n = 1_000_000
data, queries = make_data(n, nq)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(n); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data)
tn = {"warp": dict(wide_min_ef=0, wl_min_ef=0), "wl": dict(wide_min_ef=0, wl_min_ef=1),
      "wide1": dict(wide_min_ef=1, wide_warps=1, wide_prefetch=0), "wide2p": dict(wide_min_ef=1, wide_warps=2, wide_prefetch=1),
      "wide4p": dict(wide_min_ef=1, wide_warps=4, wide_prefetch=1), "wide2s": dict(wide_min_ef=1, wide_warps=2, wide_prefetch=1),
      "wide3s": dict(wide_min_ef=1, wide_warps=3, wide_prefetch=1), "wide2": dict(wide_min_ef=1, wide_warps=2, wide_prefetch=0),
      "bm2": dict(wide_min_ef=1, wide_warps=2, wide_prefetch=1, wide_bitmap_min_ef=1), "default": dict()}[variant]
for k, v in tn.items():
    ix.set_tuning(k, v)
ix.set_min_nn(ef)
q_dev = torch.from_numpy(queries).cuda()
ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda"); d = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
for _ in range(2):
    ix.knn_query_device(q_dev.data_ptr(), nq, 10, ids.data_ptr(), d.data_ptr())
    print(ef, variant, ix.stats()["last_search_ms"], "ms")
