"""One C3-shaped search launch (1M x 960 cosine, M=32, k=100, ef=192) for ncu.  Usage: python tools/c3_once.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import hnswindex  # noqa: E402

N = 1000000
data, q = bench.make_data(N, 10000, dim=960, rank=32)
ix = hnswindex.Index(960, "cosine")
ix.set_collection_size(N); ix.set_max_edges(32); ix.set_max_candidates(200); ix.set_allow_removals(False); ix.set_distribution_rate(1 / np.log(32))
ix.add(data)
del data
ix.set_min_nn(192)
qd = torch.from_numpy(q).cuda()
ids = torch.empty((10000, 100), dtype=torch.int32, device="cuda"); dd = torch.empty((10000, 100), dtype=torch.float32, device="cuda")
for _ in range(2):
    ix.knn_query_device(qd.data_ptr(), 10000, 100, ids.data_ptr(), dd.data_ptr())
    print(ix.stats()["last_search_ms"], "ms")
