"""Does query ORDER matter?  The same 10k C2 queries in their natural order and sorted by the id of their nearest neighbour (a proxy
for 'queries that walk the same region of the graph run at the same time'): kernel time per batch."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
import hnswindex  # noqa: E402

n, nq = 1_000_000, 10_000
data, q = bench.make_data(n, nq)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(n); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data)
for ef in (48, 128):
    ix.set_min_nn(ef)
    ids0, _ = ix.knn_query(q, 10)
    # proxies for locality: nearest neighbour's id is meaningless spatially (ids are insertion order), so use a projection:
    # order by the first principal direction of the data, and by a 2-level key (sign pattern of 8 random projections)
    rng = np.random.default_rng(0)
    proj = rng.standard_normal((128, 12)).astype(np.float32)
    bits = (q @ proj > 0).astype(np.int64)
    key = (bits * (1 << np.arange(12))).sum(1)
    orders = {"natural": np.arange(nq), "lsh12": np.argsort(key, kind="stable"), "first_pc": np.argsort(q @ np.linalg.svd(data[:20000] - data[:20000].mean(0), full_matrices=False)[2][0])}
    # and the ideal: group by the layer-0 entry region = sort by nearest neighbour's coordinates along the same projection
    nn_key = ((data[ids0[:, 0]] @ proj > 0).astype(np.int64) * (1 << np.arange(12))).sum(1)
    orders["lsh12_of_nn"] = np.argsort(nn_key, kind="stable")
    out = {"ef": ef}
    ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda"); d = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
    for name, o in orders.items():
        qd = torch.from_numpy(np.ascontiguousarray(q[o])).cuda()
        ms = []
        for _ in range(6):
            ix.knn_query_device(qd.data_ptr(), nq, 10, ids.data_ptr(), d.data_ptr())
            ms.append(ix.stats()["last_search_ms"])
        out[name] = round(float(np.median(ms[1:])), 4)
    print(json.dumps(out), flush=True)
