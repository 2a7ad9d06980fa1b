"""hnsw_knn_query with page-locked / pageable 10k-query batches on the C2 index: QPS per number of pieces (tuning key sync_pieces)."""
import ctypes as ct
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
import hnswindex  # noqa: E402

N, nq, K = 1000000, 10000, 10
data, q = bench.make_data(N, nq * 4)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data); ix.set_min_nn(48)
f32p, i32p = ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)
q_pin = [torch.from_numpy(q[b * nq:(b + 1) * nq]).pin_memory() for b in range(4)]
ids_pin = torch.empty((nq, K), dtype=torch.int32).pin_memory(); d_pin = torch.empty((nq, K), dtype=torch.float32).pin_memory()
q_np = [np.ascontiguousarray(q[b * nq:(b + 1) * nq]) for b in range(4)]
for pieces in (1, 2, 3, 4, 6, 8):
    ix.set_tuning("sync_pieces", pieces)
    row = {"pieces": pieces}
    for name in ("pinned", "pageable"):
        def step(i):
            if name == "pinned":
                assert hnswindex.lib.hnsw_knn_query(ix._h, ct.cast(q_pin[i % 4].data_ptr(), f32p), nq, 128, K, ct.cast(ids_pin.data_ptr(), i32p),
                                                    ct.cast(d_pin.data_ptr(), f32p)) == 0
            else:
                ix.knn_query(q_np[i % 4], K)
        for i in range(5):
            step(i)
        t = time.perf_counter()
        for i in range(40):
            step(i)
        row[name + "_qps"] = round(40 * nq / (time.perf_counter() - t))
    print(json.dumps(row), flush=True)
