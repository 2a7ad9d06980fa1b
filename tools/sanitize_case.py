"""Smallest case that touches every kernel: for compute-sanitizer (one tool per gpurun call)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np
import hnswindex
rng = np.random.default_rng(0)
for metric, dim in (("sq_euclid", 32), ("cosine", 37), ("ucosine", 136)):
    data = rng.random((600, dim), dtype=np.float32)
    if metric == "ucosine": data /= np.linalg.norm(data, axis=1, keepdims=True)
    ix = hnswindex.Index(dim, metric)
    ix.set_collection_size(64); ix.set_max_edges(6); ix.set_max_candidates(70)
    ix.add(data[:300]); 
    for v in data[300:320]: ix.add(v)
    ix.add(data[320:])
    for k, ef in ((1, 5), (10, 40), (3, 100), (2, 300)):
        ix.set_min_nn(ef); ids, d = ix.knn_query(data[:64], k)
        assert (ids[:, 0] >= 0).all()
    e_ids, _ = ix.exact_knn(data[:40], 5)
    assert (e_ids[:, 0] == np.arange(40)).all()
    r_ids, r_d = ix.range_query(data[:32], 0.5 if metric == "sq_euclid" else 0.05)
print("sanitize case ok")
