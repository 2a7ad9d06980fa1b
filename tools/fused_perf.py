"""Exploration: one C4 shard on one GPU -- plain device-resident search vs the fused search+scatter+wait+merge path (world = 1)."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np, torch
import hnswindex
from hnswindex import partitioned as P
from tools.config_perf import lowrank  # noqa

n, dim, nq, k = 1250000, 96, 10000, 10
data = lowrank(n, dim, 16, 1234, normalize=True); q = lowrank(nq, dim, 16, 4321, normalize=True)
ix = hnswindex.Index(dim, "ucosine")
ix.set_collection_size(n); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data); ix.set_min_nn(32)
qd = torch.from_numpy(q).cuda()
ids = torch.empty((nq, k), dtype=torch.int32, device="cuda"); dd = torch.empty((nq, k), dtype=torch.float32, device="cuda")
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
plain = t(lambda: ix.knn_query_device(qd.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr()))
kms = ix.stats()["last_search_ms"]
pix = P.PartitionedIndex(0, 1, n, index=ix)
gather = t(lambda: pix.knn_query_device(qd, k))
pix.enable_fused(nq, k)
fused = t(lambda: pix.knn_query_device(qd, k))
print(json.dumps({"plain_wall_ms": round(plain, 3), "plain_kernel_ms": round(kms, 3), "gather_path_ms": round(gather, 3), "fused_ms": round(fused, 3)}))
