"""Exploration (torchrun, N ranks): per-rank timing of plain search / all-gather path / fused path on C4-like shards."""
import os, sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np, torch
import torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import hnswindex
from hnswindex import bindings as hb, partitioned as P
from tools.config_perf import lowrank
hb.set_device(local)
n, dim, nq, k = 400000, 96, 10000, 10
data = lowrank(n, dim, 16, 1234 + rank, normalize=True); q = lowrank(nq, dim, 16, 4321, normalize=True)
ix = hnswindex.Index(dim, "ucosine")
ix.set_collection_size(n); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data); ix.set_min_nn(32)
qd = torch.from_numpy(q).cuda()
ids = torch.empty((nq, k), dtype=torch.int32, device="cuda"); dd = torch.empty((nq, k), dtype=torch.float32, device="cuda")
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
out = {"rank": rank}
out["plain_ms"] = round(t(lambda: ix.knn_query_device(qd.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr())), 3)
out["plain_kernel_ms"] = round(ix.stats()["last_search_ms"], 3)
pix = P.PartitionedIndex(rank, world, n * world, index=ix)
a = torch.zeros((nq, k), dtype=torch.int32, device="cuda")
def ag():
    lst = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(lst, a)
out["allgather_only_ms"] = round(t(ag), 3)
big = torch.empty((world, nq, k), dtype=torch.int32, device="cuda")
out["allgather_into_tensor_ms"] = round(t(lambda: dist.all_gather_into_tensor(big, a)), 3)
out["gather_path_ms"] = round(t(lambda: pix.knn_query_device(qd, k)), 3)
pix.enable_fused(nq, k)
out["fused_ms"] = round(t(lambda: pix.knn_query_device(qd, k)), 3)
print(json.dumps(out), flush=True)
dist.barrier(); dist.destroy_process_group()
