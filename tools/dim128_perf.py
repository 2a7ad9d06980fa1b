"""Default routing on the C2 data (1M x 128) per search width with and without the constant-trip-count instantiations (Tuning dim128):
QPS of device-resident batches by CUDA-event kernel time, answers compared.  Usage: python tools/dim128_perf.py [n] [nq]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import torch  # noqa: E402
import hnswindex  # noqa: E402
from bench import make_data  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
data, queries = make_data(n, nq)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(n); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
t = time.time(); ix.add(data); print(f"build {time.time() - t:.1f}s kernel {ix.stats()['last_build_ms']:.1f} ms", flush=True)
q_dev = torch.from_numpy(queries).cuda()
ids = torch.empty((nq, 10), dtype=torch.int32, device="cuda"); d = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
for ef in (48, 96, 128, 192, 256, 320, 384, 512, 768, 1024, 2048, 4096):
    m = nq if ef <= 1024 else 3000
    ix.set_min_nn(ef)
    row = {"ef": ef, "queries": m}
    got = {}
    for rep in range(2):
        for flag in (0, 1):
            ix.set_tuning("dim128", flag)
            best = 1e9
            for _ in range(3):
                ix.knn_query_device(q_dev.data_ptr(), m, 10, ids.data_ptr(), d.data_ptr())
                best = min(best, ix.stats()["last_search_ms"])
            got[flag] = (ids[:m].cpu().numpy().copy(), d[:m].cpu().numpy().copy())
            row[f"qps_dim128_{flag}"] = max(row.get(f"qps_dim128_{flag}", 0), round(m / (best / 1e3)))
    row["same"] = bool(np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1].view(np.uint32), got[1][1].view(np.uint32)))
    row["gain"] = round(row["qps_dim128_1"] / row["qps_dim128_0"], 3)
    print(json.dumps(row), flush=True)
