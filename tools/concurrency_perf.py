"""hnsw_knn_query from T host threads at once (pageable numpy buffers, C2 index): aggregate QPS per batch size.
Usage: python tools/concurrency_perf.py [n]"""
import json
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np  # noqa: E402
import bench  # noqa: E402
import hnswindex  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
data, q = bench.make_data(N, 40000)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data); ix.set_min_nn(48)
import torch  # noqa: E402
pinned = len(sys.argv) > 2 and sys.argv[2] == "pinned"
qpin = torch.from_numpy(q).pin_memory() if pinned else None
for nb in (1, 16, 256, 2048, 10000):
    row = {"batch": nb, "buffers": "pinned" if pinned else "pageable"}
    for T in (1, 2, 4, 8):
        reps = max(3, min(200, 20000 // nb))
        barrier = threading.Barrier(T + 1)

        def work(t):
            qq = q[(t * nb) % 30000:(t * nb) % 30000 + nb]
            if pinned:
                import ctypes as ct
                f32p, i32p = ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)
                a = qpin[(t * nb) % 30000:(t * nb) % 30000 + nb]
                oi = torch.empty((nb, 10), dtype=torch.int32).pin_memory(); od = torch.empty((nb, 10), dtype=torch.float32).pin_memory()

                def call():
                    assert hnswindex.lib.hnsw_knn_query(ix._h, ct.cast(a.data_ptr(), f32p), nb, 128, 10, ct.cast(oi.data_ptr(), i32p), ct.cast(od.data_ptr(), f32p)) == 0
            else:
                def call():
                    ix.knn_query(qq, 10)
            call()
            barrier.wait()
            for _ in range(reps):
                call()
        th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        for x in th:
            x.start()
        barrier.wait()
        t0 = time.perf_counter()
        for x in th:
            x.join()
        row[f"T{T}_qps"] = round(T * reps * nb / (time.perf_counter() - t0))
    print(json.dumps(row), flush=True)
