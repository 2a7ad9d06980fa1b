"""Exploration: recall after removing 10 % of an index -- oracle (sequential reference algorithm) vs the GPU engine with
different chunk sizes, all starting from the same oracle-built graph."""
import sys, time, json, tempfile, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np
import bench, hnswindex
from oracle import oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
data, q = bench.make_data(N, 2000)
ref = O.OracleIndex("sq_euclid", max_edges=16, max_candidates=200, collection_size=N)
ref.add(data, threads=1)
path = os.path.join(tempfile.mkdtemp(), "g.hnsw")
ref.serialize(path, packed=True)
rng = np.random.default_rng(1)
victims = rng.permutation(N)[:N // 10].astype(np.int32)
def recall(ix, ef=48):
    ix.set_min_nn(ef)
    return ix
def gpu_truth(g):
    return g.exact_knn(q, 10)[0]
g0 = hnswindex.Index.deserialize("sq_euclid", path); g0.set_min_nn(48)
t0 = gpu_truth(g0)
r0 = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(g0.knn_query(q, 10)[0], t0)])
print(json.dumps({"before": round(float(r0), 4)}), flush=True)
t = time.time(); ref.remove(victims); tr = time.time() - t
ref.set_min_nn(48)
gone = np.zeros(N, bool); gone[victims] = True
for name, tune in (("gpu default chunks", {}), ("gpu ratio 64", {"remove_batch_ratio": 64}), ("gpu ratio 256", {"remove_batch_ratio": 256}), ("gpu one by one", {"remove_batch_max": 1})):
    g = hnswindex.Index.deserialize("sq_euclid", path); g.set_min_nn(48)
    for k, v in tune.items(): g.set_tuning(k, v)
    t = time.time(); g.remove(victims); tg = time.time() - t
    truth = gpu_truth(g)
    rg = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(g.knn_query(q, 10)[0], truth)])
    if name == "gpu default chunks":
        ro = np.mean([len(set(a) & set(b)) / 10 for a, b in zip(ref.knn_query(q, 10, threads=8)[0], truth)])
        print(json.dumps({"oracle sequential": round(float(ro), 4), "seconds": round(tr, 2)}), flush=True)
    deg = (g.export_layer0()[1][~gone] >= 0).sum(1).mean()
    print(json.dumps({name: round(float(rg), 4), "seconds": round(tg, 2), "avg_degree0": round(float(deg), 2)}), flush=True)
