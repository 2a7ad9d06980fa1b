"""Exploration: BatchRangeQuery throughput on a C2-shaped index vs the oracle."""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np
import bench, hnswindex
from oracle import oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
data, q = bench.make_data(N, 10000)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N); ix.set_max_edges(16); ix.set_max_candidates(200); ix.set_allow_removals(False)
ix.add(data)
_, d = ix.exact_knn(q[:500], 50)
for name, radius in (("~10 results", float(np.median(d[:, 9]))), ("~50 results", float(np.median(d[:, 49])))):
    ix.range_query(q[:100], radius)
    t = time.perf_counter(); ids, dd = ix.range_query(q, radius); el = time.perf_counter() - t
    sizes = np.array([len(x) for x in ids])
    print(json.dumps({"case": name, "radius": round(radius, 3), "queries": len(q), "seconds": round(el, 4), "qps": round(len(q) / el),
                      "mean_results": round(float(sizes.mean()), 1), "max_results": int(sizes.max())}), flush=True)
import ctypes as ct
n = len(q)
ids_pp, dists_pp, counts = (ct.c_void_p * n)(), (ct.c_void_p * n)(), (ct.c_int * n)()
qq = np.ascontiguousarray(q)
for rep in range(2):
    t = time.perf_counter()
    rc = hnswindex.lib.hnsw_range_query(ix._h, qq.ctypes.data_as(ct.POINTER(ct.c_float)), n, 128, ct.c_float(radius), ids_pp, dists_pp, counts)
    el = time.perf_counter() - t
    hnswindex.lib.hnsw_free_results(ids_pp, dists_pp, n)
print(json.dumps({"c_abi_only_seconds": round(el, 4), "c_abi_qps": round(n / el)}), flush=True)
import tempfile, os
p = os.path.join(tempfile.mkdtemp(), "r.hnsw"); ix.serialize(p, packed=True)
ref = O.OracleIndex.deserialize("sq_euclid", p)
t = time.perf_counter(); oi, od = ref.range_query(q[:2000], radius); el = time.perf_counter() - t
print(json.dumps({"oracle_qps_all_threads": round(2000 / el), "same_sets": bool(all(set(a.tolist()) == set(b.tolist()) for a, b in zip(oi, ids[:2000])))}))
