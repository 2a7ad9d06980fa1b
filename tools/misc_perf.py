"""Exploration: latency / throughput of the secondary entry points on a 200k x 128 index, next to the oracle."""
import sys, time, json, tempfile, os
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "hnswindex.net_b200"))
import numpy as np
import bench, hnswindex
from oracle import oracle as O
N = 200000
data, q = bench.make_data(N + 2000, 4000)
ix = hnswindex.Index(128, "sq_euclid")
ix.set_collection_size(N); ix.set_max_edges(16); ix.set_max_candidates(200)
ix.add(data[:N])
p = os.path.join(tempfile.mkdtemp(), "m.hnsw"); t = time.perf_counter(); ix.serialize(p, packed=True); ts = time.perf_counter() - t
t = time.perf_counter(); ref = O.OracleIndex.deserialize("sq_euclid", p); to = time.perf_counter() - t
t = time.perf_counter(); ix2 = hnswindex.Index.deserialize("sq_euclid", p); td = time.perf_counter() - t
print(json.dumps({"serialize_s": round(ts, 3), "gpu_deserialize_s": round(td, 3), "oracle_deserialize_s": round(to, 3)}), flush=True)
def tm(fn, reps=3):
    fn(); t = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t) / reps
out = {}
out["single_add_ms_gpu"] = round(tm(lambda: ix.add(data[N:N + 1]), 50) * 1e3, 3)
out["single_add_ms_oracle"] = round(tm(lambda: ref.add(data[N:N + 1]), 50) * 1e3, 3)
out["add_100_ms_gpu"] = round(tm(lambda: ix.add(data[N + 100:N + 200]), 3) * 1e3, 3)
out["multilayer_4000q_ms_gpu"] = round(tm(lambda: ix.multilayer_knn_query(q, 10)) * 1e3, 2)
t = time.perf_counter(); ref.multilayer_knn_query(q[:500], 10); out["multilayer_500q_ms_oracle_1thread"] = round((time.perf_counter() - t) * 1e3, 2)
out["layer1_query_4000q_ms_gpu"] = round(tm(lambda: ix.knn_query_layer(q, 10, 1)) * 1e3, 2)
out["get_info_ms"] = round(tm(lambda: ix.get_info(), 1) * 1e3, 1)
out["components_ms"] = round(tm(lambda: ix.connected_component_counts(), 1) * 1e3, 1)
out["single_remove_ms_gpu"] = round(tm(lambda: None, 1) * 1e3, 3)
ids = list(range(1000, 1050))
t = time.perf_counter()
for i in ids: ix.remove([i])
out["single_remove_ms_gpu"] = round((time.perf_counter() - t) / len(ids) * 1e3, 3)
t = time.perf_counter()
for i in ids: ref.remove([i])
out["single_remove_ms_oracle"] = round((time.perf_counter() - t) / len(ids) * 1e3, 3)
print(json.dumps(out), flush=True)
