"""
World-size-2 gloo tests (CPU) of the partitioned-index host logic (SURVEY.md 8e mode 2):
range split, global ids, all-gather layout and the merge order.  The local searcher is the oracle,
so the expected answer is "the oracle on the same two sub-graphs + the same merge".
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _data():
    rng = np.random.default_rng(17)
    data = rng.standard_normal((3001, 24)).astype(np.float32)
    data /= np.linalg.norm(data, axis=1, keepdims=True)
    q = rng.standard_normal((64, 24)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return data, q


def _worker(rank, world, port, out_dir):
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root)); sys.path.insert(0, str(root / "hnswindex.net_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    import importlib.util
    spec = importlib.util.spec_from_file_location("partitioned", root / "hnswindex.net_b200" / "hnswindex" / "partitioned.py")
    P = importlib.util.module_from_spec(spec); spec.loader.exec_module(P)
    data, q = _data()
    s, e = P.shard_range(len(data), rank, world)
    sub = O.OracleIndex("ucosine", max_candidates=60)
    sub.add(data[s:e])
    sub.set_min_nn(40)
    pix = P.PartitionedIndex(rank, world, len(data), local_search=lambda qq, k: sub.knn_query(qq, k))
    ids, d = pix.knn_query(q, 10)
    np.save(os.path.join(out_dir, f"ids{rank}.npy"), ids); np.save(os.path.join(out_dir, f"d{rank}.npy"), d)
    dist.barrier()
    dist.destroy_process_group()


def test_partitioned_query_world2_gloo(tmp_path, oracle):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    ids0, ids1 = np.load(tmp_path / "ids0.npy"), np.load(tmp_path / "ids1.npy")
    d0, d1 = np.load(tmp_path / "d0.npy"), np.load(tmp_path / "d1.npy")
    assert np.array_equal(ids0, ids1) and np.array_equal(d0.view(np.uint32), d1.view(np.uint32))  # every rank: same merged answer

    # expected: the oracle on the same two sub-graphs, merged in one process
    import importlib.util
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("partitioned", root / "hnswindex.net_b200" / "hnswindex" / "partitioned.py")
    P = importlib.util.module_from_spec(spec); spec.loader.exec_module(P)
    data, q = _data()
    parts_ids, parts_d = [], []
    for r in range(world):
        s, e = P.shard_range(len(data), r, world)
        sub = oracle.OracleIndex("ucosine", max_candidates=60)
        sub.add(data[s:e]); sub.set_min_nn(40)
        i, d = sub.knn_query(q, 10)
        parts_ids.append(np.where(i >= 0, i + s, -1)); parts_d.append(d)
    exp_ids, exp_d = P.merge_topk_numpy(np.stack(parts_ids), np.stack(parts_d), 10)
    assert np.array_equal(ids0, exp_ids) and np.array_equal(d0.view(np.uint32), exp_d.view(np.uint32))
    # rows ascending, ids global and unique, recall vs brute force is high
    assert (np.diff(d0, axis=1) >= 0).all()
    assert ids0.min() >= 0 and ids0.max() < len(data)
    full = oracle.OracleIndex("ucosine", max_candidates=8, max_edges=4); full.add(data)
    gt, _ = full.exact_knn(q, 10)
    rec = np.mean([len(set(ids0[i]) & set(gt[i])) / 10 for i in range(len(q))])
    assert rec > 0.95, rec


def test_shard_range_and_merge_edge_cases():
    import importlib.util
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    spec = importlib.util.spec_from_file_location("partitioned", root / "hnswindex.net_b200" / "hnswindex" / "partitioned.py")
    P = importlib.util.module_from_spec(spec); spec.loader.exec_module(P)
    for n, w in ((10, 3), (8, 8), (3, 8), (1000001, 8)):
        rs = [P.shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
        assert max(e - s for s, e in rs) - min(e - s for s, e in rs) <= 1
    # padding (-1 / NaN) sorts last, ties break by id, negative distances sort first
    ids = np.array([[[5, 9, -1]], [[7, 2, -1]]], dtype=np.int32)
    d = np.array([[[0.5, 0.7, np.nan]], [[-0.1, 0.5, np.nan]]], dtype=np.float32)
    oi, od = P.merge_topk_numpy(ids, d, 5)
    assert oi.tolist() == [[7, 2, 5, 9, -1]]
    assert np.isnan(od[0, 4]) and od[0, 0] == np.float32(-0.1)
