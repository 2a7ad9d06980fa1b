"""GPU tests of K7 (top-k merge) and of the partitioned-index path on one device (world = 1 and emulated parts)."""
import ctypes as ct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hnswindex():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import __graft_entry__ as ge
    ge.build()
    import hnswindex
    return hnswindex


@pytest.mark.parametrize("parts,nq,k", [(2, 37, 10), (8, 200, 10), (8, 64, 100), (3, 5, 1)])
def test_merge_kernel_matches_numpy_statement(hnswindex, parts, nq, k):
    import torch
    from hnswindex import partitioned as P
    rng = np.random.default_rng(parts * 1000 + k)
    d = np.sort(rng.standard_normal((parts, nq, k)).astype(np.float32), axis=2)  # cosine-like: negatives allowed
    ids = rng.permutation(parts * nq * k).astype(np.int32).reshape(parts, nq, k)
    # some short lists: padded with -1 / NaN at the end
    for p in range(parts):
        cut = rng.integers(0, k + 1, size=nq)
        for qi in range(nq):
            ids[p, qi, cut[qi]:] = -1
            d[p, qi, cut[qi]:] = np.nan
    # a few exact ties across parts
    d[1 % parts, :, 0] = d[0, :, 0]
    exp_ids, exp_d = P.merge_topk_numpy(ids, d, k)
    t_ids, t_d = torch.from_numpy(ids).cuda(), torch.from_numpy(d).cuda()
    o_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    o_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    rc = hnswindex.lib.hnsw_merge_topk_device(ct.c_void_p(t_ids.data_ptr()), ct.c_void_p(t_d.data_ptr()), parts, nq, k,
                                              ct.c_void_p(o_ids.data_ptr()), ct.c_void_p(o_d.data_ptr()), None)
    assert rc == 0, hnswindex.last_error()
    assert np.array_equal(o_ids.cpu().numpy(), exp_ids)
    got = o_d.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(exp_d))
    assert np.array_equal(got[~np.isnan(got)].view(np.uint32), exp_d[~np.isnan(exp_d)].view(np.uint32))


def test_partitioned_emulated_on_one_gpu_matches_oracle_partitions(hnswindex, oracle, tmp_path):
    # two sub-graphs on ONE device (ranks emulated sequentially), merged by the kernel; expected = the oracle on the
    # same sub-graphs + the numpy statement of the merge
    import torch
    from hnswindex import partitioned as P
    rng = np.random.default_rng(23)
    data = rng.standard_normal((4000, 96)).astype(np.float32)
    data /= np.linalg.norm(data, axis=1, keepdims=True)
    q = rng.standard_normal((128, 96)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    world, k = 2, 10
    g_parts, o_parts = [], []
    for r in range(world):
        s, e = P.shard_range(len(data), r, world)
        sub = oracle.OracleIndex("ucosine", max_candidates=100)
        sub.add(data[s:e])
        sub.set_min_nn(32)
        path = tmp_path / f"part{r}.hnsw"
        sub.serialize(path, packed=True)
        gpu = hnswindex.Index.deserialize("ucosine", path)
        gpu.set_min_nn(32)
        gi, gd = gpu.knn_query(q, k)
        oi, od = sub.knn_query(q, k)
        g_parts.append((np.where(gi >= 0, gi + s, -1).astype(np.int32), gd))
        o_parts.append((np.where(oi >= 0, oi + s, -1).astype(np.int32), od))
    exp_ids, exp_d = P.merge_topk_numpy(np.stack([p[0] for p in o_parts]), np.stack([p[1] for p in o_parts]), k)
    t_ids = torch.from_numpy(np.stack([p[0] for p in g_parts])).cuda()
    t_d = torch.from_numpy(np.stack([p[1] for p in g_parts])).cuda()
    o_ids = torch.empty((len(q), k), dtype=torch.int32, device="cuda")
    o_d = torch.empty((len(q), k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    assert hnswindex.lib.hnsw_merge_topk_device(ct.c_void_p(t_ids.data_ptr()), ct.c_void_p(t_d.data_ptr()), world, len(q), k,
                                                ct.c_void_p(o_ids.data_ptr()), ct.c_void_p(o_d.data_ptr()), None) == 0
    same = (o_ids.cpu().numpy() == exp_ids).all(axis=1).mean()
    assert same >= 0.995, same
    assert np.array_equal(o_d.cpu().numpy().view(np.uint32), exp_d.view(np.uint32))


def test_partitioned_index_world1_cuda_path(hnswindex):
    from hnswindex import partitioned as P
    rng = np.random.default_rng(5)
    data = rng.random((3000, 32), dtype=np.float32)
    ix = hnswindex.Index(32, "sq_euclid")
    ix.set_collection_size(3000)
    ix.add(data)
    pix = P.PartitionedIndex(0, 1, 3000, index=ix)
    a = pix.knn_query(data[:100], 5)
    b = ix.knn_query(data[:100], 5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_partitioned_fused_world1_matches_plain_query(hnswindex):
    # the fused search + scatter path (CUDA IPC buffer, flags, wait + merge) with a single rank: same answers as knn_query
    from hnswindex import partitioned as P
    rng = np.random.default_rng(6)
    data = rng.random((5000, 48), dtype=np.float32)
    ix = hnswindex.Index(48, "sq_euclid")
    ix.set_collection_size(5000)
    ix.add(data)
    ix.set_min_nn(40)
    pix = P.PartitionedIndex(0, 1, 5000, index=ix)
    pix.enable_fused(600, 7)
    for n in (600, 1, 333, 600, 17):
        a = pix.knn_query(data[:n], 7)
        b = ix.knn_query(data[:n], 7)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    with pytest.raises(ValueError):
        pix.knn_query(data[:601], 7)
    pix.disable_fused()
    a = pix.knn_query(data[:50], 7)
    b = ix.knn_query(data[:50], 7)
    assert np.array_equal(a[0], b[0])


def test_partitioned_fused_two_gpus_matches_all_gather_path(hnswindex):
    import subprocess
    import sys
    from pathlib import Path
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(root / "tools" / "partition_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "partition_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
