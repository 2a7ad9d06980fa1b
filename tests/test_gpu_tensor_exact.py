"""
GPU tests of the tensor-core exact top-k (hnsw_exact_knn_tf32, csrc/exact_tf32_kernels.cu): a 3xTF32 GEMM ranks all
pairs approximately, k + slack candidates per query are evaluated with the bit-exact metric, uncertified queries go
through the FFMA scan -- so the answers must equal hnsw_exact_knn's to the bit (which test_gpu_parity.py pins on the
oracle's brute force).  The GEMM exists three times: tcgen05.mma + TMEM for rows of at most 128 floats -- warp-specialised
(4 producer + 4 consumer warps, tuning key "exact_tcgen05" = 2, the default) and the earlier 4-warp kernel (= 1) -- and
mma.sync (every other shape, or "exact_tcgen05" = 0); all three are covered.
"""
import time

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


def _index(hnswindex, metric, data, **kw):
    ix = hnswindex.Index(data.shape[1], metric)
    ix.set_max_edges(kw.get("max_edges", 4))
    ix.set_max_candidates(kw.get("max_candidates", 8))
    ix.set_collection_size(max(1, len(data)))
    ix.add(data)
    return ix


def _same(ix, q, k):
    """scan == tensor-core variant for both GEMM kernels; returns the larger number of fall-backs"""
    a_ids, a_d = ix.exact_knn(q, k)
    fb = 0
    for tcgen05 in (2, 1, 0):
        ix.set_tuning("exact_tcgen05", tcgen05)
        b_ids, b_d = ix.exact_knn(q, k, tensor_cores=True)
        assert np.array_equal(a_ids, b_ids), f"ids differ in {(a_ids != b_ids).any(axis=1).sum()} of {len(q)} rows (tcgen05={tcgen05})"
        assert np.array_equal(a_d.view(np.uint32), b_d.view(np.uint32)), f"distances are not bit-identical (tcgen05={tcgen05})"
        fb = max(fb, ix.last_exact_fallbacks)
    ix.set_tuning("exact_tcgen05", 2)
    return fb


def test_tcgen05_kernel_covers_the_headline_shape(hnswindex):
    from hnswindex.bindings import lib
    assert lib.hnsw_exact_tcgen05_supported(128, 10) == 1   # C2: 1M x 128, k = 10
    assert lib.hnsw_exact_tcgen05_supported(96, 10) == 1    # C4
    assert lib.hnsw_exact_tcgen05_supported(960, 100) == 0  # C3 rows are too long for one block's shared memory: mma.sync


@pytest.mark.parametrize("metric,dim", [("sq_euclid", 128), ("cosine", 100), ("ucosine", 96), ("euclid", 64), ("sq_euclid", 3),
                                        ("sq_euclid", 200), ("cosine", 960)])
def test_tensor_core_exact_equals_scan(hnswindex, metric, dim):
    rng = np.random.default_rng(11)
    n = 5000 if dim < 900 else 3000
    data = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((130, dim)).astype(np.float32)  # not a multiple of the 64 / 32 / 16 query tile
    if metric == "ucosine":
        data /= np.linalg.norm(data, axis=1, keepdims=True)
        q /= np.linalg.norm(q, axis=1, keepdims=True)
    ix = _index(hnswindex, metric, data)
    for k in (1, 10, 100):
        fb = _same(ix, q, k)
        if dim > 3:  # (in 3 dimensions near-ties are dense; any number of fall-backs is legitimate)
            assert fb <= len(q) // 4, f"{fb} of {len(q)} queries fell back to the scan ({metric}, dim {dim}, k {k})"


def test_tensor_core_exact_skips_removed_items_and_pads_small_indexes(hnswindex):
    rng = np.random.default_rng(12)
    data = rng.standard_normal((3000, 64)).astype(np.float32)
    q = data[:80] + 0.01 * rng.standard_normal((80, 64)).astype(np.float32)
    ix = _index(hnswindex, "sq_euclid", data)
    ix.remove(np.arange(0, 3000, 3, dtype=np.int32))  # every third item, incl. the nearest neighbours of a third of the queries
    _same(ix, q, 10)
    ids, _ = ix.exact_knn(q, 10, tensor_cores=True)
    assert (ids % 3 != 0).all()
    tiny = _index(hnswindex, "cosine", data[:5])
    ids, d = tiny.exact_knn(q[:7], 10, tensor_cores=True)
    assert (ids[:, 5:] == -1).all() and np.isnan(d[:, 5:]).all() and (ids[:, :5] >= 0).all()
    _same(tiny, q[:7], 10)


def test_tensor_core_exact_on_clustered_data_200k(hnswindex):
    # the bench's data recipe: low intrinsic dimension, large norms against small neighbour distances --
    # the hard case for the certificate
    rng = np.random.default_rng(1234)
    A = rng.standard_normal((16, 128)).astype(np.float32)
    z = rng.standard_normal((202000, 16)).astype(np.float32)
    x = (z @ A + 0.25 * rng.standard_normal((202000, 128)).astype(np.float32)).astype(np.float32)
    data, q = x[:200000], x[200000:]
    ix = _index(hnswindex, "sq_euclid", data, max_edges=8, max_candidates=32)
    fb = _same(ix, q, 10)
    t0 = time.time()
    ix.exact_knn(q, 10)
    t1 = time.time()
    ix.exact_knn(q, 10, tensor_cores=True)
    t2 = time.time()
    print(f"\nexact top-10, 2000 queries x 200k x 128: scan {t1 - t0:.3f} s, tensor cores {t2 - t1:.3f} s, fall-backs {fb}")
    assert fb <= 200
