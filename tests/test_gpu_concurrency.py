"""
"Thread-safe per operation type" (the reference's README.md:64-65): hnsw_knn_query from several host threads at once.

hnsw_knn_query holds the index's lock SHARED and runs on its own query context (stream + scratch), so concurrent callers
overlap on the GPU; Add / Remove / Serialize and the accessors take the lock exclusively or shared as needed
(hnsw_ids / hnsw_count / hnsw_export_* used to read containers that Add / Remove reallocate).
"""
import threading

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


def test_concurrent_knn_queries_equal_serial_answers(hnswindex):
    rng = np.random.default_rng(5)
    data = rng.random((20000, 64), dtype=np.float32)
    ix = hnswindex.Index(64, "sq_euclid")
    ix.set_collection_size(20000); ix.set_allow_removals(False); ix.set_min_nn(40)
    ix.add(data)
    jobs = []
    for t in range(8):
        q = rng.random((int(rng.integers(1, 3000)), 64), dtype=np.float32)
        k = int(rng.integers(1, 30))
        jobs.append((q, k, ix.knn_query(q, k)))
    errors, barrier = [], threading.Barrier(len(jobs))

    def worker(q, k, want):
        try:
            barrier.wait()
            for _ in range(20):
                ids, d = ix.knn_query(q, k)
                assert np.array_equal(ids, want[0]) and np.array_equal(d.view(np.uint32), want[1].view(np.uint32))
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))
    threads = [threading.Thread(target=worker, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


def test_queries_accessors_and_updates_from_different_threads(hnswindex):
    rng = np.random.default_rng(6)
    data = rng.random((12000, 32), dtype=np.float32)
    ix = hnswindex.Index(32, "cosine")
    ix.set_collection_size(1000); ix.set_min_nn(16)
    first = ix.add(data[:4000])
    stop, errors = threading.Event(), []

    def querier(seed):
        r = np.random.default_rng(seed)
        try:
            while not stop.is_set():
                q = r.random((int(r.integers(1, 400)), 32), dtype=np.float32)
                n_before = ix.count
                ids, d = ix.knn_query(q, 5)
                assert ids.shape == (len(q), 5) and (ids >= 0).all() and (ids < 12000).all()
                assert (np.diff(d, axis=1) >= 0).all()
                assert n_before >= 3000
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    def reader():
        try:
            while not stop.is_set():
                ids = ix.ids()
                assert len(set(ids.tolist())) == len(ids)
                ix.stats(); ix.params(); ix.entry_point(); ix.top_layer()
                levels, rows, ep = ix.export_layer0()
                assert rows.shape[0] == len(levels) and 0 <= ep < len(levels)
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))
    threads = [threading.Thread(target=querier, args=(s,)) for s in range(4)] + [threading.Thread(target=reader)]
    for t in threads:
        t.start()
    try:
        for s in range(4000, 12000, 1000):        # capacity doubles several times while the queries run
            ix.add(data[s:s + 1000])
        ix.remove(first[:1000])
        ix.add(data[:500])                         # re-uses removed slots
    finally:
        stop.set()
        for t in threads:
            t.join()
    assert not errors, errors[:3]
    assert ix.count == 12000 - 1000 + 500
