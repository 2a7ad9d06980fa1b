"""
CPU tests of the drop-in boundary: the shared library loads, exports every symbol that
include/hnsw_b200.h declares (and every symbol the reference's bindings.py binds), fails loudly
without a GPU, and its snapshot codec round-trips files written by the oracle.  No compute calls.
"""
import ctypes as ct
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

# the 16 + 1 exports of bindings/HNSWIndex.Native/HNSWIndexExports.cs:27-273
REFERENCE_EXPORTS = [
    "hnsw_get_last_error_utf8", "hnsw_create", "hnsw_free", "hnsw_add", "hnsw_remove", "hnsw_knn_query",
    "hnsw_range_query", "hnsw_free_results", "hnsw_set_collection_size", "hnsw_set_max_edges",
    "hnsw_set_max_candidates", "hnsw_set_remove_max_candidates", "hnsw_set_distribution_rate",
    "hnsw_set_random_seed", "hnsw_set_min_nn", "hnsw_set_allow_removals",
]


@pytest.fixture(scope="module")
def hnswindex():
    import __graft_entry__ as ge
    ge.build()
    import hnswindex
    return hnswindex


def test_library_at_reference_location(hnswindex):
    # bindings.py:35-41: <package>/artifacts/native/<rid>/HNSWIndex.Native.so
    assert hnswindex.LIB_PATH.name == "HNSWIndex.Native.so"
    assert hnswindex.LIB_PATH.parent.name == "linux-x64"
    assert hnswindex.LIB_PATH.exists()


def test_exports_every_declared_symbol(hnswindex):
    header = (ROOT / "include" / "hnsw_b200.h").read_text()
    declared = sorted(set(re.findall(r"HNSW_API\s+[\w\s\*]+?\b(hnsw_\w+)\s*\(", header)))
    assert len(declared) >= 30
    L = ct.CDLL(str(hnswindex.LIB_PATH))
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    for s in REFERENCE_EXPORTS:
        assert s in declared, s


def test_no_torch_or_cxx_types_in_abi():
    header = (ROOT / "include" / "hnsw_b200.h").read_text()
    code = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # declarations only, comments stripped
    assert "torch" not in code.lower() and "std::" not in code and "at::" not in code and "#include <c" not in code


def test_unknown_metric_sets_last_error(hnswindex):
    # HNSWIndexExports.cs:58-59,64: unknown metric -> null handle + last error
    h = hnswindex.lib.hnsw_create(b"manhattan")
    assert not h
    assert "Unsupported distance metric" in hnswindex.last_error()


def test_null_handle_conventions(hnswindex):
    L = hnswindex.lib
    # HNSWIndexExports.cs:78,105,122: handle == 0 -> return 0 without touching the outputs
    ids = np.full(4, 77, dtype=np.int32)
    d = np.full(4, 5.0, dtype=np.float32)
    q = np.zeros((2, 8), dtype=np.float32)
    assert L.hnsw_knn_query(None, q.ctypes.data_as(ct.POINTER(ct.c_float)), 2, 8, 2, ids.ctypes.data_as(ct.POINTER(ct.c_int)),
                            d.ctypes.data_as(ct.POINTER(ct.c_float))) == 0
    assert (ids == 77).all() and (d == 5.0).all()
    assert L.hnsw_add(None, q.ctypes.data_as(ct.POINTER(ct.c_float)), 2, 8, ids.ctypes.data_as(ct.POINTER(ct.c_int))) == 0
    assert L.hnsw_remove(None, ids.ctypes.data_as(ct.POINTER(ct.c_int)), 2) == 0
    L.hnsw_free(None)
    assert L.hnsw_count(None) == -1


def test_last_error_buffer_protocol(hnswindex):
    L = hnswindex.lib
    assert not L.hnsw_create(b"nope")
    need = L.hnsw_get_last_error_utf8(None, 0)
    assert need > 0
    small = ct.create_string_buffer(8)
    assert L.hnsw_get_last_error_utf8(small, 8) == need  # returns the needed length, writes bufLen-1 bytes + NUL
    assert len(small.value) == 7
    full = ct.create_string_buffer(need + 1)
    L.hnsw_get_last_error_utf8(full, need + 1)
    assert len(full.value) == need


def test_fails_loudly_without_gpu(hnswindex):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ix = hnswindex.Index(8)
    with pytest.raises(RuntimeError, match="hnsw_create failed"):
        ix.add(np.zeros((2, 8), dtype=np.float32))  # no CPU fallback


@pytest.mark.parametrize("packed_in,packed_out", [(False, False), (False, True), (True, False), (True, True)])
def test_codec_roundtrip_against_oracle(hnswindex, oracle, tmp_path, packed_in, packed_out):
    # oracle-written snapshot -> product codec (read + write) -> oracle reads it back: same answers
    v = oracle.random_vectors(24, 600)
    ix = oracle.OracleIndex("cosine", max_edges=6, max_candidates=40, random_seed=11, collection_size=100)
    ix.add(v)
    a, b = tmp_path / "a.bin", tmp_path / "b.bin"
    ix.serialize(a, packed=packed_in)
    n = hnswindex.lib.hnsw_snapshot_transcode(str(a).encode(), str(b).encode(), int(packed_out))
    assert n == 600, hnswindex.last_error()
    ix2 = oracle.OracleIndex.deserialize("cosine", b)
    assert ix2.params() == ix.params()
    assert ix2.entry_point == ix.entry_point and ix2.capacity == ix.capacity and ix2.count == 600
    assert np.array_equal(ix.items(), ix2.items())
    for i in range(0, 600, 37):
        for layer in range(ix.node_level(i) + 1):
            assert np.array_equal(ix.row(i, layer), ix2.row(i, layer))
            assert sorted(ix.row(i, layer, 1)) == sorted(ix2.row(i, layer, 1))  # in-edges = transpose, order free
    q = oracle.random_vectors(24, 50, seed=5)
    ia, da = ix.knn_query(q, 7)
    ib, db = ix2.knn_query(q, 7)
    assert np.array_equal(ia, ib) and np.array_equal(da.view(np.uint32), db.view(np.uint32))


@pytest.mark.parametrize("packed", [False, True])
def test_codec_keeps_removal_state(hnswindex, oracle, tmp_path, packed):
    # a snapshot with removed nodes (IsRemoved, ActiveNodes in ActiveSet order, RemovedIndexes) survives the product codec
    v = oracle.random_vectors(16, 400)
    ix = oracle.OracleIndex("sq_euclid", max_edges=6)
    ids = ix.add(v)
    ix.remove(ids[5:120:3])
    a, b = tmp_path / "a.bin", tmp_path / "b.bin"
    ix.serialize(a, packed=packed)
    assert hnswindex.lib.hnsw_snapshot_transcode(str(a).encode(), str(b).encode(), int(packed)) == 400, hnswindex.last_error()
    ix2 = oracle.OracleIndex.deserialize("sq_euclid", b)
    assert ix2.count == ix.count == 400 - len(ids[5:120:3])
    assert np.array_equal(ix2.active_ids(), ix.active_ids())
    assert [ix2.is_removed(i) for i in range(400)] == [ix.is_removed(i) for i in range(400)]
    assert ix2.entry_point == ix.entry_point
    q = oracle.random_vectors(16, 40, seed=3)
    ia, da = ix.knn_query(q, 6)
    ib, db = ix2.knn_query(q, 6)
    assert np.array_equal(ia, ib) and np.array_equal(da.view(np.uint32), db.view(np.uint32))
    # the removed-slot stack comes back in the same order: both hand out the same ids next
    assert ix2.add(v[:4]).tolist() == ix.add(v[:4]).tolist()


def test_codec_byte_identical_to_oracle_writer_when_no_stale_slots(hnswindex, oracle, tmp_path):
    # Two independent writers of the same layout must agree byte for byte once the oracle's stale
    # Buffer slots are removed (AllowRemovals=false, transcode of a transcode is a fixed point).
    v = oracle.random_vectors(16, 300)
    ix = oracle.OracleIndex("sq_euclid", allow_removals=False)
    ix.add(v)
    a, b, c = tmp_path / "a.bin", tmp_path / "b.bin", tmp_path / "c.bin"
    ix.serialize(a)
    assert hnswindex.lib.hnsw_snapshot_transcode(str(a).encode(), str(b).encode(), 0) == 300
    assert hnswindex.lib.hnsw_snapshot_transcode(str(b).encode(), str(c).encode(), 0) == 300
    assert b.read_bytes() == c.read_bytes()
    ix2 = oracle.OracleIndex.deserialize("sq_euclid", b)
    d = tmp_path / "d.bin"
    ix2.serialize(d)   # oracle re-writes what it read: Buffers now hold exactly Count entries
    assert d.read_bytes() == b.read_bytes()


def test_codec_rejects_garbage(hnswindex, tmp_path):
    p = tmp_path / "bad.bin"
    p.write_bytes(b"\x0a\xff\xff\xff\xff\x0f garbage")
    assert hnswindex.lib.hnsw_snapshot_transcode(str(p).encode(), str(tmp_path / "o.bin").encode(), 0) == -1
    assert "snapshot" in hnswindex.last_error()
    assert hnswindex.lib.hnsw_snapshot_transcode(str(tmp_path / "missing").encode(), str(tmp_path / "o.bin").encode(), 0) == -1


def test_pending_parameter_setters_return_zero(hnswindex):
    L = hnswindex.lib
    assert L.hnsw_set_max_edges(16) == 0 and L.hnsw_set_min_nn(5) == 0 and L.hnsw_set_max_candidates(100) == 0
    assert L.hnsw_set_distribution_rate(0.36067376) == 0 and L.hnsw_set_allow_removals(True) == 0
    assert L.hnsw_set_collection_size(65536) == 0 and L.hnsw_set_random_seed(31337) == 0
    assert L.hnsw_set_remove_max_candidates(100) == 0
