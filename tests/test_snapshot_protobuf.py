"""
The snapshot bytes, pinned by a parser the builder did not write.

google.protobuf (stock implementation in this image) parses files produced by the PRODUCT writer (csrc/snapshot.cpp,
reached on the CPU through hnsw_snapshot_transcode, on the GPU through hnsw_serialize) using the five protobuf-net
contracts of the reference declared as a schema (tests/pb_schema.py).  Checked:
  * every field the reference's classes hold comes out with the value the index has;
  * protobuf-net's conventions: zero-valued members are NOT written, repeated scalars are unpacked by default, a
    ConcurrentStack is written top first, EdgeList.Count sits beside a Buffer that may be longer (Node.cs:33-36);
  * re-serialising the parsed message with google.protobuf gives the product's bytes back, byte for byte, for both the
    unpacked (default) and the packed form -- so field order, varint widths (negative int32 = 10 bytes), fixed32 floats and
    length prefixes are exactly what a conforming protobuf writer emits;
  * a file WRITTEN by google.protobuf from scratch (with stale Buffer slots beyond Count, as the reference's arrays have)
    is read by the product and by the oracle.
"""
import ctypes as ct
import math
import struct
from pathlib import Path

import numpy as np
import pytest

import pb_schema

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def hnswindex():
    import __graft_entry__ as ge
    ge.build()
    import hnswindex
    return hnswindex


@pytest.fixture(scope="module")
def M():
    return pb_schema.build_messages(packed=False)


@pytest.fixture(scope="module")
def MP():
    return pb_schema.build_messages(packed=True)


def transcode(hnswindex, src, dst, packed):
    n = hnswindex.lib.hnsw_snapshot_transcode(str(src).encode(), str(dst).encode(), int(packed))
    assert n >= 0, hnswindex.last_error()
    return n


def check_against_oracle(snap, ix, removed_expected=()):
    """Every member of the parsed snapshot against the oracle index the file came from."""
    P, D = snap.Parameters, snap.DataSnapshot
    want = ix.params()
    assert (P.MaxEdges, P.MinNN, P.MaxCandidates, P.RemoveMaxCandidates, P.CollectionSize, P.RandomSeed, P.AllowRemovals) == \
        (want["max_edges"], want["min_nn"], want["max_candidates"], want["remove_max_candidates"], want["collection_size"],
         want["random_seed"], want["allow_removals"])
    assert P.DistributionRate == want["distribution_rate"]
    n = ix.length
    assert D.Length == n and D.Capacity == ix.capacity and D.Count == ix.count and len(D.Nodes) == n and len(D.Items) == n
    assert D.EntryPointId == ix.entry_point
    assert list(D.ActiveNodes) == list(ix.active_ids())
    items = ix.items()
    for i, node in enumerate(D.Nodes):
        assert node.Id == i
        assert node.IsRemoved == bool(ix.is_removed(i))
        lv = ix.node_level(i)
        assert len(node.OutEdges) == lv + 1
        assert len(node.InEdges) == (lv + 1 if want["allow_removals"] else 0)
        for layer in range(lv + 1):
            e = node.OutEdges[layer]
            assert e.Count <= len(e.Buffer)                                          # Buffer at capacity, Count beside it
            if ix.is_removed(i):
                # The reference leaves a removed node's own OutEdges behind (GraphConnector.cs:277-288 only detaches the
                # neighbours' InEdges); nothing reads them again before the slot is re-used (GraphData.cs:84-91 NewNode).
                # The product keeps removed rows empty, which is the same graph for every later operation.
                assert e.Count == 0
                continue
            assert list(e.Buffer)[:e.Count] == list(ix.row(i, layer))
            if want["allow_removals"]:
                ie = node.InEdges[layer]
                assert sorted(list(ie.Buffer)[:ie.Count]) == sorted(ix.row(i, layer, 1))  # in-edges: a set, order free
        assert np.array_equal(np.array(D.Items[i].Values, dtype=np.float32).view(np.uint32), items[i].view(np.uint32))
    removed = [i for i in range(n) if ix.is_removed(i)]
    assert sorted(D.RemovedIndexes) == removed
    if removed_expected:
        assert list(D.RemovedIndexes) == list(removed_expected)


def oracle_params_dict(ix):
    p = ix.params()
    return p if isinstance(p, dict) else p._asdict()


@pytest.mark.parametrize("allow_removals", [True, False])
@pytest.mark.parametrize("packed", [False, True])
def test_product_writer_parsed_by_google_protobuf(hnswindex, oracle, M, MP, tmp_path, packed, allow_removals):
    v = oracle.random_vectors(20, 500)
    ix = oracle.OracleIndex("cosine", max_edges=5, max_candidates=30, random_seed=-7 if not allow_removals else 4242,
                            collection_size=64, allow_removals=allow_removals, min_nn=9)
    if not allow_removals:
        ix = oracle.OracleIndex("cosine", max_edges=5, max_candidates=30, random_seed=4243, collection_size=64,
                                allow_removals=False, min_nn=9)
    ix.add(v)
    a, b = tmp_path / "oracle.bin", tmp_path / "product.bin"
    ix.serialize(a)
    assert transcode(hnswindex, a, b, packed) == 500
    raw = b.read_bytes()
    snap = (MP if packed else M)["HNSWIndexSnapshot"]()
    snap.ParseFromString(raw)
    check_against_oracle(snap, ix)
    # google.protobuf writes the same bytes from the parsed object: the product's file IS canonical protobuf
    assert snap.SerializeToString() == raw
    # and the other schema (packed <-> unpacked) reads it too, giving the same object
    other = (M if packed else MP)["HNSWIndexSnapshot"]()
    other.ParseFromString(raw)
    again = (MP if packed else M)["HNSWIndexSnapshot"]()
    again.ParseFromString(other.SerializeToString())
    assert again.SerializeToString() == raw


def test_wire_conventions_of_protobuf_net(hnswindex, oracle, tmp_path):
    """What the object model hides: zero members absent, unpacked scalars, field order, negative varints, root without prefix."""
    v = oracle.random_vectors(8, 40)
    ix = oracle.OracleIndex("sq_euclid", max_edges=4, random_seed=-5, allow_removals=False, collection_size=40)
    ix.add(v)
    a, b = tmp_path / "a.bin", tmp_path / "b.bin"
    ix.serialize(a)
    transcode(hnswindex, a, b, 0)
    root = pb_schema.scan_fields(b.read_bytes())
    assert [(f, w) for f, w, _ in root] == [(1, 2), (2, 2)]                       # HNSWIndexSnapshot.cs:12-16, no length prefix
    params = pb_schema.scan_fields(root[0][2])
    assert [f for f, _, _ in params] == [1, 2, 3, 4, 5, 6, 7]                     # AllowRemovals = false is omitted (field 8)
    seed = [x for x in params if x[0] == 7][0]
    assert seed[2] == (1 << 64) - 5                                               # int32 -5 as a sign-extended 10-byte varint
    assert struct.unpack("<d", [x for x in params if x[0] == 2][0][2])[0] == 1.0 / math.log(16.0)
    data = pb_schema.scan_fields(root[1][2])
    nums = [f for f, _, _ in data]
    assert nums == sorted(nums)                                                   # members in field order
    assert nums.count(1) == 40 and nums.count(3) == 40 and nums.count(2) == 40    # ActiveNodes unpacked: one varint per id
    assert all(w == 0 for f, w, _ in data if f == 2)
    assert 4 not in nums                                                          # empty RemovedIndexes: nothing written
    ep = ix.entry_point
    assert ([x[2] for x in data if x[0] == 5] == [ep]) if ep != 0 else (5 not in nums)
    node0 = pb_schema.scan_fields([x for x in data if x[0] == 1][0][2])
    assert 1 not in [f for f, _, _ in node0]                                      # Node.Id = 0 omitted
    assert 2 not in [f for f, _, _ in node0] and 4 not in [f for f, _, _ in node0]  # IsRemoved false, no InEdges
    node7 = pb_schema.scan_fields([x for x in data if x[0] == 1][7][2])
    assert node7[0] == (1, 0, 7)
    edge = pb_schema.scan_fields([x for x in node7 if x[0] == 3][0][2])
    assert all(w == 0 for f, w, _ in edge) and [f for f, _, _ in edge][-1] == 2   # Buffer ints unpacked, then Count
    item = pb_schema.scan_fields([x for x in data if x[0] == 3][5][2])
    assert [(f, w) for f, w, _ in item] == [(1, 5)] * 8                           # floats unpacked fixed32
    assert np.array_equal(np.frombuffer(b"".join(x[2] for x in item), dtype="<f4"), ix.items()[5])
    # packed form: one length-delimited run per repeated scalar
    transcode(hnswindex, a, tmp_path / "p.bin", 1)
    rootp = pb_schema.scan_fields((tmp_path / "p.bin").read_bytes())
    datap = pb_schema.scan_fields(rootp[1][2])
    assert [w for f, w, _ in datap if f == 2] == [2]
    itemp = pb_schema.scan_fields([x for x in datap if x[0] == 3][5][2])
    assert [(f, w, len(p)) for f, w, p in itemp] == [(1, 2, 32)]


def test_removal_state_bytes(hnswindex, oracle, M, tmp_path):
    v = oracle.random_vectors(16, 300)
    ix = oracle.OracleIndex("sq_euclid", max_edges=6)
    ids = ix.add(v)
    gone = [int(i) for i in ids[10:100:7]]
    ix.remove(gone)
    a, b = tmp_path / "a.bin", tmp_path / "b.bin"
    ix.serialize(a)
    assert transcode(hnswindex, a, b, 0) == 300
    raw = b.read_bytes()
    snap = M["HNSWIndexSnapshot"]()
    snap.ParseFromString(raw)
    # ConcurrentStack<int> enumerates from its top: the id removed LAST is written first and handed out first (GraphData.cs:84-91)
    check_against_oracle(snap, ix, removed_expected=list(reversed(gone)))
    assert snap.SerializeToString() == raw
    assert ix.add(v[:1])[0] == gone[-1]


def test_file_written_by_google_protobuf_is_read_by_product_and_oracle(hnswindex, oracle, M, tmp_path):
    """The other direction: a snapshot authored by a third party (with stale Buffer slots beyond Count, like the reference's
    capacity-sized arrays) loads in both readers, and what the product writes back parses to the same graph."""
    v = oracle.random_vectors(12, 200)
    ix = oracle.OracleIndex("ucosine", max_edges=4, max_candidates=20, collection_size=256)
    ix.add(oracle.normalize_rows_like_reference(v))
    a = tmp_path / "a.bin"
    ix.serialize(a)
    snap = M["HNSWIndexSnapshot"]()
    snap.ParseFromString(a.read_bytes())
    rng = np.random.default_rng(5)
    for node in snap.DataSnapshot.Nodes:                  # pad every Buffer to a reference-like capacity with garbage ids
        for e in list(node.OutEdges) + list(node.InEdges):
            cnt = e.Count
            del e.Buffer[cnt:]
            e.Buffer.extend(int(x) for x in rng.integers(0, 200, size=rng.integers(0, 4)))
    third = tmp_path / "third_party.bin"
    third.write_bytes(snap.SerializeToString())
    ix2 = oracle.OracleIndex.deserialize("ucosine", third)
    back = tmp_path / "back.bin"
    assert transcode(hnswindex, third, back, 0) == 200
    ix3 = oracle.OracleIndex.deserialize("ucosine", back)
    q = oracle.normalize_rows_like_reference(oracle.random_vectors(12, 40, seed=9))
    r1, r2, r3 = ix.knn_query(q, 5), ix2.knn_query(q, 5), ix3.knn_query(q, 5)
    assert np.array_equal(r1[0], r2[0]) and np.array_equal(r1[0], r3[0])
    assert np.array_equal(r1[1].view(np.uint32), r3[1].view(np.uint32))
    snap3 = M["HNSWIndexSnapshot"]()
    snap3.ParseFromString(back.read_bytes())
    check_against_oracle(snap3, ix)


@pytest.mark.gpu
@pytest.mark.parametrize("packed", [False, True])
def test_gpu_written_snapshot_parsed_by_google_protobuf(hnswindex, M, MP, tmp_path, packed):
    """hnsw_serialize on a GPU-built index with removed nodes: parsed fields equal hnsw_export_* / hnsw_ids / parameters."""
    rng = np.random.default_rng(3)
    data = rng.random((1500, 24), dtype=np.float32)
    ix = hnswindex.Index(24, "sq_euclid")
    ix.set_max_edges(6); ix.set_max_candidates(40); ix.set_collection_size(100); ix.set_random_seed(99); ix.set_min_nn(11)
    ids = ix.add(data)
    ix.remove(ids[5:400:9])
    path = tmp_path / "gpu.bin"
    ix.serialize(str(path), packed=packed)
    raw = path.read_bytes()
    snap = (MP if packed else M)["HNSWIndexSnapshot"]()
    snap.ParseFromString(raw)
    assert snap.SerializeToString() == raw
    P, D = snap.Parameters, snap.DataSnapshot
    assert (P.MaxEdges, P.MinNN, P.MaxCandidates, P.CollectionSize, P.RandomSeed, P.AllowRemovals) == (6, 11, 40, 100, 99, True)
    levels, rows0, entry = ix.export_layer0()
    n = len(levels)
    assert D.Length == n == 1500 and D.Capacity >= n and D.EntryPointId == entry == ix.entry_point()
    removed = set(int(i) for i in ids[5:400:9])
    assert D.Count == n - len(removed) == ix.count
    assert list(D.ActiveNodes) == list(ix.ids())
    assert list(D.RemovedIndexes) == [int(i) for i in ids[5:400:9]][::-1]
    upper = {l: ix.export_layer(l) for l in range(1, int(levels.max()) + 1)}
    indeg = np.zeros(n, dtype=np.int64)
    for i, node in enumerate(D.Nodes):
        assert node.Id == i and node.IsRemoved == (i in removed)
        assert len(node.OutEdges) == levels[i] + 1 == len(node.InEdges)
        row = rows0[i][rows0[i] >= 0]
        assert list(node.OutEdges[0].Buffer)[:node.OutEdges[0].Count] == row.tolist()
        indeg[row] += 1
        for l in range(1, levels[i] + 1):
            r = upper[l][i]
            assert list(node.OutEdges[l].Buffer)[:node.OutEdges[l].Count] == r[r >= 0].tolist()
        assert np.array_equal(np.array(D.Items[i].Values, dtype=np.float32), data[i])
    assert [node.InEdges[0].Count for node in D.Nodes] == indeg.tolist()        # in-edges = transpose of the out-edges


def test_crafted_snapshots_are_rejected(hnswindex, oracle, M, tmp_path):
    """RemovedIndexes / EntryPointId / Capacity of a file are validated before the engine indexes its
    arrays with them -- a later Add pops RemovedIndexes and writes levels / rows / vectors at that slot."""
    v = oracle.random_vectors(8, 60)
    ix = oracle.OracleIndex("sq_euclid", max_edges=4)
    ids = ix.add(v)
    ix.remove([int(i) for i in ids[3:20:4]])
    a = tmp_path / "a.bin"
    ix.serialize(a)
    good = M["HNSWIndexSnapshot"]()
    good.ParseFromString(a.read_bytes())
    assert transcode(hnswindex, a, tmp_path / "ok.bin", 0) == 60

    def rejected(mutate, what):
        s = M["HNSWIndexSnapshot"]()
        s.CopyFrom(good)
        mutate(s.DataSnapshot)
        bad = tmp_path / "bad.bin"
        bad.write_bytes(s.SerializeToString())
        rc = hnswindex.lib.hnsw_snapshot_transcode(str(bad).encode(), str(tmp_path / "o.bin").encode(), 0)
        assert rc == -1, what
        assert what in hnswindex.last_error(), hnswindex.last_error()

    def dup(d): d.RemovedIndexes[1] = d.RemovedIndexes[0]
    def oob(d): d.RemovedIndexes[0] = 60
    def neg(d): d.RemovedIndexes[0] = -3
    def live(d): d.RemovedIndexes[0] = 0                     # slot 0 is alive
    def short(d): del d.RemovedIndexes[0]
    def entry_oob(d): d.EntryPointId = 1000
    def entry_removed(d): d.EntryPointId = 3
    def capacity(d): d.Capacity = 2 ** 31 - 1
    def active_dup(d): d.ActiveNodes[1] = d.ActiveNodes[0]
    rejected(dup, "RemovedIndexes")
    rejected(oob, "RemovedIndexes")
    rejected(neg, "RemovedIndexes")
    rejected(live, "RemovedIndexes")
    rejected(short, "RemovedIndexes")
    rejected(entry_oob, "EntryPointId")
    rejected(entry_removed, "entry point")
    rejected(capacity, "Capacity")
    rejected(active_dup, "ActiveNodes")

    # an index whose every node was removed: EntryPointId must still be -1 or a slot of the file
    ix2 = oracle.OracleIndex("sq_euclid", max_edges=4)
    ids2 = ix2.add(v[:5])
    ix2.remove([int(i) for i in ids2])
    e = tmp_path / "empty.bin"
    ix2.serialize(e)
    s = M["HNSWIndexSnapshot"]()
    s.ParseFromString(e.read_bytes())
    assert s.DataSnapshot.Count == 0
    assert transcode(hnswindex, e, tmp_path / "o2.bin", 0) == 5
    s.DataSnapshot.EntryPointId = 77
    bad = tmp_path / "bad2.bin"
    bad.write_bytes(s.SerializeToString())
    assert hnswindex.lib.hnsw_snapshot_transcode(str(bad).encode(), str(tmp_path / "o.bin").encode(), 0) == -1
    assert "EntryPointId" in hnswindex.last_error()
