"""
oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes loader for the CPU restatement of HNSWIndex.Net (oracle/hnsw_oracle.hpp).
Importable only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product never imports this module.
"""
import ctypes as ct
import math
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "liboracle.so"


def build(force=False):
    """Compile oracle/_build/liboracle.so with the committed Makefile."""
    srcs = [_HERE / "oracle_capi.cpp", _HERE / "hnsw_oracle.hpp", _HERE / "snapshot_codec.hpp"]
    if not force and _LIB_PATH.exists() and all(_LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs):
        return _LIB_PATH
    subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = ct.CDLL(str(_LIB_PATH))
    f32p, i32p, u64p = ct.POINTER(ct.c_float), ct.POINTER(ct.c_int), ct.POINTER(ct.c_uint64)
    L.orc_last_error.restype = ct.c_char_p
    L.orc_have_avx.restype = ct.c_int
    L.orc_dotnet_next.argtypes = [ct.c_int, ct.c_int, i32p]
    L.orc_dotnet_next_single.argtypes = [ct.c_int, ct.c_int, f32p]
    L.orc_draw_levels.argtypes = [ct.c_int, ct.c_double, ct.c_int, i32p]
    L.orc_metric.restype = ct.c_float
    L.orc_metric.argtypes = [ct.c_char_p, f32p, f32p, ct.c_int, ct.c_int]
    L.orc_metric_many.argtypes = [ct.c_char_p, f32p, f32p, ct.c_int, ct.c_int, ct.c_int, f32p]
    L.orc_cos_norm.restype = ct.c_float
    L.orc_cos_norm.argtypes = [f32p, ct.c_int]
    L.orc_create.restype = ct.c_void_p
    L.orc_create.argtypes = [ct.c_char_p, ct.c_int, ct.c_double, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int]
    L.orc_free.argtypes = [ct.c_void_p]
    L.orc_use_scalar_metric.argtypes = [ct.c_void_p, ct.c_int]
    L.orc_set_min_nn.argtypes = [ct.c_void_p, ct.c_int]
    for name in ("orc_count", "orc_dim", "orc_entry_point", "orc_capacity", "orc_length"):
        getattr(L, name).argtypes = [ct.c_void_p]
        getattr(L, name).restype = ct.c_int
    L.orc_get_params.argtypes = [ct.c_void_p, i32p, ct.POINTER(ct.c_double)]
    L.orc_add.argtypes = [ct.c_void_p, f32p, ct.c_int, ct.c_int, i32p, ct.c_int, u64p]
    L.orc_add_frozen_batches.argtypes = [ct.c_void_p, f32p, ct.c_int, ct.c_int, i32p, ct.c_int, ct.c_int]
    L.orc_knn_query.argtypes = [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_int, i32p, f32p, ct.c_int, u64p]
    L.orc_knn_query_counters.argtypes = [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_int, u64p]
    L.orc_remove.argtypes = [ct.c_void_p, i32p, ct.c_int]
    L.orc_active_ids.argtypes = [ct.c_void_p, i32p, ct.c_int]
    L.orc_is_removed.argtypes = [ct.c_void_p, ct.c_int]
    L.orc_knn_query_layer.argtypes = [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, i32p, f32p, ct.c_int]
    L.orc_multilayer_knn_query.argtypes = [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, i32p, f32p]
    L.orc_range_query.restype = ct.c_longlong
    L.orc_range_query.argtypes = [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_float, ct.POINTER(ct.c_longlong), i32p, f32p, ct.c_int]
    L.orc_exact_knn.argtypes = [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_int, i32p, f32p, ct.c_int]
    L.orc_node_level.argtypes = [ct.c_void_p, ct.c_int]
    L.orc_node_row.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, i32p, ct.c_int]
    L.orc_export_layer.argtypes = [ct.c_void_p, ct.c_int, ct.c_int, i32p, i32p]
    L.orc_get_items.argtypes = [ct.c_void_p, f32p]
    L.orc_serialize.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_int]
    L.orc_deserialize.restype = ct.c_void_p
    L.orc_deserialize.argtypes = [ct.c_char_p, ct.c_char_p]
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_float))


def _ip(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_int))


def _up(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_uint64))


def dotnet_next(seed, n):
    out = np.empty(n, dtype=np.int32)
    lib().orc_dotnet_next(seed, n, _ip(out))
    return out


def dotnet_next_single(seed, n):
    out = np.empty(n, dtype=np.float32)
    lib().orc_dotnet_next_single(seed, n, _fp(out))
    return out


def random_vectors(dim, count, seed=65537):
    """Tests/Utils.cs:35-49: new Random(65537).NextSingle() i.i.d."""
    return dotnet_next_single(seed, dim * count).reshape(count, dim).copy()


def normalize_rows_like_reference(v):
    """Tests/Utils.cs:9-30: fp32 sequential magnitude, 1f/mag, in-place scale."""
    v = _f32(v).copy()
    for row in v:
        mag = np.float32(0.0)
        for x in row:
            mag = np.float32(mag + np.float32(x * x))
        norm = np.float32(1.0) / np.float32(math.sqrt(float(mag)))
        row *= norm
    return v


def draw_levels(seed, rate, n):
    out = np.empty(n, dtype=np.int32)
    lib().orc_draw_levels(seed, rate, n, _ip(out))
    return out


def metric(name, a, b, impl=0):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_metric(name.encode(), _fp(a), _fp(b), a.shape[-1], impl))


def metric_many(name, a, b, impl=0):
    a, b = _f32(a), _f32(b)
    out = np.empty(a.shape[0], dtype=np.float32)
    lib().orc_metric_many(name.encode(), _fp(a), _fp(b), a.shape[1], a.shape[0], impl, _fp(out))
    return out


class OracleIndex:
    """CPU restatement of HNSWIndex<float[], float> (HNSWIndex.cs)."""

    def __init__(self, metric="sq_euclid", max_edges=16, distribution_rate=1.0 / math.log(16.0), min_nn=5,
                 max_candidates=100, collection_size=65536, random_seed=31337, allow_removals=True, _handle=None):
        self.metric = metric
        if _handle is not None:
            self._h = _handle
        else:
            self._h = lib().orc_create(metric.encode(), max_edges, distribution_rate, min_nn, max_candidates,
                                       collection_size, random_seed, int(allow_removals))
            if not self._h:
                raise RuntimeError(lib().orc_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_free(self._h)
            self._h = None

    # -- properties
    @property
    def count(self):
        return lib().orc_count(self._h)

    @property
    def dim(self):
        return lib().orc_dim(self._h)

    @property
    def entry_point(self):
        return lib().orc_entry_point(self._h)

    @property
    def capacity(self):
        return lib().orc_capacity(self._h)

    @property
    def length(self):
        return lib().orc_length(self._h)

    def params(self):
        ints = np.empty(7, dtype=np.int32)
        rate = ct.c_double()
        lib().orc_get_params(self._h, _ip(ints), ct.byref(rate))
        keys = ["max_edges", "min_nn", "max_candidates", "remove_max_candidates", "collection_size", "random_seed",
                "allow_removals"]
        d = dict(zip(keys, [int(x) for x in ints]))
        d["distribution_rate"] = rate.value
        return d

    def set_min_nn(self, v):
        lib().orc_set_min_nn(self._h, v)

    def use_scalar_metric(self, on=True):
        lib().orc_use_scalar_metric(self._h, int(on))

    # -- operations
    def add(self, vecs, threads=1, counters=False):
        v = _f32(vecs)
        if v.ndim == 1:
            v = v.reshape(1, -1)
        ids = np.empty(v.shape[0], dtype=np.int32)
        c = np.zeros(3, dtype=np.uint64)
        rc = lib().orc_add(self._h, _fp(v), v.shape[0], v.shape[1], _ip(ids), threads, _up(c) if counters else None)
        if rc < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return (ids, c) if counters else ids

    def add_frozen_batches(self, vecs, ratio=4, bmax=4096):
        v = _f32(vecs)
        ids = np.empty(v.shape[0], dtype=np.int32)
        rc = lib().orc_add_frozen_batches(self._h, _fp(v), v.shape[0], v.shape[1], _ip(ids), ratio, bmax)
        if rc < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return ids

    def knn_query(self, queries, k, threads=1, counters=False):
        q = _f32(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        ids = np.empty((q.shape[0], k), dtype=np.int32)
        dists = np.empty((q.shape[0], k), dtype=np.float32)
        c = np.zeros(3, dtype=np.uint64)
        rc = lib().orc_knn_query(self._h, _fp(q), q.shape[0], q.shape[1], k, _ip(ids), _fp(dists), threads,
                                 _up(c) if counters else None)
        if rc < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return (ids, dists, c) if counters else (ids, dists)

    def remove(self, ids):
        """HNSWIndex.Remove for each id in order (HNSWIndex.cs:83-100), single-threaded."""
        a = np.ascontiguousarray(np.asarray(ids, dtype=np.int32).reshape(-1))
        if lib().orc_remove(self._h, _ip(a), a.shape[0]) < 0:
            raise RuntimeError(lib().orc_last_error().decode())

    def active_ids(self):
        n = lib().orc_active_ids(self._h, None, 0)
        out = np.empty(max(n, 1), dtype=np.int32)
        lib().orc_active_ids(self._h, _ip(out), n)
        return out[:n]

    def is_removed(self, i):
        return bool(lib().orc_is_removed(self._h, int(i)))

    def knn_query_layer(self, queries, k, layer, threads=1):
        """BatchKnnQuery(..., layer) -- HNSWIndex.cs:107-137."""
        q = _f32(queries)
        ids = np.empty((q.shape[0], k), dtype=np.int32)
        dists = np.empty((q.shape[0], k), dtype=np.float32)
        rc = lib().orc_knn_query_layer(self._h, _fp(q), q.shape[0], q.shape[1], k, int(layer), _ip(ids), _fp(dists), threads)
        if rc < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return ids, dists

    def multilayer_knn_query(self, queries, k, max_layer=2 ** 31 - 1, min_layer=0):
        """MultiLayerKnnQuery (HNSWIndex.cs:173-187) per query: (ids, dists) of shape [n, layers, k], -1 / NaN padded."""
        q = _f32(queries)
        layers = lib().orc_multilayer_knn_query(self._h, _fp(q), min(1, q.shape[0]), q.shape[1], k, int(max_layer), int(min_layer), 0, None, None)
        if layers < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        ids = np.empty((q.shape[0], layers, k), dtype=np.int32)
        dists = np.empty((q.shape[0], layers, k), dtype=np.float32)
        rc = lib().orc_multilayer_knn_query(self._h, _fp(q), q.shape[0], q.shape[1], k, int(max_layer), int(min_layer), layers, _ip(ids), _fp(dists))
        if rc < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return ids, dists

    def knn_query_counters(self, queries, k):
        """Per-query (ndist, nexp, row_words) -- the roofline's algorithmic work."""
        q = _f32(queries)
        out = np.zeros((q.shape[0], 3), dtype=np.uint64)
        rc = lib().orc_knn_query_counters(self._h, _fp(q), q.shape[0], q.shape[1], k, _up(out))
        if rc < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return out

    def range_query(self, queries, radius, threads=os.cpu_count() or 1):
        """HNSWIndex.BatchRangeQuery (HNSWIndex.cs:160-168): per-query (ids, dists), ascending by distance."""
        q = _f32(queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        n = q.shape[0]
        off = np.zeros(n + 1, dtype=np.int64)
        offp = off.ctypes.data_as(ct.POINTER(ct.c_longlong))
        total = lib().orc_range_query(self._h, _fp(q), n, q.shape[1], float(radius), offp, None, None, threads)
        if total < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        ids = np.empty(max(total, 1), dtype=np.int32)
        d = np.empty(max(total, 1), dtype=np.float32)
        lib().orc_range_query(self._h, _fp(q), n, q.shape[1], float(radius), offp, _ip(ids), _fp(d), threads)
        return [ids[off[i]:off[i + 1]].copy() for i in range(n)], [d[off[i]:off[i + 1]].copy() for i in range(n)]

    def exact_knn(self, queries, k, threads=os.cpu_count() or 1):
        q = _f32(queries)
        ids = np.empty((q.shape[0], k), dtype=np.int32)
        dists = np.empty((q.shape[0], k), dtype=np.float32)
        rc = lib().orc_exact_knn(self._h, _fp(q), q.shape[0], q.shape[1], k, _ip(ids), _fp(dists), threads)
        if rc < 0:
            raise RuntimeError(lib().orc_last_error().decode())
        return ids, dists

    # -- inspection
    def node_level(self, i):
        return lib().orc_node_level(self._h, i)

    def row(self, i, layer, which=0):
        buf = np.empty(4096, dtype=np.int32)
        n = lib().orc_node_row(self._h, i, layer, which, _ip(buf), buf.size)
        return None if n < 0 else buf[:n].copy()

    def export_layer(self, layer, stride):
        n = self.length
        rows = np.empty((n, stride), dtype=np.int32)
        levels = np.empty(n, dtype=np.int32)
        maxdeg = lib().orc_export_layer(self._h, layer, stride, _ip(rows), _ip(levels))
        return rows, levels, maxdeg

    def items(self):
        out = np.empty((self.length, self.dim), dtype=np.float32)
        lib().orc_get_items(self._h, _fp(out))
        return out

    # -- snapshot (HNSWIndex.cs:210-229)
    def serialize(self, path, packed=False):
        if lib().orc_serialize(self._h, str(path).encode(), int(packed)) < 0:
            raise RuntimeError(lib().orc_last_error().decode())

    @classmethod
    def deserialize(cls, metric, path):
        h = lib().orc_deserialize(metric.encode(), str(path).encode())
        if not h:
            raise RuntimeError(lib().orc_last_error().decode())
        return cls(metric=metric, _handle=h)


def self_recall(index, vectors, ids=None, k=1, threads=1):
    """Tests/Utils.cs:54-70 for k=1: is the query's own vector the top result."""
    res, _ = index.knn_query(vectors, k, threads=threads)
    ids = np.arange(len(vectors), dtype=np.int32) if ids is None else ids
    return float((res[:, 0] == ids).sum()) / len(vectors)
