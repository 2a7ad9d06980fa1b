"""
ctypes host for the B200-native HNSW engine.

Same public surface as the reference's Python package (`bindings/bindings.py:142-597` of
HNSWIndex.Net): ``Index(dim, metric)``, the ``set_*`` configuration methods that must precede the
first ``add``, ``add``, ``remove``, ``knn_query``, ``range_query`` -- so the reference's own tests
(`bindings/__tests__/*.py`) read the same against this package.  On top of that it exposes what the
C# class has but the reference's native shim lacks (``serialize`` / ``deserialize`` / ``count``) and
the engine's measurement hooks.

The shared library is the one `hnswindex.net_b200/build.py` produces at
``artifacts/native/linux-x64/HNSWIndex.Native.so`` -- the location the reference package loads
its own native library from (`bindings.py:35-41`).  There is no CPU fallback: without the
library this module raises on import, without a CUDA device ``hnsw_create`` fails loudly.
"""
import ctypes as ct
import platform
import sys
from pathlib import Path
from typing import List, Tuple

import numpy as np

_METRICS = ("sq_euclid", "cosine", "ucosine", "euclid")  # "euclid" = EuclideanMetric.Compute (C# only in the reference)


def _runtime_id() -> str:
    arch = platform.machine().lower()
    if platform.system() == "Linux":
        return "linux-arm64" if arch in ("aarch64", "arm64") else "linux-x64"
    raise RuntimeError("the B200 engine is built for Linux only")


LIB_PATH = Path(__file__).resolve().parent / "artifacts" / "native" / _runtime_id() / "HNSWIndex.Native.so"


def _load():
    if not LIB_PATH.exists():
        raise FileNotFoundError(f"Native library missing {LIB_PATH} -- run `python hnswindex.net_b200/build.py`")
    L = ct.CDLL(str(LIB_PATH))
    f32p, i32p, vpp = ct.POINTER(ct.c_float), ct.POINTER(ct.c_int), ct.POINTER(ct.c_void_p)
    sig = {
        "hnsw_get_last_error_utf8": (ct.c_int, [ct.c_void_p, ct.c_int]),
        "hnsw_create": (ct.c_void_p, [ct.c_char_p]),
        "hnsw_free": (None, [ct.c_void_p]),
        "hnsw_add": (ct.c_int, [ct.c_void_p, f32p, ct.c_int, ct.c_int, i32p]),
        "hnsw_remove": (ct.c_int, [ct.c_void_p, i32p, ct.c_int]),
        "hnsw_knn_query": (ct.c_int, [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_int, i32p, f32p]),
        "hnsw_range_query": (ct.c_int, [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_float, vpp, vpp, i32p]),
        "hnsw_free_results": (None, [vpp, vpp, ct.c_int]),
        "hnsw_set_collection_size": (ct.c_int, [ct.c_int]),
        "hnsw_set_max_edges": (ct.c_int, [ct.c_int]),
        "hnsw_set_max_candidates": (ct.c_int, [ct.c_int]),
        "hnsw_set_remove_max_candidates": (ct.c_int, [ct.c_int]),
        "hnsw_set_distribution_rate": (ct.c_int, [ct.c_float]),
        "hnsw_set_random_seed": (ct.c_int, [ct.c_int]),
        "hnsw_set_min_nn": (ct.c_int, [ct.c_int]),
        "hnsw_set_allow_removals": (ct.c_int, [ct.c_bool]),
        # additions
        "hnsw_count": (ct.c_int, [ct.c_void_p]),
        "hnsw_dim": (ct.c_int, [ct.c_void_p]),
        "hnsw_serialize": (ct.c_int, [ct.c_void_p, ct.c_char_p, ct.c_int]),
        "hnsw_deserialize": (ct.c_void_p, [ct.c_char_p, ct.c_char_p]),
        "hnsw_index_set_min_nn": (ct.c_int, [ct.c_void_p, ct.c_int]),
        "hnsw_index_get_params": (ct.c_int, [ct.c_void_p, i32p, ct.POINTER(ct.c_double)]),
        "hnsw_set_device": (ct.c_int, [ct.c_int]),
        "hnsw_device_count": (ct.c_int, []),
        "hnsw_knn_query_device": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p]),
        "hnsw_knn_query_device_async": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p]),
        "hnsw_synchronize": (ct.c_int, [ct.c_void_p]),
        "hnsw_async_join": (ct.c_int, [ct.c_void_p]),
        "hnsw_set_stream": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
        "hnsw_exact_knn": (ct.c_int, [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_int, i32p, f32p]),
        "hnsw_exact_knn_tf32": (ct.c_int, [ct.c_void_p, f32p, ct.c_int, ct.c_int, ct.c_int, i32p, f32p, i32p]),
        "hnsw_exact_tcgen05_supported": (ct.c_int, [ct.c_int, ct.c_int]),
        "hnsw_merge_topk_device": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p, ct.c_void_p]),
        "hnsw_enable_stats": (ct.c_int, [ct.c_void_p, ct.c_int]),
        "hnsw_get_stats": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
        "hnsw_knn_query_layer": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p]),
        "hnsw_multilayer_knn_query": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p, ct.c_void_p]),
        "hnsw_ids": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int]),
        "hnsw_range_query_flat": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int, ct.c_float, ct.c_void_p, ct.c_void_p, ct.c_void_p]),
        "hnsw_free_flat": (None, [ct.c_void_p, ct.c_void_p]),
        "hnsw_ipc_alloc": (ct.c_int, [ct.c_size_t, ct.c_void_p, ct.c_void_p]),
        "hnsw_ipc_open": (ct.c_int, [ct.c_void_p, ct.c_void_p]),
        "hnsw_ipc_close": (ct.c_int, [ct.c_void_p]),
        "hnsw_ipc_free": (ct.c_int, [ct.c_void_p]),
        "hnsw_knn_query_scatter": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_void_p]),
        "hnsw_partition_sync": (ct.c_int, [ct.c_void_p]),
        "hnsw_wait_and_merge_device_async": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_uint, ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int,
                                                        ct.c_void_p, ct.c_void_p]),
        "hnsw_wait_and_merge_device": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_uint, ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int,
                                                  ct.c_void_p, ct.c_void_p]),
        "hnsw_top_layer": (ct.c_int, [ct.c_void_p]),
        "hnsw_entry_point": (ct.c_int, [ct.c_void_p]),
        "hnsw_get_info": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int]),
        "hnsw_connected_components": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int]),
        "hnsw_get_build_profile": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int]),
        "hnsw_get_wide_profile": (ct.c_int, [ct.c_void_p, ct.c_void_p, ct.c_int]),
        "hnsw_set_tuning": (ct.c_int, [ct.c_void_p, ct.c_char_p, ct.c_int]),
        "hnsw_export_layer0": (ct.c_int, [ct.c_void_p, i32p, i32p, i32p, i32p]),
        "hnsw_export_layer": (ct.c_int, [ct.c_void_p, ct.c_int, i32p, i32p]),
        "hnsw_export_layer0_n": (ct.c_int, [ct.c_void_p, ct.c_int, i32p, i32p, i32p, i32p]),
        "hnsw_export_layer_n": (ct.c_int, [ct.c_void_p, ct.c_int, ct.c_int, i32p, i32p]),
        "hnsw_snapshot_transcode": (ct.c_int, [ct.c_char_p, ct.c_char_p, ct.c_int]),
        "hnsw_version": (ct.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return L


lib = _load()


class Scatter(ct.Structure):
    """hnsw_scatter of include/hnsw_b200.h (partitioned index, fused search + exchange)."""
    _fields_ = [("world", ct.c_int), ("rank", ct.c_int), ("idOffset", ct.c_int), ("epoch", ct.c_uint),
                ("ids", ct.POINTER(ct.c_int) * 16), ("dists", ct.POINTER(ct.c_float) * 16), ("flags", ct.POINTER(ct.c_uint) * 16)]


class _LayerInfo(ct.Structure):
    _fields_ = [("layerId", ct.c_int), ("nodesCount", ct.c_int), ("maxOutEdges", ct.c_int), ("minOutEdges", ct.c_int),
                ("maxInEdges", ct.c_int), ("minInEdges", ct.c_int), ("avgOutEdges", ct.c_double), ("avgInEdges", ct.c_double),
                ("outEdgesMedian", ct.c_int), ("inEdgesMedian", ct.c_int)]


class _Stats(ct.Structure):
    _fields_ = [("last_search_ms", ct.c_double), ("last_build_ms", ct.c_double), ("ndist", ct.c_ulonglong),
                ("nexp", ct.c_ulonglong), ("row_words", ct.c_ulonglong), ("search_launches", ct.c_ulonglong),
                ("build_launches", ct.c_ulonglong)]


def last_error() -> str:
    n = lib.hnsw_get_last_error_utf8(None, 0)
    if n <= 0:
        return ""
    buf = ct.create_string_buffer(n + 1)
    lib.hnsw_get_last_error_utf8(buf, len(buf))
    return buf.value.decode("utf-8", "replace")


def _rows_f32(x, dim=None) -> np.ndarray:
    a = np.asarray(x, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.ndim != 2:
        raise ValueError("expected a 2D array of shape (n, dim) or a 1D vector")
    if dim is not None and a.shape[1] != dim:
        raise ValueError(f"expected dim={dim}, got {a.shape[1]}")
    return np.ascontiguousarray(a)


def _fp(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_float))


def _ip(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_int))


def _check(status):
    if status < 0:
        raise RuntimeError(last_error())


class Index:
    """HNSW index over float32 vectors of one dimensionality, resident on one B200.

    Parameters mirror the reference: ``Index(dim, metric="sq_euclid" | "cosine" | "ucosine")`` (+ ``"euclid"``, the C# library's EuclideanMetric).
    The native index is created lazily by the first ``add`` (or by ``deserialize``); the
    ``set_*`` methods configure that creation and therefore must be called before it.
    """

    def __init__(self, dim: int, metric: str = "sq_euclid"):
        self.dim = int(dim)
        self.metric = metric
        self._initialized = False
        self._h = None
        self.last_exact_fallbacks = 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.hnsw_free(h)

    def _create(self):
        h = lib.hnsw_create(self.metric.encode("utf-8"))
        if not h:
            raise RuntimeError("hnsw_create failed: " + last_error())
        self._h, self._initialized = h, True

    # ---- configuration of the NEXT created index (HNSWParameters.cs:13-55) ----------------------
    def set_collection_size(self, init_size: int):
        _check(lib.hnsw_set_collection_size(int(init_size)))

    def set_max_edges(self, max_conn: int):
        _check(lib.hnsw_set_max_edges(int(max_conn)))

    def set_max_candidates(self, max_candidates: int):
        _check(lib.hnsw_set_max_candidates(int(max_candidates)))

    def set_remove_max_candidates(self, rem_max_candidates: int):
        _check(lib.hnsw_set_remove_max_candidates(int(rem_max_candidates)))

    def set_distribution_rate(self, dist_rate: float):
        _check(lib.hnsw_set_distribution_rate(float(dist_rate)))

    def set_random_seed(self, random_seed: int):
        _check(lib.hnsw_set_random_seed(int(random_seed)))

    def set_min_nn(self, min_nn: int):
        """Before the first add: MinNN of the index to be created.  Afterwards: change MinNN (= ef) live,
        as mutating HNSWParameters.MinNN does in the C# API."""
        if self._initialized:
            _check(lib.hnsw_index_set_min_nn(self._h, int(min_nn)))
        else:
            _check(lib.hnsw_set_min_nn(int(min_nn)))

    def set_allow_removals(self, allow_removals: bool):
        _check(lib.hnsw_set_allow_removals(bool(allow_removals)))

    # ---- operations --------------------------------------------------------------------------------
    def add(self, vecs) -> np.ndarray:
        """Insert vectors; returns their int32 ids (HNSWIndex.Add, HNSWIndex.cs:70-78)."""
        if not self._initialized:
            self._create()
        a = _rows_f32(vecs, self.dim)
        out = np.empty(a.shape[0], dtype=np.int32)
        rc = lib.hnsw_add(self._h, _fp(a), int(a.shape[0]), int(a.shape[1]), _ip(out))
        _check(rc)
        return out[:rc].copy()

    def remove(self, ids) -> None:
        arr = np.ascontiguousarray(np.asarray(ids, dtype=np.int32).ravel())
        if arr.size == 0:
            return
        _check(lib.hnsw_remove(self._h, _ip(arr), int(arr.size)))

    def knn_query(self, queries, k: int) -> Tuple[np.ndarray, np.ndarray]:
        """Batched k-NN (HNSWIndex.BatchKnnQuery, HNSWIndex.cs:129-137): (ids[n,k] int32, dists[n,k] float32)."""
        q = _rows_f32(queries, self.dim)
        n = int(q.shape[0])
        ids = np.empty((n, k), dtype=np.int32)
        dists = np.empty((n, k), dtype=np.float32)
        _check(lib.hnsw_knn_query(self._h, _fp(q), n, self.dim, int(k), _ip(ids), _fp(dists)))
        return ids, dists

    def range_query(self, queries, radius: float) -> Tuple[List[np.ndarray], List[np.ndarray]]:
        """BatchRangeQuery (HNSWIndex.cs:160-168): per query the ids and distances within `radius`, ascending.  Uses the
        flat variant of the export (one buffer per output) -- the per-query marshalling of the reference's binding
        (bindings.py:523-597) would cost more than the search."""
        q = _rows_f32(queries, self.dim)
        n = int(q.shape[0])
        if self._h is None or n == 0:
            return [np.empty(0, dtype=np.int32) for _ in range(n)], [np.empty(0, dtype=np.float32) for _ in range(n)]
        offsets = np.zeros(n + 1, dtype=np.int64)
        ids_p, dists_p = ct.POINTER(ct.c_int)(), ct.POINTER(ct.c_float)()
        _check(lib.hnsw_range_query_flat(self._h, _fp(q), n, self.dim, float(radius), offsets.ctypes.data_as(ct.POINTER(ct.c_longlong)),
                                         ct.byref(ids_p), ct.byref(dists_p)))
        try:
            total = int(offsets[n])
            if total > 0:
                ids_all = np.ctypeslib.as_array(ids_p, shape=(total,)).copy()
                d_all = np.ctypeslib.as_array(dists_p, shape=(total,)).copy()
            else:
                ids_all, d_all = np.empty(0, dtype=np.int32), np.empty(0, dtype=np.float32)
        finally:
            lib.hnsw_free_flat(ids_p, dists_p)
        cut = offsets[1:-1]
        return np.split(ids_all, cut), np.split(d_all, cut)

    # ---- what the C# class has and the reference's native shim lacks ------------------------------
    @property
    def count(self) -> int:
        return lib.hnsw_count(self._h) if self._h else 0

    def serialize(self, path, packed: bool = False) -> None:
        """HNSWIndex.Serialize (HNSWIndex.cs:210-217)."""
        _check(lib.hnsw_serialize(self._h, str(path).encode(), int(packed)))

    @classmethod
    def deserialize(cls, metric: str, path) -> "Index":
        """HNSWIndex.Deserialize (HNSWIndex.cs:222-229)."""
        h = lib.hnsw_deserialize(metric.encode(), str(path).encode())
        if not h:
            raise RuntimeError("hnsw_deserialize failed: " + last_error())
        ix = cls(lib.hnsw_dim(h), metric)
        ix._h, ix._initialized = h, True
        return ix

    def params(self) -> dict:
        ints = np.empty(7, dtype=np.int32)
        rate = ct.c_double()
        _check(lib.hnsw_index_get_params(self._h, _ip(ints), ct.byref(rate)))
        keys = ["max_edges", "min_nn", "max_candidates", "remove_max_candidates", "collection_size", "random_seed", "allow_removals"]
        d = dict(zip(keys, (int(x) for x in ints)))
        d["distribution_rate"] = rate.value
        return d

    # ---- C# surface the reference's native shim does not export (HNSWIndex.cs) ---------------------------
    def knn_query_layer(self, queries, k: int, layer: int) -> Tuple[np.ndarray, np.ndarray]:
        """BatchKnnQuery(queries, k, layer: layer) -- HNSWIndex.cs:129-137."""
        q = _rows_f32(queries, self.dim)
        n = int(q.shape[0])
        ids = np.empty((n, k), dtype=np.int32)
        dists = np.empty((n, k), dtype=np.float32)
        if self._h is None:
            ids.fill(-1); dists.fill(np.nan)
            return ids, dists
        _check(lib.hnsw_knn_query_layer(self._h, _fp(q), n, self.dim, int(k), int(layer), _ip(ids), _fp(dists)))
        return ids, dists

    def multilayer_knn_query(self, queries, k: int, max_layer: int = 2 ** 31 - 1, min_layer: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        """MultiLayerKnnQuery (HNSWIndex.cs:173-187) for every query: (ids, dists) of shape [n, layers, k]; a layer's row is
        its ascending result without the first entry, padded with -1 / NaN."""
        q = _rows_f32(queries, self.dim)
        n = int(q.shape[0])
        layers = lib.hnsw_multilayer_knn_query(self._h, None, 0, self.dim, int(k), int(max_layer), int(min_layer), None, None) if self._h else 0
        _check(layers)
        ids = np.empty((n, layers, k), dtype=np.int32)
        dists = np.empty((n, layers, k), dtype=np.float32)
        if layers > 0 and n > 0:
            _check(lib.hnsw_multilayer_knn_query(self._h, _fp(q), n, self.dim, int(k), int(max_layer), int(min_layer), _ip(ids), _fp(dists)))
        return ids, dists

    def ids(self) -> np.ndarray:
        """Ids() (HNSWIndex.cs:242-245): active ids in the reference's ActiveSet order."""
        if not self._h:
            return np.zeros(0, dtype=np.int32)
        n = lib.hnsw_ids(self._h, None, 0)
        _check(n)
        out = np.zeros(max(n, 1), dtype=np.int32)
        _check(lib.hnsw_ids(self._h, _ip(out), n))
        return out[:n]

    def top_layer(self) -> int:
        return lib.hnsw_top_layer(self._h) if self._h else -1

    def entry_point(self) -> int:
        return lib.hnsw_entry_point(self._h) if self._h else -1

    def get_info(self) -> List[dict]:
        """GetInfo (HNSWIndex.cs:192-196): one dict per layer with the fields of HNSWInfo.LayerInfo."""
        if not self._h:
            return []
        n = lib.hnsw_get_info(self._h, None, 0)
        _check(n)
        arr = (_LayerInfo * max(n, 1))()
        _check(lib.hnsw_get_info(self._h, arr, n))
        return [{f: getattr(arr[i], f) for f, _ in _LayerInfo._fields_} for i in range(n)]

    def connected_component_counts(self) -> np.ndarray:
        """GetConnectedComponentCounts (HNSWIndex.cs:202-205)."""
        if not self._h:
            return np.zeros(0, dtype=np.int32)
        n = lib.hnsw_connected_components(self._h, None, 0)
        _check(n)
        out = np.zeros(max(n, 1), dtype=np.int32)
        _check(lib.hnsw_connected_components(self._h, _ip(out), n))
        return out[:n]

    # ---- engine extras -------------------------------------------------------------------------------
    def exact_knn(self, queries, k: int, tensor_cores: bool = False) -> Tuple[np.ndarray, np.ndarray]:
        """Brute-force exact top-k on the GPU (recall ground truth).

        tensor_cores=True ranks all pairs with a 3xTF32 GEMM first and evaluates only k + slack candidates per
        query with the exact metric; the answers are the same, `last_exact_fallbacks` tells how many queries could
        not be certified and went through the plain scan."""
        q = _rows_f32(queries, self.dim)
        n = int(q.shape[0])
        ids = np.empty((n, k), dtype=np.int32)
        dists = np.empty((n, k), dtype=np.float32)
        if tensor_cores:
            fb = np.zeros(1, dtype=np.int32)
            _check(lib.hnsw_exact_knn_tf32(self._h, _fp(q), n, self.dim, int(k), _ip(ids), _fp(dists), _ip(fb)))
            self.last_exact_fallbacks = int(fb[0])
        else:
            _check(lib.hnsw_exact_knn(self._h, _fp(q), n, self.dim, int(k), _ip(ids), _fp(dists)))
        return ids, dists

    def knn_query_device(self, q_ptr: int, n: int, k: int, ids_ptr: int, dists_ptr: int) -> None:
        """Queries / outputs already in this GPU's memory (raw device pointers, e.g. tensor.data_ptr())."""
        _check(lib.hnsw_knn_query_device(self._h, ct.c_void_p(q_ptr), int(n), self.dim, int(k), ct.c_void_p(ids_ptr), ct.c_void_p(dists_ptr)))

    def knn_query_device_async(self, q_ptr: int, n: int, k: int, ids_ptr: int, dists_ptr: int) -> None:
        """Queue a device-resident batch on the index's stream without waiting; synchronize() waits."""
        _check(lib.hnsw_knn_query_device_async(self._h, ct.c_void_p(q_ptr), int(n), self.dim, int(k), ct.c_void_p(ids_ptr), ct.c_void_p(dists_ptr)))

    def synchronize(self) -> None:
        _check(lib.hnsw_synchronize(self._h))

    def async_join(self) -> None:
        """Stream-ordered: the index's stream waits for the batches queued on the auxiliary stream (no host wait)."""
        _check(lib.hnsw_async_join(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        _check(lib.hnsw_set_stream(self._h, ct.c_void_p(cuda_stream)))

    def enable_stats(self, on: bool = True) -> None:
        _check(lib.hnsw_enable_stats(self._h, int(on)))

    def stats(self) -> dict:
        s = _Stats()
        _check(lib.hnsw_get_stats(self._h, ct.byref(s)))
        return {f: getattr(s, f) for f, _ in s._fields_}

    _BUILD_PROFILE_FIELDS = ("inserts", "cyc_descent", "cyc_search", "cyc_select", "ndist_search", "ndist_select", "cyc_link_kernel",
                             "cyc_prune", "prunes", "ndist_prune", "cyc_lock", "cyc_select_kernel", "links")

    def build_profile(self, reset: bool = False) -> dict:
        """Cycle / work counters of the construction kernels (per-warp SM cycles summed over warps)."""
        out = (ct.c_ulonglong * 16)()
        _check(lib.hnsw_get_build_profile(self._h, out, int(reset)))
        return {f: int(out[i]) for i, f in enumerate(self._BUILD_PROFILE_FIELDS)}

    _WIDE_PROFILE_FIELDS = ("queries", "expansions", "spec_issued", "spec_used", "cyc_descent", "cyc_select_filter", "cyc_wait_spec",
                            "cyc_wait_eval", "cyc_insert", "cyc_beam")

    def wide_profile(self, reset: bool = False) -> dict:
        """Leader-warp counters of the CTA-per-query wide-beam kernel (SM cycles summed over queries)."""
        out = (ct.c_ulonglong * 16)()
        _check(lib.hnsw_get_wide_profile(self._h, out, int(reset)))
        return {f: int(out[i]) for i, f in enumerate(self._WIDE_PROFILE_FIELDS)}

    def set_tuning(self, key: str, value: int) -> None:
        _check(lib.hnsw_set_tuning(self._h, key.encode(), int(value)))

    def export_layer0(self):
        """(levels[n], rows0[n, stride0] with -1 padding, entry point).  Safe while other threads Add: the copy is bounded by
        the buffers' size and repeated when the index grew in between."""
        stride, ep = ct.c_int(), ct.c_int()
        n = lib.hnsw_export_layer0_n(self._h, 0, None, None, ct.byref(stride), ct.byref(ep))
        _check(n)
        while True:
            levels = np.empty(n, dtype=np.int32)
            rows = np.empty((n, stride.value), dtype=np.int32)
            m = lib.hnsw_export_layer0_n(self._h, n, _ip(levels), _ip(rows), ct.byref(stride), ct.byref(ep))
            _check(m)
            if m <= n:
                return levels[:m], rows[:m], ep.value
            n = m

    def export_layer(self, layer: int):
        stride = ct.c_int()
        n = lib.hnsw_export_layer_n(self._h, int(layer), 0, None, ct.byref(stride))
        _check(n)
        while True:
            rows = np.empty((n, stride.value), dtype=np.int32)
            m = lib.hnsw_export_layer_n(self._h, int(layer), n, _ip(rows), ct.byref(stride))
            _check(m)
            if m <= n:
                return rows[:m]
            n = m


def set_device(device: int) -> None:
    _check(lib.hnsw_set_device(int(device)))


def device_count() -> int:
    return lib.hnsw_device_count()
