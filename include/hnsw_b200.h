/*
 * hnsw_b200.h -- C ABI of the B200-native HNSW engine.
 *
 * Drop-in boundary: the library exports exactly the cdecl symbols that the reference's NativeAOT
 * shim `bindings/HNSWIndex.Native/HNSWIndexExports.cs` exports (and that `bindings/bindings.py`
 * binds with ctypes, lines 46-119), with the same argument meaning, return codes and error
 * behaviour, so the reference's Python package and any P/Invoke host work unchanged when the
 * shared object is installed as  <package>/artifacts/native/linux-x64/HNSWIndex.Native.so.
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Part 1 = the reference's 16 exports.  Part 2 = additions the C# surface has but the reference's
 * native shim lacks (Serialize/Deserialize/Count), plus device selection, resident-buffer entry
 * points and measurement hooks.  All citations are relative to /root/reference.
 *
 * Errors never cross the ABI as exceptions: functions return -1 (or a null handle) and the text
 * is available from hnsw_get_last_error_utf8, as in HNSWIndexExports.cs:11-12,27-39.
 * Threading: as the reference (README.md:64-65) -- operations of one type may run concurrently with each other.  Every
 * handle carries a reader/writer lock: hnsw_knn_query / hnsw_knn_query_layer and the read-only accessors (hnsw_ids,
 * hnsw_count, hnsw_top_layer, hnsw_entry_point, hnsw_get_stats, hnsw_index_get_params) take it shared -- concurrent
 * hnsw_knn_query calls from several host threads each run on their own stream and scratch and overlap on the GPU --, every
 * other call on the handle takes it exclusively, so mixing operation types from different threads is safe too (they queue).
 */
#ifndef HNSW_B200_H
#define HNSW_B200_H

#include <stdbool.h>
#include <stdint.h>

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define HNSW_API __declspec(dllexport)
#else
#define HNSW_API __attribute__((visibility("default")))
#endif

typedef void* hnsw_handle;

/* ------------------------------------------------------------------ Part 1: reference exports */

/* HNSWIndexExports.cs:27-39.  Returns the UTF-8 byte length of the last error (excluding NUL);
 * if buf != NULL and bufLen > 0, writes at most bufLen-1 bytes and a terminating NUL. */
HNSW_API int hnsw_get_last_error_utf8(char* buf, int bufLen);

/* HNSWIndexExports.cs:41-65.  metric: "sq_euclid" | "cosine" | "ucosine".  Consumes the pending
 * process-global parameters (hnsw_set_*) and resets them to defaults.  NULL on error. */
HNSW_API hnsw_handle hnsw_create(const char* distanceMetric);

/* HNSWIndexExports.cs:67-73.  No-op on NULL. */
HNSW_API void hnsw_free(hnsw_handle handle);

/* HNSWIndexExports.cs:75-100 -> HNSWIndex.Add(List) (HNSWIndex.cs:70-78).  vectors: row-major
 * [count][dim] fp32, caller-owned (copied before return).  Writes `count` ids (slot numbers; -1 for
 * a rejected item).  Returns the number written, 0 if handle/vectors is NULL or count/dim <= 0,
 * -1 on error.  The dimension is fixed by the first call. */
HNSW_API int hnsw_add(hnsw_handle handle, const float* vectors, int count, int dim, int* outIds);

/* HNSWIndexExports.cs:102-117 -> HNSWIndex.Remove(List<int>) (HNSWIndex.cs:83-100, GraphConnector.cs:53-167).
 * Returns 0 (also for NULL handle / NULL ids / count <= 0, like the reference), -1 on error: AllowRemovals is
 * false, an id is out of range, already removed or listed twice.  Removed slots are handed out again by later
 * hnsw_add calls, most recently removed first (GraphData.cs:84-91). */
HNSW_API int hnsw_remove(hnsw_handle handle, const int* ids, int count);
/* HNSWIndex.Ids() (HNSWIndex.cs:242-245): the active ids in the reference's ActiveSet order.  Returns their number
 * (= hnsw_count) and writes at most `capacity` of them. */
HNSW_API int hnsw_ids(hnsw_handle handle, int* outIds, int capacity);

/* HNSWIndexExports.cs:119-149 -> HNSWIndex.BatchKnnQuery (HNSWIndex.cs:129-137).  outIds/outDists:
 * caller-owned row-major [count][k]; each row ascending by distance, padded with id -1 / NaN when
 * fewer than k results exist.  Returns 0, or -1 on error; returns 0 without writing if handle is NULL. */
HNSW_API int hnsw_knn_query(hnsw_handle handle, const float* vectors, int count, int dim, int k, int* outIds, float* outDists);

/* HNSWIndexExports.cs:151-197 -> HNSWIndex.BatchRangeQuery (HNSWIndex.cs:160-168).  The callee
 * allocates outIds[i]/outDists[i] (malloc) and sets counts[i]; release with hnsw_free_results. */
HNSW_API int hnsw_range_query(hnsw_handle handle, const float* vectors, int count, int dim, float range, void** outIds, void** outDists,
                              int* counts);

/* HNSWIndexExports.cs:199-217. */
HNSW_API void hnsw_free_results(void** idsArray, void** distsArray, int count);
/* BatchRangeQuery with one allocation per output instead of two per query: offsets[count + 1] (caller-owned) delimit
 * the queries' slices of *outIds / *outDists, each slice ascending by distance.  Release with hnsw_free_flat. */
HNSW_API int hnsw_range_query_flat(hnsw_handle handle, const float* vectors, int count, int dim, float range, long long* offsets,
                                   int** outIds, float** outDists);
HNSW_API void hnsw_free_flat(int* ids, float* dists);

/* HNSWIndexExports.cs:219-273: mutate the pending parameters used by the NEXT hnsw_create
 * (HNSWParameters.cs:13-55).  Return 0, or -1 on error. */
HNSW_API int hnsw_set_collection_size(int collectionSize);
HNSW_API int hnsw_set_max_edges(int maxEdges);
HNSW_API int hnsw_set_max_candidates(int maxCandidates);
HNSW_API int hnsw_set_remove_max_candidates(int maxCandidates);
HNSW_API int hnsw_set_distribution_rate(float distRate);
HNSW_API int hnsw_set_random_seed(int seed);
HNSW_API int hnsw_set_min_nn(int minNN);
HNSW_API int hnsw_set_allow_removals(bool allowRemovals);

/* ------------------------------------------------------------------ Part 2: additions */

/* HNSWIndex.Count (HNSWIndex.cs:250). -1 on NULL handle. */
HNSW_API int hnsw_count(hnsw_handle handle);
/* dimension fixed by the first add / by the snapshot, -1 if none yet */
HNSW_API int hnsw_dim(hnsw_handle handle);

/* HNSWIndex.Serialize (HNSWIndex.cs:210-217): writes the protobuf-net snapshot layout.
 * packed != 0 writes repeated scalars packed (smaller/faster; both forms are accepted on read). */
HNSW_API int hnsw_serialize(hnsw_handle handle, const char* path, int packed);
/* HNSWIndex.Deserialize (HNSWIndex.cs:222-229).  NULL on error. */
HNSW_API hnsw_handle hnsw_deserialize(const char* distanceMetric, const char* path);

/* HNSWParameters is a live object in the C# API (KnnQuery reads parameters.MinNN on every call,
 * HNSWIndex.cs:115): change MinNN (= ef) of an existing index. */
HNSW_API int hnsw_index_set_min_nn(hnsw_handle handle, int minNN);
/* read back the parameters of an index: ints = {MaxEdges, MinNN, MaxCandidates, RemoveMaxCandidates,
 * CollectionSize, RandomSeed, AllowRemovals}, rate = DistributionRate */
HNSW_API int hnsw_index_get_params(hnsw_handle handle, int* ints7, double* rate);

/* KnnQuery / BatchKnnQuery with the reference's optional `layer` argument (HNSWIndex.cs:107,129): the search runs on
 * that layer (entry found by descending the layers above it).  -1 + last error when the layer does not exist. */
HNSW_API int hnsw_knn_query_layer(hnsw_handle handle, const float* vectors, int count, int dim, int k, int layer, int* outIds, float* outDists);
/* MultiLayerKnnQuery (HNSWIndex.cs:173-187), batched over queries: one search of width k per layer from
 * min(top layer, maxLayer) down to minLayer.  Returns the number of layers L = min(top, maxLayer) + 1 (0 for an empty
 * index) or -1; outIds / outDists are [count][L][k], each row the layer's ascending result WITHOUT its first entry (as
 * the reference), padded with -1 / NaN; layers below minLayer are all padding.  NULL outputs: query L only. */
HNSW_API int hnsw_multilayer_knn_query(hnsw_handle handle, const float* vectors, int count, int dim, int k, int maxLayer, int minLayer,
                                       int* outIds, float* outDists);
HNSW_API int hnsw_top_layer(hnsw_handle handle);   /* MaxLayer of the entry point, -1 when empty */
HNSW_API int hnsw_entry_point(hnsw_handle handle); /* EntryPointId, -1 when empty */

/* GetInfo (HNSWIndex.cs:192-196, HNSWInfo.cs): per-layer degree statistics.  Returns the number of layers; fills at
 * most maxLayers entries.  In-edge fields are 0 when AllowRemovals is false, as in the reference. */
typedef struct hnsw_layer_info {
    int layerId, nodesCount, maxOutEdges, minOutEdges, maxInEdges, minInEdges;
    double avgOutEdges, avgInEdges;
    int outEdgesMedian, inEdgesMedian;
} hnsw_layer_info;
HNSW_API int hnsw_get_info(hnsw_handle handle, hnsw_layer_info* out, int maxLayers);
/* GetConnectedComponentCounts (HNSWIndex.cs:202-205, GraphNavigator.cs:331-419): weakly connected components per layer. */
HNSW_API int hnsw_connected_components(hnsw_handle handle, int* counts, int maxLayers);

/* One process per GPU: select the CUDA device used by indexes created afterwards on this thread. */
HNSW_API int hnsw_set_device(int device);
HNSW_API int hnsw_device_count(void);

/* Same as hnsw_knn_query with queries/outputs already resident in the index's GPU memory
 * (dense [count][dim] fp32 -> [count][k]).  Synchronises before returning. */
HNSW_API int hnsw_knn_query_device(hnsw_handle handle, const float* dQueries, int count, int dim, int k, int* dOutIds, float* dOutDists);
/* Variant for serving loops: the batch is only queued (batches queued back to back keep the GPU busy across host
 * jitter).  Consecutive batches alternate between the index's stream and an internal auxiliary stream so that they
 * overlap in their tails: give them distinct output buffers.  hnsw_synchronize waits (on the host) for everything queued
 * so far; hnsw_async_join only makes the index's stream wait for the auxiliary one (for callers that time or chain work
 * on that stream).  Every other call on the handle waits for queued batches first. */
HNSW_API int hnsw_knn_query_device_async(hnsw_handle handle, const float* dQueries, int count, int dim, int k, int* dOutIds, float* dOutDists);
HNSW_API int hnsw_synchronize(hnsw_handle handle);
HNSW_API int hnsw_async_join(hnsw_handle handle);
/* Run the index's work on a caller-owned CUDA stream (cudaStream_t as void*); NULL = own stream. */
HNSW_API int hnsw_set_stream(hnsw_handle handle, void* cudaStream);

/* Exact brute-force top-k over the indexed vectors with the index's metric (recall ground truth). */
HNSW_API int hnsw_exact_knn(hnsw_handle handle, const float* queries, int count, int dim, int k, int* outIds, float* outDists);

/* The same answers (ids and distance bits) through the tensor cores: a 3xTF32 split-precision GEMM ranks all pairs
 * approximately and keeps k + slack candidates per query, which are then evaluated with the metric's exact fp32
 * arithmetic; a query whose k-th exact distance is not provably below everything outside its candidates is answered by
 * the FFMA scan of hnsw_exact_knn instead.  *outFallbacks (may be null) = number of such queries. */
HNSW_API int hnsw_exact_knn_tf32(hnsw_handle handle, const float* queries, int count, int dim, int k, int* outIds, float* outDists,
                                 int* outFallbacks);
/* With hnsw_set_tuning(handle, "exact_tcgen05", 1) the GEMM of hnsw_exact_knn_tf32 runs as tcgen05.mma kind::tf32 with the
 * accumulator in tensor memory instead of mma.sync, for the shapes this returns 1 for (rows of at most 128 floats and a
 * k whose candidate lists fit in shared memory); other shapes keep using the mma.sync kernel.  -1 on error.
 * ("exact_tcgen05": 2 = warp-specialised tcgen05 kernel, the default; 1 = the earlier 4-warp tcgen05 kernel; 0 forces mma.sync.) */
HNSW_API int hnsw_exact_tcgen05_supported(int dim, int k);

/* Merge `parts` per-partition top-k results (device memory, [parts][count][k], rows ascending,
 * -1 padded) into [count][k] on the current device. */
HNSW_API int hnsw_merge_topk_device(const int* dInIds, const float* dInDists, int parts, int count, int k, int* dOutIds, float* dOutDists,
                                    void* cudaStream);

/* Partitioned index (one sub-graph per GPU, SURVEY 8e mode 2) without a collective in the data path: buffers shared
 * between the ranks of one node through CUDA IPC, and a search whose epilogue writes every query's local top-k (ids +
 * idOffset) into slot `rank` of EVERY peer's gather buffer over NVLink, then raises flags[peer][rank] = epoch.
 *   hnsw_ipc_alloc   cudaMalloc (zeroed) + cudaIpcGetMemHandle -> 64-byte handle to send to the other ranks
 *   hnsw_ipc_open    map a peer's buffer (cudaIpcOpenMemHandle); hnsw_ipc_close / hnsw_ipc_free undo them
 *   hnsw_knn_query_scatter      stream-ordered, does not synchronise; ids[p] / dists[p] are [world][count][k], flags[p] is
 *                               [world] unsigned; p = 0..world-1 includes this rank's own buffer
 *   hnsw_wait_and_merge_device  ONE kernel: waits on the device until all `world` flags of this rank reached `epoch` -- at most
 *                               `peer_timeout_ms` of wall time (hnsw_set_tuning, default 10000) --, merges its gather buffer into
 *                               [count][k], synchronises; -1 if a peer never arrived.  In that case every output row of the
 *                               batch is -1 / NaN (never stale data), later batches fail at once instead of waiting again, and
 *                               the exchange must be set up again on all ranks (fresh buffers, epoch 0, a barrier): the late
 *                               peer may still write into a buffer parity the waiter has moved past.
 * Epochs must increase by one per query; alternate two gather buffers by epoch parity. */
#define HNSW_MAX_PEERS 16
typedef struct hnsw_scatter {
    int world, rank, idOffset;
    unsigned epoch;
    int* ids[HNSW_MAX_PEERS];
    float* dists[HNSW_MAX_PEERS];
    unsigned* flags[HNSW_MAX_PEERS];
} hnsw_scatter;
HNSW_API int hnsw_ipc_alloc(size_t bytes, void** devPtr, unsigned char* handle64);
HNSW_API int hnsw_ipc_open(const unsigned char* handle64, void** devPtr);
HNSW_API int hnsw_ipc_close(void* devPtr);
HNSW_API int hnsw_ipc_free(void* devPtr);
HNSW_API int hnsw_knn_query_scatter(hnsw_handle handle, const float* dQueries, int count, int dim, int k, const hnsw_scatter* sc);
HNSW_API int hnsw_wait_and_merge_device(hnsw_handle handle, const unsigned* dFlags, int world, unsigned epoch, const int* dInIds,
                                        const float* dInDists, int count, int k, int* dOutIds, float* dOutDists);
/* The same without synchronising (several query batches can be queued back to back; results stay on the device), and
 * the matching wait: hnsw_partition_sync returns -1 if a peer never delivered. */
HNSW_API int hnsw_wait_and_merge_device_async(hnsw_handle handle, const unsigned* dFlags, int world, unsigned epoch, const int* dInIds,
                                              const float* dInDists, int count, int k, int* dOutIds, float* dOutDists);
HNSW_API int hnsw_partition_sync(hnsw_handle handle);

/* Measurement hooks. */
typedef struct hnsw_stats {
    double last_search_ms;              /* CUDA-event time of the last search kernel */
    double last_build_ms;               /* CUDA-event time of the last hnsw_add (all batches) */
    unsigned long long ndist;           /* last search: distance evaluations (needs stats enabled) */
    unsigned long long nexp;            /* last search: rows expanded */
    unsigned long long row_words;       /* last search: sum of (degree + 1) over expansions */
    unsigned long long search_launches; /* kernels launched since creation */
    unsigned long long build_launches;
} hnsw_stats;
HNSW_API int hnsw_enable_stats(hnsw_handle handle, int on);
HNSW_API int hnsw_get_stats(hnsw_handle handle, hnsw_stats* out);
/* Construction profile since creation / the last reset, 16 words: {inserts, descent cycles, beam-search cycles,
 * heuristic cycles, search distance evaluations, heuristic evaluations, link-kernel warp cycles, overflow-prune
 * cycles, overflow prunes, prune evaluations, lock-wait cycles, select-kernel warp cycles, reverse links, 0...}.
 * Cycles are per-warp SM clock cycles summed over warps. */
HNSW_API int hnsw_get_build_profile(hnsw_handle handle, unsigned long long* out16, int reset);
/* Profile of the CTA-per-query wide-beam search kernel on the handle's device since the last reset, 16 words: {queries,
 * expansions, speculative evaluations issued, speculative evaluations used, then leader-warp SM cycles: descent, select +
 * filter, waiting for a speculation, waiting for an on-demand evaluation, insertion, beam loop total, 0...}. */
HNSW_API int hnsw_get_wide_profile(hnsw_handle handle, unsigned long long* out16, int reset);
/* tuning knobs: "vis_size_search", "vis_size_build", "blocks_per_sm", "build_batch_max", "build_batch_ratio",
 * "flags", "zero_copy" (1 = page-locked caller buffers of hnsw_knn_query are accessed in place by the kernel),
 * "wl_min_ef" / "wide_min_ef" (search width from which the non-shifting list / the CTA-per-query kernel are used),
 * "wide_warps", "wide_prefetch", "wide_max_small_batch", "wide_bitmap_min_ef", "wide_bitmap_max_nodes", "peer_timeout_ms",
 * "build_tail_fill" (1 = the link kernel of an insertion batch starts under the tail of its select kernel),
 * "vis_global_search" (1 = exact visited bitmap for searches over rows longer than 128 floats); the full list with
 * defaults and measurements is struct Tuning in csrc/kernels.hpp.  Unknown keys return -1. */
HNSW_API int hnsw_set_tuning(hnsw_handle handle, const char* key, int value);

/* Copy the graph to host memory (tests / inspection).  levels[count]; rows0[count][stride0] (-1 padded).
 * Pass NULL pointers to query sizes only.  Returns count, or -1.  These are two-call interfaces (size, then copy): when
 * another thread may Add in between, use the _n forms, which copy at most maxNodes nodes and return the current count
 * (call again with larger buffers when it exceeds maxNodes). */
HNSW_API int hnsw_export_layer0(hnsw_handle handle, int* levels, int* rows0, int* stride0, int* entryPoint);
HNSW_API int hnsw_export_layer0_n(hnsw_handle handle, int maxNodes, int* levels, int* rows0, int* stride0, int* entryPoint);
/* rows of one upper layer: rows[count][strideU], all -1 for nodes below that layer */
HNSW_API int hnsw_export_layer(hnsw_handle handle, int layer, int* rows, int* stride);
HNSW_API int hnsw_export_layer_n(hnsw_handle handle, int layer, int maxNodes, int* rows, int* stride);

/* Host-only: read a snapshot file and write it again (packed or unpacked).  Exercises the codec
 * without a GPU; used by the CPU test-suite. */
HNSW_API int hnsw_snapshot_transcode(const char* inPath, const char* outPath, int packed);

/* library identification: returns "hnsw-b200 <version> sm_100a" */
HNSW_API const char* hnsw_version(void);

#ifdef __cplusplus
}
#endif
#endif /* HNSW_B200_H */
