// capi.cu -- the C ABI (include/hnsw_b200.h) over hnsw::Index.
// Mirrors bindings/HNSWIndex.Native/HNSWIndexExports.cs: process-global pending parameters that the
// next hnsw_create consumes and resets (:16,50-61), process-global last error (:11-12), -1 / NULL on error.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/hnsw_b200.h"
#include "engine.hpp"

using namespace hnsw;

namespace {
std::mutex g_mu;
std::string g_last_error;
Parameters g_pending;

void set_error(const std::string& what) {
    std::lock_guard<std::mutex> l(g_mu);
    g_last_error = what;
}
MetricKind parse_metric(const char* m) {
    const std::string s(m ? m : "");
    if (s == "sq_euclid") return SQ_EUCLID;
    if (s == "cosine") return COSINE;
    if (s == "ucosine") return UCOSINE;
    if (s == "euclid") return EUCLID;  // addition: EuclideanMetric.Compute of the C# library (the reference's native shim stops at three names)
    throw std::invalid_argument("Unsupported distance metric: " + s);  // HNSWIndexExports.cs:59
}
Index* get(hnsw_handle h) {
    if (!h) throw std::invalid_argument("Invalid handle");  // HNSWIndexExports.cs:20
    return reinterpret_cast<Index*>(h);
}
}  // namespace

#define HNSW_TRY try {
#define HNSW_CATCH(ret)                                          \
    }                                                            \
    catch (const std::exception& e) { set_error(e.what()); return ret; } \
    catch (...) { set_error("unknown error"); return ret; }

extern "C" {

int hnsw_get_last_error_utf8(char* buf, int bufLen) {
    std::lock_guard<std::mutex> l(g_mu);
    const int need = (int)g_last_error.size();
    if (bufLen > 0 && buf != nullptr) {
        const int n = std::min(need, std::max(0, bufLen - 1));
        std::memcpy(buf, g_last_error.data(), (size_t)n);
        buf[n] = 0;
    }
    return need;
}

hnsw_handle hnsw_create(const char* distanceMetric) {
    HNSW_TRY
    const MetricKind m = parse_metric(distanceMetric);
    Parameters p;
    {
        std::lock_guard<std::mutex> l(g_mu);
        p = g_pending;
    }
    Index* ix = new Index(m, p);
    {
        std::lock_guard<std::mutex> l(g_mu);
        g_pending = Parameters();  // reset parameters for next instance (HNSWIndexExports.cs:61)
    }
    return ix;
    HNSW_CATCH(nullptr)
}

void hnsw_free(hnsw_handle handle) {
    if (!handle) return;
    try { delete reinterpret_cast<Index*>(handle); } catch (const std::exception& e) { set_error(e.what()); }
}

int hnsw_add(hnsw_handle handle, const float* vectors, int count, int dim, int* outIds) {
    if (!handle) return 0;
    HNSW_TRY
    if (vectors == nullptr || count <= 0 || dim <= 0) return 0;
    if (outIds == nullptr) {  // (the reference would fault here; the ids are simply not reported)
        std::vector<int> ids((size_t)count);
        return get(handle)->add(vectors, count, dim, ids.data());
    }
    return get(handle)->add(vectors, count, dim, outIds);
    HNSW_CATCH(-1)
}

int hnsw_remove(hnsw_handle handle, const int* ids, int count) {
    if (!handle) return 0;
    HNSW_TRY
    if (ids == nullptr || count <= 0) return 0;
    get(handle)->remove(ids, count);  // throws when AllowRemovals is false (HNSWIndex.cs:85-86)
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_ids(hnsw_handle handle, int* outIds, int capacity) {
    HNSW_TRY
    const std::vector<int> a = get(handle)->active_ids();  // a copy taken under the index's lock
    for (int i = 0; i < (int)a.size() && i < capacity && outIds != nullptr; ++i) outIds[i] = a[i];
    return (int)a.size();
    HNSW_CATCH(-1)
}

int hnsw_knn_query(hnsw_handle handle, const float* vectors, int count, int dim, int k, int* outIds, float* outDists) {
    if (!handle) return 0;
    HNSW_TRY
    if (count > 0 && (vectors == nullptr || outIds == nullptr || outDists == nullptr)) throw std::invalid_argument("null buffer");
    get(handle)->knn_query(vectors, count, dim, k, outIds, outDists);
    return 0;
    HNSW_CATCH(-1)
}

int hnsw_knn_query_layer(hnsw_handle handle, const float* vectors, int count, int dim, int k, int layer, int* outIds, float* outDists) {
    if (!handle) return 0;
    HNSW_TRY
    if (count > 0 && (vectors == nullptr || outIds == nullptr || outDists == nullptr)) throw std::invalid_argument("null buffer");
    get(handle)->knn_query(vectors, count, dim, k, outIds, outDists, layer);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_multilayer_knn_query(hnsw_handle handle, const float* vectors, int count, int dim, int k, int maxLayer, int minLayer, int* outIds,
                              float* outDists) {
    HNSW_TRY
    return get(handle)->multilayer_knn_query(vectors, count, dim, k, maxLayer, minLayer, outIds, outDists);
    HNSW_CATCH(-1)
}
int hnsw_top_layer(hnsw_handle handle) { return handle ? reinterpret_cast<Index*>(handle)->top_layer() : -1; }
int hnsw_entry_point(hnsw_handle handle) { return handle ? reinterpret_cast<Index*>(handle)->entry_point() : -1; }
int hnsw_get_info(hnsw_handle handle, hnsw_layer_info* out, int maxLayers) {
    HNSW_TRY
    const auto info = get(handle)->get_info();
    for (int i = 0; i < (int)info.size() && i < maxLayers && out != nullptr; ++i) {
        const auto& s = info[i];
        out[i].layerId = s.LayerId; out[i].nodesCount = s.NodesCount;
        out[i].maxOutEdges = s.MaxOutEdges; out[i].minOutEdges = s.MinOutEdges;
        out[i].maxInEdges = s.MaxInEdges; out[i].minInEdges = s.MinInEdges;
        out[i].avgOutEdges = s.AvgOutEdges; out[i].avgInEdges = s.AvgInEdges;
        out[i].outEdgesMedian = s.OutEdgesMedian; out[i].inEdgesMedian = s.InEdgesMedian;
    }
    return (int)info.size();
    HNSW_CATCH(-1)
}
int hnsw_connected_components(hnsw_handle handle, int* counts, int maxLayers) {
    HNSW_TRY
    const auto c = get(handle)->connected_component_counts();
    for (int i = 0; i < (int)c.size() && i < maxLayers && counts != nullptr; ++i) counts[i] = c[i];
    return (int)c.size();
    HNSW_CATCH(-1)
}

// BatchRangeQuery with ONE allocation per output: offsets[count + 1] (caller-owned) delimit the queries' slices of
// *outIds / *outDists (malloc'ed here, released with hnsw_free_flat).
int hnsw_range_query_flat(hnsw_handle handle, const float* vectors, int count, int dim, float range, long long* offsets, int** outIds,
                          float** outDists) {
    if (!handle) return 0;
    HNSW_TRY
    if (count > 0 && (vectors == nullptr || offsets == nullptr || outIds == nullptr || outDists == nullptr)) throw std::invalid_argument("null buffer");
    *outIds = nullptr; *outDists = nullptr;
    std::vector<std::vector<std::pair<float, int>>> res;
    get(handle)->range_query(vectors, count, dim, range, res);
    long long total = 0;
    for (int i = 0; i < count; ++i) { offsets[i] = total; total += (long long)res[i].size(); }
    if (count >= 0) offsets[std::max(count, 0)] = total;
    int* ids = static_cast<int*>(std::malloc(sizeof(int) * (size_t)std::max<long long>(total, 1)));
    float* dists = static_cast<float*>(std::malloc(sizeof(float) * (size_t)std::max<long long>(total, 1)));
    if (!ids || !dists) { std::free(ids); std::free(dists); throw std::bad_alloc(); }
    for (int i = 0; i < count; ++i) {
        long long o = offsets[i];
        for (const auto& pr : res[i]) { ids[o] = pr.second; dists[o] = pr.first; ++o; }
    }
    *outIds = ids; *outDists = dists;
    return 0;
    HNSW_CATCH(-1)
}
void hnsw_free_flat(int* ids, float* dists) {
    std::free(ids);
    std::free(dists);
}

int hnsw_range_query(hnsw_handle handle, const float* vectors, int count, int dim, float range, void** outIds, void** outDists, int* counts) {
    if (!handle) return 0;
    try {
        if (count > 0 && (vectors == nullptr || outIds == nullptr || outDists == nullptr || counts == nullptr)) throw std::invalid_argument("null buffer");
        for (int i = 0; i < count; ++i) { outIds[i] = nullptr; outDists[i] = nullptr; counts[i] = 0; }
        std::vector<std::vector<std::pair<float, int>>> res;
        get(handle)->range_query(vectors, count, dim, range, res);
        for (int i = 0; i < count; ++i) {
            const int n = (int)res[i].size();
            // caller frees with hnsw_free_results (Marshal.AllocHGlobal in HNSWIndexExports.cs:172-173)
            int* ids = n > 0 ? (int*)std::malloc(sizeof(int) * (size_t)n) : nullptr;
            float* dists = n > 0 ? (float*)std::malloc(sizeof(float) * (size_t)n) : nullptr;
            if (n > 0 && (!ids || !dists)) { std::free(ids); std::free(dists); throw std::bad_alloc(); }
            for (int j = 0; j < n; ++j) { ids[j] = res[i][j].second; dists[j] = res[i][j].first; }
            outIds[i] = ids; outDists[i] = dists; counts[i] = n;
        }
        return 0;
    } catch (const std::exception& e) {
        // HNSWIndexExports.cs:187-196: free everything handed out so far, zero the counts
        for (int i = 0; i < count; ++i) {
            if (outIds && outIds[i]) { std::free(outIds[i]); outIds[i] = nullptr; }
            if (outDists && outDists[i]) { std::free(outDists[i]); outDists[i] = nullptr; }
            if (counts) counts[i] = 0;
        }
        set_error(e.what());
        return -1;
    }
}

void hnsw_free_results(void** idsArray, void** distsArray, int count) {
    if (idsArray == nullptr && distsArray == nullptr) return;
    for (int i = 0; i < count; ++i) {
        if (idsArray && idsArray[i]) { std::free(idsArray[i]); idsArray[i] = nullptr; }
        if (distsArray && distsArray[i]) { std::free(distsArray[i]); distsArray[i] = nullptr; }
    }
}

int hnsw_set_collection_size(int v) { std::lock_guard<std::mutex> l(g_mu); g_pending.CollectionSize = v; return 0; }
int hnsw_set_max_edges(int v) { std::lock_guard<std::mutex> l(g_mu); g_pending.MaxEdges = v; return 0; }
int hnsw_set_max_candidates(int v) { std::lock_guard<std::mutex> l(g_mu); g_pending.MaxCandidates = v; return 0; }
int hnsw_set_remove_max_candidates(int v) { std::lock_guard<std::mutex> l(g_mu); g_pending.RemoveMaxCandidates = v; return 0; }
int hnsw_set_distribution_rate(float v) { std::lock_guard<std::mutex> l(g_mu); g_pending.DistributionRate = (double)v; return 0; }
int hnsw_set_random_seed(int v) { std::lock_guard<std::mutex> l(g_mu); g_pending.RandomSeed = v; return 0; }
int hnsw_set_min_nn(int v) { std::lock_guard<std::mutex> l(g_mu); g_pending.MinNN = v; return 0; }
int hnsw_set_allow_removals(bool v) { std::lock_guard<std::mutex> l(g_mu); g_pending.AllowRemovals = v; return 0; }

// ------------------------------------------------------------------------------------ additions
int hnsw_count(hnsw_handle handle) { return handle ? reinterpret_cast<Index*>(handle)->count_locked() : -1; }
int hnsw_dim(hnsw_handle handle) { return handle ? reinterpret_cast<Index*>(handle)->dim() : -1; }

int hnsw_serialize(hnsw_handle handle, const char* path, int packed) {
    HNSW_TRY
    if (!path) throw std::invalid_argument("null path");
    get(handle)->serialize(path, packed != 0);
    return 0;
    HNSW_CATCH(-1)
}
hnsw_handle hnsw_deserialize(const char* distanceMetric, const char* path) {
    HNSW_TRY
    if (!path) throw std::invalid_argument("null path");
    return Index::from_snapshot(parse_metric(distanceMetric), path);
    HNSW_CATCH(nullptr)
}
int hnsw_index_set_min_nn(hnsw_handle handle, int minNN) {
    HNSW_TRY
    if (minNN < 1) throw std::invalid_argument("MinNN must be at least 1");
    get(handle)->set_min_nn(minNN);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_index_get_params(hnsw_handle handle, int* ints7, double* rate) {
    HNSW_TRY
    const Parameters p = get(handle)->params();
    if (ints7) {
        ints7[0] = p.MaxEdges; ints7[1] = p.MinNN; ints7[2] = p.MaxCandidates; ints7[3] = p.RemoveMaxCandidates;
        ints7[4] = p.CollectionSize; ints7[5] = p.RandomSeed; ints7[6] = p.AllowRemovals ? 1 : 0;
    }
    if (rate) *rate = p.DistributionRate;
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_set_device(int device) {
    HNSW_TRY
    HNSW_CUDA_CHECK(cudaSetDevice(device));
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int hnsw_knn_query_device(hnsw_handle handle, const float* dQueries, int count, int dim, int k, int* dOutIds, float* dOutDists) {
    HNSW_TRY
    get(handle)->knn_query_device(dQueries, count, dim, k, dOutIds, dOutDists);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_knn_query_device_async(hnsw_handle handle, const float* dQueries, int count, int dim, int k, int* dOutIds, float* dOutDists) {
    HNSW_TRY
    if (count > 0 && (dQueries == nullptr || dOutIds == nullptr || dOutDists == nullptr)) throw std::invalid_argument("null buffer");
    get(handle)->knn_query_device_async(dQueries, count, dim, k, dOutIds, dOutDists);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_async_join(hnsw_handle handle) {
    HNSW_TRY
    get(handle)->async_join();
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_synchronize(hnsw_handle handle) {
    HNSW_TRY
    get(handle)->synchronize();
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_set_stream(hnsw_handle handle, void* cudaStream) {
    HNSW_TRY
    get(handle)->set_stream(reinterpret_cast<cudaStream_t>(cudaStream));
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_exact_knn(hnsw_handle handle, const float* queries, int count, int dim, int k, int* outIds, float* outDists) {
    HNSW_TRY
    if (count > 0 && (queries == nullptr || outIds == nullptr || outDists == nullptr)) throw std::invalid_argument("null buffer");
    get(handle)->exact_knn(queries, count, dim, k, outIds, outDists);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_exact_knn_tf32(hnsw_handle handle, const float* queries, int count, int dim, int k, int* outIds, float* outDists, int* outFallbacks) {
    HNSW_TRY
    if (count > 0 && (queries == nullptr || outIds == nullptr || outDists == nullptr)) throw std::invalid_argument("null buffer");
    const int fb = get(handle)->exact_knn_tf32(queries, count, dim, k, outIds, outDists);
    if (outFallbacks) *outFallbacks = fb;
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_exact_tcgen05_supported(int dim, int k) {
    HNSW_TRY
    return exact_tcgen05_supported(dim, k) ? 1 : 0;
    HNSW_CATCH(-1)
}
int hnsw_merge_topk_device(const int* dInIds, const float* dInDists, int parts, int count, int k, int* dOutIds, float* dOutDists, void* cudaStream) {
    HNSW_TRY
    if (parts < 1 || k < 1) throw std::invalid_argument("parts and k must be positive");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cudaStream);
    launch_merge_topk(dInIds, dInDists, parts, count, k, dOutIds, dOutDists, s);
    HNSW_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
    HNSW_CATCH(-1)
}
// ---- partitioned index: CUDA IPC buffers + fused search/scatter (partition_kernels.cu)
int hnsw_ipc_alloc(size_t bytes, void** devPtr, unsigned char* handle64) {
    HNSW_TRY
    if (!devPtr || !handle64 || bytes == 0) throw std::invalid_argument("null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    HNSW_CUDA_CHECK(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); HNSW_CUDA_CHECK(e); }
    std::memcpy(handle64, &h, 64);
    *devPtr = p;
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_ipc_open(const unsigned char* handle64, void** devPtr) {
    HNSW_TRY
    if (!devPtr || !handle64) throw std::invalid_argument("null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    HNSW_CUDA_CHECK(cudaIpcOpenMemHandle(devPtr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_ipc_close(void* devPtr) {
    HNSW_TRY
    if (devPtr) HNSW_CUDA_CHECK(cudaIpcCloseMemHandle(devPtr));
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_ipc_free(void* devPtr) {
    HNSW_TRY
    if (devPtr) HNSW_CUDA_CHECK(cudaFree(devPtr));
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_knn_query_scatter(hnsw_handle handle, const float* dQueries, int count, int dim, int k, const hnsw_scatter* sc) {
    HNSW_TRY
    if (!sc || !dQueries) throw std::invalid_argument("null argument");
    if (sc->world < 1 || sc->world > HNSW_MAX_PEERS || sc->rank < 0 || sc->rank >= sc->world) throw std::invalid_argument("bad world / rank");
    ScatterArgs a{};
    a.world = sc->world; a.rank = sc->rank; a.id_offset = sc->idOffset; a.epoch = sc->epoch;
    for (int p = 0; p < sc->world; ++p) {
        if (!sc->ids[p] || !sc->dists[p] || !sc->flags[p]) throw std::invalid_argument("null peer pointer");
        a.ids[p] = sc->ids[p]; a.dists[p] = sc->dists[p]; a.flags[p] = sc->flags[p];
    }
    get(handle)->knn_query_scatter(dQueries, count, dim, k, a);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_wait_and_merge_device(hnsw_handle handle, const unsigned* dFlags, int world, unsigned epoch, const int* dInIds, const float* dInDists,
                               int count, int k, int* dOutIds, float* dOutDists) {
    HNSW_TRY
    if (world < 1 || k < 1) throw std::invalid_argument("world and k must be positive");
    get(handle)->wait_and_merge(dFlags, world, epoch, dInIds, dInDists, count, k, dOutIds, dOutDists, true);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_wait_and_merge_device_async(hnsw_handle handle, const unsigned* dFlags, int world, unsigned epoch, const int* dInIds,
                                     const float* dInDists, int count, int k, int* dOutIds, float* dOutDists) {
    HNSW_TRY
    if (world < 1 || k < 1) throw std::invalid_argument("world and k must be positive");
    get(handle)->wait_and_merge(dFlags, world, epoch, dInIds, dInDists, count, k, dOutIds, dOutDists, false);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_partition_sync(hnsw_handle handle) {
    HNSW_TRY
    get(handle)->partition_sync();
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_enable_stats(hnsw_handle handle, int on) {
    HNSW_TRY
    get(handle)->enable_stats(on != 0);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_get_stats(hnsw_handle handle, hnsw_stats* out) {
    HNSW_TRY
    if (!out) throw std::invalid_argument("null stats");
    const IndexStats s = get(handle)->stats();
    out->last_search_ms = s.last_search_ms; out->last_build_ms = s.last_build_ms;
    out->ndist = s.ndist; out->nexp = s.nexp; out->row_words = s.row_words;
    out->search_launches = s.search_launches; out->build_launches = s.build_launches;
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_get_build_profile(hnsw_handle handle, unsigned long long* out16, int reset) {
    HNSW_TRY
    if (!out16) throw std::invalid_argument("null output");
    get(handle)->build_profile(out16, reset != 0);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_get_wide_profile(hnsw_handle handle, unsigned long long* out16, int reset) {
    HNSW_TRY
    if (!out16) throw std::invalid_argument("null output");
    HNSW_CUDA_CHECK(cudaSetDevice(get(handle)->device()));
    wide_profile(out16, reset != 0);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_set_tuning(hnsw_handle handle, const char* key, int value) {
    HNSW_TRY
    get(handle)->set_tuning(key ? key : "", value);
    return 0;
    HNSW_CATCH(-1)
}
int hnsw_export_layer0_n(hnsw_handle handle, int maxNodes, int* levels, int* rows0, int* stride0, int* entryPoint) {
    HNSW_TRY
    Index* ix = get(handle);
    HostGraph g;
    ix->export_graph_locked(g, false);
    if (stride0) *stride0 = g.stride0;
    if (entryPoint) *entryPoint = g.entry;
    const size_t n = (size_t)std::max(0, std::min(g.length, maxNodes));
    if (levels) std::memcpy(levels, g.levels.data(), n * sizeof(int));
    if (rows0) std::memcpy(rows0, g.adj0.data(), n * (size_t)g.stride0 * sizeof(int));
    return g.length;
    HNSW_CATCH(-1)
}
int hnsw_export_layer_n(hnsw_handle handle, int layer, int maxNodes, int* rows, int* stride) {
    HNSW_TRY
    if (layer < 1) throw std::invalid_argument("use hnsw_export_layer0 for layer 0");
    Index* ix = get(handle);
    HostGraph g;
    ix->export_graph_locked(g, false);
    if (stride) *stride = g.strideU;
    const int n = std::max(0, std::min(g.length, maxNodes));
    if (rows)
        for (int i = 0; i < n; ++i) {
            int* dst = rows + (size_t)i * g.strideU;
            if (g.levels[i] >= layer) std::memcpy(dst, g.adjU.data() + ((size_t)g.upper_off[i] + (layer - 1)) * g.strideU, (size_t)g.strideU * sizeof(int));
            else for (int j = 0; j < g.strideU; ++j) dst[j] = -1;
        }
    return g.length;
    HNSW_CATCH(-1)
}
int hnsw_export_layer0(hnsw_handle handle, int* levels, int* rows0, int* stride0, int* entryPoint) {
    return hnsw_export_layer0_n(handle, 0x7FFFFFFF, levels, rows0, stride0, entryPoint);
}
int hnsw_export_layer(hnsw_handle handle, int layer, int* rows, int* stride) { return hnsw_export_layer_n(handle, layer, 0x7FFFFFFF, rows, stride); }
int hnsw_snapshot_transcode(const char* inPath, const char* outPath, int packed) {
    HNSW_TRY
    if (!inPath || !outPath) throw std::invalid_argument("null path");
    HostGraph g;
    read_snapshot(inPath, g);
    write_snapshot(g, outPath, packed != 0);
    return g.length;
    HNSW_CATCH(-1)
}
const char* hnsw_version(void) { return "hnsw-b200 0.1.0 sm_100a"; }

}  // extern "C"
