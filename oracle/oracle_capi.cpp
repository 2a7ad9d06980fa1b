// =============================================================================
// oracle/oracle_capi.cpp -- TEST INFRASTRUCTURE ONLY (see hnsw_oracle.hpp).
// extern "C" surface over the CPU restatement so tests/ and bench.py's
// cpu_baseline leg can drive it through ctypes.  Symbols are prefixed orc_ so
// they can never be mistaken for (or collide with) the product's hnsw_* ABI.
// =============================================================================
#include <chrono>
#include <cstring>
#include <string>

#include "hnsw_oracle.hpp"
#include "snapshot_codec.hpp"

using namespace hnsw_oracle;

static thread_local std::string g_err;

static Metric parse_metric(const char* m) {
    std::string s(m ? m : "");
    if (s == "sq_euclid") return SQ_EUCLID;
    if (s == "cosine") return COSINE;
    if (s == "ucosine") return UCOSINE;
    if (s == "euclid") return EUCLID;
    throw std::runtime_error("Unsupported distance metric: " + s);
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
int orc_have_avx() { return ORACLE_HAVE_AVX; }

// ---- .NET System.Random ------------------------------------------------------
void orc_dotnet_next(int seed, int n, int* out) {
    DotNetRandom r(seed);
    for (int i = 0; i < n; i++) out[i] = r.Next();
}
void orc_dotnet_next_single(int seed, int n, float* out) {
    DotNetRandom r(seed);
    for (int i = 0; i < n; i++) out[i] = r.NextSingle();
}
// levels drawn as GraphData.GetRandomLayer does (GraphData.cs:211-219)
void orc_draw_levels(int seed, double rate, int n, int* out) {
    Params p; p.RandomSeed = seed; p.DistributionRate = rate;
    Index ix(SQ_EUCLID, p);
    for (int i = 0; i < n; i++) out[i] = ix.GetRandomLayer();
}

// ---- metrics -------------------------------------------------------------------
// impl: 0 = scalar 8-lane emulation, 1 = AVX intrinsics (if built with them)
float orc_metric(const char* metric, const float* a, const float* b, int len, int impl) {
    Metric m = parse_metric(metric);
    return impl == 0 ? metric_scalar(m, a, b, len) : metric_fast(m, a, b, len);
}
void orc_metric_many(const char* metric, const float* a, const float* b, int len, int n, int impl, float* out) {
    Metric m = parse_metric(metric);
    for (int i = 0; i < n; i++)
        out[i] = impl == 0 ? metric_scalar(m, a + (size_t)i * len, b + (size_t)i * len, len) : metric_fast(m, a + (size_t)i * len, b + (size_t)i * len, len);
}
// pre-sqrt norm accumulator of one vector in the cosine metric's order
float orc_cos_norm(const float* a, int len) { return cos_acc_scalar(a, a, len).na; }

// ---- index -----------------------------------------------------------------------
void* orc_create(const char* metric, int maxEdges, double distRate, int minNN, int maxCandidates, int collectionSize, int seed, int allowRemovals) {
    try {
        Params p;
        p.MaxEdges = maxEdges; p.DistributionRate = distRate; p.MinNN = minNN; p.MaxCandidates = maxCandidates;
        p.CollectionSize = collectionSize; p.RandomSeed = seed; p.AllowRemovals = allowRemovals != 0;
        return new Index(parse_metric(metric), p);
    } catch (std::exception& e) { g_err = e.what(); return nullptr; }
}
void orc_free(void* h) { delete (Index*)h; }
void orc_use_scalar_metric(void* h, int on) { ((Index*)h)->use_scalar_metric = on != 0; }
void orc_set_min_nn(void* h, int v) { ((Index*)h)->p.MinNN = v; }
int orc_count(void* h) { return ((Index*)h)->Count(); }
int orc_dim(void* h) { return ((Index*)h)->dim; }
int orc_entry_point(void* h) { return ((Index*)h)->EntryPointId; }
int orc_capacity(void* h) { return ((Index*)h)->Capacity; }
int orc_length(void* h) { return ((Index*)h)->Length; }
void orc_get_params(void* h, int* ints /*7*/, double* rate) {
    Params& p = ((Index*)h)->p;
    ints[0] = p.MaxEdges; ints[1] = p.MinNN; ints[2] = p.MaxCandidates; ints[3] = p.RemoveMaxCandidates;
    ints[4] = p.CollectionSize; ints[5] = p.RandomSeed; ints[6] = p.AllowRemovals ? 1 : 0;
    *rate = p.DistributionRate;
}

// threads <= 1: serial Add in input order (deterministic); > 1: Parallel.For stand-in
int orc_add(void* h, const float* v, int count, int dim, int* outIds, int threads, uint64_t* counters /*3 or null*/) {
    try {
        Counters c;
        ((Index*)h)->AddBatch(v, count, dim, outIds, threads, counters ? &c : nullptr);
        if (counters) { counters[0] = c.ndist; counters[1] = c.nexp; counters[2] = c.row_words; }
        return count;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
int orc_add_frozen_batches(void* h, const float* v, int count, int dim, int* outIds, int ratio, int bmax) {
    try { ((Index*)h)->AddBatchFrozen(v, count, dim, outIds, ratio, bmax); return count; }
    catch (std::exception& e) { g_err = e.what(); return -1; }
}
// HNSWIndex.Remove(int) applied to ids in order (HNSWIndex.cs:83-100, single-threaded)
int orc_remove(void* h, const int* ids, int count) {
    try {
        Index* ix = (Index*)h;
        for (int i = 0; i < count; i++) ix->Remove(ids[i]);
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
int orc_active_ids(void* h, int* out, int cap) {
    Index* ix = (Index*)h;
    int n = (int)ix->ActiveIds.size();
    for (int i = 0; i < n && i < cap; i++) out[i] = ix->ActiveIds[i];
    return n;
}
int orc_is_removed(void* h, int id) { return ((Index*)h)->Nodes[id]->IsRemoved ? 1 : 0; }
int orc_knn_query(void* h, const float* q, int count, int dim, int k, int* outIds, float* outDists, int threads, uint64_t* counters /*3 or null*/) {
    try {
        Index* ix = (Index*)h;
        if (dim != ix->dim) throw std::runtime_error("dimension mismatch");
        Counters c;
        ix->BatchKnnQuery(q, count, k, outIds, outDists, threads, counters ? &c : nullptr);
        if (counters) { counters[0] = c.ndist; counters[1] = c.nexp; counters[2] = c.row_words; }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// BatchKnnQuery(queries, k, layer: L) -- HNSWIndex.cs:129-137
int orc_knn_query_layer(void* h, const float* q, int count, int dim, int k, int layer, int* outIds, float* outDists, int threads) {
    try {
        Index* ix = (Index*)h;
        if (dim != ix->dim) throw std::runtime_error("dimension mismatch");
        ix->BatchKnnQuery(q, count, k, outIds, outDists, threads, nullptr, layer);
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// MultiLayerKnnQuery (HNSWIndex.cs:173-187) per query; out arrays [count][layers][k] padded with -1 / NaN, a layer's
// row holds at most k-1 entries (the first one is dropped by the reference).  Returns the number of layers.
int orc_multilayer_knn_query(void* h, const float* q, int count, int dim, int k, int maxLayer, int minLayer, int layers, int* outIds, float* outDists) {
    try {
        Index* ix = (Index*)h;
        if (dim != ix->dim) throw std::runtime_error("dimension mismatch");
        int got = 0;
        for (int i = 0; i < count; i++) {
            auto res = ix->MultiLayerKnnQuery(q + (size_t)i * dim, k, maxLayer, minLayer);
            got = (int)res.size();
            if (outIds == nullptr) break;
            if (got != layers) throw std::runtime_error("layer count mismatch");
            for (int l = 0; l < layers; l++)
                for (int j = 0; j < k; j++) {
                    const size_t o = ((size_t)i * layers + l) * k + j;
                    if (j < (int)res[l].size()) { outIds[o] = res[l][j].Id; outDists[o] = res[l][j].Dist; }
                    else { outIds[o] = -1; outDists[o] = std::numeric_limits<float>::quiet_NaN(); }
                }
        }
        return got;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// per-query counters (ndist, nexp, row_words) for the roofline's algorithmic bytes
int orc_knn_query_counters(void* h, const float* q, int count, int dim, int k, uint64_t* perQuery /*count*3*/) {
    try {
        Index* ix = (Index*)h;
        if (dim != ix->dim) throw std::runtime_error("dimension mismatch");
        for (int i = 0; i < count; i++) {
            Counters c;
            ix->KnnQuery(q + (size_t)i * dim, k, &c);
            perQuery[3 * i] = c.ndist; perQuery[3 * i + 1] = c.nexp; perQuery[3 * i + 2] = c.row_words;
        }
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
// HNSWIndex.BatchRangeQuery (HNSWIndex.cs:160-168): results concatenated; offsets[count+1]
// two-call protocol: pass ids == null to get the total size
long long orc_range_query(void* h, const float* q, int count, int dim, float range, long long* offsets, int* ids, float* dists, int threads) {
    try {
        Index* ix = (Index*)h;
        if (dim != ix->dim) throw std::runtime_error("dimension mismatch");
        std::vector<std::vector<NodeDistance>> res(count);
        std::atomic<int> next{0};
        auto work = [&]() { for (;;) { int i = next.fetch_add(1); if (i >= count) break; res[i] = ix->RangeQuery(q + (size_t)i * dim, range); } };
        std::vector<std::thread> pool;
        for (int t = 0; t < std::max(1, threads); t++) pool.emplace_back(work);
        for (auto& th : pool) th.join();
        long long total = 0;
        for (int i = 0; i < count; i++) { offsets[i] = total; total += (long long)res[i].size(); }
        offsets[count] = total;
        if (ids && dists)
            for (int i = 0; i < count; i++)
                for (size_t j = 0; j < res[i].size(); j++) { ids[offsets[i] + j] = res[i][j].Id; dists[offsets[i] + j] = res[i][j].Dist; }
        return total;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}
int orc_exact_knn(void* h, const float* q, int count, int dim, int k, int* outIds, float* outDists, int threads) {
    try {
        Index* ix = (Index*)h;
        if (dim != ix->dim) throw std::runtime_error("dimension mismatch");
        ix->ExactKnn(q, count, k, outIds, outDists, threads);
        return 0;
    } catch (std::exception& e) { g_err = e.what(); return -1; }
}

// ---- graph inspection ----------------------------------------------------------------
int orc_node_level(void* h, int id) { return ((Index*)h)->Nodes[id]->MaxLayer(); }
int orc_node_row(void* h, int id, int layer, int which /*0 out,1 in*/, int* buf, int cap) {
    Node* n = ((Index*)h)->Nodes[id].get();
    auto& lists = which == 0 ? n->OutEdges : n->InEdges;
    if (layer >= (int)lists.size()) return -1;
    EdgeList& e = lists[layer];
    for (int i = 0; i < e.Count && i < cap; i++) buf[i] = e.Buffer[i];
    return e.Count;
}
// flat export of one layer's rows: rows[id*stride .. ] padded with -1, returns max degree seen
int orc_export_layer(void* h, int layer, int stride, int* rows /*Length*stride*/, int* levels /*Length or null*/) {
    Index* ix = (Index*)h;
    int maxdeg = 0;
    for (int i = 0; i < ix->Length; i++) {
        Node* n = ix->Nodes[i].get();
        if (levels) levels[i] = n->MaxLayer();
        int* r = rows + (size_t)i * stride;
        for (int j = 0; j < stride; j++) r[j] = -1;
        if (layer <= n->MaxLayer()) {
            EdgeList& e = n->OutEdges[layer];
            maxdeg = std::max(maxdeg, e.Count);
            for (int j = 0; j < e.Count && j < stride; j++) r[j] = e.Buffer[j];
        }
    }
    return maxdeg;
}
void orc_get_items(void* h, float* out) {
    Index* ix = (Index*)h;
    std::memcpy(out, ix->Items.data(), ix->Items.size() * sizeof(float));
}

// ---- snapshot ------------------------------------------------------------------------
int orc_serialize(void* h, const char* path, int packed) {
    try { serialize(*(Index*)h, path, packed != 0); return 0; }
    catch (std::exception& e) { g_err = e.what(); return -1; }
}
void* orc_deserialize(const char* metric, const char* path) {
    try { return deserialize(parse_metric(metric), path).release(); }
    catch (std::exception& e) { g_err = e.what(); return nullptr; }
}

}  // extern "C"
