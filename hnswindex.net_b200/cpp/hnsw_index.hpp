// hnsw_index.hpp -- header-only C++ host over the C ABI (include/hnsw_b200.h).
// Mirrors HNSWIndex<float[], float> of the reference (src/HNSWIndex/HNSWIndex.cs) for the accelerated
// path, for C++ applications that link or dlopen HNSWIndex.Native.so.  Throws std::runtime_error with
// the library's last-error text where the C# class throws.
#pragma once
#include <cmath>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hnsw_b200.h"

namespace hnswindex {

struct HNSWParameters {  // HNSWParameters.cs:13-55
    int MaxEdges = 16;
    double DistributionRate = 1.0 / std::log(16.0);
    int MinNN = 5;
    int MaxCandidates = 100;
    int RemoveMaxCandidates = 100;
    int CollectionSize = 65536;
    int RandomSeed = 31337;
    bool AllowRemovals = true;
};

struct KNNResult {  // KNNResult.cs (Label omitted: the caller owns the vectors)
    int Id;
    float Distance;
};

inline std::string last_error() {
    const int n = hnsw_get_last_error_utf8(nullptr, 0);
    if (n <= 0) return {};
    std::string s((size_t)n + 1, '\0');
    hnsw_get_last_error_utf8(&s[0], n + 1);
    s.resize((size_t)n);
    return s;
}

class HNSWIndex {
public:
    // metric: "sq_euclid" | "cosine" | "ucosine" (HNSWIndexExports.cs:47-57)
    explicit HNSWIndex(const std::string& metric, const HNSWParameters& p = HNSWParameters()) : params_(p) {
        static std::mutex create_mu;  // pending parameters are process-global
        std::lock_guard<std::mutex> l(create_mu);
        hnsw_set_max_edges(p.MaxEdges);
        hnsw_set_distribution_rate((float)p.DistributionRate);
        hnsw_set_min_nn(p.MinNN);
        hnsw_set_max_candidates(p.MaxCandidates);
        hnsw_set_remove_max_candidates(p.RemoveMaxCandidates);
        hnsw_set_collection_size(p.CollectionSize);
        hnsw_set_random_seed(p.RandomSeed);
        hnsw_set_allow_removals(p.AllowRemovals);
        h_ = hnsw_create(metric.c_str());
        if (!h_) throw std::runtime_error("hnsw_create failed: " + last_error());
    }
    ~HNSWIndex() { hnsw_free(h_); }
    HNSWIndex(const HNSWIndex&) = delete;
    HNSWIndex& operator=(const HNSWIndex&) = delete;
    HNSWIndex(HNSWIndex&& o) noexcept : h_(o.h_), params_(o.params_) { o.h_ = nullptr; }

    // HNSWIndex.Add(List) (HNSWIndex.cs:70-78); vectors row-major [count][dim]
    std::vector<int> Add(const float* vectors, int count, int dim) {
        std::vector<int> ids((size_t)count);
        if (hnsw_add(h_, vectors, count, dim, ids.data()) < 0) throw std::runtime_error(last_error());
        return ids;
    }
    // HNSWIndex.BatchKnnQuery (HNSWIndex.cs:129-137)
    std::vector<std::vector<KNNResult>> BatchKnnQuery(const float* queries, int count, int dim, int k) {
        std::vector<int> ids((size_t)count * k);
        std::vector<float> d((size_t)count * k);
        if (hnsw_knn_query(h_, queries, count, dim, k, ids.data(), d.data()) < 0) throw std::runtime_error(last_error());
        std::vector<std::vector<KNNResult>> out((size_t)count);
        for (int i = 0; i < count; ++i)
            for (int j = 0; j < k && ids[(size_t)i * k + j] >= 0; ++j) out[i].push_back({ids[(size_t)i * k + j], d[(size_t)i * k + j]});
        return out;
    }
    std::vector<KNNResult> KnnQuery(const float* query, int dim, int k) { return BatchKnnQuery(query, 1, dim, k)[0]; }
    // BatchKnnQuery(..., layer) (HNSWIndex.cs:129-137)
    std::vector<std::vector<KNNResult>> BatchKnnQuery(const float* queries, int count, int dim, int k, int layer) {
        std::vector<int> ids((size_t)count * k);
        std::vector<float> d((size_t)count * k);
        if (hnsw_knn_query_layer(h_, queries, count, dim, k, layer, ids.data(), d.data()) < 0) throw std::runtime_error(last_error());
        std::vector<std::vector<KNNResult>> out((size_t)count);
        for (int i = 0; i < count; ++i)
            for (int j = 0; j < k && ids[(size_t)i * k + j] >= 0; ++j) out[i].push_back({ids[(size_t)i * k + j], d[(size_t)i * k + j]});
        return out;
    }
    // HNSWIndex.MultiLayerKnnQuery (HNSWIndex.cs:173-187): result[layer], first entry of each layer's list dropped as in the reference
    std::vector<std::vector<KNNResult>> MultiLayerKnnQuery(const float* query, int dim, int k, int maxLayer = 2147483647, int minLayer = 0) {
        const int layers = hnsw_multilayer_knn_query(h_, query, 0, dim, k, maxLayer, minLayer, nullptr, nullptr);
        if (layers < 0) throw std::runtime_error(last_error());
        std::vector<std::vector<KNNResult>> out((size_t)layers);
        if (layers == 0) return out;
        std::vector<int> ids((size_t)layers * k);
        std::vector<float> d((size_t)layers * k);
        if (hnsw_multilayer_knn_query(h_, query, 1, dim, k, maxLayer, minLayer, ids.data(), d.data()) < 0) throw std::runtime_error(last_error());
        for (int l = 0; l < layers; ++l)
            for (int j = 0; j < k && ids[(size_t)l * k + j] >= 0; ++j) out[l].push_back({ids[(size_t)l * k + j], d[(size_t)l * k + j]});
        return out;
    }
    // HNSWIndex.BatchRangeQuery (HNSWIndex.cs:160-168)
    std::vector<std::vector<KNNResult>> BatchRangeQuery(const float* queries, int count, int dim, float range) {
        std::vector<void*> pi((size_t)count, nullptr), pd((size_t)count, nullptr);
        std::vector<int> counts((size_t)count, 0);
        if (hnsw_range_query(h_, queries, count, dim, range, pi.data(), pd.data(), counts.data()) < 0) throw std::runtime_error(last_error());
        std::vector<std::vector<KNNResult>> out((size_t)count);
        for (int i = 0; i < count; ++i)
            for (int j = 0; j < counts[i]; ++j) out[i].push_back({static_cast<int*>(pi[i])[j], static_cast<float*>(pd[i])[j]});
        hnsw_free_results(pi.data(), pd.data(), count);
        return out;
    }
    // HNSWIndex.Remove(List<int>) (HNSWIndex.cs:83-100)
    void Remove(const std::vector<int>& ids) {
        if (!ids.empty() && hnsw_remove(h_, ids.data(), (int)ids.size()) < 0) throw std::runtime_error(last_error());
    }
    // HNSWIndex.Ids() (HNSWIndex.cs:242-245)
    std::vector<int> Ids() const {
        const int n = hnsw_ids(h_, nullptr, 0);
        std::vector<int> ids((size_t)(n > 0 ? n : 0));
        if (n > 0) hnsw_ids(h_, ids.data(), n);
        return ids;
    }
    // HNSWIndex.GetInfo / GetConnectedComponentCounts (HNSWIndex.cs:192-205)
    std::vector<hnsw_layer_info> GetInfo() const {
        const int n = hnsw_get_info(h_, nullptr, 0);
        if (n < 0) throw std::runtime_error(last_error());
        std::vector<hnsw_layer_info> info((size_t)n);
        if (n > 0) hnsw_get_info(h_, info.data(), n);
        return info;
    }
    std::vector<int> GetConnectedComponentCounts() const {
        const int n = hnsw_connected_components(h_, nullptr, 0);
        if (n < 0) throw std::runtime_error(last_error());
        std::vector<int> c((size_t)n);
        if (n > 0) hnsw_connected_components(h_, c.data(), n);
        return c;
    }
    // HNSWIndex.Serialize / Deserialize (HNSWIndex.cs:210-229)
    void Serialize(const std::string& path) const {
        if (hnsw_serialize(h_, path.c_str(), 0) < 0) throw std::runtime_error(last_error());
    }
    static HNSWIndex Deserialize(const std::string& metric, const std::string& path) {
        hnsw_handle h = hnsw_deserialize(metric.c_str(), path.c_str());
        if (!h) throw std::runtime_error("hnsw_deserialize failed: " + last_error());
        return HNSWIndex(h);
    }
    int Count() const { return hnsw_count(h_); }  // HNSWIndex.cs:250
    void SetMinNN(int v) { if (hnsw_index_set_min_nn(h_, v) < 0) throw std::runtime_error(last_error()); }
    hnsw_handle native_handle() const { return h_; }

private:
    explicit HNSWIndex(hnsw_handle h) : h_(h) {}
    hnsw_handle h_ = nullptr;
    HNSWParameters params_;
};

}  // namespace hnswindex
