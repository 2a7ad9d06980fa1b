// engine.hpp -- host side of the B200 HNSW engine: the object behind an hnsw_* handle.
//
// Mirrors HNSWIndex<float[], float> (HNSWIndex.cs) for the accelerated path:
//   ctor(parameters)            HNSWIndex.cs:20-30     -> Index(metric, params)
//   Add(List<float[]>)          HNSWIndex.cs:70-78     -> add()
//   BatchKnnQuery               HNSWIndex.cs:129-137   -> knn_query()
//   Serialize / Deserialize     HNSWIndex.cs:210-229   -> serialize() / from_snapshot()
//   Count                       HNSWIndex.cs:250       -> count()
// GraphData's arrays of objects (GraphData.cs:17-25) become flat device arrays
// (common.cuh GraphView); AddItem's slot allocation, level draw and x2 growth
// (GraphData.cs:79-118, 211-219) stay on the host because they are sequential
// scalar bookkeeping.  All distance work runs on the GPU; there is no CPU fallback.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.hpp"
#include "snapshot.hpp"

namespace hnsw {

// System.Random(seed) of the .NET BCL (seeded legacy algorithm), needed so that levels drawn by
// GraphData.GetRandomLayer (GraphData.cs:211-219) come out identical for a given RandomSeed.
class DotNetRandom {
public:
    explicit DotNetRandom(int seed);
    float NextSingle();
private:
    int InternalSample();
    int seed_array_[56];
    int inext_, inextp_;
};

struct IndexStats {
    double last_search_ms = 0.0;    // device time of the last search kernel (CUDA events)
    double last_build_ms = 0.0;     // device time of the last add() (all batches)
    unsigned long long ndist = 0;   // distance evaluations of the last search (if stats enabled)
    unsigned long long nexp = 0;
    unsigned long long row_words = 0;
    unsigned long long search_launches = 0, build_launches = 0;  // kernels launched since creation
    unsigned long long remove_launches = 0;
};

class Index {
public:
    Index(MetricKind metric, const Parameters& params);
    ~Index();
    Index(const Index&) = delete;
    Index& operator=(const Index&) = delete;

    static Index* from_snapshot(MetricKind metric, const std::string& path);

    int add(const float* vectors, int count, int dim, int* out_ids);
    // HNSWIndex.Remove(List<int>) (HNSWIndex.cs:83-100); throws when AllowRemovals is false
    void remove(const int* ids, int count);
    // HNSWIndex.Ids() (HNSWIndex.cs:242-245): active ids in ActiveSet order
    std::vector<int> active_ids() { std::shared_lock<std::shared_mutex> l(mu_); return active_dense_; }
    // `layer`: HNSWIndex.cs:107/129 optional argument (0 = base layer)
    void knn_query(const float* queries, int count, int dim, int k, int* out_ids, float* out_dists, int layer = 0);
    // MultiLayerKnnQuery (HNSWIndex.cs:173-187), batched; returns the layer count, outputs [count][layers][k]
    int multilayer_knn_query(const float* queries, int count, int dim, int k, int max_layer, int min_layer, int* out_ids, float* out_dists);
    int top_layer() { std::shared_lock<std::shared_mutex> l(mu_); return top_; }
    int entry_point() { std::shared_lock<std::shared_mutex> l(mu_); return entry_; }
    // queries / outputs already resident in device memory of this index's GPU
    void knn_query_device(const float* d_queries, int count, int dim, int k, int* d_ids, float* d_dists);
    // the same, only queued on the index's stream (several batches back to back); synchronize() waits for them
    // Consecutive async batches alternate between the index's stream and an auxiliary one, so the next batch's blocks
    // take the SM slots the previous batch's persistent blocks free up in its tail (about 10 % more throughput).
    void knn_query_device_async(const float* d_queries, int count, int dim, int k, int* d_ids, float* d_dists);
    void synchronize();
    void async_join();  // stream-ordered: the index's stream waits for the auxiliary one (no host wait)
    // partitioned index: stream-ordered search + scatter of the local top-k to all peers (no synchronisation), and the
    // matching wait-for-peers + merge of this rank's gather buffer (synchronises; throws if a peer never arrives)
    void knn_query_scatter(const float* d_queries, int count, int dim, int k, const ScatterArgs& sc);
    void wait_and_merge(const unsigned* d_flags, int world, unsigned epoch, const int* d_in_ids, const float* d_in_dists, int count, int k,
                        int* d_out_ids, float* d_out_dists, bool sync);
    void partition_sync();
    void exact_knn(const float* queries, int count, int dim, int k, int* out_ids, float* out_dists);
    // the same answers through the 3xTF32 tensor-core filter + bit-exact re-ranking (exact_tf32_kernels.cu); returns the
    // number of queries whose certificate failed and that were answered by the FFMA scan instead (all of them when
    // dim / k do not fit the variant)
    int exact_knn_tf32(const float* queries, int count, int dim, int k, int* out_ids, float* out_dists);
    // BatchRangeQuery (HNSWIndex.cs:160-168): per query, (id, distance) ascending by distance
    void range_query(const float* queries, int count, int dim, float range, std::vector<std::vector<std::pair<float, int>>>& out);
    void serialize(const std::string& path, bool packed);

    int count() const { return (int)active_dense_.size(); }  // GraphData.Count: active (not removed) items (callers hold mu_)
    int count_locked() { std::shared_lock<std::shared_mutex> l(mu_); return (int)active_dense_.size(); }
    int length() const { return length_; }                     // GraphData.Length: slots in use
    int dim() const { return dim_; }
    Parameters params() { std::shared_lock<std::shared_mutex> l(mu_); return params_; }
    void set_min_nn(int v) { Exclusive l(mu_); params_.MinNN = v; }
    void set_tuning(const std::string& key, int value);
    // run on a caller-owned CUDA stream (e.g. torch's current stream) instead of the index's own
    void set_stream(cudaStream_t s);
    void enable_stats(bool on) { Exclusive l(mu_); collect_stats_ = on; }
    IndexStats stats() { std::shared_lock<std::shared_mutex> g(mu_); std::lock_guard<std::mutex> l(stats_mu_); return stats_; }
    // cycle / work counters of the construction kernels since creation or the last reset (build_kernels.cu BuildProf)
    void build_profile(unsigned long long out[16], bool reset);
    void export_graph_locked(HostGraph& g, bool with_vectors = true);  // device -> host copy of the whole index

    // HNSWInfo.LayerInfo (HNSWInfo.cs:19-53): degree statistics of one layer, computed from the exported adjacency
    struct LayerInfo {
        int LayerId, NodesCount, MaxOutEdges, MinOutEdges, MaxInEdges, MinInEdges;
        double AvgOutEdges, AvgInEdges;
        int OutEdgesMedian, InEdgesMedian;
    };
    std::vector<LayerInfo> get_info();                      // HNSWIndex.GetInfo, HNSWIndex.cs:192-196
    std::vector<int> connected_component_counts();          // HNSWIndex.GetConnectedComponentCounts, GraphNavigator.cs:331-419
    int device() const { return device_; }

private:
    void ensure_dim(int dim);
    void check_usable() const;   // throws when an earlier Add failed after its commit point
    void add_committed(const float* vectors, int count, int dim, const std::vector<int>& lv, const std::vector<int>& ids_new,
                       const std::vector<int>& src_row, int nfresh, int nreused);
    std::string poisoned_;
    void export_graph(HostGraph& g, bool with_vectors = true);  // (caller holds mu_ exclusively)
    void exact_scan_locked(const float* queries, int count, int k, int* out_ids, float* out_dists);
    void ensure_capacity(int needed);
    void ensure_upper_rows(int needed);
    void ensure_query_scratch(int count, int k);
    void ensure_stats(int count);
    void upload_graph(const HostGraph& g);
    GraphView view() const;
    void fetch_query_stats(int count);
    void active_add(int id);
    void active_remove(int id);
    void remove_chunk(const int* ids, int n);
    void replace_entry_point(const std::vector<unsigned char>& in_chunk, int chunk_size);
    std::vector<int> read_row(int id, int layer);

    MetricKind metric_;
    Parameters params_;
    Tuning tuning_;
    DotNetRandom rng_;
    int device_ = 0;
    int dim_ = -1, pitch_ = 0;
    int stride0_, strideU_;
    int length_ = 0, capacity_ = 0, dev_capacity_ = 0;
    int entry_ = -1, top_ = -1;
    bool collect_stats_ = false;
    bool zero_copy_ = true;        // page-locked caller buffers are read / written in place by the kernel
    IndexStats stats_;
    int build_batch_max_ = 16384;  // new nodes per build launch (upper bound; 24576 is 5 % faster at 1M but costs 0.2 pt of recall@10 at 400k)
    int build_batch_ratio_ = 16;   // ... and at most 1/ratio of the nodes already linked
    int build_ramp_ratio_ = 4;     // ... 1/this while fewer than build_ramp_until_ nodes are linked: until a batch fills the GPU
    int build_ramp_until_ = 40000; // every batch costs one insertion's latency whatever its size (measured: 1M build 976 -> 854 ms
                                   // kernel time, recall@10 unchanged); applied only when the call
                                   // leaves an index of at least 4 x build_ramp_until_ nodes

    std::vector<int> levels_, upper_off_;
    // removal state (GraphData.RemovedIndexes, ActiveSet, Node.IsRemoved)
    std::vector<unsigned char> removed_;
    std::vector<int> active_dense_, active_sparse_, removed_stack_;
    int remove_batch_max_ = 4096, remove_batch_ratio_ = 16;
    int* d_rm_index_ = nullptr;   // [capacity] see GraphView::rm_index
    int* d_new_ids_ = nullptr;    // ids of the nodes of the running add()
    size_t new_ids_cap_ = 0;
    int* d_ready_ = nullptr;      // per node: the epoch of the insertion batch that wrote its rows (tail fill, build_kernels.cu)
    int ready_cap_ = 0;
    int build_epoch_ = 0;
    int upper_rows_ = 0, upper_rows_cap_ = 0;

    float* d_vectors_ = nullptr;
    double* d_norms_ = nullptr;
    int* d_adj0_ = nullptr;
    int* d_upper_off_ = nullptr;
    int* d_adjU_ = nullptr;
    int* d_locks_ = nullptr;
    int* d_levels_ = nullptr;
    unsigned* d_counters_ = nullptr;  // [0..2] build, [3] search, [4..5] remove, [6..7] scatter, [8] second query stream
    u64* d_build_prof_ = nullptr;     // [16]
    void *rm_small_ = nullptr, *rm_aff_ = nullptr;   // remove() scratch (grow-only)
    size_t rm_small_bytes_ = 0, rm_aff_bytes_ = 0;
    // range_query scratch (grow-only): per-query key slices, visited tables, small arrays, packed results
    void *rq_keys_ = nullptr, *rq_hash_ = nullptr, *rq_small_ = nullptr, *rq_dense_ = nullptr;
    size_t rq_keys_bytes_ = 0, rq_hash_bytes_ = 0, rq_small_bytes_ = 0, rq_dense_bytes_ = 0;

    float* d_q_ = nullptr;
    int* d_ids_ = nullptr;
    float* d_dists_ = nullptr;
    QueryStats* d_stats_ = nullptr;
    size_t q_cap_ = 0, out_cap_ = 0, stats_cap_ = 0;

    cudaStream_t stream_ = nullptr, own_stream_ = nullptr, aux_stream_ = nullptr;
    cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr;
    unsigned async_seq_ = 0;
    std::atomic<bool> aux_dirty_{false};   // work queued on aux_stream_ that the primary stream has not been joined with
    void join_aux_locked(bool host_wait);
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    // Locking ("thread-safe per operation type", README.md:64-65 of the reference): mu_ is held SHARED by hnsw_knn_query -- any
    // number of host threads search at once, each with its own stream and scratch (QueryCtx) -- and EXCLUSIVELY by everything
    // that changes the index, its tuning or the primary scratch (Add, Remove, Serialize, the device-pointer entry points, ...).
    std::shared_mutex mu_;
    typedef std::unique_lock<std::shared_mutex> Exclusive;
    struct QueryCtx {
        cudaStream_t stream = nullptr, aux = nullptr;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
        float* d_q = nullptr;
        int* d_ids = nullptr;
        float* d_dists = nullptr;
        QueryStats* d_stats = nullptr;
        unsigned* d_counters = nullptr;  // [16]
        unsigned char* h_stage = nullptr;  // page-locked, device-mapped scratch for small calls (kSmallCallBytes): no copy commands at all
        unsigned char* d_stage = nullptr;  // ... its device alias
        size_t q_cap = 0, out_cap = 0, stats_cap = 0;
        bool busy = false, created = false;
    };
    static constexpr int kQueryContexts = 4;
    static constexpr size_t kSmallCallBytes = 64 * 1024;  // queries + results of a call that goes through the mapped scratch
    QueryCtx qctx_[kQueryContexts];  // [0] runs on stream_ (set_stream applies), the others on their own streams
    std::mutex ctx_mu_, stats_mu_;
    std::condition_variable ctx_cv_;
    int acquire_ctx();
    void release_ctx(int i);
    void prepare_ctx(QueryCtx& c, int count, int k, bool want_stats);
};

}  // namespace hnsw
