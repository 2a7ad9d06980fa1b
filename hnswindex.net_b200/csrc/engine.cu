// engine.cu -- host side of the B200 HNSW engine (see engine.hpp).
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>

namespace hnsw {

// ---------------------------------------------------------------------------------------------
// .NET System.Random(seed): Knuth subtractive generator as published in dotnet/runtime
// (Random.CompatPrng).  NextSingle() = (float)(InternalSample() * (1.0 / int.MaxValue)),
// redrawn when it rounds to 1.0f.
// ---------------------------------------------------------------------------------------------
DotNetRandom::DotNetRandom(int seed) {
    const int kMax = std::numeric_limits<int>::max();
    const int sub = seed == std::numeric_limits<int>::min() ? kMax : std::abs(seed);
    int mj = 161803398 - sub;
    std::memset(seed_array_, 0, sizeof(seed_array_));
    seed_array_[55] = mj;
    int mk = 1, ii = 0;
    for (int i = 1; i < 55; ++i) {
        ii += 21;
        if (ii >= 55) ii -= 55;
        seed_array_[ii] = mk;
        mk = mj - mk;
        if (mk < 0) mk += kMax;
        mj = seed_array_[ii];
    }
    for (int round = 0; round < 4; ++round)
        for (int i = 1; i < 56; ++i) {
            int n = i + 30;
            if (n >= 55) n -= 55;
            seed_array_[i] = (int)((unsigned)seed_array_[i] - (unsigned)seed_array_[1 + n]);
            if (seed_array_[i] < 0) seed_array_[i] += kMax;
        }
    inext_ = 0;
    inextp_ = 21;
}
int DotNetRandom::InternalSample() {
    const int kMax = std::numeric_limits<int>::max();
    if (++inext_ >= 56) inext_ = 1;
    if (++inextp_ >= 56) inextp_ = 1;
    int r = (int)((unsigned)seed_array_[inext_] - (unsigned)seed_array_[inextp_]);
    if (r == kMax) r--;
    if (r < 0) r += kMax;
    seed_array_[inext_] = r;
    return r;
}
float DotNetRandom::NextSingle() {
    for (;;) {
        const float f = (float)(InternalSample() * (1.0 / std::numeric_limits<int>::max()));
        if (f < 1.0f) return f;
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static void grow_device(T*& p, size_t old_n, size_t new_n, int fill_byte, cudaStream_t s) {
    T* np = nullptr;
    HNSW_CUDA_CHECK(cudaMalloc(&np, std::max<size_t>(new_n, 1) * sizeof(T)));
    if (fill_byte >= 0) HNSW_CUDA_CHECK(cudaMemsetAsync(np, fill_byte, std::max<size_t>(new_n, 1) * sizeof(T), s));
    if (p && old_n) HNSW_CUDA_CHECK(cudaMemcpyAsync(np, p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (p) {
        HNSW_CUDA_CHECK(cudaStreamSynchronize(s));
        HNSW_CUDA_CHECK(cudaFree(p));
    }
    p = np;
}

static int seed_or_random(int seed) {
    if (seed >= 0) return seed;  // GraphData.cs:42 -- negative = unseeded
    std::random_device rd;
    return (int)(rd() & 0x7fffffff);
}

Index::Index(MetricKind metric, const Parameters& params) : metric_(metric), params_(params), rng_(seed_or_random(params.RandomSeed)) {
    if (params_.MaxEdges < 1) throw std::invalid_argument("MaxEdges must be at least 1");
    if (params_.MaxEdges > 256) throw std::invalid_argument("MaxEdges above 256 is not supported by the B200 engine");
    if (params_.MaxCandidates < 1) throw std::invalid_argument("MaxCandidates must be at least 1");
    if (params_.MinNN < 1) params_.MinNN = 1;
    HNSW_CUDA_CHECK(cudaGetDevice(&device_));
    HNSW_CUDA_CHECK(cudaStreamCreateWithFlags(&own_stream_, cudaStreamNonBlocking));
    stream_ = own_stream_;
    HNSW_CUDA_CHECK(cudaEventCreate(&ev0_));
    HNSW_CUDA_CHECK(cudaEventCreate(&ev1_));
    HNSW_CUDA_CHECK(cudaStreamCreateWithFlags(&aux_stream_, cudaStreamNonBlocking));
    HNSW_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
    HNSW_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
    HNSW_CUDA_CHECK(cudaMalloc(&d_counters_, 16 * sizeof(unsigned)));
    HNSW_CUDA_CHECK(cudaMemset(d_counters_, 0, 16 * sizeof(unsigned)));
    HNSW_CUDA_CHECK(cudaMalloc(&d_build_prof_, 16 * sizeof(u64)));
    HNSW_CUDA_CHECK(cudaMemset(d_build_prof_, 0, 16 * sizeof(u64)));
    stride0_ = stride_for(2 * params_.MaxEdges);
    strideU_ = stride_for(params_.MaxEdges);
    capacity_ = std::max(1, params_.CollectionSize);
}

Index::~Index() {
    cudaSetDevice(device_);
    cudaFree(d_vectors_); cudaFree(d_norms_); cudaFree(d_adj0_); cudaFree(d_upper_off_); cudaFree(d_adjU_);
    cudaFree(d_locks_); cudaFree(d_levels_); cudaFree(d_counters_); cudaFree(d_build_prof_);
    cudaFree(rq_keys_); cudaFree(rq_hash_); cudaFree(rq_small_); cudaFree(rq_dense_);
    cudaFree(rm_small_); cudaFree(rm_aff_); cudaFree(d_rm_index_); cudaFree(d_new_ids_); cudaFree(d_ready_);
    cudaFree(d_q_); cudaFree(d_ids_); cudaFree(d_dists_); cudaFree(d_stats_);
    for (QueryCtx& c : qctx_) {
        cudaFree(c.d_q); cudaFree(c.d_ids); cudaFree(c.d_dists); cudaFree(c.d_stats); cudaFree(c.d_counters);
        if (c.h_stage) cudaFreeHost(c.h_stage);
        if (c.ev0) cudaEventDestroy(c.ev0);
        if (c.ev1) cudaEventDestroy(c.ev1);
        if (c.ev_fork) cudaEventDestroy(c.ev_fork);
        if (c.ev_join) cudaEventDestroy(c.ev_join);
        if (c.stream) cudaStreamDestroy(c.stream);
        if (c.aux) cudaStreamDestroy(c.aux);
    }
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (ev_fork_) cudaEventDestroy(ev_fork_);
    if (ev_join_) cudaEventDestroy(ev_join_);
    if (aux_stream_) cudaStreamDestroy(aux_stream_);
    if (own_stream_) cudaStreamDestroy(own_stream_);
}

void Index::set_tuning(const std::string& key, int value) {
    if (key == "vis_size_search") tuning_.vis_size_search = value;
    else if (key == "vis_size_build") tuning_.vis_size_build = value;
    else if (key == "vis_global_search") tuning_.vis_global_search = value;
    else if (key == "blocks_per_sm") tuning_.blocks_per_sm = value;
    else if (key == "flags") tuning_.flags = value;
    else if (key == "zero_copy") zero_copy_ = value != 0;
    else if (key == "build_batch_max") build_batch_max_ = std::max(1, value);
    else if (key == "build_batch_ratio") build_batch_ratio_ = std::max(1, value);
    else if (key == "build_ramp_ratio") build_ramp_ratio_ = std::max(1, value);
    else if (key == "build_ramp_until") build_ramp_until_ = std::max(0, value);
    else if (key == "async_overlap") tuning_.async_overlap = value;
    else if (key == "sync_pieces") tuning_.sync_pieces = value;
    else if (key == "build_overlap") tuning_.build_overlap = value;
    else if (key == "build_specialized") tuning_.build_specialized = value;
    else if (key == "dim128") tuning_.dim128 = value;
    else if (key == "build_flags") tuning_.build_flags = value;
    else if (key == "build_tail_fill") tuning_.build_tail_fill = value;
    else if (key == "exact_tcgen05") tuning_.exact_tcgen05 = value;
    else if (key == "wl_min_ef") tuning_.wl_min_ef = value;
    else if (key == "wide_min_ef") tuning_.wide_min_ef = value;
    else if (key == "wide_warps") tuning_.wide_warps = value;
    else if (key == "wide_prefetch") tuning_.wide_prefetch = value;
    else if (key == "wide_max_small_batch") tuning_.wide_max_small_batch = value;
    else if (key == "wide_bitmap_min_ef") tuning_.wide_bitmap_min_ef = value;
    else if (key == "wide_bitmap_max_nodes") tuning_.wide_bitmap_max_nodes = value;
    else if (key == "peer_timeout_ms") tuning_.peer_timeout_ms = std::max(1, value);
    else if (key == "remove_batch_max") remove_batch_max_ = std::max(1, value);
    else if (key == "remove_batch_ratio") remove_batch_ratio_ = std::max(1, value);
    else throw std::invalid_argument("unknown tuning key: " + key);
}

GraphView Index::view() const {
    GraphView g;
    g.vectors = d_vectors_;
    g.norm_sqrt = d_norms_;
    g.adj0 = d_adj0_;
    g.adjU = d_adjU_;
    g.upper_off = d_upper_off_;
    g.locks = d_locks_;
    g.rm_index = d_rm_index_;
    g.dim = dim_;
    g.pitch = pitch_;
    g.stride0 = stride0_;
    g.strideU = strideU_;
    g.max_edges = params_.MaxEdges;
    g.entry = entry_;
    g.top_layer = top_;
    g.count = length_;
    return g;
}

void Index::build_profile(unsigned long long out[16], bool reset) {
    Exclusive lock(mu_);
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    HNSW_CUDA_CHECK(cudaMemcpy(out, d_build_prof_, 16 * sizeof(u64), cudaMemcpyDeviceToHost));
    if (reset) HNSW_CUDA_CHECK(cudaMemset(d_build_prof_, 0, 16 * sizeof(u64)));
}

void Index::ensure_dim(int dim) {
    if (dim_ < 0) {
        dim_ = dim;
        pitch_ = (dim + 3) & ~3;
    } else if (dim != dim_) {
        throw std::invalid_argument("vector dimension " + std::to_string(dim) + " does not match the index dimension " + std::to_string(dim_));
    }
}

// GraphData.cs:98-111: Capacity doubles whenever Length exceeds it
void Index::ensure_capacity(int needed) {
    if (needed > (1 << 30)) throw std::runtime_error("the index cannot hold more than 2^30 items");
    int cap = std::max(1, capacity_);
    while (needed > cap) cap *= 2;
    if (cap <= dev_capacity_) { capacity_ = cap; return; }
    // capacity_ / dev_capacity_ only change once every array has grown: a failed allocation leaves the index as it was
    const size_t o = (size_t)dev_capacity_, n = (size_t)cap;
    grow_device(d_vectors_, o * pitch_, n * pitch_, 0, stream_);
    if (metric_ == COSINE) grow_device(d_norms_, o, n, 0, stream_);
    grow_device(d_adj0_, o * stride0_, n * stride0_, 0xFF, stream_);
    grow_device(d_upper_off_, o, n, 0xFF, stream_);
    grow_device(d_locks_, o, n, 0, stream_);
    grow_device(d_rm_index_, o, n, 0xFF, stream_);
    grow_device(d_levels_, o, n, 0, stream_);
    capacity_ = cap;
    dev_capacity_ = cap;
}

void Index::ensure_upper_rows(int needed) {
    if (needed <= upper_rows_cap_) return;
    int cap = std::max(1024, upper_rows_cap_);
    while (cap < needed) cap *= 2;
    grow_device(d_adjU_, (size_t)upper_rows_cap_ * strideU_, (size_t)cap * strideU_, 0xFF, stream_);
    upper_rows_cap_ = cap;
}

void Index::ensure_query_scratch(int count, int k) {
    const size_t qn = (size_t)count * dim_, on = (size_t)count * k;
    if (qn > q_cap_) { cudaFree(d_q_); d_q_ = nullptr; HNSW_CUDA_CHECK(cudaMalloc(&d_q_, qn * sizeof(float))); q_cap_ = qn; }
    if (on > out_cap_) {
        cudaFree(d_ids_); cudaFree(d_dists_); d_ids_ = nullptr; d_dists_ = nullptr;
        HNSW_CUDA_CHECK(cudaMalloc(&d_ids_, on * sizeof(int)));
        HNSW_CUDA_CHECK(cudaMalloc(&d_dists_, on * sizeof(float)));
        out_cap_ = on;
    }
}

void Index::ensure_stats(int count) {
    if ((size_t)count > stats_cap_) {
        cudaFree(d_stats_); d_stats_ = nullptr;
        HNSW_CUDA_CHECK(cudaMalloc(&d_stats_, (size_t)count * sizeof(QueryStats)));
        stats_cap_ = count;
    }
}

void Index::set_stream(cudaStream_t s) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    stream_ = s ? s : own_stream_;
}

// ---------------------------------------------------------------------------------------------
// Add (HNSWIndex.cs:70-78).  Levels and slots are assigned in input order, i.e. the ids equal the
// ones the reference hands out when items are added one after another.
// ---------------------------------------------------------------------------------------------
int Index::add(const float* vectors, int count, int dim, int* out_ids) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    ensure_dim(dim);
    std::vector<int> lv, ids_new, src_row;
    lv.reserve(count); ids_new.reserve(count); src_row.reserve(count);
    int nfresh = 0, nreused = 0;
    size_t stack_top = removed_stack_.size();  // slots are taken from the stack only once nothing below can fail
    for (int i = 0; i < count; ++i) {
        // GraphData.GetRandomLayer, GraphData.cs:211-219
        const float r = rng_.NextSingle();
        const double x = -std::log((double)r) * params_.DistributionRate;
        if (!(x < 2147483648.0) || x < 0.0) {  // r == 0: the reference's cast goes negative and the item is dropped (:81-82)
            out_ids[i] = -1;
            continue;
        }
        int level = (int)x;
        if (level > 30) level = 30;
        int id;
        if (params_.AllowRemovals && stack_top > 0) {  // "Search for empty spot first", GraphData.cs:84-91
            id = removed_stack_[--stack_top];
            ++nreused;
        } else {
            id = length_ + nfresh++;
        }
        out_ids[i] = id;
        ids_new.push_back(id);
        lv.push_back(level);
        src_row.push_back(i);
    }
    const int nnew = (int)ids_new.size();
    if (nnew == 0) return count;
    const int first = length_;
    ensure_capacity(first + nfresh);
    // ---- commit point.  Everything above can fail without a trace (the level draws are consumed, as in the reference).  From
    // here on the host bookkeeping (slots, levels, active set, upper rows) runs ahead of the device work; if that work fails
    // (allocation, launch, a node lock that could not be taken) the index no longer describes a linked graph, so the handle
    // is marked unusable instead of answering from a half-inserted batch.
    removed_stack_.resize(stack_top);
    try {
        add_committed(vectors, count, dim, lv, ids_new, src_row, nfresh, nreused);
    } catch (const std::exception& e) {
        poisoned_ = e.what();
        throw std::runtime_error(std::string("Add failed after the index had been modified; the handle is unusable now (free it and "
                                             "rebuild or deserialize): ") + e.what());
    }
    return count;
}

void Index::check_usable() const {
    if (!poisoned_.empty())
        throw std::runtime_error("this index is unusable: an earlier Add failed half-way (" + poisoned_ + "); free it and rebuild or deserialize");
}

void Index::add_committed(const float* vectors, int count, int dim, const std::vector<int>& lv, const std::vector<int>& ids_new,
                          const std::vector<int>& src_row, int nfresh, int nreused) {
    const int nnew = (int)ids_new.size();
    const int first = length_;
    // vectors -> slab rows
    if (nreused == 0 && nnew == count) {
        HNSW_CUDA_CHECK(cudaMemcpy2DAsync(d_vectors_ + (size_t)first * pitch_, (size_t)pitch_ * 4, vectors, (size_t)dim * 4, (size_t)dim * 4,
                                          (size_t)nnew, cudaMemcpyHostToDevice, stream_));
    }
    // (re-used slots / dropped items: the batch is staged once and scattered by a kernel further down, together with the levels)
    // NewNode (GraphData.cs:224-242): rows for layers 0..level.  A reused slot gets fresh upper rows; its layer-0 row was
    // emptied when the old item was removed.
    levels_.resize(first + nfresh);
    upper_off_.resize(first + nfresh);
    removed_.resize(first + nfresh, 0);
    for (int j = 0; j < nnew; ++j) {
        const int id = ids_new[j];
        levels_[id] = lv[j];
        upper_off_[id] = lv[j] > 0 ? upper_rows_ : -1;
        upper_rows_ += lv[j];
        removed_[id] = 0;
        active_add(id);
    }
    ensure_upper_rows(upper_rows_);
    if ((size_t)nnew > new_ids_cap_) {
        cudaFree(d_new_ids_); d_new_ids_ = nullptr; new_ids_cap_ = 0;
        HNSW_CUDA_CHECK(cudaMalloc(&d_new_ids_, (size_t)nnew * sizeof(int)));
        new_ids_cap_ = nnew;
    }
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_new_ids_, ids_new.data(), (size_t)nnew * 4, cudaMemcpyHostToDevice, stream_));
    if (ready_cap_ < dev_capacity_) {  // one word per node: the epoch of the batch that wrote its rows (tail fill); epochs start at 1
        cudaFree(d_ready_); d_ready_ = nullptr; ready_cap_ = 0;
        HNSW_CUDA_CHECK(cudaMalloc(&d_ready_, (size_t)dev_capacity_ * sizeof(int)));
        HNSW_CUDA_CHECK(cudaMemsetAsync(d_ready_, 0, (size_t)dev_capacity_ * sizeof(int), stream_));
        ready_cap_ = dev_capacity_;
    }
    if (nreused == 0 && nnew == count) {
        HNSW_CUDA_CHECK(cudaMemcpyAsync(d_levels_ + first, levels_.data() + first, (size_t)nfresh * 4, cudaMemcpyHostToDevice, stream_));
        HNSW_CUDA_CHECK(cudaMemcpyAsync(d_upper_off_ + first, upper_off_.data() + first, (size_t)nfresh * 4, cudaMemcpyHostToDevice, stream_));
        if (metric_ == COSINE) launch_item_norms(d_vectors_, pitch_, dim_, first, nfresh, d_norms_, stream_);
    }
    if (nreused > 0 || nnew != count) {
        // One staged copy of the input batch, then a kernel scatters every accepted row to its slot and records the level and
        // first upper row of exactly the touched ids; norms in one launch over the id list (re-adding 100k items after removals
        // used to cost 100k copies + 100k one-item launches and a re-upload of the whole level arrays).
        std::vector<int> meta((size_t)nnew * 3);
        for (int j = 0; j < nnew; ++j) {
            meta[j] = src_row[j];
            meta[(size_t)nnew + j] = lv[j];
            meta[(size_t)2 * nnew + j] = upper_off_[ids_new[j]];
        }
        float* d_stage = nullptr;
        int* d_meta = nullptr;
        HNSW_CUDA_CHECK(cudaMallocAsync(&d_stage, (size_t)count * dim * sizeof(float), stream_));
        HNSW_CUDA_CHECK(cudaMallocAsync(&d_meta, meta.size() * sizeof(int), stream_));
        HNSW_CUDA_CHECK(cudaMemcpyAsync(d_stage, vectors, (size_t)count * dim * sizeof(float), cudaMemcpyHostToDevice, stream_));
        HNSW_CUDA_CHECK(cudaMemcpyAsync(d_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
        launch_scatter_new_items(d_stage, d_meta, d_new_ids_, d_meta + nnew, d_meta + 2 * (size_t)nnew, nnew, dim, pitch_, d_vectors_, d_levels_,
                                 d_upper_off_, stream_);
        launch_set_rm_index(d_rm_index_, d_new_ids_, nnew, false, -1, stream_);  // alive again
        if (metric_ == COSINE) launch_item_norms_ids(d_vectors_, pitch_, dim_, d_new_ids_, nnew, d_norms_, stream_);
        HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));  // `meta` and the caller's buffer are read by the copies above
        HNSW_CUDA_CHECK(cudaFreeAsync(d_stage, stream_));
        HNSW_CUDA_CHECK(cudaFreeAsync(d_meta, stream_));
    }
    length_ = first + nfresh;

    HNSW_CUDA_CHECK(cudaEventRecord(ev0_, stream_));
    int done = 0;
    bool links_on_aux = false;
    const bool ramp = (long long)(first + nfresh) >= 4ll * build_ramp_until_;
    int linked = this->count() - nnew;  // nodes a new node can reach (everything active before this call, then batch by batch)
    while (done < nnew) {
        const int v = ids_new[done];
        if (entry_ < 0) {  // GraphConnector.cs:28-33: the first node becomes the entry point, no links
            entry_ = v;
            top_ = levels_[v];
            ++done;
            ++linked;
            continue;
        }
        int batch = 1;
        const bool new_top = levels_[v] > top_;
        if (!new_top) {
            // ramp-up: while the graph is too small to fill the GPU every batch costs the latency of ONE insertion whatever its
            // size, so the early batches may take a larger share of the linked nodes (build_ramp_ratio_ below build_ramp_until_).
            // Nodes of one batch cannot link to each other, which a small index would keep as a recall loss (2 000 nodes built
            // at 1/4: self-recall 0.80 instead of 0.86) -- so only when this call leaves an index at least 4x beyond the ramp,
            // whose later insertions repair the early rows through their reverse links (1M: recall@10 0.9657 both ways)
            const int ratio = (ramp && linked < build_ramp_until_) ? std::min(build_ramp_ratio_, build_batch_ratio_) : build_batch_ratio_;
            const int b = std::max(1, std::min(build_batch_max_, linked / ratio));
            while (done + batch < nnew && batch < b && levels_[ids_new[done + batch]] <= top_) ++batch;
        }
        // nodes of the same batch cannot reach each other: nothing points at them before the link kernel runs
        const bool overlap = tuning_.build_overlap != 0 && batch >= 2048 && !new_top;
        if (!overlap && links_on_aux) {  // small / top-raising batches run strictly after everything before them
            HNSW_CUDA_CHECK(cudaEventRecord(ev_join_, aux_stream_));
            HNSW_CUDA_CHECK(cudaStreamWaitEvent(stream_, ev_join_, 0));
            links_on_aux = false;
        }
        launch_build_batch((int)metric_, view(), d_levels_, d_new_ids_ + done, batch, params_.MaxCandidates, d_counters_, d_build_prof_,
                           tuning_, stream_, overlap ? aux_stream_ : stream_, ev_fork_, d_ready_, ++build_epoch_);
        links_on_aux = links_on_aux || overlap;
        stats_.build_launches += 2;
        if (new_top) {  // GraphConnector.cs:36-41
            entry_ = v;
            top_ = levels_[v];
        }
        done += batch;
        linked += batch;
    }
    if (links_on_aux) {
        HNSW_CUDA_CHECK(cudaEventRecord(ev_join_, aux_stream_));
        HNSW_CUDA_CHECK(cudaStreamWaitEvent(stream_, ev_join_, 0));
    }
    HNSW_CUDA_CHECK(cudaEventRecord(ev1_, stream_));
    int err = 0;
    HNSW_CUDA_CHECK(cudaMemcpyAsync(&err, d_counters_ + 2, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    float ms = 0.f;
    HNSW_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
    stats_.last_build_ms = ms;
    if (err != 0) {
        HNSW_CUDA_CHECK(cudaMemset(d_counters_ + 2, 0, sizeof(int)));
        throw std::runtime_error("build: a node lock could not be acquired (index may be inconsistent)");
    }
}

// ActiveSet.cs:69-93
void Index::active_add(int id) {
    if ((int)active_sparse_.size() <= id) active_sparse_.resize((size_t)id + 1, -1);
    active_sparse_[id] = (int)active_dense_.size();
    active_dense_.push_back(id);
}
void Index::active_remove(int id) {
    const int idx = active_sparse_[id];
    const int last_id = active_dense_.back();
    active_dense_.pop_back();
    if (idx < (int)active_dense_.size()) {  // move last into the removed slot
        active_dense_[idx] = last_id;
        active_sparse_[last_id] = idx;
    }
    active_sparse_[id] = -1;
}

// ---------------------------------------------------------------------------------------------
// Remove (HNSWIndex.cs:83-100, GraphConnector.cs:53-167, GraphData.cs:124-199) in chunks; see remove_kernels.cu
// ---------------------------------------------------------------------------------------------
std::vector<int> Index::read_row(int id, int layer) {
    const int stride = layer == 0 ? stride0_ : strideU_;
    const int* src = layer == 0 ? d_adj0_ + (size_t)id * stride0_ : d_adjU_ + ((size_t)upper_off_[id] + (layer - 1)) * strideU_;
    std::vector<int> row(stride);
    HNSW_CUDA_CHECK(cudaMemcpyAsync(row.data(), src, (size_t)stride * 4, cudaMemcpyDeviceToHost, stream_));
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    int n = 0;
    while (n < stride && row[n] >= 0) ++n;
    row.resize(n);
    return row;
}

// ReplaceEntryPointIfNeeded (GraphConnector.cs:72-85) for an entry point that is part of the chunk: walk its layers from
// the top; the first layer with a surviving neighbour hands the role to the neighbour with most links on that layer
// (GraphData.TryReplaceEntryPoint, :146-166); otherwise the highest active node takes over (ForceReplaceEntryPoint,
// :172-189 -- the reference may pick the node being removed there; the chunk is excluded here).
void Index::replace_entry_point(const std::vector<unsigned char>& in_chunk, int chunk_size) {
    const int ep = entry_;
    for (int layer = levels_[ep]; layer >= 0; --layer) {
        int best = -1, best_deg = -1;
        for (int nb : read_row(ep, layer)) {
            if (in_chunk[nb]) continue;
            const int deg = (int)read_row(nb, layer).size();
            if (deg > best_deg) { best_deg = deg; best = nb; }
        }
        if (best >= 0) {
            entry_ = best;
            top_ = levels_[best];
            return;
        }
    }
    if (count() - chunk_size <= 0) {
        entry_ = -1;
        top_ = -1;
        return;
    }
    int best = -1, best_layer = -1;
    for (int id : active_dense_)
        if (!in_chunk[id] && levels_[id] > best_layer) { best_layer = levels_[id]; best = id; }
    entry_ = best;
    top_ = best_layer;
}

void Index::remove_chunk(const int* ids, int n) {
    // membership of the chunk: a byte map for large chunks, a linear scan for the one-id-per-call case
    std::vector<unsigned char> in_chunk;
    bool entry_in_chunk = false;
    for (int i = 0; i < n && !entry_in_chunk; ++i) entry_in_chunk = ids[i] == entry_;
    if (entry_in_chunk) {
        in_chunk.assign(length_, 0);
        for (int i = 0; i < n; ++i) in_chunk[ids[i]] = 1;
        replace_entry_point(in_chunk, n);
    }

    // work items: (node, layer) for every layer of every removed node, a node's items consecutive from layer 0
    std::vector<int> item_node, item_layer, item_first(n);
    for (int i = 0; i < n; ++i) {
        item_first[i] = (int)item_node.size();
        for (int l = 0; l <= levels_[ids[i]]; ++l) { item_node.push_back(ids[i]); item_layer.push_back(l); }
    }
    const int nitems = (int)item_node.size();
    const int rmc = std::max(1, params_.RemoveMaxCandidates);
    const size_t max_rows = (size_t)length_ + (size_t)upper_rows_;
    // device scratch, kept between calls (grow-only): ids | item_node | item_layer | item_first | cand, and the affected rows
    auto grow = [&](void*& p, size_t& have, size_t need) {
        if (need <= have) return;
        cudaFree(p); p = nullptr; have = 0;
        HNSW_CUDA_CHECK(cudaMalloc(&p, need + need / 2));
        have = need + need / 2;
    };
    grow(rm_small_, rm_small_bytes_, ((size_t)n * 2 + (size_t)nitems * 2 + (size_t)nitems * rmc) * 4);
    grow(rm_aff_, rm_aff_bytes_, max_rows * sizeof(int2));
    int* d_ids = static_cast<int*>(rm_small_);
    int* d_item_first = d_ids + n;
    int* d_item_node = d_item_first + n;
    int* d_item_layer = d_item_node + nitems;
    int* d_cand = d_item_layer + nitems;
    int2* d_aff = static_cast<int2*>(rm_aff_);
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_ids, ids, (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_item_first, item_first.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_item_node, item_node.data(), (size_t)nitems * 4, cudaMemcpyHostToDevice, stream_));
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_item_layer, item_layer.data(), (size_t)nitems * 4, cudaMemcpyHostToDevice, stream_));
    launch_set_rm_index(d_rm_index_, d_ids, n, true, 0, stream_);
    const GraphView g = view();
    launch_remove_search((int)metric_, g, d_item_node, d_item_layer, nitems, rmc, d_cand, d_counters_ + 4, stream_);
    launch_find_affected(g, d_levels_, length_, d_aff, d_counters_ + 5, stream_);
    unsigned naff = 0;
    HNSW_CUDA_CHECK(cudaMemcpyAsync(&naff, d_counters_ + 5, sizeof(unsigned), cudaMemcpyDeviceToHost, stream_));
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    launch_relink((int)metric_, g, d_aff, (int)naff, d_item_first, d_cand, rmc, d_counters_ + 4, stream_);
    launch_clear_rows(g, d_item_node, d_item_layer, nitems, stream_);
    launch_set_rm_index(d_rm_index_, d_ids, n, false, -2, stream_);
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    stats_.remove_launches += 6;
    for (int i = 0; i < n; ++i) {  // GraphData.RemoveItem, GraphData.cs:124-128
        removed_[ids[i]] = 1;
        removed_stack_.push_back(ids[i]);
        active_remove(ids[i]);
    }
}

void Index::remove(const int* ids, int count) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    if (!params_.AllowRemovals) throw std::runtime_error("Removals are disabled in this index instance.");  // HNSWIndex.cs:85-86
    if (count <= 0) return;
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    {
        std::vector<int> sorted(ids, ids + count);
        std::sort(sorted.begin(), sorted.end());
        for (int i = 0; i < count; ++i) {
            const int id = sorted[i];
            if (id < 0 || id >= length_) throw std::out_of_range("remove: item index " + std::to_string(id) + " is out of range");
            if (removed_[id]) throw std::invalid_argument("remove: item " + std::to_string(id) + " has already been removed");
            if (i > 0 && sorted[i - 1] == id) throw std::invalid_argument("remove: item " + std::to_string(id) + " is listed twice");
        }
    }
    int done = 0;
    while (done < count) {
        const int chunk = std::max(1, std::min(std::min(remove_batch_max_, this->count() / remove_batch_ratio_), count - done));
        remove_chunk(ids + done, chunk);
        done += chunk;
    }
}

// ---------------------------------------------------------------------------------------------
// BatchKnnQuery (HNSWIndex.cs:129-137) + result marshalling of HNSWIndexExports.cs:135-145
// ---------------------------------------------------------------------------------------------
static void fill_empty(int* ids, float* dists, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        ids[i] = -1;
        dists[i] = std::numeric_limits<float>::quiet_NaN();
    }
}

void Index::fetch_query_stats(int count) {
    std::vector<QueryStats> h(count);
    HNSW_CUDA_CHECK(cudaMemcpyAsync(h.data(), d_stats_, (size_t)count * sizeof(QueryStats), cudaMemcpyDeviceToHost, stream_));
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    stats_.ndist = stats_.nexp = stats_.row_words = 0;
    for (auto& q : h) { stats_.ndist += q.ndist; stats_.nexp += q.nexp; stats_.row_words += q.row_words; }
}

// Device-visible alias of a host buffer when it is page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory):
// the kernel can then read or write it in place over PCIe, so the transfer overlaps the search instead of
// preceding / following it.  Pageable memory returns nullptr and takes the staged cudaMemcpy path.
template <typename T>
static T* pinned_alias(const T* host) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || a.devicePointer == nullptr) return nullptr;
    return reinterpret_cast<T*>(a.devicePointer);
}

// ---- query contexts: stream + scratch of one in-flight hnsw_knn_query ----------------------------------------------
int Index::acquire_ctx() {
    std::unique_lock<std::mutex> l(ctx_mu_);
    for (;;) {
        for (int i = 0; i < kQueryContexts; ++i)
            if (!qctx_[i].busy) {
                qctx_[i].busy = true;
                return i;
            }
        ctx_cv_.wait(l);
    }
}
void Index::release_ctx(int i) {
    {
        std::lock_guard<std::mutex> l(ctx_mu_);
        qctx_[i].busy = false;
    }
    ctx_cv_.notify_one();
}
// called by the thread that holds the context
void Index::prepare_ctx(QueryCtx& c, int count, int k, bool want_stats) {
    if (!c.created) {
        HNSW_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        HNSW_CUDA_CHECK(cudaStreamCreateWithFlags(&c.aux, cudaStreamNonBlocking));
        HNSW_CUDA_CHECK(cudaEventCreate(&c.ev0));
        HNSW_CUDA_CHECK(cudaEventCreate(&c.ev1));
        HNSW_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
        HNSW_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
        HNSW_CUDA_CHECK(cudaMalloc(&c.d_counters, 16 * sizeof(unsigned)));
        HNSW_CUDA_CHECK(cudaMemset(c.d_counters, 0, 16 * sizeof(unsigned)));
        HNSW_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&c.h_stage), kSmallCallBytes, cudaHostAllocMapped));
        HNSW_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c.d_stage), c.h_stage, 0));
        c.created = true;
    }
    const size_t qn = (size_t)count * dim_, on = (size_t)count * k;
    if (qn > c.q_cap) { cudaFree(c.d_q); c.d_q = nullptr; c.q_cap = 0; HNSW_CUDA_CHECK(cudaMalloc(&c.d_q, qn * sizeof(float))); c.q_cap = qn; }
    if (on > c.out_cap) {
        cudaFree(c.d_ids); cudaFree(c.d_dists); c.d_ids = nullptr; c.d_dists = nullptr; c.out_cap = 0;
        HNSW_CUDA_CHECK(cudaMalloc(&c.d_ids, on * sizeof(int)));
        HNSW_CUDA_CHECK(cudaMalloc(&c.d_dists, on * sizeof(float)));
        c.out_cap = on;
    }
    if (want_stats && (size_t)count > c.stats_cap) {
        cudaFree(c.d_stats); c.d_stats = nullptr; c.stats_cap = 0;
        HNSW_CUDA_CHECK(cudaMalloc(&c.d_stats, (size_t)count * sizeof(QueryStats)));
        c.stats_cap = count;
    }
}

// BatchKnnQuery (HNSWIndex.cs:129-137).  Runs under the SHARED lock: host threads calling it at the same time each get a
// query context (stream, scratch, work counters) and their kernels share the GPU; Add / Remove / Serialize wait for them.
void Index::knn_query(const float* queries, int count, int dim, int k, int* out_ids, float* out_dists, int layer) {
    if (aux_dirty_.load()) {  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
        Exclusive ex(mu_);
        join_aux_locked(true);
    }
    std::shared_lock<std::shared_mutex> shared(mu_);
    check_usable();
    if (count <= 0 || k < 1) return;
    if (this->count() <= 0) {  // HNSWIndex.cs:109
        fill_empty(out_ids, out_dists, (size_t)count * k);
        return;
    }
    // the reference indexes OutEdges[layer] of a node found on that layer: IndexOutOfRangeException beyond the top layer
    if (layer < 0 || layer > top_) throw std::out_of_range("layer " + std::to_string(layer) + " does not exist (top layer is " + std::to_string(top_) + ")");
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    if (dim != dim_) throw std::invalid_argument("vector dimension " + std::to_string(dim) + " does not match the index dimension " + std::to_string(dim_));
    const bool want_stats = collect_stats_;
    const int ci = acquire_ctx();
    struct Release {
        Index* ix; int i;
        ~Release() { ix->release_ctx(i); }
    } release{this, ci};
    QueryCtx& c = qctx_[ci];
    prepare_ctx(c, count, k, want_stats);
    cudaStream_t st0 = ci == 0 ? stream_ : c.stream;  // context 0 honours set_stream()
    const int ef = std::max(params_.MinNN, k);  // HNSWIndex.cs:115
    const float* q_src = zero_copy_ ? pinned_alias(queries) : nullptr;
    int* ids_dst = zero_copy_ ? pinned_alias(out_ids) : nullptr;
    float* d_dst = zero_copy_ ? pinned_alias(out_dists) : nullptr;
    // Small calls (KnnQuery for one item is the C# API's default shape): three copy commands around a 180 us kernel are a tenth
    // of the call.  Queries and results go through the context's page-locked, device-mapped scratch instead -- the kernel reads
    // and writes it in place, the host side is two memcpy of a few hundred bytes.
    const size_t q_bytes = (size_t)count * dim * sizeof(float), o_bytes = (size_t)count * k * sizeof(int);
    const size_t q_room = (q_bytes + 255) & ~(size_t)255, o_room = (o_bytes + 255) & ~(size_t)255;
    const bool small_call = zero_copy_ && !want_stats && q_room + 2 * o_room <= kSmallCallBytes && (q_src == nullptr || ids_dst == nullptr || d_dst == nullptr);
    if (small_call) {
        std::memcpy(c.h_stage, queries, q_bytes);
        q_src = reinterpret_cast<const float*>(c.d_stage);
        ids_dst = reinterpret_cast<int*>(c.d_stage + q_room);
        d_dst = reinterpret_cast<float*>(c.d_stage + q_room + o_room);
    }
    const bool direct_out = ids_dst != nullptr && d_dst != nullptr;
    int* ids_out = direct_out ? ids_dst : c.d_ids;
    float* d_out = direct_out ? d_dst : c.d_dists;
    // Pageable queries (what numpy hands over through the reference's own bindings.py) are staged by the driver, and the
    // host thread is held while that happens: a large batch is therefore uploaded in four pieces, each piece's search
    // launched as soon as it is on the device and alternating between the two streams, so the upload of piece i+1 runs
    // under the search of piece i.
    // Page-locked batches need no upload, but the same split pays there too: a persistent launch ends with a tail in which
    // its last queries keep only part of the SM slots busy, and the next piece's blocks (other stream) move in as they free up.
    int pieces = 1;
    if (!want_stats && layer == 0 && count >= 4096 && tuning_.async_overlap != 0)
        pieces = tuning_.sync_pieces > 0 ? tuning_.sync_pieces : (q_src ? 2 : 4);
    pieces = std::max(1, std::min(pieces, 16));
    HNSW_CUDA_CHECK(cudaEventRecord(c.ev0, st0));
    if (pieces == 1) {
        if (!q_src) {
            HNSW_CUDA_CHECK(cudaMemcpyAsync(c.d_q, queries, (size_t)count * dim * 4, cudaMemcpyHostToDevice, st0));
            q_src = c.d_q;
        }
        launch_knn_search((int)metric_, view(), q_src, count, k, ef, layer, ids_out, d_out, c.d_counters + 3, want_stats ? c.d_stats : nullptr,
                          tuning_, st0);
    } else {
        HNSW_CUDA_CHECK(cudaEventRecord(c.ev_fork, st0));
        HNSW_CUDA_CHECK(cudaStreamWaitEvent(c.aux, c.ev_fork, 0));
        const GraphView g = view();
        for (int p = 0; p < pieces; ++p) {
            const int q0 = (int)((long long)count * p / pieces), q1 = (int)((long long)count * (p + 1) / pieces);
            cudaStream_t st = (p & 1) ? c.aux : st0;
            const float* qp = q_src ? q_src + (size_t)q0 * dim : c.d_q + (size_t)q0 * dim;
            if (!q_src)
                HNSW_CUDA_CHECK(cudaMemcpyAsync(c.d_q + (size_t)q0 * dim, queries + (size_t)q0 * dim, (size_t)(q1 - q0) * dim * 4, cudaMemcpyHostToDevice, st));
            launch_knn_search((int)metric_, g, qp, q1 - q0, k, ef, 0, ids_out + (size_t)q0 * k, d_out + (size_t)q0 * k,
                              c.d_counters + ((p & 1) ? 8 : 3), nullptr, tuning_, st);
        }
        HNSW_CUDA_CHECK(cudaEventRecord(c.ev_join, c.aux));
        HNSW_CUDA_CHECK(cudaStreamWaitEvent(st0, c.ev_join, 0));
    }
    HNSW_CUDA_CHECK(cudaEventRecord(c.ev1, st0));
    if (!direct_out) {
        HNSW_CUDA_CHECK(cudaMemcpyAsync(out_ids, c.d_ids, (size_t)count * k * 4, cudaMemcpyDeviceToHost, st0));
        HNSW_CUDA_CHECK(cudaMemcpyAsync(out_dists, c.d_dists, (size_t)count * k * 4, cudaMemcpyDeviceToHost, st0));
    }
    std::vector<QueryStats> hstats;
    if (want_stats) {
        hstats.resize(count);
        HNSW_CUDA_CHECK(cudaMemcpyAsync(hstats.data(), c.d_stats, (size_t)count * sizeof(QueryStats), cudaMemcpyDeviceToHost, st0));
    }
    HNSW_CUDA_CHECK(cudaStreamSynchronize(st0));
    if (small_call) {
        std::memcpy(out_ids, c.h_stage + q_room, o_bytes);
        std::memcpy(out_dists, c.h_stage + q_room + o_room, o_bytes);
    }
    float ms = 0.f;
    HNSW_CUDA_CHECK(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    {
        std::lock_guard<std::mutex> l(stats_mu_);
        stats_.last_search_ms = ms;
        stats_.search_launches += pieces;
        if (want_stats) {
            stats_.ndist = stats_.nexp = stats_.row_words = 0;
            for (auto& q : hstats) { stats_.ndist += q.ndist; stats_.nexp += q.nexp; stats_.row_words += q.row_words; }
        }
    }
}

void Index::knn_query_device(const float* d_queries, int count, int dim, int k, int* d_ids, float* d_dists) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    if (count <= 0 || k < 1) return;
    if (this->count() <= 0) throw std::runtime_error("knn_query_device on an empty index");
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    ensure_dim(dim);
    if (collect_stats_) ensure_stats(count);
    const int ef = std::max(params_.MinNN, k);
    HNSW_CUDA_CHECK(cudaEventRecord(ev0_, stream_));
    launch_knn_search((int)metric_, view(), d_queries, count, k, ef, 0, d_ids, d_dists, d_counters_ + 3, collect_stats_ ? d_stats_ : nullptr,
                      tuning_, stream_);
    HNSW_CUDA_CHECK(cudaEventRecord(ev1_, stream_));
    stats_.search_launches += 1;
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    float ms = 0.f;
    HNSW_CUDA_CHECK(cudaEventElapsedTime(&ms, ev0_, ev1_));
    stats_.last_search_ms = ms;
    if (collect_stats_) fetch_query_stats(count);
}

// MultiLayerKnnQuery (HNSWIndex.cs:173-187) for a batch of queries.  Returns the number of layers L = min(top, maxLayer) + 1;
// outputs are [count][L][k] (pass null pointers to query L only).
int Index::multilayer_knn_query(const float* queries, int count, int dim, int k, int max_layer, int min_layer, int* out_ids, float* out_dists) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    if (this->count() <= 0 || k < 1) return 0;  // HNSWIndex.cs:176
    if (max_layer < 0) throw std::out_of_range("maxLayer must not be negative");
    const int top_l = std::min(top_, max_layer);
    const int layers = top_l + 1;
    if (out_ids == nullptr || out_dists == nullptr || count <= 0) return layers;
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    ensure_dim(dim);
    ensure_query_scratch(count, k * layers);
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_q_, queries, (size_t)count * dim * 4, cudaMemcpyHostToDevice, stream_));
    launch_multilayer_search((int)metric_, view(), d_q_, count, k, top_l, std::max(min_layer, 0), d_ids_, d_dists_, d_counters_ + 3, tuning_, stream_);
    stats_.search_launches += 1;
    const size_t n = (size_t)count * layers * k;
    HNSW_CUDA_CHECK(cudaMemcpyAsync(out_ids, d_ids_, n * 4, cudaMemcpyDeviceToHost, stream_));
    HNSW_CUDA_CHECK(cudaMemcpyAsync(out_dists, d_dists_, n * 4, cudaMemcpyDeviceToHost, stream_));
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    return layers;
}

// Stream-ordered variant: the search is only queued (no host synchronisation, no event timing); synchronize() waits.
// Odd-numbered batches go to the auxiliary stream (which first waits for everything queued on the primary one when it
// is picked up again after a join), so two consecutive batches overlap: the blocks of the later one fill the SM slots
// that the persistent blocks of the earlier one give up in its tail.
void Index::knn_query_device_async(const float* d_queries, int count, int dim, int k, int* d_ids, float* d_dists) {
    Exclusive lock(mu_);
    if (count <= 0 || k < 1) return;
    if (this->count() <= 0) throw std::runtime_error("knn_query_device_async on an empty index");
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    ensure_dim(dim);
    const int ef = std::max(params_.MinNN, k);
    const bool overlap = tuning_.async_overlap != 0;
    if (overlap && !aux_dirty_) {  // first batch of a sequence: the auxiliary stream starts behind what the primary holds NOW
        HNSW_CUDA_CHECK(cudaEventRecord(ev_fork_, stream_));
        HNSW_CUDA_CHECK(cudaStreamWaitEvent(aux_stream_, ev_fork_, 0));
        aux_dirty_ = true;
        async_seq_ = 0;
    }
    const bool use_aux = overlap && (async_seq_++ & 1u) != 0;
    if (!use_aux) {
        launch_knn_search((int)metric_, view(), d_queries, count, k, ef, 0, d_ids, d_dists, d_counters_ + 3, nullptr, tuning_, stream_);
    } else {
        launch_knn_search((int)metric_, view(), d_queries, count, k, ef, 0, d_ids, d_dists, d_counters_ + 8, nullptr, tuning_, aux_stream_);
    }
    stats_.search_launches += 1;
}

void Index::join_aux_locked(bool host_wait) {
    if (!aux_dirty_) return;
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    if (host_wait) {
        HNSW_CUDA_CHECK(cudaStreamSynchronize(aux_stream_));
    } else {
        HNSW_CUDA_CHECK(cudaEventRecord(ev_join_, aux_stream_));
        HNSW_CUDA_CHECK(cudaStreamWaitEvent(stream_, ev_join_, 0));
    }
    aux_dirty_ = false;
}

void Index::async_join() {
    Exclusive lock(mu_);
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    join_aux_locked(false);
}

void Index::synchronize() {
    Exclusive lock(mu_);
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    join_aux_locked(true);
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void Index::knn_query_scatter(const float* d_queries, int count, int dim, int k, const ScatterArgs& sc) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    if (count <= 0 || k < 1) return;
    if (this->count() <= 0) throw std::runtime_error("knn_query_scatter on an empty shard");
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    ensure_dim(dim);
    const int ef = std::max(params_.MinNN, k);
    launch_knn_scatter((int)metric_, view(), d_queries, count, k, ef, sc, d_counters_ + 3, d_counters_ + 6, tuning_, stream_);
    stats_.search_launches += 1;
}

void Index::wait_and_merge(const unsigned* d_flags, int world, unsigned epoch, const int* d_in_ids, const float* d_in_dists, int count, int k,
                           int* d_out_ids, float* d_out_dists, bool sync) {
    {
        Exclusive lock(mu_);
        HNSW_CUDA_CHECK(cudaSetDevice(device_));
        int* d_err = reinterpret_cast<int*>(d_counters_ + 7);
        launch_wait_merge(d_flags, world, epoch, tuning_.peer_timeout_ms, d_err, d_in_ids, d_in_dists, count, k, d_out_ids, d_out_dists, stream_);
        stats_.search_launches += 1;
    }
    if (sync) partition_sync();
}

// Wait for everything queued by knn_query_scatter / wait_and_merge and report a peer that never delivered.
void Index::partition_sync() {
    Exclusive lock(mu_);
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    int* d_err = reinterpret_cast<int*>(d_counters_ + 7);
    int err = 0;
    HNSW_CUDA_CHECK(cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    if (err != 0) {
        HNSW_CUDA_CHECK(cudaMemset(d_err, 0, sizeof(int)));
        throw std::runtime_error("partitioned query: a peer GPU did not deliver its results within " + std::to_string(tuning_.peer_timeout_ms) +
                                 " ms (tuning key peer_timeout_ms); the outputs of the affected batches are -1 / NaN and the exchange "
                                 "must be set up again (enable_fused + barrier) before the next fused query");
    }
}

void Index::exact_knn(const float* queries, int count, int dim, int k, int* out_ids, float* out_dists) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    if (count <= 0 || k < 1) return;
    if (this->count() <= 0) {
        fill_empty(out_ids, out_dists, (size_t)count * k);
        return;
    }
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    ensure_dim(dim);
    exact_scan_locked(queries, count, k, out_ids, out_dists);
}

// K3 (exact_kernels.cu) over a host batch, in slabs that bound the scratch
void Index::exact_scan_locked(const float* queries, int count, int k, int* out_ids, float* out_dists) {
    const int dim = dim_;
    const int slab = std::max(8, std::min(count, (int)((size_t)(1u << 30) / ((size_t)32 * k * 8))));
    void* scratch = nullptr;
    const size_t sbytes = exact_knn_scratch_bytes(slab, k, length_);
    HNSW_CUDA_CHECK(cudaMalloc(&scratch, sbytes));
    try {
        for (int q0 = 0; q0 < count; q0 += slab) {
            const int n = std::min(slab, count - q0);
            ensure_query_scratch(n, k);
            HNSW_CUDA_CHECK(cudaMemcpyAsync(d_q_, queries + (size_t)q0 * dim, (size_t)n * dim * 4, cudaMemcpyHostToDevice, stream_));
            launch_exact_knn((int)metric_, view(), d_q_, n, k, d_ids_, d_dists_, scratch, sbytes, stream_);
            HNSW_CUDA_CHECK(cudaMemcpyAsync(out_ids + (size_t)q0 * k, d_ids_, (size_t)n * k * 4, cudaMemcpyDeviceToHost, stream_));
            HNSW_CUDA_CHECK(cudaMemcpyAsync(out_dists + (size_t)q0 * k, d_dists_, (size_t)n * k * 4, cudaMemcpyDeviceToHost, stream_));
            HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
        }
    } catch (...) {
        cudaFree(scratch);
        throw;
    }
    cudaFree(scratch);
}

// K3b: tensor-core filter + bit-exact re-ranking; uncertified queries go through K3.  Same answers as exact_knn.
int Index::exact_knn_tf32(const float* queries, int count, int dim, int k, int* out_ids, float* out_dists) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    if (count <= 0 || k < 1) return 0;
    if (this->count() <= 0) {
        fill_empty(out_ids, out_dists, (size_t)count * k);
        return 0;
    }
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    ensure_dim(dim);
    if (exact_tf32_block_queries(dim_, k) == 0) {  // very long rows / very large k: the FFMA scan answers everything
        exact_scan_locked(queries, count, k, out_ids, out_dists);
        return count;
    }
    const int kp = exact_tf32_candidates(k);
    const int slab = std::max(64, std::min(count, (int)((size_t)(1u << 30) / ((size_t)33 * kp * 8))));
    const size_t sbytes = exact_tf32_scratch_bytes(slab, k, length_);
    void* scratch = nullptr;
    int* d_fail = nullptr;
    std::vector<int> fail, redo;
    try {
        HNSW_CUDA_CHECK(cudaMalloc(&scratch, sbytes));
        HNSW_CUDA_CHECK(cudaMalloc(&d_fail, (size_t)slab * sizeof(int)));
        for (int q0 = 0; q0 < count; q0 += slab) {
            const int n = std::min(slab, count - q0);
            ensure_query_scratch(n, k);
            HNSW_CUDA_CHECK(cudaMemcpyAsync(d_q_, queries + (size_t)q0 * dim, (size_t)n * dim * 4, cudaMemcpyHostToDevice, stream_));
            launch_exact_knn_tf32((int)metric_, view(), d_q_, n, k, d_ids_, d_dists_, d_fail, scratch, sbytes, q0 == 0, tuning_.exact_tcgen05,
                                  stream_);
            stats_.search_launches += q0 == 0 ? 4 : 3;
            fail.resize(n);
            HNSW_CUDA_CHECK(cudaMemcpyAsync(out_ids + (size_t)q0 * k, d_ids_, (size_t)n * k * 4, cudaMemcpyDeviceToHost, stream_));
            HNSW_CUDA_CHECK(cudaMemcpyAsync(out_dists + (size_t)q0 * k, d_dists_, (size_t)n * k * 4, cudaMemcpyDeviceToHost, stream_));
            HNSW_CUDA_CHECK(cudaMemcpyAsync(fail.data(), d_fail, (size_t)n * 4, cudaMemcpyDeviceToHost, stream_));
            HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
            for (int i = 0; i < n; ++i)
                if (fail[i]) redo.push_back(q0 + i);
        }
    } catch (...) {
        cudaFree(scratch);
        cudaFree(d_fail);
        throw;
    }
    cudaFree(scratch);
    cudaFree(d_fail);
    if (!redo.empty()) {
        const size_t m = redo.size();
        std::vector<float> rq(m * dim);
        std::vector<int> rids(m * k);
        std::vector<float> rd(m * k);
        for (size_t i = 0; i < m; ++i) std::memcpy(rq.data() + i * dim, queries + (size_t)redo[i] * dim, (size_t)dim * 4);
        exact_scan_locked(rq.data(), (int)m, k, rids.data(), rd.data());
        for (size_t i = 0; i < m; ++i) {
            std::memcpy(out_ids + (size_t)redo[i] * k, rids.data() + i * k, (size_t)k * 4);
            std::memcpy(out_dists + (size_t)redo[i] * k, rd.data() + i * k, (size_t)k * 4);
        }
    }
    return (int)redo.size();
}

// ---------------------------------------------------------------------------------------------
// BatchRangeQuery (HNSWIndex.cs:160-168)
// ---------------------------------------------------------------------------------------------
void Index::range_query(const float* queries, int count, int dim, float range, std::vector<std::vector<std::pair<float, int>>>& out) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    out.assign((size_t)std::max(count, 0), {});
    if (count <= 0 || this->count() <= 0) return;  // HNSWIndex.cs:146
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    ensure_dim(dim);
    ensure_query_scratch(count, 1);
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_q_, queries, (size_t)count * dim * 4, cudaMemcpyHostToDevice, stream_));

    // Device scratch is kept between calls (grow-only).  Per round: search -> counts to the host -> offsets -> the device
    // sorts every result and packs it -> only the packed results travel back.  Queries whose result or visited table
    // overflowed are searched again with 4x the room.
    auto grow = [&](void*& p, size_t& have, size_t need) {
        if (need <= have) return;
        cudaFree(p); p = nullptr; have = 0;
        HNSW_CUDA_CHECK(cudaMalloc(&p, need));
        have = need;
    };
    std::vector<int> todo(count);
    for (int i = 0; i < count; ++i) todo[i] = i;
    int cap = 1024;
    const int cap_max = std::max(1024, length_);
    std::vector<int> cnt, flags, retry;
    std::vector<long long> off;
    std::vector<u64> dense;
    while (!todo.empty()) {
        cap = std::min(cap, cap_max);
        int hash_size = 4096;
        while (hash_size < 4 * cap) hash_size <<= 1;  // load factor <= 0.5 at 2*cap visited nodes
        const size_t per_q = (size_t)cap * 8 + (size_t)hash_size * 4 + 8;
        const int chunk = (int)std::max<size_t>(1, std::min<size_t>(todo.size(), ((size_t)768 << 20) / per_q));
        grow(rq_keys_, rq_keys_bytes_, (size_t)chunk * cap * 8);
        grow(rq_hash_, rq_hash_bytes_, (size_t)chunk * hash_size * 4);
        const size_t off_at = ((size_t)chunk * 12 + 7) & ~(size_t)7;  // counts, map, flags (int each), then 8-byte offsets
        grow(rq_small_, rq_small_bytes_, off_at + (size_t)chunk * 8);
        int* d_cnt = static_cast<int*>(rq_small_);
        int* d_map = d_cnt + chunk;
        int* d_flag = d_map + chunk;
        long long* d_off = reinterpret_cast<long long*>(static_cast<char*>(rq_small_) + off_at);
        retry.clear();
        for (size_t c0 = 0; c0 < todo.size(); c0 += chunk) {
            const int nw = (int)std::min<size_t>(chunk, todo.size() - c0);
            HNSW_CUDA_CHECK(cudaMemcpyAsync(d_map, todo.data() + c0, (size_t)nw * 4, cudaMemcpyHostToDevice, stream_));
            launch_range_search((int)metric_, view(), d_q_, d_map, nw, range, cap, hash_size, static_cast<u64*>(rq_keys_),
                                static_cast<uint32_t*>(rq_hash_), d_cnt, d_counters_ + 3, tuning_, stream_);
            stats_.search_launches += 1;
            cnt.resize(nw);
            HNSW_CUDA_CHECK(cudaMemcpyAsync(cnt.data(), d_cnt, (size_t)nw * 4, cudaMemcpyDeviceToHost, stream_));
            HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
            off.resize(nw);
            long long total = 0;
            for (int w = 0; w < nw; ++w) {
                off[w] = total;
                if (cnt[w] > 0) total += cnt[w];
            }
            if (total > 0) {
                grow(rq_dense_, rq_dense_bytes_, (size_t)total * 8);
                HNSW_CUDA_CHECK(cudaMemcpyAsync(d_off, off.data(), (size_t)nw * 8, cudaMemcpyHostToDevice, stream_));
                launch_range_finalize(static_cast<u64*>(rq_keys_), cap, d_cnt, d_off, nw, static_cast<u64*>(rq_dense_), d_flag, stream_);
                stats_.search_launches += 1;
                dense.resize((size_t)total);
                flags.resize(nw);
                HNSW_CUDA_CHECK(cudaMemcpyAsync(dense.data(), rq_dense_, (size_t)total * 8, cudaMemcpyDeviceToHost, stream_));
                HNSW_CUDA_CHECK(cudaMemcpyAsync(flags.data(), d_flag, (size_t)nw * 4, cudaMemcpyDeviceToHost, stream_));
                HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
            }
            for (int w = 0; w < nw; ++w) {
                const int qi = todo[c0 + w];
                if (cnt[w] < 0) { retry.push_back(qi); continue; }
                if (cnt[w] == 0) continue;
                u64* kq = dense.data() + off[w];
                if (flags[w]) std::sort(kq, kq + cnt[w]);  // ascending by (distance, id): OrderBy(c => c.Dist), HNSWIndex.cs:154
                auto& dst = out[qi];
                dst.resize(cnt[w]);
                for (int j = 0; j < cnt[w]; ++j) dst[j] = {key_dist(kq[j]), key_id(kq[j])};
            }
        }
        if (!retry.empty() && cap >= cap_max) throw std::runtime_error("range_query: result does not fit the index size (internal error)");
        todo.swap(retry);
        cap *= 4;
    }
}

// ---------------------------------------------------------------------------------------------
// Serialize / Deserialize (HNSWIndex.cs:210-229)
// ---------------------------------------------------------------------------------------------
void Index::export_graph(HostGraph& g, bool with_vectors) {
    HNSW_CUDA_CHECK(cudaSetDevice(device_));
    g = HostGraph();
    g.params = params_;
    g.dim = std::max(dim_, 0);
    g.length = length_;
    g.capacity = capacity_;
    g.entry = entry_;
    g.stride0 = stride0_;
    g.strideU = strideU_;
    g.levels.assign(levels_.begin(), levels_.begin() + length_);
    g.upper_off.assign(upper_off_.begin(), upper_off_.begin() + length_);
    if (count() != length_ || !removed_stack_.empty()) {  // removal state travels only when there is any
        g.removed.assign(removed_.begin(), removed_.begin() + length_);
        g.removed_stack = removed_stack_;
    }
    bool identity = (int)active_dense_.size() == length_;
    for (int i = 0; identity && i < length_; ++i) identity = active_dense_[i] == i;
    if (!identity) g.active = active_dense_;
    if (with_vectors) g.vectors.resize((size_t)length_ * g.dim);
    g.adj0.resize((size_t)length_ * stride0_);
    g.adjU.resize((size_t)upper_rows_ * strideU_);
    if (length_ > 0) {
        if (with_vectors)
            HNSW_CUDA_CHECK(cudaMemcpy2DAsync(g.vectors.data(), (size_t)g.dim * 4, d_vectors_, (size_t)pitch_ * 4, (size_t)g.dim * 4, (size_t)length_,
                                              cudaMemcpyDeviceToHost, stream_));
        HNSW_CUDA_CHECK(cudaMemcpyAsync(g.adj0.data(), d_adj0_, g.adj0.size() * 4, cudaMemcpyDeviceToHost, stream_));
        if (!g.adjU.empty()) HNSW_CUDA_CHECK(cudaMemcpyAsync(g.adjU.data(), d_adjU_, g.adjU.size() * 4, cudaMemcpyDeviceToHost, stream_));
        HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    }
}

// ---------------------------------------------------------------------------------------------
// Diagnostics, host side from the exported adjacency
// ---------------------------------------------------------------------------------------------
namespace {
const int* host_row(const HostGraph& g, int id, int layer, int& stride) {
    if (layer == 0) { stride = g.stride0; return g.adj0.data() + (size_t)id * g.stride0; }
    stride = g.strideU;
    return g.adjU.data() + ((size_t)g.upper_off[id] + (layer - 1)) * g.strideU;
}
int median_like_reference(std::vector<int>& v) {  // HNSWInfo.cs:45-51
    std::sort(v.begin(), v.end());
    const size_t n = v.size();
    return n % 2 == 0 ? (v[n / 2 - 1] + v[n / 2]) / 2 : v[n / 2];
}
}  // namespace

std::vector<Index::LayerInfo> Index::get_info() {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    std::vector<LayerInfo> out;
    if (count() <= 0) return out;
    HostGraph g;
    export_graph(g, false);
    const bool has_in = params_.AllowRemovals;
    for (int layer = 0; layer <= top_; ++layer) {
        std::vector<int> outdeg, indeg_all(g.length, 0), ids;
        for (int i = 0; i < g.length; ++i) {
            if (g.levels[i] < layer || removed_[i]) continue;  // data.ActiveIds only (HNSWIndex.cs:194)
            int stride;
            const int* row = host_row(g, i, layer, stride);
            int d = 0;
            while (d < stride && row[d] >= 0) { indeg_all[row[d]]++; ++d; }
            outdeg.push_back(d);
            ids.push_back(i);
        }
        LayerInfo li{};
        li.LayerId = layer;
        li.NodesCount = (int)ids.size();
        if (ids.empty()) { out.push_back(li); continue; }
        std::vector<int> indeg;
        for (int id : ids) indeg.push_back(indeg_all[id]);
        auto stats = [](std::vector<int>& v, int& mx, int& mn, double& avg, int& med) {
            mx = *std::max_element(v.begin(), v.end());
            mn = *std::min_element(v.begin(), v.end());
            long long sum = 0;
            for (int x : v) sum += x;
            avg = (double)sum / (double)v.size();
            med = median_like_reference(v);
        };
        stats(outdeg, li.MaxOutEdges, li.MinOutEdges, li.AvgOutEdges, li.OutEdgesMedian);
        if (has_in) stats(indeg, li.MaxInEdges, li.MinInEdges, li.AvgInEdges, li.InEdgesMedian);
        out.push_back(li);
    }
    return out;
}

std::vector<int> Index::connected_component_counts() {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    std::vector<int> counts;
    if (count() <= 0) return counts;
    HostGraph g;
    export_graph(g, false);
    counts.assign(top_ + 1, 0);
    std::vector<int> parent(g.length);
    for (int layer = 0; layer <= top_; ++layer) {
        for (int i = 0; i < g.length; ++i) parent[i] = i;
        auto find = [&](int x) {
            while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
            return x;
        };
        for (int i = 0; i < g.length; ++i) {
            if (g.levels[i] < layer || removed_[i]) continue;
            int stride;
            const int* row = host_row(g, i, layer, stride);
            for (int j = 0; j < stride && row[j] >= 0; ++j) {
                if (g.levels[row[j]] < layer) continue;  // GraphNavigator.cs:372 nodesOnLayerSet.Contains
                const int a = find(i), b = find(row[j]);
                if (a != b) parent[a] = b;
            }
        }
        int c = 0;
        for (int i = 0; i < g.length; ++i)
            if (g.levels[i] >= layer && !removed_[i] && find(i) == i) ++c;
        counts[layer] = c;
    }
    return counts;
}

void Index::export_graph_locked(HostGraph& g, bool with_vectors) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);
    export_graph(g, with_vectors);
}

void Index::serialize(const std::string& path, bool packed) {
    Exclusive lock(mu_);
    check_usable();
    join_aux_locked(true);  // batches still queued on the auxiliary stream (knn_query_device_async) finish first
    HostGraph g;
    export_graph(g);
    write_snapshot(g, path, packed);
}

void Index::upload_graph(const HostGraph& g) {
    const int n = g.length;
    if (n == 0) return;
    if (g.entry < 0 && g.active.empty() && g.removed.empty()) throw std::runtime_error("snapshot: no entry point");
    ensure_dim(g.dim);
    capacity_ = std::max(capacity_, std::max(g.capacity, 1));
    ensure_capacity(n);
    levels_ = g.levels;
    upper_off_ = g.upper_off;
    upper_rows_ = (int)(g.adjU.size() / (size_t)strideU_);
    for (int i = 0; i < n; ++i) {
        if (levels_[i] > 30) throw std::runtime_error("snapshot: node level above 30 is not supported");
        const int rows = levels_[i];
        if (rows > 0 && (upper_off_[i] < 0 || upper_off_[i] + rows > upper_rows_)) throw std::runtime_error("snapshot: inconsistent upper-layer rows");
    }
    // every link at layer l must point at a node that exists on layer l
    for (int i = 0; i < n; ++i)
        for (int l = 1; l <= levels_[i]; ++l) {
            const int* row = g.adjU.data() + ((size_t)upper_off_[i] + (l - 1)) * strideU_;
            for (int j = 0; j < strideU_ && row[j] >= 0; ++j)
                if (levels_[row[j]] < l) throw std::runtime_error("snapshot: upper-layer link to a node that is not on that layer");
        }
    ensure_upper_rows(upper_rows_);
    HNSW_CUDA_CHECK(cudaMemcpy2DAsync(d_vectors_, (size_t)pitch_ * 4, g.vectors.data(), (size_t)dim_ * 4, (size_t)dim_ * 4, (size_t)n,
                                      cudaMemcpyHostToDevice, stream_));
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_adj0_, g.adj0.data(), g.adj0.size() * 4, cudaMemcpyHostToDevice, stream_));
    if (!g.adjU.empty()) HNSW_CUDA_CHECK(cudaMemcpyAsync(d_adjU_, g.adjU.data(), g.adjU.size() * 4, cudaMemcpyHostToDevice, stream_));
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_levels_, levels_.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
    HNSW_CUDA_CHECK(cudaMemcpyAsync(d_upper_off_, upper_off_.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
    if (metric_ == COSINE) launch_item_norms(d_vectors_, pitch_, dim_, 0, n, d_norms_, stream_);
    HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    length_ = n;
    entry_ = g.entry;
    top_ = entry_ >= 0 ? levels_[entry_] : -1;
    // removal state
    removed_.assign(n, 0);
    if (!g.removed.empty()) removed_.assign(g.removed.begin(), g.removed.end());
    removed_stack_ = g.removed_stack;
    active_dense_.clear();
    active_sparse_.assign(n, -1);
    if (!g.active.empty() || !g.removed.empty()) {
        for (int id : g.active) {
            if (id < 0 || id >= n || removed_[id] || active_sparse_[id] >= 0) throw std::runtime_error("snapshot: inconsistent ActiveNodes");
            active_add(id);
        }
    } else {
        for (int i = 0; i < n; ++i) active_add(i);
    }
    std::vector<int> rm_index(n, -1);
    int nremoved = 0;
    for (int i = 0; i < n; ++i)
        if (removed_[i]) { rm_index[i] = -2; ++nremoved; }
    if (nremoved + (int)active_dense_.size() != n) throw std::runtime_error("snapshot: ActiveNodes and IsRemoved flags disagree");
    if (nremoved > 0) {
        HNSW_CUDA_CHECK(cudaMemcpyAsync(d_rm_index_, rm_index.data(), (size_t)n * 4, cudaMemcpyHostToDevice, stream_));
        HNSW_CUDA_CHECK(cudaStreamSynchronize(stream_));
    }
    if (entry_ >= 0 && removed_[entry_]) throw std::runtime_error("snapshot: the entry point is a removed node");
}

Index* Index::from_snapshot(MetricKind metric, const std::string& path) {
    HostGraph g;
    read_snapshot(path, g);
    Index* ix = new Index(metric, g.params);
    try {
        if (g.stride0 != ix->stride0_ || g.strideU != ix->strideU_) throw std::runtime_error("snapshot: stride mismatch");
        ix->upload_graph(g);
    } catch (...) {
        delete ix;
        throw;
    }
    return ix;
}

}  // namespace hnsw
