// search_kernels.cu -- K2: batched KnnQuery on sm_100a, one warp per query.
//
// Replaces HNSWIndex.BatchKnnQuery / KnnQuery (HNSWIndex.cs:107-137):
//   ef = max(MinNN, k); greedy descent over layers top..1 (GraphNavigator.cs:39-82);
//   layer-0 beam search of width ef (GraphNavigator.cs:194-256); ascending by distance,
//   first k (HNSWIndex.cs:119-123); rows padded with -1 / NaN (HNSWIndexExports.cs:135-145).
// Persistent grid: every warp pulls the next query index from a global counter, so
// long and short searches balance across the 148 SMs.
#include <math_constants.h>

#include <algorithm>
#include <mutex>
#include <set>

#include "wide.cuh"
#include "kernels.hpp"
#include "launch.hpp"

namespace hnsw {

// ANY_LAYER = false: the base-layer search of the headline path (layer is the constant 0 throughout);
// ANY_LAYER = true: KnnQuery(..., layer > 0), a separate instantiation so that the hot kernel stays untouched.
// SPEC: compile-time facts about the launch (beam.cuh kSpec*), so the C2-shaped hot kernel carries no dead code
// (its instruction footprint decides whether the loop stays inside the 32 KB L1.5 instruction cache).
template <int METRIC, bool ANY_LAYER, int SPEC>
__global__ void __launch_bounds__(kThreads, 5)
knn_search_kernel(GraphView g, const float* __restrict__ queries, int nq, int k, int ef, int layer, int vis_size, int cand_cap, int flags,
                  int* __restrict__ out_ids, float* __restrict__ out_dists, unsigned* work_counter, QueryStats* stats, uint32_t* gvis,
                  int gvis_slots) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = warp_scratch_bytes(g.pitch, ef, vis_size, cand_cap);
    WarpScratch s = carve_scratch(smem + warp * per_warp, g.pitch, ef, vis_size, cand_cap, lane);
    if (SPEC & kSpecGvis) {  // exact visited bitmap of this warp (long rows): gvis_slots words per warp
        s.vbm = gvis + ((size_t)blockIdx.x * (blockDim.x >> 5) + warp) * gvis_slots;
        s.vbm_words = gvis_slots;
    }

    for (;;) {
        unsigned qi = 0;
        if (lane == 0) qi = atomicAdd(work_counter, 1u);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= (unsigned)nq) break;

        const float* qsrc = queries + (size_t)qi * g.dim;
        for (int i = lane; i < g.pitch; i += 32) s.q[i] = i < g.dim ? __ldg(qsrc + i) : 0.f;
        __syncwarp();
        double qn = 0.0;
        if (METRIC == COSINE) qn = pair_norm_sqrt(s.q, g.dim, lane & 1);

        WarpCounters wc;
        int cur = g.entry;
        float cur_d = single_dist<METRIC, SPEC>(g, s, qn, cur, lane);
        wc.ndist++;
        const int target = ANY_LAYER ? layer : 0;
        for (int l = g.top_layer; l > target; --l) greedy_layer<METRIC, SPEC>(g, s, qn, l, cur, cur_d, lane, wc);

        int found;
        if ((SPEC & kSpecWideList) && !ANY_LAYER) {
            WideList L;
            beam_search_wl<METRIC, SPEC>(g, s, qn, cur, cur_d, ef, L, lane, wc, flags);
            found = L.n;
        } else {
            BeamState st;
            beam_search<METRIC, SPEC>(g, s, qn, target, cur, cur_d, ef, st, lane, wc, ANY_LAYER ? 0 : flags);
            found = st.n;
        }

        for (int j = lane; j < k; j += 32) {
            int id = -1;
            float d = CUDART_NAN_F;
            if (j < found) {
                const u64 e = s.list[j];
                id = key_id(e);
                d = key_dist(e);
            }
            out_ids[(size_t)qi * k + j] = id;
            out_dists[(size_t)qi * k + j] = d;
        }
        if (stats != nullptr && lane == 0) {
            QueryStats qs;
            qs.ndist = wc.ndist; qs.nexp = wc.nexp; qs.row_words = wc.row_words;
            stats[qi] = qs;
        }
        __syncwarp();
    }
}

// MultiLayerKnnQuery (HNSWIndex.cs:173-187), one warp per query: from layer `top_l` = min(top layer, maxLayer) down to
// min_layer one beam search of width k per layer, each entered from the closest result of the layer above; a layer's
// output row is its sorted result WITHOUT the first entry (at most k-1 entries, padded with -1 / NaN).
// out_* are [nq][top_l + 1][k]; rows of layers below min_layer are all padding.
template <int METRIC>
__global__ void __launch_bounds__(kThreads, 5)
multilayer_search_kernel(GraphView g, const float* __restrict__ queries, int nq, int k, int top_l, int min_layer, int vis_size, int cand_cap,
                         int* __restrict__ out_ids, float* __restrict__ out_dists, unsigned* work_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = warp_scratch_bytes(g.pitch, k, vis_size, cand_cap);
    WarpScratch s = carve_scratch(smem + warp * per_warp, g.pitch, k, vis_size, cand_cap, lane);
    const int layers = top_l + 1;

    for (;;) {
        unsigned qi = 0;
        if (lane == 0) qi = atomicAdd(work_counter, 1u);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= (unsigned)nq) break;

        const float* qsrc = queries + (size_t)qi * g.dim;
        for (int i = lane; i < g.pitch; i += 32) s.q[i] = i < g.dim ? __ldg(qsrc + i) : 0.f;
        __syncwarp();
        double qn = 0.0;
        if (METRIC == COSINE) qn = pair_norm_sqrt(s.q, g.dim, lane & 1);

        WarpCounters wc;
        int cur = g.entry;
        float cur_d = single_dist<METRIC>(g, s, qn, cur, lane);
        for (int l = g.top_layer; l > top_l; --l) greedy_layer<METRIC>(g, s, qn, l, cur, cur_d, lane, wc);  // FindEntryPointQuery(maxLayer)

        for (int layer = top_l; layer >= 0; --layer) {
            int* oi = out_ids + ((size_t)qi * layers + layer) * k;
            float* od = out_dists + ((size_t)qi * layers + layer) * k;
            int n = 0;
            if (layer >= min_layer) {
                BeamState st;
                beam_search<METRIC>(g, s, qn, layer, cur, cur_d, k, st, lane, wc, 0);
                n = st.n;
                cur = key_id(s.list[0]);
                cur_d = key_dist(s.list[0]);
            }
            for (int j = lane; j < k; j += 32) {
                int id = -1;
                float d = CUDART_NAN_F;
                if (j + 1 < n) {
                    const u64 e = s.list[j + 1];
                    id = key_id(e);
                    d = key_dist(e);
                }
                oi[j] = id;
                od[j] = d;
            }
            __syncwarp();
        }
    }
}

// one lane pair per item; 16 items per warp
__global__ void item_norms_kernel(const float* __restrict__ vectors, int pitch, int dim, int first, int count, double* __restrict__ norm_sqrt) {
    const int lane = threadIdx.x & 31;
    const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const bool active = pair < count;
    const int id = first + (active ? pair : 0);
    const double r = pair_norm_sqrt(vectors + (size_t)id * pitch, dim, lane & 1);
    if (active && (lane & 1) == 0) norm_sqrt[id] = r;
}

// the same for a list of ids (slots re-used by Add after removals)
__global__ void item_norms_ids_kernel(const float* __restrict__ vectors, int pitch, int dim, const int* __restrict__ ids, int count,
                                      double* __restrict__ norm_sqrt) {
    const int lane = threadIdx.x & 31;
    const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    const bool active = pair < count;
    const int id = __ldg(ids + (active ? pair : 0));
    const double r = pair_norm_sqrt(vectors + (size_t)id * pitch, dim, lane & 1);
    if (active && (lane & 1) == 0) norm_sqrt[id] = r;
}

// Add into re-used slots: one warp per new item copies row src_row[j] of the staged input batch into slab row ids[j]
// (padding floats zeroed) and records the item's level and first upper row.
__global__ void scatter_new_items_kernel(const float* __restrict__ src, const int* __restrict__ src_row, const int* __restrict__ ids,
                                         const int* __restrict__ lv, const int* __restrict__ uoff, int n, int dim, int pitch,
                                         float* __restrict__ vectors, int* __restrict__ levels, int* __restrict__ upper_off) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= n) return;
    const int id = __ldg(ids + j);
    const float* s = src + (size_t)__ldg(src_row + j) * dim;
    float* d = vectors + (size_t)id * pitch;
    for (int i = lane; i < pitch; i += 32) d[i] = i < dim ? __ldg(s + i) : 0.f;
    if (lane == 0) {
        levels[id] = __ldg(lv + j);
        upper_off[id] = __ldg(uoff + j);
    }
}

int device_sm_count() {
    int dev = 0, sms = 0;
    HNSW_CUDA_CHECK(cudaGetDevice(&dev));
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return sms;
}

static int pow2_at_least(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

int pick_vis_size(int requested, int ef) {
    if (requested > 0) return std::max(64, pow2_at_least(requested));
    // The table only saves re-evaluations; shared memory per warp decides how many searches an SM keeps in
    // flight, and the kernel is latency-bound: 4*ef slots (+9 % evaluations at C2) beat 16*ef (+2.5 %) by 9 %.
    return std::min(32768, std::max(256, pow2_at_least(4 * ef)));
}

// Persistent launch shape: `warps` independent warps per block (4, fewer when one warp's scratch is large),
// at most one resident wave of blocks.  Returns the grid size and sets `warps`.
template <typename Kern>
static int configure_persistent(Kern kern, size_t smem_per_warp, const Tuning& t, int work_items, int& warps) {
    return persistent_launch(kern, smem_per_warp, t.blocks_per_sm, work_items, warps);
}

// The visited bitmaps of the wide-beam and long-row searches are stream-ordered allocations of a few hundred MB per launch:
// let the device's default pool keep up to 1 GiB of freed memory instead of returning it at every synchronisation.
void retain_freed_scratch() {
    static std::mutex mu;
    static std::set<int> done;
    int dev = 0;
    HNSW_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (!done.insert(dev).second) return;
    cudaMemPool_t pool;
    HNSW_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, dev));
    uint64_t have = 0;
    HNSW_CUDA_CHECK(cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &have));
    uint64_t want = 1ull << 30;
    if (have < want) HNSW_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &want));
}

void launch_knn_search(int metric, const GraphView& g, const float* d_queries, int nq, int k, int ef, int layer, int* d_ids, float* d_dists,
                       unsigned* d_work_counter, QueryStats* d_stats, const Tuning& t, cudaStream_t stream) {
    if (nq <= 0) return;
    retain_freed_scratch();
    if (wide_search_applies(g, nq, ef, layer, t)) {
        launch_knn_wide(metric, g, d_queries, nq, k, ef, d_ids, d_dists, d_work_counter, d_stats, t, stream);
        return;
    }
    const int vis = pick_vis_size(t.vis_size_search, ef);
    // 16-bit tags hold two entries per word: for the mid-size lists the same filter fits half the words (more warps per SM)
    const int vis_mid = std::max(256, vis / 2);
    const int cand_cap = ((std::max(g.stride0, g.strideU) + 31) / 32) * 32;
    const size_t per_warp = warp_scratch_bytes(g.pitch, ef, vis, cand_cap);
    const size_t per_warp_mid = warp_scratch_bytes(g.pitch, ef, vis_mid, cand_cap);
    // exact visited bitmap for long-row searches: one bit per node per resident warp (1M nodes: 125 KB x ~1 800 warps)
    // vectors of exactly 128 floats: instantiations whose distance loops have compile-time trip counts (beam.cuh kSpecDim128);
    // not for the ef <= 64 kernel, which runs at the HBM roofline as it is
    const bool dim128 = t.dim128 != 0 && g.dim == 128 && g.pitch == 128;
    const int bitmap_words = (t.vis_global_search != 0 && g.count <= t.wide_bitmap_max_nodes) ? ((g.count + 127) / 128) * 4 : 0;
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_work_counter, 0, sizeof(unsigned), stream));
#define HNSW_LAUNCH_SEARCH(M)                                                                                                   \
    {                                                                                                                           \
        int warps = 0;                                                                                                          \
        if (layer == 0 && g.pitch <= kSegFloats && ef <= 64 && (long long)g.count <= ((long long)vis << 16)) {                  \
            constexpr int S = kSpecShortRows | kSpecSmallList | kSpecVis16;                                                     \
            const int grid = configure_persistent(knn_search_kernel<M, false, S>, per_warp, t, nq, warps);                     \
            knn_search_kernel<M, false, S><<<grid, 32 * warps, per_warp * warps, stream>>>(                                    \
                g, d_queries, nq, k, ef, 0, vis, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, nullptr, 0);                  \
        } else if (layer == 0 && t.wl_min_ef > 0 && ef >= t.wl_min_ef && g.pitch <= kSegFloats &&                               \
                   (long long)g.count <= ((long long)vis_mid << 16)) {                                                          \
            /* wide.cuh: sorted main array + 32 pending keys in registers, merged in one pass: no per-insertion shifting */    \
            if (dim128) {                                                                                                       \
                constexpr int S = kSpecShortRows | kSpecWideList | kSpecVis16 | kSpecDim128;                                    \
                const int grid = configure_persistent(knn_search_kernel<M, false, S>, per_warp_mid, t, nq, warps);             \
                knn_search_kernel<M, false, S><<<grid, 32 * warps, per_warp_mid * warps, stream>>>(                            \
                    g, d_queries, nq, k, ef, 0, vis_mid, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, nullptr, 0); \
            } else {                                                                                                            \
                constexpr int S = kSpecShortRows | kSpecWideList | kSpecVis16;                                                  \
                const int grid = configure_persistent(knn_search_kernel<M, false, S>, per_warp_mid, t, nq, warps);             \
                knn_search_kernel<M, false, S><<<grid, 32 * warps, per_warp_mid * warps, stream>>>(                            \
                    g, d_queries, nq, k, ef, 0, vis_mid, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, nullptr, 0); \
            }                                                                                                                   \
        } else if (layer == 0 && t.wl_min_ef > 0 && ef >= t.wl_min_ef && g.pitch > kSegFloats && bitmap_words == 0 &&          \
                   (long long)g.count <= ((long long)vis << 16)) {                                                              \
            constexpr int S = kSpecWideList | kSpecVis16;                                                                       \
            const int grid = configure_persistent(knn_search_kernel<M, false, S>, per_warp, t, nq, warps);                     \
            knn_search_kernel<M, false, S><<<grid, 32 * warps, per_warp * warps, stream>>>(                                    \
                g, d_queries, nq, k, ef, 0, vis, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, nullptr, 0);                  \
        } else if (layer == 0 && g.pitch <= kSegFloats && ef <= 1024 && (long long)g.count <= ((long long)vis_mid << 16)) {     \
            if (dim128) {                                                                                                       \
                constexpr int S = kSpecShortRows | kSpecMidList | kSpecVis16 | kSpecDim128;                                     \
                const int grid = configure_persistent(knn_search_kernel<M, false, S>, per_warp_mid, t, nq, warps);             \
                knn_search_kernel<M, false, S><<<grid, 32 * warps, per_warp_mid * warps, stream>>>(                            \
                    g, d_queries, nq, k, ef, 0, vis_mid, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, nullptr, 0); \
            } else {                                                                                                            \
                constexpr int S = kSpecShortRows | kSpecMidList | kSpecVis16;                                                   \
                const int grid = configure_persistent(knn_search_kernel<M, false, S>, per_warp_mid, t, nq, warps);             \
                knn_search_kernel<M, false, S><<<grid, 32 * warps, per_warp_mid * warps, stream>>>(                            \
                    g, d_queries, nq, k, ef, 0, vis_mid, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, nullptr, 0); \
            }                                                                                                                   \
        } else if (layer == 0 && g.pitch > kSegFloats && ef <= 1024 && (long long)g.count <= ((long long)vis << 16)) {          \
            /* long rows are DRAM-bound and every re-evaluation is real traffic (C3: 429 GB read for 393 GB of algorithmic      \
               bytes): 16-bit tags (same words, twice the entries) and behind them an exact per-warp visited bitmap in global  \
               memory whose look-up travels with the first segment of the row (beam.cuh batch_dist_probe) */                   \
            constexpr int S = kSpecMidList | kSpecVis16;                                                                        \
            if (bitmap_words > 0) {                                                                                             \
                const int grid = configure_persistent(knn_search_kernel<M, false, S | kSpecGvis>, per_warp, t, nq, warps);     \
                uint32_t* d_bitmap = nullptr;                                                                                   \
                HNSW_CUDA_CHECK(cudaMallocAsync(&d_bitmap, (size_t)grid * warps * bitmap_words * sizeof(uint32_t), stream));    \
                knn_search_kernel<M, false, S | kSpecGvis><<<grid, 32 * warps, per_warp * warps, stream>>>(                    \
                    g, d_queries, nq, k, ef, 0, vis, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, d_bitmap, bitmap_words); \
                const cudaError_t launch_err = cudaGetLastError(); /* back to the pool whether or not the launch took */        \
                HNSW_CUDA_CHECK(cudaFreeAsync(d_bitmap, stream));                                                               \
                HNSW_CUDA_CHECK(launch_err);                                                                                    \
            } else {                                                                                                            \
                const int grid = configure_persistent(knn_search_kernel<M, false, S>, per_warp, t, nq, warps);                 \
                knn_search_kernel<M, false, S><<<grid, 32 * warps, per_warp * warps, stream>>>(                                \
                    g, d_queries, nq, k, ef, 0, vis, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, nullptr, 0);  \
            }                                                                                                                   \
        } else if (layer == 0) {                                                                                                \
            const int grid = configure_persistent(knn_search_kernel<M, false, 0>, per_warp, t, nq, warps);                     \
            knn_search_kernel<M, false, 0><<<grid, 32 * warps, per_warp * warps, stream>>>(                                    \
                g, d_queries, nq, k, ef, 0, vis, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, nullptr, 0);                  \
        } else {                                                                                                                \
            const int grid = configure_persistent(knn_search_kernel<M, true, 0>, per_warp, t, nq, warps);                      \
            knn_search_kernel<M, true, 0><<<grid, 32 * warps, per_warp * warps, stream>>>(                                     \
                g, d_queries, nq, k, ef, layer, vis, cand_cap, t.flags, d_ids, d_dists, d_work_counter, d_stats, nullptr, 0);              \
        }                                                                                                                       \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_LAUNCH_SEARCH(SQ_EUCLID) break;
        case EUCLID: HNSW_LAUNCH_SEARCH(EUCLID) break;
        case COSINE: HNSW_LAUNCH_SEARCH(COSINE) break;
        default: HNSW_LAUNCH_SEARCH(UCOSINE) break;
    }
#undef HNSW_LAUNCH_SEARCH
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_multilayer_search(int metric, const GraphView& g, const float* d_queries, int nq, int k, int top_l, int min_layer, int* d_ids,
                              float* d_dists, unsigned* d_work_counter, const Tuning& t, cudaStream_t stream) {
    if (nq <= 0) return;
    const int vis = pick_vis_size(t.vis_size_search, k);
    const int cand_cap = ((std::max(g.stride0, g.strideU) + 31) / 32) * 32;
    const size_t per_warp = warp_scratch_bytes(g.pitch, k, vis, cand_cap);
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_work_counter, 0, sizeof(unsigned), stream));
#define HNSW_LAUNCH_ML(M)                                                                                                        \
    {                                                                                                                            \
        int warps = 0;                                                                                                           \
        const int grid = configure_persistent(multilayer_search_kernel<M>, per_warp, t, nq, warps);                             \
        multilayer_search_kernel<M><<<grid, 32 * warps, per_warp * warps, stream>>>(g, d_queries, nq, k, top_l, min_layer, vis, \
                                                                                    cand_cap, d_ids, d_dists, d_work_counter);  \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_LAUNCH_ML(SQ_EUCLID) break;
        case EUCLID: HNSW_LAUNCH_ML(EUCLID) break;
        case COSINE: HNSW_LAUNCH_ML(COSINE) break;
        default: HNSW_LAUNCH_ML(UCOSINE) break;
    }
#undef HNSW_LAUNCH_ML
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_item_norms_ids(const float* d_vectors, int pitch, int dim, const int* d_ids, int count, double* d_norm_sqrt, cudaStream_t stream) {
    if (count <= 0) return;
    const int threads = 256;
    item_norms_ids_kernel<<<(count * 2 + threads - 1) / threads, threads, 0, stream>>>(d_vectors, pitch, dim, d_ids, count, d_norm_sqrt);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_scatter_new_items(const float* d_src, const int* d_src_row, const int* d_ids, const int* d_lv, const int* d_uoff, int n, int dim,
                              int pitch, float* d_vectors, int* d_levels, int* d_upper_off, cudaStream_t stream) {
    if (n <= 0) return;
    const int threads = 256;
    const long long blocks = ((long long)n * 32 + threads - 1) / threads;
    scatter_new_items_kernel<<<(unsigned)blocks, threads, 0, stream>>>(d_src, d_src_row, d_ids, d_lv, d_uoff, n, dim, pitch, d_vectors, d_levels,
                                                                      d_upper_off);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_item_norms(const float* d_vectors, int pitch, int dim, int first, int count, double* d_norm_sqrt, cudaStream_t stream) {
    if (count <= 0) return;
    const int threads = 256;
    const int blocks = (count * 2 + threads - 1) / threads;
    item_norms_kernel<<<blocks, threads, 0, stream>>>(d_vectors, pitch, dim, first, count, d_norm_sqrt);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

}  // namespace hnsw
