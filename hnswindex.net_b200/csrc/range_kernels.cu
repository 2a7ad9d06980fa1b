// range_kernels.cu -- RangeQuery on sm_100a (first "next" row after KnnQuery / Add).
//
// Replaces HNSWIndex.RangeQuery / BatchRangeQuery (HNSWIndex.cs:144-168) and
// GraphNavigator.SearchLayerRange (GraphNavigator.cs:262-325).
// The reference pops candidates closest-first, but its candidate heap only ever receives nodes with
// d <= range (plus the entry), and the early break can only fire for the entry -- so every in-range
// node that is reachable from the entry through in-range nodes is expanded exactly once, whatever the
// order.  The result SET is therefore a plain graph traversal; the kernel runs it breadth-first over
// its own output list and the host sorts each result by distance (OrderBy, HNSWIndex.cs:154).
// The visited set must be exact here (nothing bounds the result), so it is an open-addressing table in
// global memory, one slice per query; a query whose result or table overflows is re-run with 4x space.
#include <algorithm>

#include "beam.cuh"
#include "kernels.hpp"

namespace hnsw {

__device__ __forceinline__ bool global_test_and_set(uint32_t* tab, uint32_t mask, int id) {
    const uint32_t tag = (uint32_t)id + 1u;
    uint32_t h = (tag * 0x9E3779B1u) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        const uint32_t old = atomicCAS(tab + h, 0u, tag);
        if (old == 0u) return true;
        if (old == tag) return false;
        h = (h + 1u) & mask;
    }
    return false;  // table full: caller's load-factor check reports the overflow
}

template <int METRIC>
__global__ void __launch_bounds__(kThreads, 5)
range_search_kernel(GraphView g, const float* __restrict__ queries, const int* __restrict__ qmap, int nwork, float range, int cap, int hash_size,
                    int cand_cap, u64* out_keys, uint32_t* hash, int* out_count, unsigned* work_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = warp_scratch_bytes(g.pitch, 1, 64, cand_cap);
    WarpScratch s = carve_scratch(smem + warp * per_warp, g.pitch, 1, 64, cand_cap, lane);

    for (;;) {
        unsigned wi = 0;
        if (lane == 0) wi = atomicAdd(work_counter, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= (unsigned)nwork) break;
        const int qi = qmap ? qmap[wi] : (int)wi;
        const float* qsrc = queries + (size_t)qi * g.dim;
        for (int i = lane; i < g.pitch; i += 32) s.q[i] = i < g.dim ? __ldg(qsrc + i) : 0.f;
        __syncwarp();
        double qn = 0.0;
        if (METRIC == COSINE) qn = pair_norm_sqrt(s.q, g.dim, lane & 1);

        WarpCounters wc;
        int cur = g.entry;
        float cur_d = single_dist<METRIC>(g, s, qn, cur, lane);
        for (int layer = g.top_layer; layer > 0; --layer) greedy_layer<METRIC>(g, s, qn, layer, cur, cur_d, lane, wc);

        u64* out = out_keys + (size_t)wi * cap;
        uint32_t* tab = hash + (size_t)wi * hash_size;
        const uint32_t mask = (uint32_t)hash_size - 1u;
        int n = 0, head = 0, nvis = 1;
        bool overflow = false;
        if (lane == 0) global_test_and_set(tab, mask, cur);
        bool entry_pending = !(cur_d <= range);  // GraphNavigator.cs:271: an out-of-range entry is still expanded
        if (!entry_pending) {
            if (lane == 0) out[0] = make_key(cur_d, cur);
            n = 1;
        }
        __syncwarp();
        while (!overflow) {
            int c;
            if (entry_pending) { c = cur; entry_pending = false; }
            else {
                if (head >= n) break;
                c = key_id(out[head]);
                ++head;
            }
            int stride;
            const int* row = row_ptr(g, c, 0, stride);
            int m = 0;
            for (int rb = 0; rb < stride; rb += 32) {
                const int slot = rb + lane;
                const int nb = slot < stride ? row[slot] : -1;
                bool fresh = false;
                if (nb >= 0) fresh = global_test_and_set(tab, mask, nb);
                const unsigned mk = __ballot_sync(FULL, fresh);
                if (fresh) s.cand_id[m + __popc(mk & ((1u << lane) - 1u))] = nb;
                m += __popc(mk);
            }
            __syncwarp();
            nvis += m;
            if ((size_t)nvis * 2 > (size_t)hash_size) { overflow = true; break; }
            if (m == 0) continue;
            batch_dist<METRIC>(g, s, qn, m, lane);
            for (int jb = 0; jb < m; jb += 32) {
                const int j = jb + lane;
                const bool in = j < m && s.cand_dist[j] <= range;  // GraphNavigator.cs:302
                const unsigned mk = __ballot_sync(FULL, in);
                const int pos = n + __popc(mk & ((1u << lane) - 1u));
                if (in && pos < cap) out[pos] = make_key(s.cand_dist[j], s.cand_id[j]);
                n += __popc(mk);
            }
            if (n > cap) overflow = true;
            __syncwarp();
        }
        if (lane == 0) out_count[wi] = overflow ? -1 : n;
        __syncwarp();
    }
}

// One warp per query: sorts the query's result keys ascending (bitonic sort in shared memory, up to kSortMax keys;
// larger results are copied as they are and flagged for the host) and writes them to their place in the dense output.
constexpr int kSortMax = 2048;
__global__ void __launch_bounds__(kThreads)
range_finalize_kernel(const u64* __restrict__ keys, int cap, const int* __restrict__ counts, const long long* __restrict__ offsets, int nwork,
                      u64* __restrict__ dense, int* __restrict__ unsorted_flag) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    u64* buf = reinterpret_cast<u64*>(smem) + (size_t)warp * kSortMax;
    for (int w = blockIdx.x * kWarpsPerBlock + warp; w < nwork; w += gridDim.x * kWarpsPerBlock) {
        const int n = counts[w];
        if (n <= 0) continue;
        const u64* src = keys + (size_t)w * cap;
        u64* dst = dense + offsets[w];
        if (n > kSortMax) {
            for (int i = lane; i < n; i += 32) dst[i] = src[i];
            if (lane == 0) unsorted_flag[w] = 1;
            continue;
        }
        int p2 = 32;
        while (p2 < n) p2 <<= 1;
        for (int i = lane; i < p2; i += 32) buf[i] = i < n ? src[i] : ~0ull;
        __syncwarp();
        for (int size = 2; size <= p2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int i = lane; i < (p2 >> 1); i += 32) {
                    const int lo = ((i / stride) * stride * 2) + (i % stride), hi = lo + stride;
                    const bool up = ((lo & size) == 0);
                    const u64 a = buf[lo], b = buf[hi];
                    if ((a > b) == up) { buf[lo] = b; buf[hi] = a; }
                }
                __syncwarp();
            }
        for (int i = lane; i < n; i += 32) dst[i] = buf[i];
        __syncwarp();
    }
}

void launch_range_finalize(const u64* d_keys, int cap, const int* d_counts, const long long* d_offsets, int nwork, u64* d_dense,
                           int* d_unsorted_flag, cudaStream_t stream) {
    if (nwork <= 0) return;
    const size_t smem = (size_t)kWarpsPerBlock * kSortMax * sizeof(u64);
    static bool configured = false;
    if (!configured) {
        HNSW_CUDA_CHECK(cudaFuncSetAttribute(range_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_unsorted_flag, 0, (size_t)nwork * sizeof(int), stream));
    const int grid = std::max(1, std::min((nwork + kWarpsPerBlock - 1) / kWarpsPerBlock, 148 * 3));
    range_finalize_kernel<<<grid, kThreads, smem, stream>>>(d_keys, cap, d_counts, d_offsets, nwork, d_dense, d_unsorted_flag);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_range_search(int metric, const GraphView& g, const float* d_queries, const int* d_qmap, int nwork, float range, int cap,
                         int hash_size, u64* d_out_keys, uint32_t* d_hash, int* d_out_count, unsigned* d_work_counter, const Tuning& t,
                         cudaStream_t stream) {
    if (nwork <= 0) return;
    const int cand_cap = ((std::max(g.stride0, g.strideU) + 31) / 32) * 32;
    const size_t smem = warp_scratch_bytes(g.pitch, 1, 64, cand_cap) * kWarpsPerBlock;
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_work_counter, 0, sizeof(unsigned), stream));
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_hash, 0, (size_t)nwork * hash_size * sizeof(uint32_t), stream));
    int dev = 0, sms = 0, per_sm = 0;
    HNSW_CUDA_CHECK(cudaGetDevice(&dev));
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
#define HNSW_LAUNCH_RANGE(M)                                                                                                           \
    {                                                                                                                                  \
        HNSW_CUDA_CHECK(cudaFuncSetAttribute(range_search_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
        HNSW_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, range_search_kernel<M>, kThreads, smem));               \
        if (t.blocks_per_sm > 0) per_sm = std::min(per_sm, t.blocks_per_sm);                                                           \
        const int grid = std::max(1, std::min((nwork + kWarpsPerBlock - 1) / kWarpsPerBlock, sms * std::max(per_sm, 1)));               \
        range_search_kernel<M><<<grid, kThreads, smem, stream>>>(g, d_queries, d_qmap, nwork, range, cap, hash_size, cand_cap, d_out_keys, \
                                                                 d_hash, d_out_count, d_work_counter);                                  \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_LAUNCH_RANGE(SQ_EUCLID) break;
        case EUCLID: HNSW_LAUNCH_RANGE(EUCLID) break;
        case COSINE: HNSW_LAUNCH_RANGE(COSINE) break;
        default: HNSW_LAUNCH_RANGE(UCOSINE) break;
    }
#undef HNSW_LAUNCH_RANGE
    HNSW_CUDA_CHECK(cudaGetLastError());
}

}  // namespace hnsw
