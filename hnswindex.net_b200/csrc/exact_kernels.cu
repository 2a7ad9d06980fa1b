// exact_kernels.cu -- K3: exact brute-force top-k (recall ground truth) and K7: top-k list merge.
//
// K3 has no counterpart in the reference (its tests measure self-recall only); BASELINE.json's
// north_star asks for it as the recall ground truth.  It uses the same bit-exact metric
// arithmetic as the search (metric.cuh), so exact distances equal what KnnQuery would report.
// Each lane pair streams one stored vector (float4 per lane per 8-element chunk) and evaluates
// it against TQ queries held in shared memory, so every gathered byte is reused TQ times.
//
// K7 merges per-part sorted top-k lists (graph partitions on different GPUs, or data chunks of K3).
#include <math_constants.h>

#include <algorithm>

#include "beam.cuh"
#include "kernels.hpp"

namespace hnsw {

constexpr int TQ = 8;  // queries per block tile

template <int METRIC>
__global__ void __launch_bounds__(kThreads)
exact_scan_kernel(GraphView g, const float* __restrict__ queries, int nq, int k, int chunk_items, u64* __restrict__ part_keys /*[nq][chunks][k]*/) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane & 1, grp = lane >> 1;
    const int chunk = blockIdx.x, chunks = gridDim.x;
    const int q0 = blockIdx.y * TQ;
    const int pitch4 = (g.pitch + 3) & ~3;
    float* qs = reinterpret_cast<float*>(smem);                          // [TQ][pitch4]
    double* qn = reinterpret_cast<double*>(qs + (size_t)TQ * pitch4);     // [TQ]
    u64* lists = reinterpret_cast<u64*>(qn + TQ);                          // [warps][TQ][k]
    int* ln = reinterpret_cast<int*>(lists + (size_t)kWarpsPerBlock * TQ * k);  // [warps][TQ] counts

    for (int i = threadIdx.x; i < TQ * pitch4; i += blockDim.x) {
        const int t = i / pitch4, d = i - t * pitch4;
        qs[i] = (q0 + t < nq && d < g.dim) ? __ldg(queries + (size_t)(q0 + t) * g.dim + d) : 0.f;
    }
    for (int i = threadIdx.x; i < kWarpsPerBlock * TQ; i += blockDim.x) ln[i] = 0;
    __syncthreads();
    if (METRIC == COSINE) {
        for (int t = warp; t < TQ; t += kWarpsPerBlock) {
            const double r = pair_norm_sqrt(qs + (size_t)t * pitch4, g.dim, h);
            if (lane == 0) qn[t] = r;
        }
        __syncthreads();
    }

    u64* mylists = lists + (size_t)warp * TQ * k;
    int* myn = ln + warp * TQ;
    const int begin = chunk * chunk_items, end = min(g.count, begin + chunk_items);
    const int nch = g.dim >> 3;

    for (int base = begin + warp * 16; base < end; base += 16 * kWarpsPerBlock) {
        const int id = base + grp;
        const bool active = id < end;
        const float* vec = g.vectors + (size_t)(active ? id : begin) * g.pitch;
        const float4* v4 = reinterpret_cast<const float4*>(vec) + h;
        float acc[TQ][4];
#pragma unroll
        for (int t = 0; t < TQ; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
        if (active) {
            for (int c = 0; c < nch; ++c) {
                const float4 a = ldg_f4(v4 + 2 * c);
#pragma unroll
                for (int t = 0; t < TQ; ++t) {
                    const float4 b = reinterpret_cast<const float4*>(qs + (size_t)t * pitch4)[2 * c + h];
                    if (metric_is_sq(METRIC)) {
                        float d;
                        d = __fsub_rn(a.x, b.x); acc[t][0] = __fmaf_rn(d, d, acc[t][0]);
                        d = __fsub_rn(a.y, b.y); acc[t][1] = __fmaf_rn(d, d, acc[t][1]);
                        d = __fsub_rn(a.z, b.z); acc[t][2] = __fmaf_rn(d, d, acc[t][2]);
                        d = __fsub_rn(a.w, b.w); acc[t][3] = __fmaf_rn(d, d, acc[t][3]);
                    } else {
                        acc[t][0] = __fadd_rn(acc[t][0], __fmul_rn(a.x, b.x));
                        acc[t][1] = __fadd_rn(acc[t][1], __fmul_rn(a.y, b.y));
                        acc[t][2] = __fadd_rn(acc[t][2], __fmul_rn(a.z, b.z));
                        acc[t][3] = __fadd_rn(acc[t][3], __fmul_rn(a.w, b.w));
                    }
                }
            }
        }
        double an = 0.0;
        if (METRIC == COSINE && active) an = __ldg(g.norm_sqrt + id);
        const bool live = active && __ldg(g.rm_index + id) == -1;  // removed items (HNSWIndex.Remove) are not candidates
#pragma unroll
        for (int t = 0; t < TQ; ++t) {
            const float s0 = __fadd_rn(acc[t][0], __shfl_xor_sync(FULL, acc[t][0], 1));
            const float s1 = __fadd_rn(acc[t][1], __shfl_xor_sync(FULL, acc[t][1], 1));
            const float s2 = __fadd_rn(acc[t][2], __shfl_xor_sync(FULL, acc[t][2], 1));
            const float s3 = __fadd_rn(acc[t][3], __shfl_xor_sync(FULL, acc[t][3], 1));
            float r = metric_is_sq(METRIC) ? __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3)) : __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));
            if (active) {
                const float* qt = qs + (size_t)t * pitch4;
                for (int i = nch << 3; i < g.dim; ++i) {
                    const float a = __ldg(vec + i), b = qt[i];
                    if (metric_is_sq(METRIC)) { const float d = __fsub_rn(a, b); r = __fadd_rn(r, __fmul_rn(d, d)); }
                    else r = __fadd_rn(r, __fmul_rn(a, b));
                }
                if (METRIC == COSINE) r = cosine_finish(r, an, qn[t]);
                else if (METRIC == UCOSINE) r = __fsub_rn(1.0f, r);
                else if (METRIC == EUCLID) r = __fsqrt_rn(r);
            }
            // insert into this warp's list of query t
            BeamState st;
            st.n = myn[t];
            st.ef = k;
            st.cursor = 0;
            u64* list = mylists + (size_t)t * k;
            st.far_ord = st.n == k ? key_ord(list[k - 1]) : 0xFFFFFFFFu;
            const u64 mykey = make_key(r, id);
            bool pending = live && h == 0 && (q0 + t < nq);
            for (;;) {
                const bool want = pending && (st.n < k || mykey < list[k - 1]);
                const unsigned mk = __ballot_sync(FULL, want);
                if (!mk) break;
                const int src = __ffs(mk) - 1;
                const u64 kn = __shfl_sync(FULL, mykey, src);
                if (lane == src) pending = false;
                list_insert(list, st, kn, lane);
            }
            __syncwarp();
            if (lane == 0) myn[t] = st.n;
            __syncwarp();
        }
    }
    __syncthreads();
    // fold the other warps' lists into warp 0's, then write the block's partial result
    for (int t = warp; t < TQ; t += kWarpsPerBlock) {
        if (q0 + t >= nq) continue;
        u64* list = lists + (size_t)t * k;  // warp 0's list of query t
        BeamState st;
        st.n = ln[t];
        st.ef = k;
        st.cursor = 0;
        st.far_ord = st.n == k ? key_ord(list[k - 1]) : 0xFFFFFFFFu;
        for (int w = 1; w < kWarpsPerBlock; ++w) {
            const u64* other = lists + ((size_t)w * TQ + t) * k;
            const int on = ln[w * TQ + t];
            for (int j = 0; j < on; ++j) {
                const u64 kn = other[j];
                if (st.n == k && kn >= list[k - 1]) break;  // sorted: the rest is larger
                list_insert(list, st, kn, lane);
            }
        }
        __syncwarp();
        u64* out = part_keys + ((size_t)(q0 + t) * chunks + chunk) * k;
        for (int j = lane; j < k; j += 32) out[j] = j < st.n ? list[j] : ~0ull;
    }
}

// one warp per query: merge `parts` sorted key lists of length k
__global__ void __launch_bounds__(kThreads)
merge_keys_kernel(const u64* __restrict__ part_keys, int parts, int nq, int k, int* __restrict__ out_ids, float* __restrict__ out_dists) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kWarpsPerBlock + warp;
    if (q >= nq) return;
    u64* list = reinterpret_cast<u64*>(smem) + (size_t)warp * k;
    BeamState st;
    st.n = 0; st.ef = k; st.cursor = 0; st.far_ord = 0xFFFFFFFFu;
    const u64* src = part_keys + (size_t)q * parts * k;
    const int total = parts * k;
    for (int base = 0; base < total; base += 32) {
        const int i = base + lane;
        const u64 mykey = i < total ? src[i] : ~0ull;
        bool pending = mykey != ~0ull;
        for (;;) {
            const bool want = pending && (st.n < k || mykey < list[k - 1]);
            const unsigned mk = __ballot_sync(FULL, want);
            if (!mk) break;
            const int s = __ffs(mk) - 1;
            const u64 kn = __shfl_sync(FULL, mykey, s);
            if (lane == s) pending = false;
            list_insert(list, st, kn, lane);
        }
    }
    __syncwarp();
    for (int j = lane; j < k; j += 32) {
        int id = -1;
        float d = CUDART_NAN_F;
        if (j < st.n) { id = key_id(list[j]); d = key_dist(list[j]); }
        out_ids[(size_t)q * k + j] = id;
        out_dists[(size_t)q * k + j] = d;
    }
}

// [parts][nq][k] (id, dist) lists -> [nq][parts][k] keys
__global__ void pack_keys_kernel(const int* __restrict__ ids, const float* __restrict__ dists, int parts, int nq, int k, u64* __restrict__ keys) {
    const size_t total = (size_t)parts * nq * k;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % k);
        const size_t pq = i / k;
        const int q = (int)(pq % nq), p = (int)(pq / nq);
        const int id = ids[i];
        keys[((size_t)q * parts + p) * k + j] = id < 0 ? ~0ull : make_key(dists[i], id);
    }
}

static int exact_chunks(int count) {
    // enough chunks to give every SM work on small query batches, few enough to keep the partials small
    return std::max(1, std::min(32, (count + 4095) / 4096));
}

size_t exact_knn_scratch_bytes(int nq, int k, int count) { return (size_t)nq * exact_chunks(count) * k * sizeof(u64); }

void launch_exact_knn(int metric, const GraphView& g, const float* d_queries, int nq, int k, int* d_ids, float* d_dists, void* d_scratch,
                      size_t scratch_bytes, cudaStream_t stream) {
    if (nq <= 0) return;
    const int chunks = exact_chunks(g.count);
    if (scratch_bytes < exact_knn_scratch_bytes(nq, k, g.count)) throw std::runtime_error("exact_knn: scratch too small");
    const int chunk_items = (g.count + chunks - 1) / chunks;
    const int pitch4 = (g.pitch + 3) & ~3;
    const size_t smem = (size_t)TQ * pitch4 * 4 + TQ * 8 + (size_t)kWarpsPerBlock * TQ * k * 8 + kWarpsPerBlock * TQ * 4;
    dim3 grid(chunks, (nq + TQ - 1) / TQ);
    u64* keys = reinterpret_cast<u64*>(d_scratch);
#define HNSW_LAUNCH_EXACT(M)                                                                                     \
    {                                                                                                            \
        HNSW_CUDA_CHECK(cudaFuncSetAttribute(exact_scan_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        exact_scan_kernel<M><<<grid, kThreads, smem, stream>>>(g, d_queries, nq, k, chunk_items, keys);         \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_LAUNCH_EXACT(SQ_EUCLID) break;
        case EUCLID: HNSW_LAUNCH_EXACT(EUCLID) break;
        case COSINE: HNSW_LAUNCH_EXACT(COSINE) break;
        default: HNSW_LAUNCH_EXACT(UCOSINE) break;
    }
#undef HNSW_LAUNCH_EXACT
    HNSW_CUDA_CHECK(cudaGetLastError());
    const size_t msmem = (size_t)kWarpsPerBlock * k * 8;
    HNSW_CUDA_CHECK(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    merge_keys_kernel<<<(nq + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, msmem, stream>>>(keys, chunks, nq, k, d_ids, d_dists);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_merge_keys(const u64* d_part_keys, int parts, int nq, int k, int* d_out_ids, float* d_out_dists, cudaStream_t stream) {
    if (nq <= 0) return;
    const size_t msmem = (size_t)kWarpsPerBlock * k * 8;
    HNSW_CUDA_CHECK(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    merge_keys_kernel<<<(nq + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, msmem, stream>>>(d_part_keys, parts, nq, k, d_out_ids, d_out_dists);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_merge_topk(const int* d_in_ids, const float* d_in_dists, int parts, int nq, int k, int* d_out_ids, float* d_out_dists,
                       cudaStream_t stream) {
    if (nq <= 0) return;
    u64* keys = nullptr;
    HNSW_CUDA_CHECK(cudaMallocAsync(&keys, (size_t)parts * nq * k * sizeof(u64), stream));
    const size_t total = (size_t)parts * nq * k;
    const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 8);
    pack_keys_kernel<<<blocks, 256, 0, stream>>>(d_in_ids, d_in_dists, parts, nq, k, keys);
    HNSW_CUDA_CHECK(cudaGetLastError());
    const size_t msmem = (size_t)kWarpsPerBlock * k * 8;
    HNSW_CUDA_CHECK(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    merge_keys_kernel<<<(nq + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, msmem, stream>>>(keys, parts, nq, k, d_out_ids, d_out_dists);
    HNSW_CUDA_CHECK(cudaGetLastError());
    HNSW_CUDA_CHECK(cudaFreeAsync(keys, stream));
}

}  // namespace hnsw
