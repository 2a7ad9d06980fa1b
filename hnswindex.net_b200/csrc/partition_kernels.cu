// partition_kernels.cu -- partitioned index (SURVEY 8e mode 2): search fused with the exchange of the per-shard top-k.
//
// Every GPU owns one sub-graph and searches ALL queries on it.  Instead of "search, then ncclAllGather, then merge",
// the search kernel's epilogue stores a query's local top-k (ids shifted to global ids) straight into EVERY peer's
// gather buffer over NVLink (peer pointers obtained with CUDA IPC), so the transfer of query i overlaps the search
// of query i+1 and no collective is launched.  Completion is signalled with one flag per (peer, source rank):
//   every warp:   result stores to all peers -> __threadfence_system() -> atomicAdd(done)
//   last warp:    flags[peer][rank] = epoch  (system-scope release) for every peer
//   wait+merge:   one kernel, every warp polls flags[rank][*] >= epoch (acquire, wall-clock deadline), then merges its query.
// Buffers are double-buffered by epoch parity; a rank cannot run two epochs ahead of a peer because each epoch's wait
// needs that peer's flag of the same epoch, which the peer raises only after finishing the previous epoch's merge
// (stream order).
#include <math_constants.h>

#include <algorithm>

#include "beam.cuh"
#include "kernels.hpp"
#include "launch.hpp"

namespace hnsw {

template <int METRIC, int SPEC>
__global__ void __launch_bounds__(kThreads, 5)
knn_scatter_kernel(GraphView g, const float* __restrict__ queries, int nq, int k, int ef, int vis_size, int cand_cap, int flags,
                   ScatterArgs sc, unsigned* work_counter, unsigned* done_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = warp_scratch_bytes(g.pitch, ef, vis_size, cand_cap);
    WarpScratch s = carve_scratch(smem + warp * per_warp, g.pitch, ef, vis_size, cand_cap, lane);

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
        for (int l = g.top_layer; l > 0; --l) greedy_layer<METRIC, SPEC>(g, s, qn, l, cur, cur_d, lane, wc);
        BeamState st;
        beam_search<METRIC, SPEC>(g, s, qn, 0, cur, cur_d, ef, st, lane, wc, flags);

        // epilogue: this shard's top-k of query qi, global ids, into slot `rank` of every peer's gather buffer
        const size_t row = ((size_t)sc.rank * nq + qi) * k;
        for (int j = lane; j < k; j += 32) {
            int id = -1;
            float d = CUDART_NAN_F;
            if (j < st.n) {
                const u64 e = s.list[j];
                id = key_id(e) + sc.id_offset;
                d = key_dist(e);
            }
            for (int p = 0; p < sc.world; ++p) {
                sc.ids[p][row + j] = id;
                sc.dists[p][row + j] = d;
            }
        }
        __syncwarp();
    }
    // signal: all of this warp's peer stores are ordered before its arrival; the last warp raises the flags
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        const unsigned total = gridDim.x * (blockDim.x >> 5);
        if (atomicAdd(done_counter, 1u) == total - 1u) {
            *done_counter = 0u;  // ready for the next launch (stream-ordered)
            __threadfence_system();
            for (int p = 0; p < sc.world; ++p) {
                volatile unsigned* f = sc.flags[p] + sc.rank;
                *f = sc.epoch;
            }
            __threadfence_system();
        }
    }
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Wait + merge in ONE kernel, one warp per query: lane p polls rank p's flag of this epoch (acquire, system scope) until it
// arrives or `timeout_ns` of wall time have passed (%globaltimer: independent of clocks and of how long a poll takes), then
// the warp merges the `world` sorted top-k lists of its query straight out of the gather buffer (peer stores land in this
// GPU's L2; they are read past L1) -- no separate wait, pack and merge launches, no scratch allocation per batch.
// A peer that never arrives does not hang the GPU and does not leak stale data: *error is set and every output row of the
// batch is poisoned (-1 / NaN).  Once *error is set, later batches poison immediately instead of waiting again: after a
// timeout the epoch protocol is out of step (the late peer may still write into a buffer parity the waiter has moved past)
// and the exchange has to be set up again (PartitionedIndex.enable_fused + barrier).
__global__ void __launch_bounds__(kThreads)
wait_merge_kernel(const unsigned* flags, int world, unsigned epoch, unsigned long long timeout_ns, int* error, const int* in_ids,
                  const float* in_dists, int nq, int k, int* __restrict__ out_ids, float* __restrict__ out_dists) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * kWarpsPerBlock + warp;
    if (q >= nq) return;
    bool ok = *reinterpret_cast<volatile int*>(error) == 0;
    if (ok && lane < world) {
        const unsigned* f = flags + lane;
        const unsigned long long t0 = global_ns();
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int)(v - epoch) >= 0) break;  // a fast peer may already have raised the next epoch
            if (global_ns() - t0 > timeout_ns) { ok = false; break; }
            __nanosleep(100);
        }
    }
    ok = __all_sync(FULL, ok);
    if (!ok) {
        if (lane == 0) atomicExch(error, 1);
        for (int j = lane; j < k; j += 32) {
            out_ids[(size_t)q * k + j] = -1;
            out_dists[(size_t)q * k + j] = CUDART_NAN_F;
        }
        return;
    }
    u64* list = reinterpret_cast<u64*>(smem) + (size_t)warp * k;
    BeamState st;
    st.n = 0; st.ef = k; st.cursor = 0; st.far_ord = 0xFFFFFFFFu;
    const int total = world * k;
    for (int base = 0; base < total; base += 32) {
        const int i = base + lane;
        u64 mykey = ~0ull;
        if (i < total) {
            const int p = i / k, j = i - p * k;
            const size_t at = ((size_t)p * nq + q) * k + j;
            const int id = __ldcg(in_ids + at);
            if (id >= 0) mykey = make_key(__ldcg(in_dists + at), id);
        }
        bool pending = mykey != ~0ull;
        for (;;) {
            const bool want = pending && (st.n < k || mykey < list[k - 1]);
            const unsigned mk = __ballot_sync(FULL, want);
            if (!mk) break;
            const int src = __ffs(mk) - 1;
            const u64 kn = __shfl_sync(FULL, mykey, src);
            if (lane == src) pending = false;
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

template <typename Kern>
static int scatter_grid(Kern kern, size_t smem_per_warp, const Tuning& t, int work_items, int& warps) {
    return persistent_launch(kern, smem_per_warp, t.blocks_per_sm, work_items, warps);
}

int pick_vis_size(int requested, int ef);  // search_kernels.cu

void launch_knn_scatter(int metric, const GraphView& g, const float* d_queries, int nq, int k, int ef, const ScatterArgs& sc,
                        unsigned* d_work_counter, unsigned* d_done_counter, const Tuning& t, cudaStream_t stream) {
    if (nq <= 0) return;
    if (sc.world < 1 || sc.world > kMaxPeers) throw std::invalid_argument("scatter: world size out of range");
    const int vis = pick_vis_size(t.vis_size_search, ef);
    const int cand_cap = ((std::max(g.stride0, g.strideU) + 31) / 32) * 32;
    const size_t per_warp = warp_scratch_bytes(g.pitch, ef, vis, cand_cap);
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_work_counter, 0, sizeof(unsigned), stream));
#define HNSW_LAUNCH_SCATTER(M)                                                                                                 \
    {                                                                                                                          \
        int warps = 0;                                                                                                         \
        if (g.pitch <= kSegFloats && ef <= 64 && (long long)g.count <= ((long long)vis << 16)) {                               \
            constexpr int S = kSpecShortRows | kSpecSmallList | kSpecVis16;                                                    \
            const int grid = scatter_grid(knn_scatter_kernel<M, S>, per_warp, t, nq, warps);                                  \
            knn_scatter_kernel<M, S><<<grid, 32 * warps, per_warp * warps, stream>>>(g, d_queries, nq, k, ef, vis, cand_cap,  \
                                                                                    t.flags, sc, d_work_counter, d_done_counter); \
        } else {                                                                                                               \
            const int grid = scatter_grid(knn_scatter_kernel<M, 0>, per_warp, t, nq, warps);                                  \
            knn_scatter_kernel<M, 0><<<grid, 32 * warps, per_warp * warps, stream>>>(g, d_queries, nq, k, ef, vis, cand_cap,  \
                                                                                    t.flags, sc, d_work_counter, d_done_counter); \
        }                                                                                                                      \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_LAUNCH_SCATTER(SQ_EUCLID) break;
        case EUCLID: HNSW_LAUNCH_SCATTER(EUCLID) break;
        case COSINE: HNSW_LAUNCH_SCATTER(COSINE) break;
        default: HNSW_LAUNCH_SCATTER(UCOSINE) break;
    }
#undef HNSW_LAUNCH_SCATTER
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_wait_merge(const unsigned* d_flags, int world, unsigned epoch, int timeout_ms, int* d_error, const int* d_in_ids,
                       const float* d_in_dists, int nq, int k, int* d_out_ids, float* d_out_dists, cudaStream_t stream) {
    if (nq <= 0) return;
    if (world < 1 || world > 32) throw std::invalid_argument("wait_merge: world size out of range");
    const size_t msmem = (size_t)kWarpsPerBlock * k * 8;
    if (msmem > 48 * 1024)  // (per device and per call: only k > 1536 needs the opt-in at all)
        HNSW_CUDA_CHECK(cudaFuncSetAttribute(wait_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
    const unsigned long long timeout_ns = (unsigned long long)std::max(1, timeout_ms) * 1000000ull;
    wait_merge_kernel<<<(nq + kWarpsPerBlock - 1) / kWarpsPerBlock, kThreads, msmem, stream>>>(d_flags, world, epoch, timeout_ns, d_error, d_in_ids,
                                                                                              d_in_dists, nq, k, d_out_ids, d_out_dists);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

}  // namespace hnsw
