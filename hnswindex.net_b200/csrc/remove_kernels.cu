// remove_kernels.cu -- batched Remove on sm_100a.
//
// Replaces HNSWIndex.Remove (HNSWIndex.cs:83-100) -> GraphConnector.RemoveNodeConnections /
// RemoveConnectionsAtLayer (GraphConnector.cs:53-167) for a CHUNK of removed nodes at once:
//   the reference, per removed node r and layer:   here, per chunk:
//     SearchLayer(r, layer, RemoveMaxCandidates,     remove_search_kernel: one warp per (r, layer), beam search from r of
//                 filter id != r)                      width RemoveMaxCandidates + 1 (r itself holds one slot), r and every
//                                                      node removed in the same chunk dropped from the result
//     affected = r.InEdges[layer]                    find_affected_kernel: the engine keeps no in-edge lists; one streaming
//                                                      pass over all rows finds every row that names a removed node
//     per affected node a: candidates = a's          relink_kernel: one warp per affected (a, layer) row: remaining
//       remaining neighbours + the search result,      neighbours + the search results of ALL removed neighbours of the row,
//       RelativeNeighborPruning, new out-row           same heuristic (Heuristic.cs:11-46), the row is rewritten once --
//                                                      each row has exactly one writer, so no locks are needed
//     (in-edge bookkeeping, GraphLocker)             none: in-edges are the transpose, produced when a snapshot is written
// With one node per chunk this is the reference's sequential Remove (same candidate sets, same selected SETS; the
// order inside a rewritten row differs, which only matters under exact distance ties).
#include <algorithm>

#include "build_scratch.cuh"
#include "kernels.hpp"
#include "launch.hpp"

namespace hnsw {

// rm_index[id]: -1 alive, >= 0 position in the chunk being removed, -2 removed earlier
__global__ void set_rm_index_kernel(int* rm_index, const int* __restrict__ ids, int n, bool use_index, int value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rm_index[ids[i]] = use_index ? i : value;
}

template <int METRIC>
__global__ void __launch_bounds__(kThreads, 4)
remove_search_kernel(GraphView g, const int* __restrict__ item_node, const int* __restrict__ item_layer, int nitems, int rmc, int list_cap,
                     int vis_size, int cand_cap, int* __restrict__ out_cand, unsigned* work_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = warp_scratch_bytes(g.pitch, list_cap, vis_size, cand_cap);
    WarpScratch s = carve_scratch(smem + warp * per_warp, g.pitch, list_cap, vis_size, cand_cap, lane);
    for (;;) {
        unsigned wi = 0;
        if (lane == 0) wi = atomicAdd(work_counter, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= (unsigned)nitems) break;
        const int r = __ldg(item_node + wi), layer = __ldg(item_layer + wi);
        load_vector(g, r, s.q, lane);
        double qn = 0.0;
        if (METRIC == COSINE) qn = __ldg(g.norm_sqrt + r);
        WarpCounters wc;
        const float d0 = single_dist<METRIC>(g, s, qn, r, lane);
        BeamState st;
        beam_search<METRIC>(g, s, qn, layer, r, d0, rmc + 1, st, lane, wc, 0);
        // result without r and without anything else removed in this chunk, ascending, at most rmc entries
        int* out = out_cand + (size_t)wi * rmc;
        int nout = 0;
        for (int base = 0; base < st.n; base += 32) {
            const int idx = base + lane;
            int id = -1;
            if (idx < st.n) id = key_id(s.list[idx]);
            const bool keep = id >= 0 && id != r && __ldg(g.rm_index + id) == -1;
            const unsigned mk = __ballot_sync(FULL, keep);
            const int pos = nout + __popc(mk & ((1u << lane) - 1u));
            if (keep && pos < rmc) out[pos] = id;
            nout += __popc(mk);
        }
        nout = min(nout, rmc);
        for (int j = nout + lane; j < rmc; j += 32) out[j] = -1;
        __syncwarp();
    }
}

// One warp per node (grid-stride): every row of the node that names a node of the chunk is appended to `affected`
// as (node, layer).  Rows of removed nodes are skipped.
__global__ void find_affected_kernel(GraphView g, const int* __restrict__ levels, int length, int2* __restrict__ affected, unsigned* n_affected) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < length; i += warps) {
        if (__ldg(g.rm_index + i) != -1) continue;
        const int level = __ldg(levels + i);
        for (int layer = 0; layer <= level; ++layer) {
            int stride;
            const int* row = row_ptr(g, i, layer, stride);
            bool hit = false;
            for (int rb = 0; rb < stride; rb += 32) {
                const int slot = rb + lane;
                const int nb = slot < stride ? row[slot] : -1;
                if (nb >= 0 && __ldg(g.rm_index + nb) >= 0) hit = true;
            }
            if (__any_sync(FULL, hit) && lane == 0) {
                const unsigned at = atomicAdd(n_affected, 1u);
                affected[at] = make_int2(i, layer);
            }
        }
    }
}

template <int METRIC>
__global__ void __launch_bounds__(kThreads, 4)
relink_kernel(GraphView g, const int2* __restrict__ affected, int naffected, const int* __restrict__ item_first, const int* __restrict__ cand,
              int rmc, int list_cap, int cand_cap, int sel_cap, unsigned* work_counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = build_scratch_bytes(g.pitch, list_cap, 64, cand_cap, sel_cap);
    BuildScratch b = carve_build(smem + warp * per_warp, g.pitch, list_cap, 64, cand_cap, sel_cap, lane);
    WarpScratch& s = b.s;
    for (;;) {
        unsigned wi = 0;
        if (lane == 0) wi = atomicAdd(work_counter, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= (unsigned)naffected) break;
        const int a = affected[wi].x, layer = affected[wi].y;
        const int maxE = layer == 0 ? 2 * g.max_edges : g.max_edges;
        load_vector(g, a, s.q, lane);
        double an = 0.0;
        if (METRIC == COSINE) an = __ldg(g.norm_sqrt + a);
        int stride;
        int* row = row_ptr_mut(g, a, layer, stride);

        BeamState st;
        st.n = 0; st.ef = list_cap; st.far_ord = 0xFFFFFFFFu; st.cursor = 0;
        unsigned ndist = 0;
        // candidates [0, m) of s.cand_id -> distances to a -> sorted list (list_insert drops repeated ids)
        auto take = [&](int m) {
            if (m <= 0) return;
            batch_dist<METRIC>(g, s, an, m, lane);
            for (int j = 0; j < m; ++j) list_insert(s.list, st, make_key(s.cand_dist[j], s.cand_id[j]), lane);
        };
        // 1. the neighbours that stay (GraphConnector.cs:104-120); removed ones are remembered in b.sel
        int nkept = 0, nrem = 0;
        for (int rb = 0; rb < stride; rb += 32) {
            const int slot = rb + lane;
            const int nb = slot < stride ? row[slot] : -1;
            const int ri = nb >= 0 ? __ldg(g.rm_index + nb) : -1;
            const bool kept = nb >= 0 && ri == -1, gone = nb >= 0 && ri >= 0;
            const unsigned mk = __ballot_sync(FULL, kept), mg = __ballot_sync(FULL, gone);
            if (kept) s.cand_id[nkept + __popc(mk & ((1u << lane) - 1u))] = nb;
            if (gone) b.sel[nrem + __popc(mg & ((1u << lane) - 1u))] = ri;
            nkept += __popc(mk);
            nrem += __popc(mg);
        }
        __syncwarp();
        take(nkept);
        // 2. the search results of every removed neighbour of this row (:122-129), 32 at a time
        for (int t = 0; t < nrem; ++t) {
            const int* cl = cand + (size_t)(__ldg(item_first + b.sel[t]) + layer) * rmc;
            for (int c0 = 0; c0 < rmc; c0 += 32) {
                const int j = c0 + lane;
                const int c = j < rmc ? __ldg(cl + j) : -1;
                const bool ok = c >= 0 && c != a;
                const unsigned mk = __ballot_sync(FULL, ok);
                if (ok) s.cand_id[__popc(mk & ((1u << lane) - 1u))] = c;
                __syncwarp();
                take(__popc(mk));
                if (__ballot_sync(FULL, c < 0 && j < rmc)) break;  // lists are -1 padded at the end
            }
        }
        // 3. Heuristic.cs:11-46 and the new row (:131-164: the outcome of the remove/add pass is exactly newOut)
        int nsel;
        if (st.n < maxE) {
            nsel = st.n;
            for (int j = lane; j < nsel; j += 32) b.sel[j] = key_id(s.list[j]);
            __syncwarp();
        } else {
            nsel = select_neighbors<METRIC>(g, b, s.list, st.n, maxE, lane, ndist);
        }
        for (int j = lane; j < stride; j += 32) row[j] = j < nsel ? b.sel[j] : -1;
        __syncwarp();
    }
}

// rows of the removed nodes become empty (the reference leaves them stale; nothing reaches them either way)
__global__ void clear_rows_kernel(GraphView g, const int* __restrict__ item_node, const int* __restrict__ item_layer, int nitems) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nitems) return;
    int stride;
    int* row = row_ptr_mut(g, item_node[w], item_layer[w], stride);
    for (int j = lane; j < stride; j += 32) row[j] = -1;
}

static int pow2_ge(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <typename Kern>
static int persistent_grid(Kern kern, size_t smem_per_warp, int work_items, int& warps) {
    return persistent_launch(kern, smem_per_warp, 0, work_items, warps);
}

void launch_set_rm_index(int* d_rm_index, const int* d_ids, int n, bool use_index, int value, cudaStream_t stream) {
    if (n <= 0) return;
    set_rm_index_kernel<<<(n + 255) / 256, 256, 0, stream>>>(d_rm_index, d_ids, n, use_index, value);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_remove_search(int metric, const GraphView& g, const int* d_item_node, const int* d_item_layer, int nitems, int rmc, int* d_cand,
                          unsigned* d_work_counter, cudaStream_t stream) {
    if (nitems <= 0) return;
    const int list_cap = rmc + 1;
    const int vis = std::min(2048, std::max(256, pow2_ge(2 * list_cap)));
    const int cand_cap = ((std::max(g.stride0, g.strideU) + 31) / 32) * 32;
    const size_t per_warp = warp_scratch_bytes(g.pitch, list_cap, vis, cand_cap);
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_work_counter, 0, sizeof(unsigned), stream));
#define HNSW_LAUNCH_RS(M)                                                                                                    \
    {                                                                                                                        \
        int w = 0;                                                                                                           \
        const int grid = persistent_grid(remove_search_kernel<M>, per_warp, nitems, w);                                     \
        remove_search_kernel<M><<<grid, 32 * w, per_warp * w, stream>>>(g, d_item_node, d_item_layer, nitems, rmc, list_cap, \
                                                                        vis, cand_cap, d_cand, d_work_counter);             \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_LAUNCH_RS(SQ_EUCLID) break;
        case EUCLID: HNSW_LAUNCH_RS(EUCLID) break;
        case COSINE: HNSW_LAUNCH_RS(COSINE) break;
        default: HNSW_LAUNCH_RS(UCOSINE) break;
    }
#undef HNSW_LAUNCH_RS
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_find_affected(const GraphView& g, const int* d_levels, int length, int2* d_affected, unsigned* d_n_affected, cudaStream_t stream) {
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_n_affected, 0, sizeof(unsigned), stream));
    if (length <= 0) return;
    int dev = 0, sms = 0;
    HNSW_CUDA_CHECK(cudaGetDevice(&dev));
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = std::max(1, std::min((length + 7) / 8, sms * 8));
    find_affected_kernel<<<blocks, 256, 0, stream>>>(g, d_levels, length, d_affected, d_n_affected);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_relink(int metric, const GraphView& g, const int2* d_affected, int naffected, const int* d_item_first, const int* d_cand, int rmc,
                   unsigned* d_work_counter, cudaStream_t stream) {
    if (naffected <= 0) return;
    const int maxE0 = 2 * g.max_edges;
    const int list_cap = maxE0 + 2 * rmc;  // one removed neighbour never overflows it; more of them keep the closest
    const int cand_cap = ((std::max(std::max(g.stride0, g.strideU), 32) + 31) / 32) * 32;
    const int sel_cap = std::max(maxE0, std::max(g.stride0, g.strideU));
    const size_t per_warp = build_scratch_bytes(g.pitch, list_cap, 64, cand_cap, sel_cap);
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_work_counter, 0, sizeof(unsigned), stream));
#define HNSW_LAUNCH_RL(M)                                                                                                       \
    {                                                                                                                           \
        int w = 0;                                                                                                              \
        const int grid = persistent_grid(relink_kernel<M>, per_warp, naffected, w);                                            \
        relink_kernel<M><<<grid, 32 * w, per_warp * w, stream>>>(g, d_affected, naffected, d_item_first, d_cand, rmc, list_cap, \
                                                                 cand_cap, sel_cap, d_work_counter);                            \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_LAUNCH_RL(SQ_EUCLID) break;
        case EUCLID: HNSW_LAUNCH_RL(EUCLID) break;
        case COSINE: HNSW_LAUNCH_RL(COSINE) break;
        default: HNSW_LAUNCH_RL(UCOSINE) break;
    }
#undef HNSW_LAUNCH_RL
    HNSW_CUDA_CHECK(cudaGetLastError());
}

void launch_clear_rows(const GraphView& g, const int* d_item_node, const int* d_item_layer, int nitems, cudaStream_t stream) {
    if (nitems <= 0) return;
    const int blocks = (nitems * 32 + 255) / 256;
    clear_rows_kernel<<<blocks, 256, 0, stream>>>(g, d_item_node, d_item_layer, nitems);
    HNSW_CUDA_CHECK(cudaGetLastError());
}

}  // namespace hnsw
