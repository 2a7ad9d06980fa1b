// build_scratch.cuh -- device helpers shared by the construction (build_kernels.cu) and removal (remove_kernels.cu)
// kernels: the per-warp scratch with a second vector slot, and RelativeNeighborPruning (Heuristic.cs:11-46).
#pragma once
#include "beam.cuh"

namespace hnsw {

struct BuildScratch {
    WarpScratch s;
    float* cvec;  // [2][pitch] vector buffers of the heuristic (the vector under test, double-buffered)
    int* sel;     // [sel_cap] selected neighbour ids
};



// The heuristic's two vector slots are only live between two beam searches, when the visited table is dead (every beam
// search starts by clearing it): when the table is large enough the slots live inside it (1 KB per warp at dim 128, which
// is what lets 18 instead of 16 warps of the select kernel share an SM).
__host__ __device__ inline bool cvec_in_vis(int pitch, int vis_size) { return (size_t)vis_size * 4 >= (size_t)((pitch + 3) & ~3) * 8; }
__host__ __device__ inline size_t build_scratch_bytes(int pitch, int cap, int vis_size, int cand_cap, int sel_cap) {
    const size_t cvec = cvec_in_vis(pitch, vis_size) ? 0 : (size_t)((pitch + 3) & ~3) * 8;
    return warp_scratch_bytes(pitch, cap, vis_size, cand_cap) + ((cvec + (size_t)sel_cap * 4 + 127) & ~(size_t)127);
}
__device__ __forceinline__ BuildScratch carve_build(unsigned char* base, int pitch, int cap, int vis_size, int cand_cap, int sel_cap, int lane) {
    BuildScratch b;
    b.s = carve_scratch(base, pitch, cap, vis_size, cand_cap, lane);
    base += warp_scratch_bytes(pitch, cap, vis_size, cand_cap);
    if (cvec_in_vis(pitch, vis_size)) {
        b.cvec = reinterpret_cast<float*>(b.s.vis);  // 16-byte aligned (carve_scratch)
    } else {
        b.cvec = reinterpret_cast<float*>(base);
        base += (size_t)((pitch + 3) & ~3) * 8;
    }
    b.sel = reinterpret_cast<int*>(base);
    return b;
}

__device__ __forceinline__ void load_vector(const GraphView& g, int id, float* dst, int lane) {
    const float4* src = reinterpret_cast<const float4*>(g.vectors + (size_t)id * g.pitch);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = lane; i < (g.pitch >> 2); i += 32) d4[i] = __ldg(src + i);
    __syncwarp();
}

// Heuristic.cs:11-46 on the sorted list[0..n): returns the number selected (ids in b.sel, ascending
// by distance).  Caller handles the n < maxE early-return case.
// Long rows (more than one staging segment): candidate by candidate, the selected rows re-staged for each.
template <int METRIC>
__device__ __forceinline__ int select_neighbors_long(const GraphView& g, BuildScratch& b, const u64* list, int n, int maxE, int lane,
                                                     unsigned& ndist) {
    int nsel = 0;
    for (int i = 0; i < n && nsel < maxE; ++i) {
        const u64 e = list[i];
        const int c = key_id(e);
        const float cd = key_dist(e);
        bool ok = true;
        if (nsel > 0) {
            load_vector(g, c, b.cvec, lane);
            double cn = 0.0;
            if (METRIC == COSINE) cn = __ldg(g.norm_sqrt + c);
            // 16 selected neighbours per pass, closest first; stop at the first pass that rejects the candidate
            // (most rejections come from the closest neighbours, so the later passes are rarely needed)
            for (int s0 = 0; s0 < nsel && ok; s0 += kStageRows) {
                const int m = min(kStageRows, nsel - s0);
                batch_dist_to<METRIC>(g, b.s.stg, b.cvec, cn, b.sel + s0, b.s.cand_dist, m, lane);
                ndist += m;
                const bool rej = lane < m && b.s.cand_dist[lane] < cd;  // strict, Heuristic.cs:34
                ok = !__any_sync(FULL, rej);
                __syncwarp();
            }
        }
        if (ok) {
            if (lane == 0) b.sel[nsel] = c;
            ++nsel;
            __syncwarp();
        }
    }
    return nsel;
}

// distance between one staged row (this lane pair's) and a vector in shared memory; both lanes of the pair get it
template <int METRIC, int SPEC = 0>
__device__ __forceinline__ float staged_row_dist(const GraphView& g, const float* row, const float* q, double an, double qn, bool active, int h) {
    constexpr bool SQ = metric_is_sq(METRIC);
    const int nfull = (SPEC & kSpecDim128) ? 16 : g.dim >> 3, ntail = (SPEC & kSpecDim128) ? 0 : g.dim - (nfull << 3);
    PairAcc acc = {0.f, 0.f, 0.f, 0.f};
    if (active) pair_accumulate<SQ>(acc, row, q, nfull, h);
    float r = pair_hsum<SQ>(acc);
    if (active && ntail > 0) r = pair_tail<SQ>(r, row + (nfull << 3), q + (nfull << 3), ntail);
    return metric_finish<METRIC>(r, an, qn);
}

// Rows of one staging segment: the sorted candidates are taken in WINDOWS of 16.  A window's rows are staged once
// (one lane pair per candidate); every neighbour selected before the window is streamed through a double-buffered
// vector slot (cp.async, the next one in flight under the current pass) and tested against all 16 rows in one pass;
// then the window is resolved in order, a newly selected candidate being tested against the later ones straight out
// of the staging buffer.  A candidate is dropped at the first selected neighbour that is strictly closer to it than
// the base vector is (Heuristic.cs:31-35), so it is evaluated against exactly the neighbours the reference tries.
template <int METRIC, int SPEC = 0>
__device__ __forceinline__ int select_neighbors_windowed(const GraphView& g, BuildScratch& b, const u64* list, int n, int maxE, int lane,
                                                         unsigned& ndist) {
    WarpScratch& s = b.s;
    const int h = lane & 1, grp = lane >> 1;
    const int pitch4 = g.pitch >> 2;
    int nsel = 0;
    for (int w0 = 0; w0 < n && nsel < maxE; w0 += kStageRows) {
        const int nw = min(kStageRows, n - w0);
        if (lane < nw) s.cand_id[lane] = key_id(list[w0 + lane]);
        __syncwarp();
        stage_issue(g, s.stg, s.cand_id, 0, nw, 0, g.pitch, lane);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int nsel0 = nsel;
        if (nsel0 > 0 && lane < pitch4) cp_async_16(b.cvec + lane * 4, g.vectors + (size_t)b.sel[0] * g.pitch + lane * 4);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const bool have = grp < nw;
        const u64 mykey = have ? list[w0 + grp] : 0ull;
        const int myid = key_id(mykey);
        const float cd = key_dist(mykey);
        double an = 0.0;
        if (METRIC == COSINE && have) an = __ldg(g.norm_sqrt + myid);
        const float* row = s.stg.buf + grp * kSegPitch;
        bool rej = !have;
        // neighbours selected in earlier windows
        for (int j = 0; j < nsel0; ++j) {
            if (__all_sync(FULL, rej)) break;
            float* qv = b.cvec + (j & 1) * g.pitch;
            if (j + 1 < nsel0) {
                if (lane < pitch4) cp_async_16(b.cvec + ((j + 1) & 1) * g.pitch + lane * 4, g.vectors + (size_t)b.sel[j + 1] * g.pitch + lane * 4);
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();
            double qn = 0.0;
            if (METRIC == COSINE) qn = __ldg(g.norm_sqrt + b.sel[j]);
            const bool act = !rej;
            ndist += __popc(__ballot_sync(FULL, act && h == 0));
            const float d = staged_row_dist<METRIC, SPEC>(g, row, qv, an, qn, act, h);
            if (act && d < cd) rej = true;  // strict, Heuristic.cs:34
            __syncwarp();                   // the slot of j is free for the copy issued two iterations later
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        // the window itself, in order
        for (int gi = 0; gi < nw; ++gi) {
            if (__shfl_sync(FULL, (int)rej, gi * 2)) continue;
            if (lane == 0) b.sel[nsel] = s.cand_id[gi];
            ++nsel;
            if (nsel == maxE) break;
            const bool act = !rej && grp > gi;
            const unsigned later = __ballot_sync(FULL, act && h == 0);
            if (!later) continue;
            ndist += __popc(later);
            double qn = 0.0;
            if (METRIC == COSINE) qn = __ldg(g.norm_sqrt + s.cand_id[gi]);
            const float d = staged_row_dist<METRIC, SPEC>(g, row, s.stg.buf + gi * kSegPitch, an, qn, act, h);
            if (act && d < cd) rej = true;
        }
        __syncwarp();  // staged rows, cand_id and sel are consistent before the next window overwrites them
    }
    return nsel;
}

// SPEC & kSpecShortRows: the caller knows at compile time that rows fit one staging segment (the long-row code is dropped)
template <int METRIC, int SPEC = 0>
__device__ __forceinline__ int select_neighbors(const GraphView& g, BuildScratch& b, const u64* list, int n, int maxE, int lane, unsigned& ndist) {
#ifndef HNSW_STAGE_BULK
    if ((SPEC & kSpecShortRows) || g.pitch <= kSegFloats) return select_neighbors_windowed<METRIC, SPEC>(g, b, list, n, maxE, lane, ndist);
#endif
    if (SPEC & kSpecShortRows) return 0;  // (unreachable)
    return select_neighbors_long<METRIC>(g, b, list, n, maxE, lane, ndist);
}

__device__ __forceinline__ int* row_ptr_mut(const GraphView& g, int id, int layer, int& stride) {
    if (layer == 0) {
        stride = g.stride0;
        return g.adj0 + (size_t)id * g.stride0;
    }
    stride = g.strideU;
    return g.adjU + ((size_t)__ldg(g.upper_off + id) + (size_t)(layer - 1)) * g.strideU;
}

}  // namespace hnsw
