// beam.cuh -- warp-level HNSW traversal primitives (one warp = one query / one insert).
//
// Replaces, for TVector=float[] / TDistance=float:
//   GraphNavigator.FindEntryAtLayer     GraphNavigator.cs:51-82   (greedy descent)
//   GraphNavigator.SearchLayer(Query)   GraphNavigator.cs:123-256 (ef beam search)
//   BinaryHeap / DistanceComparer       BinaryHeap.cs, DistanceComparer.cs
//   VisitedList / VisitedListPool       VisitedListPool.cs
//
// The reference keeps a max-heap `top` (cap ef) and a min-heap `cand`.  Everything
// ever pushed to `cand` was pushed to `top` at the same time, and whatever `top`
// evicts is >= far for the rest of the search, so it can never be expanded before
// the break at GraphNavigator.cs:218.  Hence the two heaps are equivalent (up to
// exact fp32 ties) to ONE sorted list of the best <= ef entries carrying an
// "expanded" bit: expand the closest unexpanded entry until none is left.
// The list lives in shared memory as 64-bit keys (ordered distance | id | flag).
//
// Visited set: the reference stamps a short[capacity] per search.  Here a small
// 2-way bucketed table of recently seen ids in shared memory filters most repeats.
// It may forget (never invents): a forgotten node is re-evaluated, then either
// rejected again (its distance is >= far, which only shrinks) or recognised as a
// duplicate of a list entry (same distance bits, same id) -- results are identical
// to an exact visited set, only the number of distance evaluations can grow.
#pragma once
#include "common.cuh"
#include "metric.cuh"
#include "stage.cuh"

namespace hnsw {

struct WarpScratch {
    Stager stg;        // staged neighbour rows + mbarrier (stage.cuh)
    float* q;          // [pitch] query / inserted vector (16 B aligned)
    u64* list;         // [cap] sorted candidate keys
    uint32_t* vis;     // [vis_size] recently visited ids: 2-way buckets of 32-bit ids + 1, or one word = one bucket of two
                       // 16-bit tags (vis16_test_and_set) when the index is small enough
    int* cand_id;      // [cand_cap] fresh neighbour ids of the current expansion
    float* cand_dist;  // [cand_cap]
    int vis_size;      // power of two >= 64
    int vis_shift;     // 32 - log2(vis_size / 2)
    int cap;           // list capacity (>= ef)
    uint32_t* vbm;     // optional EXACT visited bitmap of this warp in global memory (one bit per node; long-row searches), or null
    int vbm_words;     // its size in 32-bit words (multiple of 4)
    // (keep this struct small: two more words and ptxas compiles the 96-register construction kernel in a mode that runs 2.4x slower)
};

constexpr int kPrefetchRows = 1;  // flags bit 0: L2-prefetch the adjacency rows of accepted candidates
constexpr int kGuessRow = 2;      // flags bit 1: load the likely next candidate's row while the last pass is inserted

struct WarpCounters {
    unsigned ndist = 0, nexp = 0, row_words = 0;
};

__host__ __device__ inline size_t warp_scratch_bytes(int pitch, int cap, int vis_size, int cand_cap) {
    size_t b = kStageBytes;
    b += (size_t)((pitch + 3) & ~3) * 4;
    b += ((size_t)cap * 8 + 15) & ~(size_t)15;  // the visited table behind the list stays 16-byte aligned
    b += (size_t)vis_size * 4;
    b += (size_t)cand_cap * 8;
    return (b + 127) & ~(size_t)127;
}
// Called once per warp at kernel start (initialises the warp's mbarrier).
__device__ __forceinline__ WarpScratch carve_scratch(unsigned char* base, int pitch, int cap, int vis_size, int cand_cap, int lane) {
    WarpScratch s;
    stager_init(s.stg, base, lane);
    base += kStageBytes;
    s.q = reinterpret_cast<float*>(base);
    base += (size_t)((pitch + 3) & ~3) * 4;
    s.list = reinterpret_cast<u64*>(base);
    base += ((size_t)cap * 8 + 15) & ~(size_t)15;
    s.vis = reinterpret_cast<uint32_t*>(base);
    base += (size_t)vis_size * 4;
    s.cand_id = reinterpret_cast<int*>(base);
    base += (size_t)cand_cap * 4;
    s.cand_dist = reinterpret_cast<float*>(base);
    s.vis_size = vis_size;
    s.vis_shift = 32 - (31 - __clz(vis_size >> 1));
    s.cap = cap;
    s.vbm = nullptr;
    s.vbm_words = 0;
    return s;
}

__device__ __forceinline__ const int* row_ptr(const GraphView& g, int id, int layer, int& stride) {
    if (layer == 0) {
        stride = g.stride0;
        return g.adj0 + (size_t)id * g.stride0;
    }
    stride = g.strideU;
    return g.adjU + ((size_t)__ldg(g.upper_off + id) + (size_t)(layer - 1)) * g.strideU;
}

// ----------------------------------------------------------------------------
// distances of ids[0..m) to the vector q (shared memory) -> out[0..m): 16 rows per pass, staged in
// shared memory by asynchronous full-row copies (stage.cuh), one lane pair per row (metric.cuh)
// ----------------------------------------------------------------------------
// start copying floats [f0, f0+nfl) of rows ids[base .. base+nrows) into the staging buffer (asynchronous)
__device__ __forceinline__ void stage_issue(const GraphView& g, Stager& stg, const int* ids, int base, int nrows, int f0, int nfl, int lane) {
#ifdef HNSW_STAGE_BULK
    const int copy_id = lane < nrows ? ids[base + lane] : 0;  // lane r stages row r
    if (lane == 0) stage_expect(stg, (uint32_t)(nrows * nfl * 4));
    __syncwarp();
    if (lane < nrows) stage_copy(stg, stg.buf + lane * kSegPitch, g.vectors + (size_t)copy_id * g.pitch + f0, (uint32_t)(nfl * 4));
#else
    // Lane r fetches the id of row r once; every row then costs one SHFL, one 32x32->64 multiply-add for the address
    // and one LDGSTS (16 B per lane = the whole row per instruction).
    const int myid = lane < nrows ? ids[base + lane] : 0;
    const char* src0 = reinterpret_cast<const char*>(g.vectors) + (size_t)(f0 + lane * 4) * 4;
    const uint32_t dst0 = stg.buf_s + lane * 16;
    const unsigned pitch_bytes = (unsigned)g.pitch * 4u;
    const bool act = lane * 4 < nfl;  // lanes beyond the row length copy nothing
    const int limit = act ? nrows : 0;
    // fully unrolled: a rolled x4 loop has a third of the code but measured 2.5 % slower (gpurun_out/sweep13.log)
#pragma unroll
    for (int r = 0; r < kStageRows; ++r) {
        const unsigned id = (unsigned)__shfl_sync(FULL, myid, r);
        if (r < limit) cp_async_16s(dst0 + r * (kSegPitch * 4), src0 + (size_t)id * pitch_bytes);
    }
#endif
}
// the same for a subset of the rows: bit 2r of `keep` set = row r is still wanted (lane-pair ballot form)
__device__ __forceinline__ void stage_issue_kept(const GraphView& g, Stager& stg, const int* ids, int base, int nrows, int f0, int nfl, int lane,
                                                 unsigned keep) {
    const int myid = lane < nrows ? ids[base + lane] : 0;
    const char* src0 = reinterpret_cast<const char*>(g.vectors) + (size_t)(f0 + lane * 4) * 4;
    const uint32_t dst0 = stg.buf_s + lane * 16;
    const unsigned pitch_bytes = (unsigned)g.pitch * 4u;
    if (lane * 4 >= nfl) keep = 0u;  // lanes beyond the row length copy nothing
#pragma unroll
    for (int r = 0; r < kStageRows; ++r) {
        const unsigned id = (unsigned)__shfl_sync(FULL, myid, r);
        if ((keep >> (2 * r)) & 1u) cp_async_16s(dst0 + r * (kSegPitch * 4), src0 + (size_t)id * pitch_bytes);
    }
}
// wait until the copies started by stage_issue have landed and are visible to the whole warp
__device__ __forceinline__ void stage_complete(Stager& stg) {
#ifdef HNSW_STAGE_BULK
    stage_wait(stg);
#else
    cp_async_wait_all();
    __syncwarp();
#endif
}

// SPEC (compile-time knowledge that lets the hot kernels drop dead code; 0 = generic):
//   bit 0: rows fit one staging segment (pitch <= kSegFloats);  bit 1: the list never exceeds 64 entries (ef <= 64)
//   bit 2: the visited table holds 16-bit tags (vis16_test_and_set): every id of the index is below vis_size * 65536
//   bit 3: the list never exceeds 1024 entries (ef <= 1024): the binary-search insertion path is dead code
//   bit 5: the search kernel attaches an exact per-warp visited bitmap in global memory behind the shared-memory table
//          (long rows: WarpScratch::vbm, batch_dist_probe; the CTA-per-query kernel has its own use of the bit)
//   bit 6: the list is wide.cuh's non-shifting WideList (sorted main array + 32 pending keys in registers), layer 0 only
//   bit 7: vectors of exactly 128 floats (dim == pitch == 128): constant trip counts in the distance loops (construction)
constexpr int kSpecShortRows = 1, kSpecSmallList = 2, kSpecVis16 = 4, kSpecMidList = 8, kSpecGvis = 32, kSpecWideList = 64, kSpecDim128 = 128;

// rows of at most kSegFloats floats (one segment), already staged: distances of ids[base .. base+nrows) -> out
template <int METRIC, int SPEC = 0>
__device__ __forceinline__ void compute_staged(const GraphView& g, Stager& stg, const float* q, double q_norm_sqrt, const int* ids, float* out,
                                               int base, int nrows, int lane) {
    constexpr bool SQ = metric_is_sq(METRIC);
    const int h = lane & 1, grp = lane >> 1;
    // kSpecDim128: the chunk loop has a compile-time trip count (the construction kernel is issue-bound: its loop control counts)
    const int nfull = (SPEC & kSpecDim128) ? 16 : g.dim >> 3, ntail = (SPEC & kSpecDim128) ? 0 : g.dim - (nfull << 3);
    const bool active = grp < nrows;
    double an = 0.0;
    if (METRIC == COSINE && active) an = __ldg(g.norm_sqrt + ids[base + grp]);
    PairAcc acc = {0.f, 0.f, 0.f, 0.f};
    const float* row = stg.buf + grp * kSegPitch;
    if (active) pair_accumulate<SQ>(acc, row, q, nfull, h);
    float r = pair_hsum<SQ>(acc);
    if (active && ntail > 0) r = pair_tail<SQ>(r, row + (nfull << 3), q + (nfull << 3), ntail);
    if (active && h == 0) out[base + grp] = metric_finish<METRIC>(r, an, q_norm_sqrt);
    __syncwarp();  // staged rows are free again, distances are visible
}


template <int METRIC, int SPEC = 0>
__device__ __forceinline__ void batch_dist_to(const GraphView& g, Stager& stg, const float* q, double q_norm_sqrt, const int* ids, float* out,
                                              int m, int lane) {
    constexpr bool SQ = metric_is_sq(METRIC);
    const int h = lane & 1, grp = lane >> 1;
    const int nfull = g.dim >> 3;                 // complete 8-element chunks
    const int ntail = g.dim - (nfull << 3);       // scalar remainder
    for (int base = 0; base < m; base += kStageRows) {
        const int nrows = min(kStageRows, m - base);
        if ((SPEC & kSpecShortRows) || g.pitch <= kSegFloats) {
            stage_issue(g, stg, ids, base, nrows, 0, g.pitch, lane);
            stage_complete(stg);
            compute_staged<METRIC, SPEC>(g, stg, q, q_norm_sqrt, ids, out, base, nrows, lane);
            continue;
        }
        if (SPEC & kSpecShortRows) continue;  // (unreachable; lets the compiler drop the long-row code)
        // long rows: segments of kSegFloats floats, lane accumulators carried across segments
        const bool active = grp < nrows;
        double an = 0.0;
        if (METRIC == COSINE && active) an = __ldg(g.norm_sqrt + ids[base + grp]);
        PairAcc acc = {0.f, 0.f, 0.f, 0.f};
        float r = 0.f;
        for (int f0 = 0; f0 < g.pitch; f0 += kSegFloats) {
            const int nfl = min(kSegFloats, g.pitch - f0);  // floats of this segment, multiple of 4
            stage_issue(g, stg, ids, base, nrows, f0, nfl, lane);
            stage_complete(stg);
            const float* row = stg.buf + grp * kSegPitch;
            const int nch = min(nfull - (f0 >> 3), kSegFloats >> 3);
            if (active && nch > 0) pair_accumulate<SQ>(acc, row, q + f0, nch, h);
            if (f0 + kSegFloats >= g.pitch) {  // last segment: it also holds the scalar remainder
                r = pair_hsum<SQ>(acc);
                if (active && ntail > 0) r = pair_tail<SQ>(r, row + ((nfull << 3) - f0), q + (nfull << 3), ntail);
            }
            __syncwarp();  // all lanes are done with the staged rows before the next copies overwrite them
        }
        if (active && h == 0) out[base + grp] = metric_finish<METRIC>(r, an, q_norm_sqrt);
    }
    __syncwarp();
}
// Long rows with the warp's exact visited bitmap (kSpecGvis searches): the filter in front of this call is the forgetful
// shared-memory table, so some of ids[0..m) have been evaluated before -- on 1M x 960 that is 9 % of all evaluations, each
// 3.8 KB of DRAM traffic.  An exact test costs a round trip to L2 / DRAM, which this kernel cannot afford once per expansion,
// so it travels under work that is needed anyway: the bitmap words of ALL candidates are requested up front; the first pass
// (16 rows) starts at once and drops the candidates found visited after the first segment of their rows (1/8 of a 960-float
// row read instead of all of it: ids[j] = -1, no distance); by then every word has arrived, and the remaining candidates are
// compacted (stable: stored order is kept) before their passes are formed, so those passes are full and fewer.
// Fresh candidates are marked.  m is updated; returns the number of dropped candidates.
template <int METRIC, int SPEC>
__device__ __forceinline__ int batch_dist_probe(const GraphView& g, Stager& stg, const float* q, double q_norm_sqrt, int* ids, float* out, int& m,
                                                uint32_t* vbm, int lane) {
    constexpr bool SQ = metric_is_sq(METRIC);
    const int h = lane & 1, grp = lane >> 1;
    const int nfull = g.dim >> 3;
    const int ntail = g.dim - (nfull << 3);
    int dropped = 0;
    // requests: pass 0's rows by lane pair, the rest one candidate per lane (two chunks in registers; more are tested later)
    const int nrows = min(kStageRows, m);
    const bool mine = grp < nrows;
    const int id = mine ? ids[grp] : 0;
    uint32_t word = 0u;
    if (mine && h == 0) word = __ldcg(vbm + (id >> 5));
    const int j0 = kStageRows + lane, j1 = kStageRows + 32 + lane;
    const int id0 = j0 < m ? ids[j0] : -1, id1 = j1 < m ? ids[j1] : -1;
    uint32_t w0 = 0u, w1 = 0u;
    if (id0 >= 0) w0 = __ldcg(vbm + (id0 >> 5));
    if (id1 >= 0) w1 = __ldcg(vbm + (id1 >> 5));
    {   // pass 0
        bool active = mine;
        unsigned keep = __ballot_sync(FULL, mine && h == 0);
        double an = 0.0;
        if (METRIC == COSINE && active) an = __ldg(g.norm_sqrt + id);
        PairAcc acc = {0.f, 0.f, 0.f, 0.f};
        float r = 0.f;
        for (int f0 = 0; f0 < g.pitch; f0 += kSegFloats) {
            const int nfl = min(kSegFloats, g.pitch - f0);
            stage_issue_kept(g, stg, ids, 0, nrows, f0, nfl, lane, keep);
            stage_complete(stg);
            const float* row = stg.buf + grp * kSegPitch;
            const int nch = min(nfull - (f0 >> 3), kSegFloats >> 3);
            if (active && nch > 0) pair_accumulate<SQ>(acc, row, q + f0, nch, h);
            if (f0 == 0) {
                const bool seen = mine && h == 0 && ((word >> (id & 31)) & 1u);
                if (mine && h == 0 && !seen) atomicOr(vbm + (id >> 5), 1u << (id & 31));  // result unused: a reduction
                keep = __ballot_sync(FULL, mine && h == 0 && !seen);
                active = mine && ((keep >> (lane & ~1)) & 1u);
                dropped += nrows - __popc(keep);
            }
            if (f0 + kSegFloats >= g.pitch) {
                r = pair_hsum<SQ>(acc);
                if (active && ntail > 0) r = pair_tail<SQ>(r, row + ((nfull << 3) - f0), q + (nfull << 3), ntail);
            }
            __syncwarp();
        }
        if (mine && h == 0) {
            if (active) out[grp] = metric_finish<METRIC>(r, an, q_norm_sqrt);
            else ids[grp] = -1;
        }
    }
    if (m <= kStageRows) {
        __syncwarp();
        return dropped;
    }
    // compact the rest: stable, fresh ones marked
    int mnew = kStageRows;
    {
        const bool f0 = id0 >= 0 && !((w0 >> (id0 & 31)) & 1u);
        const bool f1 = id1 >= 0 && !((w1 >> (id1 & 31)) & 1u);
        if (f0) atomicOr(vbm + (id0 >> 5), 1u << (id0 & 31));
        if (f1) atomicOr(vbm + (id1 >> 5), 1u << (id1 & 31));
        const unsigned m0 = __ballot_sync(FULL, f0), m1 = __ballot_sync(FULL, f1);
        __syncwarp();  // every lane has read its ids before they are overwritten
        if (f0) ids[mnew + __popc(m0 & ((1u << lane) - 1u))] = id0;
        mnew += __popc(m0);
        if (f1) ids[mnew + __popc(m1 & ((1u << lane) - 1u))] = id1;
        mnew += __popc(m1);
        __syncwarp();
    }
    for (int jb = kStageRows + 64; jb < m; jb += 32) {  // rows longer than 80 neighbours: tested chunk by chunk (one round trip each)
        const int j = jb + lane;
        const int idj = j < m ? ids[j] : -1;
        bool fr = false;
        if (idj >= 0) fr = !((atomicOr(vbm + (idj >> 5), 1u << (idj & 31)) >> (idj & 31)) & 1u);
        const unsigned mk = __ballot_sync(FULL, fr);
        __syncwarp();
        if (fr) ids[mnew + __popc(mk & ((1u << lane) - 1u))] = idj;
        mnew += __popc(mk);
        __syncwarp();
    }
    dropped += m - mnew;
    m = mnew;
    if (m > kStageRows) batch_dist_to<METRIC, SPEC>(g, stg, q, q_norm_sqrt, ids + kStageRows, out + kStageRows, m - kStageRows, lane);
    __syncwarp();
    return dropped;
}
template <int METRIC, int SPEC = 0>
__device__ __forceinline__ void batch_dist(const GraphView& g, WarpScratch& s, double q_norm_sqrt, int m, int lane) {
    batch_dist_to<METRIC, SPEC>(g, s.stg, s.q, q_norm_sqrt, s.cand_id, s.cand_dist, m, lane);
}

template <int METRIC, int SPEC = 0>
__device__ __forceinline__ float single_dist(const GraphView& g, WarpScratch& s, double q_norm_sqrt, int id, int lane) {
    __syncwarp();
    if (lane == 0) s.cand_id[0] = id;
    __syncwarp();
    batch_dist<METRIC, SPEC>(g, s, q_norm_sqrt, 1, lane);
    const float d = s.cand_dist[0];
    __syncwarp();
    return d;
}

// ----------------------------------------------------------------------------
// sorted list in shared memory
// ----------------------------------------------------------------------------
struct BeamState {
    int n;             // entries in list
    int ef;            // capacity in use
    uint32_t far_ord;  // ordered distance of list[ef-1] when full, else 0xFFFFFFFF
    int cursor;        // every entry below cursor is expanded
};

__device__ __forceinline__ int list_lower_bound(const u64* list, int n, u64 key, int lane) {
    if (n <= 256) {
        int pos = 0;
        for (int base = 0; base < n; base += 32) {
            const int idx = base + lane;
            const bool lt = idx < n && list[idx] < key;
            const unsigned mk = __ballot_sync(FULL, lt);
            pos += __popc(mk);
            if (mk != FULL) break;  // sorted: nothing smaller beyond this chunk
        }
        return pos;
    }
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (list[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Insert `key` (flag bit clear) keeping the list sorted, dropping the last entry when
// full.  Returns false when the id is already present (duplicate) or falls off the end.
template <int SPEC = 0>
__device__ __forceinline__ bool list_insert(u64* list, BeamState& st, u64 key, int lane) {
    const int n = st.n;
    if ((SPEC & kSpecSmallList) || n <= 64) {
        // whole list in two registers per lane: one pass finds the position, detects duplicates and shifts
        const u64 v0 = lane < n ? list[lane] : ~0ull;
        const u64 v1 = lane + 32 < n ? list[lane + 32] : ~0ull;
        const int pos = __popc(__ballot_sync(FULL, v0 < key)) + __popc(__ballot_sync(FULL, v1 < key));
        const u64 at = __shfl_sync(FULL, pos < 32 ? v0 : v1, pos & 31);  // list[pos], ~0 beyond the end
        if ((at >> 1) == (key >> 1)) return false;                       // same distance bits + id
        if (pos >= st.ef) return false;
        const int last = (n == st.ef) ? n - 1 : n;                       // entries [pos, last) move up by one
        if (lane >= pos && lane < last) list[lane + 1] = v0;
        if (lane + 32 >= pos && lane + 32 < last) list[lane + 33] = v1;
        if (lane == 0) list[pos] = key;
        __syncwarp();
        if (n < st.ef) st.n = n + 1;
        st.far_ord = (st.n == st.ef) ? key_ord(list[st.ef - 1]) : 0xFFFFFFFFu;
        if (pos < st.cursor) st.cursor = pos;
        return true;
    }
    if (SPEC & kSpecSmallList) return false;  // (unreachable)
    if (n <= 1024) {
        // Two-level search (construction: ef = MaxCandidates ~ 100-200): lane c looks at the LAST entry of chunk c
        // (32 entries per chunk), the ballot names the first chunk that is not entirely below the key, a second
        // ballot over that chunk gives the position.  Only the chunks from there to the end are shifted.
        const int nch = (n + 31) >> 5;
        u64 bnd = ~0ull;
        if (lane < nch) bnd = list[min(lane * 32 + 31, n - 1)];
        const int ci = __popc(__ballot_sync(FULL, bnd < key));  // chunks entirely below the key
        int pos = n;
        u64 v = ~0ull, at = ~0ull;
        if (ci < nch) {
            const int idx = ci * 32 + lane;
            if (idx < n) v = list[idx];
            const int off = __popc(__ballot_sync(FULL, v < key));  // < 32: the chunk's last entry is >= key
            pos = ci * 32 + off;
            at = __shfl_sync(FULL, v, off);  // list[pos]
        }
        if ((at >> 1) == (key >> 1)) return false;  // same distance bits + id
        if (pos >= st.ef) return false;
        const int last = (n == st.ef) ? n - 1 : n;  // entries [pos, last) move up by one
        __syncwarp();                               // every lane has read its entries before any is overwritten
        if (last > pos) {
            for (int c = (last - 1) >> 5; c > ci; --c) {  // top chunk first: its successor has already moved
                const int idx = c * 32 + lane;
                u64 w = 0;
                if (idx < last) w = list[idx];
                __syncwarp();
                if (idx < last) list[idx + 1] = w;
            }
            const int idx = ci * 32 + lane;
            if (idx >= pos && idx < last) list[idx + 1] = v;
        }
        if (lane == 0) list[pos] = key;
        __syncwarp();
        if (n < st.ef) st.n = n + 1;
        st.far_ord = (st.n == st.ef) ? key_ord(list[st.ef - 1]) : 0xFFFFFFFFu;
        if (pos < st.cursor) st.cursor = pos;
        return true;
    }
    if (SPEC & kSpecMidList) return false;  // (unreachable)
    const int pos = list_lower_bound(list, n, key, lane);
    if (pos < n && (list[pos] >> 1) == (key >> 1)) return false;  // same distance bits + id
    if (pos >= st.ef) return false;
    const int last = (n == st.ef) ? n - 1 : n;  // entries [pos, last) move up by one
    __syncwarp();
    for (int hi = last; hi > pos; hi -= 32) {
        const int idx = hi - 1 - lane;
        u64 v = 0;
        if (idx >= pos) v = list[idx];
        __syncwarp();
        if (idx >= pos) list[idx + 1] = v;
        __syncwarp();
    }
    if (lane == 0) list[pos] = key;
    __syncwarp();
    if (n < st.ef) st.n = n + 1;
    st.far_ord = (st.n == st.ef) ? key_ord(list[st.ef - 1]) : 0xFFFFFFFFu;
    if (pos < st.cursor) st.cursor = pos;
    return true;
}

__device__ __forceinline__ void vis_clear(const WarpScratch& s, int lane) {
    for (int i = lane; i < s.vis_size; i += 32) s.vis[i] = 0u;
    __syncwarp();
}
// true if `id` was not in the table (and records it)
__device__ __forceinline__ bool vis_test_and_set(const WarpScratch& s, int id) {
    const uint32_t tag = (uint32_t)id + 1u;
    const uint32_t b = (tag * 0x9E3779B1u) >> s.vis_shift;  // multiplicative hash, top bits
    const uint32_t v0 = s.vis[2 * b], v1 = s.vis[2 * b + 1];
    if (v0 == tag || v1 == tag) return false;
    s.vis[2 * b + 1] = v0;
    s.vis[2 * b] = tag;
    return true;
}
// Same table with 16-bit entries -- twice the entries in the same shared memory -- usable while every id is below
// vis_size * 65536: x -> x * A mod 2^(16+b) (A odd) is a bijection of the ids, so (bucket = top b bits, tag = low 16
// bits) still identifies an id exactly.  One 32-bit word is one 2-way bucket: a lookup is one LDS, an update one STS.
// Tag 0 marks an empty way; the 1/65536 of the ids whose tag is 0 are simply never recorded (re-evaluated, harmless).
__device__ __forceinline__ bool vis16_test_and_set(const WarpScratch& s, int id) {
    const uint32_t z = (uint32_t)id * 0x9E3779B1u;
    const uint32_t tag = z & 0xFFFFu;
    const uint32_t b = (z >> 16) & (uint32_t)(s.vis_size - 1);
    const uint32_t w = s.vis[b];
    if (tag == 0u) return true;
    if ((w & 0xFFFFu) == tag || (w >> 16) == tag) return false;
    s.vis[b] = (w << 16) | tag;
    return true;
}

// ----------------------------------------------------------------------------
// Greedy descent on one upper layer (GraphNavigator.cs:51-82): scan the whole
// row of the current node in stored order, move to the strict minimum (first
// occurrence on ties), repeat until a scan brings no improvement.
// ----------------------------------------------------------------------------
template <int METRIC, int SPEC = 0>
__device__ __forceinline__ void greedy_layer(const GraphView& g, WarpScratch& s, double qn, int layer, int& cur, float& cur_d,
                                             int lane, WarpCounters& wc) {
    bool changed = true;
    while (changed) {
        changed = false;
        int stride;
        const int* row = row_ptr(g, cur, layer, stride);
        int m = 0;
        for (int rb = 0; rb < stride; rb += 32) {
            const int slot = rb + lane;
            const int nb = slot < stride ? row[slot] : -1;
            const unsigned mk = __ballot_sync(FULL, nb >= 0);
            if (nb >= 0) s.cand_id[m + __popc(mk & ((1u << lane) - 1u))] = nb;
            m += __popc(mk);
        }
        __syncwarp();
        wc.nexp++;
        wc.row_words += m + 1;
        if (m == 0) break;
        wc.ndist += m;
        batch_dist<METRIC, SPEC>(g, s, qn, m, lane);
        u64 best = ~0ull;
        for (int j = lane; j < m; j += 32) {
            const u64 k = ((u64)f2ord(s.cand_dist[j]) << 32) | (u64)(uint32_t)j;
            best = k < best ? k : best;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const u64 other = __shfl_xor_sync(FULL, best, o);
            best = other < best ? other : best;
        }
        const int bj = (int)(uint32_t)best;
        const float bd = s.cand_dist[bj];
        const int bid = s.cand_id[bj];
        __syncwarp();
        if (bd < cur_d) {
            cur_d = bd;
            cur = bid;
            changed = true;
        }
    }
}

// Insert candidates [lo, hi) of the current expansion in stored order; a candidate is accepted iff
// |top| < ef or d < far (strict), re-evaluated after every insertion (GraphNavigator.cs:233-246).
template <int SPEC = 0>
__device__ __forceinline__ void insert_candidates(const GraphView& g, WarpScratch& s, BeamState& st, int lo, int hi, int layer, int lane, int flags) {
    for (int jb = lo; jb < hi; jb += 32) {
        const int j = jb + lane;
        const bool have = j < hi && (!(SPEC & kSpecGvis) || s.cand_id[j] >= 0);  // kSpecGvis: batch_dist_probe drops visited candidates
        const u64 mykey = have ? make_key(s.cand_dist[j], s.cand_id[j]) : ~0ull;
        bool pending = have;
        for (;;) {
            const bool want = pending && (st.n < st.ef || key_ord(mykey) < st.far_ord);
            const unsigned mk = __ballot_sync(FULL, want);
            if (!mk) break;
            const int src = __ffs(mk) - 1;
            const u64 kn = __shfl_sync(FULL, mykey, src);
            if (lane == src) pending = false;
            const bool inserted = list_insert<SPEC>(s.list, st, kn, lane);
            // optional: pull the adjacency row of an accepted candidate into L2 (measured: no gain at C2)
            if (inserted && lane == src && layer == 0 && (flags & kPrefetchRows)) {
                const char* rp = reinterpret_cast<const char*>(g.adj0 + (size_t)key_id(kn) * g.stride0);
                for (int off = 0; off < g.stride0 * 4; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + off));
            }
        }
    }
}

// ----------------------------------------------------------------------------
// Beam search of width ef on one layer (GraphNavigator.cs:194-256).  On return
// s.list[0..st.n) holds the result sorted ascending by (distance, id).
// ----------------------------------------------------------------------------
template <int METRIC, int SPEC = 0>
__device__ __forceinline__ void beam_search(const GraphView& g, WarpScratch& s, double qn, int layer, int entry, float entry_d,
                                            int ef, BeamState& st, int lane, WarpCounters& wc, int flags = kPrefetchRows) {
    st.n = 0;
    st.ef = ef;
    st.far_ord = 0xFFFFFFFFu;
    st.cursor = 0;
    vis_clear(s, lane);
    if ((SPEC & kSpecGvis) && s.vbm != nullptr) {
        uint4* t = reinterpret_cast<uint4*>(s.vbm);
        for (int i = lane; i < (s.vbm_words >> 2); i += 32) __stcg(t + i, make_uint4(0u, 0u, 0u, 0u));
        __syncwarp();
    }
    // compile-time choice: two inlined variants under a run-time branch cost this loop a factor of two on sm_100a
    constexpr bool vis16 = (SPEC & kSpecVis16) != 0;
    list_insert<SPEC>(s.list, st, make_key(entry_d, entry), lane);
    if (lane == 0) {
        if (vis16) vis16_test_and_set(s, entry); else vis_test_and_set(s, entry);
        if ((SPEC & kSpecGvis) && s.vbm != nullptr) atomicOr(s.vbm + (entry >> 5), 1u << (entry & 31));
    }
    __syncwarp();
    int guess_id = -1, guess_nb = -1;  // adjacency word [lane] of the candidate expected to be expanded next

    for (;;) {
        // closest unexpanded entry
        int sel = -1;
        for (int base = st.cursor & ~31; base < st.n; base += 32) {
            const int idx = base + lane;
            const bool un = idx < st.n && !(s.list[idx] & 1ull);
            const unsigned mk = __ballot_sync(FULL, un);
            if (mk) { sel = base + __ffs(mk) - 1; break; }
        }
        if (sel < 0) break;
        const u64 e = s.list[sel];
        const int c = key_id(e);
        __syncwarp();
        if (lane == 0) s.list[sel] = e | 1ull;
        __syncwarp();
        st.cursor = sel + 1;

        // row of c: keep the neighbours not seen recently, in stored order
        int stride;
        const int* row = row_ptr(g, c, layer, stride);
        int m = 0, deg = 0;
        for (int rb = 0; rb < stride; rb += 32) {
            const int slot = rb + lane;
            int nb;
            if (guess_id == c) nb = guess_nb;  // row already fetched under the previous insertion (stride <= 32)
            else nb = slot < stride ? row[slot] : -1;
            deg += __popc(__ballot_sync(FULL, nb >= 0));
            bool fresh = false;
            if (nb >= 0) fresh = vis16 ? vis16_test_and_set(s, nb) : vis_test_and_set(s, nb);
            const unsigned mk = __ballot_sync(FULL, fresh);
            if (fresh) s.cand_id[m + __popc(mk & ((1u << lane) - 1u))] = nb;
            m += __popc(mk);
        }
        __syncwarp();
        wc.nexp++;
        wc.row_words += deg + 1;
        if (m == 0) continue;
        wc.ndist += m;

        if ((SPEC & kSpecShortRows) || g.pitch <= kSegFloats) {
            // Software pipeline over passes of 16 rows: the copies of pass p+1 are in flight while the
            // candidates of pass p are inserted (stored order is kept: pass p is inserted before p+1).
            stage_issue(g, s.stg, s.cand_id, 0, min(kStageRows, m), 0, g.pitch, lane);
            for (int base = 0; base < m; base += kStageRows) {
                const int nrows = min(kStageRows, m - base);
                stage_complete(s.stg);
                compute_staged<METRIC, SPEC>(g, s.stg, s.q, qn, s.cand_id, s.cand_dist, base, nrows, lane);
                const int nbase = base + kStageRows;
                if (nbase < m) {
                    stage_issue(g, s.stg, s.cand_id, nbase, min(kStageRows, m - nbase), 0, g.pitch, lane);
                } else if ((flags & kGuessRow) && layer == 0 && g.stride0 <= 32) {
                    // Last pass: unless one of its candidates lands in front of it, the next expansion is the
                    // closest entry still unexpanded now -- start loading its row under the insertion work.
                    guess_id = -1;
                    for (int b2 = st.cursor & ~31; b2 < st.n; b2 += 32) {
                        const int idx = b2 + lane;
                        const bool un = idx < st.n && !(s.list[idx] & 1ull);
                        const unsigned mk = __ballot_sync(FULL, un);
                        if (mk) { guess_id = key_id(s.list[b2 + __ffs(mk) - 1]); break; }
                    }
                    if (guess_id >= 0) guess_nb = lane < g.stride0 ? __ldg(g.adj0 + (size_t)guess_id * g.stride0 + lane) : -1;
                }
                insert_candidates<SPEC>(g, s, st, base, base + nrows, layer, lane, flags);
            }
        } else if (!(SPEC & kSpecShortRows)) {
            if ((SPEC & kSpecGvis) && s.vbm != nullptr) wc.ndist -= batch_dist_probe<METRIC, SPEC>(g, s.stg, s.q, qn, s.cand_id, s.cand_dist, m, s.vbm, lane);
            else batch_dist<METRIC, SPEC>(g, s, qn, m, lane);
            insert_candidates<SPEC>(g, s, st, 0, m, layer, lane, flags);
        }
    }
}

}  // namespace hnsw
