// wide.cuh -- the candidate list of a WIDE beam (ef from a few hundred to ~16k): a sorted main array that never shifts
// per insertion + a 32-entry pending buffer held in one register per lane.
//
// Same object as beam.cuh's single sorted list (BinaryHeap.cs / GraphNavigator.cs:194-256: the reference's result
// max-heap and candidate min-heap folded into ONE ordered set of the best <= ef entries with an "expanded" bit), same
// decisions in the same order -- only the bookkeeping differs:
//   * an accepted candidate goes into the PENDING buffer: lane i holds the i-th smallest pending key, an insertion is one
//     ballot + one shuffle, no shared-memory traffic;
//   * when 32 entries are pending they are merged into the main array in ONE pass (every main entry above the smallest
//     pending key moves up by the number of pending keys below it, top chunk first, in place): the O(ef / 32) chunk moves
//     that beam.cuh's list_insert pays per accepted candidate are paid once per 32 of them;
//   * the entry the reference's max-heap would evict is the larger of main[n-1] and the largest pending key: evicting from
//     the main array is n--, evicting from the buffer clears one lane;
//   * far (GraphNavigator.cs:233-246, re-read after every insertion) = the larger of the two when n + pending == ef;
//   * the closest unexpanded entry is the smaller of (first unexpanded main entry at or after the cursor, first
//     unexpanded pending lane).
// All functions are executed by ONE warp (all 32 lanes); every member except `pend` is warp-uniform.
#pragma once
#include "beam.cuh"

namespace hnsw {

constexpr int kPendCap = 32;

struct WideList {
    u64* main;          // shared memory [ef]: entries [0, n) sorted ascending by (ordered distance | id | expanded flag)
    int n;              // live entries of main
    int pc;             // pending entries (lanes [0, pc) of `pend`)
    int ef;             // capacity of the set: n + pc <= ef
    int cursor;         // every main entry below cursor is expanded
    u64 pend;           // lane i: i-th smallest pending key, ~0 for lanes >= pc
    u64 main_last;      // main[n - 1] (0 when n == 0)
    uint32_t far_ord;   // ordered distance of the largest entry when the set is full, else 0xFFFFFFFF
};

__device__ __forceinline__ void wl_init(WideList& L, u64* main, int ef) {
    L.main = main;
    L.n = 0;
    L.pc = 0;
    L.ef = ef;
    L.cursor = 0;
    L.pend = ~0ull;
    L.main_last = 0ull;
    L.far_ord = 0xFFFFFFFFu;
}

__device__ __forceinline__ void wl_update_far(WideList& L) {
    if (L.n + L.pc < L.ef) {
        L.far_ord = 0xFFFFFFFFu;
        return;
    }
    u64 mx = L.n > 0 ? L.main_last : 0ull;
    if (L.pc > 0) {
        const u64 pm = __shfl_sync(FULL, L.pend, L.pc - 1);
        mx = pm > mx ? pm : mx;
    }
    L.far_ord = key_ord(mx);
}

// per-lane binary search: number of entries of a[0..n) below key (lanes search independently)
__device__ __forceinline__ int lane_lower_bound(const u64* a, int n, u64 key) {
    int lo = 0, len = n;
    while (len > 0) {
        const int half = len >> 1;
        if (a[lo + half] < key) {
            lo += half + 1;
            len -= half + 1;
        } else {
            len = half;
        }
    }
    return lo;
}

// Merge the pending keys into the main array (one pass, in place).
// Seen from the OUTPUT array (n + pc entries): pending key j lands at P_j = (main entries below it) + j -- distinct,
// increasing positions; the main entries fill the remaining positions in order.  So for an output chunk of 32 positions
// one OR-reduction of the lanes' P_j bits gives the chunk's "pending slot" mask, and main entry -> position is a popcount:
// no per-entry search.  Chunks are handled from the top down (a chunk's sources lie at or below its own positions, i.e.
// never in a chunk that was already overwritten), four chunks per iteration so that their shared-memory latencies overlap.
__device__ __forceinline__ void wl_merge(WideList& L, int lane) {
    const int n = L.n, pc = L.pc;
    if (pc == 0) return;
    int P = 0x7FFFFFFF;  // output position of my pending key
    if (lane < pc) P = lane_lower_bound(L.main, n, L.pend) + lane;
    const int P0 = __shfl_sync(FULL, P, 0);
    const int total = n + pc;
    const int c_lo = P0 >> 5;  // output chunks below this one are untouched
    __syncwarp();
    auto chunk_src = [&](int c, unsigned& occ, int& below) {
        const int off = P - (c << 5);
        occ = __reduce_or_sync(FULL, (off >= 0 && off < 32) ? (1u << off) : 0u);
        below = __popc(__ballot_sync(FULL, off < 0));  // pending slots in lower chunks (lanes >= pc have P = INT_MAX)
    };
    // four chunks per iteration: their shared-memory round trips overlap (a merge of a 16k-entry list is 512 chunks)
    for (int c = (total - 1) >> 5; c >= c_lo; c -= 4) {
        u64 x[4];
        bool mv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int cc = c - u;
            mv[u] = false;
            x[u] = 0ull;
            if (cc >= c_lo) {  // warp-uniform
                unsigned occ;
                int bel;
                chunk_src(cc, occ, bel);
                const int o = (cc << 5) + lane;
                mv[u] = o < total && o >= P0 && !((occ >> lane) & 1u);
                if (mv[u]) x[u] = L.main[o - bel - __popc(occ & ((1u << lane) - 1u))];
            }
        }
        __syncwarp();  // all four chunks' sources are in registers before any position is overwritten
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (mv[u]) L.main[((c - u) << 5) + lane] = x[u];
        __syncwarp();
    }
    if (lane < pc) L.main[P] = L.pend;
    __syncwarp();
    L.n = total;
    L.pc = 0;
    L.pend = ~0ull;
    L.main_last = L.main[total - 1];
    const int lb0 = P0;  // = main entries below the smallest pending key
    if (lb0 < L.cursor) L.cursor = lb0;
}

// Insert a key that passed the acceptance test (|set| < ef or d < far) and is not a duplicate.  Returns true when the
// pending buffer was merged into the main array afterwards.
__device__ __forceinline__ bool wl_insert(WideList& L, u64 kn, int lane) {
    const int pos = __popc(__ballot_sync(FULL, L.pend < kn));  // lanes >= pc hold ~0
    const u64 up = __shfl_up_sync(FULL, L.pend, 1);
    if (lane > pos) L.pend = up;
    else if (lane == pos) L.pend = kn;
    L.pc++;
    if (L.n + L.pc > L.ef) {  // evict the largest entry of the set (what the reference's max-heap pops)
        const u64 pm = __shfl_sync(FULL, L.pend, L.pc - 1);
        if (L.n == 0 || pm > L.main_last) {
            if (lane == L.pc - 1) L.pend = ~0ull;
            L.pc--;
        } else {
            L.n--;
            L.main_last = L.n > 0 ? L.main[L.n - 1] : 0ull;
            if (L.cursor > L.n) L.cursor = L.n;
        }
    }
    wl_update_far(L);
    if (L.pc == kPendCap) {
        wl_merge(L, lane);
        return true;
    }
    return false;
}

// Closest unexpanded entry WITHOUT marking it (used to guess the next expansion); ~0 when there is none.
// `second` (optional): the unexpanded main-array entry after it in the same chunk, or ~0 -- a cheap hint for prefetching.
__device__ __forceinline__ u64 wl_peek(const WideList& L, int lane, u64* second = nullptr) {
    u64 mkey = ~0ull, skey = ~0ull;
    for (int base = L.cursor & ~31; base < L.n; base += 32) {
        const int idx = base + lane;
        const u64 v = idx < L.n ? L.main[idx] : 1ull;
        unsigned mk = __ballot_sync(FULL, !(v & 1ull));
        if (mk) {
            mkey = __shfl_sync(FULL, v, __ffs(mk) - 1);
            mk &= mk - 1;
            if (second != nullptr && mk) skey = __shfl_sync(FULL, v, __ffs(mk) - 1);
            break;
        }
    }
    const unsigned pk = __ballot_sync(FULL, lane < L.pc && !(L.pend & 1ull));
    if (pk) {
        const u64 pkey = __shfl_sync(FULL, L.pend, __ffs(pk) - 1);
        if (pkey < mkey) {
            skey = mkey;
            mkey = pkey;
        } else if (pkey < skey) {
            skey = pkey;
        }
    }
    if (second != nullptr) *second = skey;
    return mkey;
}

// Closest unexpanded entry: marks it expanded and returns its key; false when every entry is expanded.
__device__ __forceinline__ bool wl_select(WideList& L, int lane, u64& out) {
    int sel = -1;
    u64 mkey = ~0ull;
    for (int base = L.cursor & ~31; base < L.n; base += 32) {
        const int idx = base + lane;
        const u64 v = idx < L.n ? L.main[idx] : 1ull;
        const unsigned mk = __ballot_sync(FULL, !(v & 1ull));
        if (mk) {
            const int l = __ffs(mk) - 1;
            sel = base + l;
            mkey = __shfl_sync(FULL, v, l);
            break;
        }
    }
    L.cursor = sel >= 0 ? sel : L.n;
    const unsigned pk = __ballot_sync(FULL, lane < L.pc && !(L.pend & 1ull));
    int pl = -1;
    u64 pkey = ~0ull;
    if (pk) {
        pl = __ffs(pk) - 1;
        pkey = __shfl_sync(FULL, L.pend, pl);
    }
    if (sel < 0 && pl < 0) return false;
    if (pkey < mkey) {
        if (lane == pl) L.pend |= 1ull;
        out = pkey;
    } else {
        __syncwarp();
        if (lane == 0) L.main[sel] = mkey | 1ull;
        __syncwarp();
        if (sel == L.n - 1) L.main_last = mkey | 1ull;
        L.cursor = sel + 1;
        out = mkey;
    }
    return true;
}

// Is (distance bits, id) of `key` already in the main array?  Warp-cooperative 32-ary search: every step the lanes look at
// 32 evenly spaced pivots and one ballot names the bucket (x 32 narrower), the last step compares a whole bucket at once:
// 2 shared-memory round trips up to 1024 entries, 3 up to 32k -- against ~log2(n) dependent ones for a binary search.
__device__ __forceinline__ bool wl_contains(const WideList& L, u64 key, int lane) {
    int lo = 0, len = L.n;
    while (len > 32) {
        const int step = (len + 31) >> 5;
        const int idx = min(lo + (lane + 1) * step, lo + len) - 1;
        const int below = __popc(__ballot_sync(FULL, L.main[idx] < key));  // pivots below the key (an expanded twin is key | 1)
        const int nlo = lo + below * step;
        len = min(step, lo + len - nlo);
        lo = nlo;
    }
    u64 v = ~0ull;
    if (lane < len) v = L.main[lo + lane];
    return __any_sync(FULL, (v >> 1) == (key >> 1));
}

// Candidates [lo, hi) of one expansion (ids / distances in shared memory), in stored order; a candidate is accepted iff
// |set| < ef or d < far (strict), far re-read after every insertion (GraphNavigator.cs:233-246).  A node the forgetful
// visited table let through a second time passes that test only if it is still in the set (far only shrinks), so every
// candidate that passes is looked up by (distance bits, id) in the pending buffer (one vote) and in the main array
// (wl_contains) before it is inserted.
// EXACT_VISITED: the caller's visited set never forgets (global bitmap), so no node is evaluated twice and the lookups are
// skipped.
template <bool EXACT_VISITED = false>
__device__ __forceinline__ void wl_insert_candidates(WideList& L, const int* cand_id, const float* cand_dist, int lo, int hi, int lane) {
    for (int jb = lo; jb < hi; jb += 32) {
        const int j = jb + lane;
        const bool have = j < hi;
        const u64 mykey = have ? make_key(cand_dist[j], cand_id[j]) : ~0ull;
        bool pending = have;
        for (;;) {
            const bool want = pending && (L.n + L.pc < L.ef || key_ord(mykey) < L.far_ord);
            const unsigned mk = __ballot_sync(FULL, want);
            if (!mk) break;
            const int src = __ffs(mk) - 1;
            const u64 kn = __shfl_sync(FULL, mykey, src);
            if (lane == src) pending = false;
            if (!EXACT_VISITED) {
                if (__any_sync(FULL, lane < L.pc && (L.pend >> 1) == (kn >> 1))) continue;  // already pending
                if (wl_contains(L, kn, lane)) continue;                                      // already in the array
            }
            wl_insert(L, kn, lane);
        }
    }
}

// ----------------------------------------------------------------------------
// beam.cuh's beam_search (GraphNavigator.cs:194-256, layer 0) on the non-shifting list, for the one-warp-per-query
// kernels: same software pipeline (the copies of pass p+1 fly while pass p is inserted, the likely next row is fetched
// under the last pass's insertions), same visited table, same order of decisions.  On return s.list[0..L.n) holds the
// result sorted ascending.
// ----------------------------------------------------------------------------
template <int METRIC, int SPEC>
__device__ __forceinline__ void beam_search_wl(const GraphView& g, WarpScratch& s, double qn, int entry, float entry_d, int ef, WideList& L,
                                               int lane, WarpCounters& wc, int flags) {
    constexpr bool vis16 = (SPEC & kSpecVis16) != 0;
    wl_init(L, s.list, ef);
    vis_clear(s, lane);
    wl_insert(L, make_key(entry_d, entry), lane);
    if (lane == 0) {
        if (vis16) vis16_test_and_set(s, entry); else vis_test_and_set(s, entry);
    }
    __syncwarp();
    int guess_id = -1, guess_nb = -1;
    for (;;) {
        u64 e;
        if (!wl_select(L, lane, e)) break;
        const int c = key_id(e);
        const int* row = g.adj0 + (size_t)c * g.stride0;
        int m = 0, deg = 0;
        for (int rb = 0; rb < g.stride0; rb += 32) {
            const int slot = rb + lane;
            int nb;
            if (guess_id == c && rb == 0) nb = guess_nb;
            else nb = slot < g.stride0 ? __ldg(row + slot) : -1;
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
            stage_issue(g, s.stg, s.cand_id, 0, min(kStageRows, m), 0, g.pitch, lane);
            for (int base = 0; base < m; base += kStageRows) {
                const int nrows = min(kStageRows, m - base);
                stage_complete(s.stg);
                compute_staged<METRIC, SPEC>(g, s.stg, s.q, qn, s.cand_id, s.cand_dist, base, nrows, lane);
                const int nbase = base + kStageRows;
                if (nbase < m) {
                    stage_issue(g, s.stg, s.cand_id, nbase, min(kStageRows, m - nbase), 0, g.pitch, lane);
                } else if (flags & kGuessRow) {
                    const u64 pk = wl_peek(L, lane);
                    guess_id = pk != ~0ull ? key_id(pk) : -1;
                    if (guess_id >= 0) guess_nb = lane < g.stride0 ? __ldg(g.adj0 + (size_t)guess_id * g.stride0 + lane) : -1;
                }
                wl_insert_candidates(L, s.cand_id, s.cand_dist, base, base + nrows, lane);
            }
        } else if (!(SPEC & kSpecShortRows)) {
            batch_dist<METRIC, SPEC>(g, s, qn, m, lane);
            if (flags & kGuessRow) {
                const u64 pk = wl_peek(L, lane);
                guess_id = pk != ~0ull ? key_id(pk) : -1;
                if (guess_id >= 0) guess_nb = lane < g.stride0 ? __ldg(g.adj0 + (size_t)guess_id * g.stride0 + lane) : -1;
            }
            wl_insert_candidates(L, s.cand_id, s.cand_dist, 0, m, lane);
        }
    }
    wl_merge(L, lane);
}

}  // namespace hnsw
