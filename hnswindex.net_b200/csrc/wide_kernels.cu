// wide_kernels.cu -- K2w: KnnQuery with a WIDE beam, one CTA per query.
//
// Replaces the same reference code as search_kernels.cu (HNSWIndex.cs:107-137, GraphNavigator.cs:39-82,194-256) for
// ef = max(MinNN, k) from a few hundred up to ~24k, and for calls of a handful of queries.  The reference's heaps grow
// without bound (BinaryHeap.cs:109-113), so any MinNN costs it the same per step; the one-warp-per-query kernel does not
// scale that way: its sorted list shifts O(ef / 32) chunks per accepted candidate and one warp's list + visited table leave
// 2-3 warps per SM at ef = 4096.  Here one CTA holds ONE query's state (query vector, candidate list, visited set) once:
//   * warp 0 (the LEADER) owns the control flow of GraphNavigator.SearchLayerQuery: pick the closest unexpanded entry,
//     filter its adjacency row through the visited set, insert the evaluated neighbours in stored order (wide.cuh: sorted
//     main array + 32-entry pending buffer in registers, merged in one pass);
//   * the other warps (HELPERS, one by default) evaluate distances -- and they evaluate the NEXT expansion ahead of time,
//     while the leader still inserts the current one (see the comment above the kernel);
//   * the visited set is the shared-memory table of beam.cuh or, from ef 1024, an exact bitmap in global memory.
// The order of every decision (strict d < far re-read after each insertion, stored order, first occurrence on ties) is the
// one of beam.cuh, so results and the nexp / row_words counters are identical to the one-warp kernel and to the oracle.
#include <math_constants.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

#include "kernels.hpp"
#include "wide.cuh"

namespace hnsw {

constexpr int kCmdEval = 0, kCmdSpec = 1, kCmdDone = 2;

// Leader-side profile of the CTA-per-query kernel, summed over all queries since the last reset (SM cycles of the leader
// warp; two atomics per counter per QUERY, so it is always on): hnsw_get_wide_profile().
//   0 queries  1 expansions  2 speculations issued  3 speculations used  4 cyc descent  5 cyc select + filter
//   6 cyc waiting for a speculation  7 cyc waiting for an on-demand evaluation  8 cyc insert  9 cyc beam loop total
// (the compiler may move the clock reads across the barriers: the split between 5..8 is indicative, 9 is exact)
__device__ unsigned long long g_wide_prof[16];

struct WideCtrl {
    unsigned qi;       // query handled by this CTA
    int cmd;           // what the helper warps do next (kCmd*)
    int m;             // kCmdEval: number of ids in the authoritative list
    int node;          // kCmdSpec: the node whose expansion is evaluated ahead of time
    int buf;           // kCmdSpec: which of the two speculative buffers to fill
    int spec_node[2];  // filled by the helpers: node and list length of each speculative buffer
    int spec_m[2];
    int spec_deg[2];   // ... and the degree of that node (for the row_words counter)
    int pad;
};

// shared memory of one CTA: [helpers' stagers][q][list][visited table][auth ids | auth dists | 2 x (spec ids | spec dists)][ctrl]
__host__ __device__ inline size_t wide_smem_bytes(int warps, int pitch, int ef, int vis_size, int cand_cap) {
    size_t b = (size_t)(warps - 1) * kStageBytes;
    b += (size_t)((pitch + 3) & ~3) * 4;
    b += (size_t)((ef + 1) & ~1) * 8;
    b += (size_t)vis_size * 4;
    b += (size_t)cand_cap * 4 * 6;
    b += sizeof(WideCtrl);
    return (b + 127) & ~(size_t)127;
}

// read-only look at the visited table (the helpers' speculative filter): would test_and_set call `id` fresh?
template <bool VIS16>
__device__ __forceinline__ bool vis_would_be_fresh(const uint32_t* vis, int vis_size, int vis_shift, int id) {
    if (VIS16) {
        const uint32_t z = (uint32_t)id * 0x9E3779B1u, tag = z & 0xFFFFu;
        const uint32_t w = vis[(z >> 16) & (uint32_t)(vis_size - 1)];
        return tag == 0u || !((w & 0xFFFFu) == tag || (w >> 16) == tag);
    }
    const uint32_t tag = (uint32_t)id + 1u;
    const uint32_t b = (tag * 0x9E3779B1u) >> vis_shift;
    return !(vis[2 * b] == tag || vis[2 * b + 1] == tag);
}

// Leader / helper pipeline (2..4 warps per CTA, one CTA per query):
//   warp 0, the LEADER, takes every decision of GraphNavigator.SearchLayerQuery in the reference's order: closest unexpanded
//   entry, its row filtered through the visited table (the only place the table is written), candidates inserted in stored
//   order with far re-read after each one (wide.cuh);
//   warps 1.., the HELPERS, evaluate distances.  While the leader inserts the candidates of expansion i, the helpers already
//   evaluate expansion i+1 SPECULATIVELY: the leader names the entry that will be expanded next unless one of the pending
//   candidates lands in front of it; the helpers read that node's row, filter it through the table read-only (nobody writes
//   the table during an insertion phase, so they see exactly what the leader will see), stage the fresh rows and reduce the
//   distances into one of two speculative buffers.  When the leader then selects that very node, the helpers' id list IS
//   what its own filter would produce (same row, same visited set: only the leader writes the set, and only in this step),
//   so it takes the list and the distances over and merely records the new ids in the set; otherwise (wrong guess) it
//   filters the row itself and the helpers evaluate that list on demand.  A wrong guess leaves no trace -- nothing of it
//   reaches the list or the visited set --, so results and counters are those of the one-warp kernels; only the latency of
//   row + filter + gather + reduction moves off the leader's critical path.
// SPEC: kSpecShortRows (rows fit one staging segment), kSpecVis16 (16-bit visited tags), kSpecGvis (the visited set is an
// EXACT bitmap of this CTA in global memory -- bitmap + (size_t)blockIdx.x * bitmap_words, one bit per node, L2-resident:
// one atomicOr per neighbour instead of a shared-memory table, no re-evaluations at all, and the shared memory of the table
// goes to more resident queries); SPECULATE: evaluate ahead
// Registers: with only a thread bound ptxas settles on 64 registers and 12-68 bytes of spills outside the hot loops.  Stated
// budgets were measured (1M x 128, full batches): 80 registers (no spills) is no faster at any width; 96 or 128 cost resident
// CTAs where shared memory is not the limit (ef 384-1024: -10 ... -18 %) and change nothing from ef 2048 on or for one query.
template <int METRIC, int SPEC, bool SPECULATE>
__global__ void __launch_bounds__(kThreads)
knn_wide_kernel(GraphView g, const float* __restrict__ queries, int nq, int k, int ef, int vis_size, int cand_cap, int* __restrict__ out_ids,
                float* __restrict__ out_dists, unsigned* work_counter, QueryStats* stats, uint32_t* bitmap, int bitmap_words) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5, H = W - 1;
    constexpr bool vis16 = (SPEC & kSpecVis16) != 0;
    constexpr bool gvis = (SPEC & kSpecGvis) != 0;
    uint32_t* bm = gvis ? bitmap + (size_t)blockIdx.x * bitmap_words : nullptr;

    unsigned char* p = smem;
    Stager stg;  // helpers: their own staging buffer; the leader borrows helper 1's for the descent (the helpers are idle then)
    stager_init(stg, p + (size_t)(warp == 0 ? 0 : warp - 1) * kStageBytes, lane);
    p += (size_t)H * kStageBytes;
    float* q = reinterpret_cast<float*>(p);
    p += (size_t)((g.pitch + 3) & ~3) * 4;
    u64* list = reinterpret_cast<u64*>(p);
    p += (size_t)((ef + 1) & ~1) * 8;
    uint32_t* vis = reinterpret_cast<uint32_t*>(p);
    p += (size_t)vis_size * 4;
    int* cand_id = reinterpret_cast<int*>(p);
    float* cand_dist = reinterpret_cast<float*>(p + (size_t)cand_cap * 4);
    int* spec_id0 = reinterpret_cast<int*>(p + (size_t)cand_cap * 8);
    p += (size_t)cand_cap * 24;
    WideCtrl* ctrl = reinterpret_cast<WideCtrl*>(p);
    const int vis_shift = 32 - (31 - __clz(vis_size >> 1));

    // the leader's view of the shared scratch for the warp-level primitives of beam.cuh
    WarpScratch s;
    s.stg = stg;
    s.q = q;
    s.list = list;
    s.vis = vis;
    s.cand_id = cand_id;
    s.cand_dist = cand_dist;
    s.vis_size = vis_size;
    s.vis_shift = vis_shift;
    s.cap = ef;
    s.vbm = nullptr;
    s.vbm_words = 0;

    for (;;) {
        __syncthreads();  // the previous query is finished by every warp
        if (threadIdx.x == 0) ctrl->qi = atomicAdd(work_counter, 1u);
        __syncthreads();
        const unsigned qi = ctrl->qi;
        if (qi >= (unsigned)nq) break;
        const float* qsrc = queries + (size_t)qi * g.dim;
        for (int i = threadIdx.x; i < g.pitch; i += blockDim.x) q[i] = i < g.dim ? __ldg(qsrc + i) : 0.f;
        for (int i = threadIdx.x; i < vis_size; i += blockDim.x) vis[i] = 0u;
        if (gvis) {
            uint4* b4 = reinterpret_cast<uint4*>(bm);
            for (int i = threadIdx.x; i < (bitmap_words >> 2); i += blockDim.x) __stcg(b4 + i, make_uint4(0u, 0u, 0u, 0u));
            __threadfence();
        }
        __syncthreads();
        double qn = 0.0;
        if (METRIC == COSINE) qn = pair_norm_sqrt(q, g.dim, lane & 1);  // every warp computes the same value

        if (warp != 0) {
            // ---------------- helpers: serve commands until the query is done
            for (;;) {
                __syncthreads();  // (X) a command is published
                const int cmd = ctrl->cmd;
                if (cmd == kCmdDone) break;
                if (cmd == kCmdEval) {
                    const int m = ctrl->m;
                    for (int base = kStageRows * (warp - 1); base < m; base += kStageRows * H)
                        batch_dist_to<METRIC, SPEC>(g, stg, q, qn, cand_id + base, cand_dist + base, min(kStageRows, m - base), lane);
                } else {
                    const int node = ctrl->node, b = ctrl->buf;
                    int* sid = spec_id0 + (size_t)b * 2 * cand_cap;
                    float* sd = reinterpret_cast<float*>(sid + cand_cap);
                    if (warp == 1) {
                        const int* row = g.adj0 + (size_t)node * g.stride0;
                        int m = 0, deg = 0;
                        for (int rb = 0; rb < g.stride0; rb += 32) {
                            const int slot = rb + lane;
                            const int nb = slot < g.stride0 ? __ldg(row + slot) : -1;
                            deg += __popc(__ballot_sync(FULL, nb >= 0));
                            bool fresh = nb >= 0;
                            if (fresh) {
                                if (gvis) fresh = !((__ldcg(bm + (nb >> 5)) >> (nb & 31)) & 1u);
                                else fresh = vis_would_be_fresh<vis16>(vis, vis_size, vis_shift, nb);
                            }
                            const unsigned mk = __ballot_sync(FULL, fresh);
                            if (fresh) sid[m + __popc(mk & ((1u << lane) - 1u))] = nb;
                            m += __popc(mk);
                        }
                        if (lane == 0) {
                            ctrl->spec_m[b] = m;
                            ctrl->spec_deg[b] = deg;
                            ctrl->spec_node[b] = node;
                        }
                        __syncwarp();
                    }
                    if (H > 1) asm volatile("bar.sync 1, %0;" ::"r"(32 * H) : "memory");  // the helpers among themselves
                    const int m = ctrl->spec_m[b];
                    for (int base = kStageRows * (warp - 1); base < m; base += kStageRows * H)
                        batch_dist_to<METRIC, SPEC>(g, stg, q, qn, sid + base, sd + base, min(kStageRows, m - base), lane);
                }
                __syncthreads();  // (Y) the command is served
            }
            continue;
        }

        // ---------------- leader
        WarpCounters wc;
        WideList L;
        unsigned pf_issued = 0, pf_used = 0;
        long long pf_sel = 0, pf_wspec = 0, pf_weval = 0, pf_ins = 0;
        const long long t_start = clock64();
        int cur = g.entry;
        float cur_d = single_dist<METRIC, SPEC>(g, s, qn, cur, lane);
        wc.ndist++;
        for (int l = g.top_layer; l > 0; --l) greedy_layer<METRIC, SPEC>(g, s, qn, l, cur, cur_d, lane, wc);
        wl_init(L, list, ef);
        wl_insert(L, make_key(cur_d, cur), lane);
        if (lane == 0) {
            if (gvis) atomicOr(bm + (cur >> 5), 1u << (cur & 31));
            else if (vis16) vis16_test_and_set(s, cur);
            else vis_test_and_set(s, cur);
        }
        __syncwarp();
        int guess_id = -1, guess_nb = -1;
        bool outstanding = false;  // a speculative command is being served
        int spec_buf = 0;          // ... into this buffer
        const long long t_beam = clock64();
        for (;;) {
            u64 e;
            long long t0 = clock64();
            const bool more = wl_select(L, lane, e);
            if (!more) {
                if (outstanding) __syncthreads();  // (Y) of the speculation nobody needs
                break;
            }
            const int c = key_id(e);
            // The speculation must be complete before the table is written again: the helpers' read-only filter has to see
            // the table exactly as this (authoritative) filter is about to see it.
            bool spec_ready = false;
            if (outstanding) {
                const long long tw = clock64();
                __syncthreads();  // (Y)
                pf_wspec += clock64() - tw;
                t0 += clock64() - tw;
                outstanding = false;
                spec_ready = true;
            }
            int m = 0, deg = 0;
            const int* ids_use = cand_id;
            const float* dist = cand_dist;
            bool have = false;
            if (spec_ready && ctrl->spec_node[spec_buf] == c) {
                // The guessed node: the helpers filtered this very row against the visited set as it still is (only this warp
                // writes it, and only here), so their list IS the authoritative one.  The leader only records the new ids --
                // bitmap: fire-and-forget reductions, no L2 round trip on its critical path; table: one LDS + STS per id -- and
                // does not read the row again.
                const int* sid = spec_id0 + (size_t)spec_buf * 2 * cand_cap;
                m = ctrl->spec_m[spec_buf];
                deg = ctrl->spec_deg[spec_buf];
                for (int j = lane; j < m; j += 32) {
                    const int id = sid[j];
                    if (gvis) atomicOr(bm + (id >> 5), 1u << (id & 31));
                    else if (vis16) vis16_test_and_set(s, id);
                    else vis_test_and_set(s, id);
                }
                __syncwarp();
                ids_use = sid;
                dist = reinterpret_cast<const float*>(sid + cand_cap);
                have = true;
                pf_used++;
            } else {
            // row of c: keep the neighbours not seen recently, in stored order (the authoritative, table-updating filter)
            const int* row = g.adj0 + (size_t)c * g.stride0;
            for (int rb = 0; rb < g.stride0; rb += 32) {
                const int slot = rb + lane;
                int nb;
                if (guess_id == c && rb == 0) nb = guess_nb;  // fetched under the previous insertion
                else nb = slot < g.stride0 ? __ldg(row + slot) : -1;
                deg += __popc(__ballot_sync(FULL, nb >= 0));
                bool fresh = false;
                if (nb >= 0) {
                    if (gvis) fresh = !((atomicOr(bm + (nb >> 5), 1u << (nb & 31)) >> (nb & 31)) & 1u);
                    else fresh = vis16 ? vis16_test_and_set(s, nb) : vis_test_and_set(s, nb);
                }
                const unsigned mk = __ballot_sync(FULL, fresh);
                if (fresh) cand_id[m + __popc(mk & ((1u << lane) - 1u))] = nb;
                m += __popc(mk);
            }
            __syncwarp();
            }
            wc.nexp++;
            wc.row_words += deg + 1;
            wc.ndist += m;
            // distances: taken over from the speculation above, else evaluated on demand
            pf_sel += clock64() - t0;
            if (!have && m > 0) {
                const long long tw = clock64();
                if (lane == 0) {
                    ctrl->cmd = kCmdEval;
                    ctrl->m = m;
                }
                __syncthreads();  // (X)
                __syncthreads();  // (Y)
                pf_weval += clock64() - tw;
            }
            const long long t_ins = clock64();
            // the entry most likely to be expanded next, unless one of this expansion's candidates lands in front of it
            u64 pk2;
            const u64 pk = wl_peek(L, lane, &pk2);
            guess_id = pk != ~0ull ? key_id(pk) : -1;
            // the row of the entry after it will most likely be wanted one expansion later: pull it into L2 now
            if (pk2 != ~0ull && lane == 0) {
                const char* rp = reinterpret_cast<const char*>(g.adj0 + (size_t)key_id(pk2) * g.stride0);
                for (int off = 0; off < g.stride0 * 4; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + off));
            }
            if (guess_id >= 0) {
                if (SPECULATE) {
                    spec_buf ^= 1;
                    if (lane == 0) {
                        ctrl->cmd = kCmdSpec;
                        ctrl->node = guess_id;
                        ctrl->buf = spec_buf;
                    }
                    __syncthreads();  // (X)
                    outstanding = true;
                    pf_issued++;
                }
                guess_nb = lane < g.stride0 ? __ldg(g.adj0 + (size_t)guess_id * g.stride0 + lane) : -1;
            }
            if (m > 0) wl_insert_candidates<gvis>(L, ids_use, dist, 0, m, lane);
            pf_ins += clock64() - t_ins;
        }
        if (lane == 0) {
            const long long t_end = clock64();
            atomicAdd(&g_wide_prof[0], 1ull);
            atomicAdd(&g_wide_prof[1], (unsigned long long)wc.nexp);
            atomicAdd(&g_wide_prof[2], (unsigned long long)pf_issued);
            atomicAdd(&g_wide_prof[3], (unsigned long long)pf_used);
            atomicAdd(&g_wide_prof[4], (unsigned long long)(t_beam - t_start));
            atomicAdd(&g_wide_prof[5], (unsigned long long)pf_sel);
            atomicAdd(&g_wide_prof[6], (unsigned long long)pf_wspec);
            atomicAdd(&g_wide_prof[7], (unsigned long long)pf_weval);
            atomicAdd(&g_wide_prof[8], (unsigned long long)pf_ins);
            atomicAdd(&g_wide_prof[9], (unsigned long long)(t_end - t_beam));
        }
        if (lane == 0) ctrl->cmd = kCmdDone;
        __syncthreads();  // (X) releases the helpers to the next query

        wl_merge(L, lane);
        for (int j = lane; j < k; j += 32) {
            int id = -1;
            float d = CUDART_NAN_F;
            if (j < L.n) {
                const u64 ee = list[j];
                id = key_id(ee);
                d = key_dist(ee);
            }
            out_ids[(size_t)qi * k + j] = id;
            out_dists[(size_t)qi * k + j] = d;
        }
        if (stats != nullptr && lane == 0) {
            QueryStats qs;
            qs.ndist = wc.ndist; qs.nexp = wc.nexp; qs.row_words = wc.row_words;
            stats[qi] = qs;
        }
    }
}

static int pow2_floor(size_t v) {
    int p = 1;
    while ((size_t)p * 2 <= v) p <<= 1;
    return p;
}

struct WideShape {
    int per_sm, sms, max_optin;
};

template <typename Kern>
static WideShape wide_shape(Kern kern, int threads, size_t smem) {
    static std::mutex mu;
    static std::map<std::tuple<int, const void*, int, size_t>, WideShape> cache;
    static std::map<std::pair<int, const void*>, size_t> optin;
    int dev = 0;
    HNSW_CUDA_CHECK(cudaGetDevice(&dev));
    const void* fn = reinterpret_cast<const void*>(kern);
    const auto key = std::make_tuple(dev, fn, threads, smem);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    WideShape sh;
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&sh.max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&sh.sms, cudaDevAttrMultiProcessorCount, dev));
    size_t& have = optin[std::make_pair(dev, fn)];
    if (smem > have) {
        HNSW_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        have = smem;
    }
    sh.per_sm = 0;
    HNSW_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sh.per_sm, kern, threads, smem));
    if (sh.per_sm < 1) sh.per_sm = 1;
    cache.emplace(key, sh);
    return sh;
}

void wide_profile(unsigned long long out[16], bool reset) {
    HNSW_CUDA_CHECK(cudaMemcpyFromSymbol(out, g_wide_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {};
        HNSW_CUDA_CHECK(cudaMemcpyToSymbol(g_wide_prof, z, sizeof(z)));
    }
}

bool wide_search_applies(const GraphView& g, int nq, int ef, int layer, const Tuning& t) {
    if (layer != 0 || t.wide_min_ef <= 0) return false;
    if (ef >= t.wide_min_ef) return true;
    // tiny calls: a handful of queries cannot fill the GPU with one warp each; one CTA per query answers sooner
    return t.wide_max_small_batch > 0 && nq <= t.wide_max_small_batch;
}

void launch_knn_wide(int metric, const GraphView& g, const float* d_queries, int nq, int k, int ef, int* d_ids, float* d_dists,
                     unsigned* d_work_counter, QueryStats* d_stats, const Tuning& t, cudaStream_t stream) {
    if (nq <= 0) return;
    int dev = 0, max_optin = 0;
    HNSW_CUDA_CHECK(cudaGetDevice(&dev));
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int cand_cap = ((g.stride0 + 31) / 32) * 32;
    // warps per CTA: the leader + 1 helper (2 helpers for rows of more than 32 neighbours).  With full 10k-query batches on
    // 1M x 128 a second helper never pays for the resident queries its staging buffer costs (ef 1024: 254k vs 242k QPS)
    // exact visited bitmap in global memory instead of the shared-memory table: one bit per node per resident query must stay
    // L2-resident, so only up to wide_bitmap_max_nodes (2M nodes = 256 KB per query); larger indexes keep the table
    const bool bitmap_range = t.wide_prefetch != 0 && t.wide_bitmap_min_ef > 0 && ef >= t.wide_bitmap_min_ef && g.count <= t.wide_bitmap_max_nodes;
    // ... but a call with fewer queries than SMs leaves room for a second helper on every SM: a single query at ef 48 is answered in
    // 191 instead of 238 us of kernel time with it (a third helper adds nothing)
    int warps = t.wide_warps > 0 ? t.wide_warps : ((g.stride0 > 32 || nq <= device_sm_count()) ? 3 : 2);
    warps = std::max(2, std::min(kWarpsPerBlock, warps));
    // (the second helper's staging buffer must not be what pushes a very wide list out of shared memory)
    if (t.wide_warps <= 0 && warps == 3 && g.stride0 <= 32 && wide_smem_bytes(3, g.pitch, ef, bitmap_range ? 0 : 64, cand_cap) > (size_t)max_optin) warps = 2;
    // visited table: 4 ef entries when they fit (the setting of the one-warp kernel), else what the shared memory leaves
    int vis = t.vis_size_search > 0 ? std::max(64, pow2_floor((size_t)t.vis_size_search)) : std::max(256, pow2_floor((size_t)2 * ef));
    const bool can16 = (long long)g.count <= (64ll << 16);
    if (!can16 && t.vis_size_search <= 0) vis *= 2;  // 32-bit entries: one per word
    while (vis > 64 && wide_smem_bytes(warps, g.pitch, ef, vis, cand_cap) > (size_t)max_optin) vis >>= 1;
    const size_t smem = wide_smem_bytes(warps, g.pitch, ef, vis, cand_cap);
    if (smem > (size_t)max_optin)
        throw std::runtime_error("search width (MinNN / k) too large for shared memory: one query needs " + std::to_string(smem) +
                                 " B, the SM offers " + std::to_string(max_optin));
    const bool v16 = (long long)g.count <= ((long long)vis << 16);
    const bool shortrows = g.pitch <= kSegFloats;
    const bool dim128 = t.dim128 != 0 && g.dim == 128 && g.pitch == 128;  // constant trip counts in the distance loops (kSpecDim128)
    const bool prefetch = t.wide_prefetch != 0;  // speculative evaluation of the next expansion
    // exact bitmap in global memory instead of the shared-memory table (speculating kernels, wide beams)
    const bool use_bitmap = bitmap_range;
    const int bitmap_words = use_bitmap ? ((g.count + 127) / 128) * 4 : 0;
    size_t smem_bm = 0;
    if (use_bitmap) {
        smem_bm = wide_smem_bytes(warps, g.pitch, ef, 0, cand_cap);
        if (smem_bm > (size_t)max_optin) throw std::runtime_error("search width (MinNN / k) too large for shared memory");
    }
    uint32_t* d_bitmap = nullptr;
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_work_counter, 0, sizeof(unsigned), stream));
#define HNSW_WIDE_BM(M, S)                                                                                                         \
    {                                                                                                                              \
        const WideShape sh = wide_shape(knn_wide_kernel<M, S, true>, 32 * warps, smem_bm);                                         \
        int per_sm = sh.per_sm;                                                                                                    \
        if (t.blocks_per_sm > 0) per_sm = std::min(per_sm, t.blocks_per_sm);                                                       \
        const int grid = std::max(1, std::min(nq, sh.sms * per_sm));                                                               \
        HNSW_CUDA_CHECK(cudaMallocAsync(&d_bitmap, (size_t)grid * bitmap_words * sizeof(uint32_t), stream));                       \
        knn_wide_kernel<M, S, true><<<grid, 32 * warps, smem_bm, stream>>>(g, d_queries, nq, k, ef, 0, cand_cap, d_ids, d_dists,   \
                                                                           d_work_counter, d_stats, d_bitmap, bitmap_words);       \
        const cudaError_t launch_err = cudaGetLastError(); /* the bitmap goes back to the pool whether or not the launch took */ \
        HNSW_CUDA_CHECK(cudaFreeAsync(d_bitmap, stream));                                                                          \
        HNSW_CUDA_CHECK(launch_err);                                                                                               \
    }
    if (use_bitmap) {
        switch (metric) {
            case SQ_EUCLID: if (shortrows && dim128) HNSW_WIDE_BM(SQ_EUCLID, kSpecShortRows | kSpecGvis | kSpecDim128) else if (shortrows) HNSW_WIDE_BM(SQ_EUCLID, kSpecShortRows | kSpecGvis) else HNSW_WIDE_BM(SQ_EUCLID, kSpecGvis) break;
            case EUCLID: if (shortrows && dim128) HNSW_WIDE_BM(EUCLID, kSpecShortRows | kSpecGvis | kSpecDim128) else if (shortrows) HNSW_WIDE_BM(EUCLID, kSpecShortRows | kSpecGvis) else HNSW_WIDE_BM(EUCLID, kSpecGvis) break;
            case COSINE: if (shortrows && dim128) HNSW_WIDE_BM(COSINE, kSpecShortRows | kSpecGvis | kSpecDim128) else if (shortrows) HNSW_WIDE_BM(COSINE, kSpecShortRows | kSpecGvis) else HNSW_WIDE_BM(COSINE, kSpecGvis) break;
            default: if (shortrows && dim128) HNSW_WIDE_BM(UCOSINE, kSpecShortRows | kSpecGvis | kSpecDim128) else if (shortrows) HNSW_WIDE_BM(UCOSINE, kSpecShortRows | kSpecGvis) else HNSW_WIDE_BM(UCOSINE, kSpecGvis) break;
        }
        return;
    }
#undef HNSW_WIDE_BM
#define HNSW_WIDE_GO(M, S, P)                                                                                                      \
    {                                                                                                                              \
        const WideShape sh = wide_shape(knn_wide_kernel<M, S, P>, 32 * warps, smem);                                               \
        int per_sm = sh.per_sm;                                                                                                    \
        if (t.blocks_per_sm > 0) per_sm = std::min(per_sm, t.blocks_per_sm);                                                       \
        const int grid = std::max(1, std::min(nq, sh.sms * per_sm));                                                               \
        knn_wide_kernel<M, S, P><<<grid, 32 * warps, smem, stream>>>(g, d_queries, nq, k, ef, vis, cand_cap, d_ids, d_dists,       \
                                                                     d_work_counter, d_stats, nullptr, 0);                         \
    }
#define HNSW_WIDE_SPEC(M)                                                                                                          \
    {                                                                                                                              \
        if (shortrows && v16 && prefetch && dim128) HNSW_WIDE_GO(M, kSpecShortRows | kSpecVis16 | kSpecDim128, true)                  \
        else if (shortrows && v16) { if (prefetch) HNSW_WIDE_GO(M, kSpecShortRows | kSpecVis16, true) else HNSW_WIDE_GO(M, kSpecShortRows | kSpecVis16, false) } \
        else if (shortrows) { if (prefetch) HNSW_WIDE_GO(M, kSpecShortRows, true) else HNSW_WIDE_GO(M, kSpecShortRows, false) }     \
        else if (v16) { if (prefetch) HNSW_WIDE_GO(M, kSpecVis16, true) else HNSW_WIDE_GO(M, kSpecVis16, false) }                   \
        else { if (prefetch) HNSW_WIDE_GO(M, 0, true) else HNSW_WIDE_GO(M, 0, false) }                                              \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_WIDE_SPEC(SQ_EUCLID) break;
        case EUCLID: HNSW_WIDE_SPEC(EUCLID) break;
        case COSINE: HNSW_WIDE_SPEC(COSINE) break;
        default: HNSW_WIDE_SPEC(UCOSINE) break;
    }
#undef HNSW_WIDE_SPEC
#undef HNSW_WIDE_GO
    HNSW_CUDA_CHECK(cudaGetLastError());
}

}  // namespace hnsw
