// build_kernels.cu -- K4/K5/K6: batched insertion on sm_100a, one warp per new node.
//
// Replaces the linking half of HNSWIndex.Add (HNSWIndex.cs:55-78):
//   GraphConnector.AddNewConnections   GraphConnector.cs:172-181  (descent, then per layer top..0)
//   GraphConnector.ConnectAtLayer      GraphConnector.cs:187-217  (SearchLayer with ef = MaxCandidates,
//                                                                  heuristic, own row, reverse links)
//   Heuristic.RelativeNeighborPruning  Heuristic.cs:11-46
//   GraphConnector.PruneOverflow       GraphConnector.cs:222-262
// The reference inserts with Parallel.For under per-node Monitor locks.  Here a batch of new
// nodes is handled by two kernels:
//   select_kernel : every new node searches the graph as it stood before the batch (nodes of the
//                   same batch are unreachable, exactly like nodes still being linked in the
//                   reference), runs the heuristic and writes its own rows;
//   link_kernel   : for every selected neighbour n, under a per-node spin lock (atomicCAS on
//                   locks[n]), append the new id to n's row, re-running the heuristic on
//                   overflow so MaxEdges(layer) is honoured.
// Tail fill (default): a warp of the select kernel handles ~6 insertions of ~1.7 ms each per full batch, so for the last
// ~10 % of the kernel's time the SMs drain.  The link kernel is therefore a programmatic dependent launch: its blocks move
// into the slots the first finishing select blocks leave, and each work item waits for its node's ready flag (written by the
// select warp after the node's rows) instead of for the whole select kernel.  The last insertions of a batch can then see
// reverse links of its first ones -- as threads of the reference's Parallel.For do.
#include <algorithm>

#include "build_scratch.cuh"
#include "kernels.hpp"
#include "launch.hpp"

namespace hnsw {

// Cycle / work counters of the construction kernels (d_build_prof, 16 words; Index::build_profile()).
enum BuildProf : int {
    BP_INSERTS = 0, BP_CYC_DESCENT, BP_CYC_SEARCH, BP_CYC_SELECT, BP_NDIST_SEARCH, BP_NDIST_SELECT,
    BP_CYC_LINK, BP_CYC_PRUNE, BP_PRUNES, BP_NDIST_PRUNE, BP_CYC_LOCK, BP_CYC_SELECT_KERNEL, BP_LINKS, BP_WORDS
};

template <int METRIC, int SPEC>
__device__ __forceinline__ void build_select_body(const GraphView& g, const int* __restrict__ levels, const int* __restrict__ new_ids, int count, int efc,
                                                  int list_cap, int vis_size, int cand_cap, int sel_cap, unsigned* work_counter, u64* prof, int flags,
                                                  int* ready, int epoch) {
    extern __shared__ __align__(128) unsigned char smem[];
    // Tail fill (ready != null): the link kernel that follows in the stream is a programmatic dependent launch.  Its blocks
    // may become resident as soon as every block of THIS grid has started, i.e. into the slots the first finishing select
    // blocks leave, and each of its work items waits for its node's ready flag instead of for this whole kernel.
    if (ready != nullptr) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = build_scratch_bytes(g.pitch, list_cap, vis_size, cand_cap, sel_cap);
    BuildScratch b = carve_build(smem + warp * per_warp, g.pitch, list_cap, vis_size, cand_cap, sel_cap, lane);
    WarpScratch& s = b.s;
    // Profile counters go to global memory as they arise (a few atomics per 1.7 ms insertion): nothing of them is carried
    // across the searches of this issue- and register-bound kernel.

    for (;;) {
        unsigned wi = 0;
        if (lane == 0) wi = atomicAdd(work_counter, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= (unsigned)count) break;
        const int v = __ldg(new_ids + wi);
        const int level = __ldg(levels + v);
        long long t0 = clock64();
        load_vector(g, v, s.q, lane);
        double qn = 0.0;
        if (METRIC == COSINE) qn = __ldg(g.norm_sqrt + v);

        WarpCounters wc;
        int cur = g.entry;
        float cur_d = single_dist<METRIC, SPEC>(g, s, qn, cur, lane);
        for (int layer = g.top_layer; layer > level; --layer) greedy_layer<METRIC, SPEC>(g, s, qn, layer, cur, cur_d, lane, wc);
        long long t1 = clock64();
        if (lane == 0) {
            atomicAdd(prof + BP_INSERTS, (u64)1);
            atomicAdd(prof + BP_CYC_DESCENT, (u64)(t1 - t0));
            atomicAdd(prof + BP_CYC_SELECT_KERNEL, (u64)(t1 - t0));  // busy cycles of the warp: descent + searches + heuristic
        }

        for (int layer = min(level, g.top_layer); layer >= 0; --layer) {
            BeamState st;
            t0 = clock64();
            beam_search<METRIC, SPEC>(g, s, qn, layer, cur, cur_d, efc, st, lane, wc, flags);
            t1 = clock64();
            if (lane == 0) {
                atomicAdd(prof + BP_CYC_SEARCH, (u64)(t1 - t0));
                atomicAdd(prof + BP_CYC_SELECT_KERNEL, (u64)(t1 - t0));
            }
            unsigned nd_select = 0;
            const int maxE = layer == 0 ? 2 * g.max_edges : g.max_edges;
            int nsel;
            int next;
            float next_d;
            if (st.n < maxE) {
                // Heuristic.cs:13-18 returns the candidates unsorted; ConnectAtLayer then continues from
                // element [0] of that array, the root of the max-heap = the FARTHEST candidate.
                nsel = st.n;
                for (int j = lane; j < nsel; j += 32) b.sel[j] = key_id(s.list[j]);
                next = key_id(s.list[st.n - 1]);
                next_d = key_dist(s.list[st.n - 1]);
                __syncwarp();
            } else {
                nsel = select_neighbors<METRIC, SPEC>(g, b, s.list, st.n, maxE, lane, nd_select);
                next = key_id(s.list[0]);  // the closest candidate is always selected first
                next_d = key_dist(s.list[0]);
            }
            if (lane == 0) {
                const u64 dt = (u64)(clock64() - t1);
                atomicAdd(prof + BP_CYC_SELECT, dt);
                atomicAdd(prof + BP_CYC_SELECT_KERNEL, dt);
                atomicAdd(prof + BP_NDIST_SELECT, (u64)nd_select);
            }
            int stride;
            int* row = row_ptr_mut(g, v, layer, stride);
            for (int j = lane; j < stride; j += 32) row[j] = j < nsel ? b.sel[j] : -1;
            __syncwarp();
            cur = next;
            cur_d = next_d;
        }
        if (lane == 0) atomicAdd(prof + BP_NDIST_SEARCH, (u64)wc.ndist);
        if (ready != nullptr) {
            // every lane's row stores, then the flag (release).  Indexed by node id, not by batch position: `v` is live here anyway
            __threadfence();
            __syncwarp();
            if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ready + v), "r"(epoch) : "memory");
        }
        __syncwarp();
    }
}

// Four blocks of four warps per SM.  (A shape of six three-warp blocks -- 18 instead of 16 resident warps -- was the default for
// a while: under __launch_bounds__(96, 6) ptxas settled on 96 registers, and on that edge the compile was bimodal: one mode
// moved the loop control of the distance loops to the uniform datapath and ran 4 % faster than this shape, the other -- reached
// through unrelated edits such as two more words in WarpScratch or one more value live across the searches -- reported no spills
// either and ran 2.4x slower.  With a stated budget of 112 registers the shape equals this one; the constant trip counts of
// kSpecDim128 give what the lucky mode gave, reliably.  Removed.)
template <int METRIC, int SPEC>
__global__ void __launch_bounds__(kThreads, 4)
build_select_kernel(GraphView g, const int* __restrict__ levels, const int* __restrict__ new_ids, int count, int efc, int list_cap, int vis_size, int cand_cap,
                    int sel_cap, unsigned* work_counter, u64* prof, int flags, int* ready, int epoch) {
    build_select_body<METRIC, SPEC>(g, levels, new_ids, count, efc, list_cap, vis_size, cand_cap, sel_cap, work_counter, prof, flags, ready, epoch);
}

__device__ __forceinline__ bool lock_node(int* locks, int n, int lane, int* error_flag) {
    int ok = 1;
    if (lane == 0) {
        unsigned spins = 0;
        while (atomicCAS(locks + n, 0, 1) != 0) {
            __nanosleep(128);
            if (++spins > (1u << 22)) { ok = 0; atomicExch(error_flag, 1); break; }  // never hang the GPU
        }
    }
    ok = __shfl_sync(FULL, ok, 0);
    __threadfence();
    return ok != 0;
}
__device__ __forceinline__ void unlock_node(int* locks, int n, int lane) {
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(locks + n, 0);
}

template <int METRIC, int SPEC = 0>
__global__ void __launch_bounds__(kThreads, 4)
build_link_kernel(GraphView g, const int* __restrict__ levels, const int* __restrict__ new_ids, int count, int list_cap, int cand_cap, int sel_cap,
                  unsigned* work_counter, int* error_flag, u64* prof, const int* ready, int epoch, const unsigned* sel_counter,
                  unsigned sel_fetches) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = build_scratch_bytes(g.pitch, list_cap, 64, cand_cap, sel_cap);
    BuildScratch b = carve_build(smem + warp * per_warp, g.pitch, list_cap, 64, cand_cap, sel_cap, lane);
    WarpScratch& s = b.s;
    long long cyc_prune = 0, cyc_lock = 0;
    unsigned n_prunes = 0, nd_prune = 0, n_links = 0;
    const long long t_begin = clock64();

    // one reverse link: append v to the row of n on `layer`, or re-run the heuristic when the row is full
    auto link_one = [&](int v, int n, int layer) {
        const int maxE = layer == 0 ? 2 * g.max_edges : g.max_edges;
        const long long tl = clock64();
        if (!lock_node(g.locks, n, lane, error_flag)) return;
        cyc_lock += clock64() - tl;
        ++n_links;
        int stride;
        int* row = row_ptr_mut(g, n, layer, stride);
        // current row of n, read past L1 (other SMs update it under the same lock)
        int cnt = 0;
        bool dup = false;
        for (int rb = 0; rb < stride; rb += 32) {
            const int slot = rb + lane;
            const int nb = slot < stride ? __ldcg(row + slot) : -1;
            if (nb >= 0) s.cand_id[slot] = nb;
            cnt += __popc(__ballot_sync(FULL, nb >= 0));
            dup = dup || __any_sync(FULL, nb == v);
        }
        __syncwarp();
        if (dup) {
            // only with tail fill: n is a node of this very batch that had already been linked to v (v's link reached n's row
            // first and was then read back as one of n's own targets) -- the edge exists, nothing to do
        } else if (cnt < maxE) {
            if (lane == 0) row[cnt] = v;  // EdgeList.Add, GraphConnector.cs:207
        } else {
            // PruneOverflow (GraphConnector.cs:222-262): heuristic over the row plus the new link
            const long long tp = clock64();
            ++n_prunes;
            if (lane == 0) s.cand_id[cnt] = v;
            const int m = cnt + 1;
            load_vector(g, n, s.q, lane);
            double nn = 0.0;
            if (METRIC == COSINE) nn = __ldg(g.norm_sqrt + n);
            batch_dist<METRIC, SPEC>(g, s, nn, m, lane);
            nd_prune += m;
            BeamState st;
            st.n = 0; st.ef = list_cap; st.far_ord = 0xFFFFFFFFu; st.cursor = 0;
            for (int j = 0; j < m; ++j) list_insert(s.list, st, make_key(s.cand_dist[j], s.cand_id[j]), lane);
            const int nsel = select_neighbors<METRIC, SPEC>(g, b, s.list, st.n, maxE, lane, nd_prune);
            for (int j = lane; j < stride; j += 32) row[j] = j < nsel ? b.sel[j] : -1;
            cyc_prune += clock64() - tp;
        }
        unlock_node(g.locks, n, lane);
    };

    // Work items: [0, count x stride0) = one (new node, layer-0 slot) pair each -- a reverse link is a short, latency-bound
    // critical section (lock, read the row, append or prune), so a warp per PAIR instead of a warp per node walking its
    // 32 slots one after another keeps ~30x more of them in flight and balances the rare, long prunes; a new node's own
    // rows are never a link target inside its batch, so the pairs are independent.  [count x stride0, + count) = the upper
    // layers of one new node each (6 % of the nodes have any).
    const unsigned pairs = (unsigned)count * (unsigned)g.stride0;
    for (;;) {
        unsigned wi = 0;
        if (lane == 0) wi = atomicAdd(work_counter, 1u);
        wi = __shfl_sync(FULL, wi, 0);
        if (wi >= pairs + (unsigned)count) {
            // Tail fill: this kernel must not end before every select warp has done its last (empty) fetch from the select work
            // counter -- the next batch resets that word.  The one warp that draws the first empty index of THIS queue waits
            // until the select counter shows all of them: count real fetches + one empty fetch per select warp.
            if (ready != nullptr && wi == pairs + (unsigned)count && lane == 0) {
                unsigned spins = 0, nap = 64, seen;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(sel_counter) : "memory");
                    if (seen >= sel_fetches) break;
                    __nanosleep(nap);
                    if (nap < 4096) nap <<= 1;
                    if (++spins > (1u << 20)) { atomicExch(error_flag, 1); break; }
                }
            }
            break;
        }
        const int v = __ldg(new_ids + (wi < pairs ? wi / (unsigned)g.stride0 : wi - pairs));
        if (ready != nullptr) {  // tail fill: this node's rows may still be in the making (select kernel still running)
            int ok = 1;
            if (lane == 0) {
                unsigned spins = 0, nap = 256;
                int f;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(ready + v) : "memory");
                    if (f == epoch) break;
                    __nanosleep(nap);
                    if (nap < 4096) nap <<= 1;
                    if (++spins > (1u << 20)) { ok = 0; atomicExch(error_flag, 1); break; }  // ~4 s: never hang the GPU
                }
            }
            ok = __shfl_sync(FULL, ok, 0);
            if (!ok) break;
        }
        if (wi < pairs) {
            const int n = __ldcg(g.adj0 + (size_t)v * g.stride0 + wi % (unsigned)g.stride0);
            if (n >= 0) link_one(v, n, 0);
        } else {
            const int level = __ldg(levels + v);
            for (int layer = min(level, g.top_layer); layer >= 1; --layer) {
                int vstride;
                const int* vrow = row_ptr_mut(g, v, layer, vstride);
                for (int t = 0; t < vstride; ++t) {
                    const int n = __ldcg(vrow + t);
                    if (n < 0) break;
                    link_one(v, n, layer);
                }
            }
        }
        __syncwarp();
    }
    if (lane == 0 && n_links > 0) {
        atomicAdd(prof + BP_CYC_LINK, (u64)(clock64() - t_begin));
        atomicAdd(prof + BP_CYC_PRUNE, (u64)cyc_prune);
        atomicAdd(prof + BP_CYC_LOCK, (u64)cyc_lock);
        atomicAdd(prof + BP_PRUNES, (u64)n_prunes);
        atomicAdd(prof + BP_NDIST_PRUNE, (u64)nd_prune);
        atomicAdd(prof + BP_LINKS, (u64)n_links);
    }
}

static int pow2_ge(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

template <typename Kern>
static int persistent_grid(Kern kern, size_t smem_per_warp, const Tuning& t, int work_items, int& warps) {
    return persistent_launch(kern, smem_per_warp, t.blocks_per_sm, work_items, warps);
}

// d_work_counter: [0] select work, [1] link work, [2] error flag
void launch_build_batch(int metric, const GraphView& g, const int* d_levels, const int* d_new_ids, int count, int efc, unsigned* d_work_counter,
                        u64* d_build_prof, const Tuning& t, cudaStream_t stream,
                        cudaStream_t link_stream, cudaEvent_t select_done, int* d_ready, int epoch) {
    if (count <= 0) return;
    if (link_stream == nullptr) link_stream = stream;
    // tail fill: both kernels in one stream; not for a handful of nodes (one Add per call: nothing to fill, and the waiting link
    // blocks cost the lone select warp 70 us)
    if (link_stream != stream || t.build_tail_fill == 0 || count < 64) d_ready = nullptr;
    const int maxE0 = 2 * g.max_edges;
    const int list_cap_sel = std::max(efc, 1);
    const int list_cap_link = maxE0 + 1;
    const int cand_cap = ((std::max(g.stride0, g.strideU) + 1 + 31) / 32) * 32;
    const int sel_cap = std::max(maxE0, list_cap_sel < maxE0 ? list_cap_sel : maxE0);
    // small table: shared memory per warp decides the occupancy of this latency-bound kernel (see search_kernels.cu)
    int vis = t.vis_size_build > 0 ? std::max(64, pow2_ge(t.vis_size_build)) : std::min(2048, std::max(256, pow2_ge(2 * efc)));
    // 16-bit tags (beam.cuh vis16_test_and_set) hold two entries per word: half the words keep the same filter and let a
    // fourth block fit on the SM
    if (t.vis_size_build <= 0 && vis > 256 && (long long)g.count <= ((long long)(vis / 2) << 16)) vis /= 2;
    const bool vis16 = (long long)g.count <= ((long long)vis << 16);
    const size_t smem_sel = build_scratch_bytes(g.pitch, list_cap_sel, vis, cand_cap, sel_cap);   // per warp
    const size_t smem_link = build_scratch_bytes(g.pitch, list_cap_link, 64, cand_cap, sel_cap);  // per warp
    unsigned* sel_counter = d_work_counter;
    int sel_warps = 0;  // warps of the select launch (the link kernel waits for the last fetch of every one of them, see there)
    HNSW_CUDA_CHECK(cudaMemsetAsync(sel_counter, 0, sizeof(unsigned), stream));
    HNSW_CUDA_CHECK(cudaMemsetAsync(d_work_counter + 1, 0, sizeof(unsigned), link_stream));
    int* err = reinterpret_cast<int*>(d_work_counter + 2);
#define HNSW_LAUNCH_LINK(M, SPECV)                                                                                             \
    {                                                                                                                           \
        constexpr int SL = SPECV;                                                                                               \
        const int g2 = persistent_grid(build_link_kernel<M, SL>, smem_link, t, (int)std::min<long long>((long long)count * (g.stride0 + 1), 1 << 30), w2); \
        if (d_ready != nullptr) {                                                                                               \
            /* programmatic dependent launch: resident as soon as all select blocks have started and SM slots are free */      \
            cudaLaunchConfig_t cfg = {};                                                                                        \
            cfg.gridDim = dim3((unsigned)g2);                                                                                   \
            cfg.blockDim = dim3(32u * (unsigned)w2);                                                                            \
            cfg.dynamicSmemBytes = smem_link * w2;                                                                              \
            cfg.stream = stream;                                                                                                \
            cudaLaunchAttribute attr[1];                                                                                        \
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                                    \
            attr[0].val.programmaticStreamSerializationAllowed = 1;                                                             \
            cfg.attrs = attr;                                                                                                   \
            cfg.numAttrs = 1;                                                                                                   \
            cudaError_t pe = cudaLaunchKernelEx(&cfg, build_link_kernel<M, SL>, g, d_levels, d_new_ids, count, list_cap_link, cand_cap, sel_cap, \
                                                d_work_counter + 1, err, d_build_prof, (const int*)d_ready, epoch, (const unsigned*)sel_counter, \
                                                (unsigned)(count + sel_warps));            \
            if (pe == cudaErrorNotSupported || pe == cudaErrorInvalidValue) {                                                   \
                /* a driver without programmatic stream serialization: plain launch, the ready flags are then always set */    \
                (void)cudaGetLastError();                                                                                       \
                cfg.numAttrs = 0;                                                                                               \
                pe = cudaLaunchKernelEx(&cfg, build_link_kernel<M, SL>, g, d_levels, d_new_ids, count, list_cap_link, cand_cap, sel_cap, \
                                        d_work_counter + 1, err, d_build_prof, (const int*)d_ready, epoch, (const unsigned*)sel_counter, \
                                                (unsigned)(count + sel_warps));                    \
            }                                                                                                                   \
            HNSW_CUDA_CHECK(pe);                                                                                                \
        } else {                                                                                                                \
            build_link_kernel<M, SL><<<g2, 32 * w2, smem_link * w2, link_stream>>>(g, d_levels, d_new_ids, count, list_cap_link, cand_cap, sel_cap, \
                                                                               d_work_counter + 1, err, d_build_prof, nullptr, 0, nullptr, 0u); \
        }                                                                                                                       \
    }
#define HNSW_LAUNCH_SELECT_C5(M, SPECV)                                                                                        \
    {                                                                                                                           \
        constexpr int S = SPECV;                                                                                                \
        const int g1 = persistent_grid(build_select_kernel<M, S>, smem_sel, t, count, w1);                                     \
        sel_warps = g1 * w1;                                                                                                    \
        build_select_kernel<M, S><<<g1, 32 * w1, smem_sel * w1, stream>>>(                                                      \
            g, d_levels, d_new_ids, count, efc, list_cap_sel, vis, cand_cap, sel_cap, sel_counter, d_build_prof, t.build_flags, d_ready, epoch); \
    }
#define HNSW_LAUNCH_BUILD(M)                                                                                                    \
    {                                                                                                                           \
        int w1 = 0, w2 = 0;                                                                                                     \
        if (vis16 && g.pitch <= kSegFloats && list_cap_sel <= 1024 && t.build_specialized != 0) {                                \
            /* the C5 shape: rows of one staging segment, MaxCandidates <= 1024, 16-bit tags -- no long-row / large-list code; \
               and for vectors of exactly 128 floats the distance loops have compile-time trip counts */                       \
            if (g.dim == 128 && g.pitch == 128 && t.dim128 != 0)                                                          \
                HNSW_LAUNCH_SELECT_C5(M, (kSpecVis16 | kSpecShortRows | kSpecMidList | kSpecDim128))                            \
            else                                                                                                                \
                HNSW_LAUNCH_SELECT_C5(M, (kSpecVis16 | kSpecShortRows | kSpecMidList))                                          \
        } else if (vis16) {                                                                                                     \
            const int g1 = persistent_grid(build_select_kernel<M, kSpecVis16>, smem_sel, t, count, w1);                        \
            sel_warps = g1 * w1;                                                                                                \
            build_select_kernel<M, kSpecVis16><<<g1, 32 * w1, smem_sel * w1, stream>>>(                                        \
                g, d_levels, d_new_ids, count, efc, list_cap_sel, vis, cand_cap, sel_cap, sel_counter, d_build_prof, t.build_flags, d_ready, epoch); \
        } else {                                                                                                                \
            const int g1 = persistent_grid(build_select_kernel<M, 0>, smem_sel, t, count, w1);                                 \
            sel_warps = g1 * w1;                                                                                                \
            build_select_kernel<M, 0><<<g1, 32 * w1, smem_sel * w1, stream>>>(                                                 \
                g, d_levels, d_new_ids, count, efc, list_cap_sel, vis, cand_cap, sel_cap, sel_counter, d_build_prof, t.build_flags, d_ready, epoch); \
        }                                                                                                                       \
        if (link_stream != stream) { /* the links of this batch follow its select kernel, on their own stream */             \
            HNSW_CUDA_CHECK(cudaEventRecord(select_done, stream));                                                             \
            HNSW_CUDA_CHECK(cudaStreamWaitEvent(link_stream, select_done, 0));                                                 \
        }                                                                                                                       \
        if (g.dim == 128 && g.pitch == 128 && t.dim128 != 0) HNSW_LAUNCH_LINK(M, (kSpecShortRows | kSpecDim128))                 \
        else HNSW_LAUNCH_LINK(M, 0)                                                                                             \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_LAUNCH_BUILD(SQ_EUCLID) break;
        case EUCLID: HNSW_LAUNCH_BUILD(EUCLID) break;
        case COSINE: HNSW_LAUNCH_BUILD(COSINE) break;
        default: HNSW_LAUNCH_BUILD(UCOSINE) break;
    }
#undef HNSW_LAUNCH_BUILD
#undef HNSW_LAUNCH_SELECT_C5
#undef HNSW_LAUNCH_LINK
    HNSW_CUDA_CHECK(cudaGetLastError());
}

}  // namespace hnsw
