// kernels.hpp -- host-callable launchers of the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace hnsw {

struct Tuning {
    int vis_size_search = 0;  // 0 = derive from ef
    int vis_size_build = 0;
    int vis_global_search = 0;  // 1: long-row (dim > 128) one-warp searches keep an EXACT visited bitmap per resident warp in global memory
                                // behind the shared-memory table; the look-ups travel with the first 128 floats of the first 16 rows
                                // (beam.cuh batch_dist_probe), the rest of the candidates is compacted before its passes are formed.
                                // 1M x 960, k = 100, ef 192: evaluations 11 288 -> 10 210 per query (the oracle's 10 211), but the
                                // bitmaps of ~1 800 resident warps (222 MB) do not stay in L2 and their 32-byte DRAM accesses cost
                                // what the saved rows gave: 72.2 vs 68.6-69.7 ms, 68.8 with a 256-word table -- off
    int build_overlap = 0;    // large insertion batches: run the link kernel on a second stream so that the next batch's select
                              // kernel fills the SM slots freed in the tails.  Off: -4.5 % build time, but the next batch
                              // then sees the previous batch's reverse links only partly and recall@10 drops 0.2 pt
    int build_tail_fill = 1;  // the link kernel of a batch is a programmatic dependent launch that moves into the SM slots the
                              // select kernel's first finishing blocks leave and waits for each node's ready flag, instead of
                              // starting when the last insertion of the batch has ended (a batch is ~6 insertions deep per warp, so
                              // the select kernel's tail is ~8 % of it)
    int build_flags = 2;        // beam-search flags of the construction search (bit 0: L2-prefetch rows of accepted candidates, bit 1: guess-row load)
    int build_specialized = 1;  // the C5-shaped select kernel instantiation (short rows, MaxCandidates <= 1024, 16-bit tags)
    int dim128 = 1;             // vectors of exactly 128 floats: kernel instantiations whose distance loops have compile-time trip counts
                                // (construction, reverse links, searches with ef > 64; the issue-bound kernels feel their loop control)
    int sync_pieces = 0;      // hnsw_knn_query: pieces a large batch is split into, alternating two streams (0 = 2 for page-locked, 4 for pageable buffers)
    int async_overlap = 1;    // hnsw_knn_query_device_async alternates two streams (consecutive batches overlap in their tails)
    int blocks_per_sm = 0;    // 0 = occupancy query
    int exact_tcgen05 = 2;    // exact_knn_tf32: pass A on tcgen05 / TMEM where the shape allows (rows <= 128 floats), else mma.sync;
                              // 2 = warp-specialised kernel (4 producer + 4 consumer warps; 49.8 ms for 10k x 1M x 128 on B200),
                              // 1 = the earlier 4-warp tcgen05 kernel (68.8 ms), 0 = always mma.sync
    int wl_min_ef = 256;      // one-warp-per-query searches with ef at or above this keep their candidates in wide.cuh's non-shifting
                              // list (0 = never): the shifting list of beam.cuh costs O(ef / 32) chunk moves per accepted candidate
    int wide_min_ef = 384;     // layer-0 searches with ef = max(MinNN, k) at or above this run on the CTA-per-query kernel with the
                              // non-shifting list (wide_kernels.cu); 0 = never
    int wide_warps = 0;       // warps per CTA of that kernel: the leader + helpers (0 = 2, or 3 for rows of more than 32 neighbours)
    int wide_prefetch = 1;    // its helper warps evaluate the likely next expansion ahead of time (speculative, never changes results)
    int wide_bitmap_min_ef = 1024;  // CTA-per-query searches at or above this width keep an EXACT visited bitmap in global memory (one bit per
                                    // node per resident query, L2-resident) instead of the shared-memory table: no re-evaluations, no duplicate
                                    // look-ups, and the table's shared memory goes to more resident queries (3 000-query batches on 1M x 128:
                                    // ef 1024 255k -> 299-333k QPS, 2048 85k -> 133k, 4096 25.7k -> 40k, 8192 4.3k -> 13.8k; below 1024 the table wins); 0 = never
    int wide_bitmap_max_nodes = 2097152;  // ... only for indexes up to this many nodes (the bitmaps of all resident queries should stay in L2)
    int wide_max_small_batch = 4;  // calls with at most this many queries use the CTA-per-query kernel whatever ef is: one query alone
                                   // is answered ~15 % sooner (177 vs 215 us of kernel time at ef = 48 on the 1M x 128 index)
    int peer_timeout_ms = 10000;  // partitioned index: how long the wait+merge kernel waits for a peer GPU's results before it
                                  // gives up (error + poisoned outputs).  Wall time; raise it for very large / high-ef batches
    int flags = 2;            // bit 0: L2-prefetch rows of accepted candidates (no gain); bit 1: guess-row load
};

// per-query counters written by the search kernel when requested
struct QueryStats {
    unsigned ndist, nexp, row_words;
};

int device_sm_count();

// K2: batched KnnQuery (HNSWIndex.cs:107-137).  queries: device [nq][dim] dense fp32;
// out_ids/out_dists: device [nq][k], rows ascending, padded with -1 / NaN.
// `layer`: the layer searched (HNSWIndex.cs:107 `layer` argument; 0 = base layer), descent stops above it.
void launch_knn_search(int metric, const GraphView& g, const float* d_queries, int nq, int k, int ef, int layer, int* d_ids, float* d_dists,
                       unsigned* d_work_counter, QueryStats* d_stats, const Tuning& t, cudaStream_t stream);
// lets the device's default memory pool keep freed scratch (the per-launch visited bitmaps) instead of releasing it at every sync
void retain_freed_scratch();

// K2w: the same search with one CTA per query and a list that does not shift per insertion (wide_kernels.cu); layer 0 only.
// launch_knn_search routes to it when wide_search_applies().
bool wide_search_applies(const GraphView& g, int nq, int ef, int layer, const Tuning& t);
void wide_profile(unsigned long long out[16], bool reset);  // leader-side cycle / speculation counters of K2w (current device)
void launch_knn_wide(int metric, const GraphView& g, const float* d_queries, int nq, int k, int ef, int* d_ids, float* d_dists,
                     unsigned* d_work_counter, QueryStats* d_stats, const Tuning& t, cudaStream_t stream);

// MultiLayerKnnQuery (HNSWIndex.cs:173-187): out_* are [nq][top_l + 1][k], rows of layers below min_layer all padding,
// each row the layer's sorted result without its first entry.
void launch_multilayer_search(int metric, const GraphView& g, const float* d_queries, int nq, int k, int top_l, int min_layer, int* d_ids,
                              float* d_dists, unsigned* d_work_counter, const Tuning& t, cudaStream_t stream);

// per-item sqrt((double)normA) in the cosine metric's accumulation order
void launch_item_norms(const float* d_vectors, int pitch, int dim, int first, int count, double* d_norm_sqrt, cudaStream_t stream);
void launch_item_norms_ids(const float* d_vectors, int pitch, int dim, const int* d_ids, int count, double* d_norm_sqrt, cudaStream_t stream);
// Add into re-used slots: rows src_row[j] of the staged batch d_src [.][dim] -> slab rows d_ids[j]; levels / upper_off of those ids
void launch_scatter_new_items(const float* d_src, const int* d_src_row, const int* d_ids, const int* d_lv, const int* d_uoff, int n, int dim,
                              int pitch, float* d_vectors, int* d_levels, int* d_upper_off, cudaStream_t stream);

// K4+K5: for the new nodes d_new_ids[0..count): descent, per-layer beam search (ef = MaxCandidates) against the
// graph as it stands, relative-neighbour selection, write the node's own rows.
// K6: reverse links with per-node locks and overflow pruning.
// d_build_prof: 16 u64 cycle / work counters accumulated by the kernels (build_kernels.cu BuildProf).
// d_ready / epoch (tail fill, Tuning::build_tail_fill): one word per node id, set to `epoch` by the select kernel
// when the node's rows are written; the link kernel then starts under the select kernel's tail and waits per node.
void launch_build_batch(int metric, const GraphView& g, const int* d_levels, const int* d_new_ids, int count, int ef_construction,
                        unsigned* d_work_counter, u64* d_build_prof, const Tuning& t, cudaStream_t stream,
                        cudaStream_t link_stream = nullptr, cudaEvent_t select_done = nullptr, int* d_ready = nullptr, int epoch = 0);

// RangeQuery (HNSWIndex.cs:144-168): traversal of all in-range nodes reachable from the layer-0 entry.  Work item w handles
// query qmap[w] (or w when qmap is null); out_keys [nwork][cap] unsorted (distance|id) keys, out_count[w] = result size or -1
// when cap / the visited table (hash [nwork][hash_size], power of two) overflowed.
void launch_range_search(int metric, const GraphView& g, const float* d_queries, const int* d_qmap, int nwork, float range, int cap,
                         int hash_size, u64* d_out_keys, uint32_t* d_hash, int* d_out_count, unsigned* d_work_counter, const Tuning& t,
                         cudaStream_t stream);

// per query: sort the result keys ascending on the device and pack them at d_offsets[w] of d_dense (results larger
// than 2048 keys are packed unsorted and flagged in d_unsorted_flag for the host)
void launch_range_finalize(const u64* d_keys, int cap, const int* d_counts, const long long* d_offsets, int nwork, u64* d_dense,
                           int* d_unsorted_flag, cudaStream_t stream);

// K3: exact brute-force top-k (recall ground truth), same metric arithmetic.
void launch_exact_knn(int metric, const GraphView& g, const float* d_queries, int nq, int k, int* d_ids, float* d_dists,
                      void* d_scratch, size_t scratch_bytes, cudaStream_t stream);
size_t exact_knn_scratch_bytes(int nq, int k, int count);

// K3b: the same answers through a 3xTF32 tensor-core GEMM filter + bit-exact re-ranking (exact_tf32_kernels.cu).
// d_fail[q] = 1 where the certificate did not hold (the caller re-runs those queries with launch_exact_knn).
// refresh_aux: recompute the per-item norms kept at the start of the scratch (needed after the index changed).
int exact_tf32_candidates(int k);
int exact_tf32_block_queries(int dim, int k);  // 64 / 32 / 16, or 0 when dim / k do not fit the variant
size_t exact_tf32_scratch_bytes(int nq, int k, int count);
// tcgen05_mode: 0 = mma.sync; 1 = pass A as tcgen05.mma kind::tf32 with the accumulator in TMEM (rows <= 128 floats, else
// the mma.sync kernel); 2 = the warp-specialised tcgen05 kernel (4 producer + 4 consumer warps; the default)
bool exact_tcgen05_supported(int dim, int k);
void launch_exact_knn_tf32(int metric, const GraphView& g, const float* d_queries, int nq, int k, int* d_ids, float* d_dists, int* d_fail,
                           void* d_scratch, size_t scratch_bytes, bool refresh_aux, int tcgen05_mode, cudaStream_t stream);
// per query: merge `parts` sorted key lists [nq][parts][k] (~0 padded) into ids / distances [nq][k] (-1 / NaN padded)
void launch_merge_keys(const u64* d_part_keys, int parts, int nq, int k, int* d_out_ids, float* d_out_dists, cudaStream_t stream);

// Remove (GraphConnector.cs:53-167) for one chunk of removed nodes, see remove_kernels.cu.
//   rm_index[id] = position of id in the chunk (use_index) or `value`
void launch_set_rm_index(int* d_rm_index, const int* d_ids, int n, bool use_index, int value, cudaStream_t stream);
//   per (removed node, layer) item: beam search from the node, width RemoveMaxCandidates + 1 -> d_cand [nitems][rmc] (-1 padded)
void launch_remove_search(int metric, const GraphView& g, const int* d_item_node, const int* d_item_layer, int nitems, int rmc, int* d_cand,
                          unsigned* d_work_counter, cudaStream_t stream);
//   every live row naming a node of the chunk -> d_affected (node, layer), count in *d_n_affected
void launch_find_affected(const GraphView& g, const int* d_levels, int length, int2* d_affected, unsigned* d_n_affected, cudaStream_t stream);
//   rewrite each affected row: remaining neighbours + search results of its removed neighbours, heuristic
void launch_relink(int metric, const GraphView& g, const int2* d_affected, int naffected, const int* d_item_first, const int* d_cand, int rmc,
                   unsigned* d_work_counter, cudaStream_t stream);
void launch_clear_rows(const GraphView& g, const int* d_item_node, const int* d_item_layer, int nitems, cudaStream_t stream);

// Partitioned index (partition_kernels.cu): search whose epilogue stores each query's top-k (global ids) into slot
// `rank` of every peer's gather buffer [world][nq][k] (peer pointers from CUDA IPC), then raises flags[peer][rank].
constexpr int kMaxPeers = 16;
struct ScatterArgs {
    int world, rank, id_offset;
    unsigned epoch;
    int* ids[kMaxPeers];
    float* dists[kMaxPeers];
    unsigned* flags[kMaxPeers];
};
void launch_knn_scatter(int metric, const GraphView& g, const float* d_queries, int nq, int k, int ef, const ScatterArgs& sc,
                        unsigned* d_work_counter, unsigned* d_done_counter, const Tuning& t, cudaStream_t stream);
// wait (bounded by timeout_ms of wall time) until d_flags[0..world) have all reached `epoch`, then merge the gather buffer
// in_* [world][nq][k] into out_* [nq][k] -- one kernel.  On timeout (or when *d_error is already set) *d_error = 1 and
// every output row is -1 / NaN.
void launch_wait_merge(const unsigned* d_flags, int world, unsigned epoch, int timeout_ms, int* d_error, const int* d_in_ids,
                       const float* d_in_dists, int nq, int k, int* d_out_ids, float* d_out_dists, cudaStream_t stream);

// K7: merge `parts` sorted top-k lists per query: in_* are [parts][nq][k], out_* [nq][k]
void launch_merge_topk(const int* d_in_ids, const float* d_in_dists, int parts, int nq, int k, int* d_out_ids, float* d_out_dists,
                       cudaStream_t stream);

}  // namespace hnsw
