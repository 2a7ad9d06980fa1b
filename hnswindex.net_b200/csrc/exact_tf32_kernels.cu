// exact_tf32_kernels.cu -- K3b: exact top-k through a 3xTF32 tensor-core GEMM (the optional variant of
// BASELINE.json's "exact top-k for recall checks").  No counterpart in the reference.
//
// The brute-force scan is a GEMM (queries x stored vectors x dim), so the dot products go through the
// tensor cores: every fp32 operand is split into hi = tf32(x) and lo = tf32(x - hi) and the product is
// accumulated as  lo*hi + hi*lo + hi*hi  (three TF32 MMAs with fp32 accumulation, the lo*lo term is below
// fp32 resolution).  That reproduces fp32 dot products to ~1e-6, NOT to the bit, and the reference's metric
// arithmetic is a fixed fp32 chain (metric.cuh), so the tensor-core pass is only a FILTER:
//
//   A  exact_tf32_scan_kernel   approximate distances of all pairs, per query the kp = k + slack best
//                               (per data chunk; merged by merge_keys_kernel);
//   B  exact_tf32_rerank_kernel the kp candidates are evaluated again with the bit-exact metric, sorted,
//                               cut to k -- and CERTIFIED: if the k-th exact distance is below the kp-th
//                               approximate distance minus the worst-case error of pass A, no vector outside
//                               the candidates can belong to the top k.  Queries that cannot be certified
//                               (near-ties denser than the slack) are flagged and re-run by the FFMA scan.
//
// The answers are therefore identical (ids and distance bits) to exact_scan_kernel's.
//
// Pass A, per block: TQ = 16 * WM queries (raw fp32 in shared memory) against tiles of 64 stored vectors
// streamed through a cp.async double buffer in K-chunks of 128 floats; 4 warps as WM x WN, each owning a
// 16 x (64 / WN) slab of the tile as m16n8k8 accumulators in registers.  Scores below the query's current
// kp-th best go to a block-wide queue in shared memory and are folded into per-query sorted lists between
// tiles (rare once the lists have warmed up).
#include <math_constants.h>

#include <algorithm>

#include "beam.cuh"
#include "kernels.hpp"

namespace hnsw {

namespace {

constexpr int kTileN = 64;    // stored vectors per tile
constexpr int kChunkK = 128;  // floats of every vector per pipeline stage
constexpr int kPitchB = kChunkK + 4;  // shared-memory row pitch of a stage: = 4 (mod 8) -> conflict-free fragment loads

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// x = hi + lo with hi = tf32(x) (11 significant bits) and lo = tf32(x - hi); x - hi is exact in fp32
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(__fsub_rn(x, __uint_as_float(hi)));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// worst-case error of the 3xTF32 dot product relative to |q| |x|: operand splitting (<= 1.25 * 2^-20) plus one
// truncating fp32 accumulation per MMA of the chain (3 * dim / 8 MMAs, <= 2^-22 each), with a factor 1.5 of slack
__host__ __device__ inline float tf32_dot_error(int dim) { return (0.5625f * (float)dim + 12.0f) * 2.384185791015625e-07f; }

}  // namespace

// Per stored vector: what pass A needs besides the dot product -- |x|^2 (sq_euclid / euclid), 1 / |x| (cosine; 0 for a
// null vector), 0 (ucosine) -- NaN for removed items (they are no candidates), and the largest |x|^2 of the index.
template <int METRIC>
__global__ void __launch_bounds__(kThreads) exact_tf32_aux_kernel(GraphView g, float* __restrict__ xaux, unsigned* __restrict__ xmax2_bits) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float mx = 0.f;
    for (int id = blockIdx.x * kWarpsPerBlock + warp; id < g.count; id += gridDim.x * kWarpsPerBlock) {
        const float* v = g.vectors + (size_t)id * g.pitch;
        float s = 0.f;
        for (int i = lane; i < g.dim; i += 32) { const float a = __ldg(v + i); s = fmaf(a, a, s); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        const bool live = __ldg(g.rm_index + id) == -1;
        if (live) mx = fmaxf(mx, s);
        if (lane == 0) {
            float a;
            if (!live) a = CUDART_NAN_F;
            else if (metric_is_sq(METRIC)) a = s;
            else if (METRIC == COSINE) a = s > 1e-36f ? rsqrtf(s) : 0.f;
            else a = 0.f;
            xaux[id] = a;
        }
    }
    if (lane == 0 && mx > 0.f) atomicMax(xmax2_bits, __float_as_uint(mx));  // non-negative floats order like their bits
}

// Pass A.  part_keys [nq][chunks][kp]: per (query, data chunk) the kp best (approximate distance | id) keys, ascending,
// ~0 padded.
template <int METRIC, int WM>
__global__ void __launch_bounds__(kThreads, 1)
exact_tf32_scan_kernel(GraphView g, const float* __restrict__ queries, int nq, int kp, int chunk_items, const float* __restrict__ xaux,
                       u64* __restrict__ part_keys) {
    constexpr int TQ = 16 * WM;             // queries per block
    constexpr int WN = kWarpsPerBlock / WM;  // warps side by side along the tile
    constexpr int NTW = (kTileN / 8) / WN;   // m16n8 accumulator tiles per warp
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gq = lane >> 2, tq = lane & 3;  // fragment coordinates ("groupID", "threadID_in_group")
    const int wm = warp / WN, wn = warp % WN;
    const int chunk = blockIdx.x, chunks = gridDim.x;
    const int q0 = blockIdx.y * TQ;
    const int dim8 = (g.dim + 7) & ~7;
    const int pitchA = dim8 + 4;

    float* As = reinterpret_cast<float*>(smem);                       // [TQ][pitchA] queries, zero beyond dim
    float* Bs = As + (size_t)TQ * pitchA;                             // [2][kTileN][kPitchB]
    u64* lists = reinterpret_cast<u64*>(Bs + 2 * kTileN * kPitchB);   // [TQ][kp]
    u64* qkeys = lists + (size_t)TQ * kp;                             // [TQ * kTileN] queue of (key) ...
    float* qa = reinterpret_cast<float*>(qkeys + TQ * kTileN);        // [TQ] |q|^2 or 1/|q|
    uint32_t* thr = reinterpret_cast<uint32_t*>(qa + TQ);             // [TQ] ordered kp-th best so far (0xFFFFFFFF: list not full)
    int* ln = reinterpret_cast<int*>(thr + TQ);                       // [TQ] list lengths
    int* qcount = ln + TQ;                                            // [1]
    unsigned char* qrows = reinterpret_cast<unsigned char*>(qcount + 1);  // [TQ * kTileN] ... and the query row of every queued key

    for (int i = tid; i < TQ * pitchA; i += kThreads) {
        const int r = i / pitchA, d = i - r * pitchA;
        As[i] = (q0 + r < nq && d < g.dim) ? __ldg(queries + (size_t)(q0 + r) * g.dim + d) : 0.f;
    }
    for (int i = tid; i < 2 * kTileN * kPitchB; i += kThreads) Bs[i] = 0.f;
    for (int i = tid; i < TQ; i += kThreads) { thr[i] = 0xFFFFFFFFu; ln[i] = 0; }
    if (tid == 0) *qcount = 0;
    __syncthreads();
    for (int r = warp; r < TQ; r += kWarpsPerBlock) {
        float s = 0.f;
        for (int i = lane; i < g.dim; i += 32) { const float a = As[(size_t)r * pitchA + i]; s = fmaf(a, a, s); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if (lane == 0) qa[r] = metric_is_sq(METRIC) ? s : (METRIC == COSINE ? (s > 1e-36f ? rsqrtf(s) : 0.f) : 0.f);
    }

    const int begin = chunk * chunk_items, end = min(g.count, begin + chunk_items);
    const int ntiles = begin < end ? (end - begin + kTileN - 1) / kTileN : 0;
    const int nkc = (dim8 + kChunkK - 1) / kChunkK;
    const int nitems = ntiles * nkc;

    // stage loader: item = (tile, K-chunk); rows past `end` re-read the chunk's last vector and are masked later
    auto load_item = [&](int item, int stage) {
        const int tile = item / nkc, kc = (item - tile * nkc) * kChunkK;
        const int width = min(kChunkK, g.pitch - kc);  // floats present in the slab rows for this chunk (multiple of 4)
        const int upr = width >> 2;
        float* dst = Bs + (size_t)stage * kTileN * kPitchB;
        const int base = begin + tile * kTileN;
        for (int u = tid; u < kTileN * upr; u += kThreads) {
            const int r = u / upr, c4 = u - r * upr;
            const int id = min(base + r, end - 1);
            cp_async16(dst + (size_t)r * kPitchB + c4 * 4, g.vectors + (size_t)id * g.pitch + kc + c4 * 4);
        }
        if (width & 7) {  // the last 8-wide step reads 4 floats past the row: keep them zero
            for (int r = tid; r < kTileN; r += kThreads) *reinterpret_cast<float4*>(dst + (size_t)r * kPitchB + width) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        cp_async_commit();
    };

    float acc[NTW][4];
    if (nitems > 0) load_item(0, 0);
    for (int it = 0; it < nitems; ++it) {
        const int stage = it & 1;
        if (it + 1 < nitems) { load_item(it + 1, stage ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const int tile = it / nkc, kci = it - tile * nkc, kc = kci * kChunkK;
        if (kci == 0) {
#pragma unroll
            for (int j = 0; j < NTW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
        }
        const int ksteps = (min(kChunkK, dim8 - kc) + 7) >> 3;
        const float* a_lo_row = As + (size_t)(wm * 16 + gq) * pitchA + kc + tq;
        const float* a_hi_row = a_lo_row + 8 * pitchA;
        const float* b_base = Bs + (size_t)stage * kTileN * kPitchB + (size_t)(wn * NTW * 8 + gq) * kPitchB + tq;
        for (int ks = 0; ks < ksteps; ++ks) {
            uint32_t ah[4], al[4];
            split_tf32(a_lo_row[ks * 8], ah[0], al[0]);
            split_tf32(a_hi_row[ks * 8], ah[1], al[1]);
            split_tf32(a_lo_row[ks * 8 + 4], ah[2], al[2]);
            split_tf32(a_hi_row[ks * 8 + 4], ah[3], al[3]);
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const float* bp = b_base + (size_t)j * 8 * kPitchB + ks * 8;
                uint32_t bh[2], bl[2];
                split_tf32(bp[0], bh[0], bl[0]);
                split_tf32(bp[4], bh[1], bl[1]);
                mma_tf32(acc[j], al, bh);  // small terms first
                mma_tf32(acc[j], ah, bl);
                mma_tf32(acc[j], ah, bh);
            }
        }
        if (kci == nkc - 1) {
            // scores of this tile -> queue (only what can still enter the query's list)
            const int base = begin + tile * kTileN;
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const int col = wn * NTW * 8 + j * 8 + 2 * tq;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = wm * 16 + gq + (e >> 1) * 8;
                    const int id = base + col + (e & 1);
                    if (id >= end || q0 + row >= nq) continue;
                    const float xa = __ldg(xaux + id);
                    if (xa != xa) continue;  // removed item
                    const float c = acc[j][e];
                    float s;
                    if (metric_is_sq(METRIC)) s = (qa[row] + xa) - 2.0f * c;
                    else if (METRIC == COSINE) s = 1.0f - c * (qa[row] * xa);
                    else s = 1.0f - c;
                    if (f2ord(s) <= thr[row]) {
                        const int slot = atomicAdd(qcount, 1);
                        qkeys[slot] = make_key(s, id);
                        qrows[slot] = (unsigned char)row;
                    }
                }
            }
            __syncthreads();
            const int nqueued = *qcount;
            __syncthreads();
            if (tid == 0) *qcount = 0;
            // fold the queue into the lists: query row r belongs to warp r % 4, so every list has one writer
            for (int b0 = 0; b0 < nqueued; b0 += 32) {
                const int e = b0 + lane;
                u64 key = ~0ull;
                int row = 0;
                bool pending = false;
                if (e < nqueued) { key = qkeys[e]; row = qrows[e]; pending = (row & (kWarpsPerBlock - 1)) == warp; }
                for (;;) {
                    bool want = false;
                    if (pending) {
                        want = ln[row] < kp || key < lists[(size_t)row * kp + kp - 1];
                        pending = want;
                    }
                    const unsigned mk = __ballot_sync(FULL, want);
                    if (!mk) break;
                    const int src = __ffs(mk) - 1;
                    const u64 kn = __shfl_sync(FULL, key, src);
                    const int rw = __shfl_sync(FULL, row, src);
                    if (lane == src) pending = false;
                    u64* list = lists + (size_t)rw * kp;
                    BeamState st;
                    st.n = ln[rw];
                    st.ef = kp;
                    st.cursor = 0;
                    st.far_ord = st.n == kp ? key_ord(list[kp - 1]) : 0xFFFFFFFFu;
                    __syncwarp();
                    list_insert(list, st, kn, lane);
                    __syncwarp();
                    if (lane == 0) { ln[rw] = st.n; thr[rw] = st.far_ord; }
                    __syncwarp();
                }
            }
        }
        __syncthreads();  // stage fully read (and the lists settled) before the next iteration overwrites it
    }

    for (int r = warp; r < TQ; r += kWarpsPerBlock) {
        if (q0 + r >= nq) continue;
        const u64* list = lists + (size_t)r * kp;
        const int n = ln[r];
        u64* out = part_keys + ((size_t)(q0 + r) * chunks + chunk) * kp;
        for (int j = lane; j < kp; j += 32) out[j] = j < n ? list[j] : ~0ull;
    }
}

// Pass B, one warp per query: bit-exact distances of the kp candidates (lane pairs, 16 candidates per pass), top k,
// certificate.  cand_ids / cand_approx [nq][kp] ascending by approximate distance, -1 / NaN padded.
template <int METRIC>
__global__ void __launch_bounds__(kThreads)
exact_tf32_rerank_kernel(GraphView g, const float* __restrict__ queries, int nq, int k, int kp, const int* __restrict__ cand_ids,
                         const float* __restrict__ cand_approx, const unsigned* __restrict__ xmax2_bits, int* __restrict__ out_ids,
                         float* __restrict__ out_dists, int* __restrict__ fail) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, h = lane & 1, grp = lane >> 1;
    const int q = blockIdx.x * kWarpsPerBlock + warp;
    if (q >= nq) return;
    const int pitch4 = (g.pitch + 3) & ~3;
    float* qs = reinterpret_cast<float*>(smem) + (size_t)warp * pitch4;
    u64* list = reinterpret_cast<u64*>(reinterpret_cast<float*>(smem) + (size_t)kWarpsPerBlock * pitch4) + (size_t)warp * k;
    for (int d = lane; d < pitch4; d += 32) qs[d] = d < g.dim ? __ldg(queries + (size_t)q * g.dim + d) : 0.f;
    __syncwarp();
    double qn_sqrt = 0.0;
    if (METRIC == COSINE) qn_sqrt = pair_norm_sqrt(qs, g.dim, h);
    float qn2 = 0.f;
    for (int i = lane; i < g.dim; i += 32) qn2 = fmaf(qs[i], qs[i], qn2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qn2 += __shfl_xor_sync(FULL, qn2, o);

    const int* cids = cand_ids + (size_t)q * kp;
    int ncand = 0;
    for (int b0 = 0; b0 < kp; b0 += 32) ncand += __popc(__ballot_sync(FULL, b0 + lane < kp && cids[b0 + lane] >= 0));
    BeamState st;
    st.n = 0; st.ef = k; st.cursor = 0; st.far_ord = 0xFFFFFFFFu;
    const int nch = g.dim >> 3;
    for (int b0 = 0; b0 < ncand; b0 += 16) {
        const int ci = b0 + grp;
        const bool active = ci < ncand;
        const int id = cids[active ? ci : 0];
        const float* vec = g.vectors + (size_t)id * g.pitch;
        PairAcc acc = {0.f, 0.f, 0.f, 0.f};
        float r;
        if (metric_is_sq(METRIC)) {
            pair_accumulate<true>(acc, vec, qs, nch, h);
            r = pair_hsum<true>(acc);
            r = pair_tail<true>(r, vec + (nch << 3), qs + (nch << 3), g.dim - (nch << 3));
        } else {
            pair_accumulate<false>(acc, vec, qs, nch, h);
            r = pair_hsum<false>(acc);
            r = pair_tail<false>(r, vec + (nch << 3), qs + (nch << 3), g.dim - (nch << 3));
        }
        double an = 0.0;
        if (METRIC == COSINE) an = __ldg(g.norm_sqrt + id);
        r = metric_finish<METRIC>(r, an, qn_sqrt);
        const u64 mykey = make_key(r, id);
        bool pending = active && h == 0;
        for (;;) {
            const bool want = pending && (st.n < k || mykey < list[k - 1]);
            const unsigned mk = __ballot_sync(FULL, want);
            if (!mk) break;
            const int src = __ffs(mk) - 1;
            const u64 kn = __shfl_sync(FULL, mykey, src);
            if (lane == src) pending = false;
            list_insert(list, st, kn, lane);
            __syncwarp();
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
    if (lane == 0) {
        bool ok = true;
        if (ncand >= kp) {  // otherwise the candidates are ALL live items
            const float xmax2 = __uint_as_float(*xmax2_bits);
            const float cd = tf32_dot_error(g.dim);
            const float chain = ((float)(g.dim >> 3) + 16.0f) * 2.384185791015625e-07f;  // fp32 chains of both sides
            float err;
            if (metric_is_sq(METRIC)) err = (cd + chain + (float)g.dim * 1.1920928955078125e-07f) * (qn2 + xmax2);
            else if (METRIC == COSINE) err = cd + chain;
            else err = (cd + chain) * fmaxf(1.0f, sqrtf(qn2 * xmax2));
            const float bound = cand_approx[(size_t)q * kp + kp - 1] - err;  // every vector outside the candidates is at least this far
            float dk = st.n == k ? key_dist(list[k - 1]) : CUDART_INF_F;  // kp > k candidates: the list is full
            if (METRIC == EUCLID) dk = dk * dk * (1.0f + 4.76837158203125e-07f);
            ok = dk < bound;
        }
        fail[q] = ok ? 0 : 1;
    }
}


// ---------------------------------------------------------------------------------------------
// Pass A on the 5th-generation tensor cores (tcgen05 + TMEM), opt-in (tuning key "exact_tcgen05") for rows of at most
// 128 floats.  One block = 128 queries = the 128 TMEM lanes of one accumulator D[128 x 64] (64 TMEM columns).  Both
// operands are split into tf32 hi / lo parts when they are written to shared memory, in the K-major no-swizzle
// canonical layout of a UMMA shared-memory descriptor: 8-row x 16-byte core matrices, 128 B each, consecutive along K
// (leading byte offset 128), 8-row groups Kt * 32 bytes apart (stride byte offset).  Per tile of 64 stored vectors
// one elected thread issues 3 x Kt / 8 `tcgen05.mma.cta_group::1.kind::tf32` (lo*hi, hi*lo, hi*hi per K-step of 8)
// and commits them to an mbarrier; every thread then owns one query row: `tcgen05.ld.32x32b` brings its 64 dot products
// to registers, and the few that beat the row's kp-th best are inserted into the row's sorted list by that thread
// alone.  Same part_keys output as exact_tf32_scan_kernel, so merge, re-ranking and certificate are shared.
// Two accumulators alternate, so the epilogue of tile t - 1 runs under the MMAs of tile t, and the rows of tile t + 1 are
// on their way from global memory into registers meanwhile; the operand stage itself is single-buffered (at 128 floats
// per row the two split operand tiles and the lists fill the 227 KB), which is what a TMA-fed multi-stage ring would lift.
namespace tc5 {

constexpr int kM = 128;  // queries per block (UMMA M, TMEM lanes)
constexpr int kN = 64;   // stored vectors per tile (UMMA N, TMEM columns)
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32 [4,6) = 1, A / B = tf32 [7,10) / [10,13) = 2, both K-major,
// N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), no swizzle: start address, leading / stride byte offsets
// (all >> 4), version 1 at [46,48)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) |
           ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && clock64() - t0 > 4000000000ll) __trap();  // ~2 s: never hang the device on a lost completion
    }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// byte offset of the 16-byte piece (row r, floats 4 * k4 .. 4 * k4 + 3) in the canonical K-major layout
__device__ __forceinline__ uint32_t piece_off(int r, int k4, int kt) { return (uint32_t)(r >> 3) * (uint32_t)(kt * 32) + (uint32_t)k4 * 128u + (uint32_t)(r & 7) * 16u; }
__device__ __forceinline__ void store_split(unsigned char* hi_base, unsigned char* lo_base, uint32_t off, float4 v) {
    uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
    split_tf32(v.x, h0, l0);
    split_tf32(v.y, h1, l1);
    split_tf32(v.z, h2, l2);
    split_tf32(v.w, h3, l3);
    *reinterpret_cast<uint4*>(hi_base + off) = make_uint4(h0, h1, h2, h3);
    *reinterpret_cast<uint4*>(lo_base + off) = make_uint4(l0, l1, l2, l3);
}

}  // namespace tc5

template <int METRIC>
__global__ void __launch_bounds__(tc5::kM, 1)
exact_tc5_scan_kernel(GraphView g, const float* __restrict__ queries, int nq, int kp, int chunk_items, const float* __restrict__ xaux,
                      u64* __restrict__ part_keys) {
    using namespace tc5;
    extern __shared__ __align__(16) unsigned char smem[];  // descriptors without swizzle need 16 B alignment only
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x, chunks = gridDim.x;
    const int q0 = blockIdx.y * kM;
    const int kt = (g.dim + 15) & ~15;  // staged K extent, zero padded
    const int pitchL = kp | 1;           // odd list pitch (in keys): rows of different threads spread over the banks

    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + (size_t)kM * kt * 4;
    unsigned char* b_hi = a_lo + (size_t)kM * kt * 4;
    unsigned char* b_lo = b_hi + (size_t)kN * kt * 4;
    u64* lists = reinterpret_cast<u64*>(b_lo + (size_t)kN * kt * 4);  // [kM][pitchL]
    float* xa_s = reinterpret_cast<float*>(lists + (size_t)kM * pitchL);  // [3][kN] per-item norms of the tiles in flight
    u64* bar = reinterpret_cast<u64*>(xa_s + 3 * kN);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(bar);

    // one warp tile = 8 rows x 4 pieces: coalesced 64 B runs in global memory, one conflict-free 512 B run in shared memory
    const int rr = lane & 7, kk = lane >> 3;
    const int kgroups = kt >> 4;
    for (int wt = warp; wt < (kM / 8) * kgroups; wt += kM / 32) {
        const int r = (wt / kgroups) * 8 + rr, k4 = (wt % kgroups) * 4 + kk;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < nq) {
            const float* src = queries + (size_t)(q0 + r) * g.dim;
            const int d = k4 * 4;
            if (d < g.dim) v.x = __ldg(src + d);
            if (d + 1 < g.dim) v.y = __ldg(src + d + 1);
            if (d + 2 < g.dim) v.z = __ldg(src + d + 2);
            if (d + 3 < g.dim) v.w = __ldg(src + d + 3);
        }
        store_split(a_hi, a_lo, piece_off(r, k4, kt), v);
    }
    float qa = 0.f;  // |q|^2 or 1/|q| of this thread's query row
    if (q0 + tid < nq) {
        const float* src = queries + (size_t)(q0 + tid) * g.dim;
        float s = 0.f;
        for (int d = 0; d < g.dim; ++d) { const float a = __ldg(src + d); s = fmaf(a, a, s); }
        qa = metric_is_sq(METRIC) ? s : (METRIC == COSINE ? (s > 1e-36f ? rsqrtf(s) : 0.f) : 0.f);
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "r"(2 * kN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    // This thread's query row: the kp best keys so far as a binary MAX-HEAP in shared memory (an insertion replaces the
    // root and sifts down at most log2(kp) levels; a sorted insertion is a chain of up to kp dependent shared-memory round
    // trips and a rescan for the maximum costs kp loads per insertion -- the ncu source pages showed either to dominate
    // the kernel); sorted once at the end.
    u64* list = lists + (size_t)tid * pitchL;
    int n = 0;
    u64 maxkey = 0;
    float thr_f = CUDART_INF_F;  // distance of the kp-th best, +inf while the list is not full
    const int begin = chunk * chunk_items, end = min(g.count, begin + chunk_items);
    const int ntiles = begin < end ? (end - begin + kN - 1) / kN : 0;
    const uint32_t sbo = (uint32_t)kt * 32u;
    const uint64_t dah0 = smem_desc((uint32_t)__cvta_generic_to_shared(a_hi), 128u, sbo), dal0 = smem_desc((uint32_t)__cvta_generic_to_shared(a_lo), 128u, sbo);
    const uint64_t dbh0 = smem_desc((uint32_t)__cvta_generic_to_shared(b_hi), 128u, sbo), dbl0 = smem_desc((uint32_t)__cvta_generic_to_shared(b_lo), 128u, sbo);
    const int b_tiles = (kN / 8) * kgroups;  // warp tiles of one stage: at most 64, i.e. 16 per warp

    // per-thread pieces of a stage, fixed for the whole kernel: byte offset in shared memory (-1: none) and float offset
    // from the tile's first row in the slab (-1: beyond the row pitch, stays zero)
    int poff[16], goff[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int wt = warp + i * (kM / 32);
        poff[i] = goff[i] = -1;
        if (wt < b_tiles) {
            const int r = (wt / kgroups) * 8 + rr, k4 = (wt % kgroups) * 4 + kk;
            poff[i] = (int)piece_off(r, k4, kt);
            if (k4 * 4 < g.pitch) goff[i] = r * g.pitch + k4 * 4;
        }
    }

    // Software pipeline over the tiles, iteration t:
    //   stage(t): the rows prefetched into registers one iteration earlier are split and written to shared memory
    //   MMA(t) is issued into accumulator t & 1 (asynchronous), the rows of tile t + 1 are requested from global memory,
    //   epilogue(t - 1) drains accumulator (t - 1) & 1 while MMA(t) runs, then everybody waits for MMA(t).
    float4 pre[16];
    float xa_pre = CUDART_NAN_F;
#define HNSW_TC5_PREFETCH(TILE)                                                                                              \
    {                                                                                                                        \
        const int pbase = begin + (TILE) * kN;                                                                               \
        const float* tb = g.vectors + (size_t)pbase * g.pitch;                                                               \
        if (pbase + kN <= end) {                                                                                             \
            _Pragma("unroll") for (int i = 0; i < 16; ++i) {                                                                 \
                pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);                                                                    \
                if (goff[i] >= 0) pre[i] = ldg_f4(reinterpret_cast<const float4*>(tb + goff[i]));                            \
            }                                                                                                                \
        } else { /* last, partial tile: rows past the end re-read the last vector and are masked by their NaN norm */       \
            _Pragma("unroll") for (int i = 0; i < 16; ++i) {                                                                 \
                const int wt = warp + i * (kM / 32);                                                                         \
                pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);                                                                    \
                if (goff[i] >= 0) {                                                                                          \
                    const int r = (wt / kgroups) * 8 + rr, k4 = (wt % kgroups) * 4 + kk;                                     \
                    const int id = min(pbase + r, end - 1);                                                                  \
                    pre[i] = ldg_f4(reinterpret_cast<const float4*>(g.vectors + (size_t)id * g.pitch) + k4);                 \
                }                                                                                                            \
            }                                                                                                                \
        }                                                                                                                    \
        if (tid < kN) xa_pre = pbase + tid < end ? __ldg(xaux + pbase + tid) : CUDART_NAN_F;                                 \
    }
    if (ntiles > 0) HNSW_TC5_PREFETCH(0)

    for (int tile = 0; tile <= ntiles; ++tile) {
        if (tile < ntiles) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (poff[i] >= 0) store_split(b_hi, b_lo, (uint32_t)poff[i], pre[i]);
            if (tid < kN) xa_s[(tile % 3) * kN + tid] = xa_pre;  // three slots: read one iteration later, rewritten three later
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the tensor core reads shared memory through the async proxy
        }
        tc_fence_before();  // (this thread's tcgen05.ld of the previous iteration are complete)
        __syncthreads();
        if (tile < ntiles) {
            if (tid == 0) {
                tc_fence_after();
                const uint32_t d = tmem_d + (uint32_t)(tile & 1) * kN;
                for (int ks = 0; ks < (kt >> 3); ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 16);  // 8 floats along K = two core matrices = 256 B, in 16 B units
                    umma_tf32(d, dal0 + adv, dbh0 + adv, ks > 0 ? 1u : 0u);  // small terms first
                    umma_tf32(d, dah0 + adv, dbl0 + adv, 1u);
                    umma_tf32(d, dah0 + adv, dbh0 + adv, 1u);
                }
                umma_commit(bar_addr);
            }
            if (tile + 1 < ntiles) HNSW_TC5_PREFETCH(tile + 1)
        }
        __syncwarp();  // tcgen05.ld is warp-collective
        if (tile > 0) {
            const int base = begin + (tile - 1) * kN;
            const float* xa_t = xa_s + ((tile - 1) % 3) * kN;  // NaN for removed items and past the chunk: never below a threshold
            const uint32_t d = tmem_d + (uint32_t)((tile - 1) & 1) * kN + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < kN; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(d + (uint32_t)c0, v);
                if (q0 + tid >= nq) continue;
                float xa[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 x4 = *reinterpret_cast<const float4*>(xa_t + c0 + j);
                    xa[j] = x4.x; xa[j + 1] = x4.y; xa[j + 2] = x4.z; xa[j + 3] = x4.w;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float c = __uint_as_float(v[j]);
                    float s;
                    if (metric_is_sq(METRIC)) s = fmaf(-2.0f, c, qa + xa[j]);
                    else if (METRIC == COSINE) s = fmaf(-c, qa * xa[j], 1.0f);
                    else s = (xa[j] + 1.0f) - c;  // xa is 0 for a live item
                    if (!(s <= thr_f)) continue;
                    const u64 key = make_key(s, base + c0 + j);
                    if (n < kp) {
                        list[n] = key;
                        if (++n == kp) {  // full: turn the array into a max-heap (largest key at list[0])
                            for (int h0 = kp / 2 - 1; h0 >= 0; --h0) {
                                const u64 val = list[h0];
                                int i = h0;
                                for (;;) {
                                    int ch = 2 * i + 1;
                                    if (ch >= kp) break;
                                    u64 cv = list[ch];
                                    if (ch + 1 < kp) { const u64 rv = list[ch + 1]; if (rv > cv) { cv = rv; ++ch; } }
                                    if (cv <= val) break;
                                    list[i] = cv;
                                    i = ch;
                                }
                                list[i] = val;
                            }
                            maxkey = list[0];
                            thr_f = key_dist(maxkey);
                        }
                    } else if (key < maxkey) {  // replace the root and sift the new key down: at most log2(kp) levels
                        int i = 0;
                        for (;;) {
                            int ch = 2 * i + 1;
                            if (ch >= kp) break;
                            u64 cv = list[ch];
                            if (ch + 1 < kp) { const u64 rv = list[ch + 1]; if (rv > cv) { cv = rv; ++ch; } }
                            if (cv <= key) break;
                            list[i] = cv;
                            i = ch;
                        }
                        list[i] = key;
                        maxkey = list[0];
                        thr_f = key_dist(maxkey);
                    }
                }
            }
        }
        if (tile < ntiles) {
            mbar_wait(bar_addr, (uint32_t)(tile & 1));
            __syncwarp();
            tc_fence_after();
        }
    }
#undef HNSW_TC5_PREFETCH
    if (q0 + tid < nq) {
        for (int a = 1; a < n; ++a) {  // sort this row's keys ascending (once)
            const u64 key = list[a];
            int i = a;
            while (i > 0 && list[i - 1] > key) { list[i] = list[i - 1]; --i; }
            list[i] = key;
        }
        u64* out = part_keys + ((size_t)(q0 + tid) * chunks + chunk) * kp;
        for (int j = 0; j < kp; ++j) out[j] = j < n ? list[j] : ~0ull;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) __syncwarp();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(2 * kN));
}

// ---------------------------------------------------------------------------------------------
// EXPERIMENTAL (tuning key "exact_tcgen05" = 2; written at the end of round 1 without GPU time left to run it): the
// same pass A, warp-specialised.  256 threads: warps 0-3 only drain the accumulators (thread t = query row t, as above),
// warps 4-7 only stage the operand tile (prefetch to registers, split, store) and one of their threads issues the MMAs.
//   mma_done[b]  (count 1, tcgen05.commit)   accumulator b holds tile t (t & 1 == b) and the operand stage is free again
//   tmem_free[b] (count 128, consumer threads) accumulator b has been read and may be overwritten by tile t + 2
// The producers of tile t + 1 wait for mma_done(t), the consumers of tile t wait for mma_done(t): the epilogue of tile t
// runs while tile t + 1 is staged and multiplied.
template <int METRIC>
__global__ void __launch_bounds__(2 * tc5::kM, 1)
exact_tc5ws_scan_kernel(GraphView g, const float* __restrict__ queries, int nq, int kp, int chunk_items, const float* __restrict__ xaux,
                        u64* __restrict__ part_keys) {
    using namespace tc5;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.x, chunks = gridDim.x;
    const int q0 = blockIdx.y * kM;
    const int kt = (g.dim + 15) & ~15;
    const int pitchL = kp | 1;

    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + (size_t)kM * kt * 4;
    unsigned char* b_hi = a_lo + (size_t)kM * kt * 4;
    unsigned char* b_lo = b_hi + (size_t)kN * kt * 4;
    u64* lists = reinterpret_cast<u64*>(b_lo + (size_t)kN * kt * 4);
    float* xa_s = reinterpret_cast<float*>(lists + (size_t)kM * pitchL);  // [3][kN]
    u64* bars = reinterpret_cast<u64*>(xa_s + 3 * kN);                     // mma_done[2], tmem_free[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    const uint32_t mma_done0 = bar0, tmem_free0 = bar0 + 16;

    const int rr = lane & 7, kk = lane >> 3;
    const int kgroups = kt >> 4;
    // queries -> split A tile (all 8 warps)
    for (int wt = warp; wt < (kM / 8) * kgroups; wt += 2 * kM / 32) {
        const int r = (wt / kgroups) * 8 + rr, k4 = (wt % kgroups) * 4 + kk;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < nq) {
            const float* src = queries + (size_t)(q0 + r) * g.dim;
            const int d = k4 * 4;
            if (d < g.dim) v.x = __ldg(src + d);
            if (d + 1 < g.dim) v.y = __ldg(src + d + 1);
            if (d + 2 < g.dim) v.z = __ldg(src + d + 2);
            if (d + 3 < g.dim) v.w = __ldg(src + d + 3);
        }
        store_split(a_hi, a_lo, piece_off(r, k4, kt), v);
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(tmem_slot)), "r"(2 * kN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mma_done0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mma_done0 + 8));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tmem_free0), "r"(kM));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tmem_free0 + 8), "r"(kM));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the A tile, for the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_slot;

    const int begin = chunk * chunk_items, end = min(g.count, begin + chunk_items);
    const int ntiles = begin < end ? (end - begin + kN - 1) / kN : 0;

    if (warp >= kM / 32) {
        // ===== producers: warps 4-7 =====
        const int pw = warp - kM / 32, ptid = tid - kM;
        const uint32_t sbo = (uint32_t)kt * 32u;
        const uint64_t dah0 = smem_desc((uint32_t)__cvta_generic_to_shared(a_hi), 128u, sbo), dal0 = smem_desc((uint32_t)__cvta_generic_to_shared(a_lo), 128u, sbo);
        const uint64_t dbh0 = smem_desc((uint32_t)__cvta_generic_to_shared(b_hi), 128u, sbo), dbl0 = smem_desc((uint32_t)__cvta_generic_to_shared(b_lo), 128u, sbo);
        const int b_tiles = (kN / 8) * kgroups;
        int poff[16], goff[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int wt = pw + i * (kM / 32);
            poff[i] = goff[i] = -1;
            if (wt < b_tiles) {
                const int r = (wt / kgroups) * 8 + rr, k4 = (wt % kgroups) * 4 + kk;
                poff[i] = (int)piece_off(r, k4, kt);
                if (k4 * 4 < g.pitch) goff[i] = r * g.pitch + k4 * 4;
            }
        }
        float4 pre[16];
        float xa_pre = CUDART_NAN_F;
#define HNSW_TC5WS_PREFETCH(TILE)                                                                                            \
    {                                                                                                                        \
        const int pbase = begin + (TILE) * kN;                                                                               \
        const float* tb = g.vectors + (size_t)pbase * g.pitch;                                                               \
        if (pbase + kN <= end) {                                                                                             \
            _Pragma("unroll") for (int i = 0; i < 16; ++i) {                                                                 \
                pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);                                                                    \
                if (goff[i] >= 0) pre[i] = ldg_f4(reinterpret_cast<const float4*>(tb + goff[i]));                            \
            }                                                                                                                \
        } else {                                                                                                             \
            _Pragma("unroll") for (int i = 0; i < 16; ++i) {                                                                 \
                const int wt = pw + i * (kM / 32);                                                                           \
                pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);                                                                    \
                if (goff[i] >= 0) {                                                                                          \
                    const int r = (wt / kgroups) * 8 + rr, k4 = (wt % kgroups) * 4 + kk;                                     \
                    const int id = min(pbase + r, end - 1);                                                                  \
                    pre[i] = ldg_f4(reinterpret_cast<const float4*>(g.vectors + (size_t)id * g.pitch) + k4);                 \
                }                                                                                                            \
            }                                                                                                                \
        }                                                                                                                    \
        if (ptid < kN) xa_pre = pbase + ptid < end ? __ldg(xaux + pbase + ptid) : CUDART_NAN_F;                              \
    }
        if (ntiles > 0) HNSW_TC5WS_PREFETCH(0)
        for (int tile = 0; tile < ntiles; ++tile) {
            if (tile > 0) mbar_wait(mma_done0 + 8u * (uint32_t)((tile - 1) & 1), (uint32_t)(((tile - 1) >> 1) & 1));  // stage free
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (poff[i] >= 0) store_split(b_hi, b_lo, (uint32_t)poff[i], pre[i]);
            if (ptid < kN) xa_s[(tile % 3) * kN + ptid] = xa_pre;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the four producer warps
            if (ptid == 0) {
                if (tile >= 2) mbar_wait(tmem_free0 + 8u * (uint32_t)(tile & 1), (uint32_t)(((tile - 2) >> 1) & 1));  // accumulator drained
                tc_fence_after();
                const uint32_t d = tmem_d + (uint32_t)(tile & 1) * kN;
                for (int ks = 0; ks < (kt >> 3); ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 16);
                    umma_tf32(d, dal0 + adv, dbh0 + adv, ks > 0 ? 1u : 0u);
                    umma_tf32(d, dah0 + adv, dbl0 + adv, 1u);
                    umma_tf32(d, dah0 + adv, dbh0 + adv, 1u);
                }
                umma_commit(mma_done0 + 8u * (uint32_t)(tile & 1));
            }
            if (tile + 1 < ntiles) HNSW_TC5WS_PREFETCH(tile + 1)
        }
#undef HNSW_TC5WS_PREFETCH
    } else {
        // ===== consumers: warps 0-3, thread t owns query row t =====
        float qa = 0.f;
        if (q0 + tid < nq) {
            const float* src = queries + (size_t)(q0 + tid) * g.dim;
            float s = 0.f;
            for (int d = 0; d < g.dim; ++d) { const float a = __ldg(src + d); s = fmaf(a, a, s); }
            qa = metric_is_sq(METRIC) ? s : (METRIC == COSINE ? (s > 1e-36f ? rsqrtf(s) : 0.f) : 0.f);
        }
        u64* list = lists + (size_t)tid * pitchL;
        int n = 0;
        u64 maxkey = 0;
        float thr_f = CUDART_INF_F;
        for (int tile = 0; tile < ntiles; ++tile) {
            mbar_wait(mma_done0 + 8u * (uint32_t)(tile & 1), (uint32_t)((tile >> 1) & 1));
            __syncwarp();
            tc_fence_after();
            const int base = begin + tile * kN;
            const float* xa_t = xa_s + (tile % 3) * kN;
            const uint32_t d = tmem_d + (uint32_t)(tile & 1) * kN + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
            for (int c0 = 0; c0 < kN; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(d + (uint32_t)c0, v);
                if (q0 + tid >= nq) continue;
                float xa[16];
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 x4 = *reinterpret_cast<const float4*>(xa_t + c0 + j);
                    xa[j] = x4.x; xa[j + 1] = x4.y; xa[j + 2] = x4.z; xa[j + 3] = x4.w;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float c = __uint_as_float(v[j]);
                    float s;
                    if (metric_is_sq(METRIC)) s = fmaf(-2.0f, c, qa + xa[j]);
                    else if (METRIC == COSINE) s = fmaf(-c, qa * xa[j], 1.0f);
                    else s = (xa[j] + 1.0f) - c;
                    if (!(s <= thr_f)) continue;
                    const u64 key = make_key(s, base + c0 + j);
                    if (n < kp) {
                        list[n] = key;
                        if (++n == kp) {
                            for (int h0 = kp / 2 - 1; h0 >= 0; --h0) {
                                const u64 val = list[h0];
                                int i = h0;
                                for (;;) {
                                    int ch = 2 * i + 1;
                                    if (ch >= kp) break;
                                    u64 cv = list[ch];
                                    if (ch + 1 < kp) { const u64 rv = list[ch + 1]; if (rv > cv) { cv = rv; ++ch; } }
                                    if (cv <= val) break;
                                    list[i] = cv;
                                    i = ch;
                                }
                                list[i] = val;
                            }
                            maxkey = list[0];
                            thr_f = key_dist(maxkey);
                        }
                    } else if (key < maxkey) {
                        int i = 0;
                        for (;;) {
                            int ch = 2 * i + 1;
                            if (ch >= kp) break;
                            u64 cv = list[ch];
                            if (ch + 1 < kp) { const u64 rv = list[ch + 1]; if (rv > cv) { cv = rv; ++ch; } }
                            if (cv <= key) break;
                            list[i] = cv;
                            i = ch;
                        }
                        list[i] = key;
                        maxkey = list[0];
                        thr_f = key_dist(maxkey);
                    }
                }
            }
            __syncwarp();
            tc_fence_before();  // this thread's tcgen05.ld of the accumulator are complete ...
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" ::"r"(tmem_free0 + 8u * (uint32_t)(tile & 1)) : "memory");  // ... it may be overwritten
        }
        if (q0 + tid < nq) {
            for (int a = 1; a < n; ++a) {
                const u64 key = list[a];
                int i = a;
                while (i > 0 && list[i - 1] > key) { list[i] = list[i - 1]; --i; }
                list[i] = key;
            }
            u64* out = part_keys + ((size_t)(q0 + tid) * chunks + chunk) * kp;
            for (int j = 0; j < kp; ++j) out[j] = j < n ? list[j] : ~0ull;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) __syncwarp();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(2 * kN));
}

static size_t tc5_scan_smem(int dim, int kp) {
    const int kt = (dim + 15) & ~15;
    return (size_t)(tc5::kM + tc5::kN) * kt * 8 + (size_t)tc5::kM * (kp | 1) * 8 + 3 * tc5::kN * 4 + 48;  // + barriers, TMEM slot
}

static int tf32_chunks(int count) { return std::max(1, std::min(32, (count + 4095) / 4096)); }

// Data chunks for the tcgen05 kernel: every chunk starts its lists from scratch (k' log(N / chunks / k') insertions per
// query and chunk), so as few as fill the device evenly: the candidate with the best wave efficiency, smaller first.
static int tc5_chunks(int count, int nq, int sms) {
    const int cmax = tf32_chunks(count), qt = (nq + tc5::kM - 1) / tc5::kM;
    int best = 1;
    double best_score = -1.0;
    for (int c = 1; c <= cmax; ++c) {
        const long long ctas = (long long)c * qt, rounds = (ctas + sms - 1) / sms;
        const double score = (double)ctas / (double)(rounds * sms) - 0.004 * c;
        if (score > best_score) { best_score = score; best = c; }
    }
    return best;
}

// the tcgen05 pass A handles rows of at most 128 floats whose lists fit next to the operand tiles
bool exact_tcgen05_supported(int dim, int k) {
    if (dim < 1 || dim > 128 || k < 1) return false;
    int dev = 0, max_optin = 0;
    HNSW_CUDA_CHECK(cudaGetDevice(&dev));
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return tc5_scan_smem(dim, exact_tf32_candidates(k)) <= (size_t)max_optin;
}

// ---------------------------------------------------------------------------------------------

int exact_tf32_candidates(int k) { return k + 16 + k / 4; }

static size_t tf32_scan_smem(int dim, int kp, int tq) {
    const int dim8 = (dim + 7) & ~7;
    size_t b = (size_t)tq * (dim8 + 4) * 4;          // queries
    b += (size_t)2 * kTileN * kPitchB * 4;            // stage double buffer
    b += (size_t)tq * kp * 8;                         // lists
    b += (size_t)tq * kTileN * 8;                     // queue keys
    b += (size_t)tq * 4 * 3 + 4;                      // qa, thr, ln, qcount
    b += (size_t)tq * kTileN;                         // queue rows
    return (b + 15) & ~(size_t)15;
}

// queries per block (64 / 32 / 16) that fit the opt-in shared memory, 0 if the shape is not supported
int exact_tf32_block_queries(int dim, int k) {
    if (k < 1 || k > 1024) return 0;
    int dev = 0, max_optin = 0;
    HNSW_CUDA_CHECK(cudaGetDevice(&dev));
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int kp = exact_tf32_candidates(k);
    for (int tq = 64; tq >= 16; tq >>= 1)
        if (tf32_scan_smem(dim, kp, tq) <= (size_t)max_optin) return tq;
    return 0;
}

size_t exact_tf32_scratch_bytes(int nq, int k, int count) {
    const int kp = exact_tf32_candidates(k);
    size_t b = (size_t)nq * tf32_chunks(count) * kp * sizeof(u64);  // per-chunk partial lists
    b += (size_t)nq * kp * (sizeof(int) + sizeof(float));            // merged candidates
    b += (size_t)count * sizeof(float) + 256;                       // xaux, xmax2
    return b + 5 * 256;                                             // every piece starts on a 256 B boundary
}

void launch_exact_knn_tf32(int metric, const GraphView& g, const float* d_queries, int nq, int k, int* d_ids, float* d_dists, int* d_fail,
                           void* d_scratch, size_t scratch_bytes, bool refresh_aux, int tcgen05_mode, cudaStream_t stream) {
    if (nq <= 0) return;
    const bool use_tcgen05 = tcgen05_mode != 0 && exact_tcgen05_supported(g.dim, k);
    const int tq = exact_tf32_block_queries(g.dim, k);
    if (tq == 0) throw std::runtime_error("exact_knn_tf32: dim / k do not fit the tensor-core variant");
    if (scratch_bytes < exact_tf32_scratch_bytes(nq, k, g.count)) throw std::runtime_error("exact_knn_tf32: scratch too small");
    const int kp = exact_tf32_candidates(k);
    int sms = 148;
    {
        int dev = 0;
        HNSW_CUDA_CHECK(cudaGetDevice(&dev));
        HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int chunks = use_tcgen05 ? tc5_chunks(g.count, nq, sms) : tf32_chunks(g.count);  // (the scratch is sized for the larger)
    const int chunk_items = (g.count + chunks - 1) / chunks;
    // scratch layout (256 B aligned pieces): xmax2 | xaux | part keys | candidate ids | candidate distances
    unsigned char* p = static_cast<unsigned char*>(d_scratch);
    auto take = [&](size_t bytes) { unsigned char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    unsigned* d_xmax2 = reinterpret_cast<unsigned*>(take(256));
    float* d_xaux = reinterpret_cast<float*>(take((size_t)g.count * 4));
    u64* d_part = reinterpret_cast<u64*>(take((size_t)nq * chunks * kp * 8));
    int* d_cids = reinterpret_cast<int*>(take((size_t)nq * kp * 4));
    float* d_capx = reinterpret_cast<float*>(take((size_t)nq * kp * 4));
    if ((size_t)(p - static_cast<unsigned char*>(d_scratch)) > scratch_bytes) throw std::runtime_error("exact_knn_tf32: scratch layout overflow");

    const size_t smem = tf32_scan_smem(g.dim, kp, tq);
    dim3 grid(chunks, (nq + tq - 1) / tq);
    const size_t rsmem = (size_t)kWarpsPerBlock * (((g.pitch + 3) & ~3) * 4 + (size_t)k * 8);
    const int rblocks = (nq + kWarpsPerBlock - 1) / kWarpsPerBlock;
#define HNSW_TF32_SCAN(M, WMV)                                                                                                  \
    {                                                                                                                           \
        HNSW_CUDA_CHECK(cudaFuncSetAttribute(exact_tf32_scan_kernel<M, WMV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        exact_tf32_scan_kernel<M, WMV><<<grid, kThreads, smem, stream>>>(g, d_queries, nq, kp, chunk_items, d_xaux, d_part);    \
    }
#define HNSW_TF32_ALL(M)                                                                                                        \
    {                                                                                                                           \
        if (refresh_aux) {                                                                                                      \
            HNSW_CUDA_CHECK(cudaMemsetAsync(d_xmax2, 0, 4, stream));                                                            \
            exact_tf32_aux_kernel<M><<<sms * 8, kThreads, 0, stream>>>(g, d_xaux, d_xmax2);                                     \
            HNSW_CUDA_CHECK(cudaGetLastError());                                                                                \
        }                                                                                                                       \
        if (use_tcgen05) {                                                                                                      \
            const size_t tsmem = tc5_scan_smem(g.dim, kp);                                                                      \
            HNSW_CUDA_CHECK(cudaFuncSetAttribute(exact_tc5_scan_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem)); \
            if (tcgen05_mode == 2) { /* warp-specialised variant (default)      */                                                \
                HNSW_CUDA_CHECK(cudaFuncSetAttribute(exact_tc5ws_scan_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem)); \
                exact_tc5ws_scan_kernel<M><<<dim3(chunks, (nq + tc5::kM - 1) / tc5::kM), 2 * tc5::kM, tsmem, stream>>>(g, d_queries, nq, kp, chunk_items, d_xaux, d_part); \
            } else                                                                                                              \
            exact_tc5_scan_kernel<M><<<dim3(chunks, (nq + tc5::kM - 1) / tc5::kM), tc5::kM, tsmem, stream>>>(g, d_queries, nq, kp, chunk_items, d_xaux, d_part); \
        } else if (tq == 64) HNSW_TF32_SCAN(M, 4) else if (tq == 32) HNSW_TF32_SCAN(M, 2) else HNSW_TF32_SCAN(M, 1)             \
        HNSW_CUDA_CHECK(cudaGetLastError());                                                                                    \
        launch_merge_keys(d_part, chunks, nq, kp, d_cids, d_capx, stream);                                                      \
        HNSW_CUDA_CHECK(cudaFuncSetAttribute(exact_tf32_rerank_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem)); \
        exact_tf32_rerank_kernel<M><<<rblocks, kThreads, rsmem, stream>>>(g, d_queries, nq, k, kp, d_cids, d_capx, d_xmax2, d_ids, d_dists, d_fail); \
        HNSW_CUDA_CHECK(cudaGetLastError());                                                                                    \
    }
    switch (metric) {
        case SQ_EUCLID: HNSW_TF32_ALL(SQ_EUCLID) break;
        case EUCLID: HNSW_TF32_ALL(EUCLID) break;
        case COSINE: HNSW_TF32_ALL(COSINE) break;
        default: HNSW_TF32_ALL(UCOSINE) break;
    }
#undef HNSW_TF32_ALL
#undef HNSW_TF32_SCAN
}

}  // namespace hnsw
