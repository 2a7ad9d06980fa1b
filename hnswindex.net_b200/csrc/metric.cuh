// metric.cuh -- K1: bit-exact fp32 distance device functions.
//
// Replaces the AVX+FMA branches of the reference's metrics
//   SquaredEuclideanMetric.Compute  Metrics/EuclideanMetric.cs:19-60
//   CosineMetric.Compute            Metrics/CosineMetric.cs:23-49,78-91,145-171
//   CosineMetric.UnitCompute        Metrics/CosineMetric.cs:103-118,135-141
// Those keep eight independent lane accumulators L[j] (j = AVX lane) that each
// sum elements 8c+j sequentially in c, then combine lanes with a fixed tree.
// fp32 addition is not associative, so the same eight sequential chains are
// kept here: one vector is handled by a PAIR of adjacent lanes; lane h (0/1)
// owns AVX lanes 4h..4h+3 and reads them as one float4.  A warp therefore
// evaluates 16 vectors at once.  The stored rows are NOT gathered with this
// layout from global memory (16 different 128 B lines per load instruction
// cap a B200 at ~1.8 TB/s, profiles/r01_microbench_gather_*.log): whole rows
// are bulk-copied to shared memory first (stage.cuh) and read from there.
// Explicit _rn intrinsics forbid FMA contraction / reassociation.
#pragma once
#include "common.cuh"

namespace hnsw {

__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

// Lane accumulators of one lane pair: lane h holds L[4h..4h+3].
struct PairAcc {
    float a0, a1, a2, a3;
};

// Accumulate `nch` 8-element chunks of a stored vector against the query, both read from SHARED memory
// (the stored row was staged there by a bulk copy, see stage.cuh).  Sequential in c, as the AVX loop.
//   SQ : L[j] = fma(d, d, L[j]), d = a - b       (EuclideanMetric.cs:29-30)
//   DOT: L[j] = L[j] + a*b   (mul, then add)     (CosineMetric.cs:37-38, 114-115)
// Blackwell packed fp32x2 arithmetic (SASS FADD2 / FFMA2): two IEEE round-to-nearest operations per
// instruction, bit-identical to the scalar ones.  Only used for the sq_euclid chain (sub, then a real
// FMA); the cosine chains need mul-then-add WITHOUT fusion and ptxas contracts mul.rn.f32x2 +
// add.rn.f32x2 into FFMA2, so they stay scalar.
__device__ __forceinline__ u64 pack_f32x2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f32x2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub_f32x2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma_f32x2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <bool SQ>
__device__ __forceinline__ void pair_accumulate(PairAcc& acc, const float* row, const float* q, int nch, int h) {
    const float4* v4 = reinterpret_cast<const float4*>(row) + h;
    const float4* q4 = reinterpret_cast<const float4*>(q) + h;
    if (SQ) {
        u64 lo = pack_f32x2(acc.a0, acc.a1), hi = pack_f32x2(acc.a2, acc.a3);
#pragma unroll 4
        for (int c = 0; c < nch; ++c) {
            const float4 a = v4[2 * c];
            const float4 b = q4[2 * c];
            const u64 dlo = sub_f32x2(pack_f32x2(a.x, a.y), pack_f32x2(b.x, b.y));
            const u64 dhi = sub_f32x2(pack_f32x2(a.z, a.w), pack_f32x2(b.z, b.w));
            lo = fma_f32x2(dlo, dlo, lo);
            hi = fma_f32x2(dhi, dhi, hi);
        }
        unpack_f32x2(lo, acc.a0, acc.a1);
        unpack_f32x2(hi, acc.a2, acc.a3);
    } else {
#pragma unroll 4
        for (int c = 0; c < nch; ++c) {
            const float4 a = v4[2 * c];
            const float4 b = q4[2 * c];
            acc.a0 = __fadd_rn(acc.a0, __fmul_rn(a.x, b.x));
            acc.a1 = __fadd_rn(acc.a1, __fmul_rn(a.y, b.y));
            acc.a2 = __fadd_rn(acc.a2, __fmul_rn(a.z, b.z));
            acc.a3 = __fadd_rn(acc.a3, __fmul_rn(a.w, b.w));
        }
    }
}

// Horizontal reduction (all 32 lanes must call): s[j] = L[j] + L[j+4] with the partner lane, then
//   SQ : (s0+s1)+(s2+s3)  -- two hadd            (EuclideanMetric.cs:45-49)
//   DOT: (s0+s2)+(s1+s3)  -- movehl + shuffle    (CosineMetric.cs:145-171)
template <bool SQ>
__device__ __forceinline__ float pair_hsum(const PairAcc& acc) {
    const float s0 = __fadd_rn(acc.a0, __shfl_xor_sync(FULL, acc.a0, 1));
    const float s1 = __fadd_rn(acc.a1, __shfl_xor_sync(FULL, acc.a1, 1));
    const float s2 = __fadd_rn(acc.a2, __shfl_xor_sync(FULL, acc.a2, 1));
    const float s3 = __fadd_rn(acc.a3, __shfl_xor_sync(FULL, acc.a3, 1));
    return SQ ? __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3)) : __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));
}

// Scalar remainder (dim % 8 elements), separate mul and add (EuclideanMetric.cs:53-57, CosineMetric.cs:78-85)
template <bool SQ>
__device__ __forceinline__ float pair_tail(float r, const float* row, const float* q, int n) {
    for (int i = 0; i < n; ++i) {
        const float a = row[i], b = q[i];
        if (SQ) {
            const float d = __fsub_rn(a, b);
            r = __fadd_rn(r, __fmul_rn(d, d));
        } else {
            r = __fadd_rn(r, __fmul_rn(a, b));
        }
    }
    return r;
}

// CosineMetric.cs:88-91 with both norms pre-rooted in fp64
__device__ __forceinline__ float cosine_finish(float dot, double sqrt_na, double sqrt_nb) {
    const float denom = __double2float_rn(__dmul_rn(sqrt_na, sqrt_nb));
    if (denom < 1e-30f) return 1.0f;
    return __fsub_rn(1.0f, __fdiv_rn(dot, denom));
}

template <int METRIC>
__device__ __forceinline__ float metric_finish(float r, double sqrt_na, double sqrt_nb) {
    if (METRIC == SQ_EUCLID) return r;
    if (METRIC == EUCLID) return __fsqrt_rn(r);  // (float)Math.Sqrt(x): the double root rounded to float is the correctly rounded float root
    if (METRIC == COSINE) return cosine_finish(r, sqrt_na, sqrt_nb);
    return __fsub_rn(1.0f, r);  // CosineMetric.cs:141
}

// Norm accumulator of one vector in the cosine metric's order (A[j] += a*a; cosine tree; tail),
// i.e. normAScalar of CosineMetric.cs:41-42,57,82 -- executed by one lane pair, src may be
// global or shared memory.  Returns sqrt((double)norm).
__device__ __forceinline__ double pair_norm_sqrt(const float* src, int dim, int h) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int nch = dim >> 3;
    const float4* v4 = reinterpret_cast<const float4*>(src) + h;
    for (int c = 0; c < nch; ++c) {
        const float4 a = v4[2 * c];
        a0 = __fadd_rn(a0, __fmul_rn(a.x, a.x));
        a1 = __fadd_rn(a1, __fmul_rn(a.y, a.y));
        a2 = __fadd_rn(a2, __fmul_rn(a.z, a.z));
        a3 = __fadd_rn(a3, __fmul_rn(a.w, a.w));
    }
    const float s0 = __fadd_rn(a0, __shfl_xor_sync(FULL, a0, 1));
    const float s1 = __fadd_rn(a1, __shfl_xor_sync(FULL, a1, 1));
    const float s2 = __fadd_rn(a2, __shfl_xor_sync(FULL, a2, 1));
    const float s3 = __fadd_rn(a3, __shfl_xor_sync(FULL, a3, 1));
    float r = __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));
    for (int i = nch << 3; i < dim; ++i) r = __fadd_rn(r, __fmul_rn(src[i], src[i]));
    return sqrt((double)r);
}

}  // namespace hnsw
