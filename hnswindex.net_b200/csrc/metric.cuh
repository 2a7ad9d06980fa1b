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
// owns AVX lanes 4h..4h+3 and reads them as one float4 (16 B, coalesced into
// a full 32 B sector per pair).  A warp therefore evaluates 16 vectors at once.
// Explicit _rn intrinsics forbid FMA contraction / reassociation.
#pragma once
#include "common.cuh"

namespace hnsw {

__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

// One pair computes Sum_lanes over the 8-lane chains; returns the horizontally
// reduced value (same bits in both lanes of the pair).  `active` false = the pair
// has no vector this pass (it still takes part in the shuffles).
//   SQ: L[j] = fma(d, d, L[j]), d = a - b ; tree (s0+s1)+(s2+s3) ; tail r += d*d (mul, add)
//   DOT: L[j] = L[j] + a*b (mul, add)     ; tree (s0+s2)+(s1+s3) ; tail r += a*b
template <bool SQ, int U>
__device__ __forceinline__ float pair_reduce(const float* __restrict__ vec, const float* q, int dim, int h, bool active) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const int nch = dim >> 3;
    if (active) {
        const float4* v4 = reinterpret_cast<const float4*>(vec) + h;
        const float4* q4 = reinterpret_cast<const float4*>(q) + h;
        int c = 0;
        for (; c + U <= nch; c += U) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ldg_f4(v4 + 2 * (c + u));
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float4 b = q4[2 * (c + u)];
                if (SQ) {
                    float d;
                    d = __fsub_rn(v[u].x, b.x); a0 = __fmaf_rn(d, d, a0);
                    d = __fsub_rn(v[u].y, b.y); a1 = __fmaf_rn(d, d, a1);
                    d = __fsub_rn(v[u].z, b.z); a2 = __fmaf_rn(d, d, a2);
                    d = __fsub_rn(v[u].w, b.w); a3 = __fmaf_rn(d, d, a3);
                } else {
                    a0 = __fadd_rn(a0, __fmul_rn(v[u].x, b.x));
                    a1 = __fadd_rn(a1, __fmul_rn(v[u].y, b.y));
                    a2 = __fadd_rn(a2, __fmul_rn(v[u].z, b.z));
                    a3 = __fadd_rn(a3, __fmul_rn(v[u].w, b.w));
                }
            }
        }
        for (; c < nch; ++c) {
            const float4 a = ldg_f4(v4 + 2 * c);
            const float4 b = q4[2 * c];
            if (SQ) {
                float d;
                d = __fsub_rn(a.x, b.x); a0 = __fmaf_rn(d, d, a0);
                d = __fsub_rn(a.y, b.y); a1 = __fmaf_rn(d, d, a1);
                d = __fsub_rn(a.z, b.z); a2 = __fmaf_rn(d, d, a2);
                d = __fsub_rn(a.w, b.w); a3 = __fmaf_rn(d, d, a3);
            } else {
                a0 = __fadd_rn(a0, __fmul_rn(a.x, b.x));
                a1 = __fadd_rn(a1, __fmul_rn(a.y, b.y));
                a2 = __fadd_rn(a2, __fmul_rn(a.z, b.z));
                a3 = __fadd_rn(a3, __fmul_rn(a.w, b.w));
            }
        }
    }
    // lower + upper half: s[j] = L[j] + L[j+4] (the partner lane holds the other half)
    const float s0 = __fadd_rn(a0, __shfl_xor_sync(FULL, a0, 1));
    const float s1 = __fadd_rn(a1, __shfl_xor_sync(FULL, a1, 1));
    const float s2 = __fadd_rn(a2, __shfl_xor_sync(FULL, a2, 1));
    const float s3 = __fadd_rn(a3, __shfl_xor_sync(FULL, a3, 1));
    float r = SQ ? __fadd_rn(__fadd_rn(s0, s1), __fadd_rn(s2, s3))   // two hadd
                 : __fadd_rn(__fadd_rn(s0, s2), __fadd_rn(s1, s3));  // movehl + shuffle(0x55)
    if (active) {
        for (int i = nch << 3; i < dim; ++i) {  // scalar remainder, both lanes redundantly
            const float a = __ldg(vec + i), b = q[i];
            if (SQ) {
                const float d = __fsub_rn(a, b);
                r = __fadd_rn(r, __fmul_rn(d, d));
            } else {
                r = __fadd_rn(r, __fmul_rn(a, b));
            }
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

template <int METRIC, int U>
__device__ __forceinline__ float pair_distance(const float* __restrict__ vec, const float* q, int dim, int h, bool active,
                                               double sqrt_na, double sqrt_nb) {
    if (METRIC == SQ_EUCLID) return pair_reduce<true, U>(vec, q, dim, h, active);
    const float dot = pair_reduce<false, U>(vec, q, dim, h, active);
    if (METRIC == COSINE) return cosine_finish(dot, sqrt_na, sqrt_nb);
    return __fsub_rn(1.0f, dot);
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
