// common.cuh -- shared definitions for the B200 (sm_100a) HNSW engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace hnsw {

typedef unsigned long long u64;

// EUCLID = sqrt of the squared Euclidean distance (EuclideanMetric.Compute, Metrics/EuclideanMetric.cs:103-109): same
// accumulation as SQ_EUCLID, the root taken once per distance
enum MetricKind : int { SQ_EUCLID = 0, COSINE = 1, UCOSINE = 2, EUCLID = 3 };
__host__ __device__ constexpr bool metric_is_sq(int m) { return m == SQ_EUCLID || m == EUCLID; }

constexpr unsigned FULL = 0xffffffffu;
constexpr int kWarpsPerBlock = 4;     // independent warps (one query / one insert each)
constexpr int kThreads = 32 * kWarpsPerBlock;

#define HNSW_CUDA_CHECK(expr)                                                                        \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");    \
    } while (0)

// Device view of the index.  Layout (DESIGN.md "Data layout in HBM"):
//   vectors   [capacity][pitch] fp32 row-major slab, rows 16 B aligned (pitch % 4 == 0)
//   norm_sqrt [capacity] fp64 sqrt of the cosine norm accumulator (cosine only)
//   adj0      [capacity][stride0] int32, layer-0 rows, unused slots = -1
//   upper_off [capacity] int32, index of the node's layer-1 row in adjU, -1 for level-0 nodes
//   adjU      [rowsU][strideU] int32, rows of layers 1..level of a node are consecutive
//   rm_index  [capacity] int32 removal state (Node.IsRemoved): nothing points at a removed node, its rows are empty
struct GraphView {
    const float* vectors;
    const double* norm_sqrt;
    int* adj0;
    int* adjU;
    const int* upper_off;
    int* locks;
    const int* rm_index;  // [capacity] -1 alive, >= 0 position in the chunk being removed, -2 removed earlier
    int dim, pitch;
    int stride0, strideU;
    int max_edges;      // M (layer >= 1); layer 0 holds 2M
    int entry, top_layer;
    int count;
};

// order-preserving float <-> uint32 (handles negative cosine distances)
__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    uint32_t b; memcpy(&b, &f, 4);
#endif
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u) {
    uint32_t b = u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu);
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f; memcpy(&f, &b, 4); return f;
#endif
}
// candidate key: [63:32] ordered distance | [31:1] node id | [0] expanded flag
__host__ __device__ __forceinline__ u64 make_key(float d, int id) { return ((u64)f2ord(d) << 32) | ((u64)(uint32_t)id << 1); }
__host__ __device__ __forceinline__ int key_id(u64 k) { return (int)((uint32_t)k >> 1); }
__host__ __device__ __forceinline__ float key_dist(u64 k) { return ord2f((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_ord(u64 k) { return (uint32_t)(k >> 32); }

}  // namespace hnsw
