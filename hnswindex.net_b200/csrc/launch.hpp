// launch.hpp -- launch shape of the persistent one-warp-per-work-item kernels, shared by all launchers.
//
// `warps` independent warps per block (4, fewer when one warp's scratch is large), at most one resident wave of blocks
// (SMs x blocks that fit per SM); work is pulled from a global counter inside the kernels.  The driver queries behind
// the shape (opt-in shared memory, occupancy) cost several microseconds each and depend only on (device, kernel,
// scratch per warp): they are answered once and remembered.  The opt-in attribute is per kernel, not per launch, so
// it is only ever raised.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>

#include "common.cuh"

namespace hnsw {

struct LaunchShape {
    int warps, per_sm, sms;
};

template <typename Kern>
inline LaunchShape launch_shape(Kern kern, size_t smem_per_warp) {
    static std::mutex mu;
    static std::map<std::tuple<int, const void*, size_t>, LaunchShape> cache;
    static std::map<std::pair<int, const void*>, size_t> optin;  // largest dynamic shared memory the kernel was opted in for
    int dev = 0;
    HNSW_CUDA_CHECK(cudaGetDevice(&dev));
    const void* fn = reinterpret_cast<const void*>(kern);
    const auto key = std::make_tuple(dev, fn, smem_per_warp);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    LaunchShape shape;
    int max_optin = 0;
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    HNSW_CUDA_CHECK(cudaDeviceGetAttribute(&shape.sms, cudaDevAttrMultiProcessorCount, dev));
    int wmax = kWarpsPerBlock;
    while (wmax > 1 && smem_per_warp * wmax > (size_t)max_optin) --wmax;  // 4, 3, 2 or 1 warps per block
    if (smem_per_warp * wmax > (size_t)max_optin)
        throw std::runtime_error("search width / MaxCandidates / MaxEdges too large for shared memory: one warp needs " +
                                 std::to_string(smem_per_warp * wmax) + " B, the SM offers " + std::to_string(max_optin));
    size_t& have = optin[std::make_pair(dev, fn)];
    if (smem_per_warp * wmax > have) {
        HNSW_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem_per_warp * wmax)));
        have = smem_per_warp * wmax;
    }
    // Blocks of fewer warps waste less of the SM's shared memory to granularity (one warp's scratch is 10-80 KB): take the
    // block size with the most resident WARPS, the larger block on ties (the C2 hot path stays at 5 x 4).
    shape.warps = wmax;
    shape.per_sm = 0;
    for (int w = wmax; w >= 1; --w) {
        int per = 0;
        HNSW_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, 32 * w, smem_per_warp * w));
        if (per * w > shape.per_sm * shape.warps) {
            shape.warps = w;
            shape.per_sm = per;
        }
    }
    if (shape.per_sm < 1) shape.per_sm = 1;
    cache.emplace(key, shape);
    return shape;
}

// grid size for `work_items` items; sets `warps` (warps per block).  blocks_per_sm_cap > 0 lowers the residency (tuning).
template <typename Kern>
inline int persistent_launch(Kern kern, size_t smem_per_warp, int blocks_per_sm_cap, int work_items, int& warps) {
    const LaunchShape s = launch_shape(kern, smem_per_warp);
    warps = s.warps;
    int per_sm = s.per_sm;
    if (blocks_per_sm_cap > 0) per_sm = std::min(per_sm, blocks_per_sm_cap);
    const int need = (work_items + warps - 1) / warps;
    return std::max(1, std::min(need, s.sms * per_sm));
}

}  // namespace hnsw
