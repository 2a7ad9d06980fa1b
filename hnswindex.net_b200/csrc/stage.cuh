// stage.cuh -- per-warp staging of stored vectors in shared memory with TMA bulk copies.
//
// A gathered neighbour row is one contiguous 16 B-aligned run of `pitch` floats.  Instead of letting
// the lanes gather it in the pair layout (one 32 B sector out of 16 different 128 B lines per load
// instruction), lane r issues ONE `cp.async.bulk.shared.global` (SASS: UBLKCP, the TMA's 1-D bulk
// copy) for row r; the copy engine streams whole lines and signals a per-warp mbarrier with the
// byte count.  All rows of a pass are in flight at once without holding registers.  The lanes then
// read the rows back with conflict-free LDS.128 (rows are staged 544 B apart: 8 banks of skew per
// row, so the 16 pairs of a warp cover all 32 banks exactly four times).
// Measured on B200 (profiles/r01_microbench_gather_staging.log): 8.4 TB/s incl. the read-back, at any
// occupancy >= 8 warps/SM, vs 1.8 TB/s for direct pair-layout loads.
//
// Rows longer than 128 floats are processed in segments of 128 floats; the lane accumulators stay
// in registers across segments, so the sequential accumulation order is unchanged.
#pragma once
#include "common.cuh"

namespace hnsw {

constexpr int kStageRows = 16;    // rows per pass = lane pairs per warp
constexpr int kSegFloats = 128;   // floats of one row staged per segment (512 B)
constexpr int kSegPitch = 136;    // floats between staged rows (544 B)
constexpr int kStageBytes = kStageRows * kSegPitch * 4 + 16;  // + the mbarrier slot

struct Stager {
    float* buf;        // [kStageRows][kSegPitch]
    uint32_t buf_s;    // shared-space address of buf
    uint32_t bar;      // shared-space address of the warp's mbarrier
    uint32_t phase;    // parity of the phase to wait for next
};

__device__ __forceinline__ void stager_init(Stager& st, unsigned char* base, int lane) {
    st.buf = reinterpret_cast<float*>(base);
    st.buf_s = (uint32_t)__cvta_generic_to_shared(base);
    st.bar = (uint32_t)__cvta_generic_to_shared(base + kStageRows * kSegPitch * 4);
    st.phase = 0;
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st.bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
}

// lane 0 only: one arrival + the number of bytes the following bulk copies will deliver
__device__ __forceinline__ void stage_expect(const Stager& st, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st.bar), "r"(bytes) : "memory");
}
// one lane: global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16 B aligned)
__device__ __forceinline__ void stage_copy(const Stager& st, float* dst, const float* src, uint32_t bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src), "r"(bytes),
                 "r"(st.bar)
                 : "memory");
}
// all lanes: wait until every byte announced by stage_expect has landed
__device__ __forceinline__ void stage_wait(Stager& st) {
    uint32_t done = 0, spins = 0;
    while (!done) {
        if (++spins > (1u << 24)) __trap();  // a lost copy must fail the launch, never hang the GPU
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(st.bar), "r"(st.phase)
            : "memory");
    }
    st.phase ^= 1u;
}

// ---- alternative: Ampere-style cp.async (SASS LDGSTS), 16 B per lane, one warp instruction per row -------------
// Same full-row access pattern and bandwidth as the bulk copy in the microbenchmark, but ~5 instead of ~10
// issue slots per staged row (ptxas serialises per-lane UBLKCP through ELECT/R2UR because the instruction
// takes uniform operands), which matters once the search kernel is issue-bound.
__device__ __forceinline__ void cp_async_16(float* dst, const float* src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
// same with the destination already a shared-space address and the source a byte pointer
__device__ __forceinline__ void cp_async_16s(uint32_t dst_s, const char* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace hnsw
