// gather_bw2.cu -- pair-layout load variants + cp.async full-row staging. See gather_bw.cu.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
__device__ __forceinline__ unsigned hash32(unsigned x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ float4 ld_na(const float4* p){ float4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x),"=f"(r.y),"=f"(r.z),"=f"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ float4 ld_cg(const float4* p){ return __ldcg(p); }
__device__ __forceinline__ float4 ld_ca(const float4* p){ return __ldg(p); }
__device__ __forceinline__ float4 ld_cs(const float4* p){ return __ldcs(p); }

template<int MODE, int U>
__global__ void gather_pair(const float4* __restrict__ slab, unsigned nrows, int iters, float* out){
    const int lane = threadIdx.x & 31; const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int h = lane & 1, grp = lane >> 1;
    float acc = 0.f; unsigned s = warp * 2654435761u + 777u;
    for (int it = 0; it < iters; ++it){
        s = hash32(s + it);
        unsigned row = hash32(s + grp * 97u) % nrows;
        const float4* p = slab + (size_t)row * 32 + h;
        for (int c = 0; c < 16; c += U){
            float4 v[U];
            #pragma unroll
            for (int u = 0; u < U; ++u) v[u] = MODE == 0 ? ld_ca(p + 2*(c+u)) : MODE == 1 ? ld_cg(p + 2*(c+u)) : MODE == 2 ? ld_na(p + 2*(c+u)) : ld_cs(p + 2*(c+u));
            #pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
// cp.async: each warp stages 16 random rows (full-row coalesced 16 B per lane) into padded smem, waits, then reads
// them back in the pair layout (LDS.128) and accumulates -- the complete proposed scheme.
constexpr int PITCH4 = 34;  // 544 B per staged row: conflict-free pair-layout reads
template<int ROWS>
__global__ void gather_cpasync(const float4* __restrict__ slab, unsigned nrows, int iters, float* out){
    extern __shared__ float4 smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5; const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4* st = smem + (size_t)wib * ROWS * PITCH4;
    const int h = lane & 1, grp = lane >> 1;
    float acc = 0.f; unsigned s = warp * 2654435761u + 4242u;
    for (int it = 0; it < iters; ++it){
        s = hash32(s + it);
        #pragma unroll
        for (int r = 0; r < ROWS; ++r){
            unsigned row = hash32(s + r * 97u) % nrows;
            const float4* src = slab + (size_t)row * 32 + lane;
            unsigned dst = (unsigned)__cvta_generic_to_shared(st + r * PITCH4 + lane);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src));
        }
        asm volatile("cp.async.commit_group;"); asm volatile("cp.async.wait_group 0;");
        __syncwarp();
        for (int base = 0; base < ROWS; base += 16){
            const float4* p = st + (base + grp) * PITCH4 + h;
            #pragma unroll
            for (int c = 0; c < 16; ++c){ float4 v = p[2*c]; acc += v.x + v.y + v.z + v.w; }
        }
        __syncwarp();
    }
    if (acc == 123.456f) out[0] = acc;
}
// cp.async.bulk (TMA 1-D bulk copy): lane r issues the 512 B copy of row r, completion on a per-warp mbarrier.
template<int ROWS>
__global__ void gather_bulk(const float4* __restrict__ slab, unsigned nrows, int iters, float* out){
    extern __shared__ float4 smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5; const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4* st = smem + (size_t)wib * ROWS * PITCH4;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)4 * ROWS * PITCH4);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(bars + wib);
    if (lane == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int h = lane & 1, grp = lane >> 1;
    float acc = 0.f; unsigned s = warp * 2654435761u + 4242u; unsigned phase = 0;
    for (int it = 0; it < iters; ++it){
        s = hash32(s + it);
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(ROWS * 512) : "memory");
        __syncwarp();
        for (int r = lane; r < ROWS; r += 32){
            unsigned row = hash32(s + r * 97u) % nrows;
            const float4* src = slab + (size_t)row * 32;
            unsigned dst = (unsigned)__cvta_generic_to_shared(st + r * PITCH4);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];" :: "r"(dst), "l"(src), "r"(bar) : "memory");
        }
        unsigned done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        phase ^= 1;
        for (int base = 0; base < ROWS; base += 16){
            const float4* p = st + (base + grp) * PITCH4 + h;
            #pragma unroll
            for (int c = 0; c < 16; ++c){ float4 v = p[2*c]; acc += v.x + v.y + v.z + v.w; }
        }
        __syncwarp();
    }
    if (acc == 123.456f) out[0] = acc;
}
int main(int argc, char** argv){
    const unsigned nrows = argc > 1 ? atoi(argv[1]) : 1000000; const size_t bytes = (size_t)nrows * 512;
    float4* slab; float* out; CK(cudaMalloc(&slab, bytes)); CK(cudaMalloc(&out, 4)); CK(cudaMemset(slab, 1, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto run = [&](const char* name, auto kern, int warps_per_sm, double rows_per_warp_iter, int iters, size_t smem){
        int blocks = sms * warps_per_sm / 4;
        if (smem) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<blocks, 128, smem>>>(slab, nrows, 4, out); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); kern<<<blocks, 128, smem>>>(slab, nrows, iters, out); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double total = (double)blocks * 4 * iters * rows_per_warp_iter * 512.0;
        printf("%-34s warps/SM=%2d  %8.1f GB/s  (%.3f ms)\n", name, warps_per_sm, total / ms / 1e6, ms);
    };
    for (int w : {8, 12, 16, 20, 24, 32}){
        run("pair ld.ca (ldg) U=8", gather_pair<0,8>, w, 16, 500, 0);
        run("pair ld.cg U=8", gather_pair<1,8>, w, 16, 500, 0);
        run("pair ld.nc.L1::no_allocate U=8", gather_pair<2,8>, w, 16, 500, 0);
        run("pair ld.cs U=8", gather_pair<3,8>, w, 16, 500, 0);
        run("pair ld.nc.L1::no_allocate U=16", gather_pair<2,16>, w, 16, 500, 0);
        if (w * 16 * 544 <= 220 * 1024) run("cp.async 16 rows + pair LDS", gather_cpasync<16>, w, 16, 500, (size_t)4 * 16 * 544);
        if (w * 32 * 544 <= 220 * 1024) run("cp.async 32 rows + pair LDS", gather_cpasync<32>, w, 32, 250, (size_t)4 * 32 * 544);
        if (w * 16 * 544 <= 220 * 1024) run("bulk(TMA) 16 rows + pair LDS", gather_bulk<16>, w, 16, 500, (size_t)4 * 16 * 544 + 64);
        if (w * 32 * 544 <= 220 * 1024) run("bulk(TMA) 32 rows + pair LDS", gather_bulk<32>, w, 32, 250, (size_t)4 * 32 * 544 + 64);
    }
    return 0;
}
