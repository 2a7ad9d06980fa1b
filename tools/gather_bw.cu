// gather_bw.cu -- microbenchmark: ceiling of random 512 B row gathers on B200 for the two load layouts.
//   mode 0: full-row coalesced (one LDG.128 per lane covers one 512 B row per warp instruction)
//   mode 1: pair layout of metric.cuh (lane pair per row, float4 per lane per 32 B chunk), U chunks in flight
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__device__ __forceinline__ unsigned hash32(unsigned x){ x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template<int ROWS_IN_FLIGHT>
__global__ void gather_full(const float4* __restrict__ slab, unsigned nrows, int iters, float* out){
    const int lane = threadIdx.x & 31; const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float acc = 0.f; unsigned s = warp * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it){
        float4 v[ROWS_IN_FLIGHT];
        #pragma unroll
        for (int r = 0; r < ROWS_IN_FLIGHT; ++r){ s = hash32(s + r + 1); unsigned row = s % nrows; v[r] = __ldg(slab + (size_t)row * 32 + lane); }
        #pragma unroll
        for (int r = 0; r < ROWS_IN_FLIGHT; ++r) acc += v[r].x + v[r].y + v[r].z + v[r].w;
    }
    if (acc == 123.456f) out[0] = acc;
}
template<int U>
__global__ void gather_pair(const float4* __restrict__ slab, unsigned nrows, int iters, float* out){
    const int lane = threadIdx.x & 31; const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int h = lane & 1, grp = lane >> 1;
    float acc = 0.f; unsigned s = warp * 2654435761u + 777u;
    for (int it = 0; it < iters; ++it){
        s = hash32(s + it);
        unsigned row = hash32(s + grp * 97u) % nrows;           // 16 different rows per warp
        const float4* p = slab + (size_t)row * 32 + h;
        for (int c = 0; c < 16; c += U){
            float4 v[U];
            #pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(p + 2 * (c + u));
            #pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}
int main(int argc, char** argv){
    const unsigned nrows = argc > 1 ? atoi(argv[1]) : 1000000; const size_t bytes = (size_t)nrows * 512;
    float4* slab; float* out; CK(cudaMalloc(&slab, bytes)); CK(cudaMalloc(&out, 4)); CK(cudaMemset(slab, 1, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto run = [&](const char* name, auto kern, int warps_per_sm, double rows_per_warp_iter, int iters){
        int blocks = sms * warps_per_sm / 4;
        kern<<<blocks, 128>>>(slab, nrows, 4, out); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); kern<<<blocks, 128>>>(slab, nrows, iters, out); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double total = (double)blocks * 4 * iters * rows_per_warp_iter * 512.0;
        printf("%-28s warps/SM=%2d  %8.1f GB/s  (%.3f ms)\n", name, warps_per_sm, total / ms / 1e6, ms);
    };
    for (int w : {8, 16, 24, 32, 48, 64}){
        run("full-row 8 rows in flight", gather_full<8>, w, 8, 2000);
        run("full-row 16 rows in flight", gather_full<16>, w, 16, 1000);
        run("pair U=8 (16 rows/warp)", gather_pair<8>, w, 16, 1000);
        run("pair U=16 (16 rows/warp)", gather_pair<16>, w, 16, 1000);
        run("pair U=4 (16 rows/warp)", gather_pair<4>, w, 16, 1000);
    }
    return 0;
}
