// Microbenchmark that decided the P2G design (DESIGN.md "Kernel design notes"): throughput of shared-memory
// int atomics (ATOMS.ADD) and loads under the address patterns P2G/G2P produce, on all 148 SMs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int TILE = 4096;
constexpr int ITERS = 2048;

// pattern: 0 = 32 distinct consecutive words (conflict-free), 1 = groups of 8 lanes share a word (8 particles/cell,
// sorted), 2 = groups of 2, 3 = all lanes one word, 4 = pseudo-random words, 5 = stride-4 words (AoS cell, 8 banks)
__device__ __forceinline__ int addr_of(int pattern, int lane, int it)
{
    switch (pattern) {
        case 0: return (lane + it * 37) & (TILE - 1);
        case 1: return ((lane >> 3) * 11 + it * 37) & (TILE - 1);
        case 2: return ((lane >> 1) + it * 37) & (TILE - 1);
        case 3: return (it * 37) & (TILE - 1);
        case 4: return ((lane * 2654435761u + it * 40503u) >> 7) & (TILE - 1);
        default: return (lane * 4 + it * 37) & (TILE - 1);
    }
}

template <int MODE>  // 0 = ATOMS.ADD, 1 = LDS, 2 = float CAS atomicAdd
__global__ void __launch_bounds__(256) k(int pattern, int* out, long long* cycles)
{
    __shared__ int tile[TILE];
    for (int i = threadIdx.x; i < TILE; i += 256) tile[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int acc = 0;
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
        const int a = addr_of(pattern, lane, it + (threadIdx.x >> 5) * 101);
        if (MODE == 0) atomicAdd(&tile[a], it);
        else if (MODE == 1) acc += tile[a];
        else atomicAdd(reinterpret_cast<float*>(&tile[a]), 1.0f);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 256 + threadIdx.x] = acc + tile[threadIdx.x];
}

int main()
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int* out; long long* cyc;
    const int ctas_per_sm = 4;
    const int grid = nsm * ctas_per_sm;
    cudaMalloc(&out, grid * 256 * sizeof(int));
    cudaMalloc(&cyc, grid * sizeof(long long));
    const char* mode_name[3] = {"ATOMS.ADD(int)", "LDS", "atomicAdd(float) CAS"};
    const char* pat_name[6] = {"distinct", "8-lanes-same-word", "2-lanes-same-word", "all-same-word", "random", "stride-4"};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("SMs=%d, %d CTAs x 256 threads per SM, %d warp-instructions per warp\n", nsm, ctas_per_sm, ITERS);
    for (int mode = 0; mode < 3; ++mode)
        for (int pat = 0; pat < 6; ++pat) {
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<grid, 256>>>(pat, out, cyc);
                else if (mode == 1) k<1><<<grid, 256>>>(pat, out, cyc);
                else k<2><<<grid, 256>>>(pat, out, cyc);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            long long c0; cudaMemcpy(&c0, cyc, sizeof(c0), cudaMemcpyDeviceToHost);
            const double warp_instr_per_sm = (double)ctas_per_sm * 8 * ITERS;
            printf("%-22s %-20s %8.3f ms  %7.2f ns/warp-instr/SM  (CTA0 cycles/instr/warp-set %.2f)\n", mode_name[mode], pat_name[pat],
                   best, best * 1e6 / warp_instr_per_sm, (double)c0 / ITERS);
        }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
