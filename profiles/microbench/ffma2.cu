// ffma2.cu -- does the packed fp32 FMA of sm_100 (fma.rn.f32x2, __ffma2_rn) double the FMAs per issue slot?
// Same arithmetic (16 independent accumulators per thread, 4096 rounds) as scalar FFMA and as FFMA2.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_scalar(float* out, float a, float b, int rounds)
{
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = threadIdx.x * 1e-3f + k;
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = fmaf(acc[k], a, b);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_packed(float* out, float a, float b, int rounds)
{
    float2 acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float2(threadIdx.x * 1e-3f + 2 * k, threadIdx.x * 1e-3f + 2 * k + 1);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = __ffma2_rn(acc[k], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: the shape of the P2G inner loop -- per 2 packed FMAs one scalar integer instruction competes for issue
__global__ void k_scalar_mix(float* out, float a, float b, int rounds, int* iout)
{
    float acc[16];
    int z = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = threadIdx.x * 1e-3f + k;
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { acc[k] = fmaf(acc[k], a, b); if ((k & 3) == 3) z = (z ^ r) + k; }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = z;
}

__global__ void k_packed_mix(float* out, float a, float b, int rounds, int* iout)
{
    float2 acc[8];
    int z = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = make_float2(threadIdx.x * 1e-3f + 2 * k, threadIdx.x * 1e-3f + 2 * k + 1);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int r = 0; r < rounds; ++r) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] = __ffma2_rn(acc[k], a2, b2); if (k & 1) z = (z ^ r) + k; }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = z;
}

int main()
{
    const int blocks = 148 * 8, threads = 256, rounds = 4096;
    float* out; int* iout;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&iout, sizeof(int) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double flop = 2.0 * 16 * rounds * (double)blocks * threads;
    for (int which = 0; which < 4; ++which) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_scalar<<<blocks, threads>>>(out, 1.0001f, 0.5f, rounds);
            else if (which == 1) k_packed<<<blocks, threads>>>(out, 1.0001f, 0.5f, rounds);
            else if (which == 2) k_scalar_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f, rounds, iout);
            else k_packed_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f, rounds, iout);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        const char* name[] = {"FFMA   (scalar)", "FFMA2  (packed)", "FFMA  + 1 int / 4 FMA", "FFMA2 + 1 int / 4 FMA"};
        printf("%-24s %8.3f ms  %7.2f TFLOP/s\n", name[which], best, flop / best * 1e-9);
    }
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
