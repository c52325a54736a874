// Throughput of packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu ; run on a B200.
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = threadIdx.x * 0.001f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], a, b);
        }
    } else {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(acc);
        float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
        const unsigned long long ar = *reinterpret_cast<unsigned long long*>(&a2), br = *reinterpret_cast<unsigned long long*>(&b2);
        unsigned long long v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = p[i];
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fma2(v[i], ar, br);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) p[i] = v[i];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int iters = 20000;
    for (int blocks_per_sm = 1; blocks_per_sm <= 4; blocks_per_sm *= 2) {
        for (int mode = 0; mode < 2; ++mode) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (mode == 0) k<0><<<148 * blocks_per_sm, 256>>>(out, iters, 0.999f, 0.001f);
                else k<1><<<148 * blocks_per_sm, 256>>>(out, iters, 0.999f, 0.001f);
                cudaEventRecord(b);
                cudaEventSynchronize(b);
            }
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            const double fma = 148.0 * blocks_per_sm * 256 * 32.0 * iters;
            printf("%s  %d x 8 warps/SM: %.3f ms  %.1f TFLOP/s  (%.1f FMA/clk/SM at 1.9 GHz)\n", mode ? "FFMA2" : "FFMA ",
                   blocks_per_sm, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.9e9);
        }
    }
    return 0;
}
