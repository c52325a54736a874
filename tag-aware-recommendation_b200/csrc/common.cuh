// Shared helpers for libtagrec_b200 (sm_100a).  Error reporting, launch accounting, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/tagrec_b200.h"

namespace tagrec {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s (%s:%d)", what, file, line);
    g_last_error = buf;
    return code;
}

#define TAGREC_REQUIRE(cond, msg) \
    do { if (!(cond)) return ::tagrec::fail(TAGREC_EINVAL, msg, __FILE__, __LINE__); } while (0)

#define TAGREC_CUDA(expr) \
    do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
        return ::tagrec::fail(TAGREC_ECUDA, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// Every kernel launch of the library goes through this so bench.py can report "gpu_launches".
#define TAGREC_LAUNCH(kernel, grid, block, smem, stream, ...) \
    do { kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__); \
         ::tagrec::g_launches.fetch_add(1, std::memory_order_relaxed); \
         TAGREC_CUDA(cudaGetLastError()); } while (0)

constexpr int kSMs = 148;   // B200

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ void fma4(float4& acc, float s, const float4& x) {
    acc.x = fmaf(s, x.x, acc.x);
    acc.y = fmaf(s, x.y, acc.y);
    acc.z = fmaf(s, x.z, acc.z);
    acc.w = fmaf(s, x.w, acc.w);
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// Sum over the 16 lanes of a half warp (mask = lanes of that half).
__device__ __forceinline__ float half_sum(float v, unsigned mask) {
    v += __shfl_xor_sync(mask, v, 8, 16);
    v += __shfl_xor_sync(mask, v, 4, 16);
    v += __shfl_xor_sync(mask, v, 2, 16);
    v += __shfl_xor_sync(mask, v, 1, 16);
    return v;
}

// 128-bit vector reduction to global memory (sm_90+: red.global.add.v4.f32).
__device__ __forceinline__ void red_add4(float4* addr, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

}  // namespace tagrec
