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

// Packed fp32 FMA (fma.rn.f32x2 -> SASS FFMA2): two IEEE fp32 FMAs per instruction.  Measured on B200
// (tools/probes/ffma2_probe.cu): the same 128 FMA/clk/SM as scalar FFMA, i.e. half the issue slots per flop; results
// are bit-identical to two fmaf.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// fma4 (the gather kernels' accumulate) measured NO different packed or scalar — K1 60.4 vs 60.2 ms per layer on the
// 1.88e9-nnz graph, K5 / K6 model steps within noise: those kernels wait on memory, not on issue slots.  Scalar kept.
#ifndef TAGREC_FMA4_PACKED
#define TAGREC_FMA4_PACKED 0
#endif
__device__ __forceinline__ void fma4(float4& acc, float s, const float4& x) {
#if TAGREC_FMA4_PACKED
    const unsigned long long sd = pack2(s, s);
    const unsigned long long lo = ffma2(pack2(x.x, x.y), sd, pack2(acc.x, acc.y));
    const unsigned long long hi = ffma2(pack2(x.z, x.w), sd, pack2(acc.z, acc.w));
    unpack2(lo, acc.x, acc.y);
    unpack2(hi, acc.z, acc.w);
#else
    acc.x = fmaf(s, x.x, acc.x);
    acc.y = fmaf(s, x.y, acc.y);
    acc.z = fmaf(s, x.z, acc.z);
    acc.w = fmaf(s, x.w, acc.w);
#endif
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

// Where an output row is stored.  n == 0: the local table only.  n >= 1: the same table on n ranks of one NVSwitch
// domain — base[r] is rank r's copy mapped into this process (peer memory over NVLink 5), or a single NVLS
// multicast address (n == 1) that the switch replicates to every rank.  This is the all-gather of the sharded path,
// fused into the SpMM epilogue: rows cross NVLink while the next rows are still being gathered from HBM.
struct Mirror {
    int n;
    float* base[TAGREC_MAX_PEERS];
};

__device__ __forceinline__ void store_row(float* local, const Mirror& m, int64_t o, const float4& v) {
    if (m.n == 0) {
        reinterpret_cast<float4*>(local)[o] = v;
        return;
    }
#pragma unroll
    for (int p = 0; p < TAGREC_MAX_PEERS; ++p)
        if (p < m.n) reinterpret_cast<float4*>(m.base[p])[o] = v;
}

// One element of torch.optim.Adam's single-tensor update (amsgrad = False, maximize = False); shared by adam.cu and the
// optimizer epilogue of K1's last backward launch so that both produce the same bits.
__device__ __forceinline__ void adam_update(float& pp, float gg, float& mm, float& vv, float b1, float b2, float eps,
                                            float wd, float step_size, float inv_sqrt_bc2) {
    if (wd != 0.f) gg = fmaf(wd, pp, gg);
    mm = mm + (gg - mm) * (1.f - b1);
    vv = b2 * vv + (1.f - b2) * gg * gg;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    pp = pp - step_size * (mm / denom);
}

inline int set_mirror(Mirror& m, const tagrec_mirror_t* src) {
    m.n = 0;
    if (!src || src->n == 0) return TAGREC_OK;
    TAGREC_REQUIRE(src->n >= 1 && src->n <= TAGREC_MAX_PEERS, "mirror: bad rank count");
    m.n = src->n;
    for (int p = 0; p < src->n; ++p) {
        TAGREC_REQUIRE(src->base[p], "mirror: null base pointer");
        m.base[p] = static_cast<float*>(src->base[p]);
    }
    return TAGREC_OK;
}

}  // namespace tagrec
