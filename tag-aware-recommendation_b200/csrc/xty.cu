// K8 — skinny X^T Y:  out[a, b] = sum_r x[r, a] * y[r, b]  for tall tables (n = 1e4 .. 1e7 rows) and a, b <= 64.
// sm_100a.
//
// This is the weight gradient of every small dense layer on the path — the attention projections of TGCN
// (model/tgcn.py:26-31: [N, 64] x [64, 32]), DisenGCN's factor projection (model/disengcn.py:25: [N, 64] x [64, 64]) —
// which torch hands to cuBLAS as a [a x n] x [n x b] GEMM.  With a 64 x 32 output there is one CTA tile of work and a
// K loop of n: cuBLAS runs it on a few SMs (91 us per launch at n = 69 K, 180 us for DisenGCN; 12 % and 18 % of those
// models' steps, profiles/r1_models_launches.md).  It is a reduction over rows, so: every CTA streams a slice of the
// rows through shared memory (cp.async, double buffered), keeps the whole a x b result in registers (4 x 4 per
// thread) and adds it to the output once (red.global.add.v4).  Bound: HBM, 4 (a + b) bytes per row.
#include <algorithm>

#include "common.cuh"

namespace tagrec {

constexpr int XR = 32;            // rows per stage (2 stages x (x + y) x 64 floats = 32 KB of static smem)

__device__ __forceinline__ void xty_cp16(void* smem, const void* gmem, bool ok) {
    const int sz = ok ? 16 : 0;   // src-size 0: zero-fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)),
                 "l"(gmem), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(256)
xty_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n, int a, int b, float* __restrict__ out) {
    __shared__ __align__(16) float Xs[2][XR * 64];
    __shared__ __align__(16) float Ys[2][XR * 64];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int a4 = a >> 2, b4 = b >> 2;
    const int64_t n_chunks = (n + XR - 1) / XR;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    auto load = [&](int64_t chunk, int buf) {
        const int64_t r0 = chunk * XR;
        for (int idx = tid; idx < XR * a4; idx += 256) {
            const int r = idx / a4, c = idx % a4;
            const bool ok = r0 + r < n;
            xty_cp16(&Xs[buf][r * a + 4 * c], x + (ok ? (r0 + r) * a + 4 * c : 0), ok);
        }
        for (int idx = tid; idx < XR * b4; idx += 256) {
            const int r = idx / b4, c = idx % b4;
            const bool ok = r0 + r < n;
            xty_cp16(&Ys[buf][r * b + 4 * c], y + (ok ? (r0 + r) * b + 4 * c : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const bool active = ty < a4 && tx < b4;
    int64_t chunk = blockIdx.x;
    if (chunk < n_chunks) load(chunk, 0);
    for (int it = 0; chunk < n_chunks; chunk += gridDim.x, ++it) {
        const int buf = it & 1;
        const int64_t nxt = chunk + gridDim.x;
        if (nxt < n_chunks) {
            load(nxt, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (active) {
#pragma unroll 8
            for (int r = 0; r < XR; ++r) {
                const float4 xv = *reinterpret_cast<const float4*>(&Xs[buf][r * a + 4 * ty]);
                const float4 yv = *reinterpret_cast<const float4*>(&Ys[buf][r * b + 4 * tx]);
                const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
                const float yb[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], yb[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            red_add4(reinterpret_cast<float4*>(out + (size_t)(4 * ty + i) * b + 4 * tx),
                     make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
}

}  // namespace tagrec

using namespace tagrec;

extern "C" int tagrec_xty_acc(const float* x, const float* y, int64_t n, int a, int b, float* out, int accumulate,
                              void* stream) {
    TAGREC_REQUIRE(x && y && out, "null pointer");
    TAGREC_REQUIRE(a >= 4 && a <= 64 && a % 4 == 0 && b >= 4 && b <= 64 && b % 4 == 0,
                   "both widths must be multiples of 4 in 4..64");
    TAGREC_REQUIRE(n >= 0, "negative row count");
    if (!accumulate) TAGREC_CUDA(cudaMemsetAsync(out, 0, (size_t)a * b * 4, (cudaStream_t)stream));
    if (n == 0) return TAGREC_OK;
    const int64_t n_chunks = (n + XR - 1) / XR;
    TAGREC_LAUNCH(xty_kernel, (unsigned)std::min<int64_t>(n_chunks, (int64_t)kSMs * 4), 256, 0, stream, x, y, n, a, b, out);
    return TAGREC_OK;
}

extern "C" int tagrec_xty(const float* x, const float* y, int64_t n, int a, int b, float* out, void* stream) {
    return tagrec_xty_acc(x, y, n, a, b, out, 0, stream);
}
