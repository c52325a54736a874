// K6 — the dense half of an NGCF layer, fused (64 -> 64 layers).  sm_100a.
//
// Replaces model/ngcf.py:77-86 per layer: two adds/muls, two matmuls against (W + b) — the bias is broadcast-added
// to the WEIGHT matrix, not to the output (ngcf.py:78,82) — two LeakyReLU(0.2), the sum, and F.normalize,
// i.e. ~10 torch kernels each streaming N x 64 tables, plus their autograd.  Here: one forward pass and one
// backward pass over the rows.
//
//   forward   x1 = nei + e, x2 = nei * e;  s = lrelu(x1 (W1+b1)), t = lrelu(x2 (W2+b2));  out = s + t;
//             nrm = out / max(||out||_2, 1e-12)                      writes out, nrm, s, t   (s, t: saved activations)
//   backward  g  = g_out + J_normalize(out)^T g_nrm;  gs = g * lrelu'(s), gt = g * lrelu'(t)
//             gx1 = gs (W1+b1)^T, gx2 = gt (W2+b2)^T;  g_nei = gx1 + gx2 * e;  g_e = gx1 + gx2 * nei
//             d(W1+b1) = (nei+e)^T gs, d(W2+b2) = (nei*e)^T gt   (outer products over the tile's rows, accumulated in
//             registers over all tiles of a block, one atomicAdd per element and block at the end)
//             writes g_nei, g_e, dW1, dW2 (+ gs, gt on request)
//
// Mapping: a block owns 64-row tiles; the two 64x64 weight matrices live in smem for the whole kernel; a thread
// owns a 4x4 micro-tile of the 64x64 output tile and walks the contraction index in float4 steps (sequential fp32
// accumulation).  Memory-bound: 6 (fwd) / 10 (bwd) row-tables of 256 B per row.
#include <algorithm>

#include "common.cuh"

namespace tagrec {

constexpr int ND = 64;          // layer width
constexpr int NR = 64;          // rows per tile
constexpr int NLD = ND + 4;     // smem row pitch (floats): keeps float4 alignment, spreads rows over banks

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.2f * x; }

// acc[r][c] += sum_i X[(ty*4+r)][i] * W[i][tx*4+c]
// Packed FMAs (FFMA2, common.cuh): the column pairs come straight out of the float4 loads of W, the x operand is
// broadcast; every output's own accumulation order over i is unchanged (bit-identical to the scalar form).
__device__ __forceinline__ void tile_mm(const float* __restrict__ Xs, const float* __restrict__ Ws, int ty, int tx,
                                        float (&acc)[4][4]) {
    unsigned long long p[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        p[r][0] = pack2(acc[r][0], acc[r][1]);
        p[r][1] = pack2(acc[r][2], acc[r][3]);
    }
#pragma unroll 4
    for (int i = 0; i < ND; i += 4) {
        float4 x[4], w[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) x[r] = *reinterpret_cast<const float4*>(Xs + (ty * 4 + r) * NLD + i);
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = *reinterpret_cast<const float4*>(Ws + (i + k) * NLD + tx * 4);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float xv[4] = {x[r].x, x[r].y, x[r].z, x[r].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned long long xd = pack2(xv[k], xv[k]);
                p[r][0] = ffma2(xd, pack2(w[k].x, w[k].y), p[r][0]);
                p[r][1] = ffma2(xd, pack2(w[k].z, w[k].w), p[r][1]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        unpack2(p[r][0], acc[r][0], acc[r][1]);
        unpack2(p[r][1], acc[r][2], acc[r][3]);
    }
}

__device__ __forceinline__ float sum16(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

__global__ void __launch_bounds__(256)
ngcf_dense_fwd_kernel(const float* __restrict__ nei, const float* __restrict__ e, const float* __restrict__ W1,
                      const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                      int64_t n, float* __restrict__ out, float* __restrict__ nrm, float* __restrict__ s_act,
                      float* __restrict__ t_act) {
    extern __shared__ __align__(16) float sm[];
    float* W1s = sm;                 // [64][NLD]  W1 + b1
    float* W2s = W1s + ND * NLD;
    float* X1s = W2s + ND * NLD;     // [64][NLD]  nei + e
    float* X2s = X1s + NR * NLD;     //            nei * e
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int idx = tid; idx < ND * ND; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        W1s[i * NLD + j] = __ldg(W1 + idx) + __ldg(b1 + j);
        W2s[i * NLD + j] = __ldg(W2 + idx) + __ldg(b2 + j);
    }
    const int64_t tiles = (n + NR - 1) / NR;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t r0 = tile * NR;
        __syncthreads();
        for (int idx = tid; idx < NR * (ND / 4); idx += 256) {
            const int r = idx >> 4, c4 = idx & 15;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (r0 + r < n) {
                a = __ldg(reinterpret_cast<const float4*>(nei + (r0 + r) * ND) + c4);
                b = __ldg(reinterpret_cast<const float4*>(e + (r0 + r) * ND) + c4);
            }
            *reinterpret_cast<float4*>(X1s + r * NLD + 4 * c4) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            *reinterpret_cast<float4*>(X2s + r * NLD + 4 * c4) = make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
        }
        __syncthreads();
        float s[4][4] = {}, t[4][4] = {};
        tile_mm(X1s, W1s, ty, tx, s);
        tile_mm(X2s, W2s, ty, tx, t);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int64_t row = r0 + ty * 4 + r;
            float4 sv = make_float4(lrelu(s[r][0]), lrelu(s[r][1]), lrelu(s[r][2]), lrelu(s[r][3]));
            float4 tv = make_float4(lrelu(t[r][0]), lrelu(t[r][1]), lrelu(t[r][2]), lrelu(t[r][3]));
            float4 o = make_float4(sv.x + tv.x, sv.y + tv.y, sv.z + tv.z, sv.w + tv.w);
            const float nn = fmaxf(sqrtf(sum16(dot4(o, o))), 1e-12f);     // F.normalize(p=2, eps=1e-12)
            if (row < n) {
                const int64_t off = row * (ND / 4) + tx;
                reinterpret_cast<float4*>(out)[off] = o;
                reinterpret_cast<float4*>(nrm)[off] = make_float4(o.x / nn, o.y / nn, o.z / nn, o.w / nn);
                reinterpret_cast<float4*>(s_act)[off] = sv;
                reinterpret_cast<float4*>(t_act)[off] = tv;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
ngcf_dense_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ g_nrm, int64_t g_nrm_ld,
                      const float* __restrict__ out, const float* __restrict__ s_act, const float* __restrict__ t_act,
                      const float* __restrict__ nei, const float* __restrict__ e, const float* __restrict__ W1,
                      const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                      int64_t n, float* __restrict__ g_nei, float* __restrict__ g_e, float* __restrict__ gs_out,
                      float* __restrict__ gt_out, float* __restrict__ dw1, float* __restrict__ dw2) {
    extern __shared__ __align__(16) float sm[];
    float* W1t = sm;                 // [64][NLD]  (W1 + b1)^T : W1t[j][i]
    float* W2t = W1t + ND * NLD;
    float* Gs = W2t + ND * NLD;      // [64][NLD]  gs tile
    float* Gt = Gs + NR * NLD;
    float* Ns = Gt + NR * NLD;       // [64][NLD]  nei tile
    float* Es = Ns + NR * NLD;       // [64][NLD]  e tile
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    for (int idx = tid; idx < ND * ND; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        W1t[j * NLD + i] = __ldg(W1 + idx) + __ldg(b1 + j);
        W2t[j * NLD + i] = __ldg(W2 + idx) + __ldg(b2 + j);
    }
    // weight gradients of this block: dW1[i][j] = sum_r (nei+e)[r][i] gs[r][j], dW2 likewise with nei*e and gt;
    // thread (ty, tx) owns rows 4ty..4ty+3, columns 4tx..4tx+3, accumulated over all tiles of the block
    unsigned long long a1p[4][2] = {}, a2p[4][2] = {};     // weight-gradient blocks, column pairs packed (FFMA2)
    const int64_t tiles = (n + NR - 1) / NR;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t r0 = tile * NR;
        __syncthreads();
        // ---- g = g_out + J^T g_nrm ; gs, gt, nei, e -> smem (16 lanes per row, 4 rows per pass) ----
        for (int rr = ty; rr < NR; rr += 16) {
            const int64_t row = r0 + rr;
            float4 gs4 = make_float4(0.f, 0.f, 0.f, 0.f), gt4 = gs4, n4 = gs4, e4 = gs4;
            if (row < n) {          // uniform per half-warp (16 lanes share rr)
                const int64_t off = row * (ND / 4) + tx;
                const float4 o = __ldg(reinterpret_cast<const float4*>(out) + off);
                const float4 gn = __ldg(reinterpret_cast<const float4*>(g_nrm + row * g_nrm_ld) + tx);
                float4 g = g_out ? __ldg(reinterpret_cast<const float4*>(g_out) + off) : make_float4(0.f, 0.f, 0.f, 0.f);
                n4 = __ldg(reinterpret_cast<const float4*>(nei) + off);
                e4 = __ldg(reinterpret_cast<const float4*>(e) + off);
                const unsigned hm = 0xffffu << (16 * ((tid >> 4) & 1));
                const float ss = half_sum(dot4(o, o), hm);
                const float dt = half_sum(dot4(o, gn), hm);
                const float nn = sqrtf(ss);
                if (nn >= 1e-12f) {
                    const float proj = dt / nn;
                    g.x += (gn.x - (o.x / nn) * proj) / nn;
                    g.y += (gn.y - (o.y / nn) * proj) / nn;
                    g.z += (gn.z - (o.z / nn) * proj) / nn;
                    g.w += (gn.w - (o.w / nn) * proj) / nn;
                } else {
                    g.x += gn.x / 1e-12f; g.y += gn.y / 1e-12f; g.z += gn.z / 1e-12f; g.w += gn.w / 1e-12f;
                }
                const float4 sa = __ldg(reinterpret_cast<const float4*>(s_act) + off);
                const float4 ta = __ldg(reinterpret_cast<const float4*>(t_act) + off);
                gs4 = make_float4(sa.x > 0.f ? g.x : 0.2f * g.x, sa.y > 0.f ? g.y : 0.2f * g.y,
                                  sa.z > 0.f ? g.z : 0.2f * g.z, sa.w > 0.f ? g.w : 0.2f * g.w);
                gt4 = make_float4(ta.x > 0.f ? g.x : 0.2f * g.x, ta.y > 0.f ? g.y : 0.2f * g.y,
                                  ta.z > 0.f ? g.z : 0.2f * g.z, ta.w > 0.f ? g.w : 0.2f * g.w);
                if (gs_out) reinterpret_cast<float4*>(gs_out)[off] = gs4;
                if (gt_out) reinterpret_cast<float4*>(gt_out)[off] = gt4;
            }
            *reinterpret_cast<float4*>(Gs + rr * NLD + 4 * tx) = gs4;
            *reinterpret_cast<float4*>(Gt + rr * NLD + 4 * tx) = gt4;
            *reinterpret_cast<float4*>(Ns + rr * NLD + 4 * tx) = n4;
            *reinterpret_cast<float4*>(Es + rr * NLD + 4 * tx) = e4;
        }
        __syncthreads();
        float gx1[4][4] = {}, gx2[4][4] = {};
        tile_mm(Gs, W1t, ty, tx, gx1);     // gx1[r][i] = sum_j gs[r][j] * (W1+b1)[i][j]
        tile_mm(Gt, W2t, ty, tx, gx2);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int rr = ty * 4 + r;
            const int64_t row = r0 + rr;
            if (row < n) {
                const int64_t off = row * (ND / 4) + tx;
                const float4 a = *reinterpret_cast<const float4*>(Ns + rr * NLD + 4 * tx);
                const float4 b = *reinterpret_cast<const float4*>(Es + rr * NLD + 4 * tx);
                reinterpret_cast<float4*>(g_nei)[off] =
                    make_float4(fmaf(gx2[r][0], b.x, gx1[r][0]), fmaf(gx2[r][1], b.y, gx1[r][1]),
                                fmaf(gx2[r][2], b.z, gx1[r][2]), fmaf(gx2[r][3], b.w, gx1[r][3]));
                reinterpret_cast<float4*>(g_e)[off] =
                    make_float4(fmaf(gx2[r][0], a.x, gx1[r][0]), fmaf(gx2[r][1], a.y, gx1[r][1]),
                                fmaf(gx2[r][2], a.z, gx1[r][2]), fmaf(gx2[r][3], a.w, gx1[r][3]));
            }
        }
        if (dw1) {      // outer products over the tile's rows (rows past n hold zeros)
#pragma unroll 4
            for (int r = 0; r < NR; ++r) {
                const float4 nn4 = *reinterpret_cast<const float4*>(Ns + r * NLD + 4 * ty);
                const float4 ee4 = *reinterpret_cast<const float4*>(Es + r * NLD + 4 * ty);
                const float4 gs4 = *reinterpret_cast<const float4*>(Gs + r * NLD + 4 * tx);
                const float4 gt4 = *reinterpret_cast<const float4*>(Gt + r * NLD + 4 * tx);
                const float x1[4] = {nn4.x + ee4.x, nn4.y + ee4.y, nn4.z + ee4.z, nn4.w + ee4.w};
                const float x2[4] = {nn4.x * ee4.x, nn4.y * ee4.y, nn4.z * ee4.z, nn4.w * ee4.w};
                const unsigned long long gs01 = pack2(gs4.x, gs4.y), gs23 = pack2(gs4.z, gs4.w);
                const unsigned long long gt01 = pack2(gt4.x, gt4.y), gt23 = pack2(gt4.z, gt4.w);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const unsigned long long x1d = pack2(x1[i], x1[i]), x2d = pack2(x2[i], x2[i]);
                    a1p[i][0] = ffma2(x1d, gs01, a1p[i][0]);
                    a1p[i][1] = ffma2(x1d, gs23, a1p[i][1]);
                    a2p[i][0] = ffma2(x2d, gt01, a2p[i][0]);
                    a2p[i][1] = ffma2(x2d, gt23, a2p[i][1]);
                }
            }
        }
    }
    if (dw1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 v1, v2;
            unpack2(a1p[i][0], v1.x, v1.y);
            unpack2(a1p[i][1], v1.z, v1.w);
            unpack2(a2p[i][0], v2.x, v2.y);
            unpack2(a2p[i][1], v2.z, v2.w);
            red_add4(reinterpret_cast<float4*>(dw1 + (ty * 4 + i) * ND + tx * 4), v1);
            red_add4(reinterpret_cast<float4*>(dw2 + (ty * 4 + i) * ND + tx * 4), v2);
        }
    }
}

constexpr size_t kNgcfSmem = (size_t)(2 * ND + 2 * NR) * NLD * sizeof(float);
constexpr size_t kNgcfSmemBwd = (size_t)(2 * ND + 4 * NR) * NLD * sizeof(float);

}  // namespace tagrec

using namespace tagrec;

extern "C" int tagrec_ngcf_dense_fwd(const float* nei, const float* e, const float* w1, const float* b1,
                                     const float* w2, const float* b2, int64_t n, int dim, float* out, float* nrm,
                                     float* s_act, float* t_act, void* stream) {
    TAGREC_REQUIRE(nei && e && w1 && b1 && w2 && b2 && out && nrm && s_act && t_act, "null pointer");
    TAGREC_REQUIRE(dim == ND, "the fused NGCF layer is built for 64 -> 64 layers");
    if (n == 0) return TAGREC_OK;
    TAGREC_CUDA(cudaFuncSetAttribute(ngcf_dense_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNgcfSmem));
    const int64_t tiles = (n + NR - 1) / NR;
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, 2 * kSMs);
    TAGREC_LAUNCH(ngcf_dense_fwd_kernel, grid, 256, kNgcfSmem, stream, nei, e, w1, b1, w2, b2, n, out, nrm, s_act, t_act);
    return TAGREC_OK;
}

extern "C" int tagrec_ngcf_dense_bwd(const float* g_out, const float* g_nrm, int64_t g_nrm_ld, const float* out,
                                     const float* s_act, const float* t_act, const float* nei, const float* e,
                                     const float* w1, const float* b1, const float* w2, const float* b2, int64_t n,
                                     int dim, float* g_nei, float* g_e, float* gs, float* gt, float* dw1, float* dw2,
                                     void* stream) {
    TAGREC_REQUIRE(g_nrm && out && s_act && t_act && nei && e && w1 && b1 && w2 && b2 && g_nei && g_e, "null pointer");
    TAGREC_REQUIRE((dw1 == nullptr) == (dw2 == nullptr), "dw1 and dw2 go together");
    TAGREC_REQUIRE(dim == ND, "the fused NGCF layer is built for 64 -> 64 layers");
    TAGREC_REQUIRE(g_nrm_ld >= ND && g_nrm_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(g_nrm) & 15) == 0,
                   "g_nrm rows must be 16-byte aligned");
    if (n == 0) return TAGREC_OK;
    TAGREC_CUDA(cudaFuncSetAttribute(ngcf_dense_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)kNgcfSmemBwd));
    const int64_t tiles = (n + NR - 1) / NR;
    const unsigned grid = (unsigned)std::min<int64_t>(tiles, 2 * kSMs);
    TAGREC_LAUNCH(ngcf_dense_bwd_kernel, grid, 256, kNgcfSmemBwd, stream, g_out, g_nrm, g_nrm_ld, out, s_act, t_act, nei,
                  e, w1, b1, w2, b2, n, g_nei, g_e, gs, gt, dw1, dw2);
    return TAGREC_OK;
}
