// K4 — TGCN neighbour attention over padded neighbour tables (gather - attend - reduce, and its scatter backward).
// sm_100a.
//
// Replaces model/tgcn.py:20-37 (Attention1.forward): the reference materialises [N, k, 64] + [N, k, 10] gathers,
// repeats the node row k times, runs two [N*k, .] x [., 32] matmuls, a softmax over k and a weighted sum — and its
// autograd scatters [N, k, 64] gradients back with index_put_(accumulate) (45 % of the reference's step).
//
// Algebra used here: the attention logits split into per-node, per-weight-id and per-neighbour 32-d projections
//     relu([e_v | e_w] W1 + e_j W2 + b) . v  =  relu(PV[v] + WW[w] + PJ[j]) . v,
//     PV = e_v W1[:64] + b,   WW = e_w W1[64:],   PJ = e_j W2            (three small dense GEMMs, done by the caller)
// so one gathered neighbour costs a 128 B PJ row for the logit and a 256 B embedding row for the weighted sum, and
// nothing of size [N, k, .] is ever written.  Index 0 of a table entry means "padding": zero projection, zero row,
// but it still takes part in the softmax (tgcn.py:21-24,35).
//
// Forward: one warp per node; lane d owns attention dim d (32 = dim_atten), then lane s owns slot s of the softmax
// (k <= 32), then lane l owns floats 2l, 2l+1 of the 64-d output.  Backward: same walk; gradients of the gathered
// rows leave through red.global.add (v2 for the 64-d rows), the tiny per-weight-id table is reduced in shared
// memory first (every node hits the same <= 64 rows).
#include "common.cuh"

namespace tagrec {

constexpr int AD = 32;        // dim_atten
constexpr int ED = 64;        // embedding dim
constexpr int MAXW = 64;      // weight ids staged in smem by the backward (edge multiplicities are small integers; larger
                              // ids — a user who applied one tag more than 64 times — add straight to global memory)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// v[f] per lane, f < 32  ->  on lane f: the sum over all lanes of their v[f]
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int o = 16 >> s;
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (i < o) {
                const float send = up ? v[i] : v[i + o];
                const float keep = up ? v[i + o] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
    }
    return v[0];
}
__device__ __forceinline__ void red_add2(float2* addr, float2 v) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x), "f"(v.y) : "memory");
}

__global__ void __launch_bounds__(256)
nbr_attention_fwd_kernel(const float* __restrict__ pv, const float* __restrict__ ww, const float* __restrict__ pj,
                         const float* __restrict__ ej, const float* __restrict__ vvec, const int64_t* __restrict__ nbr,
                         const int64_t* __restrict__ nbw, int64_t n, int k, int64_t ld, float* __restrict__ out,
                         float* __restrict__ att) {
    const int lane = threadIdx.x & 31;
    const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (v >= n) return;
    const float pvd = __ldg(pv + v * AD + lane);
    const float vd = __ldg(vvec + lane);
    int64_t jl = 0, wl = 0;                       // lane s holds slot s of the tables
    if (lane < k) {
        jl = __ldg(nbr + v * ld + lane);
        wl = __ldg(nbw + v * ld + lane);
    }
    // logit of slot s = sum over the 32 attention dims (lanes) of relu(h) v: the k per-lane partials are reduced with
    // ONE transposing butterfly (31 shuffles; lane s ends up with slot s) instead of k warp reductions of 5
    float part[32];
#pragma unroll
    for (int s = 0; s < 32; ++s) {
        part[s] = 0.f;
        if (s < k) {                                  // warp-uniform
            const int64_t j = __shfl_sync(0xffffffffu, jl, s);
            const int64_t w = __shfl_sync(0xffffffffu, wl, s);
            float h = pvd;
            if (w > 0) h += __ldg(ww + (w - 1) * AD + lane);
            if (j > 0) h += __ldg(pj + (j - 1) * AD + lane);
            part[s] = fmaxf(h, 0.f) * vd;
        }
    }
    const float x = warp_transpose_sum32(part, lane);
    const float logit = lane < k ? x : -INFINITY;
    const float m = warp_max(logit);
    const float ex = lane < k ? expf(logit - m) : 0.f;
    const float a = ex / warp_sum(ex);
    if (lane < k) att[v * k + lane] = a;
    float2 acc = make_float2(0.f, 0.f);
    for (int s = 0; s < k; ++s) {
        const int64_t j = __shfl_sync(0xffffffffu, jl, s);
        const float as = __shfl_sync(0xffffffffu, a, s);
        if (j > 0) {
            const float2 e = __ldg(reinterpret_cast<const float2*>(ej + (j - 1) * ED) + lane);
            acc.x = fmaf(as, e.x, acc.x);
            acc.y = fmaf(as, e.y, acc.y);
        }
    }
    reinterpret_cast<float2*>(out + v * ED)[lane] = acc;
}

__global__ void __launch_bounds__(256)
nbr_attention_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ att, const float* __restrict__ pv,
                         const float* __restrict__ ww, const float* __restrict__ pj, const float* __restrict__ ej,
                         const float* __restrict__ vvec, const int64_t* __restrict__ nbr,
                         const int64_t* __restrict__ nbw, int64_t n, int k, int64_t ld, int n_w,
                         float* __restrict__ g_pv, float* __restrict__ g_ww, float* __restrict__ g_pj,
                         float* __restrict__ g_ej, float* __restrict__ g_v) {
    __shared__ float s_ww[MAXW * AD];
    __shared__ float s_v[AD];
    const int n_stage = n_w < MAXW ? n_w : MAXW;
    for (int i = threadIdx.x; i < n_stage * AD; i += blockDim.x) s_ww[i] = 0.f;
    if (threadIdx.x < AD) s_v[threadIdx.x] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t v = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (v < n) {
        const float pvd = __ldg(pv + v * AD + lane);
        const float vd = __ldg(vvec + lane);
        const float2 go = __ldg(reinterpret_cast<const float2*>(g_out + v * ED) + lane);
        int64_t jl = 0, wl = 0;
        float a = 0.f;
        if (lane < k) {
            jl = __ldg(nbr + v * ld + lane);
            wl = __ldg(nbw + v * ld + lane);
            a = __ldg(att + v * k + lane);
        }
        // g_a[s] = <g_out[v], e_j(s)>;  scatter a_s * g_out[v] into g_ej
        float ga = 0.f;
        for (int s = 0; s < k; ++s) {
            const int64_t j = __shfl_sync(0xffffffffu, jl, s);
            const float as = __shfl_sync(0xffffffffu, a, s);
            float d = 0.f;
            if (j > 0) {
                const float2 e = __ldg(reinterpret_cast<const float2*>(ej + (j - 1) * ED) + lane);
                d = go.x * e.x + go.y * e.y;
                red_add2(reinterpret_cast<float2*>(g_ej + (j - 1) * ED) + lane, make_float2(as * go.x, as * go.y));
            }
            d = warp_sum(d);                          // (the one-butterfly form of the forward measured 15 % slower here)
            if (lane == s) ga = d;
        }
        const float dot = warp_sum(a * ga);
        const float gx = a * (ga - dot);                     // softmax backward; lane s holds slot s
        float gpv = 0.f, gv = 0.f;
        for (int s = 0; s < k; ++s) {
            const int64_t j = __shfl_sync(0xffffffffu, jl, s);
            const int64_t w = __shfl_sync(0xffffffffu, wl, s);
            const float gxs = __shfl_sync(0xffffffffu, gx, s);
            float h = pvd;
            if (w > 0) h += __ldg(ww + (w - 1) * AD + lane);
            if (j > 0) h += __ldg(pj + (j - 1) * AD + lane);
            gv = fmaf(gxs, fmaxf(h, 0.f), gv);
            const float gh = h > 0.f ? gxs * vd : 0.f;
            gpv += gh;
            if (w > 0) {
                if (w <= MAXW) atomicAdd(&s_ww[(w - 1) * AD + lane], gh);
                else atomicAdd(g_ww + (w - 1) * AD + lane, gh);
            }
            if (j > 0) atomicAdd(g_pj + (j - 1) * AD + lane, gh);
        }
        g_pv[v * AD + lane] = gpv;
        atomicAdd(&s_v[lane], gv);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_stage * AD; i += blockDim.x)
        if (s_ww[i] != 0.f) atomicAdd(g_ww + i, s_ww[i]);
    if (threadIdx.x < AD) atomicAdd(g_v + threadIdx.x, s_v[threadIdx.x]);
}

}  // namespace tagrec

using namespace tagrec;

extern "C" int tagrec_nbr_attention_fwd(const float* pv, const float* ww, const float* pj, const float* ej,
                                        const float* v, const int64_t* nbr, const int64_t* nbw, int64_t n, int k,
                                        int64_t ld, int dim, int dim_atten, float* out, float* att, void* stream) {
    TAGREC_REQUIRE(pv && ww && pj && ej && v && nbr && nbw && out && att, "null pointer");
    TAGREC_REQUIRE(dim == ED && dim_atten == AD, "neighbour attention is built for dim 64 / dim_atten 32");
    TAGREC_REQUIRE(k >= 1 && k <= 32 && ld >= k, "neighbor_k must be in 1..32");
    if (n == 0) return TAGREC_OK;
    TAGREC_LAUNCH(nbr_attention_fwd_kernel, (unsigned)((n + 7) / 8), 256, 0, stream, pv, ww, pj, ej, v, nbr, nbw, n, k, ld,
                  out, att);
    return TAGREC_OK;
}

extern "C" int tagrec_nbr_attention_bwd(const float* g_out, const float* att, const float* pv, const float* ww,
                                        const float* pj, const float* ej, const float* v, const int64_t* nbr,
                                        const int64_t* nbw, int64_t n, int k, int64_t ld, int n_w, int dim,
                                        int dim_atten, float* g_pv, float* g_ww, float* g_pj, float* g_ej, float* g_v,
                                        void* stream) {
    TAGREC_REQUIRE(g_out && att && pv && ww && pj && ej && v && nbr && nbw && g_pv && g_ww && g_pj && g_ej && g_v,
                   "null pointer");
    TAGREC_REQUIRE(dim == ED && dim_atten == AD, "neighbour attention is built for dim 64 / dim_atten 32");
    TAGREC_REQUIRE(k >= 1 && k <= 32 && ld >= k, "neighbor_k must be in 1..32");
    TAGREC_REQUIRE(n_w >= 1, "n_w (rows of the edge-weight embedding table) must be positive");
    if (n == 0) return TAGREC_OK;
    TAGREC_LAUNCH(nbr_attention_bwd_kernel, (unsigned)((n + 7) / 8), 256, 0, stream, g_out, att, pv, ww, pj, ej, v, nbr,
                  nbw, n, k, ld, n_w, g_pv, g_ww, g_pj, g_ej, g_v);
    return TAGREC_OK;
}
