// K7a — TGCN type-level attention + vector-level convolution, per node (forward and backward).  sm_100a.
//
// Replaces model/tgcn.py:78-84 (BasicLayer._atten2: stack, [N,3,64] x [64,32] matmul, relu, matmul, softmax over the
// three node types, scale) and the vector-level branch of _conv (tgcn.py:92-98: Conv2d(1 -> V, (j, 64)), j = 1..3,
// relu, concat) together with their autograd graphs — about 40 small torch kernels and 18 skinny SIMT GEMMs (output
// width 8) per layer in the first version of the TGCN path.  Both are per-node maps of a [3, 64] stack:
//     h_r = relu(x_r U + q),  b = softmax_r(h_r . p),  z_r = b_r x_r                      (r = user, item, tag slot)
//     y_j[ch, pos] = relu(sum_{r<j, d} w_j[ch, r*64 + d] z_{pos+r}[d])                       (pos = 0 .. 3-j)
// One warp per node; lane l owns attention dim l (its column of U lives in 64 registers) and embedding dims 2l, 2l+1.
// The 6V conv outputs are reduced across lanes with a transposing butterfly (31 shuffles per 32 outputs).
// Outputs: z [n,3,64] and xf [n,6V] — the inputs of K7 (tgcn_tail.cu).  The backward takes K7's g_z and g_xf and
// produces the gradients of the three input tables and of U, q, p, w_1..3 (per-CTA shared-memory sums, then atomics).
#include <algorithm>

#include "common.cuh"

namespace tagrec {

constexpr int MW = 64;            // embedding width
constexpr int MA = 32;            // dim_atten

__device__ __forceinline__ float mix_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// v[f] per lane, f < 32  ->  on lane f: the sum over all lanes of their v[f]
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int o = 16 >> s;                      // also the number of values that survive this round
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

// compile-time decode of vector-level feature f: conv j (1..3), channel, position; features are ordered conv_1
// (channel-major, 3 positions), conv_2 (2 positions), conv_3 (1 position) — the reshape/cat order of tgcn.py:95-98
template <int V>
struct VecFeat {
    __host__ __device__ static constexpr int j(int f) { return f < 3 * V ? 1 : (f < 5 * V ? 2 : 3); }
    __host__ __device__ static constexpr int ch(int f) { return f < 3 * V ? f / 3 : (f < 5 * V ? (f - 3 * V) / 2 : f - 5 * V); }
    __host__ __device__ static constexpr int pos(int f) { return f < 3 * V ? f % 3 : (f < 5 * V ? (f - 3 * V) % 2 : 0); }
};

template <int V>
struct MixSmem {
    float w1[V * MW];
    float w2[V * 2 * MW];
    float w3[V * 3 * MW];
    __device__ float* w(int j) { return j == 1 ? w1 : (j == 2 ? w2 : w3); }
};

// h_r[lane], logits and softmax weights of one node; xs = this warp's staged [3][64] rows; ucol(d) = U[d][lane]
template <typename UCol>
__device__ __forceinline__ void type_attention(const float* __restrict__ xs, UCol ucol, float ql, float pl,
                                               float (&h)[3], float (&b)[3]) {
    float lg[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float a0 = ql, a1 = 0.f, a2 = 0.f, a3 = 0.f;    // four independent chains (FMA latency, few warps per SM)
#pragma unroll
        for (int d4 = 0; d4 < MW / 4; ++d4) {
            const float4 xv = *reinterpret_cast<const float4*>(xs + r * MW + 4 * d4);
            a0 = fmaf(xv.x, ucol(4 * d4 + 0), a0);
            a1 = fmaf(xv.y, ucol(4 * d4 + 1), a1);
            a2 = fmaf(xv.z, ucol(4 * d4 + 2), a2);
            a3 = fmaf(xv.w, ucol(4 * d4 + 3), a3);
        }
        const float a = (a0 + a1) + (a2 + a3);
        h[r] = a;
        lg[r] = mix_warp_sum(fmaxf(a, 0.f) * pl);
    }
    const float m = fmaxf(lg[0], fmaxf(lg[1], lg[2]));
    const float e0 = expf(lg[0] - m), e1 = expf(lg[1] - m), e2 = expf(lg[2] - m);
    const float inv = 1.f / (e0 + e1 + e2);
    b[0] = e0 * inv;
    b[1] = e1 * inv;
    b[2] = e2 * inv;
}

template <int V>
__global__ void __launch_bounds__(256)
tgcn_mix_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ x2,
                    const float* __restrict__ U, const float* __restrict__ q, const float* __restrict__ p,
                    const float* __restrict__ wv1, const float* __restrict__ wv2, const float* __restrict__ wv3,
                    int64_t n, float* __restrict__ z, float* __restrict__ xf) {
    constexpr int NF = 6 * V;
    __shared__ MixSmem<V> ws;
    __shared__ __align__(16) float XS[8][3 * MW];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < V * MW; i += 256) ws.w1[i] = __ldg(wv1 + i);
    for (int i = tid; i < V * 2 * MW; i += 256) ws.w2[i] = __ldg(wv2 + i);
    for (int i = tid; i < V * 3 * MW; i += 256) ws.w3[i] = __ldg(wv3 + i);
    float ucol[MW];
#pragma unroll
    for (int d = 0; d < MW; ++d) ucol[d] = __ldg(U + d * MA + lane);
    const float ql = __ldg(q + lane), pl = __ldg(p + lane);
    __syncthreads();
    const float* xr[3] = {x0, x1, x2};
    float* xs = XS[warp];
    for (int64_t node = (int64_t)blockIdx.x * 8 + warp; node < n; node += (int64_t)gridDim.x * 8) {
        float2 x[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) x[r] = __ldg(reinterpret_cast<const float2*>(xr[r] + node * MW) + lane);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 3; ++r) *reinterpret_cast<float2*>(xs + r * MW + 2 * lane) = x[r];
        __syncwarp();
        float h[3], b[3];
        type_attention(xs, [&](int d) { return ucol[d]; }, ql, pl, h, b);
        float2 zz[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            zz[r] = make_float2(b[r] * x[r].x, b[r] * x[r].y);
            *reinterpret_cast<float2*>(z + (node * 3 + r) * MW + 2 * lane) = zz[r];
        }
        float v[32];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const int j = VecFeat<V>::j(f), ch = VecFeat<V>::ch(f), ps = VecFeat<V>::pos(f);
            const float* wrow = ws.w(j) + ch * j * MW;
            float part = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                if (r < j) {
                    const float2 w = *reinterpret_cast<const float2*>(wrow + r * MW + 2 * lane);
                    part = fmaf(w.x, zz[ps + r].x, part);
                    part = fmaf(w.y, zz[ps + r].y, part);
                }
            }
            v[f & 31] = part;
            if ((f & 31) == 31 || f == NF - 1) {
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (k > (f & 31)) v[k] = 0.f;
                const float s = warp_transpose_sum(v, lane);
                const int base = f & ~31;
                if (base + lane < NF) xf[node * NF + base + lane] = fmaxf(s, 0.f);
            }
        }
    }
}

template <int V>
__global__ void __launch_bounds__(256, 2)
tgcn_mix_bwd_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ x2,
                    const float* __restrict__ U, const float* __restrict__ q, const float* __restrict__ p,
                    const float* __restrict__ wv1, const float* __restrict__ wv2, const float* __restrict__ wv3,
                    int64_t n, const float* __restrict__ g_z, const float* __restrict__ g_xf,
                    const float* __restrict__ xf, float* __restrict__ g_x0, float* __restrict__ g_x1,
                    float* __restrict__ g_x2, float* __restrict__ gh_out, float* __restrict__ g_q, float* __restrict__ g_p) {
    constexpr int NF = 6 * V;
    __shared__ MixSmem<V> ws;                      // weights
    __shared__ __align__(16) float UT[MA * MW];    // UT[l][d] = U[d][l]
    __shared__ float US[MW * MA];                  // U itself: lane l reads column l conflict-free
    __shared__ __align__(16) float XS[8][3 * MW];
    __shared__ float GQP[2 * MA];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < V * MW; i += 256) ws.w1[i] = __ldg(wv1 + i);
    for (int i = tid; i < V * 2 * MW; i += 256) ws.w2[i] = __ldg(wv2 + i);
    for (int i = tid; i < V * 3 * MW; i += 256) ws.w3[i] = __ldg(wv3 + i);
    for (int i = tid; i < MW * MA; i += 256) {
        const float u = __ldg(U + i);
        UT[(i % MA) * MW + i / MA] = u;
        US[i] = u;
    }
    if (tid < 2 * MA) GQP[tid] = 0.f;
    const float ql = __ldg(q + lane), pl = __ldg(p + lane);
    float gq = 0.f, gp = 0.f;
    __syncthreads();
    const float* xr[3] = {x0, x1, x2};
    float* gxr[3] = {g_x0, g_x1, g_x2};
    float* xs = XS[warp];
    for (int64_t node = (int64_t)blockIdx.x * 8 + warp; node < n; node += (int64_t)gridDim.x * 8) {
        float2 x[3], gz[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            x[r] = __ldg(reinterpret_cast<const float2*>(xr[r] + node * MW) + lane);
            gz[r] = __ldg(reinterpret_cast<const float2*>(g_z + (node * 3 + r) * MW) + lane);
        }
        float gy[2] = {0.f, 0.f};                   // lane holds the masked gradient of features lane and 32 + lane
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int f = 32 * k + lane;
            if (f < NF) {
                const float y = __ldg(xf + node * NF + f);
                gy[k] = y > 0.f ? __ldg(g_xf + node * NF + f) : 0.f;
            }
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 3; ++r) *reinterpret_cast<float2*>(xs + r * MW + 2 * lane) = x[r];
        __syncwarp();
        float h[3], b[3];
        type_attention(xs, [&](int d) { return US[d * MA + lane]; }, ql, pl, h, b);
        // vector-level conv, transposed: into g_z (the weight gradients are T5's, below); one channel of each conv per
        // iteration (rolled: the fully unrolled form spills)
        auto gyf = [&](int f) { return __shfl_sync(0xffffffffu, f < 32 ? gy[0] : gy[1], f & 31); };
#pragma unroll 1
        for (int ch = 0; ch < V; ++ch) {
#pragma unroll
            for (int ps = 0; ps < 3; ++ps) {                       // conv_1: kernel (1, 64), 3 positions
                const float g = gyf(3 * ch + ps);
                const float2 w = *reinterpret_cast<const float2*>(ws.w1 + ch * MW + 2 * lane);
                gz[ps].x = fmaf(w.x, g, gz[ps].x);
                gz[ps].y = fmaf(w.y, g, gz[ps].y);
            }
#pragma unroll
            for (int ps = 0; ps < 2; ++ps) {                       // conv_2: kernel (2, 64), 2 positions
                const float g = gyf(3 * V + 2 * ch + ps);
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const float2 w = *reinterpret_cast<const float2*>(ws.w2 + (ch * 2 + r) * MW + 2 * lane);
                    gz[ps + r].x = fmaf(w.x, g, gz[ps + r].x);
                    gz[ps + r].y = fmaf(w.y, g, gz[ps + r].y);
                }
            }
            const float g3 = gyf(5 * V + ch);                      // conv_3: kernel (3, 64), 1 position
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float2 w = *reinterpret_cast<const float2*>(ws.w3 + (ch * 3 + r) * MW + 2 * lane);
                gz[r].x = fmaf(w.x, g3, gz[r].x);
                gz[r].y = fmaf(w.y, g3, gz[r].y);
            }
        }
        // type-level attention backward
        float gb[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) gb[r] = mix_warp_sum(gz[r].x * x[r].x + gz[r].y * x[r].y);
        const float s = b[0] * gb[0] + b[1] * gb[1] + b[2] * gb[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float gl = b[r] * (gb[r] - s);    // gradient of logit r
            const float gh = h[r] > 0.f ? gl * pl : 0.f;
            gp = fmaf(gl, fmaxf(h[r], 0.f), gp);
            gq += gh;
            float2 gx = make_float2(b[r] * gz[r].x, b[r] * gz[r].y);
            gh_out[((int64_t)r * n + node) * MA + lane] = gh;        // g_U = sum_r x_r^T gh_r runs on K8 afterwards
            float2 gx1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int l = 0; l < MA; l += 2) {
                const float gh0 = __shfl_sync(0xffffffffu, gh, l), gh1 = __shfl_sync(0xffffffffu, gh, l + 1);
                const float2 u0 = *reinterpret_cast<const float2*>(UT + l * MW + 2 * lane);
                const float2 u1 = *reinterpret_cast<const float2*>(UT + (l + 1) * MW + 2 * lane);
                gx.x = fmaf(gh0, u0.x, gx.x);
                gx.y = fmaf(gh0, u0.y, gx.y);
                gx1.x = fmaf(gh1, u1.x, gx1.x);
                gx1.y = fmaf(gh1, u1.y, gx1.y);
            }
            *reinterpret_cast<float2*>(gxr[r] + node * MW + 2 * lane) = make_float2(gx.x + gx1.x, gx.y + gx1.y);
        }
    }
    // per-CTA sums, then one atomic per parameter element and CTA
    atomicAdd(&GQP[lane], gq);
    atomicAdd(&GQP[MA + lane], gp);
    __syncthreads();
    if (tid < MA) {
        atomicAdd(g_q + tid, GQP[tid]);
        atomicAdd(g_p + tid, GQP[MA + tid]);
    }
}

// T5: gradients of the vector-level conv weights, a row reduction over the nodes:
//     g_w_j[ch][k] = sum_node sum_pos gy[node][f(j, ch, pos)] * zflat[node][pos*64 + k],    gy = g_xf * (xf > 0)
// Thread c owns column k of conv j for all V channels (64 + 128 + 192 = 384 columns); tiles of 32 nodes are staged in
// shared memory (shared-memory float atomics are CAS loops on this architecture — the first version of the backward,
// which accumulated these sums with them, spent most of its time there).
template <int V>
__global__ void __launch_bounds__(384)
tgcn_vecw_kernel(const float* __restrict__ z, const float* __restrict__ g_xf, const float* __restrict__ xf, int64_t n,
                 float* __restrict__ g_wv1, float* __restrict__ g_wv2, float* __restrict__ g_wv3) {
    constexpr int NF = 6 * V, T = 32;
    __shared__ __align__(16) float Zs[T][3 * MW];
    __shared__ __align__(16) float Gs[T][NF];
    const int c = threadIdx.x;
    const int j = c < MW ? 1 : (c < 3 * MW ? 2 : 3);
    const int k = c < MW ? c : (c < 3 * MW ? c - MW : c - 3 * MW);
    float acc[V];
#pragma unroll
    for (int ch = 0; ch < V; ++ch) acc[ch] = 0.f;
    const int64_t n_chunks = (n + T - 1) / T;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t n0 = chunk * T;
        __syncthreads();
        for (int idx = c; idx < T * (3 * MW / 4); idx += 384) {
            const int node = idx / (3 * MW / 4), q4 = idx % (3 * MW / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + node < n) v = __ldg(reinterpret_cast<const float4*>(z + (n0 + node) * 3 * MW) + q4);
            *reinterpret_cast<float4*>(&Zs[node][4 * q4]) = v;
        }
        for (int idx = c; idx < T * (NF / 4); idx += 384) {
            const int node = idx / (NF / 4), q4 = idx % (NF / 4);
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + node < n) {
                g = __ldg(reinterpret_cast<const float4*>(g_xf + (n0 + node) * NF) + q4);
                const float4 y = __ldg(reinterpret_cast<const float4*>(xf + (n0 + node) * NF) + q4);
                g.x = y.x > 0.f ? g.x : 0.f;
                g.y = y.y > 0.f ? g.y : 0.f;
                g.z = y.z > 0.f ? g.z : 0.f;
                g.w = y.w > 0.f ? g.w : 0.f;
            }
            *reinterpret_cast<float4*>(&Gs[node][4 * q4]) = g;
        }
        __syncthreads();
        if (j == 1) {
#pragma unroll 4
            for (int node = 0; node < T; ++node) {
                const float z0 = Zs[node][k], z1 = Zs[node][MW + k], z2 = Zs[node][2 * MW + k];
#pragma unroll
                for (int ch = 0; ch < V; ++ch)
                    acc[ch] = fmaf(Gs[node][3 * ch + 2], z2, fmaf(Gs[node][3 * ch + 1], z1, fmaf(Gs[node][3 * ch], z0, acc[ch])));
            }
        } else if (j == 2) {
#pragma unroll 4
            for (int node = 0; node < T; ++node) {
                const float z0 = Zs[node][k], z1 = Zs[node][MW + k];
#pragma unroll
                for (int ch = 0; ch < V; ++ch)
                    acc[ch] = fmaf(Gs[node][3 * V + 2 * ch + 1], z1, fmaf(Gs[node][3 * V + 2 * ch], z0, acc[ch]));
            }
        } else {
#pragma unroll 4
            for (int node = 0; node < T; ++node) {
                const float z0 = Zs[node][k];
#pragma unroll
                for (int ch = 0; ch < V; ++ch) acc[ch] = fmaf(Gs[node][5 * V + ch], z0, acc[ch]);
            }
        }
    }
    float* out = j == 1 ? g_wv1 : (j == 2 ? g_wv2 : g_wv3);
#pragma unroll
    for (int ch = 0; ch < V; ++ch)
        if (acc[ch] != 0.f) atomicAdd(out + ch * j * MW + k, acc[ch]);
}

static int mix_check(int64_t n, int dim, int dim_atten, int V) {
    TAGREC_REQUIRE(dim == MW && dim_atten == MA, "the TGCN type attention is built for dim 64 / dim_atten 32");
    TAGREC_REQUIRE(V == 4 || V == 8, "num_vec_conv must be 4 or 8");
    TAGREC_REQUIRE(n >= 0, "negative row count");
    return TAGREC_OK;
}

}  // namespace tagrec

using namespace tagrec;

extern "C" int tagrec_tgcn_mix_fwd(const float* x0, const float* x1, const float* x2, const float* U, const float* q,
                                   const float* p, const float* wv1, const float* wv2, const float* wv3, int64_t n,
                                   int dim, int dim_atten, int n_vec_conv, float* z, float* xf, void* stream) {
    TAGREC_REQUIRE(x0 && x1 && x2 && U && q && p && wv1 && wv2 && wv3 && z && xf, "null pointer");
    if (int rc = mix_check(n, dim, dim_atten, n_vec_conv)) return rc;
    if (n == 0) return TAGREC_OK;
    const unsigned grid = (unsigned)std::min<int64_t>((n + 7) / 8, (int64_t)kSMs * 8);
    if (n_vec_conv == 8)
        TAGREC_LAUNCH(tgcn_mix_fwd_kernel<8>, grid, 256, 0, stream, x0, x1, x2, U, q, p, wv1, wv2, wv3, n, z, xf);
    else
        TAGREC_LAUNCH(tgcn_mix_fwd_kernel<4>, grid, 256, 0, stream, x0, x1, x2, U, q, p, wv1, wv2, wv3, n, z, xf);
    return TAGREC_OK;
}

extern "C" size_t tagrec_tgcn_mix_workspace_bytes(int64_t n) { return (size_t)3 * n * MA * 4 + 256; }

extern "C" int tagrec_tgcn_mix_bwd(const float* x0, const float* x1, const float* x2, const float* U, const float* q,
                                   const float* p, const float* wv1, const float* wv2, const float* wv3, int64_t n,
                                   int dim, int dim_atten, int n_vec_conv, const float* z, const float* g_z,
                                   const float* g_xf, const float* xf, float* g_x0, float* g_x1, float* g_x2, float* g_U, float* g_q,
                                   float* g_p, float* g_wv1, float* g_wv2, float* g_wv3, void* workspace,
                                   size_t workspace_bytes, void* stream) {
    TAGREC_REQUIRE(x0 && x1 && x2 && U && q && p && wv1 && wv2 && wv3 && z && g_z && g_xf && xf && g_x0 && g_x1 && g_x2 &&
                       g_U && g_q && g_p && g_wv1 && g_wv2 && g_wv3, "null pointer");
    if (int rc = mix_check(n, dim, dim_atten, n_vec_conv)) return rc;
    if (n == 0) return TAGREC_OK;
    if (!workspace || workspace_bytes < tagrec_tgcn_mix_workspace_bytes(n))
        return fail(TAGREC_ENOMEM, "tgcn mix workspace too small", __FILE__, __LINE__);
    float* gh = reinterpret_cast<float*>(workspace);                  // [3][n][32]
    const unsigned grid = (unsigned)std::min<int64_t>((n + 7) / 8, (int64_t)kSMs * 2);
    const unsigned grid_w = (unsigned)std::min<int64_t>((n + 31) / 32, (int64_t)kSMs * 2);
    if (n_vec_conv == 8) {
        TAGREC_LAUNCH(tgcn_mix_bwd_kernel<8>, grid, 256, 0, stream, x0, x1, x2, U, q, p, wv1, wv2, wv3, n, g_z, g_xf, xf,
                      g_x0, g_x1, g_x2, gh, g_q, g_p);
        TAGREC_LAUNCH(tgcn_vecw_kernel<8>, grid_w, 384, 0, stream, z, g_xf, xf, n, g_wv1, g_wv2, g_wv3);
    } else {
        TAGREC_LAUNCH(tgcn_mix_bwd_kernel<4>, grid, 256, 0, stream, x0, x1, x2, U, q, p, wv1, wv2, wv3, n, g_z, g_xf, xf,
                      g_x0, g_x1, g_x2, gh, g_q, g_p);
        TAGREC_LAUNCH(tgcn_vecw_kernel<4>, grid_w, 384, 0, stream, z, g_xf, xf, n, g_wv1, g_wv2, g_wv3);
    }
    const float* xr[3] = {x0, x1, x2};
    for (int r = 0; r < 3; ++r)
        if (int rc = tagrec_xty_acc(xr[r], gh + (size_t)r * n * MA, n, MW, MA, g_U, 1, stream)) return rc;
    return TAGREC_OK;
}
