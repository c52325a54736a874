// K7 — TGCN dense tail: bit-level convolution + fusion layer of BasicLayer, fused (forward and backward).  sm_100a.
//
// Replaces model/tgcn.py:86-106 (BasicLayer._conv / _fusion) for the bit-level branch: the reference materialises
//     bit_e = relu(Conv2d(1 -> C, (3,1))(z))      [N, C*64]   (C = num_bit_conv = 32: 8 KB per node)
//     y     = cat([bit_e, vec_e])                  [N, C*64 + E]   (E = 6 * num_vec_conv = 48 vector-level features)
//     out   = relu(y Wf + bf)                      [N, 64]
// i.e. an [N, 2096] fp32 matrix that is written, rectified, concatenated, read by the GEMM and kept for autograd —
// 0.94 GB per layer at N = 112 K, re-read four more times by the backward (93 % of the TGCN step was these passes plus
// the SIMT GEMMs around them, profiles/r1_models_launches.md).  Here the 2096 features never leave the SM:
//     feature (c, d) of node v  =  relu(wb[c,0] z[v,0,d] + wb[c,1] z[v,1,d] + wb[c,2] z[v,2,d])
// is generated from the [3, 64] stack z[v] (the output of the type-level attention) into shared memory, one 64-feature
// chunk (= one conv channel) at a time, and consumed by a 128-node x 64-output fp32 register-tile GEMM against the
// matching 64 rows of Wf.  The E vector-level features (2 % of the width) are computed by the caller and enter as
// one more chunk read from global memory (`xf`).
//
//   T1 tgcn_tail_fwd_kernel     out = relu([bit(z) | xf] Wf + bf)                               2*N*F*64 flop, F = C*64+E
//   T2 tgcn_tail_bwd_z_kernel   g_pre = g_out * (out > 0);  per chunk G = g_pre Wf_c^T, masked by the recomputed
//                               pre-activation -> g_z, g_wb (bit-conv weight), g_xf, g_bf      same flop count
//   T3 tgcn_tail_bwd_w_kernel   g_Wf = [bit(z) | xf]^T g_pre  (features regenerated, 4 chunks per CTA)   same flop count
//
// All arithmetic fp32 FMA (the reference computes this path in fp32; no TF32).  Bound: fp32 FMA issue (the tile GEMM
// is 32 FMA per 3 LDS.128 per thread); algorithmic HBM bytes per node are 768 (z) + 4 E (xf) + 256 (out) only.
#include <algorithm>

#include "common.cuh"

namespace tagrec {

constexpr int TW = 64;            // layer width (in_features == out_features == 64)
constexpr int TM = 128;           // nodes per tile
constexpr int TP = TM + 4;        // pitch (floats) of the feature-major tiles X[k][node]
constexpr int T3C = 4;            // chunks per CTA in T3
constexpr int ZNP = 3 * TW + 4;   // pitch (floats) of the node-major z tile of T2

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// pre-activation of one bit-level feature; the SAME expression in all three kernels (the backward's relu mask must
// agree with the forward's bit for bit)
__device__ __forceinline__ float bit_pre(float w0, float w1, float w2, float z0, float z1, float z2) {
    return fmaf(w2, z2, fmaf(w1, z1, w0 * z0));
}

// The tile GEMMs below were issue-bound on scalar FFMA (ncu: issue slots 65 % busy, FMA pipe 55 %,
// profiles/r1_tgcn_tail_ncu.md): they use the packed FMAs of common.cuh (pack2 / ffma2).
// acc[i][j] += sum_k Af[k][8 ty + i] * Wk[k][4 tx + j]      (Af feature-major pitch TP, Wk row-major pitch TW)
// p[ip][j] holds the pair (acc[2 ip][j], acc[2 ip + 1][j]): the node pairs come straight out of the float4 loads of Af,
// the weight is duplicated into both halves (4 moves per 16 FFMA2).
__device__ __forceinline__ void tile_gemm_step(const float* __restrict__ Af, const float* __restrict__ Wk, int k, int ty,
                                               int tx, unsigned long long (&p)[4][4]) {
    const float4 a0 = *reinterpret_cast<const float4*>(Af + k * TP + 8 * ty);
    const float4 a1 = *reinterpret_cast<const float4*>(Af + k * TP + 8 * ty + 4);
    const float4 w = *reinterpret_cast<const float4*>(Wk + k * TW + 4 * tx);
    const unsigned long long ap[4] = {pack2(a0.x, a0.y), pack2(a0.z, a0.w), pack2(a1.x, a1.y), pack2(a1.z, a1.w)};
    const unsigned long long wd[4] = {pack2(w.x, w.x), pack2(w.y, w.y), pack2(w.z, w.z), pack2(w.w, w.w)};
#pragma unroll
    for (int ip = 0; ip < 4; ++ip)
#pragma unroll
        for (int j = 0; j < 4; ++j) p[ip][j] = ffma2(ap[ip], wd[j], p[ip][j]);
}

__device__ __forceinline__ void tile_gemm(const float* __restrict__ Af, const float* __restrict__ Wk, int kc, int ty,
                                          int tx, float (&acc)[8][4]) {
    unsigned long long p[4][4];
#pragma unroll
    for (int ip = 0; ip < 4; ++ip)
#pragma unroll
        for (int j = 0; j < 4; ++j) p[ip][j] = pack2(acc[2 * ip][j], acc[2 * ip + 1][j]);
    if (kc == TW) {                                 // the 64-feature chunks: compile-time trip count
#pragma unroll 16
        for (int k = 0; k < TW; ++k) tile_gemm_step(Af, Wk, k, ty, tx, p);
    } else {
#pragma unroll 4
        for (int k = 0; k < kc; ++k) tile_gemm_step(Af, Wk, k, ty, tx, p);
    }
#pragma unroll
    for (int ip = 0; ip < 4; ++ip)
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack2(p[ip][j], acc[2 * ip][j], acc[2 * ip + 1][j]);
}

// rows [n0, n0 + TM) of a row-major [n, width] table -> feature-major tile X[f][node] (zero rows past n)
__device__ __forceinline__ void load_tile_fm(float* __restrict__ X, const float* __restrict__ src, int64_t n0, int64_t n,
                                             int width, int tid) {
    const int w4 = width >> 2;
    for (int idx = tid; idx < TM * w4; idx += 256) {
        const int node = idx & (TM - 1), c4 = idx >> 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n0 + node < n) v = __ldg(reinterpret_cast<const float4*>(src + (n0 + node) * width) + c4);
        X[(4 * c4 + 0) * TP + node] = v.x;
        X[(4 * c4 + 1) * TP + node] = v.y;
        X[(4 * c4 + 2) * TP + node] = v.z;
        X[(4 * c4 + 3) * TP + node] = v.w;
    }
}

// Feature generation: thread `tid` produces, for EVERY conv channel, the same 8 groups of 4 nodes x 1 dim — so its 24
// float4 of z are chunk-invariant and live in registers (the first version re-read them from shared memory for each
// of the 32 chunks: half again as many shared-memory wavefronts as the GEMM itself).
struct ZRegs {
    float4 v[8][3];
};
__device__ __forceinline__ void load_zregs(ZRegs& zr, const float* __restrict__ Z, int tid) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int idx = tid + 256 * it, d = idx >> 5, n4 = (idx & 31) * 4;
#pragma unroll
        for (int r = 0; r < 3; ++r) zr.v[it][r] = *reinterpret_cast<const float4*>(Z + (r * TW + d) * TP + n4);
    }
}
// A[d][node] = relu(bit_pre(wb[c], z[., d][node]))  for the 64 features of conv channel c
__device__ __forceinline__ void gen_bit_chunk(float* __restrict__ A, const ZRegs& zr, const float* WB, int c, int tid) {
    const float w0 = WB[3 * c], w1 = WB[3 * c + 1], w2 = WB[3 * c + 2];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int idx = tid + 256 * it, d = idx >> 5, n4 = (idx & 31) * 4;
        const float4 z0 = zr.v[it][0], z1 = zr.v[it][1], z2 = zr.v[it][2];
        float4 r;
        r.x = fmaxf(bit_pre(w0, w1, w2, z0.x, z1.x, z2.x), 0.f);
        r.y = fmaxf(bit_pre(w0, w1, w2, z0.y, z1.y, z2.y), 0.f);
        r.z = fmaxf(bit_pre(w0, w1, w2, z0.z, z1.z, z2.z), 0.f);
        r.w = fmaxf(bit_pre(w0, w1, w2, z0.w, z1.w, z2.w), 0.f);
        *reinterpret_cast<float4*>(A + d * TP + n4) = r;
    }
}

// rows [row0, row0 + rows) of a row-major [., 64] matrix -> smem W[rows][64], asynchronously
__device__ __forceinline__ void load_w_async(float* __restrict__ W, const float* __restrict__ src, int rows, int tid) {
    for (int idx = tid; idx < rows * (TW / 4); idx += 256) cp_async16(W + idx * 4, src + (size_t)idx * 4);
}

// ------------------------------------------------------------------------------------------------ T1 forward
__global__ void __launch_bounds__(256, 1)
tgcn_tail_fwd_kernel(const float* __restrict__ z, const float* __restrict__ wb, const float* __restrict__ xf,
                     const float* __restrict__ wf, const float* __restrict__ bf, int64_t n, int C, int E,
                     float* __restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    float* Z = sm;                                  // [192][TP]
    float* A = Z + 3 * TW * TP;                     // [2][64][TP]
    float* W = A + 2 * TW * TP;                     // [2][64][64]
    float* WB = W + 2 * TW * TW;                    // [C][3]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int64_t n0 = (int64_t)blockIdx.x * TM;
    const int NC = C + (E > 0 ? 1 : 0);
    for (int i = tid; i < 3 * C; i += 256) WB[i] = __ldg(wb + i);
    load_tile_fm(Z, z, n0, n, 3 * TW, tid);
    load_w_async(W, wf, C > 0 ? TW : E, tid);
    __syncthreads();
    ZRegs zr;
    load_zregs(zr, Z, tid);
    if (C > 0) gen_bit_chunk(A, zr, WB, 0, tid); else load_tile_fm(A, xf, n0, n, E, tid);
    cp_async_wait_all();
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < NC; ++c) {
        float* An = A + ((c + 1) & 1) * TW * TP;
        float* Wn = W + ((c + 1) & 1) * TW * TW;
        if (c + 1 < NC) {                           // next chunk: weights in flight, features generated
            load_w_async(Wn, wf + (size_t)(c + 1) * TW * TW, c + 1 < C ? TW : E, tid);
            if (c + 1 < C) gen_bit_chunk(An, zr, WB, c + 1, tid); else load_tile_fm(An, xf, n0, n, E, tid);
        }
        tile_gemm(A + (c & 1) * TW * TP, W + (c & 1) * TW * TW, c < C ? TW : E, ty, tx, acc);
        cp_async_wait_all();
        __syncthreads();
    }
    const float4 b = __ldg(reinterpret_cast<const float4*>(bf) + tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t node = n0 + 8 * ty + i;
        if (node >= n) continue;
        float4 o;
        o.x = fmaxf(acc[i][0] + b.x, 0.f);
        o.y = fmaxf(acc[i][1] + b.y, 0.f);
        o.z = fmaxf(acc[i][2] + b.z, 0.f);
        o.w = fmaxf(acc[i][3] + b.w, 0.f);
        *reinterpret_cast<float4*>(out + node * TW + 4 * tx) = o;
    }
}

// wft[c][o][d] = Wf[c*64 + d][o]  (rows past the table: 0) — the chunks of Wf transposed, so that T2 runs the same
// tile GEMM with k = output column
__global__ void tgcn_tail_transpose_kernel(const float* __restrict__ wf, int C, int E, float* __restrict__ wft) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int NC = C + (E > 0 ? 1 : 0);
    if (idx >= NC * TW * TW) return;
    const int c = idx / (TW * TW), o = (idx / TW) % TW, d = idx % TW;
    const bool ok = c < C || d < E;
    wft[idx] = ok ? __ldg(wf + ((size_t)c * TW + d) * TW + o) : 0.f;
}

// ------------------------------------------------------------------------------------------------ T2 backward to z
__global__ void __launch_bounds__(256, 1)
tgcn_tail_bwd_z_kernel(const float* __restrict__ g_out, const float* __restrict__ out, const float* __restrict__ z,
                       const float* __restrict__ wb, const float* __restrict__ wft, int64_t n, int C, int E,
                       float* __restrict__ g_pre, float* __restrict__ g_z, float* __restrict__ g_wb,
                       float* __restrict__ g_xf, float* __restrict__ g_bf) {
    extern __shared__ __align__(16) float sm[];
    float* Z = sm;                                  // [TM][ZNP]  z, NODE-major: the mask stage reads float4 over 4 dims
    float* GP = Z + TM * ZNP;                       // [64][TP]   g_pre, feature-major
    float* W = GP + TW * TP;                        // [2][64][64]
    float* WB = W + 2 * TW * TW;                    // [C][3]
    float* GWB = WB + 3 * C;                        // [C][3]  per-CTA sums
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15, lane = tid & 31;
    const int NC = C + (E > 0 ? 1 : 0);
    const int64_t n_tiles = (n + TM - 1) / TM;
    for (int i = tid; i < 3 * C; i += 256) {
        WB[i] = __ldg(wb + i);
        GWB[i] = 0.f;
    }
    float bf_acc = 0.f;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n0 = tile * TM;
        __syncthreads();                            // the previous tile's readers are done
        load_w_async(W, wft, TW, tid);
        for (int idx = tid; idx < TM * (3 * TW / 4); idx += 256) {
            const int node = idx / (3 * TW / 4), q4 = idx % (3 * TW / 4);
            cp_async16(Z + node * ZNP + 4 * q4, z + (min(n0 + node, n - 1)) * 3 * TW + 4 * q4);   // rows past n: g = 0
        }
        for (int idx = tid; idx < TM * (TW / 4); idx += 256) {
            const int node = idx & (TM - 1), c4 = idx >> 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + node < n) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(g_out + (n0 + node) * TW) + c4);
                const float4 o = __ldg(reinterpret_cast<const float4*>(out + (n0 + node) * TW) + c4);
                v.x = o.x > 0.f ? g.x : 0.f;
                v.y = o.y > 0.f ? g.y : 0.f;
                v.z = o.z > 0.f ? g.z : 0.f;
                v.w = o.w > 0.f ? g.w : 0.f;
                *reinterpret_cast<float4*>(g_pre + (n0 + node) * TW + 4 * c4) = v;
            }
            GP[(4 * c4 + 0) * TP + node] = v.x;
            GP[(4 * c4 + 1) * TP + node] = v.y;
            GP[(4 * c4 + 2) * TP + node] = v.z;
            GP[(4 * c4 + 3) * TP + node] = v.w;
        }
        cp_async_wait_all();
        __syncthreads();
        if (tid < TW) {
            float s = 0.f;
            for (int node = 0; node < TM; ++node) s += GP[tid * TP + node];
            bf_acc += s;
        }
        float gz[3][8][4];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) gz[r][i][j] = 0.f;
        for (int c = 0; c < NC; ++c) {
            if (c + 1 < NC) load_w_async(W + ((c + 1) & 1) * TW * TW, wft + (size_t)(c + 1) * TW * TW, TW, tid);
            float acc[8][4];                        // G[node 8 ty + i][feature 4 tx + j] of this chunk
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            tile_gemm(GP, W + (c & 1) * TW * TW, TW, ty, tx, acc);
            if (c < C) {
                const float w[3] = {WB[3 * c], WB[3 * c + 1], WB[3 * c + 2]};
                float t[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float zz[3][4];
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float4 q = *reinterpret_cast<const float4*>(Z + (8 * ty + i) * ZNP + r * TW + 4 * tx);
                        zz[r][0] = q.x; zz[r][1] = q.y; zz[r][2] = q.z; zz[r][3] = q.w;
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float pre = bit_pre(w[0], w[1], w[2], zz[0][j], zz[1][j], zz[2][j]);
                        const float g = pre > 0.f ? acc[i][j] : 0.f;
#pragma unroll
                        for (int r = 0; r < 3; ++r) {
                            gz[r][i][j] = fmaf(w[r], g, gz[r][i][j]);
                            t[r] = fmaf(g, zz[r][j], t[r]);
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    float v = t[r];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) atomicAdd(&GWB[3 * c + r], v);
                }
            } else if (4 * tx < E) {                // the vector-level chunk: its gradient goes back to the caller
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t node = n0 + 8 * ty + i;
                    if (node < n)
                        *reinterpret_cast<float4*>(g_xf + node * E + 4 * tx) =
                            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                }
            }
            cp_async_wait_all();
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t node = n0 + 8 * ty + i;
            if (node >= n) continue;
#pragma unroll
            for (int r = 0; r < 3; ++r)
                *reinterpret_cast<float4*>(g_z + (node * 3 + r) * TW + 4 * tx) =
                    make_float4(gz[r][i][0], gz[r][i][1], gz[r][i][2], gz[r][i][3]);
        }
    }
    __syncthreads();
    for (int i = tid; i < 3 * C; i += 256)
        if (GWB[i] != 0.f) atomicAdd(g_wb + i, GWB[i]);
    if (tid < TW && bf_acc != 0.f) atomicAdd(g_bf + tid, bf_acc);
}

// ------------------------------------------------------------------------------------------------ T3 backward to Wf
// grid (chunk groups, node splits).  acc[cc][i][j] = g_Wf[(c0 + cc)*64 + 4 ty + i][4 tx + j].
// Half tiles of 64 nodes, double buffered: the global loads of half h+1 (z, g_pre, xf) are issued before the FMA loop
// of half h and turned into features after it, so their latency hides behind the arithmetic.
constexpr int T3H = 64;           // nodes per half tile

struct T3Pre {                    // one thread's share of a half tile: 4 (node, 4-dim group) items
    float4 g[4], z0[4], z1[4], z2[4], e[4];
};

__device__ __forceinline__ void t3_fetch(T3Pre& p, const float* __restrict__ z, const float* __restrict__ xf,
                                         const float* __restrict__ g_pre, int64_t n0, int64_t n, int E, bool extra,
                                         int tid) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = tid + 256 * it, node = idx >> 4, d4 = idx & 15;
        const bool ok = n0 + node < n;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        p.g[it] = p.z0[it] = p.z1[it] = p.z2[it] = p.e[it] = zero;
        if (ok) {
            p.g[it] = __ldg(reinterpret_cast<const float4*>(g_pre + (n0 + node) * TW) + d4);
            const float4* zr = reinterpret_cast<const float4*>(z + (n0 + node) * 3 * TW);
            p.z0[it] = __ldg(zr + d4);
            p.z1[it] = __ldg(zr + 16 + d4);
            p.z2[it] = __ldg(zr + 32 + d4);
            if (extra && 4 * d4 < E) p.e[it] = __ldg(reinterpret_cast<const float4*>(xf + (n0 + node) * E) + d4);
        }
    }
}

__device__ __forceinline__ void t3_store(const T3Pre& p, float* __restrict__ An, float* __restrict__ Gn,
                                         const float (&w)[T3C][3], int c0, int C, int tid) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = tid + 256 * it, node = idx >> 4, d4 = idx & 15;
        *reinterpret_cast<float4*>(Gn + node * TW + 4 * d4) = p.g[it];
        const float4 z0 = p.z0[it], z1 = p.z1[it], z2 = p.z2[it];
#pragma unroll
        for (int cc = 0; cc < T3C; ++cc) {
            const int c = c0 + cc;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < C) {
                a.x = fmaxf(bit_pre(w[cc][0], w[cc][1], w[cc][2], z0.x, z1.x, z2.x), 0.f);
                a.y = fmaxf(bit_pre(w[cc][0], w[cc][1], w[cc][2], z0.y, z1.y, z2.y), 0.f);
                a.z = fmaxf(bit_pre(w[cc][0], w[cc][1], w[cc][2], z0.z, z1.z, z2.z), 0.f);
                a.w = fmaxf(bit_pre(w[cc][0], w[cc][1], w[cc][2], z0.w, z1.w, z2.w), 0.f);
            } else if (c == C) {
                a = p.e[it];
            }
            *reinterpret_cast<float4*>(An + ((size_t)cc * T3H + node) * TW + 4 * d4) = a;
        }
    }
}

__global__ void __launch_bounds__(256, 1)
tgcn_tail_bwd_w_kernel(const float* __restrict__ z, const float* __restrict__ wb, const float* __restrict__ xf,
                       const float* __restrict__ g_pre, int64_t n, int C, int E, int64_t halves_per_split,
                       float* __restrict__ g_wf) {
    extern __shared__ __align__(16) float sm[];
    float* An = sm;                                 // [2][T3C][T3H][64]   features, node-major
    float* Gn = An + 2 * T3C * T3H * TW;            // [2][T3H][64]        g_pre, node-major
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int c0 = blockIdx.x * T3C;
    const bool extra = E > 0 && c0 <= C && C < c0 + T3C;
    const int64_t n_halves = (n + T3H - 1) / T3H;
    const int64_t h_begin = (int64_t)blockIdx.y * halves_per_split;
    const int64_t h_end = min(n_halves, h_begin + halves_per_split);
    float w[T3C][3];
#pragma unroll
    for (int cc = 0; cc < T3C; ++cc)
#pragma unroll
        for (int r = 0; r < 3; ++r) w[cc][r] = (c0 + cc < C) ? __ldg(wb + 3 * (c0 + cc) + r) : 0.f;
    unsigned long long acc2[T3C][2][4];            // acc2[cc][ip][j] = rows (2 ip, 2 ip + 1) of the 4 x 4 block, packed
#pragma unroll
    for (int cc = 0; cc < T3C; ++cc)
#pragma unroll
        for (int ip = 0; ip < 2; ++ip)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[cc][ip][j] = 0ull;
    T3Pre pre;
    if (h_begin < h_end) {
        t3_fetch(pre, z, xf, g_pre, h_begin * T3H, n, E, extra, tid);
        t3_store(pre, An, Gn, w, c0, C, tid);
    }
    __syncthreads();
    for (int64_t h = h_begin; h < h_end; ++h) {
        const int buf = (int)((h - h_begin) & 1);
        const bool more = h + 1 < h_end;
        if (more) t3_fetch(pre, z, xf, g_pre, (h + 1) * T3H, n, E, extra, tid);
        const float* Ab = An + (size_t)buf * T3C * T3H * TW;
        const float* Gb = Gn + (size_t)buf * T3H * TW;
#pragma unroll 4
        for (int node = 0; node < T3H; ++node) {
            const float4 g = *reinterpret_cast<const float4*>(Gb + node * TW + 4 * tx);
            const unsigned long long gd[4] = {pack2(g.x, g.x), pack2(g.y, g.y), pack2(g.z, g.z), pack2(g.w, g.w)};
#pragma unroll
            for (int cc = 0; cc < T3C; ++cc) {
                const float4 a = *reinterpret_cast<const float4*>(Ab + ((size_t)cc * T3H + node) * TW + 4 * ty);
                const unsigned long long a01 = pack2(a.x, a.y), a23 = pack2(a.z, a.w);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc2[cc][0][j] = ffma2(a01, gd[j], acc2[cc][0][j]);
                    acc2[cc][1][j] = ffma2(a23, gd[j], acc2[cc][1][j]);
                }
            }
        }
        if (more) t3_store(pre, An + (size_t)(buf ^ 1) * T3C * T3H * TW, Gn + (size_t)(buf ^ 1) * T3H * TW, w, c0, C, tid);
        __syncthreads();
    }
    const int F = C * TW + E;
    float acc[T3C][4][4];
#pragma unroll
    for (int cc = 0; cc < T3C; ++cc)
#pragma unroll
        for (int ip = 0; ip < 2; ++ip)
#pragma unroll
            for (int j = 0; j < 4; ++j) unpack2(acc2[cc][ip][j], acc[cc][2 * ip][j], acc[cc][2 * ip + 1][j]);
#pragma unroll
    for (int cc = 0; cc < T3C; ++cc)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = (c0 + cc) * TW + 4 * ty + i;
            if (row < F)
                red_add4(reinterpret_cast<float4*>(g_wf + (size_t)row * TW + 4 * tx),
                         make_float4(acc[cc][i][0], acc[cc][i][1], acc[cc][i][2], acc[cc][i][3]));
        }
}

static size_t tail_smem_fwd(int C) { return ((size_t)3 * TW * TP + 2 * TW * TP + 2 * TW * TW + 3 * C) * 4; }
static size_t tail_smem_bwd_z(int C) { return ((size_t)TM * ZNP + TW * TP + 2 * TW * TW + 6 * C) * 4; }
static size_t tail_smem_bwd_w() { return ((size_t)2 * T3C * T3H * TW + 2 * T3H * TW) * 4; }

}  // namespace tagrec

using namespace tagrec;

namespace tagrec {
size_t tail_tc_bwd_workspace_bytes(int C);
size_t tail_tc_bwd_w_workspace_bytes(int64_t n);
}
extern "C" size_t tagrec_tgcn_tail_workspace_bytes(int64_t n, int n_bit_conv) {
    return (size_t)(n_bit_conv + 1) * TW * TW * 4 + (size_t)n * TW * 4 + 512 + tagrec::tail_tc_bwd_workspace_bytes(n_bit_conv) +
           tagrec::tail_tc_bwd_w_workspace_bytes(n);
}

static int tail_check(int64_t n, int dim, int C, int E) {
    TAGREC_REQUIRE(dim == TW, "the TGCN tail is built for 64-d layers");
    TAGREC_REQUIRE(C >= 0 && C <= 256 && E >= 0 && E <= TW && E % 4 == 0 && C + E > 0,
                   "need 0 <= num_bit_conv <= 256 and a multiple of 4, at most 64, extra features");
    TAGREC_REQUIRE(n >= 0, "negative row count");
    return TAGREC_OK;
}

namespace tagrec {      // tgcn_tail_tc.cu
size_t tail_tc_workspace_bytes(int C);
bool tail_tc_available();
int tail_fwd_tc(const float* z, const float* wb, const float* xf, const float* wf, const float* bf, int64_t n, int C,
                int E, float* out, void* workspace, void* stream);
size_t tail_tc_bwd_workspace_bytes(int C);
size_t tail_tc_bwd_w_workspace_bytes(int64_t n);
int tail_bwd_w_tc(const float* z, const float* wb, const float* xf, const float* g_pre, int64_t n, int C, int E,
                  void* workspace, float* g_wf, void* stream);
int tail_bwd_z_tc(const float* g_out, const float* out, const float* z, const float* wb, const float* wf, int64_t n,
                  int C, int E, void* workspace, float* g_pre, float* g_z, float* g_wb, float* g_xf, float* g_bf,
                  void* stream);
}  // namespace tagrec

extern "C" size_t tagrec_tgcn_tail_fwd_workspace_bytes(int n_bit_conv) { return tail_tc_workspace_bytes(n_bit_conv); }

extern "C" int tagrec_tgcn_tail_fwd_ex(const float* z, const float* wb, const float* xf, const float* wf,
                                       const float* bf, int64_t n, int dim, int n_bit_conv, int n_extra, float* out,
                                       void* workspace, size_t workspace_bytes, int path, void* stream) {
    TAGREC_REQUIRE(z && wf && bf && out && (wb || n_bit_conv == 0) && (xf || n_extra == 0), "null pointer");
    TAGREC_REQUIRE(path == TAGREC_EVAL_AUTO || path == TAGREC_EVAL_FP32 || path == TAGREC_EVAL_TF32, "bad path");
    if (int rc = tail_check(n, dim, n_bit_conv, n_extra)) return rc;
    const bool tc = path != TAGREC_EVAL_FP32 && workspace && workspace_bytes >= tail_tc_workspace_bytes(n_bit_conv) &&
                    tail_tc_available();
    TAGREC_REQUIRE(tc || path != TAGREC_EVAL_TF32, "tensor-core path needs a workspace of tagrec_tgcn_tail_fwd_workspace_bytes");
    if (!tc) return tagrec_tgcn_tail_fwd(z, wb, xf, wf, bf, n, dim, n_bit_conv, n_extra, out, stream);
    if (n == 0) return TAGREC_OK;
    return tail_fwd_tc(z, wb, xf, wf, bf, n, n_bit_conv, n_extra, out, workspace, stream);
}

extern "C" int tagrec_tgcn_tail_fwd(const float* z, const float* wb, const float* xf, const float* wf, const float* bf,
                                    int64_t n, int dim, int n_bit_conv, int n_extra, float* out, void* stream) {
    TAGREC_REQUIRE(z && wf && bf && out && (wb || n_bit_conv == 0) && (xf || n_extra == 0), "null pointer");
    if (int rc = tail_check(n, dim, n_bit_conv, n_extra)) return rc;
    if (n == 0) return TAGREC_OK;
    const size_t smem = tail_smem_fwd(n_bit_conv);
    TAGREC_CUDA(cudaFuncSetAttribute(tgcn_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TAGREC_LAUNCH(tgcn_tail_fwd_kernel, (unsigned)((n + TM - 1) / TM), 256, smem, stream, z, wb, xf, wf, bf, n, n_bit_conv,
                  n_extra, out);
    return TAGREC_OK;
}

#ifndef TAGREC_TAIL_T3_FMA
#define TAGREC_TAIL_T3_FMA 0        // 1: keep the weight-gradient pass (T3) on the FFMA2 kernel also on the tensor-core path
#endif
static constexpr bool kTailT3Fma = TAGREC_TAIL_T3_FMA != 0;

extern "C" int tagrec_tgcn_tail_bwd(const float* g_out, const float* out, const float* z, const float* wb,
                                    const float* xf, const float* wf, int64_t n, int dim, int n_bit_conv, int n_extra,
                                    void* workspace, size_t workspace_bytes, float* g_z, float* g_wb, float* g_xf,
                                    float* g_wf, float* g_bf, void* stream) {
    return tagrec_tgcn_tail_bwd_ex(g_out, out, z, wb, xf, wf, n, dim, n_bit_conv, n_extra, workspace, workspace_bytes, g_z,
                                   g_wb, g_xf, g_wf, g_bf, TAGREC_EVAL_AUTO, stream);
}

extern "C" int tagrec_tgcn_tail_bwd_ex(const float* g_out, const float* out, const float* z, const float* wb,
                                       const float* xf, const float* wf, int64_t n, int dim, int n_bit_conv,
                                       int n_extra, void* workspace, size_t workspace_bytes, float* g_z, float* g_wb,
                                       float* g_xf, float* g_wf, float* g_bf, int path, void* stream) {
    const int C = n_bit_conv, E = n_extra;
    TAGREC_REQUIRE(g_out && out && z && wf && g_z && g_wf && g_bf && ((wb && g_wb) || C == 0) && ((xf && g_xf) || E == 0),
                   "null pointer");
    TAGREC_REQUIRE(path == TAGREC_EVAL_AUTO || path == TAGREC_EVAL_FP32 || path == TAGREC_EVAL_TF32, "bad path");
    if (int rc = tail_check(n, dim, C, E)) return rc;
    const bool tc = path != TAGREC_EVAL_FP32 && tail_tc_available();
    TAGREC_REQUIRE(tc || path != TAGREC_EVAL_TF32, "tensor-core path not available");
    cudaStream_t st = (cudaStream_t)stream;
    const int F = C * TW + E;
    TAGREC_CUDA(cudaMemsetAsync(g_wf, 0, (size_t)F * TW * 4, st));
    TAGREC_CUDA(cudaMemsetAsync(g_bf, 0, TW * 4, st));
    if (C) TAGREC_CUDA(cudaMemsetAsync(g_wb, 0, (size_t)C * 3 * 4, st));
    if (n == 0) return TAGREC_OK;
    if (!workspace || workspace_bytes < tagrec_tgcn_tail_workspace_bytes(n, C))
        return fail(TAGREC_ENOMEM, "tgcn tail workspace too small", __FILE__, __LINE__);
    const int NC = C + (E > 0 ? 1 : 0);
    float* wft = reinterpret_cast<float*>(workspace);
    float* g_pre = wft + (((size_t)NC * TW * TW + 63) / 64) * 64;
    void* tc_ws = g_pre + (size_t)n * TW;
    const int64_t n_tiles = (n + TM - 1) / TM;
    if (tc) {
        if (int rc = tail_bwd_z_tc(g_out, out, z, wb, wf, n, C, E, tc_ws, g_pre, g_z, g_wb, g_xf, g_bf, stream)) return rc;
    } else {
        TAGREC_LAUNCH(tgcn_tail_transpose_kernel, (unsigned)((NC * TW * TW + 255) / 256), 256, 0, stream, wf, C, E, wft);
        const size_t smem_z = tail_smem_bwd_z(C);
        TAGREC_CUDA(cudaFuncSetAttribute(tgcn_tail_bwd_z_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_z));
        TAGREC_LAUNCH(tgcn_tail_bwd_z_kernel, (unsigned)std::min<int64_t>(n_tiles, kSMs), 256, smem_z, stream, g_out, out, z,
                      wb, wft, n, C, E, g_pre, g_z, g_wb, g_xf, g_bf);
    }
    if (tc && !kTailT3Fma) {
        void* w_ws = reinterpret_cast<unsigned char*>(tc_ws) + tail_tc_bwd_workspace_bytes(C);
        return tail_bwd_w_tc(z, wb, xf, g_pre, n, C, E, w_ws, g_wf, stream);
    }
    const int groups = (NC + T3C - 1) / T3C;
    const int64_t n_halves = (n + T3H - 1) / T3H;
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>(n_halves, kSMs / groups));
    const int64_t tps = (n_halves + splits - 1) / splits;
    splits = (n_halves + tps - 1) / tps;
    const size_t smem_w = tail_smem_bwd_w();
    TAGREC_CUDA(cudaFuncSetAttribute(tgcn_tail_bwd_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
    TAGREC_LAUNCH(tgcn_tail_bwd_w_kernel, dim3((unsigned)groups, (unsigned)splits), 256, smem_w, stream, z, wb, xf, g_pre, n,
                  C, E, tps, g_wf);
    return TAGREC_OK;
}
