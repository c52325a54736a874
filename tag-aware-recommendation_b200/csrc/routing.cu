// K5 — disentangled (multi-intent) routing propagation for DGCF and DisenGCN.  sm_100a.
//
// Replaces model/dgcf.py:68-110 (iterate_update / factor_update) and model/disengcn.py:23-46 (Layer.forward): per
// (layer, routing iteration, factor) the reference builds a fresh torch sparse tensor from edge weights it first
// copies to the HOST (`.detach().cpu()`, dgcf.py:92, disengcn.py:36), runs sparse.sum + three sparse.mm (DGCF) or
// one (DisenGCN) on a 16-d chunk, then gathers two N x 16 tables per edge for the new edge score — >= 24 sparse
// constructions and 4*K device->host copies per forward.
//
// Here the structure (CSR of the 'plain' adjacency) never moves; the K = 4 per-edge routing weights live in ONE
// [nnz, 4] float4 array and every kernel handles the four 16-d factor chunks of a 64-d row at once:
// 16 lanes own a row (one float4 each => lane sl works for factor sl / 4), so a gathered neighbour row is one
// coalesced 256 B request shared by the four factors.
//
//   R1 edge_softmax_rowsum   w[e,:] = softmax_k(logit[e,:]);  dinv[h,k] = 1/sqrt(sum_{e in row h} w[e,k])   dgcf.py:74,95-97
//   R2 edge_scale            val[e,k] = dinv[h,k] * w[e,k] * dinv[t,k]                                    dgcf.py:98-101
//   R3 spmm4                 y[h, chunk k] = res + sum_e val[p(e),k] x[t(e), chunk k]  (+ chunk-normalise,   dgcf.py:99-101,79-80
//                            + running layer mean); p = identity or the reverse-edge permutation (A^T)       disengcn.py:39-41
//   R4 edge_dot4             d[e,k] = <a[h, chunk k], b[t, chunk k]>;  logit += d  |  w = softmax_k(d)       dgcf.py:103-109, disengcn.py:31-34
//   R5 chunk_normalize       y = x / max(||x||_chunk, 1e-12)  (optionally tanh of it)                      dgcf.py:106-108
//   R6 chunk_normalize_bwd   Jacobian-transpose of R5 applied to a gradient
//   R7 csr_reverse_perm      rev[e(h,t)] = e(t,h) over the structurally symmetric CSR (backward needs the values of
//                            A^T; the routing weights are NOT symmetric — SURVEY §8 a-10)
// All are HBM/L2-bandwidth bound gathers; algorithmic bytes per edge: R3/R4 256 (row) + 16 (weights) + 4 (col).
#include "common.cuh"

namespace tagrec {

constexpr int RL = 16;     // lanes per 64-d row

struct RowCtx {
    int lane, sub, sl;
    unsigned mask;
    int64_t row;
    bool valid;
    int64_t s, e;
};

__device__ __forceinline__ RowCtx row_ctx(const int64_t* __restrict__ rowptr, int64_t n_rows) {
    RowCtx c;
    c.lane = threadIdx.x & 31;
    c.sub = c.lane >> 4;
    c.sl = c.lane & 15;
    c.mask = 0xffffu << (16 * c.sub);
    c.row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + c.sub;
    c.valid = c.row < n_rows;
    c.s = c.valid ? __ldg(rowptr + c.row) : 0;
    c.e = c.valid ? __ldg(rowptr + c.row + 1) : 0;
    return c;
}

// sum over the 4 lanes that share a factor chunk
__device__ __forceinline__ float sum4(float v, unsigned mask) {
    v += __shfl_xor_sync(mask, v, 1, 16);
    v += __shfl_xor_sync(mask, v, 2, 16);
    return v;
}

__device__ __forceinline__ float4 softmax4(float4 x) {
    const float m = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
    float4 e = make_float4(expf(x.x - m), expf(x.y - m), expf(x.z - m), expf(x.w - m));
    const float s = (e.x + e.y) + (e.z + e.w);
    return make_float4(e.x / s, e.y / s, e.z / s, e.w / s);
}

// ---------------------------------------------------------------------------------------------- R1
// Edge-parallel (hub rows of a power-law graph would serialise a row-per-sub-warp walk): one edge per thread,
// row sums through red.global.add.v4.f32 after a warp-level merge of the runs of equal rows (edges are row-sorted,
// so a warp usually covers one or two rows).  edge_row = row id of every edge (int32 [nnz], built once per graph).
__global__ void __launch_bounds__(256)
edge_softmax_rowsum_kernel(const int32_t* __restrict__ edge_row, int64_t nnz, const float4* __restrict__ logit,
                           float4* __restrict__ w, float4* __restrict__ rowsum) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int row = -1;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < nnz) {
        row = __ldg(edge_row + e);
        p = softmax4(__ldg(logit + e));
        w[e] = p;
    }
    // segmented inclusive scan over the lanes of a run of equal rows; the last lane of a run issues the reduction
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int r2 = __shfl_up_sync(0xffffffffu, row, o);
        const float x = __shfl_up_sync(0xffffffffu, p.x, o), y = __shfl_up_sync(0xffffffffu, p.y, o);
        const float z = __shfl_up_sync(0xffffffffu, p.z, o), q = __shfl_up_sync(0xffffffffu, p.w, o);
        if (lane >= o && r2 == row) { p.x += x; p.y += y; p.z += z; p.w += q; }
    }
    const int next = __shfl_down_sync(0xffffffffu, row, 1);
    if (row >= 0 && (lane == 31 || next != row)) red_add4(rowsum + row, p);
}

__global__ void __launch_bounds__(256) rowsum_to_dinv_kernel(float4* __restrict__ d, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = d[i];      // dgcf.py:96-97: 1/sqrt(rowsum), inf -> 0; rows without edges are absent => 0
    d[i] = make_float4(s.x > 0.f ? 1.f / sqrtf(s.x) : 0.f, s.y > 0.f ? 1.f / sqrtf(s.y) : 0.f,
                       s.z > 0.f ? 1.f / sqrtf(s.z) : 0.f, s.w > 0.f ? 1.f / sqrtf(s.w) : 0.f);
}

// ---------------------------------------------------------------------------------------------- R2
__global__ void __launch_bounds__(256)
edge_scale_kernel(const int32_t* __restrict__ edge_row, const int32_t* __restrict__ col, int64_t nnz,
                  const float4* __restrict__ w, const float4* __restrict__ dinv, float4* __restrict__ val) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const float4 dh = __ldg(dinv + __ldg(edge_row + e));
    const float4 p = __ldg(w + e);
    const float4 dt = __ldg(dinv + __ldg(col + e));
    val[e] = make_float4(dh.x * p.x * dt.x, dh.y * p.y * dt.y, dh.z * p.z * dt.z, dh.w * p.w * dt.w);
}

// ---------------------------------------------------------------------------------------------- R3
struct Spmm4Epi {
    const float* res;     // optional residual row table added to the sum
    float* y_raw;         // optional: un-normalised result
    float* y_norm;        // optional: per-chunk L2-normalised result
    float* mean_acc;      // optional running mean of the normalised layers (dgcf.py:59-61)
    const float* mean_x0; // first layer: mean_acc starts from this table (ego)
    int mean_first, mean_last;
    float mean_scale;
};

constexpr int LONG4 = 256;      // rows with more entries are produced piecewise (spmm4_piece_kernel)

__device__ __forceinline__ float4 spmm4_gather(const int32_t* __restrict__ col, const float* __restrict__ val,
                                               const int32_t* __restrict__ perm, const float4* __restrict__ x4,
                                               int64_t begin, int64_t end, int64_t stride, int sl, unsigned mask) {
    const int k = sl >> 2;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t base = begin; base < end; base += stride) {
        const int cnt = (int)min((int64_t)RL, end - base);
        int cj = 0;
        int64_t pj = 0;
        if (sl < cnt) {
            cj = __ldg(col + base + sl);
            pj = perm ? (int64_t)__ldg(perm + base + sl) : base + sl;
        }
        for (int j = 0; j < cnt; ++j) {
            const int cc = __shfl_sync(mask, cj, j, 16);
            const int64_t pp = __shfl_sync(mask, pj, j, 16);
            const float wv = __ldg(val + pp * 4 + k);
            fma4(acc, wv, ldg4(x4 + (int64_t)cc * RL + sl));
        }
    }
    return acc;
}

__device__ __forceinline__ void spmm4_epilogue(const Spmm4Epi& ep, int64_t row, int sl, unsigned mask, float4 acc) {
    const int64_t o = row * RL + sl;
    if (ep.res) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(ep.res) + o);
        acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
    }
    if (ep.y_raw) reinterpret_cast<float4*>(ep.y_raw)[o] = acc;
    if (ep.y_norm || ep.mean_acc) {
        const float nn = fmaxf(sqrtf(sum4(dot4(acc, acc), mask)), 1e-12f);
        const float4 yn = make_float4(acc.x / nn, acc.y / nn, acc.z / nn, acc.w / nn);
        if (ep.y_norm) reinterpret_cast<float4*>(ep.y_norm)[o] = yn;
        if (ep.mean_acc) {
            float4* m4 = reinterpret_cast<float4*>(ep.mean_acc);
            float4 a = ep.mean_first ? __ldg(reinterpret_cast<const float4*>(ep.mean_x0) + o) : m4[o];
            a.x += yn.x; a.y += yn.y; a.z += yn.z; a.w += yn.w;
            if (ep.mean_last) { a.x *= ep.mean_scale; a.y *= ep.mean_scale; a.z *= ep.mean_scale; a.w *= ep.mean_scale; }
            m4[o] = a;
        }
    }
}

__global__ void __launch_bounds__(256)
spmm4_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
             const float* __restrict__ val, const int32_t* __restrict__ perm, const float4* __restrict__ x4,
             Spmm4Epi ep, int skip_long) {
    const RowCtx c = row_ctx(rowptr, n_rows);
    if (!c.valid || (skip_long && c.e - c.s > LONG4)) return;      // masks below are per sub-warp
    const float4 acc = spmm4_gather(col, val, perm, x4, c.s, c.e, RL, c.sl, c.mask);
    spmm4_epilogue(ep, c.row, c.sl, c.mask, acc);
}

// Long rows are cut into pieces of at most LONG4_PIECE entries; one block per piece: its 16 sub-warps take alternating
// 16-edge chunks, partial rows meet in shared memory and leave through one red.global.add.v4.f32 per lane into the
// row's scratch slot; a second tiny launch runs the fused epilogue on the complete rows and re-zeroes the scratch.
__global__ void __launch_bounds__(256)
spmm4_piece_kernel(const int32_t* __restrict__ col, const int32_t* __restrict__ piece_slot,
                   const int64_t* __restrict__ piece_begin, const int64_t* __restrict__ piece_end,
                   const float* __restrict__ val, const int32_t* __restrict__ perm, const float4* __restrict__ x4,
                   float4* __restrict__ scratch) {
    __shared__ float4 part[16][RL];
    const int lane = threadIdx.x & 31, sub16 = threadIdx.x >> 4, sl = lane & 15;
    const unsigned mask = 0xffffu << (16 * ((lane >> 4) & 1));
    const int64_t s = __ldg(piece_begin + blockIdx.x), e = __ldg(piece_end + blockIdx.x);
    part[sub16][sl] = spmm4_gather(col, val, perm, x4, s + (int64_t)sub16 * RL, e, 16 * RL, sl, mask);
    __syncthreads();
    if (sub16 == 0) {
        float4 acc = part[0][sl];
#pragma unroll
        for (int i = 1; i < 16; ++i) {
            const float4 p = part[i][sl];
            acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
        }
        red_add4(scratch + (int64_t)__ldg(piece_slot + blockIdx.x) * RL + sl, acc);
    }
}

__global__ void __launch_bounds__(256)
spmm4_long_epilogue_kernel(const int32_t* __restrict__ long_rows, int64_t n_long, float4* __restrict__ scratch,
                           Spmm4Epi ep) {
    const int lane = threadIdx.x & 31, sl = lane & 15;
    const unsigned mask = 0xffffu << (16 * (lane >> 4));
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (i >= n_long) return;
    const float4 acc = __ldcg(scratch + i * RL + sl);
    __stcg(scratch + i * RL + sl, make_float4(0.f, 0.f, 0.f, 0.f));
    spmm4_epilogue(ep, __ldg(long_rows + i), sl, mask, acc);
}

// ---------------------------------------------------------------------------------------------- R4
// Edge-parallel: 16 lanes per edge gather a[head] and b[tail] (consecutive edges share the head row: L1 hits).
template <int MODE>   // 0: logit[e] += d (dgcf.py:109)   1: w[e] = softmax_k(d) (disengcn.py:33-34)
__global__ void __launch_bounds__(256)
edge_dot4_kernel(const int32_t* __restrict__ edge_row, const int32_t* __restrict__ col, int64_t nnz,
                 const float4* __restrict__ a4, const float4* __restrict__ b4, float4* __restrict__ out) {
    const int lane = threadIdx.x & 31, sl = lane & 15;
    const unsigned mask = 0xffffu << (16 * (lane >> 4));
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (e >= nnz) return;                                          // whole sub-warp leaves together
    const int64_t h = __ldg(edge_row + e), t = __ldg(col + e);
    const float d = sum4(dot4(ldg4(a4 + h * RL + sl), ldg4(b4 + t * RL + sl)), mask);
    const float d1 = __shfl_sync(mask, d, 4, 16), d2 = __shfl_sync(mask, d, 8, 16), d3 = __shfl_sync(mask, d, 12, 16);
    if (sl == 0) {
        if (MODE == 0) {
            float4 l = out[e];
            l.x += d; l.y += d1; l.z += d2; l.w += d3;
            out[e] = l;
        } else {
            out[e] = softmax4(make_float4(d, d1, d2, d3));
        }
    }
}

// ---------------------------------------------------------------------------------------------- R5 / R6
__global__ void __launch_bounds__(256)
chunk_normalize_kernel(const float4* __restrict__ x4, int64_t n_rows, int apply_tanh, float4* __restrict__ y4) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // one float4 per thread
    const bool valid = idx < n_rows * RL;
    float4 v = valid ? __ldg(x4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    float ss = dot4(v, v);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    const float nn = fmaxf(sqrtf(ss), 1e-12f);
    v = make_float4(v.x / nn, v.y / nn, v.z / nn, v.w / nn);
    if (apply_tanh) v = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
    if (valid) y4[idx] = v;
}

__global__ void __launch_bounds__(256)
chunk_normalize_bwd_kernel(const float4* __restrict__ g4, const float4* __restrict__ x4, int64_t n_rows,
                           float4* __restrict__ out4) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < n_rows * RL;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 x = valid ? __ldg(x4 + idx) : z;
    const float4 g = valid ? __ldg(g4 + idx) : z;
    float ss = dot4(x, x), dt = dot4(x, g);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    dt += __shfl_xor_sync(0xffffffffu, dt, 1);
    dt += __shfl_xor_sync(0xffffffffu, dt, 2);
    const float nn = sqrtf(ss);
    float4 o;
    if (nn >= 1e-12f) {
        const float proj = dt / nn;
        o = make_float4((g.x - (x.x / nn) * proj) / nn, (g.y - (x.y / nn) * proj) / nn, (g.z - (x.z / nn) * proj) / nn,
                        (g.w - (x.w / nn) * proj) / nn);
    } else {
        o = make_float4(g.x / 1e-12f, g.y / 1e-12f, g.z / 1e-12f, g.w / 1e-12f);
    }
    if (valid) out4[idx] = o;
}

// ---------------------------------------------------------------------------------------------- R7
__global__ void __launch_bounds__(256)
csr_reverse_perm_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                        int32_t* __restrict__ rev, int* __restrict__ missing) {
    const RowCtx c = row_ctx(rowptr, n_rows);
    if (!c.valid) return;
    for (int64_t j = c.s + c.sl; j < c.e; j += RL) {
        const int64_t t = __ldg(col + j);
        int64_t lo = __ldg(rowptr + t), hi = __ldg(rowptr + t + 1);
        const int64_t end = hi;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)__ldg(col + mid) < c.row) lo = mid + 1; else hi = mid;
        }
        if (lo < end && (int64_t)__ldg(col + lo) == c.row) rev[j] = (int32_t)lo;
        else { rev[j] = (int32_t)j; atomicAdd(missing, 1); }
    }
}

static unsigned row_grid(int64_t n_rows) { return (unsigned)((n_rows + 15) / 16); }   // 8 warps x 2 rows per block

}  // namespace tagrec

using namespace tagrec;

extern "C" int tagrec_edge_softmax_rowsum(const int32_t* edge_row, int64_t nnz, int64_t n_rows, const float* logit,
                                          float* w, float* dinv, void* stream) {
    TAGREC_REQUIRE(edge_row && logit && w && dinv, "null pointer");
    if (n_rows == 0) return TAGREC_OK;
    TAGREC_CUDA(cudaMemsetAsync(dinv, 0, (size_t)n_rows * 4 * sizeof(float), (cudaStream_t)stream));
    if (nnz > 0)
        TAGREC_LAUNCH(edge_softmax_rowsum_kernel, (unsigned)((nnz + 255) / 256), 256, 0, stream, edge_row, nnz,
                      reinterpret_cast<const float4*>(logit), reinterpret_cast<float4*>(w),
                      reinterpret_cast<float4*>(dinv));
    TAGREC_LAUNCH(rowsum_to_dinv_kernel, (unsigned)((n_rows + 255) / 256), 256, 0, stream,
                  reinterpret_cast<float4*>(dinv), n_rows);
    return TAGREC_OK;
}

extern "C" int tagrec_edge_scale(const int32_t* edge_row, const int32_t* col, int64_t nnz, const float* w,
                                 const float* dinv, float* val, void* stream) {
    TAGREC_REQUIRE(edge_row && col && w && dinv && val, "null pointer");
    if (nnz == 0) return TAGREC_OK;
    TAGREC_LAUNCH(edge_scale_kernel, (unsigned)((nnz + 255) / 256), 256, 0, stream, edge_row, col, nnz,
                  reinterpret_cast<const float4*>(w), reinterpret_cast<const float4*>(dinv), reinterpret_cast<float4*>(val));
    return TAGREC_OK;
}

extern "C" int tagrec_spmm4(const int64_t* rowptr, const int32_t* col, int64_t n_rows, const tagrec_route_plan_t* plan,
                            const float* val, const int32_t* perm, const float* x, const float* res, float* y_raw,
                            float* y_norm, float* mean_acc, const float* mean_x0, int mean_first, int mean_last,
                            float mean_scale, void* stream) {
    TAGREC_REQUIRE(rowptr && col && val && x, "null pointer");
    TAGREC_REQUIRE(y_raw || y_norm || mean_acc, "no output requested");
    TAGREC_REQUIRE(!mean_acc || !mean_first || mean_x0, "mean_first needs mean_x0");
    const int64_t n_long = plan ? plan->n_long : 0;
    if (n_long > 0)
        TAGREC_REQUIRE(plan->long_rows && plan->piece_slot && plan->piece_begin && plan->piece_end && plan->scratch &&
                           plan->n_pieces > 0, "long-row plan arrays missing");
    if (n_rows == 0) return TAGREC_OK;
    Spmm4Epi ep{res, y_raw, y_norm, mean_acc, mean_x0, mean_first, mean_last, mean_scale};
    const float4* x4 = reinterpret_cast<const float4*>(x);
    TAGREC_LAUNCH(spmm4_kernel, row_grid(n_rows), 256, 0, stream, rowptr, col, n_rows, val, perm, x4, ep,
                  (int)(n_long > 0));
    if (n_long > 0) {
        float4* scr = reinterpret_cast<float4*>(plan->scratch);
        TAGREC_LAUNCH(spmm4_piece_kernel, (unsigned)plan->n_pieces, 256, 0, stream, col, plan->piece_slot,
                      plan->piece_begin, plan->piece_end, val, perm, x4, scr);
        TAGREC_LAUNCH(spmm4_long_epilogue_kernel, (unsigned)((n_long * 16 + 255) / 256), 256, 0, stream, plan->long_rows,
                      n_long, scr, ep);
    }
    return TAGREC_OK;
}

extern "C" int tagrec_spmm4_long_threshold(void) { return LONG4; }
extern "C" int tagrec_spmm4_piece(void) { return TAGREC_ROUTE_PIECE; }

extern "C" int tagrec_edge_dot4(const int32_t* edge_row, const int32_t* col, int64_t nnz, const float* a,
                                const float* b, float* out, int mode, void* stream) {
    TAGREC_REQUIRE(edge_row && col && a && b && out, "null pointer");
    TAGREC_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (accumulate) or 1 (softmax)");
    if (nnz == 0) return TAGREC_OK;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float4* o4 = reinterpret_cast<float4*>(out);
    const unsigned grid = (unsigned)((nnz * 16 + 255) / 256);
    if (mode == 0) {
        TAGREC_LAUNCH(edge_dot4_kernel<0>, grid, 256, 0, stream, edge_row, col, nnz, a4, b4, o4);
    } else {
        TAGREC_LAUNCH(edge_dot4_kernel<1>, grid, 256, 0, stream, edge_row, col, nnz, a4, b4, o4);
    }
    return TAGREC_OK;
}

extern "C" int tagrec_chunk_normalize(const float* x, int64_t n_rows, int apply_tanh, float* y, void* stream) {
    TAGREC_REQUIRE(x && y, "null pointer");
    if (n_rows == 0) return TAGREC_OK;
    TAGREC_LAUNCH(chunk_normalize_kernel, (unsigned)((n_rows * RL + 255) / 256), 256, 0, stream,
                  reinterpret_cast<const float4*>(x), n_rows, apply_tanh, reinterpret_cast<float4*>(y));
    return TAGREC_OK;
}

extern "C" int tagrec_chunk_normalize_bwd(const float* g, const float* x, int64_t n_rows, float* out, void* stream) {
    TAGREC_REQUIRE(g && x && out, "null pointer");
    if (n_rows == 0) return TAGREC_OK;
    TAGREC_LAUNCH(chunk_normalize_bwd_kernel, (unsigned)((n_rows * RL + 255) / 256), 256, 0, stream,
                  reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(x), n_rows,
                  reinterpret_cast<float4*>(out));
    return TAGREC_OK;
}

extern "C" int tagrec_csr_reverse_perm(const int64_t* rowptr, const int32_t* col, int64_t n_rows, int32_t* rev,
                                       int32_t* missing, void* stream) {
    TAGREC_REQUIRE(rowptr && col && rev && missing, "null pointer");
    TAGREC_CUDA(cudaMemsetAsync(missing, 0, sizeof(int32_t), (cudaStream_t)stream));
    if (n_rows == 0) return TAGREC_OK;
    TAGREC_LAUNCH(csr_reverse_perm_kernel, row_grid(n_rows), 256, 0, stream, rowptr, col, n_rows, rev, missing);
    return TAGREC_OK;
}
