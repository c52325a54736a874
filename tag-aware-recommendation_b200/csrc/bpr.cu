// K2 — fused BPR step: gather (u, i+, i-) rows, dot products, softplus/logsigmoid loss, L2 regulariser and the
// scatter-add of all gradients, one kernel.  sm_100a.
//
// Replaces model/lightgcn.py:68-82 and model/ngcf.py:95-105 (three advanced-index gathers of the propagated table,
// three of the ego table, model/help/loss.py:4-12 mul_loss, loss.py:27-32 l2reg_loss = ~25 small torch kernels) and
// the index_put_(accumulate=True) kernels of their backward.
//
// Mapping: dim = 4*LPR floats; one sub-warp of LPR lanes per triple, one float4 per lane, so each gathered row is a
// single coalesced request and each scattered gradient row is ONE red.global.add.v4.f32 per lane (the sub-warp's
// whole row in one L2 reduction transaction group, no return value).  Two triples of a warp that hit the same row
// are merged before the reduction (warp-aggregated).  Loss / regulariser sums: registers -> warp shuffle -> smem ->
// one atomicAdd per block.
//
// HBM bytes per triple (roofline numerator): 24 (indices) + 6 rows gathered + 6 rows reduced = 24 + 12*4*dim
//   = 3096 B at dim 64 (SURVEY §8 d).
#include "common.cuh"

namespace tagrec {

template <int LPR>
__device__ __forceinline__ float sub_sum_b(float v, unsigned mask) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, LPR);
    return v;
}

__device__ __forceinline__ float softplus_torch(float x) {
    // torch.nn.functional.softplus(beta=1, threshold=20)
    return x > 20.f ? x : log1pf(expf(x));
}

__device__ __forceinline__ float neg_logsigmoid_neg(float x) {
    // -logsigmoid(-x) = max(x,0) + log1p(exp(-|x|))   (torch's stable form)
    return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

template <int LPR>
__global__ void __launch_bounds__(256)
bpr_kernel(const int64_t* __restrict__ triples, int64_t b, int64_t item_offset, const float4* __restrict__ f4,
           const float4* __restrict__ r4, float reg, int loss_kind, float4* __restrict__ gf4, float4* __restrict__ gr4,
           float* __restrict__ loss_out) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, sl = lane % LPR;
    const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
    const int64_t t = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
    const bool valid = t < b;
    const float inv_b = 1.f / (float)b;

    float loss_t = 0.f, reg_t = 0.f;
    int64_t u = -1, p = -1, q = -1;
    float4 gu = make_float4(0, 0, 0, 0), gp = gu, gq = gu;     // gradient rows w.r.t. the final table
    float4 ru = gu, rp = gu, rq = gu;                          // regulariser source rows
    if (valid) {
        u = __ldg(triples + 3 * t);
        p = __ldg(triples + 3 * t + 1) + item_offset;
        q = __ldg(triples + 3 * t + 2) + item_offset;
        const float4 fu = __ldg(f4 + u * LPR + sl);
        const float4 fp = __ldg(f4 + p * LPR + sl);
        const float4 fq = __ldg(f4 + q * LPR + sl);
        const float pos = sub_sum_b<LPR>(dot4(fu, fp), mask);
        const float neg = sub_sum_b<LPR>(dot4(fu, fq), mask);
        const float x = neg - pos;
        loss_t = loss_kind == 1 ? neg_logsigmoid_neg(x) : softplus_torch(x);
        const float s = inv_b / (1.f + expf(-x));              // sigmoid(x) / B
        gu = make_float4(s * (fq.x - fp.x), s * (fq.y - fp.y), s * (fq.z - fp.z), s * (fq.w - fp.w));
        gp = make_float4(-s * fu.x, -s * fu.y, -s * fu.z, -s * fu.w);
        gq = make_float4(s * fu.x, s * fu.y, s * fu.z, s * fu.w);
        if (reg != 0.f) {
            if (r4 == f4) {
                ru = fu; rp = fp; rq = fq;
            } else {
                ru = __ldg(r4 + u * LPR + sl);
                rp = __ldg(r4 + p * LPR + sl);
                rq = __ldg(r4 + q * LPR + sl);
            }
            reg_t = sub_sum_b<LPR>(dot4(ru, ru) + dot4(rp, rp) + dot4(rq, rq), mask);
        }
    }

    // ---- warp-aggregated scatter: fold rows another sub-warp of this warp also touches into the lowest one ----
    if (RPW > 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            int64_t mine = k == 0 ? u : (k == 1 ? p : q);
            float4& g = k == 0 ? gu : (k == 1 ? gp : gq);
            float4& rr = k == 0 ? ru : (k == 1 ? rp : rq);
            // users never collide with items (offset), so only same-role rows can match
#pragma unroll
            for (int o = LPR; o < 32; o <<= 1) {
                const int64_t other = __shfl_xor_sync(0xffffffffu, mine, o);
                const float4 og = make_float4(__shfl_xor_sync(0xffffffffu, g.x, o), __shfl_xor_sync(0xffffffffu, g.y, o),
                                              __shfl_xor_sync(0xffffffffu, g.z, o), __shfl_xor_sync(0xffffffffu, g.w, o));
                const float4 orr =
                    make_float4(__shfl_xor_sync(0xffffffffu, rr.x, o), __shfl_xor_sync(0xffffffffu, rr.y, o),
                                __shfl_xor_sync(0xffffffffu, rr.z, o), __shfl_xor_sync(0xffffffffu, rr.w, o));
                if (mine >= 0 && other == mine) {
                    if ((sub & (o / LPR)) == 0) {   // lower sub-warp keeps the merged row
                        g.x += og.x; g.y += og.y; g.z += og.z; g.w += og.w;
                        rr.x += orr.x; rr.y += orr.y; rr.z += orr.z; rr.w += orr.w;
                    } else {
                        if (k == 0) u = -2; else if (k == 1) p = -2; else q = -2;   // handed over
                        mine = -2;
                    }
                }
            }
        }
    }
    if (valid) {
        if (u >= 0) red_add4(gf4 + u * LPR + sl, gu);
        if (p >= 0) red_add4(gf4 + p * LPR + sl, gp);
        if (q >= 0) red_add4(gf4 + q * LPR + sl, gq);
        if (reg != 0.f && gr4) {
            const float c = reg * inv_b;
            if (u >= 0) red_add4(gr4 + u * LPR + sl, make_float4(c * ru.x, c * ru.y, c * ru.z, c * ru.w));
            if (p >= 0) red_add4(gr4 + p * LPR + sl, make_float4(c * rp.x, c * rp.y, c * rp.z, c * rp.w));
            if (q >= 0) red_add4(gr4 + q * LPR + sl, make_float4(c * rq.x, c * rq.y, c * rq.z, c * rq.w));
        }
    }

    // ---- block reduction of the two scalars (one value per sub-warp lives in lane sl == 0) ----
    __shared__ float s_loss[8], s_reg[8];
    float l = (sl == 0) ? loss_t : 0.f, r = (sl == 0) ? reg_t : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l += __shfl_xor_sync(0xffffffffu, l, o);
        r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    if (lane == 0) {
        s_loss[threadIdx.x >> 5] = l;
        s_reg[threadIdx.x >> 5] = r;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tl = 0.f, tr = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            tl += s_loss[w];
            tr += s_reg[w];
        }
        atomicAdd(loss_out, tl * inv_b);
        if (reg != 0.f) atomicAdd(loss_out + 1, 0.5f * reg * tr * inv_b);
    }
}

// Wide rows (dim > 128: NGCF's 256-d concat, ngcf.py:89; TGCN's 192-d, tgcn.py:227-229): one WARP per triple, each
// lane owns the float4 chunks sl, sl+32, ... of a row (NC = ceil(dim/128) chunks in registers).  Same arithmetic.
template <int NC>
__global__ void __launch_bounds__(256)
bpr_wide_kernel(const int64_t* __restrict__ triples, int64_t b, int64_t item_offset, const float4* __restrict__ f4,
                const float4* __restrict__ r4, int c4, float reg, int loss_kind, float4* __restrict__ gf4,
                float4* __restrict__ gr4, float* __restrict__ loss_out) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const float inv_b = 1.f / (float)b;
    float loss_t = 0.f, reg_t = 0.f;
    if (t < b) {
        const int64_t u = __ldg(triples + 3 * t);
        const int64_t p = __ldg(triples + 3 * t + 1) + item_offset;
        const int64_t q = __ldg(triples + 3 * t + 2) + item_offset;
        float4 fu[NC], fp[NC], fq[NC];
        float pos = 0.f, neg = 0.f;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int j = lane + 32 * c;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            fu[c] = j < c4 ? __ldg(f4 + u * c4 + j) : z;
            fp[c] = j < c4 ? __ldg(f4 + p * c4 + j) : z;
            fq[c] = j < c4 ? __ldg(f4 + q * c4 + j) : z;
            pos += dot4(fu[c], fp[c]);
            neg += dot4(fu[c], fq[c]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pos += __shfl_xor_sync(0xffffffffu, pos, o);
            neg += __shfl_xor_sync(0xffffffffu, neg, o);
        }
        const float x = neg - pos;
        loss_t = loss_kind == 1 ? neg_logsigmoid_neg(x) : softplus_torch(x);
        const float s = inv_b / (1.f + expf(-x));
        const bool same = r4 == f4;
        const float cr = reg * inv_b;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int j = lane + 32 * c;
            if (j >= c4) continue;
            float4 gu = make_float4(s * (fq[c].x - fp[c].x), s * (fq[c].y - fp[c].y), s * (fq[c].z - fp[c].z),
                                    s * (fq[c].w - fp[c].w));
            float4 gp = make_float4(-s * fu[c].x, -s * fu[c].y, -s * fu[c].z, -s * fu[c].w);
            float4 gq = make_float4(s * fu[c].x, s * fu[c].y, s * fu[c].z, s * fu[c].w);
            if (reg != 0.f) {
                const float4 ru = same ? fu[c] : __ldg(r4 + u * c4 + j);
                const float4 rp = same ? fp[c] : __ldg(r4 + p * c4 + j);
                const float4 rq = same ? fq[c] : __ldg(r4 + q * c4 + j);
                reg_t += dot4(ru, ru) + dot4(rp, rp) + dot4(rq, rq);
                if (gr4 == gf4) {   // the L2 term reads the table the scores read: one reduction per row
                    fma4(gu, cr, ru);
                    fma4(gp, cr, rp);
                    fma4(gq, cr, rq);
                } else if (gr4) {
                    red_add4(gr4 + u * c4 + j, make_float4(cr * ru.x, cr * ru.y, cr * ru.z, cr * ru.w));
                    red_add4(gr4 + p * c4 + j, make_float4(cr * rp.x, cr * rp.y, cr * rp.z, cr * rp.w));
                    red_add4(gr4 + q * c4 + j, make_float4(cr * rq.x, cr * rq.y, cr * rq.z, cr * rq.w));
                }
            }
            red_add4(gf4 + u * c4 + j, gu);
            red_add4(gf4 + p * c4 + j, gp);
            red_add4(gf4 + q * c4 + j, gq);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) reg_t += __shfl_xor_sync(0xffffffffu, reg_t, o);
    }
    __shared__ float s_loss[8], s_reg[8];
    if (lane == 0) {
        s_loss[threadIdx.x >> 5] = loss_t;
        s_reg[threadIdx.x >> 5] = reg_t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tl = 0.f, tr = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            tl += s_loss[w];
            tr += s_reg[w];
        }
        atomicAdd(loss_out, tl * inv_b);
        if (reg != 0.f) atomicAdd(loss_out + 1, 0.5f * reg * tr * inv_b);
    }
}

}  // namespace tagrec

using namespace tagrec;

extern "C" int tagrec_bpr_fwd_bwd(const int64_t* triples, int64_t b, int64_t item_offset, const float* final_table,
                                  const float* reg_src, int dim, float reg, int loss_kind, float* g_final,
                                  float* g_reg, float* loss_out, void* stream) {
    TAGREC_REQUIRE(triples && final_table && g_final && loss_out, "null pointer");
    TAGREC_REQUIRE(b > 0, "empty batch");
    TAGREC_REQUIRE(dim >= 4 && dim % 4 == 0 && dim <= 1024, "dim must be a multiple of 4, at most 1024");
    TAGREC_REQUIRE(reg == 0.f || reg_src, "reg != 0 needs reg_src");
    TAGREC_CUDA(cudaMemsetAsync(loss_out, 0, 2 * sizeof(float), (cudaStream_t)stream));
    if (!(dim == 32 || dim == 64 || dim == 128)) {
        const int c4 = dim / 4, nc = (c4 + 31) / 32;
        const unsigned wgrid = (unsigned)((b + 7) / 8);
        const float4* wf4 = reinterpret_cast<const float4*>(final_table);
        const float4* wr4 = reinterpret_cast<const float4*>(reg_src);
        float4* wgf4 = reinterpret_cast<float4*>(g_final);
        float4* wgr4 = reinterpret_cast<float4*>(g_reg);
#define TAGREC_BPR_WIDE(NC) TAGREC_LAUNCH((bpr_wide_kernel<NC>), wgrid, 256, 0, stream, triples, b, item_offset, wf4, \
                                          wr4, c4, reg, loss_kind, wgf4, wgr4, loss_out)
        if (nc <= 1) TAGREC_BPR_WIDE(1);
        else if (nc <= 2) TAGREC_BPR_WIDE(2);
        else if (nc <= 4) TAGREC_BPR_WIDE(4);
        else TAGREC_BPR_WIDE(8);
#undef TAGREC_BPR_WIDE
        return TAGREC_OK;
    }
    const int lpr = dim / 4, rpw = 32 / lpr;
    const int64_t warps = (b + rpw - 1) / rpw;
    const unsigned grid = (unsigned)((warps + 7) / 8);
    const float4* f4 = reinterpret_cast<const float4*>(final_table);
    const float4* r4 = reinterpret_cast<const float4*>(reg_src);
    float4* gf4 = reinterpret_cast<float4*>(g_final);
    float4* gr4 = reinterpret_cast<float4*>(g_reg);
    if (lpr == 16) {
        TAGREC_LAUNCH((bpr_kernel<16>), grid, 256, 0, stream, triples, b, item_offset, f4, r4, reg, loss_kind, gf4, gr4,
                      loss_out);
    } else if (lpr == 8) {
        TAGREC_LAUNCH((bpr_kernel<8>), grid, 256, 0, stream, triples, b, item_offset, f4, r4, reg, loss_kind, gf4, gr4,
                      loss_out);
    } else {
        TAGREC_LAUNCH((bpr_kernel<32>), grid, 256, 0, stream, triples, b, item_offset, f4, r4, reg, loss_kind, gf4, gr4,
                      loss_out);
    }
    return TAGREC_OK;
}
