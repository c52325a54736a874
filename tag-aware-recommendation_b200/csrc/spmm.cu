// K1 — CSR SpMM with fused LightGCN epilogues (forward: raw layer + row-normalised running mean; backward: the
// normalise-Jacobian term + transposed propagation in one pass).  sm_100a.
//
// Replaces model/help/adj.py:158-167 (split_mm == torch.sparse.mm, which re-coalesces the COO and calls a generic
// SpMM every call) together with model/lightgcn.py:55-60 (F.normalize, stack, mean) and their autograd.
//
// Mapping.  A table row is dim = 4*LPR floats; LPR lanes (a "sub-warp", 16 for dim 64) own one row and each lane
// keeps one float4 of it, so every gathered neighbour row is ONE coalesced 128-bit-per-lane request (256 B for
// dim 64).  A warp owns RPW = 32/LPR consecutive output rows.  Short rows run side by side (one per sub-warp);
// when that would leave lanes idle (one long + one short row) the whole warp walks the rows one after the other
// and the sub-warps take alternating index chunks.  Each sub-warp reads LPR (col,val) pairs with one coalesced
// load, prefetches the next chunk, and issues the gathers 8 at a time (8 independent LDG.128 per lane in flight).
// Rows above TAGREC_LONG_ROW nnz are cut into TAGREC_LONG_CHUNK pieces that run in the first blocks of the same
// grid; their partial sums meet in an L2-resident scratch row through red.global.add.v4.f32 and the last piece to
// arrive runs the fused epilogue (so the epilogue always sees the complete row).
//
// HBM bytes per launch (the roofline numerator): nnz*(4 col + 4 val + 4*dim gather)
//   + n_rows*(8 rowptr + 3 * 4*dim epilogue traffic)  = 264 B/nnz + 776 B/row at dim 64.
#include "common.cuh"

namespace tagrec {

enum { EPI_PLAIN = 0, EPI_FWD = 1, EPI_BWD = 2, EPI_BWD0 = 3 };

struct Epi {
    Mirror my, macc;        // mirrors of y / acc (n == 0: local only)
    float* y;               // output rows (plain / fwd: raw layer, bwd: g_out)
    const float* x0;        // fwd, first layer: source of the accumulator (E0)
    float* acc;             // fwd: running sum
    const float* e_k;       // bwd: raw layer k
    const float* g_final;   // bwd: gradient w.r.t. the final (mean) table
    const float* reg_grad;  // bwd0: optional regulariser gradient (may alias y)
    const float* upstream;  // bwd: optional 2 floats {d/dloss, d/dreg} on device
    const uint8_t* src_nz;  // optional byte per SOURCE row: 0 = the row of x is all-zero, its gather is skipped
    float scale;            // plain: beta | fwd: final_scale | bwd: 1/(L+1)
    int first, last;
};

template <int LPR>
__device__ __forceinline__ float sub_sum(float v, unsigned mask) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, LPR);
    return v;
}

// Tuning knobs (compile-time; defaults = the configuration measured best on B200, see profiles/):
//   SPMM_U     gathers issued back to back per lane before the first FMA (rows in flight per sub-warp)
//   SPMM_WPB   warps per block (small blocks: a block retires as soon as its few rows are done)
//   SPMM_MINB  __launch_bounds__ min blocks per SM (caps registers so that SPMM_WPB*SPMM_MINB warps are resident)
#ifndef SPMM_U
#define SPMM_U 8
#endif
#ifndef SPMM_WPB
#define SPMM_WPB 2
#endif
#ifndef SPMM_MINB
#define SPMM_MINB 16
#endif

// acc = sum_{j in [begin,end), chunk(j) == part (mod nparts)} val[j] * x[col[j]]   for this lane's float4 slice.
template <int LPR, bool MASKED = false>
__device__ __forceinline__ float4 gather_rows(const int32_t* __restrict__ col, const float* __restrict__ val,
                                              const float4* __restrict__ x4, int64_t begin, int64_t end, int part,
                                              int nparts, int sl, unsigned mask,
                                              const uint8_t* __restrict__ src_nz = nullptr) {
    constexpr int U = SPMM_U < LPR ? SPMM_U : LPR;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int32_t* __restrict__ colp = col + begin;
    const float* __restrict__ valp = val + begin;
    const int len = (int)(end - begin);          // a row (or chunk) never exceeds 2^31 entries
    const int step = nparts * LPR;
    int base = part * LPR;
    int c = 0;
    float v = 0.f;
    // Zero-row skipping (backward tables that are non-zero only near the batch): the lane that loaded a column id
    // also looks its row up in the byte map and parks the answer in the id's sign bit, so the broadcast below needs no
    // extra shuffle and the 256 B gather of an all-zero row is never issued.
    if (base + sl < len) {
        c = __ldcs(colp + base + sl);
        v = __ldcs(valp + base + sl);
        if (MASKED && !__ldg(src_nz + c)) c |= (int)0x80000000;
    }
    const float4* __restrict__ xs = x4 + sl;
    while (base < len) {
        const int nbase = base + step;
        int cn = 0;
        float vn = 0.f;
        if (nbase + sl < len) {  // prefetch the next index chunk while this chunk's rows are in flight
            cn = __ldcs(colp + nbase + sl);
            vn = __ldcs(valp + nbase + sl);
            if (MASKED && !__ldg(src_nz + cn)) cn |= (int)0x80000000;
        }
        const int cnt = min(LPR, len - base);
#pragma unroll
        for (int j0 = 0; j0 < LPR; j0 += U) {
            if (j0 < cnt) {
                float4 xv[U];
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const int cj = __shfl_sync(mask, c, j0 + j, LPR);
                    xv[j] = (j0 + j < cnt && (!MASKED || cj >= 0)) ? ldg4(xs + (int64_t)cj * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < U; ++j) fma4(acc, __shfl_sync(mask, v, j0 + j, LPR), xv[j]);
            }
        }
        c = cn;
        v = vn;
        base = nbase;
    }
    return acc;
}

template <int LPR, int EPI>
__device__ __forceinline__ void epilogue(const Epi& ep, int64_t r, float4 acc, int sl, unsigned mask) {
    const int64_t o = r * LPR + sl;
    if (EPI == EPI_PLAIN) {
        float4* y4 = reinterpret_cast<float4*>(ep.y);
        if (ep.scale != 0.f) {
            const float4 old = y4[o];
            acc.x = fmaf(ep.scale, old.x, acc.x);
            acc.y = fmaf(ep.scale, old.y, acc.y);
            acc.z = fmaf(ep.scale, old.z, acc.z);
            acc.w = fmaf(ep.scale, old.w, acc.w);
        }
        y4[o] = acc;
    } else if (EPI == EPI_FWD) {
        // lightgcn.py:55-60: raw layer propagates, normalised copy joins the mean
        const float ss = sub_sum<LPR>(dot4(acc, acc), mask);
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);
        store_row(ep.y, ep.my, o, acc);
        float4* a4 = reinterpret_cast<float4*>(ep.acc);
        float4 a = ep.first ? __ldg(reinterpret_cast<const float4*>(ep.x0) + o) : a4[o];
        a.x += acc.x / nrm;
        a.y += acc.y / nrm;
        a.z += acc.z / nrm;
        a.w += acc.w / nrm;
        if (ep.last) {
            a.x *= ep.scale;
            a.y *= ep.scale;
            a.z *= ep.scale;
            a.w *= ep.scale;
        }
        store_row(ep.acc, ep.macc, o, a);
    } else {
        const float up0 = ep.upstream ? __ldg(ep.upstream) : 1.f;
        const float s = ep.scale * up0;
        float4 g = __ldg(reinterpret_cast<const float4*>(ep.g_final) + o);
        g.x *= s;
        g.y *= s;
        g.z *= s;
        g.w *= s;
        float4 out;
        if (EPI == EPI_BWD) {
            // Jacobian of e / max(||e||, eps) applied to g
            const float4 e = __ldg(reinterpret_cast<const float4*>(ep.e_k) + o);
            const float ss = sub_sum<LPR>(dot4(e, e), mask);
            const float dt = sub_sum<LPR>(dot4(e, g), mask);
            const float nrm = sqrtf(ss);
            if (nrm >= 1e-12f) {
                const float proj = dt / nrm;
                out.x = (g.x - (e.x / nrm) * proj) / nrm + acc.x;
                out.y = (g.y - (e.y / nrm) * proj) / nrm + acc.y;
                out.z = (g.z - (e.z / nrm) * proj) / nrm + acc.z;
                out.w = (g.w - (e.w / nrm) * proj) / nrm + acc.w;
            } else {
                out.x = g.x / 1e-12f + acc.x;
                out.y = g.y / 1e-12f + acc.y;
                out.z = g.z / 1e-12f + acc.z;
                out.w = g.w / 1e-12f + acc.w;
            }
        } else {
            out.x = g.x + acc.x;
            out.y = g.y + acc.y;
            out.z = g.z + acc.z;
            out.w = g.w + acc.w;
            if (ep.reg_grad) {
                const float up1 = ep.upstream ? __ldg(ep.upstream + 1) : 1.f;
                const float4 rg = reinterpret_cast<const float4*>(ep.reg_grad)[o];
                out.x = fmaf(up1, rg.x, out.x);
                out.y = fmaf(up1, rg.y, out.y);
                out.z = fmaf(up1, rg.z, out.z);
                out.w = fmaf(up1, rg.w, out.w);
            }
        }
        store_row(ep.y, ep.my, o, out);
    }
}

template <int LPR>
__device__ __forceinline__ float4 combine_subs(float4 p) {
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
        p.x += __shfl_xor_sync(0xffffffffu, p.x, o);
        p.y += __shfl_xor_sync(0xffffffffu, p.y, o);
        p.z += __shfl_xor_sync(0xffffffffu, p.z, o);
        p.w += __shfl_xor_sync(0xffffffffu, p.w, o);
    }
    return p;
}

constexpr int kWarpsPerBlock = SPMM_WPB;

template <int LPR, int EPI, bool MASKED = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, SPMM_MINB)
spmm_kernel(tagrec_csr_t a, const float4* __restrict__ x4, Epi ep, int n_long_blocks, int gather) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int sub = lane / LPR;
    const int sl = lane % LPR;
    const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));

    if ((int)blockIdx.x < n_long_blocks) {
        // ---- chunks of long / column-blocked rows: partial sums meet in the row's scratch row ----
        const bool per_sub = a.chunk_lanes != 0 && RPW > 1;      // one sub-warp per chunk (short, column-window pieces)
        const int64_t unit = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
        const int64_t item = per_sub ? unit * RPW + sub : unit;
        if (item >= a.n_items) return;
        const int slot = __ldg(a.item_slot + item);
        const int64_t b = __ldg(a.item_begin + item), e = __ldg(a.item_end + item);
        float4 p;
        if (per_sub) {
            p = gather_rows<LPR, MASKED>(a.col, a.val, x4, b, e, 0, 1, sl, mask, ep.src_nz);
        } else {
            p = gather_rows<LPR, MASKED>(a.col, a.val, x4, b, e, sub, RPW, sl, mask, ep.src_nz);
            p = combine_subs<LPR>(p);
        }
        const bool writer = per_sub || sub == 0;
        float4* scr = reinterpret_cast<float4*>(a.long_scratch) + (int64_t)slot * LPR + sl;
        if (writer) red_add4(scr, p);
        __threadfence();
        __syncwarp(per_sub ? mask : 0xffffffffu);
        const int64_t r = __ldg(a.long_rows + slot);
        int nchunks;
        if (a.long_nchunks) {
            nchunks = __ldg(a.long_nchunks + slot);
        } else {
            const int64_t deg = __ldg(a.rowptr + r + 1) - __ldg(a.rowptr + r);
            nchunks = (int)((deg + a.long_chunk - 1) / a.long_chunk);
        }
        int ticket = 0;
        if (per_sub) {
            if (sl == 0) ticket = atomicAdd(a.long_counter + slot, 1);
            ticket = __shfl_sync(mask, ticket, 0, LPR);
        } else {
            if (lane == 0) ticket = atomicAdd(a.long_counter + slot, 1);
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
        }
        if (ticket != nchunks - 1) return;
        __threadfence();
        if (writer) {  // last piece: complete row sits in the scratch row; run the fused epilogue, leave it zeroed
            const float4 tot = __ldcg(scr);
            __stcg(scr, make_float4(0.f, 0.f, 0.f, 0.f));
            if (sl == 0) a.long_counter[slot] = 0;
            epilogue<LPR, EPI>(ep, r + a.row_offset, tot, sl, mask);
        }
        return;
    }

    // ---- RPW consecutive rows per warp ----
    const int64_t w = ((int64_t)blockIdx.x - n_long_blocks) * kWarpsPerBlock + wib;
    if (w * RPW >= a.n_rows) return;
    const int64_t r = w * RPW + sub;
    const bool valid = r < a.n_rows;
    int64_t s = 0, e = 0;
    if (valid) {
        s = __ldg(a.rowptr + r);
        e = __ldg(a.rowptr + r + 1);
    }
    const int64_t long_thr = (r >= a.blocked_row_begin && a.blocked_min_deg > 0) ? a.blocked_min_deg : a.long_row;
    const bool is_long = gather && (e - s) > long_thr;
    if (is_long || !gather) e = s;  // long rows are produced by the chunk blocks above

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gather) {
        bool side_by_side = true;
        if (RPW > 1) {
            const int64_t deg = e - s;
            int64_t par = 0, seq = 0;
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int64_t d = __shfl_sync(0xffffffffu, deg, i * LPR);
                par = max(par, (d + LPR - 1) / LPR);
                seq += (d + 31) / 32;
            }
            side_by_side = par <= seq;
        }
        if (side_by_side) {
            acc = gather_rows<LPR, MASKED>(a.col, a.val, x4, s, e, 0, 1, sl, mask, ep.src_nz);
        } else {
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int64_t si = __shfl_sync(0xffffffffu, s, i * LPR);
                const int64_t ei = __shfl_sync(0xffffffffu, e, i * LPR);
                float4 p = gather_rows<LPR, MASKED>(a.col, a.val, x4, si, ei, sub, RPW, sl, mask, ep.src_nz);
                p = combine_subs<LPR>(p);
                if (sub == i) acc = p;
            }
        }
    }
    if (valid && !is_long) epilogue<LPR, EPI>(ep, r + a.row_offset, acc, sl, mask);
}

template <int EPI>
static int launch(const tagrec_csr_t* a, const float* x, const Epi& ep, int dim, int gather, void* stream) {
    TAGREC_REQUIRE(a && a->rowptr && a->n_rows >= 0, "csr descriptor missing");
    TAGREC_REQUIRE(!gather || (a->col && a->val && x), "csr arrays / source table missing");
    TAGREC_REQUIRE(dim == 32 || dim == 64 || dim == 128, "dim must be 32, 64 or 128");
    if (a->n_rows == 0) return TAGREC_OK;
    const int64_t n_items = gather ? a->n_items : 0;
    if (n_items > 0)
        TAGREC_REQUIRE(a->long_rows && a->item_slot && a->item_begin && a->item_end && a->long_scratch &&
                           a->long_counter, "long-row plan arrays missing");
    const int lpr = dim / 4, rpw = 32 / lpr;
    const int64_t chunks_per_block = (int64_t)kWarpsPerBlock * ((a->chunk_lanes != 0 && rpw > 1) ? rpw : 1);
    const int64_t long_blocks = (n_items + chunks_per_block - 1) / chunks_per_block;
    const int64_t row_blocks = (a->n_rows + (int64_t)rpw * kWarpsPerBlock - 1) / ((int64_t)rpw * kWarpsPerBlock);
    const int64_t grid = long_blocks + row_blocks;
    TAGREC_REQUIRE(grid < (1ll << 31), "grid too large");
    tagrec_csr_t d = *a;
    d.n_items = n_items;
    if (d.long_row <= 0) d.long_row = TAGREC_LONG_ROW;
    if (d.long_chunk <= 0) d.long_chunk = TAGREC_LONG_CHUNK;
    if (d.blocked_min_deg <= 0) d.blocked_row_begin = INT64_MAX;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const dim3 block(kWarpsPerBlock * 32);
    if (lpr == 16 && ep.src_nz && gather && (EPI == EPI_BWD || EPI == EPI_BWD0)) {
        TAGREC_LAUNCH((spmm_kernel<16, EPI, true>), (unsigned)grid, block, 0, stream, d, x4, ep, (int)long_blocks, gather);
    } else if (lpr == 16) {
        TAGREC_LAUNCH((spmm_kernel<16, EPI>), (unsigned)grid, block, 0, stream, d, x4, ep, (int)long_blocks, gather);
    } else if (lpr == 8) {
        TAGREC_LAUNCH((spmm_kernel<8, EPI>), (unsigned)grid, block, 0, stream, d, x4, ep, (int)long_blocks, gather);
    } else {
        TAGREC_LAUNCH((spmm_kernel<32, EPI>), (unsigned)grid, block, 0, stream, d, x4, ep, (int)long_blocks, gather);
    }
    return TAGREC_OK;
}

}  // namespace tagrec

using namespace tagrec;

extern "C" int tagrec_spmm(const tagrec_csr_t* a, const float* x, float* y, int dim, float beta, void* stream) {
    TAGREC_REQUIRE(y, "y is null");
    Epi ep{};
    ep.y = y;
    ep.scale = beta;
    return launch<EPI_PLAIN>(a, x, ep, dim, 1, stream);
}

extern "C" int tagrec_lightgcn_fwd_layer(const tagrec_csr_t* a, const float* x, float* y, float* acc, int dim,
                                         int first, int last, float final_scale, void* stream) {
    return tagrec_lightgcn_fwd_layer_p2p(a, x, y, acc, dim, first, last, final_scale, nullptr, nullptr, stream);
}

extern "C" int tagrec_lightgcn_fwd_layer_p2p(const tagrec_csr_t* a, const float* x, float* y, float* acc, int dim,
                                             int first, int last, float final_scale, const tagrec_mirror_t* y_mirror,
                                             const tagrec_mirror_t* acc_mirror, void* stream) {
    TAGREC_REQUIRE(y && acc, "y/acc is null");
    Epi ep{};
    if (int rc = set_mirror(ep.my, y_mirror)) return rc;
    if (int rc = set_mirror(ep.macc, acc_mirror)) return rc;
    ep.y = y;
    ep.x0 = x;
    ep.acc = acc;
    ep.first = first;
    ep.last = last;
    ep.scale = final_scale;
    return launch<EPI_FWD>(a, x, ep, dim, 1, stream);
}

extern "C" int tagrec_lightgcn_bwd_layer(const tagrec_csr_t* a, const float* g_next, const float* e_k,
                                         const float* g_final, const float* reg_grad, const float* upstream,
                                         float inv_layers, float* g_out, int dim, void* stream) {
    return tagrec_lightgcn_bwd_layer_p2p(a, g_next, e_k, g_final, reg_grad, upstream, inv_layers, g_out, dim, nullptr,
                                         stream);
}

extern "C" int tagrec_lightgcn_bwd_layer_p2p(const tagrec_csr_t* a, const float* g_next, const float* e_k,
                                             const float* g_final, const float* reg_grad, const float* upstream,
                                             float inv_layers, float* g_out, int dim,
                                             const tagrec_mirror_t* out_mirror, void* stream) {
    return tagrec_lightgcn_bwd_layer_ex(a, g_next, nullptr, e_k, g_final, reg_grad, upstream, inv_layers, g_out, dim,
                                        out_mirror, stream);
}

extern "C" int tagrec_lightgcn_bwd_layer_ex(const tagrec_csr_t* a, const float* g_next, const uint8_t* g_next_nz,
                                            const float* e_k, const float* g_final, const float* reg_grad,
                                            const float* upstream, float inv_layers, float* g_out, int dim,
                                            const tagrec_mirror_t* out_mirror, void* stream) {
    TAGREC_REQUIRE(g_final && g_out, "g_final/g_out is null");
    Epi ep{};
    if (int rc = set_mirror(ep.my, out_mirror)) return rc;
    ep.y = g_out;
    ep.e_k = e_k;
    ep.g_final = g_final;
    ep.reg_grad = reg_grad;
    ep.upstream = upstream;
    ep.scale = inv_layers;
    ep.src_nz = (g_next && dim == 64) ? g_next_nz : nullptr;      // masked instantiation exists for dim 64
    const int gather = g_next != nullptr;
    if (e_k) return launch<EPI_BWD>(a, g_next, ep, dim, gather, stream);
    return launch<EPI_BWD0>(a, g_next, ep, dim, gather, stream);
}

namespace tagrec {
// nz[r] = 1 if any element of row r of a [n, dim] table is non-zero (one float4 per thread, dim/4 threads per row).
template <int LPR>
__global__ void __launch_bounds__(256) row_nonzero_kernel(const float4* __restrict__ t, int64_t n, uint8_t* __restrict__ nz) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < n * LPR;
    const float4 v = valid ? __ldcs(t + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned b = __ballot_sync(0xffffffffu, v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f);
    const int lane = threadIdx.x & 31;
    if (valid && lane % LPR == 0) {
        const unsigned grp = LPR == 32 ? 0xffffffffu : (((1u << LPR) - 1u) << lane);
        nz[idx / LPR] = (b & grp) ? 1 : 0;
    }
}
}  // namespace tagrec

extern "C" int tagrec_row_nonzero(const float* table, int64_t n, int dim, uint8_t* nz, void* stream) {
    TAGREC_REQUIRE(table && nz, "null pointer");
    TAGREC_REQUIRE(dim == 32 || dim == 64 || dim == 128, "dim must be 32, 64 or 128");
    if (n == 0) return TAGREC_OK;
    const int lpr = dim / 4;
    const unsigned grid = (unsigned)((n * lpr + 255) / 256);
    const float4* t4 = reinterpret_cast<const float4*>(table);
    if (lpr == 16) { TAGREC_LAUNCH((row_nonzero_kernel<16>), grid, 256, 0, stream, t4, n, nz); }
    else if (lpr == 8) { TAGREC_LAUNCH((row_nonzero_kernel<8>), grid, 256, 0, stream, t4, n, nz); }
    else { TAGREC_LAUNCH((row_nonzero_kernel<32>), grid, 256, 0, stream, t4, n, nz); }
    return TAGREC_OK;
}

namespace tagrec {
// First backward table of a BPR step, sparse form.  dL/dF is non-zero on the batch's nodes only, so
// G_L = nb(gY, E^L) is too: instead of an elementwise pass over all N rows (3 tables x N x dim floats), only the listed
// rows are produced — into a table that is all-zero otherwise — and flagged in the byte map the masked K1 launch reads.
// Duplicate nodes write identical values.  Rows outside [row_lo, row_hi) (another rank's block) are flagged, not written.
template <int LPR>
__global__ void __launch_bounds__(256)
bwd_first_sparse_kernel(const int64_t* __restrict__ nodes, int64_t n_nodes, int64_t row_lo, int64_t row_hi,
                        const float4* __restrict__ e_k, const float4* __restrict__ g_final,
                        const float* __restrict__ upstream, float inv_layers, float4* __restrict__ g_out,
                        uint8_t* __restrict__ nz) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / LPR;
    const int sl = (int)(t % LPR);
    const int lane = threadIdx.x & 31;
    const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((lane / LPR) * LPR));
    if (i >= n_nodes) return;
    const int64_t v = __ldg(nodes + i);
    if (nz && sl == 0) nz[v] = 1;
    if (v < row_lo || v >= row_hi) return;
    const float s = inv_layers * (upstream ? __ldg(upstream) : 1.f);
    const int64_t o = v * LPR + sl;
    float4 g = __ldg(g_final + o);
    g.x *= s; g.y *= s; g.z *= s; g.w *= s;
    const float4 e = __ldg(e_k + o);
    const float ss = sub_sum<LPR>(dot4(e, e), mask);
    const float dt = sub_sum<LPR>(dot4(e, g), mask);
    const float nrm = sqrtf(ss);
    float4 out;
    if (nrm >= 1e-12f) {
        const float proj = dt / nrm;
        out.x = (g.x - (e.x / nrm) * proj) / nrm;
        out.y = (g.y - (e.y / nrm) * proj) / nrm;
        out.z = (g.z - (e.z / nrm) * proj) / nrm;
        out.w = (g.w - (e.w / nrm) * proj) / nrm;
    } else {
        out.x = g.x / 1e-12f; out.y = g.y / 1e-12f; out.z = g.z / 1e-12f; out.w = g.w / 1e-12f;
    }
    g_out[o] = out;
}

template <int LPR>
__global__ void __launch_bounds__(256)
rows_zero_kernel(const int64_t* __restrict__ nodes, int64_t n_nodes, float4* __restrict__ table, uint8_t* __restrict__ nz) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / LPR;
    if (i >= n_nodes) return;
    const int64_t v = __ldg(nodes + i);
    if (table) table[v * LPR + (t % LPR)] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nz && t % LPR == 0) nz[v] = 0;
}
}  // namespace tagrec

extern "C" int tagrec_lightgcn_bwd_first_sparse(const int64_t* nodes, int64_t n_nodes, int64_t row_lo, int64_t row_hi,
                                                const float* e_k, const float* g_final, const float* upstream,
                                                float inv_layers, float* g_out, uint8_t* nz, int dim, void* stream) {
    TAGREC_REQUIRE(nodes && e_k && g_final && g_out, "null pointer");
    TAGREC_REQUIRE(dim == 32 || dim == 64 || dim == 128, "dim must be 32, 64 or 128");
    if (n_nodes == 0) return TAGREC_OK;
    const int lpr = dim / 4;
    const unsigned grid = (unsigned)((n_nodes * lpr + 255) / 256);
    const float4* e4 = reinterpret_cast<const float4*>(e_k);
    const float4* g4 = reinterpret_cast<const float4*>(g_final);
    float4* o4 = reinterpret_cast<float4*>(g_out);
    if (lpr == 16) { TAGREC_LAUNCH((bwd_first_sparse_kernel<16>), grid, 256, 0, stream, nodes, n_nodes, row_lo, row_hi, e4, g4, upstream, inv_layers, o4, nz); }
    else if (lpr == 8) { TAGREC_LAUNCH((bwd_first_sparse_kernel<8>), grid, 256, 0, stream, nodes, n_nodes, row_lo, row_hi, e4, g4, upstream, inv_layers, o4, nz); }
    else { TAGREC_LAUNCH((bwd_first_sparse_kernel<32>), grid, 256, 0, stream, nodes, n_nodes, row_lo, row_hi, e4, g4, upstream, inv_layers, o4, nz); }
    return TAGREC_OK;
}

extern "C" int tagrec_rows_zero(const int64_t* nodes, int64_t n_nodes, float* table, uint8_t* nz, int dim, void* stream) {
    TAGREC_REQUIRE(nodes && (table || nz), "null pointer");
    TAGREC_REQUIRE(dim == 32 || dim == 64 || dim == 128, "dim must be 32, 64 or 128");
    if (n_nodes == 0) return TAGREC_OK;
    const int lpr = dim / 4;
    const unsigned grid = (unsigned)((n_nodes * lpr + 255) / 256);
    float4* t4 = reinterpret_cast<float4*>(table);
    if (lpr == 16) { TAGREC_LAUNCH((rows_zero_kernel<16>), grid, 256, 0, stream, nodes, n_nodes, t4, nz); }
    else if (lpr == 8) { TAGREC_LAUNCH((rows_zero_kernel<8>), grid, 256, 0, stream, nodes, n_nodes, t4, nz); }
    else { TAGREC_LAUNCH((rows_zero_kernel<32>), grid, 256, 0, stream, nodes, n_nodes, t4, nz); }
    return TAGREC_OK;
}
