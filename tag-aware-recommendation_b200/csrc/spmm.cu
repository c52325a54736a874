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
    float* acc_in;          // optional, launches without a gather: the row's sum was accumulated HERE beforehand
                            // (tagrec_spmm_push_rows) — full-size table, global rows; consumed rows are zeroed again
    float scale;            // plain: beta | fwd: final_scale | bwd: 1/(L+1)
    int first, last;
    // bwd0 only, optional: the optimizer folded into the epilogue (tagrec_lightgcn_bwd_layer_adam)
    float* adam_p;          // parameter table the output rows are the gradient of (NULL: plain gradient output)
    float* adam_m;
    float* adam_v;
    Mirror mp;              // mirrors of adam_p
    float b1, b2, eps, wd, step_size, inv_sqrt_bc2;
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

// A lane's slice of a table row: V float4 (row = LPR * V float4; lane sl owns float4 sl, sl + LPR, ...: every one of its
// V loads is part of a coalesced LPR * 16-byte segment).  V = 1: 16 lanes per 64-float row (the round-1 mapping).
// V = 2: 8 lanes per 64-float row, four rows per warp — half the warp-instructions per stored entry (the index
// broadcast, address arithmetic and loop control are shared by two 128-bit loads and eight FMAs instead of one and four).
template <int V>
struct Slice {
    float4 v[V];
};

template <int V>
__device__ __forceinline__ Slice<V> zero_slice() {
    Slice<V> z;
#pragma unroll
    for (int i = 0; i < V; ++i) z.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return z;
}

// acc = sum_{j in [begin,end), chunk(j) == part (mod nparts)} val[j] * x[col[j]]   for this lane's slice.
// The gathers go out SPMM_U at a time and are then consumed in order.  A software-pipelined form (groups of 4, the next
// group's loads in flight while one is accumulated) was measured on the column-blocked 1 B-edge launch, where 69 % of the
// stalls are long-scoreboard: 76 ms per layer instead of 50 (fewer loads in flight, spills at the 64-register cap);
// not kept (profiles/r2_spmm_1b_ncu.md).
template <int LPR, int V, bool MASKED = false>
__device__ __forceinline__ Slice<V> gather_rows(const int32_t* __restrict__ col, const float* __restrict__ val,
                                                const float4* __restrict__ x4, int64_t begin, int64_t end, int part,
                                                int nparts, int sl, unsigned mask,
                                                const uint8_t* __restrict__ src_nz = nullptr) {
    constexpr int U0 = SPMM_U / V < 1 ? 1 : SPMM_U / V;      // entries in flight per lane (U0 * V 128-bit loads)
    constexpr int U = U0 < LPR ? U0 : LPR;
    constexpr int ROW4 = LPR * V;                            // float4 per table row
    Slice<V> acc = zero_slice<V>();
    const int32_t* __restrict__ colp = col + begin;
    const float* __restrict__ valp = val + begin;
    const int len = (int)(end - begin);          // a row (or chunk) never exceeds 2^31 entries
    const int step = nparts * LPR;
    int base = part * LPR;
    int c = 0;
    float v = 0.f;
    // Zero-row skipping (backward tables that are non-zero only near the batch): the lane that loaded a column id
    // also looks its row up in the byte map and parks the answer in the id's sign bit, so the broadcast below needs no
    // extra shuffle and the gather of an all-zero row is never issued.
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
                float4 xv[U][V];
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const int cj = __shfl_sync(mask, c, j0 + j, LPR);
                    const bool live = j0 + j < cnt && (!MASKED || cj >= 0);
                    const float4* __restrict__ src = xs + (int64_t)cj * ROW4;
#pragma unroll
                    for (int q = 0; q < V; ++q) xv[j][q] = live ? ldg4(src + q * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const float vj = __shfl_sync(mask, v, j0 + j, LPR);
#pragma unroll
                    for (int q = 0; q < V; ++q) fma4(acc.v[q], vj, xv[j][q]);
                }
            }
        }
        c = cn;
        v = vn;
        base = nbase;
    }
    return acc;
}

template <int LPR, int V>
__device__ __forceinline__ float slice_dot(const Slice<V>& a, const Slice<V>& b) {
    float d = dot4(a.v[0], b.v[0]);
#pragma unroll
    for (int q = 1; q < V; ++q) d += dot4(a.v[q], b.v[q]);
    return d;
}

template <int LPR, int V, int EPI>
__device__ __forceinline__ void epilogue(const Epi& ep, int64_t r, Slice<V> acc, int sl, unsigned mask) {
    const int64_t o0 = r * (LPR * V) + sl;           // float4 index of this lane's first slice element; + q * LPR
    if (EPI == EPI_PLAIN) {
        float4* y4 = reinterpret_cast<float4*>(ep.y);
#pragma unroll
        for (int q = 0; q < V; ++q) {
            float4 a = acc.v[q];
            if (ep.scale != 0.f) {
                const float4 old = y4[o0 + q * LPR];
                a.x = fmaf(ep.scale, old.x, a.x);
                a.y = fmaf(ep.scale, old.y, a.y);
                a.z = fmaf(ep.scale, old.z, a.z);
                a.w = fmaf(ep.scale, old.w, a.w);
            }
            y4[o0 + q * LPR] = a;
        }
    } else if (EPI == EPI_FWD) {
        // lightgcn.py:55-60: raw layer propagates, normalised copy joins the mean
        const float ss = sub_sum<LPR>(slice_dot<LPR, V>(acc, acc), mask);
        const float nrm = fmaxf(sqrtf(ss), 1e-12f);
        float4* a4 = reinterpret_cast<float4*>(ep.acc);
#pragma unroll
        for (int q = 0; q < V; ++q) {
            const int64_t o = o0 + q * LPR;
            store_row(ep.y, ep.my, o, acc.v[q]);
            float4 a = ep.first ? __ldg(reinterpret_cast<const float4*>(ep.x0) + o) : a4[o];
            a.x += acc.v[q].x / nrm;
            a.y += acc.v[q].y / nrm;
            a.z += acc.v[q].z / nrm;
            a.w += acc.v[q].w / nrm;
            if (ep.last) {
                a.x *= ep.scale;
                a.y *= ep.scale;
                a.z *= ep.scale;
                a.w *= ep.scale;
            }
            store_row(ep.acc, ep.macc, o, a);
        }
    } else {
        const float up0 = ep.upstream ? __ldg(ep.upstream) : 1.f;
        const float s = ep.scale * up0;
        Slice<V> g;
#pragma unroll
        for (int q = 0; q < V; ++q) {
            g.v[q] = __ldg(reinterpret_cast<const float4*>(ep.g_final) + o0 + q * LPR);
            g.v[q].x *= s;
            g.v[q].y *= s;
            g.v[q].z *= s;
            g.v[q].w *= s;
        }
        if (EPI == EPI_BWD) {
            // Jacobian of e / max(||e||, eps) applied to g
            Slice<V> e;
#pragma unroll
            for (int q = 0; q < V; ++q) e.v[q] = __ldg(reinterpret_cast<const float4*>(ep.e_k) + o0 + q * LPR);
            const float ss = sub_sum<LPR>(slice_dot<LPR, V>(e, e), mask);
            const float dt = sub_sum<LPR>(slice_dot<LPR, V>(e, g), mask);
            const float nrm = sqrtf(ss);
#pragma unroll
            for (int q = 0; q < V; ++q) {
                float4 out;
                if (nrm >= 1e-12f) {
                    const float proj = dt / nrm;
                    out.x = (g.v[q].x - (e.v[q].x / nrm) * proj) / nrm + acc.v[q].x;
                    out.y = (g.v[q].y - (e.v[q].y / nrm) * proj) / nrm + acc.v[q].y;
                    out.z = (g.v[q].z - (e.v[q].z / nrm) * proj) / nrm + acc.v[q].z;
                    out.w = (g.v[q].w - (e.v[q].w / nrm) * proj) / nrm + acc.v[q].w;
                } else {
                    out.x = g.v[q].x / 1e-12f + acc.v[q].x;
                    out.y = g.v[q].y / 1e-12f + acc.v[q].y;
                    out.z = g.v[q].z / 1e-12f + acc.v[q].z;
                    out.w = g.v[q].w / 1e-12f + acc.v[q].w;
                }
                store_row(ep.y, ep.my, o0 + q * LPR, out);
            }
        } else {
            const float up1 = (ep.reg_grad && ep.upstream) ? __ldg(ep.upstream + 1) : 1.f;
#pragma unroll
            for (int q = 0; q < V; ++q) {
                float4 out;
                out.x = g.v[q].x + acc.v[q].x;
                out.y = g.v[q].y + acc.v[q].y;
                out.z = g.v[q].z + acc.v[q].z;
                out.w = g.v[q].w + acc.v[q].w;
                if (ep.reg_grad) {
                    const float4 rg = reinterpret_cast<const float4*>(ep.reg_grad)[o0 + q * LPR];
                    out.x = fmaf(up1, rg.x, out.x);
                    out.y = fmaf(up1, rg.y, out.y);
                    out.z = fmaf(up1, rg.z, out.z);
                    out.w = fmaf(up1, rg.w, out.w);
                }
                if (ep.adam_p) {
                    // the gradient row is consumed where it is produced: Adam on this row, new parameters to every rank
                    const int64_t o = o0 + q * LPR;
                    float4 pp = reinterpret_cast<const float4*>(ep.adam_p)[o];
                    float4 mm = reinterpret_cast<const float4*>(ep.adam_m)[o];
                    float4 vv = reinterpret_cast<const float4*>(ep.adam_v)[o];
                    adam_update(pp.x, out.x, mm.x, vv.x, ep.b1, ep.b2, ep.eps, ep.wd, ep.step_size, ep.inv_sqrt_bc2);
                    adam_update(pp.y, out.y, mm.y, vv.y, ep.b1, ep.b2, ep.eps, ep.wd, ep.step_size, ep.inv_sqrt_bc2);
                    adam_update(pp.z, out.z, mm.z, vv.z, ep.b1, ep.b2, ep.eps, ep.wd, ep.step_size, ep.inv_sqrt_bc2);
                    adam_update(pp.w, out.w, mm.w, vv.w, ep.b1, ep.b2, ep.eps, ep.wd, ep.step_size, ep.inv_sqrt_bc2);
                    store_row(ep.adam_p, ep.mp, o, pp);
                    reinterpret_cast<float4*>(ep.adam_m)[o] = mm;
                    reinterpret_cast<float4*>(ep.adam_v)[o] = vv;
                    if (ep.y) reinterpret_cast<float4*>(ep.y)[o] = out;
                    continue;
                }
                store_row(ep.y, ep.my, o0 + q * LPR, out);
            }
        }
    }
}

template <int LPR, int V>
__device__ __forceinline__ Slice<V> combine_subs(Slice<V> p) {
#pragma unroll
    for (int q = 0; q < V; ++q) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
            p.v[q].x += __shfl_xor_sync(0xffffffffu, p.v[q].x, o);
            p.v[q].y += __shfl_xor_sync(0xffffffffu, p.v[q].y, o);
            p.v[q].z += __shfl_xor_sync(0xffffffffu, p.v[q].z, o);
            p.v[q].w += __shfl_xor_sync(0xffffffffu, p.v[q].w, o);
        }
    }
    return p;
}

constexpr int kWarpsPerBlock = SPMM_WPB;

template <int LPR, int V, int EPI, bool MASKED = false>
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
        if (a.row_sel && !__ldg(a.row_sel + __ldg(a.long_rows + slot))) return;      // row-subset launch: row not listed
        const int64_t b = __ldg(a.item_begin + item), e = __ldg(a.item_end + item);
        Slice<V> p;
        if (per_sub) {
            p = gather_rows<LPR, V, MASKED>(a.col, a.val, x4, b, e, 0, 1, sl, mask, ep.src_nz);
        } else {
            p = gather_rows<LPR, V, MASKED>(a.col, a.val, x4, b, e, sub, RPW, sl, mask, ep.src_nz);
            p = combine_subs<LPR, V>(p);
        }
        const bool writer = per_sub || sub == 0;
        float4* scr = reinterpret_cast<float4*>(a.long_scratch) + (int64_t)slot * (LPR * V) + sl;
        if (writer) {
#pragma unroll
            for (int q = 0; q < V; ++q) red_add4(scr + q * LPR, p.v[q]);
        }
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
            Slice<V> tot;
#pragma unroll
            for (int q = 0; q < V; ++q) {
                tot.v[q] = __ldcg(scr + q * LPR);
                __stcg(scr + q * LPR, make_float4(0.f, 0.f, 0.f, 0.f));
            }
            if (sl == 0) a.long_counter[slot] = 0;
            epilogue<LPR, V, EPI>(ep, r + a.row_offset, tot, sl, mask);
        }
        return;
    }

    // ---- RPW consecutive rows per warp ----
    const int64_t w = ((int64_t)blockIdx.x - n_long_blocks) * kWarpsPerBlock + wib;
    if (w * RPW >= a.n_rows) return;
    const int64_t ridx = w * RPW + sub;
    const bool in_range = ridx < a.n_rows;
    // row-subset launches: the n_rows listed rows instead of rows 0 .. n_rows-1 (negative entries are skipped: the
    // caller blanks duplicates and rows of other blocks instead of compacting the list)
    const int64_t r = (a.row_list && in_range) ? (int64_t)__ldg(a.row_list + ridx) : ridx;
    const bool valid = in_range && r >= 0;
    int64_t s = 0, e = 0;
    if (valid) {
        s = __ldg(a.rowptr + r);
        e = __ldg(a.rowptr + r + 1);
    }
    const int64_t long_thr = (r >= a.blocked_row_begin && a.blocked_min_deg > 0) ? a.blocked_min_deg : a.long_row;
    const bool is_long = gather && (e - s) > long_thr;
    if (is_long || !gather) e = s;  // long rows are produced by the chunk blocks above

    Slice<V> acc = zero_slice<V>();
    if (!gather && ep.acc_in && valid) {
        // the sum of this row was pushed into the accumulation table by its (few) non-zero sources
        float4* in4 = reinterpret_cast<float4*>(ep.acc_in) + (r + a.row_offset) * (LPR * V) + sl;
        bool any = false;
#pragma unroll
        for (int q = 0; q < V; ++q) {
            acc.v[q] = in4[q * LPR];
            any = any || acc.v[q].x != 0.f || acc.v[q].y != 0.f || acc.v[q].z != 0.f || acc.v[q].w != 0.f;
        }
        if (any) {
#pragma unroll
            for (int q = 0; q < V; ++q) in4[q * LPR] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
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
            acc = gather_rows<LPR, V, MASKED>(a.col, a.val, x4, s, e, 0, 1, sl, mask, ep.src_nz);
        } else {
#pragma unroll
            for (int i = 0; i < RPW; ++i) {
                const int64_t si = __shfl_sync(0xffffffffu, s, i * LPR);
                const int64_t ei = __shfl_sync(0xffffffffu, e, i * LPR);
                Slice<V> p = gather_rows<LPR, V, MASKED>(a.col, a.val, x4, si, ei, sub, RPW, sl, mask, ep.src_nz);
                p = combine_subs<LPR, V>(p);
                if (sub == i) acc = p;
            }
        }
    }
    if (valid && !is_long) epilogue<LPR, V, EPI>(ep, r + a.row_offset, acc, sl, mask);
}

// dim 64: lanes per row.  16 = one float4 per lane (round 1); 8 = two float4 per lane, four rows per warp.
#ifndef SPMM_LANES64
#define SPMM_LANES64 16
#endif

template <int EPI>
static int launch(const tagrec_csr_t* a, const float* x, const Epi& ep, int dim, int gather, void* stream) {
    TAGREC_REQUIRE(a && a->rowptr && a->n_rows >= 0, "csr descriptor missing");
    TAGREC_REQUIRE(!gather || (a->col && a->val && x), "csr arrays / source table missing");
    TAGREC_REQUIRE(dim == 32 || dim == 64 || dim == 128, "dim must be 32, 64 or 128");
    const int64_t n_items = gather ? a->n_items : 0;
    if (a->n_rows == 0 && n_items == 0) return TAGREC_OK;
    if (n_items > 0)
        TAGREC_REQUIRE(a->long_rows && a->item_slot && a->item_begin && a->item_end && a->long_scratch &&
                           a->long_counter, "long-row plan arrays missing");
    const int lpr = dim == 64 ? SPMM_LANES64 : dim / 4, rpw = 32 / lpr;
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
    constexpr int L64 = SPMM_LANES64, V64 = 16 / SPMM_LANES64;
    static_assert(L64 == 16 || L64 == 8, "SPMM_LANES64 must be 16 or 8");
    if (dim == 64 && ep.src_nz && gather && (EPI == EPI_BWD || EPI == EPI_BWD0)) {
        TAGREC_LAUNCH((spmm_kernel<L64, V64, EPI, true>), (unsigned)grid, block, 0, stream, d, x4, ep, (int)long_blocks, gather);
    } else if (dim == 64) {
        TAGREC_LAUNCH((spmm_kernel<L64, V64, EPI>), (unsigned)grid, block, 0, stream, d, x4, ep, (int)long_blocks, gather);
    } else if (dim == 32) {
        TAGREC_LAUNCH((spmm_kernel<8, 1, EPI>), (unsigned)grid, block, 0, stream, d, x4, ep, (int)long_blocks, gather);
    } else {
        TAGREC_LAUNCH((spmm_kernel<32, 1, EPI>), (unsigned)grid, block, 0, stream, d, x4, ep, (int)long_blocks, gather);
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

extern "C" int tagrec_lightgcn_bwd_layer_adam(const tagrec_csr_t* a, const float* g_next, const uint8_t* g_next_nz,
                                              const float* g_final, const float* reg_grad, const float* upstream,
                                              float inv_layers, float* g_out, int dim, const tagrec_adam_t* adam,
                                              void* stream) {
    TAGREC_REQUIRE(g_final && adam, "g_final/adam is null");
    TAGREC_REQUIRE(adam->param && adam->exp_avg && adam->exp_avg_sq, "adam: null table");
    TAGREC_REQUIRE(adam->step >= 1, "adam: step counts from 1");
    TAGREC_REQUIRE((((uintptr_t)adam->param | (uintptr_t)adam->exp_avg | (uintptr_t)adam->exp_avg_sq) & 15) == 0,
                   "adam: tables must be 16-byte aligned");
    Epi ep{};
    if (int rc = set_mirror(ep.mp, &adam->param_mirror)) return rc;
    ep.y = g_out;
    ep.g_final = g_final;
    ep.reg_grad = reg_grad;
    ep.upstream = upstream;
    ep.scale = inv_layers;
    ep.src_nz = (g_next && dim == 64) ? g_next_nz : nullptr;
    ep.adam_p = adam->param;
    ep.adam_m = adam->exp_avg;
    ep.adam_v = adam->exp_avg_sq;
    ep.b1 = adam->beta1;
    ep.b2 = adam->beta2;
    ep.eps = adam->eps;
    ep.wd = adam->weight_decay;
    const double bc1 = 1.0 - pow((double)adam->beta1, (double)adam->step);      // same host arithmetic as tagrec_adam_step
    const double bc2 = 1.0 - pow((double)adam->beta2, (double)adam->step);
    ep.step_size = (float)((double)adam->lr / bc1);
    ep.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    return launch<EPI_BWD0>(a, g_next, ep, dim, g_next != nullptr, stream);
}

extern "C" int tagrec_lightgcn_bwd_layer_acc(const tagrec_csr_t* a, float* acc_in, const float* e_k, const float* g_final,
                                             const float* upstream, float inv_layers, float* g_out, int dim,
                                             const tagrec_mirror_t* out_mirror, void* stream) {
    TAGREC_REQUIRE(acc_in && e_k && g_final && g_out, "null pointer");
    Epi ep{};
    if (int rc = set_mirror(ep.my, out_mirror)) return rc;
    ep.y = g_out;
    ep.e_k = e_k;
    ep.g_final = g_final;
    ep.upstream = upstream;
    ep.scale = inv_layers;
    ep.acc_in = acc_in;
    return launch<EPI_BWD>(a, nullptr, ep, dim, 0, stream);
}

namespace tagrec {
// y[col[j]] += val[j] * x[r] for every stored entry j of the listed rows r (a sub-warp of dim/4 lanes per listed row; the
// row of x in registers, one red.global.add.v4.f32 per lane and entry).  keep[i] == 0 skips a listed row (duplicates).
template <int LPR>
__global__ void __launch_bounds__(256)
spmm_push_rows_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                      const int64_t* __restrict__ rows, const uint8_t* __restrict__ keep, int64_t n_rows,
                      const float4* __restrict__ x4, float4* __restrict__ y4) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / LPR;
    const int sl = (int)(t % LPR);
    if (i >= n_rows) return;
    if (keep && !__ldg(keep + i)) return;
    const int64_t r = __ldg(rows + i);
    const float4 xv = __ldg(x4 + r * LPR + sl);
    const int64_t b = __ldg(rowptr + r), e = __ldg(rowptr + r + 1);
    for (int64_t j = b; j < e; ++j) {
        const int64_t c = __ldg(col + j);
        const float v = __ldg(val + j);
        red_add4(y4 + c * LPR + sl, make_float4(v * xv.x, v * xv.y, v * xv.z, v * xv.w));
    }
}
}  // namespace tagrec

extern "C" int tagrec_spmm_push_rows(const int64_t* rowptr, const int32_t* col, const float* val, const int64_t* rows,
                                     const uint8_t* keep, int64_t n_rows, const float* x, float* y_acc, int dim,
                                     void* stream) {
    TAGREC_REQUIRE(rowptr && col && val && rows && x && y_acc, "null pointer");
    TAGREC_REQUIRE(dim == 32 || dim == 64 || dim == 128, "dim must be 32, 64 or 128");
    if (n_rows == 0) return TAGREC_OK;
    const int lpr = dim / 4;
    const unsigned grid = (unsigned)((n_rows * lpr + 255) / 256);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y_acc);
    if (lpr == 16) { TAGREC_LAUNCH((spmm_push_rows_kernel<16>), grid, 256, 0, stream, rowptr, col, val, rows, keep, n_rows, x4, y4); }
    else if (lpr == 8) { TAGREC_LAUNCH((spmm_push_rows_kernel<8>), grid, 256, 0, stream, rowptr, col, val, rows, keep, n_rows, x4, y4); }
    else { TAGREC_LAUNCH((spmm_push_rows_kernel<32>), grid, 256, 0, stream, rowptr, col, val, rows, keep, n_rows, x4, y4); }
    return TAGREC_OK;
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
