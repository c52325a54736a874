// K3b-TC — the dense pass of the per-user AUC (eval_auc.cu, A2) on the 5th-generation tensor cores.  sm_100a, dim 64.
//
// Replaces training/utils.py:37-45 + training/basic_test.py:52-53 like eval_auc.cu does; only the pass over ALL
// (user, item) pairs changes.  For every score s the AUC needs 2 f(s) = 2 #{positives > s} + #{positives == s}, an
// integer that only depends on WHERE s falls between the user's sorted positives.  So the score does not have to be
// exact, it has to be on the right side of every positive:
//   * scores come from 3xTF32 MMAs (u_hi.i_hi + u_hi.i_lo + u_lo.i_hi, fp32 accumulation in TMEM; hi = the fp32 value
//     with its low 13 mantissa bits cleared, lo = x - hi, both exact), whose distance from the canonical fp32 dot
//     product is bounded by AT_MARGIN * ||u|| * max_i ||i||  (dropped lo.lo term 2^-20, TF32 truncation of the lo
//     operands 2 * 2^-20, accumulation order ~ 1e-5);
//   * a score whose bisection interval [P[lb-1], P[lb]) keeps that margin on both sides has the same 2 f as the exact
//     score; the others (about 2 m * margin * density of scores: a few per thousand) are re-scored exactly in the
//     canonical sequential-fmaf order — the SAME order the positives themselves and the subtraction passes A1/A3 use,
//     so the sums stay integers that cancel bit for bit, and the result equals the fp32 path's (tested).
//
// CTA = 128 users x one item split; 640 threads:
//   warp 0        TMA: raw fp32 item tile [128 x 64] into the stage (two 128B-swizzled boxes)
//   warps 18-19   converter: hi in place, lo into the stage's second half (generic proxy -> fence.proxy.async)
//   warp 1        MMA: 24 tcgen05.mma.kind::tf32 per tile (A = user rows hi | lo in TMEM, 128 columns)
//   warps 2-17    epilogue: thread = (user row, 32-column quarter of the accumulator): two register compares against
//                 [min - margin, max + margin] of the row's positives, else a branch-free bisection over the sorted
//                 positives in shared memory (k-major: conflict-free) + the margin test; uncertain columns are
//                 re-scored after the accumulator has been handed back.
// Rows are VIRTUAL rows: a user with more than 32 positives is split into several rows (same embedding, consecutive
// groups of 32 sorted positives) — 2 f is additive over such a partition (auc_vrows_kernel).
#include <float.h>

#include <algorithm>
#include <type_traits>

#include "tc_ptx.cuh"

namespace tagrec {

constexpr int AT_STAGES = 2;
constexpr int AT_NACC = 3;                  // 128 (A) + 3 * 128 TMEM columns
constexpr int AT_PC = 32;                   // positives per row staged in shared memory
// |3xTF32 score - canonical fp32 score| <= AT_MARGIN * ||u|| * max||i||.  Worst-case terms, all relative to ||u|| ||i||:
// dropped lo.lo products 2^-20 = 0.95e-6; TF32 truncation of the two lo operands 2 * 2^-20 = 1.9e-6; fp32 accumulation
// inside the tensor core over 24 instructions ~ 1.4e-6; rounding of the canonical sequential dot itself 64 * 2^-24 =
// 3.8e-6: sum 8e-6.  Measured maximum over 6e8 scores (tools/probes/eval_tc_experiments.patch, AT_EXPERIMENT 4): 0.93e-6.
constexpr float AT_MARGIN = 1.0e-5f;
constexpr int AT_THREADS = 640;             // 20 warps: TMA, MMA, 16 epilogue (4 per scheduler: the searches are latency-bound), 2 converter
constexpr int AT_STAGE_BYTES = 2 * TC_TILE_BYTES;        // hi tile + lo tile
constexpr int AT_Q = 256;                                // a warp's queue of in-range scores (of 32 rows x 32 columns)

struct AucTcArgs {
    const int64_t* users;
    int64_t nu;
    const float* user_table;
    const float* item_table;
    int64_t n_item;
    const int64_t* test_ptr;
    const float* pos_sorted;
    const int32_t* n_pos;
    const int32_t* vr_owner;                // [nv] index into users[] of the row this virtual row belongs to
    const int32_t* vr_part;                 // [nv] which group of AT_PC sorted positives of that row it searches
    const int32_t* nv;                      // device scalar: number of virtual rows
    const float* item_maxnorm;
    int64_t items_per_split;                // multiple of TC_N
    unsigned long long* acc2;
};

// Virtual rows: 2 f(s) is additive over a partition of the positives, so a user with m positives becomes
// max(1, ceil(m / AT_PC)) rows with the same embedding and consecutive groups of AT_PC sorted positives — every row's
// positives fit the shared-memory search, whatever the size of the test set.  One block; exclusive scan of the counts.
__global__ void __launch_bounds__(1024) auc_vrows_kernel(const int32_t* __restrict__ n_pos, int64_t nu,
                                                         int32_t* __restrict__ vr_owner, int32_t* __restrict__ vr_part,
                                                         int32_t* __restrict__ nv) {
    __shared__ int part_sum[1024];
    const int t = threadIdx.x;
    const int64_t per = (nu + 1023) / 1024, lo = t * per, hi = min(nu, lo + per);
    int mine = 0;
    for (int64_t w = lo; w < hi; ++w) mine += max(1, (n_pos[w] + AT_PC - 1) / AT_PC);
    part_sum[t] = mine;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {            // inclusive scan
        const int v = t >= o ? part_sum[t - o] : 0;
        __syncthreads();
        part_sum[t] += v;
        __syncthreads();
    }
    int pos = part_sum[t] - mine;
    for (int64_t w = lo; w < hi; ++w) {
        const int parts = max(1, (n_pos[w] + AT_PC - 1) / AT_PC);
        for (int k = 0; k < parts; ++k) {
            vr_owner[pos] = (int32_t)w;
            vr_part[pos] = k;
            ++pos;
        }
    }
    if (t == 1023) *nv = part_sum[1023];
}

__device__ __forceinline__ float at_dot_seq(const float* __restrict__ urow, const float* __restrict__ irow) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < TC_D / 4; ++c) {
        const float4 x = *reinterpret_cast<const float4*>(urow + 4 * c);
        const float4 y = __ldg(reinterpret_cast<const float4*>(irow) + c);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
        acc = fmaf(x.z, y.z, acc);
        acc = fmaf(x.w, y.w, acc);
    }
    return acc;
}

// exact 2 f(s) over ascending p[0], p[STRIDE], ..., m entries
template <int STRIDE>
__device__ __forceinline__ unsigned long long at_twice_f(const float* __restrict__ p, int m, float s) {
    int lo = 0, hi = m;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (p[mid * STRIDE] < s) lo = mid + 1; else hi = mid;
    }
    int eq = 0;
    while (lo + eq < m && p[(lo + eq) * STRIDE] == s) ++eq;
    return 2ull * (unsigned long long)(m - lo - eq) + (unsigned long long)eq;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
auc_tc_kernel(const __grid_constant__ CUtensorMap item_map, AucTcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS, not generic LD)
    unsigned char* St = base;                                              // [S] x (hi 32 KB | lo 32 KB)
    float* Uf = reinterpret_cast<float*>(St + AT_STAGES * AT_STAGE_BYTES); // [128][TC_UPITCH] user rows, exact fp32
    float* Ps = Uf + TC_M * TC_UPITCH;                                     // [1 + AT_PC + 1][128]: -inf row, the sorted
                                                                           // positives (+inf padded), +inf row; k-major:
                                                                           // bank = row % 32 = lane, whatever k
    uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + (AT_PC + 2) * TC_M);
    uint64_t* full = bars;                        // [S]    TMA -> converter
    uint64_t* conv = full + AT_STAGES;            // [S]    converter -> MMA
    uint64_t* sfree = conv + AT_STAGES;           // [S]    MMA (commit) -> TMA
    uint64_t* accfull = sfree + AT_STAGES;        // [NACC] MMA -> epilogue
    uint64_t* accfree = accfull + AT_NACC;        // [NACC] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + AT_NACC);
    float* Qs = reinterpret_cast<float*>(tmem_slot + 2);              // [16 warps][AT_Q] queued scores
    uint32_t* Qc = reinterpret_cast<uint32_t*>(Qs + 16 * AT_Q);       // [16][AT_Q]  (row in warp) << 5 | column
    uint32_t* Racc = Qc + 16 * AT_Q;                                  // [16][32]    per-row sums of (m - lb)
    uint32_t* Rum = Racc + 16 * 32;                                   // [16][32]    per-row masks of uncertain columns

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t u0 = (int64_t)blockIdx.x * TC_M;
    const int64_t i_begin = (int64_t)blockIdx.y * a.items_per_split;
    const int64_t i_end = min(a.n_item, i_begin + a.items_per_split);
    const int n_tiles = (int)((i_end - i_begin + TC_N - 1) / TC_N);
    const int64_t nv = __ldg(a.nv);
    if (u0 >= nv) return;                       // the grid is sized for the upper bound of the virtual-row count

    if (tid == 0) {
        for (int s = 0; s < AT_STAGES; ++s) {
            mbar_init(smem_u32(full + s), 1);
            mbar_init(smem_u32(conv + s), 2);
            mbar_init(smem_u32(sfree + s), 1);
        }
        for (int x = 0; x < AT_NACC; ++x) {
            mbar_init(smem_u32(accfull + x), 1);
            mbar_init(smem_u32(accfree + x), 16);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_base = tmem_base + 128u;

    // epilogue identity: warps 2..17; TMEM lane quarter = warp % 4; four warps (column quarters) per lane quarter
    const bool is_epi = warp >= 2 && warp < 18;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;           // column quarter 0..3 (32 columns each)
    const int row = q * 32 + lane;
    const bool valid = is_epi && (u0 + row < nv);
    int pm = 0;
    int64_t owner = 0;                          // index into users[] / acc2[]
    float unorm2 = 0.f;
    const float* pp = a.pos_sorted;
    if (is_epi) {
        const float4* urow = nullptr;
        if (valid) {
            owner = __ldg(a.vr_owner + u0 + row);
            const int part = __ldg(a.vr_part + u0 + row);
            const int64_t u = __ldg(a.users + owner);
            urow = reinterpret_cast<const float4*>(a.user_table + u * TC_D);
            pm = min(AT_PC, max(0, __ldg(a.n_pos + owner) - part * AT_PC));
            pp = a.pos_sorted + __ldg(a.test_ptr + u) + part * AT_PC;
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            uint32_t rh[32], rl[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) v = __ldg(urow + hf * 8 + c);
                unorm2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, unorm2))));
                const float xs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t hb = __float_as_uint(xs[e]) & 0xFFFFE000u;
                    rh[4 * c + e] = hb;
                    rl[4 * c + e] = __float_as_uint(xs[e] - __uint_as_float(hb));
                }
                if (half == 0) *reinterpret_cast<float4*>(Uf + (size_t)row * TC_UPITCH + 4 * (hf * 8 + c)) = v;
            }
            if (half == 0) {
                tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 32), rh);
                tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 + hf * 32), rl);
            }
        }
        if (half == 0) {
            tmem_st_wait();
            Ps[row] = -INFINITY;
            for (int j = 0; j <= AT_PC; ++j) Ps[(j + 1) * TC_M + row] = j < pm ? __ldg(pp + j) : INFINITY;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % AT_STAGES;
                mbar_wait_relaxed(smem_u32(sfree + s), ((t / AT_STAGES) & 1) ^ 1);
                const uint32_t bar = smem_u32(full + s);
                mbar_expect_tx(bar, TC_TILE_BYTES);
                const uint32_t dst = smem_u32(St + s * AT_STAGE_BYTES);
                const int row0 = (int)(i_begin + (int64_t)t * TC_N);
                tma_load_2d(dst, &item_map, bar, 0, row0);
                tma_load_2d(dst + TC_KH_BYTES, &item_map, bar, 32, row0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % AT_STAGES, r = t % AT_NACC;
                mbar_wait(smem_u32(conv + s), (t / AT_STAGES) & 1);
                mbar_wait(smem_u32(accfree + r), ((t / AT_NACC) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = acc_base + (uint32_t)(r * TC_N);
                const uint32_t bh = smem_u32(St + s * AT_STAGE_BYTES), bl = bh + TC_TILE_BYTES;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t off = (uint32_t)((kk >> 2) * TC_KH_BYTES + (kk & 3) * 32);
                    const uint32_t ah = tmem_base + (uint32_t)(kk * 8), al = ah + 64u;
                    umma_tf32_ts(d, al, umma_desc_sw128(bh + off), kk > 0);      // small terms first
                    umma_tf32_ts(d, ah, umma_desc_sw128(bl + off), 1);
                    umma_tf32_ts(d, ah, umma_desc_sw128(bh + off), 1);
                }
                umma_commit(smem_u32(sfree + s));
                umma_commit(smem_u32(accfull + r));
            }
        }
    } else if (warp >= 18) {
        // ================= converter: raw -> hi (in place) | lo =================
        const int ct = tid - 576;
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t % AT_STAGES;
            mbar_wait_relaxed(smem_u32(full + s), (t / AT_STAGES) & 1);
            unsigned char* hi = St + s * AT_STAGE_BYTES;
            unsigned char* lo = hi + TC_TILE_BYTES;
#pragma unroll 4
            for (int i = 0; i < TC_TILE_BYTES / 16 / 64; ++i) {
                const int off = (ct + 64 * i) * 16;
                const float4 x = *reinterpret_cast<const float4*>(hi + off);
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                l.x = x.x - h.x;
                l.y = x.y - h.y;
                l.z = x.z - h.z;
                l.w = x.w - h.w;
                *reinterpret_cast<float4*>(hi + off) = h;
                *reinterpret_cast<float4*>(lo + off) = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(conv + s));
        }
    } else {
        // ================= epilogue: thread = (user row, 64-column half) =================
        const float mg = valid ? AT_MARGIN * sqrtf(unorm2) * __ldg(a.item_maxnorm) + FLT_MIN : 0.f;
        float below_thr = INFINITY, above_thr = -INFINITY;      // pm == 0: never asks for a search, adds nothing
        if (pm > 0) {
            below_thr = __ldg(pp) - mg;
            above_thr = __ldg(pp + pm - 1) + mg;
        }
        const float* ps = Ps + TC_M + row;                       // positive k of this row: ps[k * TC_M], k = -1 .. AT_PC
        const float piv7 = ps[7 * TC_M], piv15 = ps[15 * TC_M], piv23 = ps[23 * TC_M];
        const float* urow_s = Uf + (size_t)row * TC_UPITCH;
        unsigned long long tot = 0;
        bool dense_mode = false;
        for (int t = 0; t < n_tiles; ++t) {
            const int r = t % AT_NACC;
            const int64_t it0 = i_begin + (int64_t)t * TC_N + half * 32;
            const int nvalid = (int)min((int64_t)32, i_end - it0);          // <= 0: nothing of this quarter is a real item
            mbar_wait_relaxed(smem_u32(accfull + r), (t / AT_NACC) & 1);
            tc_fence_after();
            const uint32_t taddr = acc_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * TC_N + half * 32);
            uint32_t v0[32];
            tmem_ld32(taddr, v0);
            tmem_ld_wait();
            // Branch-free on purpose: where a score falls differs from lane to lane, and 128 divergent branches per
            // tile (the first version) cost 3/4 of the kernel.  One warp-uniform test (is any score of the warp not
            // below every positive of its row?) skips the search for the easy tiles of a trained model; otherwise every
            // column is searched: "below all" and "above all" are just the results lb = 0 and lb = m.
            struct Scan {
                unsigned int certain;                    // sum of (m - lb) over the columns decided here
                unsigned int um;                         // columns that need the exact score
            };
            auto scan = [=](const uint32_t (&v)[32], int c0) -> Scan {
                Scan r{0u, 0u};
                const int nreal = min(32, max(0, nvalid - c0));  // zero-filled rows past the table end are not items
                float mx = __uint_as_float(v[0]);
#pragma unroll
                for (int j = 1; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
                if (!__any_sync(0xffffffffu, mx >= below_thr)) {
                    r.certain = (unsigned int)(nreal * pm);
                    return r;
                }
                // 8 columns at a time, step-major: 8 independent probe chains per thread (shared-memory latency is
                // high while the MMA and the converter stream through the same banks); the first two levels of
                // the search compare against pivots held in registers.
                unsigned int off_sum = 0, n_sure = 0;    // over the decided columns: sum of lb * TC_M, count
#pragma unroll
                for (int j0 = 0; j0 < 32; j0 += 8) {
                    const float* cur[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float s = __uint_as_float(v[j0 + g]);
                        const bool l1 = piv15 < s;
                        const bool l2 = (l1 ? piv23 : piv7) < s;
                        cur[g] = ps + (l1 ? 16 * TC_M : 0) + (l2 ? 8 * TC_M : 0);
                    }
#pragma unroll
                    for (int st = 4; st > 0; st >>= 1)
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            cur[g] += (cur[g][(st - 1) * TC_M] < __uint_as_float(v[j0 + g])) ? st * TC_M : 0;
#pragma unroll
                    for (int g = 0; g < 8; ++g) cur[g] += (cur[g][0] < __uint_as_float(v[j0 + g])) ? TC_M : 0;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int j = j0 + g;
                        const float s = __uint_as_float(v[j]);
                        const float lo_p = cur[g][-TC_M], hi_p = cur[g][0];          // sentinel rows at -1 and AT_PC
                        const bool real = j < nreal;
                        const bool sure = real && s - lo_p > mg && hi_p - s > mg;
                        const bool redo = real && !sure;
                        off_sum += sure ? (unsigned int)(cur[g] - ps) : 0u;
                        n_sure += sure ? 1u : 0u;
                        r.um |= redo ? (1u << j) : 0u;
                    }
                }
                r.certain += n_sure * (unsigned int)pm - off_sum / TC_M;
                return r;
            };
            // Which columns are in range at all?  For a trained model (AUC 0.9+) it is a few per cent, spread over
            // all rows — so the dense search above (every lane, all its 32 columns) would run for nearly every tile
            // with nearly every result "below all".  Sparse case: the warp's in-range (row, column) pairs go to a
            // queue in shared memory and are searched 32 at a time, any lane serving any row.
            // (while the dense regime lasts the classification is skipped; it is re-probed every 16th tile)
            unsigned int inr = 0;
            int n_below = 0, cnt = 0, total = AT_Q + 1;
            if (!dense_mode || (t & 15) == 0) {              // warp-uniform
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float s = __uint_as_float(v0[j]);
                    const bool real = j < nvalid;
                    const bool below = real && s < below_thr;
                    n_below += below ? 1 : 0;
                    inr |= (real && !below && s <= above_thr) ? (1u << j) : 0u;
                }
                cnt = __popc(inr);
                total = __reduce_add_sync(0xffffffffu, cnt);
                dense_mode = total > AT_Q;
            }
            Scan r0{(unsigned int)(n_below * pm), 0u};
            if (total > AT_Q) {
                r0 = scan(v0, 0);
            } else if (total > 0) {
                const int ew = warp - 2;
                float* qs = Qs + ew * AT_Q;
                uint32_t* qc = Qc + ew * AT_Q;
                uint32_t* racc = Racc + ew * 32;
                uint32_t* rum = Rum + ew * 32;
                int pos = cnt;                               // exclusive prefix sum over the lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t2 = __shfl_up_sync(0xffffffffu, pos, o);
                    pos += lane >= o ? t2 : 0;
                }
                pos -= cnt;
                racc[lane] = 0u;
                rum[lane] = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if ((inr >> j) & 1u) {
                        qs[pos] = __uint_as_float(v0[j]);
                        qc[pos] = ((uint32_t)lane << 5) | (uint32_t)j;
                        ++pos;
                    }
                }
                __syncwarp();
                for (int b0 = 0; b0 < total; b0 += 32) {
                    const bool act = b0 + lane < total;
                    const float s = act ? qs[b0 + lane] : 0.f;
                    const uint32_t code = act ? qc[b0 + lane] : ((uint32_t)lane << 5);
                    const int rl = (int)(code >> 5), col = (int)(code & 31u);
                    const int pm_r = __shfl_sync(0xffffffffu, pm, rl);
                    const float mg_r = __shfl_sync(0xffffffffu, mg, rl);
                    const float* pr = Ps + TC_M + (q * 32 + rl);
                    const float* cur = pr;
#pragma unroll
                    for (int st = AT_PC / 2; st > 0; st >>= 1) cur += (cur[(st - 1) * TC_M] < s) ? st * TC_M : 0;
                    cur += (cur[0] < s) ? TC_M : 0;
                    const float lo_p = cur[-TC_M], hi_p = cur[0];
                    const bool sure = s - lo_p > mg_r && hi_p - s > mg_r;
                    if (act) {
                        if (sure) atomicAdd(racc + rl, (uint32_t)(pm_r - (int)(cur - pr) / TC_M));
                        else atomicOr(rum + rl, 1u << col);
                    }
                }
                __syncwarp();
                r0.certain += racc[lane];
                r0.um = rum[lane];
            }
            unsigned int um = r0.um;
            tot += 2ull * (unsigned long long)r0.certain;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(accfree + r));
            while (um) {
                const int j = __ffs((int)um) - 1;
                um &= um - 1;
                const float ex = at_dot_seq(urow_s, a.item_table + (it0 + j) * TC_D);
                tot += at_twice_f<TC_M>(ps, pm, ex);
            }
        }
        if (valid && tot) atomicAdd(a.acc2 + owner, tot);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

static size_t auc_tc_smem() {
    return 1024 + (size_t)AT_STAGES * AT_STAGE_BYTES + ((size_t)TC_M * TC_UPITCH + TC_M * (AT_PC + 2) + 2) * 4 +
           (3 * AT_STAGES + 2 * AT_NACC) * 8 + 64 + (size_t)16 * (2 * AT_Q + 64) * 4;
}

bool auc_tc_available(int dim) { return dim == TC_D && encode_tiled() != nullptr; }

size_t auc_tc_workspace_bytes(int64_t nu, int64_t n_test_total) {
    const int64_t nv_max = nu + n_test_total / AT_PC + 1;
    return 64 + (size_t)nv_max * 8;
}

// Builds the virtual-row tables in ws (layout: [0] max item norm (K3b-TC), [4] row count, [16..] owner, part).
int launch_auc_vrows(const int32_t* n_pos, int64_t nu, int64_t n_test_total, void* ws, const int32_t** vr_owner,
                     const int32_t** vr_part, const int32_t** nv, int64_t* nv_max, void* stream) {
    *nv_max = nu + n_test_total / AT_PC + 1;
    int32_t* base = reinterpret_cast<int32_t*>(ws);
    TAGREC_LAUNCH(auc_vrows_kernel, 1, 1024, 0, stream, n_pos, nu, base + 16, base + 16 + *nv_max, base + 4);
    *nv = base + 4;
    *vr_owner = base + 16;
    *vr_part = base + 16 + *nv_max;
    return TAGREC_OK;
}

// ws: auc_tc_workspace_bytes(nu, n_test_total) device bytes (max item norm, virtual-row count, the two row tables)
int eval_auc_tc(const int64_t* users, int64_t nu, const float* user_table, const float* item_table, int64_t n_item,
                const int64_t* test_ptr, const float* pos_sorted, const int32_t* n_pos, int64_t n_test_total, void* ws,
                unsigned long long* acc2, void* stream) {
    TAGREC_REQUIRE((reinterpret_cast<uintptr_t>(item_table) & 15) == 0, "item table must be 16-byte aligned");
    CUtensorMap map;
    if (int rc = make_row_table_map(&map, item_table, n_item, TC_D)) return rc;
    float* maxnorm = reinterpret_cast<float*>(ws);
    const int32_t *vr_owner, *vr_part, *nv;
    int64_t nv_max = 0;
    if (int rc = launch_auc_vrows(n_pos, nu, n_test_total, ws, &vr_owner, &vr_part, &nv, &nv_max, stream)) return rc;
    if (int rc = launch_item_maxnorm(item_table, n_item, TC_D, maxnorm, stream)) return rc;
    AucTcArgs a{};
    a.users = users; a.nu = nu; a.user_table = user_table; a.item_table = item_table; a.n_item = n_item;
    a.test_ptr = test_ptr; a.pos_sorted = pos_sorted; a.n_pos = n_pos; a.item_maxnorm = maxnorm; a.acc2 = acc2;
    a.vr_owner = vr_owner; a.vr_part = vr_part; a.nv = nv;
    const int64_t grid_tiles = (nv_max + TC_M - 1) / TC_M;       // CTAs past the actual count return at once
    const int64_t user_tiles = (nu + TC_M - 1) / TC_M;           // the usual case: about one virtual row per user
    const int64_t item_tiles = (n_item + TC_N - 1) / TC_N;
    int64_t splits = user_tiles >= kSMs ? 1 : kSMs / user_tiles;
    splits = std::max<int64_t>(1, std::min<int64_t>(splits, (item_tiles + 3) / 4));
    a.items_per_split = ((item_tiles + splits - 1) / splits) * TC_N;
    splits = (n_item + a.items_per_split - 1) / a.items_per_split;
    const size_t smem = auc_tc_smem();
    TAGREC_CUDA(cudaFuncSetAttribute(auc_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TAGREC_LAUNCH(auc_tc_kernel, dim3((unsigned)grid_tiles, (unsigned)splits), AT_THREADS, smem, stream, map, a);
    return TAGREC_OK;
}

}  // namespace tagrec
