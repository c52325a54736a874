// K3-TC — full-sort evaluation on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a only.
//
// Replaces model/lightgcn.py:84-89 (predict_rating: sigmoid(U_b @ I^T), a materialised B x n_item matrix),
// training/basic_test.py:42-48 (mask train items with -1024, torch.topk) for dim_latent == 64 tables.
//
// The only dense contraction of the path: scores[u, i] = <U[u, :], I[i, :]>, K = 64.  One CTA owns 128*NH users
// (NH "halves" of 128 rows) and a contiguous range of items ("split"); it streams 128-item tiles:
//
//   warp 0      TMA producer   cp.async.bulk.tensor.2d of the item tile [128 x 64] fp32 (two 128B-swizzled boxes
//                              of 32 floats) into a ring of B stages, mbarrier complete_tx
//   warp 1      MMA issuer     one lane: 8*NH tcgen05.mma.kind::tf32 (M128 x N128 x K8) per tile, A = user tile
//                              resident in smem (K-major, SW128), B = item tile, D = fp32 accumulator in TMEM
//                              (2 accumulator stages x NH halves x 128 columns); tcgen05.commit -> mbarrier
//   warps 2..   epilogue       thread = user row: tcgen05.ld 32 columns at a time, max-tree, compare with the row's
//                              running threshold (a register).  Scores never leave the SM.
//
// Exactness.  TF32 inputs carry 10 mantissa bits, so the tensor-core score is only a FILTER: a score is a candidate
// when  s_tf32 > thr_exact - margin,  margin = 2.2e-3 * ||u||_2 * max_i ||i||_2  >=  the worst-case TF32 error
// (2 * 2^-10 * sum_k |u_k i_k|  <=  2^-9 ||u|| ||i||, plus accumulation slack).  Every candidate is re-scored in
// exact fp32 (sequential fmaf over k = 0..63 — the same canonical order as the CUDA-core path in eval_topk.cu)
// straight from the smem copies of the two rows, and only exact scores enter the row's sorted K-list and its
// threshold.  The result is therefore identical to the fp32 path: top-K by (-score, item id).
//
// Train-item masking costs no memory traffic in the common case: a per-row cursor over the user's ascending train
// row keeps the next masked item id in a register (items are visited in ascending id order).  Masked items are never
// candidates; if a user has fewer than K un-masked items the merge kernel appends the first masked ids with score
// -1024, which is exactly the (-score, id) order the reference's -1024 fill produces (basic_test.py:47).
#include <algorithm>

#include "eval_tc.cuh"
#include "tc_ptx.cuh"

namespace tagrec {




// ---------------------------------------------------------------------------------------------- the kernel
// TC_TS == 1 (default): the A operand (user rows) lives in TENSOR MEMORY (tcgen05.mma "TS" form): every epilogue
// thread stores its own 64-float user row into its TMEM lane once (tcgen05.st).  An SS-form M128 x N128 x K8 TF32
// MMA reads 4 KB of A + 4 KB of B from shared memory per 64 cycles = 128 B/clk, i.e. ALL of an SM's shared-memory
// bandwidth, leaving nothing for the TMA writes of the next item tile; with A in TMEM the MMA reads 64 B/clk and
// the 64 KB of smem the user tile occupied become two more item-tile stages.
// TC_TS == 0: A in shared memory (SS form), kept for comparison.
#ifndef TC_TS
#define TC_TS 1
#endif
#ifndef TC_INTERLEAVE
#define TC_INTERLEAVE 0      // 1: issue the two halves' MMAs alternately (independent accumulators back to back)
#endif
constexpr int TC_NACC = TC_TS ? 3 : 4;   // accumulator ring (128 TMEM columns each; A takes 128 columns in TS form)


template <int NH>
__global__ void __launch_bounds__(64 + 128 * NH, 1)
eval_tc_kernel(const __grid_constant__ CUtensorMap item_map, TcArgs a) {
    constexpr int ROWS = TC_M * NH;
    constexpr bool kTS = TC_TS != 0;
    constexpr int A_COLS = kTS ? 64 * NH : 0;                 // TMEM columns holding the user rows (one fp32 per column)
    constexpr int TMEM_COLS = 512;                            // A_COLS + TC_NACC * 128 <= 512, power of two
    constexpr int A_SMEM = kTS ? 0 : NH * TC_TILE_BYTES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS, not generic LD)
    const int S = a.stages, K = a.k;
    unsigned char* As = base;                                         // SS form only: NH x 32 KB
    unsigned char* Bs = base + A_SMEM;                                // S x 32 KB
    float* Uf = reinterpret_cast<float*>(Bs + S * TC_TILE_BYTES);     // [ROWS][TC_UPITCH] user rows, plain fp32 (re-scores)
    float* ls = Uf + (size_t)ROWS * TC_UPITCH;                        // [K][ROWS] scores (unsorted K-lists)
    int32_t* li = reinterpret_cast<int32_t*>(ls + (size_t)K * ROWS);  // [K][ROWS] item ids
    uint64_t* bars = reinterpret_cast<uint64_t*>(li + (size_t)K * ROWS);
    uint64_t* full = bars;                                    // [S]     TMA -> MMA
    uint64_t* bfree = bars + TC_MAX_STAGES;                   // [S]     epilogue -> TMA
    uint64_t* accfull = bars + 2 * TC_MAX_STAGES;             // [NACC]  MMA -> epilogue
    uint64_t* accfree = accfull + TC_NACC;                    // [NACC]  epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + TC_NACC);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t u0 = (int64_t)blockIdx.x * ROWS;
    const int split = blockIdx.y;
    const int64_t i_begin = (int64_t)split * a.items_per_split;
    const int64_t i_end = min(a.n_item, i_begin + a.items_per_split);
    const int n_tiles = (int)((i_end - i_begin + TC_N - 1) / TC_N);

    // ---- prologue: barriers, TMEM allocation ----
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(full + s), 1);
            mbar_init(smem_u32(bfree + s), 4 * NH);           // every epilogue warp releases the stage
        }
        for (int x = 0; x < TC_NACC; ++x) {
            mbar_init(smem_u32(accfull + x), 1);
            mbar_init(smem_u32(accfree + x), 4);              // the 4 warps of ONE half drain an accumulator
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if constexpr (!kTS) {     // user tile -> smem (generic proxy), swizzled exactly like a TMA SWIZZLE_128B box
        for (int idx = tid; idx < ROWS * 16; idx += blockDim.x) {
            const int row = idx >> 4, c16 = idx & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (u0 + row < a.nu) {
                const int64_t u = __ldg(a.users + u0 + row);
                v = __ldg(reinterpret_cast<const float4*>(a.user_table + u * TC_D) + c16);
            }
            *reinterpret_cast<float4*>(As + (row >> 7) * TC_TILE_BYTES + sw128_off(row & 127, c16)) = v;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // smem writes -> visible to tcgen05.mma
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_base = tmem_base + (uint32_t)A_COLS;

    // epilogue-thread identity (computed by every thread; only warps >= 2 use it)
    const int e = warp - 2;                     // 0 .. 4*NH-1
    const int h = e >> 2;
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int row = h * TC_M + q * 32 + lane;
    const bool is_epi = warp >= 2;
    const bool valid = is_epi && (u0 + row < a.nu);
    const float4* urow = nullptr;               // this thread's user row in global memory (exact re-scores)
    float unorm2 = 0.f;
    if (is_epi) {
        if (valid) urow = reinterpret_cast<const float4*>(a.user_table + __ldg(a.users + u0 + row) * TC_D);
        // row -> registers (two halves of 32 floats) -> this thread's TMEM lane, columns h*64 .. h*64+63
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) v = __ldg(urow + half * 8 + c);
                unorm2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, unorm2))));
                r[4 * c + 0] = __float_as_uint(v.x);
                r[4 * c + 1] = __float_as_uint(v.y);
                r[4 * c + 2] = __float_as_uint(v.z);
                r[4 * c + 3] = __float_as_uint(v.w);
                *reinterpret_cast<float4*>(Uf + (size_t)row * TC_UPITCH + 4 * (half * 8 + c)) = v;
            }
            if constexpr (kTS) tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 64 + half * 32), r);
        }
        if constexpr (kTS) tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();                            // user rows are in TMEM before the first MMA reads them
    tc_fence_after();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % S;
                mbar_wait(smem_u32(bfree + s), ((t / S) & 1) ^ 1);
                const uint32_t bar = smem_u32(full + s);
                mbar_expect_tx(bar, TC_TILE_BYTES);
                const uint32_t dst = smem_u32(Bs + s * TC_TILE_BYTES);
                const int row0 = (int)(i_begin + (int64_t)t * TC_N);
                tma_load_2d(dst, &item_map, bar, 0, row0);
                tma_load_2d(dst + TC_KH_BYTES, &item_map, bar, 32, row0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % S;
                mbar_wait(smem_u32(full + s), (t / S) & 1);
                const uint32_t b0 = smem_u32(Bs + s * TC_TILE_BYTES);
#if TC_INTERLEAVE
                // Consecutive MMAs into the SAME accumulator form a dependent chain (each waits for the previous
                // accumulate to retire); alternating the two halves' accumulators keeps the tensor pipe busy.
                {
                    uint32_t dd[NH];
#pragma unroll
                    for (int hh = 0; hh < NH; ++hh) {
                        const int n = t * NH + hh, r = n % TC_NACC;
                        mbar_wait(smem_u32(accfree + r), ((n / TC_NACC) & 1) ^ 1);
                        dd[hh] = acc_base + (uint32_t)(r * TC_N);
                    }
                    tc_fence_after();
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t off = (uint32_t)((kk >> 2) * TC_KH_BYTES + (kk & 3) * 32);
                        const uint64_t bd = umma_desc_sw128(b0 + off);
#pragma unroll
                        for (int hh = 0; hh < NH; ++hh) {
                            if constexpr (kTS)
                                umma_tf32_ts(dd[hh], tmem_base + (uint32_t)(hh * 64 + kk * 8), bd, kk > 0);
                            else
                                umma_tf32(dd[hh], umma_desc_sw128(smem_u32(As + hh * TC_TILE_BYTES) + off), bd, kk > 0);
                        }
                    }
#pragma unroll
                    for (int hh = 0; hh < NH; ++hh) umma_commit(smem_u32(accfull + (t * NH + hh) % TC_NACC));
                    continue;
                }
#endif
#pragma unroll
                for (int hh = 0; hh < NH; ++hh) {
                    const int n = t * NH + hh, r = n % TC_NACC;
                    mbar_wait(smem_u32(accfree + r), ((n / TC_NACC) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d = acc_base + (uint32_t)(r * TC_N);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {     // K = 8 per instruction: 4 per 128-byte swizzle row, 2 k-halves
                        const uint32_t off = (uint32_t)((kk >> 2) * TC_KH_BYTES + (kk & 3) * 32);
                        if constexpr (kTS)
                            umma_tf32_ts(d, tmem_base + (uint32_t)(hh * 64 + kk * 8), umma_desc_sw128(b0 + off), kk > 0);
                        else
                            umma_tf32(d, umma_desc_sw128(smem_u32(As + hh * TC_TILE_BYTES) + off),
                                      umma_desc_sw128(b0 + off), kk > 0);
                    }
                    umma_commit(smem_u32(accfull + r));  // arrives when every MMA issued so far has completed
                }
            }
        }
    } else {
        // ================= epilogue: thread = user row =================
        float thr = -INFINITY, thr_lo = valid ? -INFINITY : INFINITY, margin = 0.f;
        float thr_sh = -INFINITY;               // bound published by the other item splits of this user
        float* sh_slot = (valid && a.shared_thr) ? a.shared_thr + (u0 + row) : nullptr;
        int cnt = 0, minpos = 0;                // the K-list is UNSORTED; minpos = entry to evict ((score, -id) minimum)
        int64_t tc = 0, te = 0;
        int32_t nxt = INT32_MAX;                // smallest train item of this user not yet passed
        if (valid) {
            const int64_t u = __ldg(a.users + u0 + row);
            margin = TC_MARGIN * sqrtf(unorm2) * __ldg(a.item_maxnorm) + FLT_MIN;
            tc = __ldg(a.train_ptr + u);
            te = __ldg(a.train_ptr + u + 1);
            int64_t lo = tc, hi = te;           // first train item inside this split
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if ((int64_t)__ldg(a.train_items + mid) < i_begin) lo = mid + 1; else hi = mid;
            }
            tc = lo;
            nxt = tc < te ? __ldg(a.train_items + tc) : INT32_MAX;
        }
        for (int t = 0; t < n_tiles; ++t) {
            const int s = t % S;
            const int n = t * NH + h, r = n % TC_NACC;
            const int64_t it0 = i_begin + (int64_t)t * TC_N;
            const unsigned char* brow = Bs + s * TC_TILE_BYTES;
            // another split may have raised the bound (a stale read only prunes less); issued before the wait
            const float sh_new = sh_slot ? __ldcg(sh_slot) : -INFINITY;
            mbar_wait(smem_u32(accfull + r), (n / TC_NACC) & 1);
            tc_fence_after();
            const uint32_t taddr = acc_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * TC_N);
            if (sh_new > thr_sh) {
                thr_sh = sh_new;
                thr_lo = fmaxf(thr, thr_sh) - margin;
            }
            // ---- fast path: 128 TF32 scores of this row -> a 128-bit candidate mask (usually empty) ----
            uint32_t cm[TC_N / 32];
            auto scan32 = [&](const uint32_t (&v)[32]) -> uint32_t {
                float m = __uint_as_float(v[0]);
#pragma unroll
                for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                uint32_t mask = 0;
                if (m > thr_lo) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) > thr_lo) ? (1u << j) : 0u;
                }
                return mask;
            };
#pragma unroll
            for (int c = 0; c < TC_N / 32; c += 2) {      // two 32-column loads in flight per wait
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c * 32, v0);
                tmem_ld32(taddr + (c + 1) * 32, v1);
                tmem_ld_wait();
                cm[c] = scan32(v0);
                cm[c + 1] = scan32(v1);
            }
            // the accumulator is drained: hand it back to the MMA warp before the (rare, slow) candidate work
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(accfree + r));
            // ---- slow path: all lanes' candidates of this tile in lockstep ----
            uint64_t lo64 = (uint64_t)cm[0] | ((uint64_t)cm[1] << 32), hi64 = (uint64_t)cm[2] | ((uint64_t)cm[3] << 32);
            while (lo64 | hi64) {
                int il;
                if (lo64) {
                    il = __ffsll((long long)lo64) - 1;
                    lo64 &= lo64 - 1;
                } else {
                    il = 64 + __ffsll((long long)hi64) - 1;
                    hi64 &= hi64 - 1;
                }
                const int64_t item = it0 + il;
                if (item >= i_end) break;                     // zero-filled rows past the split / table end
                // train-item cursor: nxt = smallest train item >= item (galloping, then bisection)
                if ((int64_t)nxt < item) {
                    int64_t step = 1, lo = tc + 1;
                    while (lo + step < te && (int64_t)__ldg(a.train_items + lo + step) < item) {
                        lo += step;
                        step <<= 1;
                    }
                    int64_t hi = min(te, lo + step + 1);
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if ((int64_t)__ldg(a.train_items + mid) < item) lo = mid + 1; else hi = mid;
                    }
                    tc = lo;
                    nxt = tc < te ? __ldg(a.train_items + tc) : INT32_MAX;
                }
                if ((int64_t)nxt == item) continue;           // masked (basic_test.py:47)
                // exact fp32 score, canonical sequential order; item row from the smem stage, user row from its smem copy
                float ex = 0.f;
#pragma unroll
                for (int c16 = 0; c16 < 16; ++c16) {
                    const float4 uu = *reinterpret_cast<const float4*>(Uf + (size_t)row * TC_UPITCH + 4 * c16);
                    const float4 ii = *reinterpret_cast<const float4*>(brow + sw128_off(il, c16));
                    ex = fmaf(uu.x, ii.x, ex);
                    ex = fmaf(uu.y, ii.y, ex);
                    ex = fmaf(uu.z, ii.z, ex);
                    ex = fmaf(uu.w, ii.w, ex);
                }
                // items arrive in ascending id order, so on a score tie the incumbent (smaller id) stays: strict >.
                // Against the bound of ANOTHER split only strictly smaller scores may be dropped (its K-th item
                // may have a larger id than this one).
                if ((cnt < K || ex > thr) && ex >= thr_sh) {
                    const int pos = cnt < K ? cnt : minpos;
                    ls[(size_t)pos * ROWS + row] = ex;
                    li[(size_t)pos * ROWS + row] = (int32_t)item;
                    if (cnt < K) ++cnt;
                    if (cnt == K) {     // new evictee: lowest score, largest id among equals
                        float best = INFINITY;
                        int32_t besti = -1;
                        int bp = 0;
#pragma unroll 4
                        for (int j = 0; j < K; ++j) {
                            const float sj = ls[(size_t)j * ROWS + row];
                            const int32_t ij = li[(size_t)j * ROWS + row];
                            if (sj < best || (sj == best && ij > besti)) {
                                best = sj;
                                besti = ij;
                                bp = j;
                            }
                        }
                        thr = best;
                        minpos = bp;
                        thr_lo = fmaxf(thr, thr_sh) - margin;
                        if (sh_slot && thr > thr_sh) atomic_max_float(sh_slot, thr);
                    }
                }
            }
            // this warp no longer reads B stage `s`
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(bfree + s));
        }
        if (valid) {
            const size_t o = ((size_t)(u0 + row) * a.splits + split) * K;
            for (int j = 0; j < K; ++j) {
                a.part_scores[o + j] = j < cnt ? ls[(size_t)j * ROWS + row] : -INFINITY;
                a.part_ids[o + j] = j < cnt ? li[(size_t)j * ROWS + row] : -1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- wide tables
// dim = 64 * KB, KB = 2..4 (NGCF's 256-d concat, ngcf.py:89; TGCN's 192-d, tgcn.py:227-229; 128-d tables).  Same roles
// as eval_tc_kernel with one 128-user half per CTA; a tile is KB item k-blocks of [128 x 64] that stream through the
// stage ring one after the other (stages are released by tcgen05.commit as soon as the MMAs that read them retire),
// the user rows occupy 64*KB TMEM columns, the accumulator ring has two 128-column slots, and the exact re-score
// reads both rows from global memory (L2: the item rows were streamed through it a moment ago).
__global__ void __launch_bounds__(64 + 128, 1)
eval_tc_wide_kernel(const __grid_constant__ CUtensorMap item_map, TcArgs a, int KB) {
    constexpr int ROWS = TC_M;
    constexpr int NACC = 2;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (LDS, not generic LD)
    const int S = a.stages, K = a.k, D = KB * TC_D;
    unsigned char* Bs = base;                                         // S x 32 KB
    float* ls = reinterpret_cast<float*>(Bs + S * TC_TILE_BYTES);     // [K][ROWS]
    int32_t* li = reinterpret_cast<int32_t*>(ls + (size_t)K * ROWS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(li + (size_t)K * ROWS);
    uint64_t* full = bars;                            // [S]
    uint64_t* bfree = bars + TC_MAX_STAGES;           // [S]   released by tcgen05.commit
    uint64_t* accfull = bars + 2 * TC_MAX_STAGES;     // [2]
    uint64_t* accfree = accfull + NACC;               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + NACC);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t u0 = (int64_t)blockIdx.x * ROWS;
    const int split = blockIdx.y;
    const int64_t i_begin = (int64_t)split * a.items_per_split;
    const int64_t i_end = min(a.n_item, i_begin + a.items_per_split);
    const int n_tiles = (int)((i_end - i_begin + TC_N - 1) / TC_N);
    const int a_cols = KB * TC_D;                     // TMEM columns of the user rows

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(full + s), 1);
            mbar_init(smem_u32(bfree + s), 1);
        }
        for (int x = 0; x < NACC; ++x) {
            mbar_init(smem_u32(accfull + x), 1);
            mbar_init(smem_u32(accfree + x), 4);
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
    const uint32_t acc_base = tmem_base + (uint32_t)a_cols;

    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool is_epi = warp >= 2;
    const bool valid = is_epi && (u0 + row < a.nu);
    const float* urow = nullptr;
    float unorm2 = 0.f;
    if (is_epi) {
        if (valid) urow = a.user_table + __ldg(a.users + u0 + row) * (int64_t)D;
        for (int ch = 0; ch < 2 * KB; ++ch) {          // 32 floats per tcgen05.st
            uint32_t r[32];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) v = __ldg(reinterpret_cast<const float4*>(urow) + ch * 8 + c);
                unorm2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, unorm2))));
                r[4 * c + 0] = __float_as_uint(v.x);
                r[4 * c + 1] = __float_as_uint(v.y);
                r[4 * c + 2] = __float_as_uint(v.z);
                r[4 * c + 3] = __float_as_uint(v.w);
            }
            tmem_st32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), r);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        if (lane == 0) {
            for (int t = 0; t < n_tiles; ++t) {
                const int row0 = (int)(i_begin + (int64_t)t * TC_N);
                for (int kb = 0; kb < KB; ++kb) {
                    const int g = t * KB + kb, s = g % S;
                    mbar_wait(smem_u32(bfree + s), ((g / S) & 1) ^ 1);
                    const uint32_t bar = smem_u32(full + s);
                    mbar_expect_tx(bar, TC_TILE_BYTES);
                    const uint32_t dst = smem_u32(Bs + s * TC_TILE_BYTES);
                    tma_load_2d(dst, &item_map, bar, kb * TC_D, row0);
                    tma_load_2d(dst + TC_KH_BYTES, &item_map, bar, kb * TC_D + 32, row0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            for (int t = 0; t < n_tiles; ++t) {
                const int r = t % NACC;
                mbar_wait(smem_u32(accfree + r), ((t / NACC) & 1) ^ 1);
                const uint32_t d = acc_base + (uint32_t)(r * TC_N);
                for (int kb = 0; kb < KB; ++kb) {
                    const int g = t * KB + kb, s = g % S;
                    mbar_wait(smem_u32(full + s), (g / S) & 1);
                    tc_fence_after();
                    const uint32_t b0 = smem_u32(Bs + s * TC_TILE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t off = (uint32_t)((kk >> 2) * TC_KH_BYTES + (kk & 3) * 32);
                        umma_tf32_ts(d, tmem_base + (uint32_t)(kb * TC_D + kk * 8), umma_desc_sw128(b0 + off),
                                     (kb | kk) > 0);
                    }
                    umma_commit(smem_u32(bfree + s));       // the stage is free once these MMAs have read it
                }
                umma_commit(smem_u32(accfull + r));
            }
        }
    } else {
        float thr = -INFINITY, thr_lo = valid ? -INFINITY : INFINITY, margin = 0.f;
        float thr_sh = -INFINITY;
        float* sh_slot = (valid && a.shared_thr) ? a.shared_thr + (u0 + row) : nullptr;
        int cnt = 0, minpos = 0;
        int64_t tc = 0, te = 0;
        int32_t nxt = INT32_MAX;
        if (valid) {
            const int64_t u = __ldg(a.users + u0 + row);
            margin = TC_MARGIN * sqrtf(unorm2) * __ldg(a.item_maxnorm) + FLT_MIN;
            tc = __ldg(a.train_ptr + u);
            te = __ldg(a.train_ptr + u + 1);
            int64_t lo = tc, hi = te;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if ((int64_t)__ldg(a.train_items + mid) < i_begin) lo = mid + 1; else hi = mid;
            }
            tc = lo;
            nxt = tc < te ? __ldg(a.train_items + tc) : INT32_MAX;
        }
        for (int t = 0; t < n_tiles; ++t) {
            const int r = t % NACC;
            const int64_t it0 = i_begin + (int64_t)t * TC_N;
            const float sh_new = sh_slot ? __ldcg(sh_slot) : -INFINITY;
            mbar_wait(smem_u32(accfull + r), (t / NACC) & 1);
            tc_fence_after();
            if (sh_new > thr_sh) {
                thr_sh = sh_new;
                thr_lo = fmaxf(thr, thr_sh) - margin;
            }
            const uint32_t taddr = acc_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * TC_N);
            uint32_t cm[TC_N / 32];
            auto scan32 = [&](const uint32_t (&v)[32]) -> uint32_t {
                float m = __uint_as_float(v[0]);
#pragma unroll
                for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                uint32_t mask = 0;
                if (m > thr_lo) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) > thr_lo) ? (1u << j) : 0u;
                }
                return mask;
            };
#pragma unroll
            for (int c = 0; c < TC_N / 32; c += 2) {
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c * 32, v0);
                tmem_ld32(taddr + (c + 1) * 32, v1);
                tmem_ld_wait();
                cm[c] = scan32(v0);
                cm[c + 1] = scan32(v1);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(accfree + r));
            uint64_t lo64 = (uint64_t)cm[0] | ((uint64_t)cm[1] << 32), hi64 = (uint64_t)cm[2] | ((uint64_t)cm[3] << 32);
            while (lo64 | hi64) {
                int il;
                if (lo64) {
                    il = __ffsll((long long)lo64) - 1;
                    lo64 &= lo64 - 1;
                } else {
                    il = 64 + __ffsll((long long)hi64) - 1;
                    hi64 &= hi64 - 1;
                }
                const int64_t item = it0 + il;
                if (item >= i_end) break;
                if ((int64_t)nxt < item) {
                    int64_t step = 1, lo = tc + 1;
                    while (lo + step < te && (int64_t)__ldg(a.train_items + lo + step) < item) {
                        lo += step;
                        step <<= 1;
                    }
                    int64_t hi = min(te, lo + step + 1);
                    while (lo < hi) {
                        const int64_t mid = (lo + hi) >> 1;
                        if ((int64_t)__ldg(a.train_items + mid) < item) lo = mid + 1; else hi = mid;
                    }
                    tc = lo;
                    nxt = tc < te ? __ldg(a.train_items + tc) : INT32_MAX;
                }
                if ((int64_t)nxt == item) continue;
                // exact fp32 score, canonical sequential order over all D features
                const float4* irow = reinterpret_cast<const float4*>(a.item_table + item * (int64_t)D);
                const float4* ur4 = reinterpret_cast<const float4*>(urow);
                float ex = 0.f;
                for (int c16 = 0; c16 < D / 4; ++c16) {
                    const float4 uu = __ldg(ur4 + c16);
                    const float4 ii = __ldg(irow + c16);
                    ex = fmaf(uu.x, ii.x, ex);
                    ex = fmaf(uu.y, ii.y, ex);
                    ex = fmaf(uu.z, ii.z, ex);
                    ex = fmaf(uu.w, ii.w, ex);
                }
                if ((cnt < K || ex > thr) && ex >= thr_sh) {
                    const int pos = cnt < K ? cnt : minpos;
                    ls[(size_t)pos * ROWS + row] = ex;
                    li[(size_t)pos * ROWS + row] = (int32_t)item;
                    if (cnt < K) ++cnt;
                    if (cnt == K) {
                        float best = INFINITY;
                        int32_t besti = -1;
                        int bp = 0;
#pragma unroll 4
                        for (int j = 0; j < K; ++j) {
                            const float sj = ls[(size_t)j * ROWS + row];
                            const int32_t ij = li[(size_t)j * ROWS + row];
                            if (sj < best || (sj == best && ij > besti)) {
                                best = sj;
                                besti = ij;
                                bp = j;
                            }
                        }
                        thr = best;
                        minpos = bp;
                        thr_lo = fmaxf(thr, thr_sh) - margin;
                        if (sh_slot && thr > thr_sh) atomic_max_float(sh_slot, thr);
                    }
                }
            }
        }
        if (valid) {
            const size_t o = ((size_t)(u0 + row) * a.splits + split) * K;
            for (int j = 0; j < K; ++j) {
                a.part_scores[o + j] = j < cnt ? ls[(size_t)j * ROWS + row] : -INFINITY;
                a.part_ids[o + j] = j < cnt ? li[(size_t)j * ROWS + row] : -1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

__global__ void fill_f32_kernel(float* p, int64_t n, float v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// max_i ||I_i||_2 (for the TF32 error margin).  One warp per row, lanes stride over the row's float4s.
__global__ void __launch_bounds__(256) item_maxnorm_kernel(const float4* __restrict__ it, int64_t n_item, int c4,
                                                           float* out) {
    const int lane = threadIdx.x & 31;
    float best = 0.f;
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_item; r += stride) {
        float ss = 0.f;
        for (int c = lane; c < c4; c += 32) {
            const float4 v = __ldg(it + r * c4 + c);
            ss += dot4(v, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        best = fmaxf(best, ss);
    }
    if (lane == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(sqrtf(best) * 1.0001f));   // non-negative floats order as ints
}

// Merge the per-split K-best lists of one user (one warp per user); emit sigmoid scores.  If fewer than k un-masked
// items exist, the tail is the user's first train items with score -1024 (== (-score, id) order of the reference).
__global__ void __launch_bounds__(256)
eval_tc_merge_kernel(const float* __restrict__ ps, const int32_t* __restrict__ pi, const int64_t* __restrict__ users,
                     int64_t nu, int splits, int k, const int64_t* __restrict__ train_ptr,
                     const int32_t* __restrict__ train_items, int32_t* __restrict__ out_ids,
                     float* __restrict__ out_scores) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nu) return;
    const int n = splits * k;
    const float* s = ps + (size_t)w * n;
    const int32_t* id = pi + (size_t)w * n;
    int valid = 0;
    for (int i = lane; i < n; i += 32) {
        const float si = s[i];
        const int32_t ii = id[i];
        if (ii < 0) continue;
        ++valid;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float sj = s[j];
            const int32_t ij = id[j];
            rank += (ij >= 0) && ((sj > si) || (sj == si && ij < ii));
        }
        if (rank < k) {
            out_ids[(size_t)w * k + rank] = ii;
            out_scores[(size_t)w * k + rank] = 1.f / (1.f + expf(-si));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    if (valid < k) {
        const int64_t u = users[w];
        const int64_t tb = train_ptr[u], te = train_ptr[u + 1];
        for (int j = valid + lane; j < k; j += 32) {
            const int64_t src = tb + (j - valid);
            out_ids[(size_t)w * k + j] = src < te ? train_items[src] : -1;
            out_scores[(size_t)w * k + j] = src < te ? -1024.f : -INFINITY;
        }
    }
}

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static size_t tc_smem(int nh, int stages, int k) {
    const size_t a_smem = TC_TS ? 0 : (size_t)nh * TC_TILE_BYTES;
    return 1024 + a_smem + (size_t)stages * TC_TILE_BYTES + (size_t)TC_M * nh * TC_UPITCH * 4 +
           (size_t)2 * k * TC_M * nh * 4 + 256;
}

static size_t tc_smem_wide(int stages, int k) {
    return 1024 + (size_t)stages * TC_TILE_BYTES + (size_t)2 * k * TC_M * 4 + 256;
}

// TAGREC_EVAL_CG2=0 keeps the single-CTA kernel for every shape (A/B timing); =force runs the pair kernel for every
// 64-d shape its shared-memory budget allows (tests: partial tiles, tiny tables, many splits).
static int cg2_mode() {
    const char* e = getenv("TAGREC_EVAL_CG2");
    if (e && e[0] == '0') return 0;
    if (e && e[0] == 'f') return 2;
    return 1;
}

TcPlan tc_plan(int64_t nu, int64_t n_item, int dim, int k) {
    TcPlan p{};
    p.ok = false;
    if (dim < TC_D || dim % TC_D != 0 || dim > 4 * TC_D || k < 1 || k > 128 || nu < 1 || n_item < 1) return p;
    p.kb = dim / TC_D;
    const size_t budget = 227 * 1024;
    const int64_t item_tiles = (n_item + TC_N - 1) / TC_N;
    // Item splits: every split restarts its thresholds (about K * ln(items/K) exact re-scores per row and split) and
    // the merge ranks splits*K entries per user, so: at most TC_MAX_SPLITS, at least 8 tiles each, and no more than
    // fill ONE wave of the 148 SMs.
    const int64_t max_s = std::max<int64_t>(1, std::min<int64_t>(TC_MAX_SPLITS, item_tiles / 8));
    if (p.kb > 1) {     // wide tables: one 128-user half per CTA, stages released by tcgen05.commit
        p.nh = 1;
        p.stages = TC_MAX_STAGES;
        while (p.stages >= 2 && tc_smem_wide(p.stages, k) > budget) --p.stages;
        if (p.stages < 2) return p;
        p.smem = tc_smem_wide(p.stages, k);
    } else {
        // Two 128-user halves per CTA halve the L2 traffic of the item stream; with few users one half per CTA puts
        // twice as many CTAs on the machine.
        const int64_t tiles2 = (nu + 2 * TC_M - 1) / (2 * TC_M);
        p.nh = (nu > TC_M && tiles2 * max_s >= (kSMs * 3) / 4) ? 2 : 1;
        for (;; p.nh = 1) {
            p.stages = TC_MAX_STAGES;
            while (p.stages >= 2 && tc_smem(p.nh, p.stages, k) > budget) --p.stages;
            if (p.stages >= (p.nh == 2 ? 3 : 2)) break;
            if (p.nh == 1) return p;
        }
        p.smem = tc_smem(p.nh, p.stages, k);
        // CTA pairs (eval_tc2_kernel): whenever two 128-user halves per CTA would be used and the K-lists leave room for
        // at least 3 stages of the pair kernel's ring.
        const int mode = cg2_mode();
        // (below ~3 000 users the pair kernel's extra splits cost more than its instruction shape gains: measured
        // 2.51 vs 2.27 ms at 1 024 users, 2.42 vs 2.38 at 2 048, 3.24 vs 3.44 at 4 096, 4.94 vs 6.27 at 8 192)
        const int64_t pairs_avail = (nu + 2 * TC_M - 1) / (2 * TC_M);
        if ((p.nh == 2 && mode == 1 && pairs_avail >= 12) || mode == 2) {
            int st = TC_MAX_STAGES;
            while (st >= 3 && tc2_smem(st, k) > budget) --st;
            if (st >= 3 && (item_tiles >= 16 || mode == 2)) {
                p.cg2 = 1;
                p.stages = st;
                p.smem = tc2_smem(st, k);
            }
        }
    }
    if (p.cg2) {
        const int64_t pairs = (nu + 2 * TC_M - 1) / (2 * TC_M);
        const int64_t tiles256 = (n_item + 255) / 256;
        // Item splits: only as many as it takes to fill the 74 SM pairs.  More, shorter splits would even out the last
        // wave (64 user tiles: 0.875 of the one-split makespan with 8 splits on paper), but every (user tile, split)
        // unit pays the threshold start-up of its K-lists again: measured 8.3 / 8.6 / 9.4 / 10.4 / 14.5 ms for
        // 1 / 2 / 4 / 8 / 15 splits at 16 384 x 2 M (profiles/r2_eval_tc2.md).
        int64_t best_s = std::max<int64_t>(1, (kSMs / 2) / pairs);
        best_s = std::min<int64_t>(best_s, std::max<int64_t>(1, std::min<int64_t>(TC_MAX_SPLITS / 2, tiles256 / 8)));
        if (const char* e = getenv("TAGREC_EVAL_SPLITS")) {      // tuning override
            const int64_t v = atoll(e);
            if (v >= 1 && v <= TC_MAX_SPLITS / 2) best_s = std::min<int64_t>(v, std::max<int64_t>(1, tiles256));
        }
        p.splits = (int)best_s;
        p.items_per_split = ((tiles256 + p.splits - 1) / p.splits) * 256;
        p.splits = (int)((n_item + p.items_per_split - 1) / p.items_per_split);
        p.lists = 2 * p.splits;
        p.ok = true;
        return p;
    }
    const int64_t user_tiles = (nu + TC_M * p.nh - 1) / (TC_M * p.nh);
    int64_t best_s = user_tiles >= kSMs ? 1 : kSMs / user_tiles;
    best_s = std::max<int64_t>(1, std::min<int64_t>(best_s, max_s));
    p.splits = (int)best_s;
    p.items_per_split = ((item_tiles + p.splits - 1) / p.splits) * TC_N;
    p.splits = (int)((n_item + p.items_per_split - 1) / p.items_per_split);
    p.lists = p.splits;
    p.ok = true;
    return p;
}

int make_row_table_map(CUtensorMap* map, const float* table, int64_t n_rows, int dim, int box_rows) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail(TAGREC_ECUDA, "cuTensorMapEncodeTiled not available from the driver", __FILE__, __LINE__);
    const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)n_rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)dim * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(table), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TAGREC_ECUDA, "cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
    return TAGREC_OK;
}

int launch_item_maxnorm(const float* item_table, int64_t n_item, int dim, float* out, void* stream) {
    TAGREC_CUDA(cudaMemsetAsync(out, 0, 4, (cudaStream_t)stream));
    const int64_t nb = min((int64_t)kSMs * 8, (n_item + 7) / 8);
    TAGREC_LAUNCH(item_maxnorm_kernel, (unsigned)nb, 256, 0, stream, reinterpret_cast<const float4*>(item_table), n_item,
                  dim / 4, out);
    return TAGREC_OK;
}

int eval_topk_tc(const int64_t* users, int64_t nu, const float* user_table, const float* item_table, int64_t n_item,
                 const int64_t* train_ptr, const int32_t* train_items, int k, int32_t* topk_ids, float* topk_scores,
                 void* workspace, size_t workspace_bytes, void* stream, const TcPlan& p) {
    TAGREC_REQUIRE((reinterpret_cast<uintptr_t>(item_table) & 15) == 0, "item table must be 16-byte aligned");
    const size_t need = eval_tc_workspace_bytes(nu, p, k);
    if (!workspace || workspace_bytes < need) return fail(TAGREC_ENOMEM, "eval workspace too small", __FILE__, __LINE__);
    CUtensorMap map;
    const int dim = p.kb * TC_D;
    if (int rc = make_row_table_map(&map, item_table, n_item, dim)) return rc;

    TcArgs a{};
    a.users = users; a.nu = nu; a.user_table = user_table; a.item_table = item_table; a.n_item = n_item;
    a.train_ptr = train_ptr; a.train_items = train_items; a.k = k; a.splits = p.splits; a.stages = p.stages;
    a.items_per_split = p.items_per_split;
    float* maxnorm = reinterpret_cast<float*>(workspace);
    a.item_maxnorm = maxnorm;
    a.part_scores = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + 256);
    a.part_ids = reinterpret_cast<int32_t*>(a.part_scores + (size_t)nu * p.lists * k);
    // (the max over splits of their own K-th best is no tighter than one's own bound when two splits advance in
    // lockstep; it pays with many short splits, where late starters inherit the early ones' bounds)
    a.shared_thr = p.lists > 2 ? reinterpret_cast<float*>(a.part_ids + (size_t)nu * p.lists * k) : nullptr;
    if (a.shared_thr)
        TAGREC_LAUNCH(fill_f32_kernel, (unsigned)((nu + 255) / 256), 256, 0, stream, a.shared_thr, nu, -INFINITY);
    if (int rc = launch_item_maxnorm(item_table, n_item, dim, maxnorm, stream)) return rc;
    const dim3 grid((unsigned)((nu + TC_M * p.nh - 1) / (TC_M * p.nh)), (unsigned)p.splits);
    if (p.cg2) {
        if (int rc = launch_eval_tc2(&map, a, p.smem, stream)) return rc;
    } else if (p.kb > 1) {
        TAGREC_CUDA(cudaFuncSetAttribute(eval_tc_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        TAGREC_LAUNCH(eval_tc_wide_kernel, grid, 64 + 128, p.smem, stream, map, a, p.kb);
    } else if (p.nh == 2) {
        TAGREC_CUDA(cudaFuncSetAttribute(eval_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        TAGREC_LAUNCH(eval_tc_kernel<2>, grid, 64 + 256, p.smem, stream, map, a);
    } else {
        TAGREC_CUDA(cudaFuncSetAttribute(eval_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        TAGREC_LAUNCH(eval_tc_kernel<1>, grid, 64 + 128, p.smem, stream, map, a);
    }
    TAGREC_LAUNCH(eval_tc_merge_kernel, (unsigned)((nu + 7) / 8), 256, 0, stream, a.part_scores, a.part_ids, users, nu,
                  p.lists, k, train_ptr, train_items, topk_ids, topk_scores);
    return TAGREC_OK;
}

size_t eval_tc_workspace_bytes(int64_t nu, const TcPlan& p, int k) {
    return 256 + (size_t)nu * p.lists * k * 8 + (size_t)nu * 4;
}

}  // namespace tagrec
