// K3-TC2 — full-sort evaluation on CTA PAIRS: tcgen05.mma.cta_group::2, M256 x N256 x K8, kind::tf32.  sm_100a only.
//
// Same contract as eval_tc_kernel (eval_tc.cu; replaces model/lightgcn.py:84-89 + training/basic_test.py:42-48 for
// 64-d tables): TF32 scores are a FILTER, every candidate is re-scored in exact fp32 in the canonical sequential order
// and only exact scores enter the K-lists, so the lists are identical to the fp32 path's.
//
// Why pairs.  One M128 x N128 x K8 TF32 instruction keeps the tensor pipe busy for 64 cycles and costs ~116 from issue
// to issue (profiles/r1_eval_tc_experiments.md): the single-CTA kernel is bound at ~55 % of the array.  N = 256 doubles
// the work per instruction (128 array cycles for the same fixed cost), but a single CTA cannot double-buffer 256-column
// accumulators beside a TMEM-resident A operand, and the SS form with a 256-row B tile needs all of the SM's shared-
// memory bandwidth.  A CTA pair (thread-block cluster of 2 on one TPC) issues ONE instruction for both SMs: each CTA
// keeps its own 128 user rows (A, shared memory, 128B-swizzled) and HALF of the 256-item tile (B: 128 rows, its own TMA
// stream), the hardware shares the B halves, and each CTA's tensor memory receives its 128 rows x 256 columns of
// scores — two 256-column accumulators fill the 512 TMEM columns exactly.
//
//   warp 0      TMA producer (both CTAs)   its CTA's half of every item tile -> its own stage ring; complete_tx on the
//                                          LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2)
//   warp 1      MMA issuer (leader CTA)    8 x tcgen05.mma.cta_group::2 per tile; tcgen05.commit multicast to both
//                                          CTAs' "stage empty" and "accumulator full" barriers
//   warps 2-9   drain (both CTAs)          thread = (user row, 128-column half): tcgen05.ld, max-tree, one compare
//                                          with the list's filter threshold; the accumulator goes back to the leader's
//                                          MMA warp (remote mbarrier arrive) at once; candidates go into a ring in smem
//   warps 10-13 scorers (both CTAs)        exact fp32 re-scores, train-item masking (a cursor over the user's ascending
//                                          train row), K-lists and thresholds of two drain warps each
//
// Candidate path.  With every drain warp of both CTAs handing each accumulator back to ONE issuing thread, a warp that
// stops to re-score a candidate (an L2 / DRAM round trip plus the 64-step dependent fmaf chain of the canonical exact
// score) stalls the tensor pipe of two SMs: scoring candidates in the drain warps made this kernel 3x SLOWER than the
// single-CTA one (29.9 vs 10.9 ms on 16 384 x 2 M; 7.0 ms with the candidate path cut out, profiles/).  So the two
// jobs are decoupled: drain warps only filter and append (row, item) to their ring; scorer warps take up to 32 entries
// at a time — one per lane, item row from L2 (it was streamed through a moment ago), user row from the A tile — and
// then hand every result to the lane that keeps the row's list, in ring order (ascending item id per row, which the
// tie rule needs).  The filter threshold a drain thread reads (shared memory, written by the scorer) lags by the
// ring's contents: a stale threshold only lets more candidates through; acceptance uses the current exact threshold.
#include <algorithm>

#include "eval_tc.cuh"
#include "tc_ptx.cuh"

namespace tagrec {

constexpr int T2_N = 256;                   // items per tile over the pair (UMMA N)
constexpr int T2_NACC = 2;                  // accumulators (256 TMEM columns each)
constexpr int T2_DRAIN_WARPS = 8;           // warps that read the accumulators (2 per TMEM lane quarter: column halves)
#ifndef T2_SCORE_WARPS_N
#define T2_SCORE_WARPS_N 4
#endif
constexpr int T2_SCORE_WARPS = T2_SCORE_WARPS_N;   // warps that re-score candidates exactly and keep the K-lists
constexpr int T2_PER_SCORER = T2_DRAIN_WARPS / T2_SCORE_WARPS;   // drain warps served by one scorer warp
constexpr int T2_LISTS = 32 * T2_DRAIN_WARPS;   // K-lists per CTA: (user row, column half)
constexpr int T2_THREADS = 64 + 32 * (T2_DRAIN_WARPS + T2_SCORE_WARPS);
constexpr int T2_QCAP = 128;                // candidate ring entries per drain warp (power of two)
constexpr uint32_t T2_PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: the even CTA
// Instruction descriptor: D = f32, A = B = tf32, both K-major, N = 256, M = 256 (cta_group::2).
constexpr uint32_t T2_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(T2_N >> 3) << 17) |
                              ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();      // role branches end lane by lane; the non-.aligned forms below tolerate what divergence is left
    asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
// TMA load whose completion bytes go to the pair leader's barrier (same smem offset in the even CTA).
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & T2_PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(T2_IDESC), "r"(accumulate) : "memory");
}
// Arrives (once all MMAs issued so far have completed) on the barrier at this smem offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// Arrive on the pair leader's copy of a barrier (from either CTA).  The hand-off orders tensor-memory accesses only
// (tcgen05.fence::before_thread_sync precedes it); a .release.cluster arrive would add a GPU-wide MEMBAR per tile.
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & T2_PEER_MASK) : "memory");
}

// Orders this thread's shared-memory accesses for the ring hand-offs between a drain warp and its scorer warp
// (fence.acq_rel, not the sequentially-consistent membar.cta __threadfence_block() stands for).
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
eval_tc2_kernel(const __grid_constant__ CUtensorMap item_map, TcArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = a.stages, K = a.k;
    unsigned char* As = base;                                           // this CTA's 128 user rows, SW128 k-halves
    unsigned char* Bs = base + TC_TILE_BYTES;                           // S x 32 KB: this CTA's half of the item tiles
    float* ls = reinterpret_cast<float*>(Bs + S * TC_TILE_BYTES);       // [K][256] scores (unsorted K-lists)
    int32_t* li = reinterpret_cast<int32_t*>(ls + (size_t)K * T2_LISTS);
    // Candidate rings, one per drain warp (producer) read by its scorer warp (consumer).  An entry is ONE 64-bit word
    // {generation : 27 | row lane : 5 | item id : 32} written with a single st.shared.b64: the consumer recognises a
    // filled slot by its generation (slot use count + 1), so no tail counter and no fence sit between the two warps —
    // a fence in a drain warp waits for that warp's remote accumulator hand-off (~500 cycles per round, measured).
    unsigned long long* qent = reinterpret_cast<unsigned long long*>(li + (size_t)K * T2_LISTS);   // [8][T2_QCAP]
    float* thrlo_s = reinterpret_cast<float*>(qent + T2_DRAIN_WARPS * T2_QCAP);   // [256] filter threshold per list
    uint32_t* qhead = reinterpret_cast<uint32_t*>(thrlo_s + T2_LISTS);  // [8] entries consumed by the scorer warp
    uint32_t* qdone = qhead + T2_DRAIN_WARPS;                           // [8] drain warp has written its last entry
    int32_t* qidx = reinterpret_cast<int32_t*>(qdone + T2_DRAIN_WARPS); // [4 scorer warps][32] row -> batch entry
    uint64_t* bars = reinterpret_cast<uint64_t*>(qidx + T2_SCORE_WARPS * 32);
    uint64_t* full = bars;                                  // [S]  TMA (both CTAs) -> MMA; the LEADER's copy is used
    uint64_t* empty = bars + TC_MAX_STAGES;                 // [S]  MMA commit -> TMA producer of each CTA
    uint64_t* accfull = bars + 2 * TC_MAX_STAGES;           // [2]  MMA commit -> epilogue of each CTA
    uint64_t* accfree = accfull + T2_NACC;                  // [2]  epilogue warps of both CTAs -> MMA (leader's copy)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + T2_NACC);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t u0 = (int64_t)blockIdx.x * TC_M;          // cluster = blocks (2p, 2p+1): rows p*256 + rank*128
    const int split = blockIdx.y;
    const int64_t i_begin = (int64_t)split * a.items_per_split;
    const int64_t i_end = min(a.n_item, i_begin + a.items_per_split);
    const int n_tiles = (int)((i_end - i_begin + T2_N - 1) / T2_N);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(full + s), 1);
            mbar_init(smem_u32(empty + s), 1);
        }
        for (int x = 0; x < T2_NACC; ++x) {
            mbar_init(smem_u32(accfull + x), 1);
            mbar_init(smem_u32(accfree + x), 2 * T2_DRAIN_WARPS);   // the drain warps of both CTAs
        }
        for (int w = 0; w < T2_DRAIN_WARPS; ++w) qhead[w] = qdone[w] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    // user tile -> smem (generic proxy), laid out exactly like a TMA SWIZZLE_128B box pair
    for (int idx = tid; idx < TC_M * 16; idx += blockDim.x) {
        const int row = idx >> 4, c16 = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (u0 + row < a.nu) {
            const int64_t u = __ldg(a.users + u0 + row);
            v = __ldg(reinterpret_cast<const float4*>(a.user_table + u * TC_D) + c16);
        }
        *reinterpret_cast<float4*>(As + sw128_off(row, c16)) = v;
    }
    for (int idx = tid; idx < T2_DRAIN_WARPS * T2_QCAP; idx += blockDim.x) qent[idx] = 0ull;      // generation 0: empty
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    cluster_sync_all();                                     // barriers, TMEM and A tiles of BOTH CTAs are ready
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: this CTA's 128 rows of every 256-item tile =================
        if (lane == 0) {
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % S;
                mbar_wait(smem_u32(empty + s), ((t / S) & 1) ^ 1);
                const uint32_t bar = smem_u32(full + s);
                if (leader) mbar_expect_tx(bar, 2 * TC_TILE_BYTES);       // both halves report to the leader's barrier
                const uint32_t dst = smem_u32(Bs + s * TC_TILE_BYTES);
                const int row0 = (int)(i_begin + (int64_t)t * T2_N + (int64_t)rank * TC_N);
                tma_load_2d_cg2(dst, &item_map, bar, 0, row0);
                tma_load_2d_cg2(dst + TC_KH_BYTES, &item_map, bar, 32, row0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: one thread of the leader CTA drives both SMs =================
        if (leader && lane == 0) {
            const uint32_t a0 = smem_u32(As);
            for (int t = 0; t < n_tiles; ++t) {
                const int s = t % S, x = t & 1;
                mbar_wait(smem_u32(full + s), (t / S) & 1);
                mbar_wait(smem_u32(accfree + x), ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t b0 = smem_u32(Bs + s * TC_TILE_BYTES);
                const uint32_t d = tmem_base + (uint32_t)(x * T2_N);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {     // K = 8 per instruction: 4 per 128-byte swizzle row, 2 k-halves
                    const uint32_t off = (uint32_t)((kk >> 2) * TC_KH_BYTES + (kk & 3) * 32);
                    umma_tf32_cg2(d, umma_desc_sw128(a0 + off), umma_desc_sw128(b0 + off), kk > 0);
                }
                umma_commit_pair(smem_u32(empty + s));       // both CTAs' stage s may be refilled
                umma_commit_pair(smem_u32(accfull + x));     // both CTAs' epilogues may drain accumulator x
            }
        }
    } else if (warp < 2 + T2_DRAIN_WARPS) {
        // ================= drain: thread = (user row, 128-column half) =================
        const int dw = warp - 2;                   // 0..7
        const int q = warp & 3;                    // TMEM lane quarter this warp may access
        const int ch = dw >> 2;                    // column half of the 256-wide accumulator
        const int row = q * 32 + lane;             // row of this CTA's user tile
        const int et = dw * 32 + lane;             // K-list of this thread (kept by the scorer warp dw / T2_PER_SCORER)
        const bool valid = u0 + row < a.nu;
        volatile unsigned long long* my_ring = qent + dw * T2_QCAP;
        volatile uint32_t* my_head = qhead + dw;
        volatile float* my_thrlo = thrlo_s + et;
        uint32_t tail = 0, head_seen = 0;          // warp-uniform
        *my_thrlo = valid ? -INFINITY : INFINITY;  // the scorer only ever raises it
        for (int t = 0; t < n_tiles; ++t) {
            const int x = t & 1;
            const int64_t it0 = i_begin + (int64_t)t * T2_N + ch * 128;
            const float thr_lo = *my_thrlo;        // exact K-th best minus the TF32 margin, as of the scorer's last batch
            mbar_wait(smem_u32(accfull + x), (t >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(x * T2_N + ch * 128);
            // ---- 128 TF32 scores of this row -> a 128-bit candidate mask (usually empty) ----
            uint32_t cm[4];
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
            for (int c = 0; c < 4; c += 2) {       // two 32-column loads in flight per wait
                uint32_t v0[32], v1[32];
                tmem_ld32(taddr + c * 32, v0);
                tmem_ld32(taddr + (c + 1) * 32, v1);
                tmem_ld_wait();
                cm[c] = scan32(v0);
                cm[c + 1] = scan32(v1);
            }
            // the accumulator half is drained: hand it back to the leader's MMA warp before any candidate work
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(smem_u32(accfree + x));
            // ---- candidates -> this warp's ring (warp-convergent rounds: at most one entry per lane and round; the
            // common tile has none: one ballot).  Parking a tile's masks when the next accumulator is already full
            // (drain first, enqueue later) was measured and is slower (9.1 vs 8.2 ms): the extra poll costs more than
            // the skew it removes. ----
            uint64_t lo64 = (uint64_t)cm[0] | ((uint64_t)cm[1] << 32), hi64 = (uint64_t)cm[2] | ((uint64_t)cm[3] << 32);
            while (__ballot_sync(0xffffffffu, (lo64 | hi64) != 0ull)) {
                int32_t cand = -1;
                if (lo64 | hi64) {
                    int il;
                    if (lo64) {
                        il = __ffsll((long long)lo64) - 1;
                        lo64 &= lo64 - 1;
                    } else {
                        il = 64 + __ffsll((long long)hi64) - 1;
                        hi64 &= hi64 - 1;
                    }
                    const int64_t item = it0 + il;
                    if (item < i_end) cand = (int32_t)item;
                    else lo64 = hi64 = 0;                    // zero-filled rows past the split / table end
                }
                const unsigned b = __ballot_sync(0xffffffffu, cand >= 0);
                if (!b) continue;
                const uint32_t n = (uint32_t)__popc(b);
                if (tail + n - head_seen > (uint32_t)T2_QCAP) {          // ring full: wait for the scorer (start-up only)
                    uint32_t spins = 0;
                    do {
                        head_seen = *my_head;
                        if (++spins > (1u << 26)) {
                            printf("tagrec eval_tc2: candidate ring stuck (block %d,%d warp %d)\n", blockIdx.x, blockIdx.y, warp);
                            __trap();
                        }
                    } while (tail + n - head_seen > (uint32_t)T2_QCAP);
                }
                if (cand >= 0) {
                    const uint32_t seq = tail + (uint32_t)__popc(b & ((1u << lane) - 1u));
                    const unsigned long long gen = (unsigned long long)(seq / T2_QCAP + 1u);
                    my_ring[seq % T2_QCAP] = (gen << 37) | ((unsigned long long)lane << 32) | (unsigned long long)(uint32_t)cand;
                }
                tail += n;
            }
        }
        __syncwarp();
        if (lane == 1) {
            fence_cta();
            *(volatile uint32_t*)(qdone + dw) = 1u;
        }
    } else {
        // ================= scorers: exact fp32 re-scores + the K-lists of two drain warps each =================
        const int sj = warp - 2 - T2_DRAIN_WARPS;  // 0..3
        int32_t* my_idx = qidx + sj * 32;
        struct ListState {
            float thr, thr_sh, margin;
            int cnt, minpos;
            float* sh_slot;
            bool valid;
            int64_t tc, te;                        // cursor over the user's ascending train row
            int32_t nxt;                           // smallest train item of this user not yet passed
        } st[T2_PER_SCORER];
        uint32_t head[T2_PER_SCORER] = {};
#pragma unroll
        for (int w = 0; w < T2_PER_SCORER; ++w) {
            const int dw = T2_PER_SCORER * sj + w;
            const int row = ((dw + 2) & 3) * 32 + lane;
            ListState& L = st[w];
            L.valid = u0 + row < a.nu;
            L.thr = -INFINITY;
            L.thr_sh = -INFINITY;
            L.cnt = 0;
            L.minpos = 0;
            L.margin = 0.f;
            L.sh_slot = (L.valid && a.shared_thr) ? a.shared_thr + (u0 + row) : nullptr;
            L.tc = L.te = 0;
            L.nxt = INT32_MAX;
            if (L.valid) {
                const int64_t u = __ldg(a.users + u0 + row);
                L.tc = __ldg(a.train_ptr + u);
                L.te = __ldg(a.train_ptr + u + 1);
                int64_t lo = L.tc, hi = L.te;      // first train item inside this split
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if ((int64_t)__ldg(a.train_items + mid) < i_begin) lo = mid + 1; else hi = mid;
                }
                L.tc = lo;
                L.nxt = L.tc < L.te ? __ldg(a.train_items + L.tc) : INT32_MAX;
                float unorm2 = 0.f;
#pragma unroll
                for (int c16 = 0; c16 < 16; ++c16) {
                    const float4 v = *reinterpret_cast<const float4*>(As + sw128_off(row, c16));
                    unorm2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, unorm2))));
                }
                L.margin = TC_MARGIN * sqrtf(unorm2) * __ldg(a.item_maxnorm) + FLT_MIN;
            }
        }
        // Scores n (<= 32) ring entries of drain warp dw (one per lane, exact fp32, canonical order) and hands every
        // result to the lane that keeps the row's K-list, in ring order (ascending item id per row: the tie rule).
        auto process = [&](ListState& L, const int dw, uint32_t& hd, const int n, const unsigned long long ent) {
            const int et = dw * 32 + lane;
            int32_t item = 0, r = 32 + lane;                             // sentinel rows never match
            if (lane < n) {
                item = (int32_t)(uint32_t)ent;
                r = (int)((ent >> 32) & 31ull);
            }
            // Entries of the same row must enter its list one after the other, in ring order; entries of different rows
            // are inserted at the same time by their owners.  rank = how many earlier entries of the batch share the row:
            // pass p inserts every row's p-th entry (one pass for most batches).
            const unsigned same = __match_any_sync(0xffffffffu, r);
            const int rank = __popc(same & ((1u << lane) - 1u));
            const bool mine = lane < n;
            const int passes = (int)__reduce_max_sync(0xffffffffu, mine ? (unsigned)rank + 1u : 0u);
            hd += (uint32_t)n;
            __syncwarp();
            if (lane == 0) *(volatile uint32_t*)(qhead + dw) = hd;       // the slots may be refilled (their contents are
                                                                         // in registers: the ballot above depended on them)
            // another list of this user may have raised the bound (a stale read only prunes less)
            if (L.sh_slot) {
                const float sh_new = __ldcg(L.sh_slot);
                if (sh_new > L.thr_sh) L.thr_sh = sh_new;
            }
            float ex = 0.f;
            if (mine) {
                const float4* irow = reinterpret_cast<const float4*>(a.item_table + (int64_t)item * TC_D);
                float4 iv[16];
#pragma unroll
                for (int c16 = 0; c16 < 16; ++c16) iv[c16] = __ldg(irow + c16);
                const int urow = ((dw + 2) & 3) * 32 + r;
#pragma unroll
                for (int c16 = 0; c16 < 16; ++c16) {
                    const float4 uu = *reinterpret_cast<const float4*>(As + sw128_off(urow, c16));
                    ex = fmaf(uu.x, iv[c16].x, ex);
                    ex = fmaf(uu.y, iv[c16].y, ex);
                    ex = fmaf(uu.z, iv[c16].z, ex);
                    ex = fmaf(uu.w, iv[c16].w, ex);
                }
            }
            bool raised = false;
            for (int p = 0; p < passes; ++p) {
                const bool act = mine && rank == p;
                if (act) my_idx[r] = lane;
                const unsigned present = __reduce_or_sync(0xffffffffu, act ? (1u << r) : 0u);
                __syncwarp();
                // every lane that keeps a row present in this pass takes its entry's result
                const bool own = (present >> lane) & 1u;
                const int src = own ? my_idx[lane] : 0;
                const float exe = __shfl_sync(0xffffffffu, ex, src);
                const int32_t ite = __shfl_sync(0xffffffffu, item, src);
                bool take = own && (L.cnt < K || exe > L.thr) && exe >= L.thr_sh;
                if (take) {
                    // train-item cursor (items of a list arrive in ascending order): nxt = smallest train item >= ite
                    // (galloping, then bisection).  Only results that would enter the list are looked up.
                    if (L.nxt < ite) {
                        int64_t step = 1, lo = L.tc + 1;
                        while (lo + step < L.te && __ldg(a.train_items + lo + step) < ite) {
                            lo += step;
                            step <<= 1;
                        }
                        int64_t hi = min(L.te, lo + step + 1);
                        while (lo < hi) {
                            const int64_t mid = (lo + hi) >> 1;
                            if (__ldg(a.train_items + mid) < ite) lo = mid + 1; else hi = mid;
                        }
                        L.tc = lo;
                        L.nxt = L.tc < L.te ? __ldg(a.train_items + L.tc) : INT32_MAX;
                    }
                    take = L.nxt != ite;                      // masked (basic_test.py:47)
                }
                // a row's items arrive in ascending id order, so on a score tie the incumbent (smaller id) stays:
                // strict >.  Against the bound of ANOTHER list only strictly smaller scores may be dropped.
                if (take) {
                    const int pos = L.cnt < K ? L.cnt : L.minpos;
                    ls[(size_t)pos * T2_LISTS + et] = exe;
                    li[(size_t)pos * T2_LISTS + et] = ite;
                    if (L.cnt < K) ++L.cnt;
                    if (L.cnt == K) {     // new evictee: lowest score, largest id among equals
                        float best = INFINITY;
                        int32_t besti = -1;
                        int bp = 0;
#pragma unroll 4
                        for (int j = 0; j < K; ++j) {
                            const float sj2 = ls[(size_t)j * T2_LISTS + et];
                            const int32_t ij = li[(size_t)j * T2_LISTS + et];
                            if (sj2 < best || (sj2 == best && ij > besti)) {
                                best = sj2;
                                besti = ij;
                                bp = j;
                            }
                        }
                        L.thr = best;
                        L.minpos = bp;
                        raised = true;
                    }
                }
                __syncwarp();
            }
            if (L.valid && (raised || L.sh_slot)) {
                const float bound = fmaxf(L.thr, L.thr_sh);
                if (bound > -INFINITY) *(volatile float*)(thrlo_s + et) = bound - L.margin;
                if (raised && L.sh_slot && L.thr > L.thr_sh) atomic_max_float(L.sh_slot, L.thr);
            }
            __syncwarp();
        };
        int idle = 0;
        uint32_t guard = 0;
        for (;;) {
            bool progressed = false;
            bool finished = true;
#pragma unroll
            for (int w = 0; w < T2_PER_SCORER; ++w) {
                const int dw = T2_PER_SCORER * sj + w;
                // "done" is read before the entries: a set flag means every entry of this ring has been written
                const uint32_t dn = *(volatile uint32_t*)(qdone + dw);
                const uint32_t seq = head[w] + (uint32_t)lane;
                const unsigned long long ent = *(volatile unsigned long long*)(qent + dw * T2_QCAP + seq % T2_QCAP);
                const bool filled = (ent >> 37) == (unsigned long long)(seq / T2_QCAP + 1u);
                const unsigned fb = __ballot_sync(0xffffffffu, filled);
                const int avail = __ffs(~fb) - 1 < 0 ? 32 : __ffs(~fb) - 1;            // filled prefix (entries are written in order
                                                                                        // of a round, rounds in order)
                // full batches amortise the memory round trip and the dependent fmaf chain; a partial batch runs only
                // after ~5 us without one (thresholds a few tiles stale cost next to nothing) or at the end
                if (avail == 32 || (avail > 0 && (idle >= 32 || dn))) {
                    process(st[w], dw, head[w], avail, ent);
                    progressed = true;
                }
                if (!dn || avail > 0) finished = false;
            }
            if (progressed) {
                idle = 0;
                guard = 0;
                continue;
            }
            if (finished) break;
            ++idle;
            __nanosleep(128);
            if (++guard > (1u << 24)) {
                printf("tagrec eval_tc2: scorer stuck (block %d,%d warp %d)\n", blockIdx.x, blockIdx.y, warp);
                __trap();
            }
        }
#pragma unroll
        for (int w = 0; w < T2_PER_SCORER; ++w) {
            const int dw = T2_PER_SCORER * sj + w;
            const int row = ((dw + 2) & 3) * 32 + lane;
            const int ch = dw >> 2;
            const int et = dw * 32 + lane;
            if (st[w].valid) {
                const size_t o = ((size_t)(u0 + row) * (a.splits * 2) + split * 2 + ch) * K;
                for (int j = 0; j < K; ++j) {
                    a.part_scores[o + j] = j < st[w].cnt ? ls[(size_t)j * T2_LISTS + et] : -INFINITY;
                    a.part_ids[o + j] = j < st[w].cnt ? li[(size_t)j * T2_LISTS + et] : -1;
                }
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();            // neither CTA leaves (or frees TMEM) while the other may still signal / be signalled
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

size_t tc2_smem(int stages, int k) {
    return 1024 + (size_t)(1 + stages) * TC_TILE_BYTES + (size_t)2 * k * T2_LISTS * 4 + (size_t)T2_DRAIN_WARPS * T2_QCAP * 8 +
           (size_t)T2_LISTS * 4 + 2 * T2_DRAIN_WARPS * 4 + T2_SCORE_WARPS * 32 * 4 + 256;
}

int launch_eval_tc2(const void* item_map, const TcArgs& a, size_t smem, void* stream) {
    const int64_t pairs = (a.nu + 2 * TC_M - 1) / (2 * TC_M);
    const dim3 grid((unsigned)(2 * pairs), (unsigned)a.splits);
    TAGREC_CUDA(cudaFuncSetAttribute(eval_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TAGREC_LAUNCH(eval_tc2_kernel, grid, T2_THREADS, smem, stream, *reinterpret_cast<const CUtensorMap*>(item_map), a);
    return TAGREC_OK;
}

}  // namespace tagrec
