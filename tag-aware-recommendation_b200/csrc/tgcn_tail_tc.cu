// K7-TC — forward of the TGCN dense tail (tgcn_tail.cu, T1) on the 5th-generation tensor cores.  sm_100a.
//
// Same contract as tgcn_tail_fwd_kernel (model/tgcn.py:86-106: bit-level conv, concat with the vector-level features,
// 2096 -> 64 fusion layer, bias, ReLU).  The [128 nodes x 2096] x [2096 x 64] product runs as 3xTF32 tcgen05 MMAs
//     out = A_hi W_hi + A_hi W_lo + A_lo W_hi,    x_hi = x with its low 13 mantissa bits cleared, x_lo = x - x_hi
// (both exact; the dropped lo.lo term and the TF32 truncation of the lo operands are <= 3 * 2^-20 relative to |a||w|,
// the accumulation is fp32 in TMEM) — fp32-level accuracy, the tolerance of the parity tests is unchanged.
// The A operand is not in memory: feature (c, d) of a node = relu(wb[c,.] . z[node,.,d]) is GENERATED per 64-feature
// chunk by eight warps straight into the 128B-swizzled K-major shared-memory layout the MMA reads (hi tile | lo tile),
// from the node's z rows held in registers for the whole kernel.  The W chunks (pre-transposed and pre-split by a
// tiny kernel: [chunk][out][feature], hi and lo) arrive by TMA.
//
//   warp 0      TMA: per chunk 4 boxes (hi / lo x two 128-byte k-halves) of W^T              2-stage ring
//   warp 1      MMA: 24 tcgen05.mma.kind::tf32 (M128 x N64 x K8) per chunk into ONE 64-column accumulator
//   warps 2-9   generators; afterwards warps 2-5 drain the accumulator: + bias, ReLU, store
// One CTA per 128-node tile.  2 * 128 * 2096 * 64 flop per tile in 792 MMAs.
#include <algorithm>

#include "tc_ptx.cuh"

namespace tagrec {

constexpr int KT_W = 64;                        // layer width = outputs = features per chunk
constexpr int KT_M = 128;                       // nodes per tile
constexpr int KT_A_TILE = KT_M * KT_W * 4;      // 32 KB: [2 k-halves][128 rows][128 B]
constexpr int KT_B_TILE = KT_W * KT_W * 4;      // 16 KB: [2 k-halves][64 rows][128 B]
constexpr int KT_B_KH = KT_W * 128;             // 8 KB per k-half of a W tile
constexpr int KT_STAGES = 2;
constexpr int KT_THREADS = 64 + 256;
// instruction descriptor: D = f32, A = B = tf32, K-major both, N = 64, M = 128
constexpr uint32_t KT_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KT_W >> 3) << 17) |
                              ((uint32_t)(KT_M >> 4) << 24);

__device__ __forceinline__ void umma_tf32_ss64(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(KT_IDESC), "r"(accumulate) : "memory");
}

__device__ __forceinline__ float kt_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// The lo part rounded to TF32 (nearest): the tensor core would TRUNCATE the 13-bit lo operand to 10 bits, a bias of one
// sign that adds up over the 2096-term sums of this layer; rounded here, the operand is TF32-exact and the error has
// no preferred sign.
__device__ __forceinline__ float kt_lo(float x, float h) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x - h));
    return __uint_as_float(r);
}
constexpr int KT_NACC = 8;          // partial accumulators (64 TMEM columns each), summed in fp32 by the epilogue: the
                                    // tensor core's fp32 accumulation rounds toward zero, so a chain of 792 MMAs into
                                    // one accumulator carries a systematic error ~ 792 * 2^-25; 8 chains of <= 120 do not

// wt[h][c][o][d] = hi / lo part of Wf[c*64 + d][o]   (rows past the table: 0)
__global__ void tgcn_tail_split_kernel(const float* __restrict__ wf, int C, int E, float* __restrict__ wt) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int NC = C + (E > 0 ? 1 : 0);
    if (idx >= NC * KT_W * KT_W) return;
    const int c = idx / (KT_W * KT_W), o = (idx / KT_W) % KT_W, d = idx % KT_W;
    const float w = (c < C || d < E) ? __ldg(wf + ((size_t)c * KT_W + d) * KT_W + o) : 0.f;
    const float h = kt_hi(w);
    wt[idx] = h;
    wt[(size_t)NC * KT_W * KT_W + idx] = kt_lo(w, h);
}

__global__ void __launch_bounds__(KT_THREADS, 1)
tgcn_tail_fwd_tc_kernel(const __grid_constant__ CUtensorMap wt_map, const float* __restrict__ z,
                        const float* __restrict__ wb, const float* __restrict__ xf, const float* __restrict__ bf,
                        int64_t n, int C, int E, float* __restrict__ out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* As = base;                                   // [S] x (hi 32 KB | lo 32 KB)
    unsigned char* Bs = As + KT_STAGES * 2 * KT_A_TILE;         // [S] x (hi 16 KB | lo 16 KB)
    float* WB = reinterpret_cast<float*>(Bs + KT_STAGES * 2 * KT_B_TILE);      // [C][3]
    uint64_t* bars = reinterpret_cast<uint64_t*>(WB + ((3 * C + 3) & ~3));
    uint64_t* afull = bars;                     // [S] generators -> MMA
    uint64_t* bfull = afull + KT_STAGES;        // [S] TMA -> MMA
    uint64_t* sfree = bfull + KT_STAGES;        // [S] MMA (commit) -> generators, TMA
    uint64_t* accfull = sfree + KT_STAGES;      // [1] MMA -> epilogue
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfull + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * KT_M;
    const int NC = C + (E > 0 ? 1 : 0);

    if (tid == 0) {
        for (int s = 0; s < KT_STAGES; ++s) {
            mbar_init(smem_u32(afull + s), 8);
            mbar_init(smem_u32(bfull + s), 1);
            mbar_init(smem_u32(sfree + s), 1);
        }
        mbar_init(smem_u32(accfull), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 3 * C; i += KT_THREADS) WB[i] = __ldg(wb + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: W^T chunks, hi and lo =================
        if (lane == 0) {
            for (int c = 0; c < NC; ++c) {
                const int s = c % KT_STAGES;
                mbar_wait(smem_u32(sfree + s), ((c / KT_STAGES) & 1) ^ 1);
                const uint32_t bar = smem_u32(bfull + s);
                mbar_expect_tx(bar, 2 * KT_B_TILE);
                const uint32_t dst = smem_u32(Bs + s * 2 * KT_B_TILE);
#pragma unroll
                for (int h = 0; h < 2; ++h) {                   // rows of part h of chunk c: (h * NC + c) * 64 ...
                    const int row0 = (h * NC + c) * KT_W;
                    tma_load_2d(dst + h * KT_B_TILE, &wt_map, bar, 0, row0);
                    tma_load_2d(dst + h * KT_B_TILE + KT_B_KH, &wt_map, bar, 32, row0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            for (int c = 0; c < NC; ++c) {
                const int s = c % KT_STAGES;
                mbar_wait(smem_u32(afull + s), (c / KT_STAGES) & 1);
                mbar_wait(smem_u32(bfull + s), (c / KT_STAGES) & 1);
                tc_fence_after();
                const uint32_t ah = smem_u32(As + s * 2 * KT_A_TILE), al = ah + KT_A_TILE;
                const uint32_t bh = smem_u32(Bs + s * 2 * KT_B_TILE), bl = bh + KT_B_TILE;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {                 // K = 8 per instruction: 4 per swizzle row, 2 k-halves
                    const uint32_t ao = (uint32_t)((kk >> 2) * TC_KH_BYTES + (kk & 3) * 32);
                    const uint32_t bo = (uint32_t)((kk >> 2) * KT_B_KH + (kk & 3) * 32);
                    const uint32_t d = tmem_acc + (uint32_t)((c % KT_NACC) * KT_W);
                    umma_tf32_ss64(d, umma_desc_sw128(al + ao), umma_desc_sw128(bh + bo), c >= KT_NACC || kk != 0);
                    umma_tf32_ss64(d, umma_desc_sw128(ah + ao), umma_desc_sw128(bl + bo), 1);
                    umma_tf32_ss64(d, umma_desc_sw128(ah + ao), umma_desc_sw128(bh + bo), 1);
                }
                umma_commit(smem_u32(sfree + s));
            }
            umma_commit(smem_u32(accfull));
        }
    } else {
        // ================= generators: thread = 8 x (node, four consecutive dims) =================
        const int gt = tid - 64;
        float4 zr[8][3];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = gt + 256 * it, node = idx >> 4, c16 = idx & 15;
            const bool ok = n0 + node < n;
#pragma unroll
            for (int r = 0; r < 3; ++r)
                zr[it][r] = ok ? __ldg(reinterpret_cast<const float4*>(z + ((n0 + node) * 3 + r) * KT_W) + c16)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int c = 0; c < NC; ++c) {
            const int s = c % KT_STAGES;
            mbar_wait(smem_u32(sfree + s), ((c / KT_STAGES) & 1) ^ 1);
            unsigned char* hi = As + s * 2 * KT_A_TILE;
            unsigned char* lo = hi + KT_A_TILE;
            float w0 = 0.f, w1 = 0.f, w2 = 0.f;
            if (c < C) {
                w0 = WB[3 * c];
                w1 = WB[3 * c + 1];
                w2 = WB[3 * c + 2];
            }
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int idx = gt + 256 * it, node = idx >> 4, c16 = idx & 15;
                float4 f;
                if (c < C) {            // the SAME expression as bit_pre() of tgcn_tail.cu (the backward's masks)
                    f.x = fmaxf(fmaf(w2, zr[it][2].x, fmaf(w1, zr[it][1].x, w0 * zr[it][0].x)), 0.f);
                    f.y = fmaxf(fmaf(w2, zr[it][2].y, fmaf(w1, zr[it][1].y, w0 * zr[it][0].y)), 0.f);
                    f.z = fmaxf(fmaf(w2, zr[it][2].z, fmaf(w1, zr[it][1].z, w0 * zr[it][0].z)), 0.f);
                    f.w = fmaxf(fmaf(w2, zr[it][2].w, fmaf(w1, zr[it][1].w, w0 * zr[it][0].w)), 0.f);
                } else {                // the vector-level features of the caller (E <= 64 of them, zero padded)
                    f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n0 + node < n && 4 * c16 < E) f = __ldg(reinterpret_cast<const float4*>(xf + (n0 + node) * E) + c16);
                }
                const float4 h = make_float4(kt_hi(f.x), kt_hi(f.y), kt_hi(f.z), kt_hi(f.w));
                const uint32_t off = sw128_off(node, c16);
                *reinterpret_cast<float4*>(hi + off) = h;
                *reinterpret_cast<float4*>(lo + off) =
                    make_float4(kt_lo(f.x, h.x), kt_lo(f.y, h.y), kt_lo(f.z, h.z), kt_lo(f.w, h.w));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(afull + s));
        }
        // ================= epilogue: warps 2-5, thread = node row =================
        if (warp < 6) {
            const int q = warp & 3;
            const int64_t node = n0 + q * 32 + lane;
            mbar_wait(smem_u32(accfull), 0);
            tc_fence_after();
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16);
            const int nacc = min(KT_NACC, NC);
            float4* orow = reinterpret_cast<float4*>(out + node * KT_W);
            const float4* b4 = reinterpret_cast<const float4*>(bf);
#pragma unroll
            for (int half = 0; half < 2; ++half) {              // 32 output columns at a time
                float sum[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) sum[j] = 0.f;
                for (int x = 0; x < nacc; x += 2) {              // two loads in flight; pairs added first
                    uint32_t v0[32], v1[32];
                    tmem_ld32(taddr + (uint32_t)(x * KT_W + half * 32), v0);
                    if (x + 1 < nacc) tmem_ld32(taddr + (uint32_t)((x + 1) * KT_W + half * 32), v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        sum[j] += __uint_as_float(v0[j]) + (x + 1 < nacc ? __uint_as_float(v1[j]) : 0.f);
                }
                if (node < n) {
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 b = __ldg(b4 + half * 8 + c4);
                        orow[half * 8 + c4] = make_float4(fmaxf(sum[4 * c4 + 0] + b.x, 0.f), fmaxf(sum[4 * c4 + 1] + b.y, 0.f),
                                                          fmaxf(sum[4 * c4 + 2] + b.z, 0.f), fmaxf(sum[4 * c4 + 3] + b.w, 0.f));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ T2 on tcgen05
// Backward to z of the same layer (tgcn_tail.cu, T2): per 64-feature chunk  G_c = g_pre Wf_c^T  is one fresh
// 128 x 64 accumulator (24 MMAs: no long accumulation chain), read back by 16 epilogue warps (thread = node x 16
// feature dims) that apply the recomputed ReLU mask and fold it into g_z, g_wb and (last chunk) g_xf.
//   A operand: the tile's g_pre = g_out * (out > 0), split hi | lo ONCE per tile into the swizzled layout (K = outputs)
//   B operand: rows c*64 .. c*64+63 of Wf (feature-major, K = outputs: the parameter's own layout), hi / lo copies by TMA
constexpr int KZ_NACC = 4;
constexpr int KZ_THREADS = 64 + 512;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

// ws[h][c][d][o] = hi / lo part of Wf[c*64 + d][o]   (rows past the table: 0)
__global__ void tgcn_tail_split_rows_kernel(const float* __restrict__ wf, int C, int E, float* __restrict__ ws) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int NC = C + (E > 0 ? 1 : 0);
    if (idx >= NC * KT_W * KT_W) return;
    const int c = idx / (KT_W * KT_W), d = (idx / KT_W) % KT_W;
    const float w = (c < C || d < E) ? __ldg(wf + idx) : 0.f;
    const float h = kt_hi(w);
    ws[idx] = h;
    ws[(size_t)NC * KT_W * KT_W + idx] = kt_lo(w, h);
}

__global__ void __launch_bounds__(KZ_THREADS, 1)
tgcn_tail_bwd_z_tc_kernel(const __grid_constant__ CUtensorMap ws_map, const float* __restrict__ g_out,
                          const float* __restrict__ out, const float* __restrict__ z, const float* __restrict__ wb,
                          int64_t n, int C, int E, float* __restrict__ g_pre, float* __restrict__ g_z,
                          float* __restrict__ g_wb, float* __restrict__ g_xf, float* __restrict__ g_bf) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* As = base;                                   // g_pre: hi 32 KB | lo 32 KB
    unsigned char* Bs = As + 2 * KT_A_TILE;                     // [S] x (hi 16 KB | lo 16 KB)
    float* WB = reinterpret_cast<float*>(Bs + KT_STAGES * 2 * KT_B_TILE);      // [C][3]
    float* GWB = WB + ((3 * C + 3) & ~3);                       // [16 warps][C][3] per-warp partial sums of g_wb
    float* BF = GWB + 16 * ((3 * C + 3) & ~3);                  // [16][64] partial column sums of g_pre
    uint64_t* bars = reinterpret_cast<uint64_t*>(BF + 16 * KT_W);
    uint64_t* bfull = bars;                     // [S]    TMA -> MMA
    uint64_t* sfree = bfull + KT_STAGES;        // [S]    MMA (commit) -> TMA
    uint64_t* accfull = sfree + KT_STAGES;      // [NACC] MMA -> epilogue
    uint64_t* accfree = accfull + KZ_NACC;      // [NACC] epilogue -> MMA
    uint64_t* aready = accfree + KZ_NACC;       // [1]    g_pre tiles written
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aready + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * KT_M;
    const int NC = C + (E > 0 ? 1 : 0);
    const int c3 = (3 * C + 3) & ~3;

    if (tid == 0) {
        for (int s = 0; s < KT_STAGES; ++s) {
            mbar_init(smem_u32(bfull + s), 1);
            mbar_init(smem_u32(sfree + s), 1);
        }
        for (int x = 0; x < KZ_NACC; ++x) {
            mbar_init(smem_u32(accfull + x), 1);
            mbar_init(smem_u32(accfree + x), 16);
        }
        mbar_init(smem_u32(aready), 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 3 * C; i += KZ_THREADS) WB[i] = __ldg(wb + i);
    for (int i = tid; i < 16 * c3; i += KZ_THREADS) GWB[i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int c = 0; c < NC; ++c) {
                const int s = c % KT_STAGES;
                mbar_wait(smem_u32(sfree + s), ((c / KT_STAGES) & 1) ^ 1);
                const uint32_t bar = smem_u32(bfull + s);
                mbar_expect_tx(bar, 2 * KT_B_TILE);
                const uint32_t dst = smem_u32(Bs + s * 2 * KT_B_TILE);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int row0 = (h * NC + c) * KT_W;
                    tma_load_2d(dst + h * KT_B_TILE, &ws_map, bar, 0, row0);
                    tma_load_2d(dst + h * KT_B_TILE + KT_B_KH, &ws_map, bar, 32, row0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(smem_u32(aready), 0);
            const uint32_t ah = smem_u32(As), al = ah + KT_A_TILE;
            for (int c = 0; c < NC; ++c) {
                const int s = c % KT_STAGES, x = c % KZ_NACC;
                mbar_wait(smem_u32(bfull + s), (c / KT_STAGES) & 1);
                mbar_wait(smem_u32(accfree + x), ((c / KZ_NACC) & 1) ^ 1);
                tc_fence_after();
                const uint32_t bh = smem_u32(Bs + s * 2 * KT_B_TILE), bl = bh + KT_B_TILE;
                const uint32_t d = tmem_acc + (uint32_t)(x * KT_W);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t ao = (uint32_t)((kk >> 2) * TC_KH_BYTES + (kk & 3) * 32);
                    const uint32_t bo = (uint32_t)((kk >> 2) * KT_B_KH + (kk & 3) * 32);
                    umma_tf32_ss64(d, umma_desc_sw128(al + ao), umma_desc_sw128(bh + bo), kk != 0);
                    umma_tf32_ss64(d, umma_desc_sw128(ah + ao), umma_desc_sw128(bl + bo), 1);
                    umma_tf32_ss64(d, umma_desc_sw128(ah + ao), umma_desc_sw128(bh + bo), 1);
                }
                umma_commit(smem_u32(sfree + s));
                umma_commit(smem_u32(accfull + x));
            }
        }
    } else {
        const int et = tid - 64, ew = warp - 2;                  // 512 epilogue threads, 16 warps
        // ---- the tile's g_pre: to global (T3 reads it), to the operand tiles (hi | lo), column sums for g_bf ----
        {
            float4 colsum = make_float4(0.f, 0.f, 0.f, 0.f);     // this thread's 4 output columns over its 4 nodes
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int idx = et + 512 * it, node = idx >> 4, c16 = idx & 15;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n0 + node < n) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(g_out + (n0 + node) * KT_W) + c16);
                    const float4 o = __ldg(reinterpret_cast<const float4*>(out + (n0 + node) * KT_W) + c16);
                    v.x = o.x > 0.f ? g.x : 0.f;
                    v.y = o.y > 0.f ? g.y : 0.f;
                    v.z = o.z > 0.f ? g.z : 0.f;
                    v.w = o.w > 0.f ? g.w : 0.f;
                    *reinterpret_cast<float4*>(g_pre + (n0 + node) * KT_W + 4 * c16) = v;
                }
                colsum.x += v.x; colsum.y += v.y; colsum.z += v.z; colsum.w += v.w;
                const float4 h = make_float4(kt_hi(v.x), kt_hi(v.y), kt_hi(v.z), kt_hi(v.w));
                const uint32_t off = sw128_off(node, c16);
                *reinterpret_cast<float4*>(As + off) = h;
                *reinterpret_cast<float4*>(As + KT_A_TILE + off) =
                    make_float4(kt_lo(v.x, h.x), kt_lo(v.y, h.y), kt_lo(v.z, h.z), kt_lo(v.w, h.w));
            }
            // lanes l and l + 16 hold the same columns (c16 = et & 15): fold them, one slot per (warp, column)
            colsum.x += __shfl_xor_sync(0xffffffffu, colsum.x, 16);
            colsum.y += __shfl_xor_sync(0xffffffffu, colsum.y, 16);
            colsum.z += __shfl_xor_sync(0xffffffffu, colsum.z, 16);
            colsum.w += __shfl_xor_sync(0xffffffffu, colsum.w, 16);
            if (lane < 16) *reinterpret_cast<float4*>(BF + ew * KT_W + 4 * lane) = colsum;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(aready));
        }
        // ---- per chunk: thread = (node row, 16 feature dims) ----
        const int q = warp & 3, cq = ew >> 2;
        const int node_l = q * 32 + lane;
        const int64_t node = n0 + node_l;
        const bool valid = node < n;
        float zz[3][16], gz[3][16];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) v = __ldg(reinterpret_cast<const float4*>(z + (node * 3 + r) * KT_W + 16 * cq) + j4);
                zz[r][4 * j4 + 0] = v.x; zz[r][4 * j4 + 1] = v.y; zz[r][4 * j4 + 2] = v.z; zz[r][4 * j4 + 3] = v.w;
                gz[r][4 * j4 + 0] = gz[r][4 * j4 + 1] = gz[r][4 * j4 + 2] = gz[r][4 * j4 + 3] = 0.f;
            }
        float* gwb_w = GWB + ew * c3;
        for (int c = 0; c < NC; ++c) {
            const int x = c % KZ_NACC;
            mbar_wait(smem_u32(accfull + x), (c / KZ_NACC) & 1);
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(x * KT_W + 16 * cq), v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(accfree + x));
            if (c < C) {
                const float w0 = WB[3 * c], w1 = WB[3 * c + 1], w2 = WB[3 * c + 2];
                float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    // the SAME expression as bit_pre() of tgcn_tail.cu / the forward's generator
                    const float pre = fmaf(w2, zz[2][j], fmaf(w1, zz[1][j], w0 * zz[0][j]));
                    const float g = pre > 0.f ? __uint_as_float(v[j]) : 0.f;
                    gz[0][j] = fmaf(w0, g, gz[0][j]);
                    gz[1][j] = fmaf(w1, g, gz[1][j]);
                    gz[2][j] = fmaf(w2, g, gz[2][j]);
                    t0 = fmaf(g, zz[0][j], t0);
                    t1 = fmaf(g, zz[1][j], t1);
                    t2 = fmaf(g, zz[2][j], t2);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    t0 += __shfl_xor_sync(0xffffffffu, t0, o);
                    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
                    t2 += __shfl_xor_sync(0xffffffffu, t2, o);
                }
                if (lane == 0) {
                    gwb_w[3 * c] = t0;
                    gwb_w[3 * c + 1] = t1;
                    gwb_w[3 * c + 2] = t2;
                }
            } else if (valid && 16 * cq < E) {       // the vector-level chunk: its gradient goes back to the caller
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4)
                    if (16 * cq + 4 * j4 < E)
                        *reinterpret_cast<float4*>(g_xf + node * E + 16 * cq + 4 * j4) =
                            make_float4(__uint_as_float(v[4 * j4]), __uint_as_float(v[4 * j4 + 1]),
                                        __uint_as_float(v[4 * j4 + 2]), __uint_as_float(v[4 * j4 + 3]));
            }
        }
        if (valid) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4)
                    *reinterpret_cast<float4*>(g_z + (node * 3 + r) * KT_W + 16 * cq + 4 * j4) =
                        make_float4(gz[r][4 * j4], gz[r][4 * j4 + 1], gz[r][4 * j4 + 2], gz[r][4 * j4 + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    // per-CTA sums -> one atomic per parameter element
    for (int i = tid; i < 3 * C; i += KZ_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) s += GWB[w * c3 + i];
        if (s != 0.f) atomicAdd(g_wb + i, s);
    }
    if (tid < KT_W) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) s += BF[w * KT_W + tid];
        if (s != 0.f) atomicAdd(g_bf + tid, s);
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(256u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ T3 on tcgen05
// Weight gradient of the fusion layer: g_Wf[c*64 + d][o] = sum_node feature(c, d)[node] * g_pre[node][o] — a product
// whose K dimension is the NODE index, and TF32 operands must be K-major.  So z, xf and g_pre are transposed once per
// call (node-contiguous rows; g_pre also split hi | lo there), the g_pre^T tiles arrive by TMA, and the generators
// write features of TWO conv channels (M = 128 rows) for 64 nodes per stage straight into the swizzled A tiles.
// The tensor core's fp32 accumulation rounds toward zero: after every KW_FLUSH stages (96 MMAs) the accumulator is
// drained into fp32 registers of four epilogue warps (thread = feature row, 64 output columns).
constexpr int KW_FLUSH = 4;
constexpr int KW_NACC = 2;
constexpr int KW_THREADS = 64 + 256 + 128;
constexpr int KW_ST = 64;                       // nodes per stage (K = 64)

// zT[r*64 + d][node] = z[node][r][d];  xT[f][node] = xf[node][f] (rows >= E: 0);  gT[h][o][node] = hi / lo of g_pre
// Classic 32 x 32 tile transpose: grid (node tiles, 10 column tiles: 6 of z, 2 of xf, 2 of g_pre), 128-byte segments
// on both sides.
__global__ void __launch_bounds__(256)
tgcn_tail_transpose_nodes_kernel(const float* __restrict__ z, const float* __restrict__ xf, const float* __restrict__ g_pre,
                                 int64_t n, int64_t np, int E, float* __restrict__ zT, float* __restrict__ xT,
                                 float* __restrict__ gT) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t n0 = (int64_t)blockIdx.x * 32;
    const int ct = blockIdx.y;                                  // column tile
    const float* src;
    int width, col0;
    if (ct < 6) { src = z; width = 192; col0 = 32 * ct; }
    else if (ct < 8) { src = xf; width = E; col0 = 32 * (ct - 6); }
    else { src = g_pre; width = 64; col0 = 32 * (ct - 8); }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int node = ty + 8 * k;
        tile[node][tx] = (n0 + node < n && col0 + tx < width) ? __ldg(src + (n0 + node) * width + col0 + tx) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = ty + 8 * k;                               // column of this tile -> output row
        const float v = tile[tx][c];
        const size_t o = (size_t)(col0 + c) * np + n0 + tx;
        if (ct < 6) zT[o] = v;
        else if (ct < 8) xT[o] = v;
        else {
            const float h = kt_hi(v);
            gT[o] = h;
            gT[o + (size_t)64 * np] = kt_lo(v, h);
        }
    }
}

__global__ void __launch_bounds__(KW_THREADS, 1)
tgcn_tail_bwd_w_tc_kernel(const __grid_constant__ CUtensorMap gt_map, const float* __restrict__ zT,
                          const float* __restrict__ xT, const float* __restrict__ wb, int64_t np, int C, int E,
                          int64_t stages_per_split, float* __restrict__ g_wf) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* As = base;                                   // [S] x (hi 32 KB | lo 32 KB): 128 feature rows x 64 nodes
    unsigned char* Bs = As + KT_STAGES * 2 * KT_A_TILE;         // [S] x (hi 16 KB | lo 16 KB): 64 output rows x 64 nodes
    uint64_t* bars = reinterpret_cast<uint64_t*>(Bs + KT_STAGES * 2 * KT_B_TILE);
    uint64_t* afull = bars;                     // [S]    generators -> MMA
    uint64_t* bfull = afull + KT_STAGES;        // [S]    TMA -> MMA
    uint64_t* sfree = bfull + KT_STAGES;        // [S]    MMA (commit) -> generators, TMA
    uint64_t* accfull = sfree + KT_STAGES;      // [NACC] MMA -> epilogue
    uint64_t* accfree = accfull + KW_NACC;      // [NACC] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + KW_NACC);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int c0 = blockIdx.x * 2;
    const int64_t n_stages = np / KW_ST;
    const int64_t st_begin = (int64_t)blockIdx.y * stages_per_split;
    const int64_t st_end = min(n_stages, st_begin + stages_per_split);
    const int n_st = (int)max((int64_t)0, st_end - st_begin);
    const int n_groups = (n_st + KW_FLUSH - 1) / KW_FLUSH;

    if (tid == 0) {
        for (int s = 0; s < KT_STAGES; ++s) {
            mbar_init(smem_u32(afull + s), 8);
            mbar_init(smem_u32(bfull + s), 1);
            mbar_init(smem_u32(sfree + s), 1);
        }
        for (int x = 0; x < KW_NACC; ++x) {
            mbar_init(smem_u32(accfull + x), 1);
            mbar_init(smem_u32(accfree + x), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // ================= TMA: g_pre^T tiles, hi rows 0..63, lo rows 64..127 =================
        if (lane == 0) {
            for (int i = 0; i < n_st; ++i) {
                const int s = i % KT_STAGES;
                mbar_wait(smem_u32(sfree + s), ((i / KT_STAGES) & 1) ^ 1);
                const uint32_t bar = smem_u32(bfull + s);
                mbar_expect_tx(bar, 2 * KT_B_TILE);
                const uint32_t dst = smem_u32(Bs + s * 2 * KT_B_TILE);
                const int col0 = (int)((st_begin + i) * KW_ST);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    tma_load_2d(dst + h * KT_B_TILE, &gt_map, bar, col0, h * 64);
                    tma_load_2d(dst + h * KT_B_TILE + KT_B_KH, &gt_map, bar, col0 + 32, h * 64);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            for (int i = 0; i < n_st; ++i) {
                const int s = i % KT_STAGES, g = i / KW_FLUSH, x = g % KW_NACC;
                const bool first = (i % KW_FLUSH) == 0;
                mbar_wait(smem_u32(afull + s), (i / KT_STAGES) & 1);
                mbar_wait(smem_u32(bfull + s), (i / KT_STAGES) & 1);
                if (first) mbar_wait(smem_u32(accfree + x), ((g / KW_NACC) & 1) ^ 1);
                tc_fence_after();
                const uint32_t ah = smem_u32(As + s * 2 * KT_A_TILE), al = ah + KT_A_TILE;
                const uint32_t bh = smem_u32(Bs + s * 2 * KT_B_TILE), bl = bh + KT_B_TILE;
                const uint32_t d = tmem_acc + (uint32_t)(x * KT_W);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t ao = (uint32_t)((kk >> 2) * TC_KH_BYTES + (kk & 3) * 32);
                    const uint32_t bo = (uint32_t)((kk >> 2) * KT_B_KH + (kk & 3) * 32);
                    umma_tf32_ss64(d, umma_desc_sw128(al + ao), umma_desc_sw128(bh + bo), !(first && kk == 0));
                    umma_tf32_ss64(d, umma_desc_sw128(ah + ao), umma_desc_sw128(bl + bo), 1);
                    umma_tf32_ss64(d, umma_desc_sw128(ah + ao), umma_desc_sw128(bh + bo), 1);
                }
                umma_commit(smem_u32(sfree + s));
                if ((i % KW_FLUSH) == KW_FLUSH - 1 || i == n_st - 1) umma_commit(smem_u32(accfull + x));
            }
        }
    } else if (warp < 10) {
        // ================= generators: thread = 4 x (feature dim d, four consecutive nodes), both channels =================
        const int gt = tid - 64;
        float w[2][3];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int r = 0; r < 3; ++r) w[cc][r] = (c0 + cc < C) ? __ldg(wb + 3 * (c0 + cc) + r) : 0.f;
        for (int i = 0; i < n_st; ++i) {
            const int s = i % KT_STAGES;
            const int64_t node0 = (st_begin + i) * KW_ST;
            float4 z0[4], z1[4], z2[4], xe[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {                    // loads first, then the wait for the stage
                const int idx = gt + 256 * it, n4 = idx & 15, d = idx >> 4;
                const float* zp = zT + (size_t)d * np + node0 + 4 * n4;
                z0[it] = __ldg(reinterpret_cast<const float4*>(zp));
                z1[it] = __ldg(reinterpret_cast<const float4*>(zp + (size_t)64 * np));
                z2[it] = __ldg(reinterpret_cast<const float4*>(zp + (size_t)128 * np));
                xe[it] = (c0 <= C && C < c0 + 2) ? __ldg(reinterpret_cast<const float4*>(xT + (size_t)d * np + node0 + 4 * n4))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_wait(smem_u32(sfree + s), ((i / KT_STAGES) & 1) ^ 1);
            unsigned char* hi = As + s * 2 * KT_A_TILE;
            unsigned char* lo = hi + KT_A_TILE;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int idx = gt + 256 * it, n4 = idx & 15, d = idx >> 4;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = c0 + cc;
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c < C) {        // the SAME expression as bit_pre() of tgcn_tail.cu
                        f.x = fmaxf(fmaf(w[cc][2], z2[it].x, fmaf(w[cc][1], z1[it].x, w[cc][0] * z0[it].x)), 0.f);
                        f.y = fmaxf(fmaf(w[cc][2], z2[it].y, fmaf(w[cc][1], z1[it].y, w[cc][0] * z0[it].y)), 0.f);
                        f.z = fmaxf(fmaf(w[cc][2], z2[it].z, fmaf(w[cc][1], z1[it].z, w[cc][0] * z0[it].z)), 0.f);
                        f.w = fmaxf(fmaf(w[cc][2], z2[it].w, fmaf(w[cc][1], z1[it].w, w[cc][0] * z0[it].w)), 0.f);
                    } else if (c == C) {
                        f = xe[it];     // vector-level features (rows >= E of xT are zero)
                    }
                    const float4 h = make_float4(kt_hi(f.x), kt_hi(f.y), kt_hi(f.z), kt_hi(f.w));
                    const uint32_t off = sw128_off(cc * 64 + d, n4);
                    *reinterpret_cast<float4*>(hi + off) = h;
                    *reinterpret_cast<float4*>(lo + off) =
                        make_float4(kt_lo(f.x, h.x), kt_lo(f.y, h.y), kt_lo(f.z, h.z), kt_lo(f.w, h.w));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(afull + s));
        }
    } else {
        // ================= epilogue: thread = feature row of the channel pair, 64 output columns =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        float sum[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) sum[j] = 0.f;
        for (int g = 0; g < n_groups; ++g) {
            const int x = g % KW_NACC;
            mbar_wait(smem_u32(accfull + x), (g / KW_NACC) & 1);
            tc_fence_after();
            uint32_t v0[32], v1[32];
            const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(x * KT_W);
            tmem_ld32(taddr, v0);
            tmem_ld32(taddr + 32, v1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(accfree + x));
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                sum[j] += __uint_as_float(v0[j]);
                sum[32 + j] += __uint_as_float(v1[j]);
            }
        }
        const int grow = (c0 + (row >> 6)) * KT_W + (row & 63);
        if (grow < C * KT_W + E) {
#pragma unroll
            for (int j4 = 0; j4 < 16; ++j4)
                red_add4(reinterpret_cast<float4*>(g_wf + (size_t)grow * KT_W + 4 * j4),
                         make_float4(sum[4 * j4], sum[4 * j4 + 1], sum[4 * j4 + 2], sum[4 * j4 + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(128u) : "memory");
    }
}

size_t tail_tc_workspace_bytes(int C) { return (size_t)2 * (C + 1) * KT_W * KT_W * 4 + 256; }

bool tail_tc_available() { return encode_tiled() != nullptr; }

int tail_fwd_tc(const float* z, const float* wb, const float* xf, const float* wf, const float* bf, int64_t n, int C,
                int E, float* out, void* workspace, void* stream) {
    const int NC = C + (E > 0 ? 1 : 0);
    float* wt = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    TAGREC_LAUNCH(tgcn_tail_split_kernel, (unsigned)((NC * KT_W * KT_W + 255) / 256), 256, 0, stream, wf, C, E, wt);
    CUtensorMap map;
    if (int rc = make_row_table_map(&map, wt, (int64_t)2 * NC * KT_W, KT_W, KT_W)) return rc;
    const size_t smem = 1024 + (size_t)KT_STAGES * 2 * (KT_A_TILE + KT_B_TILE) + ((3 * C + 3) & ~3) * 4 +
                        (3 * KT_STAGES + 1) * 8 + 64;
    TAGREC_CUDA(cudaFuncSetAttribute(tgcn_tail_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TAGREC_LAUNCH(tgcn_tail_fwd_tc_kernel, (unsigned)((n + KT_M - 1) / KT_M), KT_THREADS, smem, stream, map, z, wb, xf, bf, n,
                  C, E, out);
    return TAGREC_OK;
}

size_t tail_tc_bwd_workspace_bytes(int C) { return (size_t)2 * (C + 1) * KT_W * KT_W * 4 + 256; }

// T2 on the tensor cores; g_wb / g_bf are accumulated into (zeroed by the caller)
int tail_bwd_z_tc(const float* g_out, const float* out, const float* z, const float* wb, const float* wf, int64_t n,
                  int C, int E, void* workspace, float* g_pre, float* g_z, float* g_wb, float* g_xf, float* g_bf,
                  void* stream) {
    const int NC = C + (E > 0 ? 1 : 0);
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    TAGREC_LAUNCH(tgcn_tail_split_rows_kernel, (unsigned)((NC * KT_W * KT_W + 255) / 256), 256, 0, stream, wf, C, E, ws);
    CUtensorMap map;
    if (int rc = make_row_table_map(&map, ws, (int64_t)2 * NC * KT_W, KT_W, KT_W)) return rc;
    const int c3 = (3 * C + 3) & ~3;
    const size_t smem = 1024 + (size_t)2 * KT_A_TILE + KT_STAGES * 2 * KT_B_TILE + ((size_t)c3 * 17 + 16 * KT_W) * 4 +
                        (2 * KT_STAGES + 2 * KZ_NACC + 1) * 8 + 64;
    TAGREC_REQUIRE(smem <= 227 * 1024, "too many bit-level conv channels for the tensor-core backward");
    TAGREC_CUDA(cudaFuncSetAttribute(tgcn_tail_bwd_z_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TAGREC_LAUNCH(tgcn_tail_bwd_z_tc_kernel, (unsigned)((n + KT_M - 1) / KT_M), KZ_THREADS, smem, stream, map, g_out, out, z,
                  wb, n, C, E, g_pre, g_z, g_wb, g_xf, g_bf);
    return TAGREC_OK;
}

static int64_t tail_tc_np(int64_t n) { return ((n + 127) / 128) * 128; }
size_t tail_tc_bwd_w_workspace_bytes(int64_t n) { return (size_t)tail_tc_np(n) * (192 + 64 + 128) * 4 + 256; }

// T3 on the tensor cores; g_wf is accumulated into (zeroed by the caller)
int tail_bwd_w_tc(const float* z, const float* wb, const float* xf, const float* g_pre, int64_t n, int C, int E,
                  void* workspace, float* g_wf, void* stream) {
    const int NC = C + (E > 0 ? 1 : 0);
    const int64_t np = tail_tc_np(n);
    float* zT = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    float* xT = zT + (size_t)192 * np;
    float* gT = xT + (size_t)64 * np;
    TAGREC_LAUNCH(tgcn_tail_transpose_nodes_kernel, dim3((unsigned)(np / 32), 10), 256, 0, stream, z, xf, g_pre, n, np, E,
                  zT, xT, gT);
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail(TAGREC_ECUDA, "cuTensorMapEncodeTiled not available from the driver", __FILE__, __LINE__);
    CUtensorMap map;                                            // [128 rows (hi | lo)][np nodes], boxes of 32 nodes x 64 rows
    const cuuint64_t gdim[2] = {(cuuint64_t)np, 128};
    const cuuint64_t gstride[1] = {(cuuint64_t)np * 4};
    const cuuint32_t box[2] = {32, 64};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, gT, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return fail(TAGREC_ECUDA, "cuTensorMapEncodeTiled failed", __FILE__, __LINE__);
    const int pairs = (NC + 1) / 2;
    const int64_t n_stages = np / KW_ST;
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>(n_stages, kSMs / pairs));
    const int64_t sps = (n_stages + splits - 1) / splits;
    splits = (n_stages + sps - 1) / sps;
    const size_t smem = 1024 + (size_t)KT_STAGES * 2 * (KT_A_TILE + KT_B_TILE) + (3 * KT_STAGES + 2 * KW_NACC) * 8 + 64;
    TAGREC_CUDA(cudaFuncSetAttribute(tgcn_tail_bwd_w_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TAGREC_LAUNCH(tgcn_tail_bwd_w_tc_kernel, dim3((unsigned)pairs, (unsigned)splits), KW_THREADS, smem, stream, map, zT, xT, wb,
                  np, C, E, sps, g_wf);
    return TAGREC_OK;
}

}  // namespace tagrec
