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

}  // namespace tagrec
