// K3b — per-user AUC of the full-sort evaluation on the device.  sm_100a.
//
// Replaces training/utils.py:37-45 (auc: sklearn.metrics.roc_auc_score over the un-masked scores of ONE user, called
// once per user with a full score row copied to the host — 85 % of the reference's evaluation time) and
// the per-user loop at training/basic_test.py:52-53.
//
// For user u let P = test items of u that are not train items (the positives), m = |P|, and N = all other un-masked
// items.  roc_auc_score == the Mann-Whitney statistic with half credit for ties:
//     AUC_u = sum_{i in N} f_u(s_ui) / (m |N|),      f_u(s) = #{p in P : s_up > s} + 1/2 #{p in P : s_up == s}.
// Instead of testing membership for every (user, item) pair, the dense pass sums f over ALL items and two small
// sparse passes subtract the train items and the positives themselves:
//     sum_{i in N} f = sum_{all i} f - sum_{i in train(u)} f - sum_{p in P} f.
// Scores are the exact fp32 dot products in the canonical sequential-fmaf order of eval_topk.cu in all three passes,
// so the subtraction cancels bit-for-bit; sums are kept as integers (2 f) and are therefore exact.
// Ranking quantity: the raw dot product (the reference ranks fp32 sigmoid(dot); sigmoid is monotone, it only merges
// dots closer than one ulp of the sigmoid into ties — a <= 1e-7 effect on AUC, documented in DESIGN.md).
//
//   A1 auc_pos_kernel       positives' scores per user (train members -> +inf = "not a positive"), rank-sorted
//   A2 auc_all_kernel       64-user x 128-item fp32 tiles (same tiling as eval_topk_kernel); per score one or two
//                           register compares against the row's [min, max] positive, else a branch-free bisection
//                           over the row's sorted positives staged in shared memory (a user with more than 32
//                           positives is several virtual rows, one group of 32 each: 2 f is additive).  The search, not the scoring, bounds this pass when
//                           positives are spread over the whole score range (a 128 x 128 / 8 x 8 variant of the tile
//                           measured SLOWER for that reason: fewer warps to hide the search latency).
//   A3 auc_finalize_kernel  subtract train / positive contributions, divide, accumulate sum and user count
#include <float.h>

#include <algorithm>

#include "common.cuh"

// packed FFMA2 in the tile loop: +5 % (dim 64) .. +10 % (dim 256) measured against the scalar form, same sums
#ifndef TAGREC_EVAL_PACKED
#define TAGREC_EVAL_PACKED 1
#endif

namespace tagrec {

constexpr int AUT = 64;       // users per block
constexpr int AIT = 128;      // items per tile
constexpr int AKC = 32;       // feature chunk
constexpr int APC = 32;       // positives per row staged in shared memory

__device__ __forceinline__ float dot_seq(const float* __restrict__ a, const float* __restrict__ b, int dim) {
    float acc = 0.f;
    for (int k = 0; k < dim; k += 4) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a + k));
        const float4 y = __ldg(reinterpret_cast<const float4*>(b + k));
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
        acc = fmaf(x.z, y.z, acc);
        acc = fmaf(x.w, y.w, acc);
    }
    return acc;
}

__device__ __forceinline__ bool in_sorted(const int32_t* a, int64_t lo, int64_t hi, int32_t x) {
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(a + lo) == x;
}

// 2 * f(s) for the ascending positives p[0..m): 2 * #{p > s} + #{p == s}
__device__ __forceinline__ unsigned long long twice_f(const float* __restrict__ p, int m, float s) {
    int lo = 0, hi = m;                       // lower bound: first index with p >= s
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(p + mid) < s) lo = mid + 1; else hi = mid;
    }
    int eq = 0;
    while (lo + eq < m && __ldg(p + lo + eq) == s) ++eq;
    return 2ull * (unsigned long long)(m - lo - eq) + (unsigned long long)eq;
}

// A1: one warp per user.  pos_raw / pos_sorted are indexed like test_items (CSR offsets of the WHOLE test set).
__global__ void __launch_bounds__(256)
auc_pos_kernel(const int64_t* __restrict__ users, int64_t nu, const float* __restrict__ ut, const float* __restrict__ it,
               int dim, const int64_t* __restrict__ train_ptr, const int32_t* __restrict__ train_items,
               const int64_t* __restrict__ test_ptr, const int32_t* __restrict__ test_items, float* __restrict__ pos_raw,
               float* __restrict__ pos_sorted, int32_t* __restrict__ n_pos) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nu) return;
    const int64_t u = users[w];
    const int64_t tb = test_ptr[u], te = test_ptr[u + 1];
    const int64_t rb = train_ptr[u], re = train_ptr[u + 1];
    int cnt = 0;
    for (int64_t j = tb + lane; j < te; j += 32) {
        const int32_t item = test_items[j];
        float s = INFINITY;
        if (!in_sorted(train_items, rb, re, item)) {
            s = dot_seq(ut + u * dim, it + (int64_t)item * dim, dim);
            ++cnt;
        }
        pos_raw[j] = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) n_pos[w] = cnt;
    __syncwarp();
    const int m = (int)(te - tb);
    for (int i = lane; i < m; i += 32) {          // rank sort (m is a user's test-set size: tens, rarely thousands)
        const float si = pos_raw[tb + i];
        int rank = 0;
        for (int j = 0; j < m; ++j) {
            const float sj = pos_raw[tb + j];
            rank += (sj < si) || (sj == si && j < i);
        }
        pos_sorted[tb + rank] = si;
    }
}

struct AucArgs {
    const int64_t* users;
    int64_t nu;
    const float* user_table;
    const float* item_table;
    int64_t n_item;
    int dim;
    const int64_t* test_ptr;
    const float* pos_sorted;
    const int32_t* n_pos;
    int64_t items_per_split;
    unsigned long long* acc2;     // [nu] sum over all items of 2 f
    // virtual rows (eval_auc_tc.cu, auc_vrows_kernel): a user with m positives is max(1, ceil(m / APC)) rows, each with
    // one group of APC sorted positives — 2 f is additive over the groups, and every row's search runs in smem
    const int32_t* vr_owner;
    const int32_t* vr_part;
    const int32_t* nv;
};

// A2
__global__ void __launch_bounds__(256) auc_all_kernel(AucArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = a.dim;
    float* Us = reinterpret_cast<float*>(smem_raw);              // [D][AUT+4] user tile, feature-major
    float* Is = Us + (size_t)D * (AUT + 4);                       // [AKC][AIT+4] item chunk
    float* Ps = Is + (size_t)AKC * (AIT + 4);                     // [AUT][APC]   sorted positives (+inf padded)
    int64_t* uid = reinterpret_cast<int64_t*>(Ps + (size_t)AUT * APC);         // [AUT] user id of the row, -1 = none
    int64_t* poff = uid + AUT;                                                 // [AUT] offset of the row's positives
    int32_t* own = reinterpret_cast<int32_t*>(poff + AUT);                     // [AUT] index into users[] / acc2[]
    int32_t* pcnt = own + AUT;                                                 // [AUT] positives of this (virtual) row
    const int tid = threadIdx.x;
    const int tu = tid >> 4, ti = tid & 15;
    const int64_t u0 = (int64_t)blockIdx.x * AUT;
    const int64_t i_begin = (int64_t)blockIdx.y * a.items_per_split;
    const int64_t i_end = min(a.n_item, i_begin + a.items_per_split);
    const int64_t nv = __ldg(a.nv);
    if (u0 >= nv) return;                       // the grid is sized for the upper bound of the virtual-row count
    if (tid < AUT) {
        uid[tid] = -1;
        own[tid] = 0;
        pcnt[tid] = 0;
        poff[tid] = 0;
        if (u0 + tid < nv) {
            const int w = __ldg(a.vr_owner + u0 + tid), part = __ldg(a.vr_part + u0 + tid);
            const int64_t u = a.users[w];
            uid[tid] = u;
            own[tid] = w;
            pcnt[tid] = min(APC, max(0, a.n_pos[w] - part * APC));
            poff[tid] = a.test_ptr[u] + (int64_t)part * APC;
        }
    }
    __syncthreads();
    for (int idx = tid; idx < AUT * APC; idx += 256) {
        const int u = idx / APC, j = idx % APC;
        Ps[idx] = j < pcnt[u] ? __ldg(a.pos_sorted + poff[u] + j) : INFINITY;
    }
    for (int idx = tid; idx < AUT * (D / 4); idx += 256) {
        const int u = idx / (D / 4), c4 = idx % (D / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (uid[u] >= 0) v = __ldg(reinterpret_cast<const float4*>(a.user_table + uid[u] * D) + c4);
        Us[(4 * c4 + 0) * (AUT + 4) + u] = v.x;
        Us[(4 * c4 + 1) * (AUT + 4) + u] = v.y;
        Us[(4 * c4 + 2) * (AUT + 4) + u] = v.z;
        Us[(4 * c4 + 3) * (AUT + 4) + u] = v.w;
    }
    // per-row positives: pointer, count, min, max
    const float* pp[4];
    int pm[4];
    float pmin[4], pmax[4];
    unsigned long long tot[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int u = 4 * tu + r;
        pm[r] = 0;
        pp[r] = a.pos_sorted;
        pmin[r] = INFINITY;
        pmax[r] = -INFINITY;
        if (uid[u] >= 0) {
            pm[r] = pcnt[u];
            pp[r] = a.pos_sorted + poff[u];
            if (pm[r] > 0) {
                pmin[r] = __ldg(pp[r]);
                pmax[r] = __ldg(pp[r] + pm[r] - 1);
            }
        }
    }
    for (int64_t it0 = i_begin; it0 < i_end; it0 += AIT) {
        unsigned long long accp[4][4];              // accp[r][cp] = scores (r, 2 cp) and (r, 2 cp + 1), packed
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) accp[r][cp] = 0ull;
        for (int k0 = 0; k0 < D; k0 += AKC) {
            __syncthreads();
            for (int idx = tid; idx < AIT * (AKC / 4); idx += 256) {
                const int i = idx / (AKC / 4), c4 = idx % (AKC / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it0 + i < i_end) v = __ldg(reinterpret_cast<const float4*>(a.item_table + (it0 + i) * D + k0) + c4);
                Is[(4 * c4 + 0) * (AIT + 4) + i] = v.x;
                Is[(4 * c4 + 1) * (AIT + 4) + i] = v.y;
                Is[(4 * c4 + 2) * (AIT + 4) + i] = v.z;
                Is[(4 * c4 + 3) * (AIT + 4) + i] = v.w;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < AKC; ++kk) {
                const float4 uu = *reinterpret_cast<const float4*>(Us + (size_t)(k0 + kk) * (AUT + 4) + 4 * tu);
                const float4 i0 = *reinterpret_cast<const float4*>(Is + (size_t)kk * (AIT + 4) + 4 * ti);
                const float4 i1 = *reinterpret_cast<const float4*>(Is + (size_t)kk * (AIT + 4) + 64 + 4 * ti);
                // packed FMAs (FFMA2, common.cuh): two scores per instruction, each score's own sequential chain over
                // k unchanged — the canonical order of the exact re-scores
#if TAGREC_EVAL_PACKED
                const unsigned long long ud[4] = {pack2(uu.x, uu.x), pack2(uu.y, uu.y), pack2(uu.z, uu.z), pack2(uu.w, uu.w)};
                const unsigned long long ip[4] = {pack2(i0.x, i0.y), pack2(i0.z, i0.w), pack2(i1.x, i1.y), pack2(i1.z, i1.w)};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cp = 0; cp < 4; ++cp) accp[r][cp] = ffma2(ud[r], ip[cp], accp[r][cp]);
#else
                const float uv[4] = {uu.x, uu.y, uu.z, uu.w};
                const float iv[8] = {i0.x, i0.y, i0.z, i0.w, i1.x, i1.y, i1.z, i1.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int cp = 0; cp < 4; ++cp) {
                        float lo, hi;
                        unpack2(accp[r][cp], lo, hi);
                        accp[r][cp] = pack2(fmaf(uv[r], iv[2 * cp], lo), fmaf(uv[r], iv[2 * cp + 1], hi));
                    }
#endif
            }
        }
        float acc[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) unpack2(accp[r][cp], acc[r][2 * cp], acc[r][2 * cp + 1]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (pm[r] == 0) continue;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int64_t item = it0 + (c < 4 ? 4 * ti + c : 64 + 4 * ti + (c - 4));
                if (item >= i_end) continue;
                const float s = acc[r][c];
                if (s < pmin[r]) tot[r] += 2ull * (unsigned long long)pm[r];      // below every positive
                else if (s > pmax[r]) {}                                         // above every positive
                else {
                    // branch-free bisection in shared memory: lb = #{p < s}, then the (rare) run of equal positives
                    const float* ps = Ps + (size_t)(4 * tu + r) * APC;
                    int lb = 0;
#pragma unroll
                    for (int st = APC / 2; st > 0; st >>= 1) lb += (ps[lb + st - 1] < s) ? st : 0;
                    lb += (ps[lb] < s) ? 1 : 0;
                    int eq = 0;
                    while (lb + eq < pm[r] && ps[lb + eq] == s) ++eq;
                    tot[r] += 2ull * (unsigned long long)(pm[r] - lb - eq) + (unsigned long long)eq;
                }
            }
        }
    }
    // the 16 threads (ti) of a half-warp share the rows 4*tu .. 4*tu+3
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        unsigned long long v = tot[r];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (ti == 0 && uid[4 * tu + r] >= 0 && v) atomicAdd(a.acc2 + own[4 * tu + r], v);
    }
}

// A3: one warp per user
__global__ void __launch_bounds__(256)
auc_finalize_kernel(const int64_t* __restrict__ users, int64_t nu, const float* __restrict__ ut,
                    const float* __restrict__ it, int64_t n_item, int dim, const int64_t* __restrict__ train_ptr,
                    const int32_t* __restrict__ train_items, const int64_t* __restrict__ test_ptr,
                    const float* __restrict__ pos_sorted, const int32_t* __restrict__ n_pos,
                    const unsigned long long* __restrict__ acc2, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nu) return;
    const int64_t u = users[w];
    const int m = n_pos[w];
    const int64_t rb = train_ptr[u], re = train_ptr[u + 1];
    const int64_t n_neg = n_item - (re - rb) - m;
    if (m == 0 || n_neg <= 0) return;          // one class only: roc_auc_score is undefined (the reference raises)
    const float* p = pos_sorted + test_ptr[u];
    unsigned long long sub = 0;
    for (int64_t j = rb + lane; j < re; j += 32)
        sub += twice_f(p, m, dot_seq(ut + u * dim, it + (int64_t)__ldg(train_items + j) * dim, dim));
    for (int j = lane; j < m; j += 32) sub += twice_f(p, m, __ldg(p + j));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sub += __shfl_xor_sync(0xffffffffu, sub, o);
    if (lane == 0) {
        const double num = 0.5 * (double)(acc2[w] - sub);
        atomicAdd(out, num / ((double)m * (double)n_neg));
        atomicAdd(out + 1, 1.0);
    }
}

static int auc_splits(int64_t nu, int64_t n_item) {
    const int64_t user_tiles = (nu + AUT - 1) / AUT;
    int64_t s = (2 * kSMs + user_tiles - 1) / user_tiles;
    const int64_t max_s = (n_item + 4 * AIT - 1) / (4 * AIT);
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return (int)s;
}

// eval_auc_tc.cu
bool auc_tc_available(int dim);
size_t auc_tc_workspace_bytes(int64_t nu, int64_t n_test_total);
int launch_auc_vrows(const int32_t* n_pos, int64_t nu, int64_t n_test_total, void* ws, const int32_t** vr_owner,
                     const int32_t** vr_part, const int32_t** nv, int64_t* nv_max, void* stream);
int eval_auc_tc(const int64_t* users, int64_t nu, const float* user_table, const float* item_table, int64_t n_item,
                const int64_t* test_ptr, const float* pos_sorted, const int32_t* n_pos, int64_t n_test_total, void* ws,
                unsigned long long* acc2, void* stream);

}  // namespace tagrec

using namespace tagrec;

extern "C" size_t tagrec_eval_auc_workspace_bytes(int64_t nu, int64_t n_test_total) {
    return 256 + (size_t)n_test_total * 8 + (size_t)nu * 12 + 64 + auc_tc_workspace_bytes(nu, n_test_total);
}

extern "C" int tagrec_eval_auc(const int64_t* users, int64_t nu, const float* user_table, const float* item_table,
                               int64_t n_item, int dim, const int64_t* train_ptr, const int32_t* train_items,
                               const int64_t* test_ptr, const int32_t* test_items, int64_t n_test_total,
                               void* workspace, size_t workspace_bytes, double* out, void* stream) {
    return tagrec_eval_auc_ex(users, nu, user_table, item_table, n_item, dim, train_ptr, train_items, test_ptr, test_items,
                              n_test_total, workspace, workspace_bytes, out, TAGREC_EVAL_AUTO, stream);
}

extern "C" int tagrec_eval_auc_ex(const int64_t* users, int64_t nu, const float* user_table, const float* item_table,
                                  int64_t n_item, int dim, const int64_t* train_ptr, const int32_t* train_items,
                                  const int64_t* test_ptr, const int32_t* test_items, int64_t n_test_total,
                                  void* workspace, size_t workspace_bytes, double* out, int path, void* stream) {
    TAGREC_REQUIRE(users && user_table && item_table && train_ptr && test_ptr && out, "null pointer");
    TAGREC_REQUIRE(path == TAGREC_EVAL_AUTO || path == TAGREC_EVAL_FP32 || path == TAGREC_EVAL_TF32, "bad path");
    const bool use_tc = path != TAGREC_EVAL_FP32 && auc_tc_available(dim);
    TAGREC_REQUIRE(use_tc || path != TAGREC_EVAL_TF32, "tensor-core AUC path needs dim 64");
    TAGREC_REQUIRE(dim >= 4 && dim % AKC == 0, "dim must be a multiple of 32");
    TAGREC_REQUIRE(n_item > 0 && n_item < (1ll << 31), "n_item out of range");
    if (nu == 0) return TAGREC_OK;
    if (!workspace || workspace_bytes < tagrec_eval_auc_workspace_bytes(nu, n_test_total))
        return fail(TAGREC_ENOMEM, "auc workspace too small", __FILE__, __LINE__);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    unsigned long long* acc2 = reinterpret_cast<unsigned long long*>(ws);          // [nu], 8-byte aligned
    float* pos_raw = reinterpret_cast<float*>(ws + (((size_t)nu * 8 + 255) / 256) * 256);
    float* pos_sorted = pos_raw + n_test_total;
    int32_t* n_pos = reinterpret_cast<int32_t*>(pos_sorted + n_test_total);
    cudaStream_t st = (cudaStream_t)stream;
    TAGREC_CUDA(cudaMemsetAsync(acc2, 0, (size_t)nu * 8, st));
    TAGREC_LAUNCH(auc_pos_kernel, (unsigned)((nu + 7) / 8), 256, 0, stream, users, nu, user_table, item_table, dim,
                  train_ptr, train_items, test_ptr, test_items, pos_raw, pos_sorted, n_pos);
    if (use_tc) {
        void* tc_ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(n_pos + nu) + 63) & ~(uintptr_t)63);
        if (int rc = eval_auc_tc(users, nu, user_table, item_table, n_item, test_ptr, pos_sorted, n_pos, n_test_total,
                                 tc_ws, acc2, stream))
            return rc;
        TAGREC_LAUNCH(auc_finalize_kernel, (unsigned)((nu + 7) / 8), 256, 0, stream, users, nu, user_table, item_table,
                      n_item, dim, train_ptr, train_items, test_ptr, pos_sorted, n_pos, acc2, out);
        return TAGREC_OK;
    }
    AucArgs a{};
    a.users = users; a.nu = nu; a.user_table = user_table; a.item_table = item_table; a.n_item = n_item; a.dim = dim;
    a.test_ptr = test_ptr; a.pos_sorted = pos_sorted; a.n_pos = n_pos; a.acc2 = acc2;
    const int splits = auc_splits(nu, n_item);
    const int64_t tiles = (n_item + AIT - 1) / AIT;
    a.items_per_split = ((tiles + splits - 1) / splits) * AIT;
    const size_t smem = ((size_t)dim * (AUT + 4) + (size_t)AKC * (AIT + 4) + (size_t)AUT * APC) * 4 + AUT * 24;
    TAGREC_REQUIRE(smem <= 227 * 1024, "dim too large for shared memory");
    TAGREC_CUDA(cudaFuncSetAttribute(auc_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* vr_ws = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(n_pos + nu) + 63) & ~(uintptr_t)63);
    int64_t nv_max = 0;
    if (int rc = launch_auc_vrows(n_pos, nu, n_test_total, vr_ws, &a.vr_owner, &a.vr_part, &a.nv, &nv_max, stream)) return rc;
    const dim3 grid((unsigned)((nv_max + AUT - 1) / AUT), (unsigned)splits);
    TAGREC_LAUNCH(auc_all_kernel, grid, 256, smem, stream, a);
    TAGREC_LAUNCH(auc_finalize_kernel, (unsigned)((nu + 7) / 8), 256, 0, stream, users, nu, user_table, item_table, n_item,
                  dim, train_ptr, train_items, test_ptr, pos_sorted, n_pos, acc2, out);
    return TAGREC_OK;
}
