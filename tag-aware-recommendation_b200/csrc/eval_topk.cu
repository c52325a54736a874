// K3 — full-sort evaluation: user-tile x item-table scoring fused with train-item masking and per-row top-K, then
// Recall / Precision / HR / NDCG sums.  sm_100a.
//
// Replaces model/lightgcn.py:84-89 (predict_rating: a SECOND full propagation + a materialised B x n_item score
// matrix), training/basic_test.py:42-48 (Python-list mask build, index_put_ of -1024, torch.topk) and
// training/utils.py:7-35 (get_label / pre_rec_k / ndcg_k on the host through a multiprocessing pool).
//
// This file is the exact-fp32 CUDA-core path (sequential fmaf over the feature dimension — the canonical score the
// tests re-derive).  Scores are never written to memory: every thread compares its scores with the row's running
// K-th best (a register), and only the rare survivors go through the mask test (binary search in the user's
// ascending train row) and into a small per-row candidate buffer in shared memory that is compacted to K entries
// by rank-selection when it fills.  Items are visited in ascending id order, so a strict `score > threshold`
// reproduces the (-score, id) order exactly.
//
// Ranking key: the raw dot product (sigmoid is monotone; the reference ranks sigmoid(dot), lightgcn.py:88).  Masked
// (train) items rank below every un-masked item, in id order — as the reference's -1024 does (basic_test.py:47).
#include <float.h>
#include <algorithm>

#include "common.cuh"
#include "eval_tc.cuh"

// packed FFMA2 in the tile loop measured +4 % at dim 64 but -6 % at dim 256 here (and +5..10 % in eval_auc.cu, where it
// is on): scalar kept for this fallback path
#ifndef TAGREC_EVAL_PACKED
#define TAGREC_EVAL_PACKED 0
#endif

namespace tagrec {

constexpr int UT = 64;        // users per block
constexpr int IT = 128;       // items per tile
constexpr int KC = 32;        // feature chunk staged per step
constexpr int KMAX = 128;     // largest supported K
constexpr float MASKED_KEY = -FLT_MAX;

struct EvalArgs {
    const int64_t* users;
    int64_t nu;
    const float* user_table;
    const float* item_table;
    int64_t n_item;
    int dim;
    const int64_t* train_ptr;
    const int32_t* train_items;
    int k;
    int splits;
    int64_t items_per_split;
    float* part_scores;   // [nu, splits, k]
    int32_t* part_ids;
};

__device__ __forceinline__ bool contains(const int32_t* a, int64_t lo, int64_t hi, int32_t x) {
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(a + lo) == x;
}

// Keep the best min(n, k) of a row's n candidates, sorted by (-score, id), in place.  One warp.
__device__ __forceinline__ void compact_row(float* s, int32_t* id, int n, int k, int lane) {
    constexpr int PER = (KMAX + IT + 31) / 32;
    float ms[PER];
    int32_t mi[PER];
    int rank[PER];
#pragma unroll
    for (int t = 0; t < PER; ++t) {
        const int i = lane + 32 * t;
        ms[t] = i < n ? s[i] : 0.f;
        mi[t] = i < n ? id[i] : 0;
        rank[t] = 0;
    }
    for (int j = 0; j < n; ++j) {
        const float sj = s[j];
        const int32_t ij = id[j];
#pragma unroll
        for (int t = 0; t < PER; ++t) rank[t] += (sj > ms[t]) || (sj == ms[t] && ij < mi[t]);
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < PER; ++t) {
        const int i = lane + 32 * t;
        if (i < n && rank[t] < k) {
            s[rank[t]] = ms[t];
            id[rank[t]] = mi[t];
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) eval_topk_kernel(EvalArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int D = a.dim;
    const int cap = a.k + IT;
    float* Us = reinterpret_cast<float*>(smem_raw);              // [D][UT+4]   user tile, feature-major
    float* Is = Us + (size_t)D * (UT + 4);                        // [KC][IT+4]  item chunk, feature-major
    float* cs = Is + (size_t)KC * (IT + 4);                       // [UT][cap]   candidate scores
    int32_t* ci = reinterpret_cast<int32_t*>(cs + (size_t)UT * cap);   // [UT][cap] candidate ids
    float* thr = reinterpret_cast<float*>(ci + (size_t)UT * cap);     // [UT]
    int* cnt = reinterpret_cast<int*>(thr + UT);                       // [UT]
    int64_t* uid = reinterpret_cast<int64_t*>(cnt + UT);               // [UT] global user ids (8B aligned by layout)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tu = tid >> 4, ti = tid & 15;
    const int64_t u0 = (int64_t)blockIdx.x * UT;
    const int split = blockIdx.y;
    const int64_t i_begin = (int64_t)split * a.items_per_split;
    const int64_t i_end = min(a.n_item, i_begin + a.items_per_split);

    if (tid < UT) {
        thr[tid] = -INFINITY;
        cnt[tid] = 0;
        uid[tid] = (u0 + tid < a.nu) ? a.users[u0 + tid] : -1;
    }
    __syncthreads();
    // user tile -> smem (feature-major), rows gathered through `users`
    for (int idx = tid; idx < UT * (D / 4); idx += 256) {
        const int u = idx / (D / 4), c4 = idx % (D / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (uid[u] >= 0) v = __ldg(reinterpret_cast<const float4*>(a.user_table + uid[u] * D) + c4);
        Us[(4 * c4 + 0) * (UT + 4) + u] = v.x;
        Us[(4 * c4 + 1) * (UT + 4) + u] = v.y;
        Us[(4 * c4 + 2) * (UT + 4) + u] = v.z;
        Us[(4 * c4 + 3) * (UT + 4) + u] = v.w;
    }
    float tl[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) tl[r] = -INFINITY;

    for (int64_t it0 = i_begin; it0 < i_end; it0 += IT) {
        unsigned long long accp[4][4];              // accp[r][cp] = scores (r, 2 cp) and (r, 2 cp + 1), packed
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) accp[r][cp] = 0ull;

        for (int k0 = 0; k0 < D; k0 += KC) {
            __syncthreads();
            // item chunk [IT items][KC feats] -> Is[feat][item]
            for (int idx = tid; idx < IT * (KC / 4); idx += 256) {
                const int i = idx / (KC / 4), c4 = idx % (KC / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it0 + i < i_end) v = __ldg(reinterpret_cast<const float4*>(a.item_table + (it0 + i) * D + k0) + c4);
                Is[(4 * c4 + 0) * (IT + 4) + i] = v.x;
                Is[(4 * c4 + 1) * (IT + 4) + i] = v.y;
                Is[(4 * c4 + 2) * (IT + 4) + i] = v.z;
                Is[(4 * c4 + 3) * (IT + 4) + i] = v.w;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < KC; ++kk) {
                const float4 uu = *reinterpret_cast<const float4*>(Us + (size_t)(k0 + kk) * (UT + 4) + 4 * tu);
                const float4 i0 = *reinterpret_cast<const float4*>(Is + (size_t)kk * (IT + 4) + 4 * ti);
                const float4 i1 = *reinterpret_cast<const float4*>(Is + (size_t)kk * (IT + 4) + 64 + 4 * ti);
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
        // ---- selection: only scores above the row's running K-th best survive ----
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int u = 4 * tu + r;
            const int64_t gu = uid[u];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int64_t item = it0 + (c < 4 ? 4 * ti + c : 64 + 4 * ti + (c - 4));
                if (item < i_end && gu >= 0 && acc[r][c] > tl[r]) {
                    float key = acc[r][c];
                    if (contains(a.train_items, a.train_ptr[gu], a.train_ptr[gu + 1], (int32_t)item)) key = MASKED_KEY;
                    if (key > tl[r]) {
                        const int pos = atomicAdd(&cnt[u], 1);
                        cs[(size_t)u * cap + pos] = key;
                        ci[(size_t)u * cap + pos] = (int32_t)item;
                    }
                }
            }
        }
        __syncthreads();
        for (int u = warp; u < UT; u += 8) {
            const int n = cnt[u];
            if (n > a.k) {
                compact_row(cs + (size_t)u * cap, ci + (size_t)u * cap, n, a.k, lane);
                if (lane == 0) {
                    cnt[u] = a.k;
                    thr[u] = cs[(size_t)u * cap + a.k - 1];
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) tl[r] = thr[4 * tu + r];
    }
    // ---- final sort of every row and write-out of this split's K best ----
    for (int u = warp; u < UT; u += 8) {
        const int n = cnt[u];
        compact_row(cs + (size_t)u * cap, ci + (size_t)u * cap, n, a.k, lane);
        if (u0 + u < a.nu) {
            const int kept = min(n, a.k);
            for (int j = lane; j < a.k; j += 32) {
                const size_t o = ((size_t)(u0 + u) * a.splits + split) * a.k + j;
                a.part_scores[o] = j < kept ? cs[(size_t)u * cap + j] : -INFINITY;
                a.part_ids[o] = j < kept ? ci[(size_t)u * cap + j] : -1;
            }
        }
    }
}

// Merge the per-split K-best lists of one user (one warp per user) and emit sigmoid scores / -1024 for masked.
__global__ void __launch_bounds__(256)
eval_merge_kernel(const float* __restrict__ ps, const int32_t* __restrict__ pi, int64_t nu, int splits, int k,
                  int32_t* __restrict__ out_ids, float* __restrict__ out_scores) {
    const int lane = threadIdx.x & 31;
    const int64_t u = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= nu) return;
    const int n = splits * k;
    const float* s = ps + (size_t)u * n;
    const int32_t* id = pi + (size_t)u * n;
    for (int i = lane; i < n; i += 32) {
        const float si = s[i];
        const int32_t ii = id[i];
        if (ii < 0) continue;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float sj = s[j];
            const int32_t ij = id[j];
            rank += (ij >= 0) && ((sj > si) || (sj == si && ij < ii));
        }
        if (rank < k) {
            out_ids[(size_t)u * k + rank] = ii;
            out_scores[(size_t)u * k + rank] = si == MASKED_KEY ? -1024.f : 1.f / (1.f + expf(-si));
        }
    }
}

__global__ void fill_missing_kernel(int32_t* ids, float* scores, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        ids[i] = -1;
        scores[i] = -INFINITY;
    }
}

// training/utils.py:7-35 — one warp per user; double sums; out[4*nk] = recall | precision | hr | ndcg (per k).
__global__ void __launch_bounds__(256)
eval_metrics_kernel(const int64_t* __restrict__ users, int64_t nu, const int32_t* __restrict__ topk, int kmax,
                    const int64_t* __restrict__ test_ptr, const int32_t* __restrict__ test_items,
                    const int32_t* __restrict__ ks, int nk, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nu) return;
    const int64_t u = users[w];
    const int64_t tb = test_ptr[u], te = test_ptr[u + 1];
    const double n_true = (double)(te - tb);
    if (te == tb) return;
    for (int q = 0; q < nk; ++q) {
        const int k = ks[q];
        double right = 0.0, dcg = 0.0, idcg = 0.0;
        for (int r = lane; r < k && r < kmax; r += 32) {
            const int32_t item = topk[(size_t)w * kmax + r];
            const double disc = 1.0 / log2((double)(r + 2));
            if (item >= 0 && contains(test_items, tb, te, item)) {
                right += 1.0;
                dcg += disc;
            }
            if ((double)r < n_true) idcg += disc;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            right += __shfl_xor_sync(0xffffffffu, right, o);
            dcg += __shfl_xor_sync(0xffffffffu, dcg, o);
            idcg += __shfl_xor_sync(0xffffffffu, idcg, o);
        }
        if (lane == 0) {
            if (idcg == 0.0) idcg = 1.0;
            atomicAdd(out + 0 * nk + q, right / n_true);
            atomicAdd(out + 1 * nk + q, right / (double)k);
            atomicAdd(out + 2 * nk + q, right > 0.0 ? 1.0 : 0.0);
            atomicAdd(out + 3 * nk + q, dcg / idcg);
        }
    }
}

static int pick_splits(int64_t nu, int64_t n_item) {
    const int64_t user_tiles = (nu + UT - 1) / UT;
    int64_t s = (2 * kSMs + user_tiles - 1) / user_tiles;          // aim at >= 2 blocks per SM
    const int64_t max_s = (n_item + 4 * IT - 1) / (4 * IT);          // at least 4 item tiles per split
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 64) s = 64;
    return (int)s;
}

}  // namespace tagrec

using namespace tagrec;

extern "C" size_t tagrec_eval_workspace_bytes(int64_t nu, int64_t n_item, int k) {
    // large enough for either path (the tensor-core plan depends on dim only through dim == 64)
    const int s = pick_splits(nu, n_item);
    size_t need = (size_t)nu * s * k * 8 + 256;
    for (int dim = 64; dim <= 256; dim += 192) {      // the split count differs between the 64-d and the wide kernel
        const TcPlan p = tc_plan(nu, n_item, dim, k);
        if (p.ok) need = std::max(need, eval_tc_workspace_bytes(nu, p, k));
    }
    return need;
}

extern "C" int tagrec_eval_plan(int64_t nu, int64_t n_item, int dim, int k, int32_t* plan) {
    TAGREC_REQUIRE(plan, "plan is null");
    const TcPlan p = tc_plan(nu, n_item, dim, k);
    plan[0] = p.ok ? 1 : 0;
    plan[1] = p.ok ? p.cg2 : 0;
    plan[2] = p.ok ? p.nh : 0;
    plan[3] = p.ok ? p.splits : 0;
    plan[4] = p.ok ? p.stages : 0;
    plan[5] = p.ok ? p.lists : 0;
    return TAGREC_OK;
}

extern "C" int tagrec_eval_topk(const int64_t* users, int64_t nu, const float* user_table, const float* item_table,
                                int64_t n_item, int dim, const int64_t* train_ptr, const int32_t* train_items, int k,
                                int32_t* topk_ids, float* topk_scores, void* workspace, size_t workspace_bytes,
                                void* stream) {
    return tagrec_eval_topk_ex(users, nu, user_table, item_table, n_item, dim, train_ptr, train_items, k, topk_ids,
                               topk_scores, workspace, workspace_bytes, TAGREC_EVAL_AUTO, stream);
}

extern "C" int tagrec_eval_topk_ex(const int64_t* users, int64_t nu, const float* user_table, const float* item_table,
                                   int64_t n_item, int dim, const int64_t* train_ptr, const int32_t* train_items,
                                   int k, int32_t* topk_ids, float* topk_scores, void* workspace,
                                   size_t workspace_bytes, int path, void* stream) {
    TAGREC_REQUIRE(users && user_table && item_table && train_ptr && topk_ids && topk_scores, "null pointer");
    TAGREC_REQUIRE(k >= 1 && k <= KMAX, "k must be in 1..128");
    TAGREC_REQUIRE(path == TAGREC_EVAL_AUTO || path == TAGREC_EVAL_FP32 || path == TAGREC_EVAL_TF32, "bad path");
    TAGREC_REQUIRE(n_item > 0 && n_item < (1ll << 31), "n_item out of range");
    if (path != TAGREC_EVAL_FP32) {
        const TcPlan p = tc_plan(nu, n_item, dim, k);
        if (p.ok)
            return eval_topk_tc(users, nu, user_table, item_table, n_item, train_ptr, train_items, k, topk_ids,
                                topk_scores, workspace, workspace_bytes, stream, p);
        TAGREC_REQUIRE(path == TAGREC_EVAL_AUTO, "tensor-core path needs dim in {64, 128, 192, 256} and k <= 128");
    }
    TAGREC_REQUIRE(dim >= 4 && dim % KC == 0, "dim must be a multiple of 32");
    TAGREC_REQUIRE(n_item > 0 && n_item < (1ll << 31), "n_item out of range");
    if (nu == 0) return TAGREC_OK;
    const int splits = pick_splits(nu, n_item);
    const size_t need = (size_t)nu * splits * k * 8 + 256;
    if (!workspace || workspace_bytes < need) return fail(TAGREC_ENOMEM, "eval workspace too small", __FILE__, __LINE__);
    EvalArgs a{};
    a.users = users; a.nu = nu; a.user_table = user_table; a.item_table = item_table; a.n_item = n_item; a.dim = dim;
    a.train_ptr = train_ptr; a.train_items = train_items; a.k = k; a.splits = splits;
    const int64_t tiles = (n_item + IT - 1) / IT;
    a.items_per_split = ((tiles + splits - 1) / splits) * IT;
    a.part_scores = reinterpret_cast<float*>(workspace);
    a.part_ids = reinterpret_cast<int32_t*>(a.part_scores + (size_t)nu * splits * k);
    const int cap = k + IT;
    const size_t smem = ((size_t)dim * (UT + 4) + (size_t)KC * (IT + 4) + (size_t)UT * cap * 2 + UT * 2) * 4 + UT * 8;
    TAGREC_REQUIRE(smem <= 227 * 1024, "dim/k too large for shared memory");
    TAGREC_CUDA(cudaFuncSetAttribute(eval_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((unsigned)((nu + UT - 1) / UT), (unsigned)splits);
    TAGREC_LAUNCH(eval_topk_kernel, grid, 256, smem, stream, a);
    const int64_t tot = nu * k;
    TAGREC_LAUNCH(fill_missing_kernel, (unsigned)((tot + 255) / 256), 256, 0, stream, topk_ids, topk_scores, tot);
    TAGREC_LAUNCH(eval_merge_kernel, (unsigned)((nu + 7) / 8), 256, 0, stream, a.part_scores, a.part_ids, nu, splits, k,
                  topk_ids, topk_scores);
    return TAGREC_OK;
}

extern "C" int tagrec_eval_metrics(const int64_t* users, int64_t nu, const int32_t* topk_ids, int kmax,
                                   const int64_t* test_ptr, const int32_t* test_items, const int32_t* ks, int nk,
                                   double* out, void* stream) {
    TAGREC_REQUIRE(users && topk_ids && test_ptr && test_items && ks && out, "null pointer");
    TAGREC_REQUIRE(nk >= 1 && kmax >= 1, "bad k list");
    if (nu == 0) return TAGREC_OK;
    TAGREC_LAUNCH(eval_metrics_kernel, (unsigned)((nu + 7) / 8), 256, 0, stream, users, nu, topk_ids, kmax, test_ptr,
                  test_items, ks, nk, out);
    return TAGREC_OK;
}
