// Interface between eval_topk.cu (dispatch, fp32 CUDA-core path) and eval_tc.cu (tcgen05 / TMEM / TMA path).
#pragma once
#include "common.cuh"

namespace tagrec {

// Shape of a tensor-core launch: NH 128-user halves per CTA, item splits, B stages — derived from (nu, n_item, k).
struct TcPlan {
    int nh, splits, stages, kb;      // kb = dim / 64 feature blocks (1: eval_tc_kernel, 2..4: eval_tc_wide_kernel)
    int cg2;                         // 1: eval_tc2_kernel (CTA pairs, tcgen05.mma.cta_group::2 M256 x N256), dim 64 only
    int lists;                       // K-lists per user handed to the merge kernel: splits (x 2 column halves for cg2)
    int64_t items_per_split;
    size_t smem;
    bool ok;
};

// Arguments shared by the tensor-core top-K kernels (eval_tc.cu, eval_tc2.cu).
struct TcArgs {
    const int64_t* users;
    int64_t nu;
    const float* user_table;
    const float* item_table;
    int64_t n_item;
    const int64_t* train_ptr;
    const int32_t* train_items;
    const float* item_maxnorm;   // device scalar: max_i ||I_i||_2
    float* shared_thr;           // [nu] or NULL: per-user lower bound of the final K-th best score, shared by the
                                 // K-lists of that user (max over lists of their own exact K-th best)
    int k;
    int splits;
    int stages;
    int64_t items_per_split;     // multiple of the tile width
    float* part_scores;          // [nu, lists, k]  raw exact dot products, -inf when absent
    int32_t* part_ids;           // [nu, lists, k]  -1 when absent
};

// eval_tc2.cu: shared-memory footprint and launch of the CTA-pair kernel (grid = 2 * ceil(nu / 256) x splits)
size_t tc2_smem(int stages, int k);
int launch_eval_tc2(const void* item_map, const TcArgs& a, size_t smem, void* stream);
TcPlan tc_plan(int64_t nu, int64_t n_item, int dim, int k);
size_t eval_tc_workspace_bytes(int64_t nu, const TcPlan& p, int k);
int eval_topk_tc(const int64_t* users, int64_t nu, const float* user_table, const float* item_table, int64_t n_item,
                 const int64_t* train_ptr, const int32_t* train_items, int k, int32_t* topk_ids, float* topk_scores,
                 void* workspace, size_t workspace_bytes, void* stream, const TcPlan& p);

}  // namespace tagrec
