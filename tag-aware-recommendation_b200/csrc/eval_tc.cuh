// Interface between eval_topk.cu (dispatch, fp32 CUDA-core path) and eval_tc.cu (tcgen05 / TMEM / TMA path).
#pragma once
#include "common.cuh"

namespace tagrec {

// Shape of a tensor-core launch: NH 128-user halves per CTA, item splits, B stages — derived from (nu, n_item, k).
struct TcPlan {
    int nh, splits, stages, kb;      // kb = dim / 64 feature blocks (1: eval_tc_kernel, 2..4: eval_tc_wide_kernel)
    int64_t items_per_split;
    size_t smem;
    bool ok;
};
TcPlan tc_plan(int64_t nu, int64_t n_item, int dim, int k);
size_t eval_tc_workspace_bytes(int64_t nu, const TcPlan& p, int k);
int eval_topk_tc(const int64_t* users, int64_t nu, const float* user_table, const float* item_table, int64_t n_item,
                 const int64_t* train_ptr, const int32_t* train_items, int k, int32_t* topk_ids, float* topk_scores,
                 void* workspace, size_t workspace_bytes, void* stream, const TcPlan& p);

}  // namespace tagrec
