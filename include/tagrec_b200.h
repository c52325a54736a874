/*
 * tagrec_b200.h — C ABI of libtagrec_b200.so (sm_100a).
 *
 * Drop-in boundary for the graph-embedding train + full-sort evaluation hot path of
 * chenzheng5555/tag-aware-recommendation.  The reference has NO FFI of its own (it is 100 % Python); each entry
 * point below names the reference call site(s) (file:line, relative to the reference root) whose device work it
 * replaces.  The reference-side binding is the ctypes stub in INTEGRATION.md / tag-aware-recommendation_b200/_lib.py.
 *
 * Conventions
 *   - every function returns 0 on success, a negative TAGREC_E* code otherwise; tagrec_last_error() gives text
 *     (thread-local);
 *   - the CALLER owns every buffer (torch allocates); the library never frees or keeps a pointer after the call;
 *   - all pointers are DEVICE pointers on the current device unless the parameter is documented "host";
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream); launches are
 *     asynchronous and never synchronise, except the functions documented "synchronises";
 *   - tables are row-major float32 [n_rows, dim]; rows must be 16-byte aligned (dim % 4 == 0);
 *   - CSR: rowptr int64 [n+1], col int32 [nnz] ascending inside a row, val float32 [nnz].
 */
#ifndef TAGREC_B200_H
#define TAGREC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAGREC_OK 0
#define TAGREC_EINVAL (-1)   /* bad argument (null pointer, unsupported dim, ...) */
#define TAGREC_ECUDA (-2)    /* a CUDA runtime call / launch failed */
#define TAGREC_ENOMEM (-3)   /* caller-provided workspace or output capacity too small */

int tagrec_version(void);
const char* tagrec_last_error(void);
/* Number of kernels this library has launched in this process (bench.py's "gpu_launches"). */
uint64_t tagrec_launch_count(void);
/* sizeof() of the descriptor structs below as THIS library was compiled: a binding (ctypes / cffi / cgo ...) asserts
 * its own struct layout against these before the first call, so a stale declaration fails at load time instead of
 * reading past a short struct.  which: 0 = tagrec_csr_t, 1 = tagrec_mirror_t, 2 = tagrec_route_plan_t, 3 = tagrec_adam_t; other -> 0. */
size_t tagrec_sizeof_struct(int which);

/* ------------------------------------------------------------------------------------------------------------
 * K0  adjacency -> CSR            replaces model/help/adj.py:7-35 (create_ui_adj / create_uit_adj, lil_matrix
 *                                 block assembly), adj.py:90-110 (bi_norm / si_norm value computation) and
 *                                 adj.py:144-150 (sp2tensor).
 * The degree power d = np.power(rowsum, p) stays on the HOST in numpy (adj.py:93,105): numpy's float32 pow is not
 * reproducible on device, and bit-exact values need that very function.
 * ---------------------------------------------------------------------------------------------------------- */

/* Bytes of workspace tagrec_csr_build_structure needs for `n_directed` = 2*(e_ui+e_ut+e_it) (+ n if self loops). */
size_t tagrec_csr_workspace_bytes(int64_t n_directed);

/* Block adjacency [[0,R],[R^T,0]] (or the 3x3 user/item/tag block matrix when e_ut+e_it > 0), duplicates summed,
 * rows ascending, columns ascending inside a row.
 *   ui_row/ui_col (e_ui), ut_row/ut_col (e_ut), it_row/it_col (e_it): int64 block-local indices (device).
 *   self_loops: 0 none | 1 add I BEFORE the degree is taken (adj.py:81 'si_norm_self')
 *                      | 2 add I with weight 1 that is NOT part of the degree (adj.py:83 'ngcf').
 * Outputs: rowptr[n+1]; col[cap], weight[cap] (integer multiplicities as float, adj.py data before normalising);
 *   degree[n] = float32 weighted row sum (adj.py:92,103); *nnz_host (HOST) = entries written.
 * Synchronises `stream` (it has to return nnz). */
int tagrec_csr_build_structure(const int64_t* ui_row, const int64_t* ui_col, int64_t e_ui,
                               const int64_t* ut_row, const int64_t* ut_col, int64_t e_ut,
                               const int64_t* it_row, const int64_t* it_col, int64_t e_it,
                               int64_t n_user, int64_t n_item, int64_t n_tag, int self_loops,
                               void* workspace, size_t workspace_bytes,
                               int64_t* rowptr, int32_t* col, float* weight, int64_t cap,
                               float* degree, int64_t* nnz_host, void* stream);

/* val[j] = (dpow[row]*w[j])*dpow[col]   mode 0  bi_norm   (adj.py:97, two roundings, left to right)
 *        =  dpow[row]*w[j]              mode 1  si_norm / si_norm_self / ngcf (diagonal of 'ngcf' stays 1)
 *        =  w[j]*dpow[col]              mode 2  transpose of mode 1 (values of A^T for the backward SpMM)
 *        =  w[j]                        mode 3  'plain'
 * self_loops as above (mode 1/2 with self_loops==2 keep the diagonal at exactly 1). */
int tagrec_csr_normalise(const int64_t* rowptr, const int32_t* col, const float* weight, const float* dpow,
                         int64_t n, int mode, int self_loops, float* val, void* stream);

/* Plan builder of the column-blocked K1 (no reference equivalent): bounds[w * n_sel + i] = index of the first stored
 * entry of row rows[i] (local row ids, ascending columns per row) whose column id is >= w * window, for w = 0 .. n_win
 * — i.e. where each L2-sized window of the gathered table starts inside each selected row. */
int tagrec_csr_window_bounds(const int64_t* rowptr, const int32_t* col, const int32_t* rows, int64_t n_sel,
                             int64_t window, int n_win, int64_t* bounds, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K1  CSR SpMM with fused epilogues      replaces model/help/adj.py:158-167 (split_mm = torch.sparse.mm) and
 *                                        its autograd transpose, plus model/lightgcn.py:54-60.
 * Long rows (> TAGREC_LONG_ROW nnz) are split into chunks whose partial sums meet in `long_scratch`
 * (n_long x dim floats, zero on entry, left zero on exit) guarded by `long_counter` (n_long ints, zero/zero).
 * long_rows[n_long] = row ids; item_slot/item_begin/item_end[n_items] = chunk -> (slot in long_rows, nnz range).
 * All five may be NULL when n_long == 0.
 * ---------------------------------------------------------------------------------------------------------- */
#define TAGREC_LONG_ROW 4096      /* defaults, tuned on the 1.9e9-nnz graph; small graphs (everything L2-resident, a few */
#define TAGREC_LONG_CHUNK 2048    /* waves of rows) balance better with 256 / 256 — the caller's plan decides */

typedef struct {
    const int64_t* rowptr;      /* n_rows + 1 entries, local (rowptr[0] == 0 for a row block) */
    const int32_t* col;         /* GLOBAL column ids: rows of the gathered table */
    const float* val;
    int64_t n_rows;             /* rows of this block */
    int64_t row_offset;         /* global id of local row 0: every epilogue table (y, acc, e_k, g_final, ...) is a
                                   full-size table indexed by row_offset + r; 0 on a single GPU */
    const int32_t* long_rows;   /* LOCAL row ids */
    const int32_t* item_slot;
    const int64_t* item_begin;
    const int64_t* item_end;
    int64_t n_long;
    int64_t n_items;
    float* long_scratch;
    int32_t* long_counter;
    int32_t long_row;           /* rows with more entries than this are "long"; 0 = TAGREC_LONG_ROW */
    int32_t long_chunk;         /* entries per chunk of a long row;             0 = TAGREC_LONG_CHUNK */
    /* Column-blocked plans (L2-windowed gathers).  A chunk may be ANY [begin, end) piece of its row:
     * the plan cuts the rows that gather from a table far larger than the L2 at column-window boundaries and lists
     * the pieces window-major, so that the chunks in flight at any time gather from one L2-sized window of the source
     * table.  long_nchunks[slot] = number of pieces of that row (NULL: ceil(degree / long_chunk), the plain plan).
     * Rows with local index >= blocked_row_begin are "long" already above blocked_min_deg entries (0: long_row).
     * chunk_lanes = 0: one warp per chunk; 1: one sub-warp (dim/4 lanes) per chunk — short pieces. */
    const int32_t* long_nchunks;
    int64_t blocked_row_begin;
    int32_t blocked_min_deg;
    int32_t chunk_lanes;
    /* Row-subset launches (round 2).  row_list != NULL: the launch produces the n_rows LISTED rows only
     * (row_list[i] = local row id, any order, no duplicates, negative entries skipped; rowptr still covers the whole
     * block) instead of rows
     * 0 .. n_rows-1; listed rows that are long are produced by their chunks in the item list (see row_sel; n_rows == 0
     * with n_items > 0 is valid).  The last LightGCN layer of a training step is needed on the batch's rows only
     * (model/lightgcn.py:59-60 feeds the mean table to the loss, which reads 3 x batch rows of it). */
    const int32_t* row_list;
    const uint8_t* row_sel;     /* with row_list: one byte per LOCAL row, non-zero = listed.  The chunk blocks of a long row
                                   that is not listed exit at once, so the caller may pass the plan's whole chunk list
                                   (no per-call list of the listed long rows' chunks has to be built) */
} tagrec_csr_t;

/* Fused compute + collective (multi-GPU, no reference equivalent — the reference is single-device): where an
 * output row of K1 is stored.  n == 0 / NULL: the local table only.  n >= 1: the same [N, dim] table on n ranks of
 * one NVSwitch domain; base[r] = rank r's copy mapped into this process (CUDA peer / symmetric memory), base[self]
 * included.  A single NVLS multicast address is expressed as n == 1.  The row block's all-gather then happens
 * inside the SpMM epilogue (stores over NVLink 5 overlap the gathers from HBM); the caller places a cross-rank
 * barrier between this launch and the first launch that reads the table. */
/* MULTI-GPU CONTRACT (replaces the tagrec_comm_* wrappers SURVEY §8(b) sketched; decided against them on purpose).
 * The library owns NO communicator and does NO rendezvous: it never calls NCCL, never opens IPC handles and keeps no
 * cross-rank state.  One process per GPU; the CALLER
 *   1. allocates every table other ranks store into ([N, dim] fp32: layer tables, mean table, gradient ping-pong
 *      tables, dL/dE0, the parameter table for tagrec_adam_step_mirror) as symmetric / peer-mapped memory — cuMemCreate
 *      + cuMemExportToShareableHandle / cuMemMap, cudaIpc*, NVSHMEM, or torch.distributed._symmetric_memory (what
 *      tag-aware-recommendation_b200/distributed.py::PeerTables uses) — and passes the mapped addresses here:
 *      base[r] = rank r's copy of the SAME table as seen from this process (base[self] = the local copy), or n == 1
 *      with base[0] = an NVLS multicast address bound to all replicas (cuMulticastCreate / BindMem);
 *   2. passes `row_offset` / `n_rows` of its contiguous row block in tagrec_csr_t (global column ids are kept);
 *   3. places a cross-rank barrier ON THE SAME STREAM between a launch that stores through a mirror and the first
 *      launch (on any rank) that reads that table — a device-side signal exchange (symmetric-memory barrier) or a
 *      1-element NCCL all-reduce; stores through a mirror are plain st.global and are visible after that barrier;
 *   4. runs its own small collectives (the 3*B-row all-reduce of the first backward table, metric sums of the sharded
 *      evaluation) with whatever library it uses; the kernels only need the result in device memory.
 * Without mirrors (NULL / n == 0) every kernel writes its local table only and the caller all-gathers row blocks
 * (ncclAllGather over [row_offset, row_offset + n_rows)) — the baseline path, also supported. */
#define TAGREC_MAX_PEERS 8
typedef struct {
    int32_t n;
    int32_t self;
    void* base[TAGREC_MAX_PEERS];
} tagrec_mirror_t;

/* y = A x (+ beta*y)                        adj.py:162,166; also A^T g when given the transposed values. */
int tagrec_spmm(const tagrec_csr_t* a, const float* x, float* y, int dim, float beta, void* stream);

/* One LightGCN layer, lightgcn.py:55-60:   y = A x   (raw, propagates)
 *   acc = (first ? x : acc) + y / max(||y||_2, 1e-12);  if (last) acc *= final_scale   [final_scale = 1/(L+1)] */
int tagrec_lightgcn_fwd_layer(const tagrec_csr_t* a, const float* x, float* y, float* acc, int dim, int first,
                              int last, float final_scale, void* stream);

int tagrec_lightgcn_fwd_layer_p2p(const tagrec_csr_t* a, const float* x, float* y, float* acc, int dim, int first,
                                  int last, float final_scale, const tagrec_mirror_t* y_mirror,
                                  const tagrec_mirror_t* acc_mirror, void* stream);

/* One backward layer of the same (closed form of autograd through lightgcn.py:55-60):
 *   gy = g_final * (upstream ? upstream[0] : 1) * inv_layers
 *   e_k != NULL :  g_out = nb(gy, e_k) + (g_next ? A g_next : 0)
 *                  nb(g,e) = (g - y (y.g)) / ||e||,  y = e/||e||   (g / 1e-12 when ||e|| < 1e-12)
 *   e_k == NULL :  g_out = gy + A g_next + (reg_grad ? (upstream ? upstream[1] : 1) * reg_grad : 0)
 * `a` must hold the values of A^T (== A for bi_norm).  g_out may alias reg_grad. */
int tagrec_lightgcn_bwd_layer(const tagrec_csr_t* a, const float* g_next, const float* e_k, const float* g_final,
                              const float* reg_grad, const float* upstream, float inv_layers, float* g_out,
                              int dim, void* stream);

int tagrec_lightgcn_bwd_layer_p2p(const tagrec_csr_t* a, const float* g_next, const float* e_k, const float* g_final,
                                  const float* reg_grad, const float* upstream, float inv_layers, float* g_out,
                                  int dim, const tagrec_mirror_t* out_mirror, void* stream);

/* Same with zero-row skipping: g_next_nz (optional) = one byte per row of g_next, 0 where that row is all-zero
 * (tagrec_row_nonzero).  The upstream gradient of a BPR batch touches 3*B rows, so the first backward tables are
 * non-zero only on the batch's nodes / their neighbours: the 256 B gathers of the other rows are never issued. */
int tagrec_lightgcn_bwd_layer_ex(const tagrec_csr_t* a, const float* g_next, const uint8_t* g_next_nz, const float* e_k,
                                 const float* g_final, const float* reg_grad, const float* upstream, float inv_layers,
                                 float* g_out, int dim, const tagrec_mirror_t* out_mirror, void* stream);
/* The LAST backward launch of a step with the optimizer folded into its epilogue (owner-sharded Adam, multi-GPU; also
 * valid on one GPU): the launch produces rows of dL/dE0 — the gradient of the embedding table itself — and each row is
 * consumed where it is produced: exp_avg / exp_avg_sq / param rows are updated with the same arithmetic as
 * tagrec_adam_step (torch.optim.Adam, amsgrad = False) and the NEW parameter row is stored through param_mirror into
 * every rank's replica, so the 3 GB parameter exchange of a step rides under the gathers of this launch instead of
 * following it (measured at 8 GPUs: 4.2 ms per step for the separate tagrec_adam_step_mirror pass, NVLink-ingest bound).
 * param / exp_avg / exp_avg_sq are FULL-size [n, dim] tables indexed by global row; g_out (optional, may be NULL)
 * still receives the gradient rows locally.  Nothing in this launch reads `param` besides the row being updated. */
typedef struct tagrec_adam_t {
    float* param;
    float* exp_avg;
    float* exp_avg_sq;
    float lr, beta1, beta2, eps, weight_decay;
    int32_t reserved;
    int64_t step;                   /* the step being taken, counts from 1 (bias corrections) */
    tagrec_mirror_t param_mirror;   /* n == 0: local table only */
} tagrec_adam_t;
int tagrec_lightgcn_bwd_layer_adam(const tagrec_csr_t* a, const float* g_next, const uint8_t* g_next_nz,
                                   const float* g_final, const float* reg_grad, const float* upstream,
                                   float inv_layers, float* g_out, int dim, const tagrec_adam_t* adam, void* stream);
/* Push form for SPARSE sources (round 2).  When the source table of a backward launch is non-zero on a handful of
 * rows only — the item-row half of the first backward launch of a BPR step sums over USER sources, and only the batch's
 * <= B users carry a gradient — gathering means scanning every stored entry of those rows to find the few that count.
 * tagrec_spmm_push_rows goes the other way: for each listed source row r (keep[i] == 0 skips a duplicate) it adds
 * val[j] * x[r] into y_acc[col[j]] for the entries j of CSR row r (for out = A^T g this is row r of A itself; y_acc is a
 * full-size table the caller keeps all-zero otherwise).  tagrec_lightgcn_bwd_layer_acc then runs the fused backward
 * epilogue of tagrec_lightgcn_bwd_layer (with e_k) on the rows of `a` WITHOUT a gather, taking each row's sum from
 * acc_in and zeroing the rows it consumed. */
int tagrec_spmm_push_rows(const int64_t* rowptr, const int32_t* col, const float* val, const int64_t* rows,
                          const uint8_t* keep, int64_t n_rows, const float* x, float* y_acc, int dim, void* stream);
int tagrec_lightgcn_bwd_layer_acc(const tagrec_csr_t* a, float* acc_in, const float* e_k, const float* g_final,
                                  const float* upstream, float inv_layers, float* g_out, int dim,
                                  const tagrec_mirror_t* out_mirror, void* stream);
/* nz[r] = (row r of the [n, dim] table has a non-zero element). */
int tagrec_row_nonzero(const float* table, int64_t n, int dim, uint8_t* nz, void* stream);

/* Sparse form of the FIRST backward table of a BPR step (same closed form as tagrec_lightgcn_bwd_layer with
 * g_next == NULL, restricted to the listed rows): dL/dF is non-zero on the batch's nodes only, hence so is
 * G_L = nb(g_final * upstream[0] * inv_layers, e_k).  nodes[n_nodes] = GLOBAL row ids (duplicates allowed).  Rows in
 * [row_lo, row_hi) are written to g_out — a table the caller keeps all-zero otherwise — and every listed row is flagged
 * in nz (the byte map tagrec_lightgcn_bwd_layer_ex takes as g_next_nz; may be NULL).  tagrec_rows_zero restores the
 * all-zero state (table and / or map) for the same list afterwards: O(batch) instead of O(N) work per step. */
int tagrec_lightgcn_bwd_first_sparse(const int64_t* nodes, int64_t n_nodes, int64_t row_lo, int64_t row_hi,
                                     const float* e_k, const float* g_final, const float* upstream, float inv_layers,
                                     float* g_out, uint8_t* nz, int dim, void* stream);
int tagrec_rows_zero(const int64_t* nodes, int64_t n_nodes, float* table, uint8_t* nz, int dim, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K2  fused BPR step            replaces model/lightgcn.py:68-82 / model/ngcf.py:95-105 (3 gathers, mul_loss,
 *                               l2reg_loss: model/help/loss.py:4-12,27-32) and their index_put_ backward.
 *   triples: int64 [b,3] row-major (u, i+, i-) — train_data/bpr_training_data.py:44; items index the table at
 *   row item_offset + i.  loss_kind 0 softplus, 1 logsigmoid (loss.py:8-11).
 *   loss_out[0] = mean loss, loss_out[1] = reg * 1/2 sum ||rows of reg_src||^2 / b  (overwritten)
 *   g_final += d loss / d final;  g_reg += d loss_out[1] / d reg_src (skipped when reg == 0 or g_reg NULL).
 * ---------------------------------------------------------------------------------------------------------- */
int tagrec_bpr_fwd_bwd(const int64_t* triples, int64_t b, int64_t item_offset, const float* final_table,
                       const float* reg_src, int dim, float reg, int loss_kind, float* g_final, float* g_reg,
                       float* loss_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K3  full-sort evaluation      replaces model/lightgcn.py:84-89 (predict_rating), training/basic_test.py:40-48
 *                               (mask train items with -1024, torch.topk) and training/utils.py:7-35 (metrics).
 *   users int64 [nu]; user_table [*,dim], item_table [n_item,dim];
 *   train_ptr int64 [n_user+1] / train_items int32 (ascending per user): items to mask;
 *   topk_ids int32 [nu,k] / topk_scores float32 [nu,k]: best k by (-score, item id), score = sigmoid(dot) as the
 *   reference ranks it (masked items rank as -1024).
 * ---------------------------------------------------------------------------------------------------------- */
int tagrec_eval_topk(const int64_t* users, int64_t nu, const float* user_table, const float* item_table,
                     int64_t n_item, int dim, const int64_t* train_ptr, const int32_t* train_items, int k,
                     int32_t* topk_ids, float* topk_scores, void* workspace, size_t workspace_bytes, void* stream);
size_t tagrec_eval_workspace_bytes(int64_t nu, int64_t n_item, int k);

/* Same, with the scoring path chosen by the caller.  Both paths return the same top-K (by (-score, id) of the exact
 * fp32 dot product, sequential fmaf over the feature index):
 *   TAGREC_EVAL_FP32  CUDA-core fp32 tiles (any dim % 32 == 0);
 *   TAGREC_EVAL_TF32  dim in {64, 128, 192, 256}: tcgen05.mma kind::tf32 (accumulators and user rows in TMEM, item
 *                     tiles by TMA) as a filter with a proven error margin, every candidate re-scored in exact fp32
 *                     (csrc/eval_tc.cu); 64-d tables with >= 256-user batches run on CTA pairs (tcgen05.mma
 *                     cta_group::2, M256 x N256 x K8; csrc/eval_tc2.cu);
 *   TAGREC_EVAL_AUTO  TF32 when the dim allows it, else FP32 (what tagrec_eval_topk does). */
#define TAGREC_EVAL_AUTO 0
#define TAGREC_EVAL_FP32 1
#define TAGREC_EVAL_TF32 2
int tagrec_eval_topk_ex(const int64_t* users, int64_t nu, const float* user_table, const float* item_table,
                        int64_t n_item, int dim, const int64_t* train_ptr, const int32_t* train_items, int k,
                        int32_t* topk_ids, float* topk_scores, void* workspace, size_t workspace_bytes, int path,
                        void* stream);
/* The launch shape tagrec_eval_topk[_ex] would use for the tensor-core path (reporting / tests; no GPU needed):
 * plan[0] = 1 if the shape runs on tensor cores, [1] = 1 for the CTA-pair kernel (cta_group::2), [2] = 128-user halves
 * per CTA, [3] = item splits, [4] = TMA stages, [5] = K-lists per user handed to the merge kernel. */
int tagrec_eval_plan(int64_t nu, int64_t n_item, int dim, int k, int32_t* plan);

/* Per-user AUC (training/utils.py:37-45 roc_auc_score over the un-masked items of one user, basic_test.py:52-53),
 * summed over `users`: out[0] += sum of AUC_u, out[1] += number of users that have both classes (the reference raises
 * for the others).  Positives = test items of u that are not train items; ties get half credit; ranking quantity =
 * exact fp32 dot (see csrc/eval_auc.cu).  test_items int32 ascending per user; n_test_total = test_ptr[n_user]. */
size_t tagrec_eval_auc_workspace_bytes(int64_t nu, int64_t n_test_total);
int tagrec_eval_auc(const int64_t* users, int64_t nu, const float* user_table, const float* item_table, int64_t n_item,
                    int dim, const int64_t* train_ptr, const int32_t* train_items, const int64_t* test_ptr,
                    const int32_t* test_items, int64_t n_test_total, void* workspace, size_t workspace_bytes, double* out,
                    void* stream);
/* same, with the path of the dense pass chosen explicitly: TAGREC_EVAL_AUTO (tensor cores when dim == 64), _FP32 (CUDA-core
 * tiles), _TF32 (3xTF32 tcgen05 MMAs + exact re-scores inside the error margin; the sums are identical to _FP32's). */
int tagrec_eval_auc_ex(const int64_t* users, int64_t nu, const float* user_table, const float* item_table,
                       int64_t n_item, int dim, const int64_t* train_ptr, const int32_t* train_items,
                       const int64_t* test_ptr, const int32_t* test_items, int64_t n_test_total, void* workspace,
                       size_t workspace_bytes, double* out, int path, void* stream);

/* metric sums over users (training/utils.py:15-35): out[4*nk] = recall|precision|hr|ndcg per k (double, +=). */
int tagrec_eval_metrics(const int64_t* users, int64_t nu, const int32_t* topk_ids, int kmax,
                        const int64_t* test_ptr, const int32_t* test_items, const int32_t* ks, int nk,
                        double* out, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K6  NGCF dense half-layer (64 -> 64)     replaces model/ngcf.py:77-86 and its autograd: two products against
 *                                          (W + b) — bias added to the WEIGHT matrix —, LeakyReLU(0.2), sum,
 *                                          F.normalize.  All tables [n, 64] row-major; w* [64, 64], b* [64].
 *   fwd: out = lrelu((nei+e)(w1+b1)) + lrelu((nei*e)(w2+b2)); nrm = out / max(||out||, 1e-12);
 *        s_act / t_act = the two activations (saved for backward).
 *   bwd: g = g_out (may be NULL) + J_normalize(out)^T g_nrm (rows g_nrm_ld floats apart: a column slice of the
 *        concatenated gradient); gs = g*lrelu'(s), gt = g*lrelu'(t) (optional outputs, may be NULL);
 *        g_nei = gs(w1+b1)^T + (gt(w2+b2)^T)*e;  g_e = gs(w1+b1)^T + (gt(w2+b2)^T)*nei;
 *        dw1 += (nei+e)^T gs, dw2 += (nei*e)^T gt  ([64, 64] each = the gradient of W and, summed over rows, of the
 *        bias that ngcf.py:78,82 adds to the WEIGHT; caller zeroes them; both NULL to skip).
 * ---------------------------------------------------------------------------------------------------------- */
int tagrec_ngcf_dense_fwd(const float* nei, const float* e, const float* w1, const float* b1, const float* w2,
                          const float* b2, int64_t n, int dim, float* out, float* nrm, float* s_act, float* t_act,
                          void* stream);
int tagrec_ngcf_dense_bwd(const float* g_out, const float* g_nrm, int64_t g_nrm_ld, const float* out,
                          const float* s_act, const float* t_act, const float* nei, const float* e, const float* w1,
                          const float* b1, const float* w2, const float* b2, int64_t n, int dim, float* g_nei,
                          float* g_e, float* gs, float* gt, float* dw1, float* dw2, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K5  disentangled routing (DGCF / DisenGCN)     replaces model/dgcf.py:68-110 (iterate_update, factor_update) and
 *                                                model/disengcn.py:29-43 (neighbour routing inside Layer.forward).
 * Structure = rowptr int64 [n+1] / col int32 [nnz] of the 'plain' adjacency (values ignored, dgcf.py:90,
 * disengcn.py:27).  Factor count is 4, rows are 64-d = 4 chunks of 16; per-edge quantities are [nnz, 4] float
 * (one float4 per edge, factor-minor).  Node tables are [n, 64] row-major, dinv is [n, 4].
 * ---------------------------------------------------------------------------------------------------------- */
/* edge_row int32 [nnz] = row id of every CSR entry (the edge-parallel kernels are immune to hub rows).
 * w[e,:] = softmax_k(logit[e,:]) (dgcf.py:74); dinv[h,k] = 1/sqrt(sum_{e in row h} w[e,k]), 0 for empty rows
 * (dgcf.py:95-97). */
int tagrec_edge_softmax_rowsum(const int32_t* edge_row, int64_t nnz, int64_t n_rows, const float* logit, float* w,
                               float* dinv, void* stream);
/* val[e,k] = dinv[h,k] * w[e,k] * dinv[t,k]: the per-factor operator D A_k D of dgcf.py:98-101 as edge values. */
int tagrec_edge_scale(const int32_t* edge_row, const int32_t* col, int64_t nnz, const float* w, const float* dinv,
                      float* val, void* stream);
/* Long rows of a power-law graph (more than tagrec_spmm4_long_threshold() entries) are cut into pieces of at most
 * TAGREC_ROUTE_PIECE entries: long_rows int32 [n_long] row ids; piece_slot int32 / piece_begin, piece_end int64
 * [n_pieces] = (slot in long_rows, entry range); scratch float [n_long, 64], zero on entry, left zero on exit. */
#define TAGREC_ROUTE_PIECE 2048
typedef struct {
    const int32_t* long_rows;
    int64_t n_long;
    const int32_t* piece_slot;
    const int64_t* piece_begin;
    const int64_t* piece_end;
    int64_t n_pieces;
    float* scratch;
} tagrec_route_plan_t;

/* y[h, chunk k] = (res ? res[h] : 0) + sum_{e=(h,t)} val[perm ? perm[e] : e, k] * x[t, chunk k]
 *   y_raw  (optional) the sum;  y_norm (optional) each 16-d chunk L2-normalised (dgcf.py:79, disengcn.py:41);
 *   mean_acc (optional) running mean of the normalised layers: (first ? mean_x0 : mean_acc) + y_norm, times
 *   mean_scale when last (dgcf.py:59-61).  perm = reverse-edge permutation => multiplies by the TRANSPOSED operator.
 *   plan may be NULL when no row exceeds the threshold. */
int tagrec_spmm4(const int64_t* rowptr, const int32_t* col, int64_t n_rows, const tagrec_route_plan_t* plan,
                 const float* val, const int32_t* perm, const float* x, const float* res, float* y_raw, float* y_norm,
                 float* mean_acc, const float* mean_x0, int mean_first, int mean_last, float mean_scale, void* stream);
int tagrec_spmm4_long_threshold(void);
int tagrec_spmm4_piece(void);
/* d[e,k] = <a[h, chunk k], b[t, chunk k]>;  mode 0: out[e,:] += d (dgcf.py:103-109, A_values += A_score)
 *                                           mode 1: out[e,:] = softmax_k(d) (disengcn.py:31-34). */
int tagrec_edge_dot4(const int32_t* edge_row, const int32_t* col, int64_t nnz, const float* a, const float* b,
                     float* out, int mode, void* stream);
/* y = x / max(||x||_2 per 16-d chunk, 1e-12), optionally tanh(y) (dgcf.py:106-108). */
int tagrec_chunk_normalize(const float* x, int64_t n_rows, int apply_tanh, float* y, void* stream);
/* out = J^T g for y = chunk_normalize(x). */
int tagrec_chunk_normalize_bwd(const float* g, const float* x, int64_t n_rows, float* out, void* stream);
/* rev[e(h,t)] = e(t,h); *missing (device int) counts edges without a reverse (0 for a symmetric structure). */
int tagrec_csr_reverse_perm(const int64_t* rowptr, const int32_t* col, int64_t n_rows, int32_t* rev, int32_t* missing,
                            void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K4  TGCN neighbour attention      replaces model/tgcn.py:20-37 (Attention1.forward: [N,k,64] gathers, repeat,
 *                                   two matmuls, softmax over k, weighted sum) and its index_put_ backward.
 * The caller supplies the three dense projections (plain GEMMs):
 *     pv [n, 32] = e_v W1[:64] + b,   ww [n_w, 32] = e_w W1[64:],   pj [n_j, 32] = e_j W2
 * nbr / nbw: int64 [n, ld] neighbour tables (data/utils.py:87-106): entry = id + 1, 0 = padding (zero row, still a
 * softmax slot, tgcn.py:21-24); the first k columns are used (tgcn.py:199).  v [32] = attention vector.
 *   fwd: att[n,k] = softmax_k(relu(pv + ww[w-1] + pj[j-1]) . v);  out[n,64] = sum_k att * ej[j-1]
 *   bwd: g_pv (written), g_ww / g_pj / g_ej / g_v (ACCUMULATED: caller zeroes them).
 * ---------------------------------------------------------------------------------------------------------- */
int tagrec_nbr_attention_fwd(const float* pv, const float* ww, const float* pj, const float* ej, const float* v,
                             const int64_t* nbr, const int64_t* nbw, int64_t n, int k, int64_t ld, int dim,
                             int dim_atten, float* out, float* att, void* stream);
int tagrec_nbr_attention_bwd(const float* g_out, const float* att, const float* pv, const float* ww, const float* pj,
                             const float* ej, const float* v, const int64_t* nbr, const int64_t* nbw, int64_t n, int k,
                             int64_t ld, int n_w, int dim, int dim_atten, float* g_pv, float* g_ww, float* g_pj,
                             float* g_ej, float* g_v, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K7  TGCN dense tail              replaces model/tgcn.py:86-106 (BasicLayer._conv bit-level branch + _fusion) and
 *                                  their autograd: the [N, 2096] feature matrix is generated and consumed on chip.
 *   z  [n, 3, 64]   output of the type-level attention (tgcn.py:78-84)
 *   wb [C, 3]       conv.bit_level.weight[:, 0, :, 0]   (C = num_bit_conv)
 *   xf [n, E]       rectified vector-level features, cat of conv_1..3 channel-major (tgcn.py:92-98), E = 6*num_vec_conv
 *   wf [C*64+E, 64] fusion weight, bf [64] fusion bias
 *   fwd: out[n,64] = relu([relu(bit_conv(z)) | xf] wf + bf)
 *   bwd: g_z [n,3,64], g_xf [n,E] written; g_wb [C,3], g_wf, g_bf [64] OVERWRITTEN (zeroed inside, then reduced).
 *        workspace: tagrec_tgcn_tail_workspace_bytes(n, C) device bytes.
 * dim must be 64; E a multiple of 4, <= 64.
 * ---------------------------------------------------------------------------------------------------------- */
size_t tagrec_tgcn_tail_workspace_bytes(int64_t n, int n_bit_conv);
int tagrec_tgcn_tail_fwd(const float* z, const float* wb, const float* xf, const float* wf, const float* bf, int64_t n,
                         int dim, int n_bit_conv, int n_extra, float* out, void* stream);
/* same forward with the path chosen explicitly (TAGREC_EVAL_AUTO / _FP32 / _TF32): the tensor-core path runs the
 * [n, C*64+E] x [C*64+E, 64] product as 3xTF32 tcgen05 MMAs (fp32-level accuracy) with the features generated straight
 * into the swizzled operand tiles; it needs tagrec_tgcn_tail_fwd_workspace_bytes(C) device bytes.  AUTO = TF32 when a
 * workspace is given. */
size_t tagrec_tgcn_tail_fwd_workspace_bytes(int n_bit_conv);
int tagrec_tgcn_tail_fwd_ex(const float* z, const float* wb, const float* xf, const float* wf, const float* bf, int64_t n,
                            int dim, int n_bit_conv, int n_extra, float* out, void* workspace, size_t workspace_bytes,
                            int path, void* stream);
int tagrec_tgcn_tail_bwd(const float* g_out, const float* out, const float* z, const float* wb, const float* xf,
                         const float* wf, int64_t n, int dim, int n_bit_conv, int n_extra, void* workspace,
                         size_t workspace_bytes, float* g_z, float* g_wb, float* g_xf, float* g_wf, float* g_bf,
                         void* stream);
/* same with the path of the z-gradient pass chosen explicitly (AUTO = tensor cores, 3xTF32 tcgen05; FP32 = FMA kernel) */
int tagrec_tgcn_tail_bwd_ex(const float* g_out, const float* out, const float* z, const float* wb, const float* xf,
                         const float* wf, int64_t n, int dim, int n_bit_conv, int n_extra, void* workspace,
                         size_t workspace_bytes, float* g_z, float* g_wb, float* g_xf, float* g_wf, float* g_bf,
                         int path, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * K7a TGCN type attention + vector-level conv     replaces model/tgcn.py:78-84 (BasicLayer._atten2) and the
 *                                  vector-level branch of _conv (tgcn.py:92-98) with their autograd graphs.
 *   x0, x1, x2 [n, 64]  the (user, item, tag) slots of one node type: its own rows and its two neighbour-attention outputs
 *   U [64, 32], q [32], p [32]      type-attention parameters;  wv_j [V, j*64] = conv.vec_level.conv_j.weight, V in {4, 8}
 *   fwd: z [n,3,64] = softmax_r(relu(x_r U + q) . p) * x_r;   xf [n, 6V] = relu(conv_1..3(z)), channel-major per conv
 *   bwd: takes the forward's z and xf; g_x0..2 written; g_U, g_q, g_p, g_wv1..3 ACCUMULATED (caller zeroes them);
 *        workspace: tagrec_tgcn_mix_workspace_bytes(n) device bytes.
 * ---------------------------------------------------------------------------------------------------------- */
int tagrec_tgcn_mix_fwd(const float* x0, const float* x1, const float* x2, const float* U, const float* q,
                        const float* p, const float* wv1, const float* wv2, const float* wv3, int64_t n, int dim,
                        int dim_atten, int n_vec_conv, float* z, float* xf, void* stream);
int tagrec_tgcn_mix_bwd(const float* x0, const float* x1, const float* x2, const float* U, const float* q,
                        const float* p, const float* wv1, const float* wv2, const float* wv3, int64_t n, int dim,
                        int dim_atten, int n_vec_conv, const float* z, const float* g_z, const float* g_xf,
                        const float* xf, float* g_x0, float* g_x1, float* g_x2, float* g_U, float* g_q, float* g_p,
                        float* g_wv1, float* g_wv2, float* g_wv3, void* workspace, size_t workspace_bytes,
                        void* stream);
size_t tagrec_tgcn_mix_workspace_bytes(int64_t n);

/* ------------------------------------------------------------------------------------------------------------
 * K8  skinny X^T Y                 out[a, b] = x[n, a]^T y[n, b]  (OVERWRITTEN), a and b multiples of 4 in 4..64.
 * The weight gradient of the small dense layers on the path (attention projections model/tgcn.py:26-31, factor
 * projection model/disengcn.py:25): a row reduction spread over all SMs instead of a one-tile GEMM with K = n.
 * ---------------------------------------------------------------------------------------------------------- */
int tagrec_xty(const float* x, const float* y, int64_t n, int a, int b, float* out, void* stream);
/* same; accumulate != 0 adds to out instead of overwriting it */
int tagrec_xty_acc(const float* x, const float* y, int64_t n, int a, int b, float* out, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * BPR negative sampler          replaces train_data/bpr_training_data.py:29-45 + train_data/utils.py:19-28,52-55.
 * Host version: bit-exact numpy-legacy MT19937 stream for cpu_core == 1 (parity mode).  All pointers HOST.
 *   state: 625 uint32 (624 words + position), advanced exactly as the parent's RandomState is (shuffle only).
 * Device version: Philox counter RNG, rejection against the user's ascending train row (throughput mode).
 * ---------------------------------------------------------------------------------------------------------- */
void tagrec_mt19937_seed(uint32_t seed, uint32_t* state);
int tagrec_sample_bpr_host(uint32_t* state, const int64_t* edges /*[e,2]*/, int64_t e, const int64_t* train_ptr,
                           const int64_t* train_items_sorted, int64_t num_item, int64_t* triples_out /*[e,3]*/);
/* TransTag / TransE negative tails, host, bit-exact with train_data/transe_training_data.py:42-70 +
 * train_data/utils.py:31-37 (sample_neg_tail) for cpu_core == 1: for every (h, r, t) draw np.random.randint(0, num)
 * until the draw is not a tail of (h, r).  group[e] = id of the (h, r) pair of triple e; group_ptr / group_items_sorted =
 * CSR of the tails per pair (ascending).  `state` (625 words) is read, NOT advanced (the worker is a fork). */
int tagrec_sample_neg_tail_host(const uint32_t* state, const int64_t* group, int64_t e, const int64_t* group_ptr,
                                const int64_t* group_items_sorted, int64_t num, int64_t* neg_out);
int tagrec_sample_bpr_device(const int64_t* edges, int64_t e, const int64_t* train_ptr, const int32_t* train_items,
                             int64_t num_item, uint64_t seed, uint64_t epoch, int64_t* triples_out, void* stream);

/* TGCN neighbour tables on the device — replaces data/utils.py:87-106 (all_neighbor_sample) and
 * data/tgcn_load.py:41-53 (a Python loop over every row with matrix[i].toarray() + np.random.choice).
 * One relation per call: rows [row_begin, row_begin + n_rows) of the tripartite CSR (integer multiplicities in
 * `weight`), restricted to columns [col_lo, col_hi).  ids[i * width + s] = (neighbour column - col_lo) + 1, 0 = padding
 * (only in empty rows); rows shorter than `width` are filled by sampling WITH replacement, longer rows contribute a
 * uniformly random width-subset in random order; wts = the matching integer edge weights.  width <= 64.  Philox
 * keyed by (seed, relation, row): the reference's distribution, not numpy's stream. */
int tagrec_neighbor_table(const int64_t* rowptr, const int32_t* col, const float* weight, int64_t row_begin,
                          int64_t n_rows, int32_t col_lo, int32_t col_hi, int width, uint64_t seed, uint32_t relation,
                          int64_t* ids, int64_t* wts, void* stream);

/* Fused dense Adam (torch.optim.Adam semantics, com.py:25): one pass over param/grad/m/v. */
int tagrec_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int64_t step, void* stream);
/* Owner-sharded form (multi-GPU): the same step on a contiguous, 16-byte aligned segment (n % 4 == 0) of a parameter
 * table this rank owns; the new values are stored through `param_mirror` (bases point at the segment's first element in
 * every rank's replica, or one NVLS multicast address) so each row of the replicated table is updated by its owner
 * only; exp_avg / exp_avg_sq stay local.  The caller separates this launch from the next reader with a cross-rank
 * barrier. */
int tagrec_adam_step_mirror(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                            float beta1, float beta2, float eps, float weight_decay, int64_t step,
                            const tagrec_mirror_t* param_mirror, void* stream);

/* CUDA-graph-capturable form: the step counter and this step's bias corrections live in device memory.
 * tagrec_adam_advance: *step_dev += 1; scal_dev[0] = lr / (1 - beta1^step), scal_dev[1] = 1 / sqrt(1 - beta2^step)
 * (once per optimizer step); tagrec_adam_step_dev: the update of one tensor with those scalars. */
int tagrec_adam_advance(int64_t* step_dev, float lr, float beta1, float beta2, float* scal_dev, void* stream);
int tagrec_adam_step_dev(float* param, const float* grad, float* m, float* v, int64_t n, float beta1, float beta2,
                         float eps, float weight_decay, const float* scal_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAGREC_B200_H */
