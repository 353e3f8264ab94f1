/*
 * bignn_b200.h -- C-ABI of the B200-native Bi-GNN bi-level message-passing path.
 *
 * Drop-in boundary: the Python layer classes behind the reference's
 * model/layers_factory.py:179-190 `layer_ctors` registry call ONLY these entry
 * points (through ctypes).  Plain pointers and sizes, no torch
 * types.  Rules that hold for every function:
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - nothing is allocated inside: outputs and workspaces are caller-owned
 *     (query sizes with the *_workspace_bytes functions);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*), nothing
 *     synchronises, so every call is CUDA-graph capturable;
 *   - the return value is 0 on success, a positive cudaError_t, or a negative
 *     BIGNN_E* code for argument errors (bignn_error_string() explains both);
 *   - indices are int32 (CSR); int64 appears only in the COO `edge_index` view
 *     the reference API exposes (src/merged_graph.py:60);
 *   - features are fp32, row-major, leading dimension given in ELEMENTS.
 *
 * Each entry point cites the reference interface it replaces (paths relative
 * to the reference tree).
 */
#ifndef BIGNN_B200_H_
#define BIGNN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BIGNN_ABI_VERSION 4

/* argument errors */
#define BIGNN_EINVAL   (-1)   /* bad size / null pointer / unsupported flag */
#define BIGNN_EALIGN   (-2)   /* pointer or leading dimension not aligned as required */
#define BIGNN_EWORKSPACE (-3) /* workspace too small */

/* activations (model/layers_util.py:60-76 create_act) */
#define BIGNN_ACT_IDENTITY 0
#define BIGNN_ACT_RELU     1
#define BIGNN_ACT_SIGMOID  2
#define BIGNN_ACT_TANH     3

/* spmm modes */
#define BIGNN_SPMM_SUM   0  /* y_i = sum_{j in N(i)} x_j                                  */
#define BIGNN_SPMM_GIN   1  /* y_i = self_coef*x_i + sum_{j in N(i), j!=i} x_j            (PyG GINConv)  */
#define BIGNN_SPMM_GCN   2  /* y_i = sum_{j!=i} dinv_i dinv_j x_j + dinv_i^2 x_i (+bias)  (PyG GCNConv)  */

/* readout styles (model/layers_aggregation.py:14-19) */
#define BIGNN_READOUT_SUM  0
#define BIGNN_READOUT_MEAN 1

int bignn_abi_version(void);
const char* bignn_error_string(int code);
/* number of kernels launched by this library since load (bench.py gpu_launches) */
int64_t bignn_launch_count(void);

/* ---------------------------------------------------------------------------
 * Merged-batch / CSR construction on the device.
 * Replaces: src/batch.py:105-144 (_merge_into_one_graph), model/layers_util.py:100-166
 * (convert_nx_to_pyg_graph / create_edge_index, per graph per step on the host) and
 * src/merged_graph.py:27-96 (MergedGraphData.from_data_list).
 *
 * Packed dataset (resident in HBM, uploaded once): atom_ptr[N+1]; nbr_ptr[sumA+1]
 * and nbr_idx[nnz] = per-graph LOCAL neighbour ids, atoms ascending, neighbours
 * ascending (== the lexicographically sorted directed COO the reference builds);
 * x_all[sumA, F].
 * `rows[G]` = dataset rows of the graphs to merge, in merged order.
 * Outputs (sizes A = sum of atoms, E = sum of directed edges, both known to the
 * caller from its host copy of the pointers):
 *   seg_ptr[G+1]  node offset of graph g   (reference ind_list[g] = (seg_ptr[g], seg_ptr[g+1]))
 *   edge_ptr[G+1] edge offset of graph g   (reference edge_ind_list)
 *   row_ptr[A+1], col_idx[E]               merged CSR (== COO sorted by (row,col))
 *   batch[A]      graph id of every node   (reference `batch`, int32 here)
 *   x[A, F]       gathered features        (may be NULL to skip)
 *   edge_index_i64[2, E], batch_i64[A]     optional reference-typed views (NULL to skip)
 * ------------------------------------------------------------------------- */
int64_t bignn_merge_build_workspace_bytes(int32_t G);
int bignn_merge_build(const int32_t* atom_ptr, const int32_t* nbr_ptr, const int32_t* nbr_idx,
                      const float* x_all, int32_t F,
                      const int32_t* rows, int32_t G,
                      int32_t* seg_ptr, int32_t* edge_ptr,
                      int32_t* row_ptr, int32_t* col_idx, int32_t* batch,
                      float* x, int64_t* edge_index_i64, int64_t* batch_i64,
                      int32_t A, int32_t E,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Deterministic row-parallel segment SpMM (sub-warp per row, 128-bit gathers,
 * neighbours accumulated in ascending order, no atomics).
 * Replaces: PyG MessagePassing.propagate = index_select + torch-scatter scatter_add
 * reached from model/layers.py:52-54 (GINConv / GCNConv), and -- because every
 * graph on this path is symmetric -- its backward as well.
 * dinv (GCN mode) = (1 + #non-self neighbours)^-1/2 from bignn_gcn_dinv.
 * act is applied after bias (GCN forward: act(conv(x)) of model/layers.py:55).
 * ------------------------------------------------------------------------- */
int bignn_gcn_dinv(const int32_t* row_ptr, const int32_t* col_idx, int32_t n_rows,
                   float* dinv, void* stream);
int bignn_spmm_f32(const int32_t* row_ptr, const int32_t* col_idx,
                   const float* X, int64_t ldx, float* Y, int64_t ldy,
                   int32_t n_rows, int32_t D, int32_t mode, float self_coef,
                   const float* dinv, const float* bias, int32_t act, void* stream);

/* Row-partitioned form (multi-GPU upper level, edges partitioned by source drug): the CSR holds only this
 * rank's n_rows rows, whose node ids are row_offset .. row_offset + n_rows - 1 in the column / feature
 * index space; X and dinv cover ALL nodes, Y only the local rows. */
int bignn_spmm_rows_f32(const int32_t* row_ptr, const int32_t* col_idx,
                        const float* X, int64_t ldx, float* Y, int64_t ldy,
                        int32_t n_rows, int32_t row_offset, int32_t D, int32_t mode, float self_coef,
                        const float* dinv, const float* bias, int32_t act, void* stream);
/* n_big: the LAST n_big entries of multi_rows are hub rows (more than BIGNN_SPMM_BIG_ITEMS work items); their
 * partial sums are added by a whole CTA in a fixed two-level order instead of one sub-warp walking all of them
 * (a drug with 150 k interactions has ~4 900 items).  n_big = 0: every multi-item row is summed in item order. */
#define BIGNN_SPMM_BIG_ITEMS 64
int bignn_spmm_planned_rows_f32(const int32_t* row_ptr, const int32_t* col_idx,
                                const int32_t* item_ptr, const int32_t* item_row, int32_t n_items, int32_t seg,
                                const int32_t* multi_rows, int32_t n_multi, int32_t n_big,
                                const float* X, int64_t ldx, float* Y, int64_t ldy,
                                int32_t n_rows, int32_t row_offset, int32_t D, int32_t mode, float self_coef,
                                const float* dinv, const float* bias, int32_t act,
                                void* workspace, int64_t workspace_bytes, void* stream);

/* Long-row variant for skewed graphs (interaction graphs with hub drugs): the caller splits every
 * row into work items of at most `seg` neighbours -- item_ptr[n_rows+1] (items per row, prefix sum,
 * every row has >= 1 item), item_row[n_items], multi_rows[n_multi] = rows with more than one item.
 * Items are processed by independent sub-warps; rows with several items are finished by a second
 * kernel that adds the per-item partial sums in item order (deterministic).  Requires D % 4 == 0,
 * D <= 512 and 16-byte aligned operands. */
int64_t bignn_spmm_planned_workspace_bytes(int32_t n_items, int32_t D);
int bignn_spmm_planned_f32(const int32_t* row_ptr, const int32_t* col_idx,
                           const int32_t* item_ptr, const int32_t* item_row, int32_t n_items, int32_t seg,
                           const int32_t* multi_rows, int32_t n_multi,
                           const float* X, int64_t ldx, float* Y, int64_t ldy,
                           int32_t n_rows, int32_t D, int32_t mode, float self_coef,
                           const float* dinv, const float* bias, int32_t act,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Dense fp32 transforms:  C[M,N] = act(op(A)[M,K] * op(B)[K,N] + bias[N])
 * op(A) = A (ta=0, A is [M,K]) or A^T (ta=1, A is [K,M]); same for B.
 * Replaces: nn.Linear inside the GIN MLP (model/layers.py:26-30) and MLP
 * (model/layers_util.py:28-33), `x @ weight` of GCNConv/GATConv, and their
 * autograd backward (dX = dY W, dW = dY^T X with a deterministic split-K over the
 * row dimension).  fp32 FMA accumulation -- the 1e-5 parity target rules out
 * single-pass TF32.
 * ------------------------------------------------------------------------- */
int64_t bignn_gemm_workspace_bytes(int32_t M, int32_t N, int32_t K, int32_t ta);
int bignn_gemm_f32(int32_t ta, int32_t tb, int32_t M, int32_t N, int32_t K,
                   const float* A, int64_t lda, const float* B, int64_t ldb,
                   float* C, int64_t ldc, const float* bias, int32_t act,
                   void* workspace, int64_t workspace_bytes, void* stream);
/* Tensor-core variant for the tall-skinny transforms of the path (tcgen05.mma kind::tf32 with TMEM
 * accumulators, fp32-accurate through a 3xTF32 hi/lo split):  C[M,N] = act(A[M,K] * op(B) + bias),
 * B stored [N,K] (b_is_nk = 1, nn.Linear layout) or [K,N] (b_is_nk = 0, PyG layout).
 * Persistent kernel, two CTAs per SM, cp.async operand staging.  Requires N <= 128, K % 4 == 0,
 * K <= 64 (K <= 96 for N <= 64), lda % 4 == 0 and a 16-byte aligned A. */
int bignn_gemm_tc_supported(int32_t M, int32_t N, int32_t K);   /* 1 if the shape has a tensor-core kernel (N<=128, K<=64; K<=96 for N<=64) */
int bignn_gemm_tc_f32(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda,
                      const float* B, int64_t ldb, int32_t b_is_nk,
                      float* C, int64_t ldc, const float* bias, int32_t act, void* stream);
/* The same GEMM with a fused activation backward in the epilogue:  C = act(A * op(B) + bias) * mask_act'(mask_y),
 * mask_y [M,N] being the OUTPUT of activation mask_act (relu: mask_y > 0).  Used for the backward-input transform
 * dT = (g W2) * relu'(t) of the GIN MLP (model/layers.py:27-29: Linear, act, Linear), which saves the separate
 * elementwise pass of autograd's ReluBackward.  mask_y = NULL: identical to bignn_gemm_tc_f32. */
int bignn_gemm_tc_masked_f32(int32_t M, int32_t N, int32_t K, const float* A, int64_t lda,
                             const float* B, int64_t ldb, int32_t b_is_nk,
                             float* C, int64_t ldc, const float* bias, int32_t act,
                             const float* mask_y, int64_t ldmy, int32_t mask_act, void* stream);
/* Weight gradients on the tensor cores:  D[Np,Nq] = P[M,Np]^T * Q[M,Nq]  (row-major, ld = Nq) and,
 * optionally, the column sums of P (colsum_of = 0) or Q (colsum_of = 1) -> colsum[] (bias gradient);
 * colsum_of = -1 skips them.  3xTF32 on tcgen05 with MN-major operands, per-CTA partials summed in a
 * fixed order (deterministic).  Requires Np, Nq <= 64 and multiples of 4, ld % 4 == 0, 16-byte
 * aligned operands. */
int bignn_dw_tc_supported(int32_t M, int32_t Np, int32_t Nq);
int64_t bignn_dw_tc_workspace_bytes(int32_t M, int32_t Np, int32_t Nq);
int bignn_dw_tc_f32(int32_t M, int32_t Np, int32_t Nq, const float* P, int64_t ldp,
                    const float* Q, int64_t ldq, float* D, int32_t colsum_of, float* colsum,
                    void* workspace, int64_t workspace_bytes, void* stream);
/* column sums out[c] = sum_r X[r,c]  (bias gradients), deterministic */
int64_t bignn_colsum_workspace_bytes(int32_t rows, int32_t cols);
int bignn_colsum_f32(const float* X, int64_t ldx, int32_t rows, int32_t cols, float* out,
                     void* workspace, int64_t workspace_bytes, void* stream);
/* dX = dY * act'(Y) elementwise from the activation OUTPUT Y (relu/sigmoid/tanh/identity) */
int bignn_act_bwd_f32(const float* Y, const float* dY, float* dX, int64_t n, int32_t act,
                      void* stream);

/* ---------------------------------------------------------------------------
 * Segmented train-mode BatchNorm1d: rows [seg_row_ptr[s], seg_row_ptr[s+1]) form
 * one independent batch (one 128-graph chunk of the all-drug pass, src/train.py:62-71),
 * so a whole all-drug pass is normalised in one launch with the reference's
 * per-chunk statistics.  Replaces torch.nn.BatchNorm1d at model/layers.py:57.
 * Statistics are accumulated in fp64, biased variance for normalisation;
 * running buffers are updated sequentially in segment order with the unbiased
 * variance (momentum, as torch).  mean/rstd are [S, C] and are what backward needs.
 * ------------------------------------------------------------------------- */
int64_t bignn_bn_workspace_bytes(int32_t S, int32_t C, int32_t parts);
int bignn_bn_seg_fwd(const float* X, int64_t ldx, float* Y, int64_t ldy,
                     const int32_t* seg_row_ptr, int32_t S, int32_t C, int32_t parts,
                     const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked,
                     float* mean, float* rstd, double* seg_stats_out,
                     void* workspace, int64_t workspace_bytes, void* stream);
/* seg_stats_out (optional, [2, S, C] fp64): per-segment batch mean and UNBIASED variance.  When the
 * chunks of one all-drug pass are sharded over several GPUs, each rank passes running_mean = NULL,
 * the ranks exchange these statistics, and every rank replays the reference's sequential momentum
 * updates (one per chunk, in chunk order) with bignn_bn_running_update. */
int bignn_bn_running_update(const double* seg_stats, const int32_t* seg_row_ptr, int32_t S, int32_t C,
                            float momentum, float* running_mean, float* running_var,
                            int64_t* num_batches_tracked, void* stream);
/* eval mode: normalise with the running buffers (one launch, no statistics) */
int bignn_bn_eval_fwd(const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t rows, int32_t C,
                      const float* gamma, const float* beta, float eps,
                      const float* running_mean, const float* running_var, void* stream);
/* dX, and dgamma/dbeta ACCUMULATED over segments in segment order (written, not added).
 * input_act (BIGNN_ACT_*): X is the output of that activation (model/layers.py:55-57 applies act, then bn); its
 * derivative is folded into dX, so dX is the gradient w.r.t. the activation's INPUT (BIGNN_ACT_IDENTITY = plain
 * BatchNorm backward).  Saves the separate elementwise pass over [rows, C].
 * From 24 segments of 64 channels on (the all-drug lower level) one thread-block cluster per segment does both passes
 * while the segment is L2-resident (csrc/bn.cu k_bn_bwd_chunk; same results bit for bit, 2 launches instead of 4). */
int bignn_bn_seg_bwd(const float* X, int64_t ldx, const float* dY, int64_t lddy,
                     float* dX, int64_t lddx,
                     const int32_t* seg_row_ptr, int32_t S, int32_t C, int32_t parts,
                     const float* gamma, const float* mean, const float* rstd,
                     float* dgamma, float* dbeta, int32_t input_act,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* Row-partitioned BatchNorm (multi-GPU upper level, SURVEY 8e: interaction-graph rows partitioned by
 * source drug, one BatchNorm batch = the rows of ALL ranks).  Three steps around one all-reduce that the
 * caller issues (NCCL, [2, C] fp64):
 *   bignn_bn_rows_sums       local (sum x, sum x^2) -- or, with dY != NULL, (sum dy, sum dy*xhat) -- over this
 *                            rank's `rows` rows, fp64, fixed part order;
 *   bignn_bn_rows_fwd_apply  mean / rstd of the whole batch from the rank-summed sums and n_total rows, one
 *                            momentum update of the (replicated) running buffers, y = bn(x) on the local rows;
 *   bignn_bn_rows_bwd_apply  dX on the local rows from the rank-summed backward sums.
 * The local backward sums ARE this rank's partial dbeta / dgamma.  With one rank the results equal
 * bignn_bn_seg_fwd / bwd with S = 1. */
int64_t bignn_bn_rows_workspace_bytes(int32_t C, int32_t parts);
int bignn_bn_rows_sums(const float* X, int64_t ldx, const float* dY, int64_t lddy, int32_t rows, int32_t C,
                       int32_t parts, const float* mean, const float* rstd, double* sums,
                       void* workspace, int64_t workspace_bytes, void* stream);
int bignn_bn_rows_fwd_apply(const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t rows, int32_t C,
                            int32_t parts, const double* sums, int64_t n_total,
                            const float* gamma, const float* beta, float eps, float momentum,
                            float* running_mean, float* running_var, int64_t* num_batches_tracked,
                            float* mean, float* rstd, void* stream);
int bignn_bn_rows_bwd_apply(const float* X, int64_t ldx, const float* dY, int64_t lddy, float* dX, int64_t lddx,
                            int32_t rows, int32_t C, int32_t parts, const float* gamma,
                            const float* mean, const float* rstd, const double* sums, int64_t n_total,
                            int32_t input_act, void* stream);

/* ---------------------------------------------------------------------------
 * Fused gather-and-score edge decoder + loss head (SURVEY 8b `pair_decoder_fwd/bwd`).
 * Replaces model/layers_link_pred.py:43-65 (F.normalize, gather of the two embedding rows of every pair, concat,
 * the mlp_concat MLP, sigmoid) and model/layers.py:79-89 (BCELoss / BCEWithLogitsLoss / CrossEntropyLoss, mean):
 * forward ONE launch, backward ONE launch (+ bignn_spmm_f32 in SUM mode over the entry CSR, the transpose of the
 * gather, to add the per-entry row gradients per drug).
 * ids [P,2] rows of H [*, ldh] (D columns).  The scorer has n_layers = 2 or 3 Linear layers (nn.Linear layout
 * W_l [n_l, n_{l-1}], n_0 = 2D), ReLU after all but the last.  head: 0 = sigmoid + BCE (y float [P]),
 * 1 = logits + BCEWithLogits (y), 2 = logits + cross entropy (labels int32 [P]).  Limits: 2D <= 256, n1 <= 16, later
 * widths <= 32 (bignn_pair_decoder_supported).  scores [P, lds] receives the LinkPred output (probabilities for head 0,
 * logits otherwise); nrm [P,2], h1 [P,n1], h2 [P,n2] are kept for the backward; loss (optional) the mean loss.
 * workspace: bignn_pair_decoder_workspace_bytes, ZERO-INITIALISED ONCE by the caller and then reused across calls (it
 * holds the completion counter of the deterministic last-CTA reduction, which the kernels reset themselves).
 * ------------------------------------------------------------------------- */
int bignn_pair_decoder_supported(int32_t D, int32_t n_layers, int32_t n1, int32_t n2, int32_t n3);
int64_t bignn_pair_decoder_workspace_bytes(int32_t P, int32_t D, int32_t n_layers, int32_t n1, int32_t n2, int32_t n3);
int bignn_pair_decoder_fwd(const float* H, int64_t ldh, const int32_t* ids, int32_t P, int32_t D, int32_t n_layers,
                           const float* W0, const float* b0, int32_t n1, const float* W1, const float* b1, int32_t n2,
                           const float* W2, const float* b2, int32_t n3, int32_t head,
                           const float* y, const int32_t* labels, float* scores, int64_t lds,
                           float* nrm, float* h1, float* h2, float* loss,
                           void* workspace, int64_t workspace_bytes, void* stream);
int bignn_pair_decoder_bwd(const float* H, int64_t ldh, const int32_t* ids, int32_t P, int32_t D, int32_t n_layers,
                           const float* W0, int32_t n1, const float* W1, int32_t n2, const float* W2, int32_t n3,
                           int32_t head, const float* y, const int32_t* labels,
                           const float* scores, int64_t lds, const float* nrm, const float* h1, const float* h2,
                           const float* dloss, float* drows, int64_t lddr,
                           float* dW0, float* db0, float* dW1, float* db1, float* dW2, float* db2,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * One lower-level GIN layer as ONE launch (SURVEY 8b `gin_layer_fwd`).
 * Replaces model/layers.py:42-57 with type='gin': PyG GINConv (index_select + scatter_add, Linear, act,
 * Linear) -> act -> the statistics pass of BatchNorm1d; the BatchNorm *apply* of this layer is not a pass at
 * all -- it is folded (centred) into the aggregation of the layer that consumes it, and into the readout:
 *     z_i = fold_a[c] * ((1+eps)(x_i - mean[c]) + sum_{j in N(i)} (x_j - mean[c])) + beta * (1 + eps + deg_i)
 *     t   = act_inner(z W1^T + b1);   Y = act_outer(t W2^T + b2)          (Y = the BatchNorm's INPUT)
 * with c = chunk of row i.  X [rows, ldx]: the producer layer's Y (or the raw features, fold_a = NULL); rows
 * zero-padded to a multiple of 4 columns.  W1 [64, din], W2 [64, 64] contiguous (nn.Linear layout).  dout must
 * be 64, din <= 64.  nnz = entries of col_idx; tile_edge_ptr [ceil(rows/128)+1] = row_ptr[128 t] (row_ptr[rows]
 * last).  chunk_row_ptr [S+1]: row ranges that are independent BatchNorm batches (the 128-graph chunks of
 * src/train.py:62-71; no edge crosses a chunk boundary); tile_chunk0 [ceil(rows/128)]: chunk of the first row
 * of every 128-row tile.  fold_mean / fold_a [S, din], fold_beta [din].
 * Z [rows, ldz] (aggregated input) and T [rows, ldt] (hidden activations) are optional outputs for the backward.
 * stat_parts (optional): fp64 [bignn_gin_layer_stat_records(rows, S), 2, 64] partial column sums / sums of
 * squares of Y per (tile, chunk), which bignn_gin_bn_finalize adds in tile order (deterministic) into
 * mean / rstd [S, 64], seg_stats_out [2, S, 64] (mean, unbiased variance: the input of
 * bignn_bn_running_update) and fold_a = gamma * rstd for the consumer.
 * ------------------------------------------------------------------------- */
int bignn_gin_layer_supported(int32_t din, int32_t dout);
int64_t bignn_gin_layer_stat_records(int32_t rows, int32_t S);
int bignn_gin_layer_fwd(int32_t rows, int32_t din, int32_t dout,
                        const int32_t* row_ptr, const int32_t* col_idx, int32_t nnz, const int32_t* tile_edge_ptr,
                        const float* X, int64_t ldx,
                        const float* fold_mean, const float* fold_a, const float* fold_beta,
                        const int32_t* chunk_row_ptr, int32_t S, const int32_t* tile_chunk0,
                        float self_coef, const float* W1, const float* b1, const float* W2, const float* b2,
                        int32_t act_inner, int32_t act_outer,
                        float* Z, int64_t ldz, float* T, int64_t ldt, float* Y, int64_t ldy,
                        double* stat_parts, void* stream);
int bignn_gin_bn_finalize(const double* stat_parts, const int32_t* chunk_row_ptr, int32_t S, int32_t C,
                          float eps, const float* gamma, float* mean, float* rstd,
                          double* seg_stats_out, float* fold_a, void* stream);

/* ---------------------------------------------------------------------------
 * Segment readout (atoms -> one row per drug), rows summed in ascending order.
 * Replaces torch-scatter scatter_mean / scatter_add at
 * model/layers_aggregation.py:17-19,34-41 and the per-row Python scatter into
 * init_x at :70-74 (dst_row[g] = gs_map row; NULL = identity).
 * out[dst_row[g], col_off : col_off+D] = pool(X[seg_ptr[g]:seg_ptr[g+1], :]).
 * ------------------------------------------------------------------------- */
int bignn_readout_fwd(const float* X, int64_t ldx, const int32_t* seg_ptr, int32_t G, int32_t D,
                      int32_t style, const int32_t* dst_row,
                      float* out, int64_t ldo, int32_t col_off, void* stream);
/* the same with the BatchNorm affine of the pooled activations folded in (the input is the BatchNorm's INPUT):
 * pool(a (x - mean) + beta) = a * pool(x - mean) + beta * (1 for mean, n for sum);
 * fold_mean / fold_a [S, D], fold_beta [D], graph_chunk [G]. */
int bignn_readout_fold_fwd(const float* X, int64_t ldx, const int32_t* seg_ptr, int32_t G, int32_t D,
                           int32_t style, const int32_t* dst_row, const float* fold_mean, const float* fold_a,
                           const float* fold_beta, const int32_t* graph_chunk, float* out, int64_t ldo,
                           int32_t col_off, void* stream);
/* gated ("attention") readout, model/layers_aggregation.py:90-94 (GMNAggregatorPairs): out[g] = sum over the atoms of
 * graph g of sigmoid(gate) * weight -- product and sum in one launch; backward (dGate, dWeight) in one launch.  D % 4 == 0. */
int bignn_readout_gated_fwd(const float* gate, int64_t ldg, const float* weight, int64_t ldw,
                            const int32_t* seg_ptr, int32_t G, int32_t D, float* out, int64_t ldo, void* stream);
int bignn_readout_gated_bwd(const float* gate, int64_t ldg, const float* weight, int64_t ldw,
                            const float* dOut, int64_t ldo, const int32_t* seg_ptr, int32_t G, int32_t D,
                            float* dGate, int64_t lddg, float* dWeight, int64_t lddw, void* stream);
int bignn_readout_bwd(const float* dOut, int64_t ldo, int32_t col_off, const int32_t* dst_row,
                      const int32_t* seg_ptr, int32_t G, int32_t D, int32_t style,
                      float* dX, int64_t lddx, int32_t accumulate, void* stream);

/* ---------------------------------------------------------------------------
 * Pair decoder front end: Z[p, 0:D] = H[id1[p]]/max(|H[id1[p]]|,1e-12), Z[p, D:2D]
 * likewise for id2 (F.normalize + gather + concat of model/layers_link_pred.py:44-61;
 * only the gathered rows are normalised -- same values).  inv_norm[P,2] is saved
 * for backward.  Backward produces per-entry row gradients dRows[2P, D]
 * (entry e = 2p+side) which the caller sums per drug with bignn_spmm_f32(SUM)
 * over the entry CSR -- deterministic, no atomics.
 * ------------------------------------------------------------------------- */
int bignn_pair_gather_norm_fwd(const float* H, int64_t ldh, const int32_t* ids /*[P,2]*/,
                               int32_t P, int32_t D, float* Z, int64_t ldz, float* inv_norm,
                               void* stream);
int bignn_pair_gather_norm_bwd(const float* H, int64_t ldh, const int32_t* ids, int32_t P, int32_t D,
                               const float* dZ, int64_t lddz, const float* inv_norm,
                               float* dRows, int64_t lddr, void* stream);

/* ---------------------------------------------------------------------------
 * Loss heads (model/layers.py:66-89): nn.BCELoss (mean, log clamped at -100) on the
 * sigmoid outputs of LinkPred (model/layers_link_pred.py:65; the sigmoid itself is
 * the activation epilogue of the last decoder GEMM), and nn.BCEWithLogitsLoss.
 * *loss is a device scalar; bwd takes the upstream scalar gradient *dloss (device).
 * ------------------------------------------------------------------------- */
int bignn_bce_fwd(const float* pred, const float* y, int32_t P, float* loss, void* stream);
int bignn_bce_bwd(const float* pred, const float* y, int32_t P, const float* dloss, float* dpred,
                  void* stream);
int bignn_bce_logits_fwd(const float* x, const float* y, int32_t P, float* loss, void* stream);
int bignn_bce_logits_bwd(const float* x, const float* y, int32_t P, const float* dloss, float* dx,
                         void* stream);

/* ---------------------------------------------------------------------------
 * One-head GAT edge-softmax message passing (PyG 1.1.2 GATConv reached from
 * model/layers.py:32-34,54).  H = x W is computed by bignn_gemm_f32.
 * att[2D] = [att_i ; att_j] (target part, source part).  group_target = 0 groups the
 * softmax by the SOURCE node (torch-geometric 1.1.x), 1 by the TARGET (>= 1.2).
 * scratch4n[4n] floats (p, q, max, sum) are produced by fwd and consumed by bwd.
 * bwd returns dH and dpq[2n] = (d p, d q); d att = [dp^T H ; dq^T H] (a GEMM).
 * ------------------------------------------------------------------------- */
/* Every neighbour pass runs over the work items of a row plan (see bignn_spmm_planned_f32: item_ptr,
 * item_row, multi_rows; rows of one item are finished in place, hub rows through per-item partials
 * combined in item order).  D <= 64, D % 4 == 0, 16-byte aligned rows.
 * Several edge types (model/layers_meta.py:61-79) are batched as ONE block-diagonal graph of
 * n / n_block blocks of n_block nodes: block b uses att[b, 2D] and bias[b, D] (n_block = n for a
 * single graph). */
int64_t bignn_gat_fwd_workspace_bytes(int32_t n_items, int32_t D);
int bignn_gat_fwd(const int32_t* row_ptr, const int32_t* col_idx,
                  const int32_t* item_ptr, const int32_t* item_row, int32_t n_items, int32_t seg,
                  const int32_t* multi_rows, int32_t n_multi, int32_t n, int32_t n_block, int32_t D,
                  const float* H, int64_t ldh, const float* att, const float* bias,
                  float negative_slope, int32_t group_target, float* out, int64_t ldo,
                  float* scratch4n, void* workspace, int64_t workspace_bytes, void* stream);
int64_t bignn_gat_bwd_workspace_bytes(int32_t n, int32_t D, int32_t n_items);
int bignn_gat_bwd(const int32_t* row_ptr, const int32_t* col_idx,
                  const int32_t* item_ptr, const int32_t* item_row, int32_t n_items, int32_t seg,
                  const int32_t* multi_rows, int32_t n_multi, int32_t n, int32_t n_block, int32_t D,
                  const float* H, int64_t ldh, const float* att, const float* bias,
                  float negative_slope, int32_t group_target, const float* out, int64_t ldo,
                  const float* dOut, int64_t lddo, const float* scratch4n,
                  float* dH, int64_t lddh, float* dpq,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Elementwise / row-wise operators.
 *   act_fwd            standalone activation (MLP with BatchNorm: act(bn(linear(x))),
 *                      model/layers_util.py:51-52)
 *   prelu_fwd/bwd      nn.PReLU (create_act 'prelu', model/layers_util.py:63-64); bwd also
 *                      writes T = dY * min(x,0)/x-part whose column sums are d slope
 *   rownorm_fwd/bwd    F.normalize(p=2, dim=1) of NodeEmbedding normalize=True (model/layers.py:60-61)
 *   gate_mul_fwd/bwd   sigmoid(gate) * weight of the gmn_aggr readout (model/layers_aggregation.py:90-93)
 *   pair_dot_fwd/bwd   dot-product scorer sigmoid(<g1, g2>) (model/layers_link_pred.py:66-67)
 *   ce_fwd/bwd         nn.CrossEntropyLoss, mean (model/layers.py:75,85-88); labels int32
 * ------------------------------------------------------------------------- */
int bignn_act_fwd_f32(const float* X, float* Y, int64_t n, int32_t act, void* stream);
/* O = A + B: the sum over edge types of NodeModelAggrByEdge (model/layers_meta.py:74-79) */
int bignn_add_f32(const float* A, const float* B, float* O, int64_t n, void* stream);
int bignn_prelu_fwd_f32(const float* X, float* Y, int64_t rows, int32_t C, const float* w, int32_t nw,
                        void* stream);
int bignn_prelu_bwd_f32(const float* X, const float* dY, float* dX, float* T, int64_t rows, int32_t C,
                        const float* w, int32_t nw, void* stream);
int bignn_rownorm_fwd_f32(const float* X, int64_t ldx, float* Y, int64_t ldy, int32_t rows, int32_t D,
                          float* nrm, void* stream);
int bignn_rownorm_bwd_f32(const float* Y, int64_t ldy, const float* dY, int64_t lddy, const float* nrm,
                          float* dX, int64_t lddx, int32_t rows, int32_t D, void* stream);
int bignn_gate_mul_fwd_f32(const float* G, const float* W, float* O, int64_t n, void* stream);
int bignn_gate_mul_bwd_f32(const float* G, const float* W, const float* dO, float* dG, float* dW,
                           int64_t n, void* stream);
int bignn_pair_dot_fwd_f32(const float* Z, int64_t ldz, int32_t P, int32_t D, float* out, int32_t act,
                           void* stream);
int bignn_pair_dot_bwd_f32(const float* Z, int64_t ldz, int32_t P, int32_t D, const float* out,
                           const float* dout, int32_t act, float* dZ, int64_t lddz, void* stream);
int bignn_ce_fwd(const float* logits, int64_t ldx, const int32_t* labels, int32_t P, int32_t K,
                 float* loss, void* stream);
int bignn_ce_bwd(const float* logits, int64_t ldx, const int32_t* labels, int32_t P, int32_t K,
                 const float* dloss, float* dlogits, int64_t lddx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BIGNN_B200_H_ */
