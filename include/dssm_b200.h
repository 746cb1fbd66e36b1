/*
 * dssm_b200.h -- C ABI of the B200-native DSSM two-tower hot path.
 *
 * The reference (MC-Zealot/dssm) has no FFI: its "interface" is a flat TensorFlow-1.x script
 * (semantic_matching/dssm/new_dssm.py) whose arithmetic lives in TF op kernels.  Every entry point
 * below therefore names the reference *call site* (file:line under /root/reference) whose TF ops it
 * replaces.  INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions:
 *   - extern "C", plain pointers and sizes; all data pointers are DEVICE pointers unless the
 *     parameter name starts with `host_`.
 *   - every call returns int: 0 = DSSM_OK, otherwise one of DSSM_ERR_*; dssm_last_error() gives the
 *     message of the last failure on the calling thread.
 *   - no hidden device allocation: ops that need scratch take (workspace, workspace_bytes) and have a
 *     *_workspace_bytes() query.  Workspaces must be 256-byte aligned.
 *   - every call takes a stream (a cudaStream_t passed as void*); nothing synchronises the host
 *     except the *_host entry points, which say so.
 *   - a dssm_tower handle is not thread-safe: one handle per device/stream.
 *   - float tensors are fp32 row-major; index tensors are int32.
 *
 * Row convention for a batch (utils/utils.py:45-61, pull_batch): the three CSR slices are stacked
 *   rows [0,B)            query_batch
 *   rows [B,2B)           doc_positive_batch
 *   rows [2B,2B+B*NEG)    doc_negative_batch  (negatives of query j are rows 2B + j*NEG .. +NEG-1)
 * R = (2+NEG)*B rows in total.  "Segment" below: q = rows [0,B), d = rows [B,R)  (the two
 * batch_normalization instances of new_dssm.py:129-130 / :151-152).
 */
#ifndef DSSM_B200_H_
#define DSSM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSSM_B200_VERSION 100

typedef void* dssm_stream_t; /* cudaStream_t */

enum {
    DSSM_OK = 0,
    DSSM_ERR_BAD_ARG = 1,   /* null pointer / negative size / unknown enum */
    DSSM_ERR_BAD_SHAPE = 2, /* shape not supported by the kernels */
    DSSM_ERR_BAD_ALIGN = 3, /* pointer not 16-byte aligned where vector access needs it */
    DSSM_ERR_CUDA = 4,      /* a CUDA runtime call failed; message holds cudaGetErrorString */
    DSSM_ERR_WORKSPACE = 5, /* workspace too small */
    DSSM_ERR_STATE = 6      /* tower used before bind / backward before forward ... */
};

enum { DSSM_ACT_NONE = 0, DSSM_ACT_RELU = 1, DSSM_ACT_TANH = 2 };

/* Arithmetic of the dense-layer contractions (FC2.. and their gradients).
 *   FP32      : FFMA, fp32 accumulate.
 *   TC_3XTF32 : tcgen05.mma kind::tf32 with error-compensated operands (x = hi + lo, three MMAs per product,
 *               fp32 accumulate in TMEM): tensor-core path that keeps the 1e-5 parity bar.
 *   TC_TF32   : the same kernels with ONE kind::tf32 MMA per product (operands rounded to tf32, fp32 accumulate): a third
 *               of the tensor work; does NOT hold 1e-5 -- measured against the oracle at C2: loss within 5e-4 relative,
 *               embeddings / cosines within 2e-3 of their scale (tests/test_gpu_tower.py::test_tf32_mode_tolerance). */
enum { DSSM_GEMM_FP32 = 0, DSSM_GEMM_TC_3XTF32 = 1, DSSM_GEMM_TC_TF32 = 2 };

#define DSSM_MAX_LAYERS 8
#define DSSM_MAX_PEERS 16 /* ranks addressable by the NVLink peer-memory exchange */

/* Hyper-parameters; names follow semantic_matching/dssm/config.py:19-28 and new_dssm.py:44. */
typedef struct dssm_config {
    int32_t TRIGRAM_D;               /* input vocabulary size (new_dssm.py:44) */
    int32_t n_layers;                /* 2 in the reference graph (L1_N, L2_N); 3 for 300-300-128 */
    int32_t layers[DSSM_MAX_LAYERS]; /* L1_N, L2_N, ... */
    int32_t NEG;                     /* config.py:28 */
    int32_t query_BS;                /* config.py:19 */
    int32_t use_bn;                  /* 1: semantic_matching/dssm, 0: semantic_matching/dssm_no_bn */
    int32_t act;                     /* DSSM_ACT_RELU (reference) or DSSM_ACT_TANH */
    int32_t loss_div_bs;             /* 1: new_dssm.py:209, 0: archive/dssm_v2.py:184 */
    int32_t gemm_mode;               /* DSSM_GEMM_* */
    float bn_eps;                    /* 1e-3, new_dssm.py:87 */
    float ema_decay;                 /* 0.5, new_dssm.py:78 */
    float gamma;                     /* 20, new_dssm.py:199 */
    float loss_eps;                  /* 0 (new_dssm.py:209) or 1e-8 (dssm_no_bn/my_dssm.py:169) */
    float learning_rate;             /* config.py:23 */
    float beta1, beta2, adam_eps;    /* tf.train.AdamOptimizer defaults .9 .999 1e-8 (new_dssm.py:217) */
} dssm_config;

const char* dssm_last_error(void);
int dssm_version(void);
/* Number of SMs of the current device (grid sizing); <0 on error. */
int dssm_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * FC1: word-hashed sparse input layer.   Replaces tf.sparse_tensor_dense_matmul(x, weight1) + bias1
 * (new_dssm.py:124-126) for the three stacked inputs at once.
 *   Y[r,:] = sum_{p in [indptr[r],indptr[r+1])} values[p] * W1[indices[p],:]  (+ b1)
 * CSR must be canonical (sorted, duplicate-free columns per row).  b1 may be NULL.
 * W1/b1/Y must be 16-byte aligned when L1 % 4 == 0 (vector path).
 */
int dssm_spmm_fwd(const int32_t* indptr, const int32_t* indices, const float* values, int32_t R, int32_t D,
                  const float* W1, const float* b1, int32_t L1, float* Y, dssm_stream_t stream);

/* Gradient of FC1 w.r.t. weight1: the dense [D,L1] tensor  dW1 = X^T dH  that TF's
 * SparseTensorDenseMatMul gradient (adjoint_a=True) produces for new_dssm.py:124-126, summed over
 * the three tower applications.  Every row of dW1 is written (rows of absent columns are zero).
 * method 0: per-batch CSC built on device + gather-by-column (each dW1 row written once),
 * method 1: zero-fill + red.global.add.v4.f32 scatter (kept as the cross-check). */
size_t dssm_spmm_bwd_dw_workspace_bytes(int32_t R, int32_t D, int32_t L1, int64_t max_nnz);
int dssm_spmm_bwd_dw(const int32_t* indptr, const int32_t* indices, const float* values, int32_t R, int32_t D,
                     const float* dH, int32_t L1, float* dW1, int32_t method, void* workspace,
                     size_t workspace_bytes, dssm_stream_t stream);

/* The two halves of method 0, for callers that overlap the gradient exchange with the gather (data-parallel
 * training): build the per-batch CSC once, then produce dW1 rows [col_begin, col_end) range by range (each range
 * a contiguous slice of dW1; `chunk` < 64 must differ between ranges issued after one build).  L1 % 4 == 0.
 * The build zero-fills dW1; the ranges then write the rows of the columns present in the batch. */
int dssm_spmm_bwd_csc_build(const int32_t* indptr, const int32_t* indices, const float* values, int32_t R, int32_t D,
                            int32_t L1, float* dW1 /* zero-filled here: rows of absent columns get no other write */,
                            void* workspace, size_t workspace_bytes, dssm_stream_t stream);
int dssm_spmm_bwd_dw_range(const float* dH, int32_t R, int32_t D, int32_t L1, float* dW1, int32_t col_begin, int32_t col_end,
                           int32_t chunk, void* workspace, size_t workspace_bytes, dssm_stream_t stream);

/* The gather fused with tf.train.AdamOptimizer on weight1 (new_dssm.py:217) for the single-GPU step: after
 * dssm_spmm_bwd_csc_build(dW1 = NULL), every dW1 row is consumed in registers and W1 / m / v are updated in place;
 * dW1 itself is not produced.  Columns absent from the batch get the zero-gradient update (dense-Adam semantics),
 * either here (absent_done = 0) or, earlier and off the critical path, by dssm_spmm_bwd_adam_absent (absent_done = 1):
 * that call needs only the CSC's column histogram and touches rows neither dssm_spmm_fwd nor the backward of the
 * same batch reads, so it may run on another stream beside them, ordered after dssm_spmm_bwd_csc_build. */
int dssm_spmm_bwd_dw_adam(const float* dH, int32_t R, int32_t D, int32_t L1, float* W1, float* m1, float* v1,
                          const float* beta_pow, float lr, float beta1, float beta2, float eps, int32_t absent_done,
                          void* workspace, size_t workspace_bytes, dssm_stream_t stream);
int dssm_spmm_bwd_adam_absent(int32_t D, int32_t L1, float* W1, float* m1, float* v1, const float* beta_pow, float lr,
                              float beta1, float beta2, float eps, void* workspace, size_t workspace_bytes,
                              dssm_stream_t stream);

/* Data-parallel exchange of dW1 over NVLink peer memory, fused with Adam (no reference counterpart; SURVEY 8e's
 * reduce-scatter -> sharded Adam -> all-gather as one kernel).  peer_dW1[r] / peer_W1[r] (HOST arrays of n_ranks device
 * pointers, r = rank order, entry `self` being this rank's own buffers) address every rank's dense dW1 [D,L1] and W1
 * [D,L1]; they must be peer-mapped (e.g. torch symmetric memory).  For the rows [row_begin,row_end) this rank owns, the
 * kernel pulls the gradient row from every rank, sums in rank order, divides by n_ranks, applies TF-Adam with the local
 * m1, v1 (full-size [D,L1] buffers; only the owned rows are touched) and writes the new weight row into every rank's W1.
 * Ordering is the caller's: a cross-device barrier after all ranks produced dW1, another after this call before W1 is
 * read again.  beta_pow is read, not advanced. */
int dssm_w1_shard_reduce_adam(const float* const* peer_dW1, float* const* peer_W1, int32_t n_ranks, int32_t self, int32_t D,
                              int32_t L1, int32_t row_begin, int32_t row_end, float* m1, float* v1, const float* beta_pow,
                              float lr, float beta1, float beta2, float eps, dssm_stream_t stream);
/* The same exchange through an NVSwitch multicast mapping of the two buffers (NVLS): mc_dW1 / mc_W1 are the MULTICAST
 * addresses of dW1 / W1 (e.g. torch symmetric memory's multicast_ptr), W1_local this rank's ordinary pointer to its own
 * W1.  multimem.ld_reduce returns the switch-side fp32 sum of a gradient row over all replicas, multimem.st replicates
 * the new weight row into all of them: per GPU and direction the wire carries |W1| instead of 2(n-1)/n |W1|. */
int dssm_w1_shard_reduce_adam_mc(const float* mc_dW1, float* mc_W1, const float* W1_local, int32_t n_ranks, int32_t D, int32_t L1,
                                 int32_t row_begin, int32_t row_end, float* m1, float* v1, const float* beta_pow, float lr,
                                 float beta1, float beta2, float eps, dssm_stream_t stream);

/* PUSH exchange (default for N > 1): instead of a local dense dW1 that the owners pull, the dW1 gather writes every finished
 * gradient row straight into the OWNER's peer-mapped slot buffer -- rank o owns the W1 rows [o*per, (o+1)*per); buffer
 * layout [n_ranks][per][L1] floats, written at [self][row - o*per] -- with a validity stamp (epoch + 1; `epoch` = the device
 * word of the flag block, see dssm_peer_*) in the owner's array [n_ranks][per] of uint32.  Rows of columns absent from the
 * batch are not written at all (no zero-fill, no traffic).  After a flag round (all ranks' pushes visible), every rank runs
 * dssm_w1_slots_reduce_adam over its own rows: valid slots summed in rank order (LOCAL loads), averaged, TF-Adam with the
 * local m1 / v1, new weight row replicated into every rank's W1 (peer stores through peer_W1, or one multimem.st per 16
 * bytes through the NVSwitch multicast mapping mc_W1 when it is not NULL). */
int dssm_w1_slots_reduce_adam(const float* slots, const uint32_t* valid, const uint32_t* epoch, float* const* peer_W1, float* mc_W1,
                              int32_t n_ranks, int32_t self, int32_t D, int32_t L1, int32_t per, float* m1, float* v1,
                              const float* beta_pow, float lr, float beta1, float beta2, float eps, dssm_stream_t stream);

/* Flag synchronisation between the ranks for a CHUNKED exchange (the dW1 gather is issued in column chunks; the owner
 * kernels of chunk k run on a second stream as soon as every rank has finished that chunk, under the gather of chunk k+1).
 * Every rank owns a peer-mapped, ZERO-FILLED block of dssm_peer_flags_bytes() bytes; host_peer_flags[r] addresses rank r's
 * block.  A step has `stride` sync points idx = 0..stride-1.  dssm_peer_signal(idx): everything enqueued before it on
 * `stream` is visible to any rank that passes dssm_peer_wait(idx) (fence + st.release.sys of a value that grows with a
 * device-side epoch; ld.acquire.sys spin on the own block).  dssm_peer_epoch_advance once per step, after the last wait.
 * All ranks must issue the same sequence; the three calls are graph-capturable. */
size_t dssm_peer_flags_bytes(void);
int dssm_peer_signal(void* const* host_peer_flags, int32_t n_ranks, int32_t self, int32_t idx, int32_t stride, dssm_stream_t stream);
int dssm_peer_wait(const void* own_flags, int32_t n_ranks, int32_t idx, int32_t stride, dssm_stream_t stream);
int dssm_peer_epoch_advance(void* own_flags, dssm_stream_t stream);
/* dssm_peer_signal(idx) followed by dssm_peer_wait(idx) in ONE launch (a full cross-rank barrier on `stream`); with
 * advance_epoch != 0 it also does the dssm_peer_epoch_advance that ends a step.  Used by the push exchange, whose two sync
 * points sit on the critical path of every data-parallel step. */
int dssm_peer_barrier(void* const* host_peer_flags, int32_t n_ranks, int32_t self, int32_t idx, int32_t stride, int32_t advance_epoch,
                      dssm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * batch_normalization(x, phase_train, out_size)  (new_dssm.py:62-88), both instances of one layer
 * (query segment rows [0,B), doc segment rows [B,R)) in one call.
 * All per-column vectors are laid out [2][L]: index 0 = query instance, 1 = doc instance.
 *   on_train != 0: mean/var = batch moments (biased variance, tf.nn.moments); if update_ema the
 *                  shadows move: ema -= (1-decay)*(ema - batch)            (new_dssm.py:78-83)
 *   on_train == 0: mean/var = shadows                                       (new_dssm.py:85-86)
 *   out: mean,var,rstd = rsqrt(var+eps), scale = gamma*rstd, shift = beta - mean*scale  (:87)
 * The normalised tensor itself is not written: consumers apply act(x*scale+shift) on load.
 * workspace (dssm_bn_workspace_bytes, shared by dssm_bn_forward and dssm_bn_act_backward): must be ZERO-FILLED once
 * before its first use -- it starts with the ticket counters of the "last block finalizes" reductions, which every
 * call leaves at zero again.  Calls sharing a workspace must be ordered (same stream).
 */
size_t dssm_bn_workspace_bytes(int32_t R, int32_t L);
int dssm_bn_forward(const float* X, int32_t R, int32_t L, int32_t B, int32_t on_train, int32_t update_ema,
                    const float* gamma, const float* beta, float* ema_mean, float* ema_var, float eps,
                    float ema_decay, float* mean, float* var, float* rstd, float* scale, float* shift,
                    void* workspace, size_t workspace_bytes, dssm_stream_t stream);

/* Y = act(X*scale + shift) per segment (tf.nn.batch_normalization + tf.nn.relu, new_dssm.py:87,
 * :134-136, :156-158).  scale/shift NULL = identity (dssm_no_bn/my_dssm.py:98-121). */
int dssm_bn_act_apply(const float* X, int32_t R, int32_t L, int32_t B, const float* scale, const float* shift,
                      int32_t act, float* Y, dssm_stream_t stream);

/* Backward of act(BN(x)) for one layer.  In: dA = dLoss/d(post-activation) [R,L] ; H = pre-BN
 * activations saved by the forward.  Out (in place over dA): dH = dLoss/dH; dgamma,dbeta [2][L]; optionally
 * db [L] = sum_r dH[r,:] (the gradient of the pre-BN bias), evaluated through the exact identity
 * sum_r dH = -gamma*rstd*dgamma*mean(xhat) per instance -- analytically zero, numerically rounding noise, exactly as
 * the column sum the reference takes.  With scale == NULL (no-BN mode) only the activation derivative is applied
 * (db is not written: use dssm_colsum). */
int dssm_bn_act_backward(float* dA, const float* H, int32_t R, int32_t L, int32_t B, int32_t act,
                         const float* gamma, const float* mean, const float* rstd, const float* scale,
                         const float* shift, float* dgamma, float* dbeta, float* db, void* workspace,
                         size_t workspace_bytes, dssm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * FC2..: dense layer.  Replaces tf.matmul(x_out, weight2) + bias2 for the three inputs
 * (new_dssm.py:146-148) with the previous layer's BN + activation fused into the A-operand load:
 *   Hout[R,N] = act(Hprev*scale + shift)[R,K] . W[K,N] + bias[N]
 * scale/shift are [2][K] (per segment) or NULL (identity); act applies also when scale is NULL.
 */
/* workspace of dssm_fc_fwd and dssm_fc_bwd_dx (the pre-split weight image of the tensor-core path); 0 for FP32 */
size_t dssm_fc_fwd_workspace_bytes(int32_t K, int32_t N, int32_t gemm_mode);
int dssm_fc_fwd(const float* Hprev, int32_t R, int32_t K, int32_t B, const float* scale, const float* shift,
                int32_t act, const float* W, const float* bias, int32_t N, float* Hout, int32_t gemm_mode,
                void* workspace, size_t workspace_bytes, dssm_stream_t stream);
/* dA[R,K] = dH[R,N] . W[K,N]^T   (gradient w.r.t. the post-activation input of the layer) */
int dssm_fc_bwd_dx(const float* dH, int32_t R, int32_t N, const float* W, int32_t K, float* dA,
                   int32_t gemm_mode, void* workspace, size_t workspace_bytes, dssm_stream_t stream);
/* dW[K,N] = act(Hprev*scale+shift)^T . dH ;  db[N] = column sums of dH.  Deterministic split-K. */
size_t dssm_fc_bwd_dw_workspace_bytes(int32_t R, int32_t K, int32_t N);
int dssm_fc_bwd_dw(const float* Hprev, int32_t R, int32_t K, int32_t B, const float* scale, const float* shift,
                   int32_t act, const float* dH, int32_t N, float* dW, float* db, int32_t gemm_mode,
                   void* workspace, size_t workspace_bytes, dssm_stream_t stream);
/* out[N] = column sums of X[R,N] (bias gradient of FC1). Deterministic. */
size_t dssm_colsum_workspace_bytes(int32_t R, int32_t N);
int dssm_colsum(const float* X, int32_t R, int32_t N, float* out, void* workspace, size_t workspace_bytes,
                dssm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Merge_Negative_Doc (new_dssm.py:160-180): doc_y = [doc_positive_y ; negatives in slot-major order]
 *   doc_y[r] = pos[r] (r < B);  doc_y[(i+1)*B + j] = neg[j*NEG + i].
 * The index form writes src[r] = source row in the stacked [pos ; neg] matrix (bit-exact contract).
 */
int dssm_merge_negative_doc(const float* doc_positive_y, const float* doc_negative_y, int32_t B, int32_t NEG,
                            int32_t L, float* doc_y, dssm_stream_t stream);
int dssm_merge_negative_doc_index(int32_t B, int32_t NEG, int32_t* src, dssm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Cosine_Similarity + Loss (new_dssm.py:182-213) and their gradient, one warp per query group.
 * Y [R,L] = stacked embeddings (query_y ; doc_positive_y ; doc_negative_y), post-activation.
 * Outputs (any may be NULL except loss_terms):
 *   query_norm_single [B]            (:187)
 *   doc_norm   [(1+NEG)*B]           (:190)   reference order: row k*B + j
 *   cos_sim_raw[(1+NEG)*B]           (:197)   reference order: row k*B + j   (0/0 -> NaN, no epsilon)
 *   cos_sim    [B,(1+NEG)]           (:199)   = gamma * raw
 *   prob       [B,(1+NEG)]           (:206)   max-subtracted softmax
 *   loss_terms [B]                   -log(prob[j,0] + loss_eps)
 *   loss       [1]                   sum(loss_terms) / (loss_div_bs ? B : 1)   (:209)
 *   dY         [R,L]                 dLoss/dY, same stacked row order as Y (NULL = forward only)
 */
int dssm_cos_softmax_loss(const float* Y, int32_t B, int32_t NEG, int32_t L, float gamma, float loss_eps,
                          int32_t loss_div_bs, float* query_norm_single, float* doc_norm, float* cos_sim_raw,
                          float* cos_sim, float* prob, float* loss_terms, float* loss, float* dY,
                          dssm_stream_t stream);

/* Same with the last layer's BN + activation (new_dssm.py:87,156-158) fused in front: reads the pre-BN activations
 * H [R,L] and the [2][L] scale/shift (NULL = identity), WRITES the embeddings Y [R,L], then proceeds as above. */
int dssm_cos_softmax_loss_fused(const float* H, const float* scale, const float* shift, int32_t act, float* Y, int32_t B,
                                int32_t NEG, int32_t L, float gamma, float loss_eps, int32_t loss_div_bs,
                                float* query_norm_single, float* doc_norm, float* cos_sim_raw, float* cos_sim, float* prob,
                                float* loss_terms, float* loss, float* dY, dssm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training (new_dssm.py:215-217): tf.train.AdamOptimizer over one flat parameter buffer.
 *   beta_pow[2] (device) holds beta1^t, beta2^t (initialised to beta1, beta2);
 *   lr_t = lr*sqrt(1-b2p)/(1-b1p); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; w -= lr_t m/(sqrt(v)+eps)
 *   g = grads * grad_scale (1/world_size after a summing all-reduce).
 * dssm_adam_advance multiplies the powers by beta1, beta2 (TF does it after all variables' updates).
 */
int dssm_adam_step(float* params, const float* grads, float* m, float* v, int64_t n, const float* beta_pow,
                   float lr, float beta1, float beta2, float eps, float grad_scale, dssm_stream_t stream);
int dssm_adam_advance(float* beta_pow, float beta1, float beta2, dssm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Accuracy / Auc (new_dssm.py:219-231).  tf.metrics.auc(labels, cos_sim_raw, num_thresholds=T): a prediction is positive
 * at threshold t iff prediction > t; the reference never resets the metric's local variables (:252), so the two 64-bit
 * histograms pos_hist / neg_hist [T+1] (device, zero-filled once by the caller) accumulate over every update.
 * predictions [n]: the first n_pos carry label 1, the rest 0 (label = [1]*B + [0]*B*NEG, :163-165, cos_sim_raw's order);
 * thresholds [T] ascending (device; TF: -1e-7, i/(T-1), 1+1e-7 as float32).  dssm_auc_result writes 3 doubles (device):
 * {auc (trapezoidal ROC, TF's 1e-6 epsilon), positives seen, negatives seen}.  T <= 2048.
 * dssm_accuracy: out[0] = mean(argmax(prob,1) == 0) over prob [B, n_classes] (:220-221). */
int dssm_auc_update(const float* predictions, int32_t n_pos, int32_t n, const float* thresholds, int32_t num_thresholds,
                    uint64_t* pos_hist, uint64_t* neg_hist, dssm_stream_t stream);
int dssm_auc_result(const uint64_t* pos_hist, const uint64_t* neg_hist, int32_t num_thresholds, double* out, dssm_stream_t stream);
int dssm_accuracy(const float* prob, int32_t B, int32_t n_classes, float* out, dssm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Corpus cosine top-k.  No reference function exists; cosine follows
 * new_dssm.py:185-197 (no epsilon), ordering follows tf.nn.top_k(sorted=True)
 * (utils/tf_ranking_utils.py:47): score descending, ties to the lower doc id; NaN ranks as -inf.
 * Scores use the sequential fp32 multiply-then-add of oracle/retrieval_oracle.py so ids are bit-exact.
 *   out_scores [nq,k] fp32, out_ids [nq,k] int32 (= id_offset + local row), sorted.
 * dssm_topk_merge merges n_parts per-shard results laid out [n_parts][nq][k] into the global top-k.
 */
size_t dssm_corpus_topk_workspace_bytes(int32_t nq, int64_t nd, int32_t d, int32_t k);
int dssm_corpus_topk(const float* Q, int32_t nq, const float* docs, int64_t nd, int32_t d, int32_t k,
                     int32_t id_offset, float* out_scores, int32_t* out_ids, void* workspace,
                     size_t workspace_bytes, dssm_stream_t stream);
/* Tensor-core path of the same contract (d must be 128): tcgen05 tf32 contraction with a per-query threshold filter in
 * the TMEM epilogue, then exact fp32 rescoring of the survivors, so ids and scores are bit-identical to
 * dssm_corpus_topk.  *overflow_flag (device int) becomes 1 if a candidate list overflowed -- the result is then
 * incomplete and the caller must rerun dssm_corpus_topk (dssm_b200/retrieval.py does). */
size_t dssm_corpus_topk_tc_workspace_bytes(int32_t nq, int64_t nd, int32_t d, int32_t k);
int dssm_corpus_topk_tc(const float* Q, int32_t nq, const float* docs, int64_t nd, int32_t d, int32_t k, int32_t id_offset,
                        float* out_scores, int32_t* out_ids, int32_t* overflow_flag, void* workspace, size_t workspace_bytes,
                        dssm_stream_t stream);
/* Corpus INDEX path (bf16 storage for the filter): dssm_corpus_index_build is run once per corpus and writes
 * [fp32 row norms | the rows normalised, rounded to bf16 and laid out as SWIZZLE_128B tensor-core tiles] into `index`
 * (dssm_corpus_index_bytes(nd, d) bytes, 1024-byte aligned; d must be 128).  dssm_corpus_topk_indexed then streams 256 B
 * per document with one bulk copy per 128-doc tile, two resident query tiles per CTA (tcgen05.mma kind::f16, accumulators
 * double-buffered in TMEM) and one compare per score; survivors are re-scored exactly from the fp32 rows `docs`, so ids and
 * scores are bit-identical to dssm_corpus_topk.  *overflow_flag as for dssm_corpus_topk_tc. */
size_t dssm_corpus_index_bytes(int64_t nd, int32_t d);
int dssm_corpus_index_build(const float* docs, int64_t nd, int32_t d, void* index, size_t index_bytes, dssm_stream_t stream);
size_t dssm_corpus_topk_indexed_workspace_bytes(int32_t nq, int32_t k);
int dssm_corpus_topk_indexed(const float* Q, int32_t nq, const float* docs, const void* index, int64_t nd, int32_t d, int32_t k,
                             int32_t id_offset, float* out_scores, int32_t* out_ids, int32_t* overflow_flag, void* workspace,
                             size_t workspace_bytes, dssm_stream_t stream);
int dssm_topk_merge(const float* part_scores, const int32_t* part_ids, int32_t n_parts, int32_t nq, int32_t k,
                    float* out_scores, int32_t* out_ids, dssm_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Tower handle: the whole graph of new_dssm.py:104-217 over caller-owned device buffers.
 *
 * Flat parameter layout (all offsets in floats, each tensor padded to a multiple of 4 floats):
 *   for l = 1..n_layers:  W{l} [in,out], b{l} [out]
 *   if use_bn: for l = 1..n_layers:  bn{l}_gamma [2][out], bn{l}_beta [2][out]   (0 = query, 1 = doc)
 * grads / m / v use the same layout.  ema: for l: bn{l}_ema_mean [2][out], bn{l}_ema_var [2][out].
 * dssm_tower_tensor_info enumerates it (kind 0 = params/grads/m/v, 1 = ema, 2 = workspace tensors of
 * the last forward: h{l}, Y, dY, cos_sim_raw, query_norm_single, doc_norm, cos_sim, prob, loss_terms, loss,
 * bn{l}_mean/var/rstd/scale/shift).
 */
typedef struct dssm_tower dssm_tower;

int dssm_tower_create(const dssm_config* cfg, dssm_tower** out);
void dssm_tower_destroy(dssm_tower* t);
int64_t dssm_tower_param_count(const dssm_tower* t); /* floats in params (= grads = m = v) */
int64_t dssm_tower_ema_count(const dssm_tower* t);   /* floats in ema */
size_t dssm_tower_workspace_bytes(const dssm_tower* t, int64_t max_nnz);
int32_t dssm_tower_num_tensors(const dssm_tower* t, int32_t kind);
int dssm_tower_tensor_info(const dssm_tower* t, int32_t kind, int32_t index, char* name, int32_t name_cap,
                           int64_t* offset_floats, int64_t* rows, int64_t* cols);
/* beta_pow: 2 floats (device) initialised by the caller to {beta1, beta2}.  bind clears the ticket counters inside
 * the workspace (a synchronous cudaMemset: call it outside stream capture). */
int dssm_tower_bind(dssm_tower* t, float* params, float* grads, float* m, float* v, float* ema, float* beta_pow,
                    void* workspace, size_t workspace_bytes, int64_t max_nnz);
/* SyncBN (data-parallel option; closes the gap to the single-process reference, whose tf.nn.moments at new_dssm.py:77 span
 * the whole batch): every rank passes the peer-mapped exchange buffers of all n_ranks replicas (HOST array of device
 * pointers, entry `rank` its own; each dssm_tower_syncbn_bytes(t, n_ranks) bytes, ZERO-FILLED once, e.g. torch symmetric
 * memory).  Training-mode BN then uses  mean = avg_r mean_r,  var = avg_r (var_r + (mean_r - mean)^2)  and the backward
 * averages the column sums [dbeta | dgamma] over the replicas -- one single-CTA kernel per BN layer and direction that
 * pushes the statistics to every peer, meets at a flag barrier in peer memory and merges in rank order (csrc/nvlink.cu).
 * All ranks must run the same sequence of training steps.  n_ranks <= 1 turns it off.  Drops captured graphs. */
size_t dssm_tower_syncbn_bytes(const dssm_tower* t, int32_t n_ranks);
int dssm_tower_set_syncbn(dssm_tower* t, int32_t n_ranks, int32_t rank, void* const* host_peer_bufs);
/* sess.run(loss / embeddings, feed_dict=pull_batch(on_train, ...)) -- new_dssm.py:276-285,
 * load_model_and_save_vector.py:60-99.  Device CSR of the stacked batch. */
int dssm_tower_forward(dssm_tower* t, const int32_t* indptr, const int32_t* indices, const float* values,
                       int32_t on_train, int32_t update_ema, dssm_stream_t stream);
/* Gradients of `loss` w.r.t. every trainable into the bound grads buffer (training-mode forward must
 * precede it with the same CSR pointers still valid). */
int dssm_tower_backward(dssm_tower* t, dssm_stream_t stream);
/* Adam on the bound buffers; grad_scale multiplies the grads first (data-parallel average). */
int dssm_tower_adam(dssm_tower* t, float grad_scale, dssm_stream_t stream);
/* Data-parallel pipeline: dssm_tower_backward_begin = everything of the backward except the dW1 gather (dense
 * layers, small gradients, CSC build); dssm_tower_backward_w1(chunk k of n) = dW1 rows of column chunk k;
 * dssm_tower_adam_range = Adam over params[offset, offset+count) without advancing the beta powers;
 * dssm_tower_adam_advance advances them once per step.  Chunk k of n covers columns [k*ceil(D/n), ...) = the
 * contiguous float range reported by dssm_tower_w1_chunk. */
int dssm_tower_backward_begin(dssm_tower* t, dssm_stream_t stream);
int dssm_tower_backward_w1(dssm_tower* t, int32_t chunk, int32_t n_chunks, dssm_stream_t stream);
int dssm_tower_w1_chunk(const dssm_tower* t, int32_t chunk, int32_t n_chunks, int64_t* offset_floats, int64_t* count_floats);
/* Push exchange on the tower: dssm_tower_set_w1_push(t, 1) makes the CSC build skip the zero-fill of the local dW1;
 * dssm_tower_backward_w1_push is the dW1 gather with the rows pushed to their owners (arguments as dssm_w1_slots_reduce_adam:
 * HOST arrays of n_ranks device pointers to every rank's slot buffer and validity array). */
int dssm_tower_set_w1_push(dssm_tower* t, int32_t enabled);
int dssm_tower_backward_w1_push(dssm_tower* t, float* const* host_peer_slots, uint32_t* const* host_peer_valid, const uint32_t* epoch,
                                int32_t n_ranks, int32_t self, int32_t per, dssm_stream_t stream);
int dssm_tower_adam_range(dssm_tower* t, int64_t offset_floats, int64_t count_floats, float grad_scale, dssm_stream_t stream);
int dssm_tower_adam_advance(dssm_tower* t, dssm_stream_t stream);
/* Training forward + dssm_tower_backward_begin on the staging CSR; dssm_tower_capture_graph_dp turns that pair into
 * one CUDA graph that dssm_tower_fwd_bwd_begin_staged replays (the per-step CPU cost of ~40 launches disappears). */
int dssm_tower_capture_graph_dp(dssm_tower* t, dssm_stream_t stream);
int dssm_tower_fwd_bwd_begin_staged(dssm_tower* t, dssm_stream_t stream);
/* sess.run(train_step, feed_dict=pull_batch(True, ...)) -- new_dssm.py:267-269: forward + backward + Adam. */
int dssm_tower_train_step(dssm_tower* t, const int32_t* indptr, const int32_t* indices, const float* values,
                          dssm_stream_t stream);
/* Same with HOST CSR buffers (pinned for async copies): uploads into the workspace staging area, runs the
 * step and copies the loss back; synchronises the stream before returning.  host_loss may be NULL
 * (then no readback and no synchronisation). */
int dssm_tower_train_step_host(dssm_tower* t, const int32_t* host_indptr, const int32_t* host_indices,
                               const float* host_values, int64_t nnz, float* host_loss, dssm_stream_t stream);
/* Capture forward+backward+Adam over the bound buffers into a CUDA graph reading the CSR from the
 * workspace staging area; afterwards dssm_tower_train_step_host / _staged replay the graph. */
int dssm_tower_capture_graph(dssm_tower* t, dssm_stream_t stream);
/* Device pointers of the staging CSR (indptr [R+1], indices [max_nnz], values [max_nnz]). */
int dssm_tower_staging(dssm_tower* t, int32_t** indptr, int32_t** indices, float** values);
/* Run one train step on whatever is in the staging CSR (graph replay when captured). */
int dssm_tower_train_step_staged(dssm_tower* t, dssm_stream_t stream);
/* Pipelined host feed (the reference's training loop holds host batches: pull_batch, utils/utils.py:45-61, fed one
 * per sess.run, new_dssm.py:261-269).  dssm_tower_train_step_host_async uploads step k's CSR from PINNED host memory
 * on the tower's own copy stream into upload buffer k%2 while step k-1 still computes, then runs the step on `stream`
 * and copies the loss to host_loss (pinned, one slot per in-flight step); it never synchronises and returns k (< 0 on
 * error, see dssm_last_error).  dssm_tower_feed_wait(k) blocks until step k is done: its loss is valid and its host
 * buffers may be reused.  Only the last two issued steps are waitable; keep at most two in flight. */
int64_t dssm_tower_train_step_host_async(dssm_tower* t, const int32_t* host_indptr, const int32_t* host_indices,
                                         const float* host_values, int64_t nnz, float* host_loss, dssm_stream_t stream);
int dssm_tower_feed_wait(dssm_tower* t, int64_t step);
/* The two halves of dssm_tower_train_step_host_async for callers that run their own step on the staging CSR (the
 * data-parallel pipeline): upload (returns the step id k), then -- after enqueuing the step on `stream` --
 * dssm_tower_feed_step_done(k) copies the loss back and records the completion event dssm_tower_feed_wait(k) waits on. */
int64_t dssm_tower_feed_upload_async(dssm_tower* t, const int32_t* host_indptr, const int32_t* host_indices,
                                     const float* host_values, int64_t nnz, dssm_stream_t stream);
int dssm_tower_feed_step_done(dssm_tower* t, int64_t step, float* host_loss, dssm_stream_t stream);
/* HOST helper of the pipelined feed (no device work): write the rows [row_lo[p], row_hi[p]) of n_parts host CSR matrices,
 * stacked in order, as ONE int32/fp32 CSR into out_* (typically pinned) -- what pull_batch's three row slices + feed
 * conversion produce (utils/utils.py:45-61,20-24) without intermediate objects.  index_kind[p]: 0 = int32, 1 = int64
 * (indptr and indices of part p); value_kind[p]: 0 = float32, 1 = float64, 2 = int64, 3 = int32 (cast to float32 like the
 * feed does).  out_indptr needs sum(row_hi - row_lo) + 1 entries, out_indices / out_values `capacity`.  Copies are split
 * over up to n_threads host threads.  Returns nnz, or -1 (dssm_last_error). */
int64_t dssm_host_stack_csr(int32_t n_parts, const void* const* part_indptr, const void* const* part_indices,
                            const void* const* part_values, const int32_t* index_kind, const int32_t* value_kind,
                            const int64_t* row_lo, const int64_t* row_hi, int32_t* out_indptr, int32_t* out_indices,
                            float* out_values, int64_t capacity, int32_t n_threads);
/* Number of kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t dssm_tower_launch_count(const dssm_tower* t);
/* One un-graphed train step on the staging CSR with CUDA events between the phases; synchronises.
 * host_phase_ms[8] = {FC1 SpMM fwd, dense fwd (BN+FC), cosine/loss, dense bwd, CSC build, dW1 gather, db1, Adam}. */
int dssm_tower_profile_step(dssm_tower* t, float* host_phase_ms, dssm_stream_t stream);
/* The same events around the step as it really runs (side stream, W1 Adam fused into the gather): "CSC build" is then
 * the time the main stream waited for the side stream at the join, "Adam" the parameters behind W1. */
int dssm_tower_profile_step_overlapped(dssm_tower* t, float* host_phase_ms, dssm_stream_t stream);
/* Timeline of that step inside a CUDA graph: a 1-thread %globaltimer stamp after every call on the main stream (each
 * stamp is a graph node, ~1.5 us).  names <- labels joined by ';', ms[i] <- milliseconds between label i-1 and label i
 * (ms[0] = 0), *n_out <- number of labels.  Needs a non-default stream; allocates a 1 KB stamp buffer; synchronises. */
int dssm_tower_profile_timeline(dssm_tower* t, char* names, int32_t names_cap, float* ms, int32_t max_n, int32_t* n_out,
                                dssm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DSSM_B200_H_ */
