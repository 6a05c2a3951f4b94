/*
 * nasrec_b200 -- C ABI of the B200-native NASRec supernet hot path.
 *
 * Drop-in boundary: these are the entry points a binding
 * inside the reference's nasrec/supernet/{supernet,modules}.py would call
 * instead of its eager PyTorch ops.  Plain device pointers, sizes and a
 * cudaStream_t (passed as void*); no torch types.  Every function
 *   - launches asynchronously on `stream`, never synchronises, never allocates,
 *   - returns 0 on success, a cudaError_t (> 0) for a CUDA failure, or a
 *     negative NASREC_E* code for a rejected argument (nothing is launched),
 *   - is CUDA-graph capturable.
 * All floating-point data is fp32, indices are int64 (as the reference's
 * cat_feats), row-major.  "Citations" name the reference code each entry point
 * replaces (paths relative to the NasRec repository root).
 *
 * Data model.  A block output is stored *compact*: dense [B, d] holds only the
 * d live columns of the reference's zero-masked [B, 1024]; sparse
 * [B, s (+8), 16] holds the s live rows (+ the 8 dense->sparse merger rows).
 * A module input (the reference's zero-padded torch.cat of earlier outputs,
 * supernet.py:530-573) is therefore a *segment list*: for every selected
 * source, where it lives and which columns of the weight's K axis it meets.
 */
#ifndef NASREC_B200_H
#define NASREC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NASREC_MAX_SEGS 16
#define NASREC_EMB_DIM 16          /* supernet.py:224 embedding_dim */
#define NASREC_ATTN_PARAMS 1696    /* floats in one Transformer node's 16-wide parameter pack */

#define NASREC_EINVAL (-1)         /* bad argument (null pointer, size out of range) */
#define NASREC_ETOOBIG (-2)        /* size beyond what this build supports */
#define NASREC_ENOSPACE (-3)       /* a caller-provided arena is too small (nasrec_net_*): grow it and retry */

/* One source of a zero-padded concat.  2-D: `ptr` is [M, >=width] with row stride
 * `ld`; 3-D: `ptr` is [B, >=width, 16] with batch stride `ld` (floats).  `width`
 * live columns (2-D) / rows (3-D) meet weight columns [w_off, w_off+width). */
typedef struct {
    const float* ptr;
    int64_t ld;
    int64_t width;
    int64_t w_off;
} nasrec_seg_t;

/* Library identity: returns a version number; *sm (if non-null) gets the compute
 * capability the kernels were compiled for (100 for sm_100a). */
int nasrec_version(int* sm);

/* GEMM arithmetic mode of every nasrec_seg_linear_* / nasrec_sproj_* entry point:
 *   0  fp32 FFMA on CUDA cores (reference-exact fp32 products);
 *   3  tcgen05 kind::tf32 tensor cores, each fp32 operand split into tf32 hi+lo, products
 *      hi*hi + hi*lo + lo*hi accumulated in fp32 Tensor Memory ("3xTF32");
 *   4  as 3 plus lo*lo;   1  plain single-pass tf32 (not fp32 parity; diagnostics only);
 *   2  bf16 compute (the reference's --use_amp path, train_utils.py:146,247-286): every GEMM operand is rounded to
 *      nearest-even bfloat16 and ONE product per k-step is accumulated in fp32 Tensor Memory -- the products a
 *      kind::f16 bf16 MMA forms, issued on the tf32 pipe from fp32 containers; master weights, Adagrad state,
 *      activations and every non-GEMM kernel stay fp32.  The weight planes then hold rn_bf16(W): rebuild them
 *      (nasrec_planes_refresh) after switching to or from this mode.
 * Process-wide; returns 0 or NASREC_EINVAL. */
int nasrec_set_gemm_mode(int mode);
int nasrec_get_gemm_mode(void);
/* Optional caller-owned device scratch (the library never allocates).  With it, tensor-core GEMM
 * launches that would occupy fewer than 64 SMs split their K range over several CTAs and finish
 * with a fixed-order reduction (deterministic).  Pass (NULL, 0) to detach. */
int nasrec_set_workspace(float* ws, int64_t nfloats);
/* Optional overlap of weight-gradient GEMMs with the dY -> dX chain: with a side stream attached, the
 * op-level backward entry points (nasrec_linear_ln_bwd, nasrec_sproj_ln_bwd) fork their wgrad / bias-grad
 * launches onto it (event-ordered after their LayerNorm backward); nasrec_side_join(stream) makes `stream`
 * wait for all forked work and must be called before the gradients are read (and before a CUDA-graph
 * capture ends).  Buffers handed to those entry points must stay allocated until the join.
 * stream == NULL detaches.  The split-K workspace is halved between the two streams while attached. */
int nasrec_set_side_stream(void* stream);
int nasrec_side_join(void* stream);
/* TMA-fed operand path of the tensor-core GEMM (default on).  A launch takes it when every operand is
 * 16-byte aligned with row strides that are multiples of 4 floats and -- for forward / dgrad -- the weight's
 * pre-split planes have been announced with nasrec_set_weight_planes; anything else takes the LDG-producer
 * kernel.  Both paths use the same arithmetic and agree to the last bit. */
int nasrec_set_gemm_tma(int on);
/* Deferred weight gradients.  dW of a linear is read only by the optimizer, so while deferral is on,
 * nasrec_seg_linear_wgrad (and the op-level backward entry points built on it) only QUEUE their work when it qualifies for
 * the TMA kernel; nasrec_wgrad_flush(stream) then runs everything queued as one batched launch -- one grid over the output
 * tiles of all queued problems (up to 64 problems per launch) -- on `stream`.  Operands (dC, the input segments, dW) must stay
 * valid and unchanged until the flush; a call with accumulate != 0 drains the queue and runs at once.  Switching deferral off
 * drops the queue (flush first).  nasrec_wgrad_defer returns the previous setting; nasrec_wgrad_pending the queue length.
 * The step executor (nasrec_net_forward_backward) does this by itself unless nasrec_net_set_defer_wgrad(net, 0). */
int nasrec_wgrad_defer(int on);
/* Host-only diagnostic: the tile width (16 / 32 / 64 / 128 output columns per CTA) and the split-K factor (1 / 2 / 4 / 8 CTAs
 * of a thread-block cluster per output tile) the planner picks for `nprob` row-major problems [M x N x K] of one launch;
 * kind: 0 forward, 1 dgrad, 2 wgrad operand layouts.  Needs no GPU. */
int nasrec_gemm_plan(int kind, int M, int N, int K, int nprob, int* bn, int* ns);
int nasrec_wgrad_flush(void* stream);
int64_t nasrec_wgrad_pending(void);
/* Contractions of at most `k` elements (one or two k-tiles: 13 dense features, 16-wide FM / DotProduct projections, 26..64
 * sparse rows) run on the CUDA-core kernel instead of the tensor-core pipeline, whose set-up (TMEM allocation, tensor maps,
 * mbarrier ring) costs more than such a problem; fp32-parity modes (3, 4) only.  Default 0 = off (the generic CUDA-core
 * kernel measured slower inside the training step).  Returns the previous value. */
int nasrec_set_small_k(int k);
/* Pre-split copies of a [rows, cols] weight W (modules.py nn.LazyLinear weights keep the reference layout, whose
 * row stride such as 1037 floats no tensor map can describe, and whose second concat source starts at column
 * nd = 13 or F = 26, which no TMA box can start at): hi = rn_tf32(W), lo = W - hi, both [rows, ldp] with
 * ldp % 4 == 0; W column c lives in plane column c + (c >= first ? shift : 0), shift = (4 - first % 4) % 4, all other
 * plane columns zero (pass first = the width of the first concat source, or 0).  The announcement is a one-entry
 * hint consumed by the GEMM entry points called next with the same W pointer (ldw == cols); W == NULL clears it.
 * Keep the planes (zero-initialised by the caller) in step with W through nasrec_planes_refresh or
 * nasrec_adagrad_multi_planes. */
int nasrec_set_weight_planes(const float* W, const float* hi, const float* lo, int64_t ldp, int rows, int cols,
                             int first);
int nasrec_planes_refresh(const float* W, int64_t ldw, int rows, int cols, int first, float* hi, float* lo,
                          int64_t ldp, void* stream);
/* Live GEMM accounting for roofline reports: what = 1 starts recording a CUDA-event pair (on the launching stream) and
 * the algorithmic flops (2 M N K over the live support) of every GEMM launch of the library; what = 0 stops; what = 2
 * stops, synchronises on the events and writes {total ms, launches, flops} to out3 (host doubles). */
int nasrec_gemm_prof(int what, double* out3);
/* Host-side launch accounting: what = 1 starts (and clears), 0 stops, 2 returns the nanoseconds spent inside
 * cudaLaunchKernelEx since the start, 3 the number of launches; what = 4 returns the number of kernels the library has
 * launched since it was loaded (always counted: bench.py's gpu_launches); what = 10 starts the per-kernel device trace
 * (a CUDA-event pair around every launch), 11 stops it, appends `name launches total_us` lines to $NASREC_TRACE_FILE and
 * returns the number of traced launches. */
int64_t nasrec_host_prof(int what);
/* which = 0: tensor-map cache hits, 1: tensor maps encoded, 2: GEMM launches that took the TMA path */
int64_t nasrec_tensor_map_stats(int which);

/* ------------------------------------------------------------------ embedding
 * a1  SuperNet._input_stem_layers_bi_output, supernet.py:404-430:
 *     out[b,f,:] = tables[f][idx[b,f],:]   (F nn.Embedding lookups + torch.stack).
 * tables: device array of F table base pointers; num_rows: device array [F].
 * err_flag (device int, may be null) is OR-ed with 1 on an out-of-range id
 * (the reference relies on PyTorch's device-side assert); such ids read row 0. */
int nasrec_emb_gather_fwd(const float* const* tables, const int64_t* num_rows, const int64_t* idx,
                          float* out, int B, int F, int* err_flag, void* stream);

/* a1 backward, replaces F embedding_dense_backward calls (nn.Embedding, supernet.py:407).
 * Deterministic: per table, (row, sample) keys are sorted; duplicate rows are summed in
 * ascending sample order.  Outputs per table f: uniq[f, 0..nuniq[f]) ascending row ids,
 * row_grad[f, u, :] the summed gradient rows, sumsq[f] = sum of squares of row_grad
 * (for the global clip norm).  seg_scratch: int32 [F, B+1].  One CTA per table, shared-memory sort: B <= 16384 and
 * best for B <= 2048; larger batches: nasrec_emb_grad_sort_reduce_big. */
int nasrec_emb_grad_sort_reduce(const int64_t* idx, const float* gout, int B, int F,
                                int64_t* uniq, int* nuniq, float* row_grad, float* sumsq,
                                int* seg_scratch, void* stream);
/* Same with the forward gather's bounds check: ids outside [0, num_rows[f]) are dropped from the reduction (they
 * would address rows that do not exist) and err_flag (device int, may be null) is OR-ed with 1. */
int nasrec_emb_grad_sort_reduce_checked(const int64_t* idx, const int64_t* num_rows, int* err_flag,
                                        const float* gout, int B, int F, int64_t* uniq, int* nuniq,
                                        float* row_grad, float* sumsq, int* seg_scratch, void* stream);

/* The same reduction for large batches (the data-parallel global batch, 16 K-sample KDD batches): multi-CTA, any B
 * (B * F < 2^31, F <= 31, rows per table < 2^27) -- one stable device radix sort of (table, row, sample) keys, head
 * flags + scan, 16 lanes per unique row (a whole CTA for rows with more than 256 duplicates, partial sums combined in a
 * fixed order), per-table sums of squares.  Deterministic; duplicate rows are summed in ascending sample order up to 256
 * duplicates, in 16 strided ascending runs beyond.  ws: nasrec_emb_grad_sort_reduce_big_ws_bytes(B, F) bytes. */
int64_t nasrec_emb_grad_sort_reduce_big_ws_bytes(int B, int F);
int nasrec_emb_grad_sort_reduce_big(const int64_t* idx, const int64_t* num_rows, int* err_flag, const float* gout,
                                    int B, int F, int64_t* uniq, int* nuniq, float* row_grad, float* sumsq, void* ws,
                                    int64_t ws_bytes, void* stream);
/* Scatter the reduced rows into dense zero-initialised [N_f,16] gradients (what
 * nn.Embedding.weight.grad holds in the reference). */
int nasrec_emb_grad_to_dense(const int64_t* uniq, const int* nuniq, const float* row_grad,
                             float* const* grad_tables, int B, int F, void* stream);

/* a17 restricted to the touched rows; identical to torch.optim.Adagrad(eps) on the dense
 * gradient because untouched rows have g == 0 (train_supernet.py:121-123):
 *   g = row_grad * clip_coef[0];  state[row] += g*g;  W[row] -= lr * g / (sqrt(state[row]) + eps). */
int nasrec_emb_rowwise_adagrad(const int64_t* uniq, const int* nuniq, const float* row_grad,
                               float* const* tables, float* const* states, int B, int F,
                               float lr, float eps, const float* clip_coef, void* stream);

/* ------------------------------------------------------- segment linear (2-D)
 * a5/a7/a8/a11/a12/a4  nn.LazyLinear applied to a zero-padded concat
 * (modules.py:171,340,385,489,578,584,740; supernet.py:598,1140):
 *   C[m,n] = sum_s sum_k A_s[m,k] * W[n_off+n, w_off_s+k] (+ bias[n_off+n]),  n < N.
 * Zero padding contributes nothing, so only the live K support is read. */
int nasrec_seg_linear_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int n_off,
                          int N, const float* bias, float* C, int64_t ldc, int M, void* stream);
/* dA_s[m,k] (+)= sum_n dC[m,n] * W[n_off+n, w_off_s+k]; segment outputs must not alias. */
int nasrec_seg_linear_dgrad(const float* dC, int64_t ldc, int N, const float* W, int64_t ldw, int n_off,
                            const nasrec_seg_t* dsegs, int nseg, int M, int accumulate, void* stream);
/* dW[n_off+n, w_off_s+k] (+)= sum_m dC[m,n] * A_s[m,k]; segments must not overlap in W. */
int nasrec_seg_linear_wgrad(const float* dC, int64_t ldc, int N, const nasrec_seg_t* segs, int nseg,
                            float* dW, int64_t ldw, int n_off, int M, int accumulate, void* stream);
/* out[n] (+)= sum_m x[m,n]  (bias gradients); deterministic. */
int nasrec_colsum(const float* x, int64_t ld, int M, int N, float* out, int accumulate, void* stream);

/* ------------------------------------------- projection along the sparse axis
 * a9/a10/a6  nn.LazyLinear on sparse.transpose(1,2) (modules.py:222-223,358-359,648):
 *   Z[b,p,e] = sum_s sum_r W[p, w_off_s+r] * X_s[b,r,e] (+ bias[p]),  p < P, e < 16. */
int nasrec_sproj_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int P,
                     const float* bias, float* Z, int64_t z_bstride, int B, void* stream);
int nasrec_sproj_dgrad(const float* dZ, int64_t dz_bstride, int P, const float* W, int64_t ldw,
                       const nasrec_seg_t* dsegs, int nseg, int B, int accumulate, void* stream);
/* workspace: nasrec_sproj_wgrad_ws_floats(P, total_width, B) floats. */
int64_t nasrec_sproj_wgrad_ws_floats(int P, int64_t total_width, int B);
int nasrec_sproj_wgrad(const float* dZ, int64_t dz_bstride, int P, const nasrec_seg_t* segs, int nseg,
                       float* dW, int64_t ldw, int B, int accumulate, float* ws, void* stream);

/* db[p] (+)= sum_{b,e} dZ[b,p,e]  (bias gradient of the projection; use_layernorm=False models). */
int nasrec_sproj_bias_grad(const float* dZ, int64_t dz_bstride, int P, int B, float* db, int accumulate,
                           void* stream);

/* ------------------------------------------------------- op-level sequences
 * One call per operator direction (same kernels as the fine-grained entry points above and
 * below, sequenced on `stream`); they exist to cut host-side call overhead.  z/dz/mean/rstd are
 * caller-provided scratch ([M,N] / [B,P,16]); a null gamma means "no LayerNorm" (plain activation);
 * null dW/dbias/dgamma/dbeta/dsegs[i].ptr mean "gradient not wanted"; dseg_accumulate[i] = 1 adds
 * into the target instead of overwriting it.  Targets must be distinct and segments must meet
 * distinct weight columns (callers fall back to the fine-grained calls otherwise). */
int nasrec_linear_ln_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int n_off, int N,
                         const float* bias, const float* gamma, const float* beta, float eps, int relu, int d_out,
                         float* z, float* y, int64_t ldy, float* mean, float* rstd, int accumulate, int M,
                         void* stream);
int nasrec_linear_ln_bwd(const float* dy, int64_t lddy, int d_out, const float* z, int M, int N, const float* gamma,
                         const float* beta, const float* mean, const float* rstd, int relu, const nasrec_seg_t* segs,
                         const nasrec_seg_t* dsegs, const int* dseg_accumulate, int nseg, const float* W, int64_t ldw,
                         int n_off, float* dW, float* dbias, float* dgamma, float* dbeta, float* dz, void* stream);
int nasrec_sproj_ln_fwd(const nasrec_seg_t* segs, int nseg, const float* W, int64_t ldw, int P, const float* bias,
                        const float* gamma, const float* beta, float eps, int relu, int p_out, float* z, float* y,
                        int64_t y_bstride, float* mean, float* rstd, int accumulate, int B, void* stream);
int nasrec_sproj_ln_bwd(const float* dy, int64_t dy_bstride, int p_out, const float* z, int B, int P, const float* gamma,
                        const float* beta, const float* mean, const float* rstd, int relu, const nasrec_seg_t* segs,
                        const nasrec_seg_t* dsegs, const int* dseg_accumulate, int nseg, const float* W, int64_t ldw,
                        float* dW, float* dbias, float* dgamma, float* dbeta, float* dz, float* ws, void* stream);

/* ------------------------------------------------------- LayerNorm epilogues
 * nn.LayerNorm(N) + activation + prefix mask (modules.py:171-181, 385-400, 489-499):
 *   y[m,j] (+)= act(LN(x[m,0:N])[j]) for j < d_out  (columns >= d_out are the masked ones
 *   and are not stored); mean/rstd [M] are saved for the backward. relu: 0/1. */
int nasrec_ln_fwd(const float* x, int64_t ldx, int M, int N, const float* gamma, const float* beta,
                  float eps, int relu, int d_out, float* y, int64_t ldy, float* mean, float* rstd,
                  int accumulate, void* stream);
/* dx [M,N] from dy [M,d_out]; dgamma/dbeta [N] (+)= column sums (zero beyond d_out). */
int nasrec_ln_bwd(const float* dy, int64_t lddy, int d_out, const float* x, int64_t ldx, int M, int N,
                  const float* gamma, const float* beta, const float* mean, const float* rstd, int relu,
                  float* dx, int64_t lddx, float* dgamma, float* dbeta, int accumulate_params, void* stream);
/* Same over the P axis of z [B,P,16] (nn.LayerNorm(P) on the transposed tensor,
 * modules.py:224-230,360,649): y[b,p,e] = act(LN_p(z[b,:,e])[p]) for p < p_out;
 * mean/rstd are [B,16]. */
int nasrec_ln3_fwd(const float* z, int64_t z_bstride, int B, int P, const float* gamma, const float* beta,
                   float eps, int relu, int p_out, float* y, int64_t y_bstride, float* mean, float* rstd,
                   int accumulate, void* stream);
int nasrec_ln3_bwd(const float* dy, int64_t dy_bstride, int p_out, const float* z, int64_t z_bstride, int B,
                   int P, const float* gamma, const float* beta, const float* mean, const float* rstd, int relu,
                   float* dz, int64_t dz_bstride, float* dgamma, float* dbeta, int accumulate_params,
                   void* stream);
/* Plain epilogue without LayerNorm (use_layernorm=False fixed models): y = act(x) (+)=. */
int nasrec_act_fwd(const float* x, int64_t ldx, int M, int N, int relu, float* y, int64_t ldy,
                   int accumulate, void* stream);
int nasrec_act_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, int M, int N, int relu,
                   float* dx, int64_t lddx, void* stream);

/* ----------------------------------------------------------------- DotProduct
 * a6  modules.py:366-383: T = [x ; y] ([B,1+P,16]); Z = T T^T; R = strict lower triangle
 * of Z in torch.tril_indices row-major order ((1,0),(2,0),(2,1),...): R [B, (P+1)P/2]. */
int nasrec_dot_tril_fwd(const float* x, int64_t ldx, const float* y, int64_t y_bstride, int P,
                        float* R, int64_t ldr, int B, void* stream);
int nasrec_dot_tril_bwd(const float* dR, int64_t ldr, const float* x, int64_t ldx, const float* y,
                        int64_t y_bstride, int P, float* dx, int64_t lddx, float* dy, int64_t dy_bstride,
                        int B, void* stream);

/* ------------------------------------------------------------ SigmoidGating
 * a7  modules.py:576-582: out[m, koff_s+k] = sigmoid(pre[m, koff_s+k]) * right_s[m,k], where the
 * right operand is a segment list and koff_s is the running sum of widths. */
int nasrec_gate_fwd(const float* pre, int64_t ldp, const nasrec_seg_t* right, int nseg, float* out,
                    int64_t ldo, int M, void* stream);
/* dpre = dout * right * s(1-s);  dright_s (+)= dout * s   (dright given as segments). */
int nasrec_gate_bwd(const float* dout, int64_t ldo, const float* pre, int64_t ldp, const nasrec_seg_t* right,
                    const nasrec_seg_t* dright, int nseg, float* dpre, int64_t lddp, int M, int accumulate,
                    void* stream);
/* out[m, w_off_s+k] (+)= A_s[m,k]: materialises a padded concat (rare no-projection corner,
 * modules.py:487-491, 582-586). */
int nasrec_concat_segs(const nasrec_seg_t* segs, int nseg, float* out, int64_t ldo, int M, int accumulate,
                       void* stream);

/* ----------------------------------------------------- FactorizationMachine3D
 * a11  modules.py:736-738: ix[b,e] = (sum_r x[b,r,e])^2 - sum_r x[b,r,e]^2, r < rows. */
int nasrec_fm_fwd(const float* x, int64_t x_bstride, int rows, float* ix, int B, void* stream);
/* dx[b,r,e] = (dx_in ? dx_in[b,r,e] : 0) + 2*dix[b,e]*(S[b,e] - x[b,r,e]). */
int nasrec_fm_bwd(const float* dix, const float* x, int64_t x_bstride, int rows, const float* dx_in,
                  int64_t dxin_bstride, float* dx, int64_t dx_bstride, int B, void* stream);

/* ------------------------------------------------------------------ Attention
 * a10  modules.py:664-688: nn.MultiheadAttention(16, heads=8, batch_first) self-attention over
 * L tokens, +residual, LayerNorm(16), FC-ReLU-FC, +residual, LayerNorm(16), row mask.
 * x is [B, s_live, 16]; tokens s_live..L-1 are the masked (all-zero) rows which still act as
 * keys/values through the in-proj bias.  params: HOST array of 12 device pointers, in order
 * in_proj_weight[48,16] in_proj_bias[48] out_proj.weight[16,16] out_proj.bias[16]
 * attn_ln.w[16] attn_ln.b[16] fc1.w[16,16] fc1.b[16] fc2.w[16,16] fc2.b[16] fc_ln.w[16] fc_ln.b[16]
 * (NASREC_ATTN_PARAMS floats in total; dparams below is one packed buffer in the same order).
 * y [B, s_live, 16]. L <= 64. */
int nasrec_attn_fwd(const float* x, int64_t x_bstride, int L, int s_live, const float* const* params,
                    float* y, int64_t y_bstride, int B, void* stream);
/* dx [B,s_live,16]; dparams [NASREC_ATTN_PARAMS] (+)=; ws: nasrec_attn_bwd_ws_floats(B) floats. */
int64_t nasrec_attn_bwd_ws_floats(int B);
int nasrec_attn_bwd(const float* dy, int64_t dy_bstride, const float* x, int64_t x_bstride, int L,
                    int s_live, const float* const* params, float* dx, int64_t dx_bstride, float* dparams,
                    int accumulate_params, float* ws, int B, void* stream);

/* ------------------------------------------------------------------ loss/step
 * a17  BCEWithLogitsLoss(mean) forward + gradient (train_supernet.py:104, train_utils.py:266):
 * loss[0] = mean(max(z,0) - z*y + log1p(exp(-|z|))); dlogits[b] = (sigmoid(z)-y)/B * grad_scale. */
int nasrec_bce_fwd_bwd(const float* logits, const float* y, int B, float grad_scale, float* loss,
                       float* dlogits, void* stream);
/* f3  evaluation metrics over n concatenated predictions (train_utils.py:158-178):
 * out3[0] = accuracy of (sigmoid(z) > 0.5) vs y; out3[1] = ROC-AUC, ties counted 1/2 (==
 * sklearn.metrics.roc_auc_score on the fp32 sigmoid outputs; exact integer pair count);
 * out3[2] = mean BCE-with-logits.  out3: 3 device doubles.  ws: nasrec_binary_metrics_ws_bytes(n). */
int64_t nasrec_binary_metrics_ws_bytes(int64_t n);
int nasrec_binary_metrics(const float* logits, const float* y, int64_t n, void* ws, int64_t ws_bytes,
                          double* out3, void* stream);
/* f2  raw batch -> model inputs (data_pipes.py:135-175, VanillaTransform{Criteo,Avazu,KDD}):
 * int_x[b,c] = log(max(0, dense_raw[b,c]) + 1);
 * cat_x[b,f] = (int(hex[b,f], 16) if non-empty else -1).fmod(num_rows[f] - 1) + 1.
 * dense_raw[b,c] sits at dense_raw[b*dense_stride_b + c*dense_stride_c] (row- or column-major raw
 * columns); field (b,f) is `width` bytes at hex + b*hex_stride_b + f*hex_stride_f, NUL-padded hex
 * digits of either case (width <= 15); a non-hex byte sets err_flag.  num_rows: F device int64.
 * Outputs are contiguous [B,nd] float and [B,F] int64.  For Avazu pass nd = 0 (all-zero dense input). */
int nasrec_input_transform(const float* dense_raw, int64_t dense_stride_b, int64_t dense_stride_c, int nd,
                           const uint8_t* hex, int64_t hex_stride_b, int64_t hex_stride_f, int width, int F,
                           const int64_t* num_rows, int64_t B, float* int_x, int64_t* cat_x, int* err_flag,
                           void* stream);
/* (grads/sizes/w/state below are HOST arrays of device pointers / element counts.)
 * Global L2 norm of n tensors + extra pre-reduced sums of squares, then the
 * clip_grad_norm_ coefficient (train_utils.py:285): out[0] = total_norm,
 * out[1] = min(1, max_norm / (total_norm + 1e-6)).  partial: >= nasrec_sumsq_ws_floats() floats. */
int64_t nasrec_sumsq_ws_floats(const int64_t* sizes, int n);
int nasrec_grad_norm_clip(const float* const* grads, const int64_t* sizes, int n, const float* extra_sumsq,
                          int n_extra, float max_norm, float* partial, float* out, void* stream);
/* torch.optim.Adagrad(lr, eps, lr_decay=0) over n dense tensors:
 *   g = grad*clip_coef[0]; state += g*g; w -= lr * g / (sqrt(state)+eps). */
int nasrec_adagrad_multi(float* const* w, const float* const* grads, float* const* state,
                         const int64_t* sizes, int n, float lr, float eps, const float* clip_coef,
                         void* stream);
/* Same, and rewrites the hi/lo planes (nasrec_set_weight_planes) of every tensor i with hi[i] != NULL from the
 * updated weights in the same pass (cols[i] = row length of w[i], first[i] / ldp[i] as for nasrec_set_weight_planes). */
int nasrec_adagrad_multi_planes(float* const* w, const float* const* grads, float* const* state,
                                const int64_t* sizes, int n, float lr, float eps, const float* clip_coef,
                                float* const* hi, float* const* lo, const int* cols, const int* first,
                                const int64_t* ldp, void* stream);

/* ---------------------------------------------------------------- step executor
 * The whole hot path behind one handle: SuperNet.forward (supernet.py:513-603), SuperNetBlock.forward
 * (:1067-1162) with every weight-sharing module of supernet/modules.py, BCE, backward, clip and Adagrad
 * (train_utils.py:262-286), sequenced on the host in C++ from (a) a one-off description of the model and
 * (b) the sampled subnet as a flat int array -- the "kernel path that takes the sampled subnet's
 * mask/choice directly".  Same kernels and launch order as the per-operator entry points above.
 *
 * desc_i: [num_blocks, nd, F, final_w, final_b, emb_param[F], then per block: num_nodes, max_dense,
 *          max_sparse, dotproduct_P, merger{W,b,ln_g,ln_b}, fm{W,b,ln_g,ln_b}, num_nodes x {type, p[24]}]
 *   (values are indices into the parameter table, -1 = absent; node types 0 FC, 1 DotProduct, 2 Sum,
 *   3 SigmoidGating, 4 EFC, 5 Transformer, 6 Zeros2D, 7 Zeros3D; p[] in state-dict order of the node).
 * w/state: HOST arrays of n_params device pointers (state = Adagrad accumulators, may be NULL for
 *   inference); numel/rows/cols/req per parameter.  d_tables/d_rows/d_tables_rw/d_states/d_err: DEVICE
 *   arrays as for nasrec_emb_gather_fwd / nasrec_emb_rowwise_adagrad.
 * choice: per block 49 ints: 5 lists {count, idx[8]} (dense_idx, sparse_idx, dense_left_idx,
 *   dense_right_idx, active_nodes) then dense_in_dims, sparse_in_dims, dense_sparse_interact, deep_fm.
 * Arenas are caller-owned device memory; NASREC_ENOSPACE asks for bigger ones
 * (nasrec_net_arena_high_water says how big).  Fixed (standalone) models are not handled here. */
void* nasrec_net_create(const int* desc_i, int desc_len, int n_params, float* const* w, float* const* state,
                        const int64_t* numel, const int* rows, const int* cols, const int* req,
                        const float* const* d_tables, const int64_t* d_rows, float* const* d_tables_rw,
                        float* const* d_states, int* d_err);
void nasrec_net_destroy(void* net);
/* (Re)attaching arenas discards the current step: nasrec_net_sparse_reduce / nasrec_net_apply then return NASREC_EINVAL
 * until the next nasrec_net_forward_backward. */
int nasrec_net_set_arenas(void* net, void* act, int64_t act_bytes, void* pgrad, int64_t pgrad_bytes);
int nasrec_net_set_requires_grad(void* net, const int* req, int n_params);
/* hi/lo planes per parameter (HOST arrays of n_params device pointers, NULL entries = none; ldp in floats, first as above): the
 * executor announces them to the GEMM entry points and keeps them in step inside nasrec_net_apply. */
int nasrec_net_set_planes(void* net, float* const* hi, float* const* lo, const int64_t* ldp, const int* first,
                          int n_params);
/* Rows of the batch nasrec_net_sparse_reduce will be given (the all-gathered global batch under data parallelism; default:
 * the step's own B).  forward_backward reserves the reduction's and the clip's scratch for that many rows up front, so the
 * two later calls never run out of arena after the step's gradients exist. */
int nasrec_net_set_reserve(void* net, int rows);
int nasrec_net_set_overlap(void* net, int on);      /* 1: join the side stream (nasrec_set_side_stream) after backward; 2: join it in nasrec_net_apply instead (the caller reads no gradient in between), so that nasrec_net_sparse_reduce overlaps the last batch of parameter gradients */
int nasrec_net_set_defer_wgrad(void* net, int on); /* queue dense weight gradients during backward, one batched launch at its end (default on) */
/* Data-parallel overlap: during nasrec_net_forward_backward, cb(offset_bytes, nbytes) is called on the host each
 * time a block's parameter gradients are final -- the byte range of the gradient bucket sealed since the last call,
 * already ordered on `stream` -- so the caller can start all-reducing it while backward continues.  NULL: off. */
typedef void (*nasrec_seal_cb_t)(int64_t offset_bytes, int64_t nbytes);
int nasrec_net_set_seal_callback(void* net, nasrec_seal_cb_t cb);
/* logits [B] for one subnet; emb_rows (optional) = [B,F,16] rows gathered once and shared by many candidates. */
int nasrec_net_forward(void* net, const int* choice, const float* int_x, const int64_t* cat_x, const float* emb_rows,
                       int B, float* logits, void* stream);
/* Batched multi-subnet evaluation (searcher_utils.py:57-104, eval_subnet_from_supernet.py:182-198): logits [n_cand, B] of
 * n_cand subnets (choices: n_cand x num_blocks x 49 ints, layout as above) on ONE batch against the resident weights.
 * A block's output depends only on its own choice and on the blocks it reads, so a block that several candidates agree
 * on (together with everything upstream of it) is computed once and shared -- always the stem, and about half of the
 * blocks for the children of one EA generation, which differ from their parent in one field of one block.  Results are
 * bit-identical to n_cand nasrec_net_forward calls.  stats (host, may be NULL): {blocks computed, blocks reused}. */
int nasrec_multi_subnet_eval(void* net, const int* choices, int n_cand, const float* int_x, const int64_t* cat_x,
                             const float* emb_rows, int B, float* logits, int* stats, void* stream);
/* forward + BCEWithLogits(mean)*grad_scale + backward.  Afterwards the dense parameter gradients lie back to
 * back in the pgrad arena (nasrec_net_grad_bucket: what data-parallel training all-reduces in place) and the
 * raw embedding gradient [B,F,16] is exposed by nasrec_net_sparse_raw. */
int nasrec_net_forward_backward(void* net, const int* choice, const float* int_x, const int64_t* cat_x, const float* y,
                                int B, float grad_scale, float* logits, float* loss, void* stream);
int nasrec_net_grad_bucket(void* net, float** ptr, int64_t* nfloats);
int nasrec_net_sparse_raw(void* net, const int64_t** cat_x, float** gout);
/* sorted-row reduction of the step's own embedding gradient (cat_all == NULL) or of an all-gathered one */
int nasrec_net_sparse_reduce(void* net, const int64_t* cat_all, const float* gout_all, int B_all, void* stream);
/* clip_grad_norm_(max_norm; <= 0: off) + Adagrad on what received a gradient; norm_out: 2 device floats */
int nasrec_net_apply(void* net, float lr, float eps, float max_norm, float* norm_out, void* stream);
int64_t nasrec_net_launches(void);
int64_t nasrec_net_arena_high_water(void* net, int which);

#ifdef __cplusplus
}
#endif
#endif /* NASREC_B200_H */
