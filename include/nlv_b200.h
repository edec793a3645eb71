/* nlv_b200 — C ABI of the B200-native (sm_100a) hot path of rlqja1107/NL-VSGG.
 *
 * Drop-in boundary.  The reference has no FFI of its own for this path: its
 * boundary is the Python module API (lib/sttran.py, lib/dsg_detr.py, lib/transformer*.py,
 * lib/evaluation_recall.py) over two native surfaces — the pybind11 extension
 * fasterRCNN/lib/model/csrc/vision.cpp:7-13 (nms, roi_align_forward/backward) and two Cython
 * modules (lib/draw_rectangles/draw_rectangles.pyx:11, lib/fpn/box_intersections_cpu/bbox.pyx:15).
 * Every entry point below names the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; all pointers are DEVICE pointers unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void*; every function only enqueues work on it;
 *   - no allocation inside: callers own all buffers (workspace sizes are documented per call);
 *   - return value NLV_OK (0) or a negative NLV_ERR_*; nlv_last_error() gives the text
 *     (thread-local).  Nothing throws, nothing falls back to the CPU.
 *   - dtype tags: NLV_F32 = 0, NLV_BF16 = 1.
 */
#ifndef NLV_B200_H_
#define NLV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLV_OK 0
#define NLV_ERR_INVALID_ARGUMENT (-1)
#define NLV_ERR_CUDA (-2)
#define NLV_ERR_UNSUPPORTED (-3)

#define NLV_F32 0
#define NLV_BF16 1

#define NLV_MAJOR_K 0  /* operand stored [MN, K] row-major (K contiguous)  */
#define NLV_MAJOR_MN 1 /* operand stored [K, MN] row-major (MN contiguous) */

const char* nlv_last_error(void);
int nlv_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long nlv_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * GEMM  D[m,n] = act( sum_k A(m,k) * B(n,k) + bias[n] ) + residual[m,n]
 * Replaces the cuBLAS calls behind nn.Linear / nn.MultiheadAttention in/out projections /
 * 1x1 conv in lib/sttran.py:336-348,370-372,381-387,404-406 and lib/transformer.py:9-13,38-42,
 * forward and backward (dX: B MN-major; dW: A and B MN-major).
 *   ab_dtype NLV_BF16 -> tcgen05/TMEM kernel (TMA-fed, fp32 accumulate); requires a,b 16-byte
 *                        aligned and lda,ldb multiples of 8.
 *   ab_dtype NLV_F32  -> exact-fp32 SIMT kernel (any shape/stride); parity mode and tiny shapes.
 * ------------------------------------------------------------------------------------------ */
/* Counter-based dropout (Philox4x32-10; csrc/philox.cuh): replaces nn.Dropout / MultiheadAttention(dropout=0.1) of
 * lib/transformer.py:9-29,38-57, lib/sttran.py:46, lib/dsg_detr.py:28,48.  thr16 = round(p * 65536) (0 = off), scale = 1 / (1 - p);
 * `stream` names the dropout site, so forward and backward regenerate identical masks from (seed, stream, element). */
typedef struct nlv_dropout {
  unsigned thr16;
  float scale;
  unsigned seed_lo, seed_hi;
  unsigned stream;
} nlv_dropout;

typedef struct nlv_gemm_args {
  const void* a;        /* A operand: [m,k] (a_major K, ld=lda) or [k,m] (a_major MN) */
  const void* b;        /* B operand: [n,k] (b_major K, ld=ldb) or [k,n] (b_major MN) */
  void* d;              /* output [m,n], row stride ldd, dtype d_dtype */
  const float* bias;    /* optional [n] */
  const void* residual; /* optional [m,n], row stride ldr, dtype r_dtype (may alias d) */
  int m, n, k;
  int lda, ldb, ldd, ldr;
  int a_major, b_major;
  int ab_dtype, d_dtype, r_dtype;
  int relu;             /* applied after bias, before residual */
  const void* gate;     /* optional [m,n] (row stride ldg, dtype gate_dtype): value kept where gate > 0, else 0 —
                           the ReLU backward fused into an input-gradient GEMM; applied before residual */
  int ldg, gate_dtype;
  float gate_scale;     /* multiplies the gated value (0 is read as 1): the 1 / (1 - p) of a dropout that followed the ReLU */
  nlv_dropout drop;     /* dropout applied after bias / ReLU, before the residual (element (row, col) of the [m,n] output) */
} nlv_gemm_args;

int nlv_gemm(const nlv_gemm_args* args, void* stream);
/* Data gradient of the 3x3 / pad 1 convolution over NHWC [r,7,7,c] maps (lib/sttran.py:342) as one implicit GEMM (the taps are 4-D TMA
 * boxes with a zero-filled halo; no column-gradient matrix, no col2im pass): dx[r*49, c_in] (fp32 or bf16) from dy bf16 [r*49, c_out]
 * and wt bf16 [c_in, 9*c_out] with k = (ky*3 + kx) * c_out + channel.  c_out % 64 == 0, c_in <= 128. */
int nlv_conv3x3_dgrad(const void* dy, long long r, int c_out, const void* wt, int c_in, void* dx, int dx_dtype, void* stream);

/* ------------------------------------------------------------------------------------------
 * Box / mask kernels
 * ------------------------------------------------------------------------------------------ */
/* lib/draw_rectangles/draw_rectangles.pyx:11-66 draw_union_boxes(bbox_pairs f32[r,8], ps) ->
 * f32[r,2,ps,ps]; bit-exact.  `offset` is added to every cell (callers pass -0.5f,
 * lib/sttran.py:166,281) */
int nlv_draw_union_boxes(const float* box_pairs, int r, int pooling_size, float offset, float* out, void* stream);
/* same, fused with the pair gather of lib/sttran.py:279-281: boxes f32[n,5] (col 0 = frame id),
 * pair_idx i64[r,2] */
int nlv_union_mask_pairs(const float* boxes, const int64_t* pair_idx, int r, int pooling_size, float offset,
                         float* out, void* stream);
/* lib/fpn/box_intersections_cpu/bbox.pyx:15-61 bbox_overlaps (float64, +1 convention) -> f64[n,k] */
int nlv_bbox_overlaps_f64(const double* boxes, int n, const double* query, int k, double* out, void* stream);


/* ------------------------------------------------------------------------------------------
 * Layout / conversion / gather kernels (pair-token construction, lib/sttran.py:381-399)
 * ------------------------------------------------------------------------------------------ */
/* 2-D strided dtype conversion (f32 <-> bf16) */
int nlv_convert(const void* src, int src_dtype, int lds, void* dst, int dst_dtype, int ldd, long long rows, int cols,
                void* stream);
/* bf16x3 operand split of an fp32 matrix (parity mode): dst gets 3 blocks along block_dim (0 rows, 1 cols);
 * pattern 0 = (hi,hi,lo) for A operands, 1 = (hi,lo,hi) for B operands */
int nlv_split3(const float* src, int lds, long long rows, int cols, void* dst_bf16, int ldd, int block_dim, int pattern,
               void* stream);
/* union_feat [r,c,7,7] (NCHW; f32 as the reference producer hands it, or bf16 from packed feature files) ->
 * [r*49, c] rows (operand of the union_func1 1x1 conv, lib/sttran.py:336,386) */
int nlv_nchw_to_rows(const void* src, int src_dtype, int r, int c, int hw, void* dst, int dst_dtype, void* stream);
/* im2col of the 2x27x27 spatial masks for Conv2d(2,128,k7,s2,p3) (lib/sttran.py:338): -> [r*196, ldd], 98 cols */
int nlv_im2col_mask(const float* masks, int r, void* dst, int dst_dtype, int ldd, void* stream);
/* im2col / col2im of the 3x3 s1 p1 conv (lib/sttran.py:342) over NHWC [r,h,w,c]; column = (ky*3+kx)*c_total + c
 * (channels innermost: both sides move whole 16-byte channel groups; the conv weight is permuted to match) */
int nlv_im2col_3x3(const void* x, int x_dtype, int r, int h, int w, int c, void* dst, int dst_dtype, void* stream);
int nlv_col2im_3x3(const void* dcol, int dtype, int r, int h, int w, int c, float* dx, void* stream);
/* MaxPool2d(3,2,1) (lib/sttran.py:341) on NHWC [r,14,14,c] -> [r,7,7,c]; argmax u8 per output */
int nlv_maxpool_fwd(const void* x, int x_dtype, int r, int c, void* y, int y_dtype, uint8_t* argmax, void* stream);
int nlv_maxpool_bwd(const float* dy, const uint8_t* argmax, int r, int c, void* dx, int dx_dtype, void* stream);
/* First stage of the spatial-mask branch on the bf16 path without the im2col matrix (csrc/maskconv.cu; lib/sttran.py:337-341):
 * nlv_mask_conv1_fwd: Conv2d(2,128,k7,s2,p3) + bias + ReLU over masks [r,2,27,27] (fp32) with w [128,98] fp32 (c,ky,kx) -> out bf16
 *   [r*196,128] (NHWC), plus the training-mode BatchNorm statistics of that map per video: pair_video[pair] = video, seg196 =
 *   int[nv+1] row offsets (as nlv_bn_stats), sums_ws = double[nv*2*128]; mean/var [nv,128] out, running statistics updated.
 *   mean == NULL: no statistics (eval mode).
 * nlv_bn_apply_maxpool: BatchNorm apply + MaxPool2d(3,2,1) in one pass, x bf16 [r*196,128] -> y bf16 [r*49,128] + argmax taps;
 *   same values as nlv_bn_apply (bf16 out) followed by nlv_maxpool_fwd (taps differ only where outputs tie after rounding).
 *   pair_video NULL: segment 0.  xmax (nullable):
 *   bf16 [r*49,128], the activation at every argmax.
 * nlv_mask_conv1_dw: dw[128,98] (assigned) = weight gradient from dy bf16 [r*196,128]; ws = float[nlv_mask_conv1_dw_ws_floats()]. */
int nlv_mask_conv1_fwd(const float* masks, const float* w, const float* bias, long long r, const int* pair_video, const int* seg196,
                       int nv, void* out, float momentum, double* sums_ws, float* mean, float* var, float* running_mean,
                       float* running_var, void* stream);
int nlv_bn_apply_maxpool(const void* x, const int* pair_video, const float* mean, const float* var, const float* w, const float* b,
                         float eps, long long r, void* y, uint8_t* argmax, void* xmax, void* stream);
/* backward of that stage from the pooled gradient dp fp32 [r*49,128]: dx bf16 [r*196,128] at the conv output (ReLU applied); dw / db
 * (BatchNorm parameter gradients) and dx_colsum (conv bias gradient, nullable) are accumulated.  xmax: the activation at every argmax
 * (bf16 [r*49,128], written by nlv_bn_apply_maxpool).  Replaces nlv_maxpool_bwd + nlv_bn_bwd_colsum(gate_by_x = 1). */
int nlv_pool_bn_bwd(const float* dp, const uint8_t* argmax, const void* x, const void* xmax, const int* pair_video, const int* seg196,
                    const int* seg49, int nv, const float* mean, const float* var, const float* w, float eps, int use_batch_stats,
                    long long r, double* sums_ws, void* dx, float* dw, float* db, float* dx_colsum, void* stream);
long long nlv_mask_conv1_dw_ws_floats(void);
int nlv_mask_conv1_dw(const void* dy, const float* masks, long long r, float* ws, float* dw, void* stream);
/* dst[i,:] = src[idx[i],:] (+ add[add_idx[i],:]); idx null = identity, idx<0 = zero row; optional 2nd output.
 * Builds the sliding-window token stream + frame position embedding (lib/transformer_wk.py:163-171) and the
 * DSG-DETR class sequences + sinusoidal encoding (lib/dsg_detr.py:545-559). */
int nlv_gather_rows(const void* src, int src_dtype, int lds, const int* idx, const float* add, const int* add_idx,
                    int ld_add, long long n_out, int cols, void* dst, int dst_dtype, int ldd, void* dst2,
                    int dst2_dtype, int ldd2, void* stream);
/* dst[r,:] = src[r,:] * row_scale[r] (fp32; dst may alias src) */
int nlv_scale_rows(const float* src, int lds, const float* row_scale, long long rows, int cols, float* dst, int ldd, void* stream);
/* dst[i,:] (+)= sum_j src[idx[i*fan+j],:] (idx<0 skipped): adjoint of a bounded-fan-out gather */
int nlv_gather_sum_rows(const float* src, int lds, const int* idx, int fan, long long n_out, int cols, float* dst,
                        int ldd, int accumulate, void* stream);
int nlv_scatter_add_rows(const float* src, int lds, const long long* idx, int idx_stride, long long n, int cols,
                         float* dst, int ldd, void* stream);
/* rel[r, 0:512|512:1024|1536:1736|1736:1936] from fo[N,1024], pair_idx i64[r,2], labels i64[N], tables [37,200] */
int nlv_assemble_tokens(const float* fo, const long long* pair_idx, const long long* labels, const float* e1,
                        const float* e2, long long r, float* rel, void* stream);
int nlv_assemble_tokens_bwd(const float* drel, const long long* pair_idx, const long long* labels, long long r,
                            float* dfo, float* de1, float* de2, void* stream);
/* lib/fpn/box_utils.py:51-63 center_size on boxes[:,1:5] -> f32[n,4] */
int nlv_center_size(const float* boxes, long long n, float* out, void* stream);
/* out[cls,c] += sum over rows of class cls of x[row,c] (bias / position-embedding gradients) */
int nlv_colsum(const void* x, int x_dtype, int ld, long long rows, int cols, const int* row_class, int n_class,
               float* out, void* stream);
int nlv_relu_mask(const void* x, int x_dtype, int ldx, const void* gate, int gate_dtype, int ldg, long long rows,
                  int cols, void* y, int y_dtype, int ldy, void* stream);
int nlv_add(const float* a, const float* b, long long n, float* y, void* stream);

/* ------------------------------------------------------------------------------------------
 * Normalisation (lib/transformer.py:15-16,45; lib/sttran.py:43,49,340,344)
 * ------------------------------------------------------------------------------------------ */
int nlv_layernorm_fwd(const float* x, long long rows, int cols, const float* w, const float* b, float eps, float* y,
                      void* y2, int y2_dtype, float* mean, float* rstd, void* stream);
int nlv_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* w,
                      long long rows, int cols, float* dx, void* dx2, int dx2_dtype, float* dw, float* db, void* stream);
int nlv_bn_stats(const void* x, int x_dtype, int ld, const int* seg, int nseg, long long rows, int c, float momentum,
                 double* sums_ws, float* mean, float* var, float* running_mean, float* running_var, void* stream);
/* row_seg[r / row_div] = statistics segment (video) of row r (row_seg NULL: segment 0); row_div = rows per indexed unit
 * (1 for per-box arrays, 49 / 196 for the per-pair conv maps) */
int nlv_bn_apply(const void* x, int x_dtype, int ldx, const int* row_seg, int row_div, const float* mean, const float* var,
                 const float* w, const float* b, float eps, int relu, long long rows, int c, void* y, int y_dtype, int ldy,
                 void* y2, int y2_dtype, int ldy2, void* stream);
int nlv_bn_bwd(const void* dy, int dy_dtype, int lddy, const void* x, int x_dtype, int ldx, const void* yout, int y_dtype, int ldy,
               const int* seg, const int* row_seg, int row_div, int nseg, const float* mean, const float* var, const float* w, float eps,
               int use_batch_stats, int gate_by_x, long long rows, int c, double* sums_ws, void* dx, int dx_dtype, int lddx,
               float* dw, float* db, void* stream);
/* nlv_bn_bwd that also accumulates the column sums of dx into dx_colsum (nullable): the bias gradient of the conv / linear layer
 * in front of the BatchNorm, produced by the kernel that writes dx instead of a separate pass */
int nlv_bn_bwd_colsum(const void* dy, int dy_dtype, int lddy, const void* x, int x_dtype, int ldx, const void* yout, int y_dtype, int ldy,
               const int* seg, const int* row_seg, int row_div, int nseg, const float* mean, const float* var, const float* w, float eps,
               int use_batch_stats, int gate_by_x, long long rows, int c, double* sums_ws, void* dx, int dx_dtype, int lddx,
               float* dw, float* db, float* dx_colsum, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused variable-length attention over contiguous segments (nn.MultiheadAttention core of
 * lib/transformer.py:22,51 and nn.TransformerEncoderLayer of lib/dsg_detr.py:502-506)
 * work: int4[n_work] = {segment first row, segment length, first row of the item inside the segment, 0}; a segment of
 * length L is covered by ceil(L/16) items.  q/k/v: row-major [rows, >= heads*hd], head h at column h*hd, hd even, <= 256.
 * lse: float[rows*heads] (written by fwd, read by bwd).  delta: float[rows*heads] workspace of bwd.
 * bf16 in / bf16 out with heads % 4 == 0 runs on mma.sync tensor cores (csrc/attn_mma.cu); every other combination of
 * fp32 / bf16 on the register-resident SIMT kernels (csrc/attn.cu; NLV_ATTN_SIMT=1 forces them).
 * ------------------------------------------------------------------------------------------ */
int nlv_attn_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                 float scale, const void* work, int n_work, void* o, int ldo, int o_dtype, float* lse, void* stream);
int nlv_attn_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                 float scale, const void* work, int n_work, const void* o, int ldo, int o_dtype, const void* dout, int lddo,
                 int do_dtype, const float* lse, float* delta, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv,
                 int dqkv_dtype, void* stream);

/* with dropout on the attention weights (drop NULL or thr16 == 0: identical to the plain entry points).  Masks are keyed by
 * (global row, head, key index inside the segment).  bf16 tensor-core path only (NLV_ERR_UNSUPPORTED otherwise). */
int nlv_attn_fwd_drop(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                      float scale, const void* work, int n_work, void* o, int ldo, int o_dtype, float* lse, const nlv_dropout* drop,
                      void* stream);
int nlv_attn_bwd_drop(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                      float scale, const void* work, int n_work, const void* o, int ldo, int o_dtype, const void* dout, int lddo,
                      int do_dtype, const float* lse, float* delta, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv,
                      int dqkv_dtype, const nlv_dropout* drop, void* stream);
/* nlv_attn_bwd_drop for a work list ordered long-first: items [0, n_long_work) are those of segments longer than 16 rows, the
 * only ones the two-kernel backward has work for (plan.py orders the lists; n_long_work < 0: unknown, every item is visited) */
int nlv_attn_bwd_sorted(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                        float scale, const void* work, int n_work, int n_long_work, const void* o, int ldo, int o_dtype, const void* dout,
                        int lddo, int do_dtype, const float* lse, float* delta, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv,
                        int dqkv_dtype, const nlv_dropout* drop, void* stream);
/* forward with the torch-1.10.1 reading of the INT key_padding_mask of lib/transformer_wk.py:154 (the mask value is ADDED to
 * the logits instead of masking): every segment keeps work[i].w padded keys in its softmax, each with logit
 * q . kpad * scale + 1 and value vpad, where kpad / vpad (f32[heads*hd]) are the K / V slices of in_proj_bias (a padded row
 * is all-zero).  kpad = vpad = NULL: identical to nlv_attn_fwd_drop.  Inference only (no backward); SIMT kernels. */
int nlv_attn_fwd_padkeys(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                         float scale, const void* work, int n_work, void* o, int ldo, int o_dtype, float* lse, const nlv_dropout* drop,
                         const float* kpad, const float* vpad, void* stream);
/* dst = keep(row, col) ? src * scale : 0 (dtype conversion allowed, dst may alias src); and the masks themselves as bytes
 * (tests): matrix sites [rows, cols], attention sites [rows, heads, nkeys] */
int nlv_dropout_apply(const void* src, int src_dtype, int lds, void* dst, int dst_dtype, int ldd, long long rows, int cols,
                      const nlv_dropout* drop, void* stream);
int nlv_dropout_mask(long long rows, int cols, const nlv_dropout* drop, unsigned char* out, void* stream);
int nlv_dropout_mask_attn(long long rows, int heads, int nkeys, const nlv_dropout* drop, unsigned char* out, void* stream);
/* LayerNorm backward in one pass over dy and x: dx, dx2 (optionally dropout-masked like nlv_layernorm_bwd_drop), dw += , db +=
 * and dprev += colsum(dx2 values) — the bias gradient of the Linear in front of the residual sum (nullable).
 * cols % 8 == 0, cols <= 2048 */
int nlv_layernorm_bwd_fused(const float* dy, const float* x, const float* mean, const float* rstd, const float* w,
                            long long rows, int cols, float* dx, void* dx2, int dx2_dtype, float* dw, float* db, float* dprev,
                            const nlv_dropout* drop, void* stream);
/* nlv_layernorm_bwd whose second output dx2 is the dropout-masked, scaled copy of dx (operand of the GEMMs behind a
 * residual dropout: d(dropout(a)) = mask * dx / (1 - p)); dx itself stays unmasked (the residual branch) */
int nlv_layernorm_bwd_drop(const float* dy, const float* x, const float* mean, const float* rstd, const float* w,
                           long long rows, int cols, float* dx, void* dx2, int dx2_dtype, float* dw, float* db,
                           const nlv_dropout* drop, void* stream);

/* ------------------------------------------------------------------------------------------
 * Heads, losses, optimiser (lib/sttran.py:404-409; tools/train_STTran.py:169-195; lib/AdamW.py:52-114)
 * ------------------------------------------------------------------------------------------ */
int nlv_heads_activation(const float* logits, long long r, float* att, float* spa, float* con, void* stream);
int nlv_ce_loss(const float* logits, int ld, int c, const long long* labels, const float* row_weight, long long rows,
                float* loss, float* dlogits, int ldd, void* stream);
int nlv_bce_sigmoid_loss(const float* logits, int ld, int c, const unsigned* label_bits, const float* row_weight,
                         long long rows, float* loss, float* dlogits, int ldd, void* stream);
int nlv_sumsq(const float* x, long long n, float* out, void* stream);
int nlv_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, const float* total_sq, float max_norm, void* p_bf16, void* stream);

/* The same update with the step counter on the device, so an iteration can be skipped without a host sync
 * (lib/utils.py:3-11 check_valid_iter + the `continue` of tools/train_STTran.py:191): state[0] = steps applied (bias
 * corrections use state[0]+1), state[1] = steps skipped.  Skipped when *skip_flag != 0 or *total_sq is not finite.
 * grad_scale multiplies every gradient (1/world after a summing all-reduce; total_sq is the norm of the summed gradients).
 * Call nlv_adamw_step_state on every parameter range of the step, then nlv_adamw_finish once. */
int nlv_adamw_step_state(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                         float weight_decay, const int* state, const float* total_sq, const int* skip_flag, float grad_scale,
                         float max_norm, void* p_bf16, void* stream);
int nlv_adamw_finish(int* state, const float* total_sq, const int* skip_flag, void* stream);
/* flag[0] = 1 if any of x[0:n] is NaN / Inf, else 0 (n small: the loss scalar) */
int nlv_flag_nonfinite(const float* x, int n, int* flag, void* stream);

/* ------------------------------------------------------------------------------------------
 * Recall@K triplet matching (lib/evaluation_recall.py:397-467, :209-353, :630-773; bbox.pyx:21-61)
 * One launch over n_frames frames (any number of videos).  Per frame f:
 *   pairs  [pair_off[f], pair_off[f+1])   : pair_sub/pair_obj index the pred_* arrays; att (softmaxed) / spa / con scores
 *   gt rels [gtrel_off[f], gtrel_off[f+1]): (sub, obj, predicate) with sub/obj local to the frame's gt boxes
 *   gt boxes [gtbox_off[f], gtbox_off[f+1]): class + box (rounded to f32 as evaluation_recall.py:765 does)
 * out: u32[n_frames, 3 protocols (with, no, semi constraint), 3 K (10,20,50), 8] = 256-bit sets of matched gt relations.
 * Limits per frame (nlv_recall_limits): 40 pairs, 256 gt relations, 64 gt boxes.
 * ------------------------------------------------------------------------------------------ */
int nlv_recall_match(int n_frames, const int* pair_off, const int* gtrel_off, const int* gtbox_off, const int* pair_sub,
                     const int* pair_obj, const float* att, const float* spa, const float* con, const float* obj_scores,
                     const int* pred_cls, const float* pred_boxes, const int* gt_rel, const int* gt_cls,
                     const float* gt_boxes, unsigned* out, void* stream);
int nlv_recall_limits(int* p_max, int* g_max, int* gb_max);

/* ------------------------------------------------------------------------------------------
 * Native surface 1 of the reference: fasterRCNN/lib/model/csrc/vision.cpp:7-13 (pybind11 `_C`)
 * ------------------------------------------------------------------------------------------ */
/* _C.roi_align_forward(input[b,c,h,w], rois[r,5], spatial_scale, ph, pw, sampling_ratio) -> [r,c,ph,pw]
 * (ROIAlign.h:11-27, ROIAlign_cuda.cu:65-122); bit-identical to the reference CPU kernel */
int nlv_roi_align_fwd(const float* input, int b, int c, int h, int w, const float* rois, int r, float spatial_scale, int ph,
                      int pw, int sampling_ratio, float* out, void* stream);
/* _C.roi_align_backward(grad, rois, scale, ph, pw, b, c, h, w, sampling_ratio) -> [b,c,h,w] (ROIAlign.h:29-45);
 * dinput must be zeroed by the caller */
int nlv_roi_align_bwd(const float* grad, const float* rois, int r, float spatial_scale, int ph, int pw, int b, int c, int h, int w,
                      int sampling_ratio, float* dinput, void* stream);
/* _C.nms(dets[n,4], scores[n], thr) (nms.h:10-28, nms.cu:23-131): `order` = argsort(scores, descending);
 * keep_flags u8[n] (zeroed by the caller) marks kept ORIGINAL indices; mask_ws u64[n*ceil(n/64)] */
int nlv_nms(const float* dets, const long long* order, int n, float thr, int strict, unsigned long long* mask_ws,
            unsigned char* keep_flags, void* stream);
/* lib/matcher.py:102-150 HungarianMatcher cost matrix for one frame (detections x live tracks) */
int nlv_track_cost(const float* det_box_xywh, const float* trk_box_xywh, const float* det_feat, const float* trk_feat, int feat_dim,
                   const float* det_dist, const float* trk_dist, int dist_dim, int n_det, int n_trk, float w_class, float w_feat,
                   float w_bbox, float w_giou, float* cost, float* cost_dist, float* cost_feat, void* stream);
/* scipy.optimize.linear_sum_assignment(cost) of lib/matcher.py:147-149 on the device (one warp, float64 duals, the same
 * shortest-augmenting-path algorithm and tie rules -> identical assignments): match_of_row[i] = column assigned to row i
 * or -1; at most 1024 rows / columns */
int nlv_lsap(const float* cost, int n_rows, int n_cols, int ld, int* match_of_row, void* stream);
/* lib/track.py:127-262 get_sequence(task="sgcls") for a batch of videos in ONE launch (one CTA per video): per key frame the
 * matcher cost (lib/matcher.py:124-145), the assignment, the tau = 0.5 accept rule, cluster / track bookkeeping in the
 * reference's order and the 50-frame track expiry.  boxes f32[N,5] (frame, x1, y1, x2, y2), cls = argmax of the
 * distribution; det_off / frame_off int[V+1]; frame_start: per video T+1 offsets relative to its first detection
 * (concatenated); frame_key: frame number per key frame; cost_off[v]: offset of video v in a cost plane of cost_elems
 * (>= max detections per frame x detections of the video).  Outputs: cluster_of_det int[N] (cluster ids in creation
 * order), n_clusters int[V], status int[V] (1: a frame exceeded the 1024-wide assignment state). */
long long nlv_track_sequence_workspace(long long n_det_total, int feat_dim, int n_cls, long long cost_elems);
int nlv_track_sequence(const float* boxes, const float* feats, int feat_dim, const int* cls, int n_cls, const int* det_off,
                       const int* frame_off, const int* frame_start, const int* frame_key, int n_videos, long long n_det_total,
                       int max_det_per_frame, float img_w, float img_h, float w_class, float w_feat, float w_bbox, float w_giou,
                       int max_gap, const long long* cost_off, long long cost_elems, void* workspace, int* cluster_of_det,
                       int* n_clusters, int* status, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small multi-segment helpers used by the model sequencer
 * ------------------------------------------------------------------------------------------ */
/* up to any number of independent contiguous conversions/copies dst[i][0:n[i]] = src[i][0:n[i]] (f32/bf16 either side) in
 * ceil(count/32) launches; src/dst/n are HOST arrays of device pointers / element counts */
int nlv_convert_multi(const void* const* src_host, void* const* dst_host, const long long* n_host, int count, int src_dtype,
                      int dst_dtype, void* stream);
/* dst[a, c, b] = src[a, b, c] (dtype conversion allowed): the conv / vr_fc weight re-orderings between the reference's
 * NCHW parameter layout and the channels-last operand layout of the kernels, and back for their gradients */
int nlv_permute_021(const void* src, int src_dtype, int a, int b, int c, void* dst, int dst_dtype, void* stream);
int nlv_zero_bytes(void* p, long long nbytes, void* stream);
/* Packed per-video feature files (nlvsgg_b200/featfile.py; replaces the per-frame dets.npy / feat.npy of
 * lib/assign_pseudo_label.py:27-45): zero-suppressed channels-last union features -> dense bf16 rows [rows, 2048].
 * bitmap u64[rows,32] (bit c of word w = channel 64 w + c is stored), off u32[rows+1] = first value of each row, vals bf16
 * (16-byte aligned, padded by 16 bytes). */
int nlv_union_unpack(const void* bitmap, const unsigned* off, const void* vals, long long rows, void* dst_bf16, void* stream);
/* the same for 12-bit stored values (featfile.py "sparse12"): value i = (base[row] + code_i) << 8 | lo[i], code_i = nibble i & 1 of
 * hx[i >> 1]; lo / hx 16-byte aligned and readable 32 bytes past their ends */
int nlv_union_unpack12(const void* bitmap, const unsigned* off, const void* lo, const void* hx, const unsigned char* base, long long rows,
                       void* dst_bf16, void* stream);
/* dst_bf16[pos[i]] = val[i] (bf16 bits): the exception list of the 12-bit encoding */
int nlv_union_patch(void* dst_bf16, const unsigned* pos, const unsigned short* val, int n, void* stream);
/* lib/assign_pseudo_label.py:934-938 create_dis on device: out f32[n,36] = conf at idx, `other` elsewhere (and at idx when
 * conf == 0); other NULL -> (1 - conf) / 35 in fp32 arithmetic */
int nlv_create_dis(const float* conf, const float* other, const int* idx, long long n, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Whole-model sequencer: one call enqueues the complete kernel sequence of
 *   lib/sttran.py:375-411 STTran.forward (object classifier :173-184, pair tokens :381-399, lib/transformer_wk.py:130-217,
 *   heads :404-409), lib/dsg_detr.py:514-572, the losses of tools/train_STTran.py:169-189 and the matching backward pass.
 * The host does no per-layer work: buffers come from ONE caller-provided workspace (bump-allocated, sized by a dry run),
 * parameters / gradients are pointer tables indexed by the NLV_P_* slots below.
 * ------------------------------------------------------------------------------------------ */
enum {
  NLV_P_OC_EMBED = 0,                                              /* object_classifier.obj_embed.weight [36,200]   */
  NLV_P_OC_BN0_W, NLV_P_OC_BN0_B, NLV_P_OC_BN0_RM, NLV_P_OC_BN0_RV, /* object_classifier.pos_embed.0 (BatchNorm1d(4))  */
  NLV_P_OC_LIN1_W, NLV_P_OC_LIN1_B,                                /* object_classifier.pos_embed.1 [128,4]         */
  NLV_P_OC_DEC0_W, NLV_P_OC_DEC0_B,                                /* object_classifier.decoder_lin.0 [1024,2376]   */
  NLV_P_OC_BN1_W, NLV_P_OC_BN1_B, NLV_P_OC_BN1_RM, NLV_P_OC_BN1_RV, /* object_classifier.decoder_lin.1               */
  NLV_P_OC_DEC3_W, NLV_P_OC_DEC3_B,                                /* object_classifier.decoder_lin.3 [37,1024]     */
  NLV_P_UNION_W, NLV_P_UNION_B,                                    /* union_func1 [256,2048,1,1]                    */
  NLV_P_CONV0_W, NLV_P_CONV0_B,                                    /* conv.0 [128,2,7,7]                            */
  NLV_P_BN2_W, NLV_P_BN2_B, NLV_P_BN2_RM, NLV_P_BN2_RV,            /* conv.2 BatchNorm2d(128)                       */
  NLV_P_CONV4_W, NLV_P_CONV4_B,                                    /* conv.4 [256,128,3,3]                          */
  NLV_P_BN6_W, NLV_P_BN6_B, NLV_P_BN6_RM, NLV_P_BN6_RV,            /* conv.6 BatchNorm2d(256)                       */
  NLV_P_SUBJ_W, NLV_P_SUBJ_B, NLV_P_OBJ_W, NLV_P_OBJ_B,            /* subj_fc / obj_fc [512,2048]                   */
  NLV_P_VR_W, NLV_P_VR_B,                                          /* vr_fc [512,12544]                             */
  NLV_P_EMB1, NLV_P_EMB2,                                          /* obj_embed / obj_embed2 [37,200]               */
  NLV_P_A_W, NLV_P_A_B, NLV_P_S_W, NLV_P_S_B, NLV_P_C_W, NLV_P_C_B, /* a/s/c_rel_compress                            */
  NLV_P_POS,                                                       /* STTran: glocal_transformer.position_embedding.weight [2,1936];
                                                                      DSG-DETR: positional_encoder.pe [max_len,1936] (buffer) */
  NLV_P_LAYER0                                                     /* first transformer layer; NLV_P_LAYER_STRIDE slots each */
};
enum {
  NLV_L_INPROJ_W = 0, NLV_L_INPROJ_B, NLV_L_OUTPROJ_W, NLV_L_OUTPROJ_B, NLV_L_LIN1_W, NLV_L_LIN1_B, NLV_L_LIN2_W, NLV_L_LIN2_B,
  NLV_L_NORMA_W, NLV_L_NORMA_B,   /* encoder: norm1; decoder: norm3 */
  NLV_L_NORMB_W, NLV_L_NORMB_B,   /* encoder: norm2; decoder: unused */
  NLV_P_LAYER_STRIDE
};
/* layer order: STTran = n_enc spatial-encoder layers, then n_dec temporal-decoder layers;
 *              DSG-DETR = local_transformer layer, then the 3 global_transformer layers (all encoder type) */

#define NLV_ARCH_STTRAN 0
#define NLV_ARCH_DSG 1
#define NLV_MODE_PREDCLS 0
#define NLV_MODE_SGCLS 1
#define NLV_MODE_SGDET 2
#define NLV_PREC_BF16 0
#define NLV_PREC_BF16X3 1
#define NLV_PREC_FP32 2

typedef struct nlv_model {
  int arch, mode, precision;
  int n_enc, n_dec;            /* STTran layer counts (DSG-DETR: 1, 3) */
  int training;                /* BatchNorm batch statistics + running-stat update, dropout */
  int n_slots;                 /* entries in the three tables below */
  const float* const* params;  /* fp32 parameters / buffers per slot (device pointers; host array) */
  const void* const* params_op;/* optional bf16 operand copies per slot kept current by the caller (trainer mirror); NULL
                                  table or NULL entries -> converted inside the forward call */
  float* grad_base;            /* backward: ONE fp32 gradient buffer, zeroed by the backward call ... */
  long long grad_elems;
  const long long* grad_offset;/* ... parameter of slot s receives its gradient at grad_base + grad_offset[s] (<0: none) */
  float dropout_p;             /* reference 0.1 (lib/transformer.py:11-57); applied only when training != 0 */
  unsigned long long seed;     /* Philox seed of this step's dropout masks */
  int additive_mask;           /* 0: bool key-padding masks (lib/transformer.py:144); 1: the int mask of
                                  lib/transformer_wk.py:154 under torch 1.10.1 (+1 added to padded keys' logits; the number of
                                  padded keys of a frame travels in the 4th field of its local work items; inference only) */
  int pe_rows;                 /* DSG-DETR: rows of the positional-encoding buffer */
  int transformer_both;        /* STTran temporal decoder output: 0 = mode 'latter' (lib/transformer_wk.py:209-215, what lib/sttran.py:358
                                  uses), 1 = mode 'both' (:197-207: a frame inside the video is the mean of its two windows) */
} nlv_model;

typedef struct nlv_batch {
  int nv;                       /* videos in the batch */
  long long n_boxes, n_pairs, n_stream;   /* N, R, Mg (rows of the sliding-window stream) */
  const void* features; int feat_dtype;   /* [N,2048] f32 (entry contract) or bf16 (packed feature files) */
  const float* boxes;                     /* [N,5] */
  const long long* labels;                /* [N] labels used for the semantic embeddings (pred_labels) */
  const float* distribution;              /* [N,36] (sgdet / sgcls) */
  const void* union_feat; int union_dtype;
  int union_rows;                         /* 0: NCHW [R,2048,7,7] (entry contract); 1: channels-last rows [R*49,2048];
                                             2: zero-suppressed rows: union_feat = bf16 values, + union_bitmap / union_off;
                                             3: the same with 12-bit values (union_hx / union_base below) */
  const float* spatial_masks;             /* [R,2,27,27] or NULL -> rasterised from boxes + pair_idx */
  const long long* pair_idx;              /* [R,2] */
  /* host-built descriptors (nlvsgg_b200/plan.py), int32 device arrays */
  const int *box_seg, *seg196, *seg49;    /* [nv+1] row offsets of every video in the box / 196-per-pair / 49-per-pair arrays */
  const int *box_row, *pair_row;          /* [N], [R]: video of a box / pair (NULL when nv == 1) */
  const int *local_work; int n_local_work;
  const int *glob_work; int n_glob_work;
  const int *stream_src, *stream_slot, *inv, *out_src, *out_inv, *passthrough; int has_passthrough;
  const int *cls_perm, *cls_iperm, *cls_pos, *cls_work; int n_cls_work;   /* DSG-DETR class sequences */
  /* packed-file inputs (nlvsgg_b200/featfile.py): occupancy words + row offsets of zero-suppressed union features;
   * (confidence, class) per box from which `distribution` is rebuilt on device when distribution == NULL */
  const void* union_bitmap; const unsigned* union_off;
  const float* dist_conf; const float* dist_other; const int* dist_idx;
  /* fused-loss labels (tools/train_STTran.py:143-167), NULL for inference */
  const long long* lab_att; const float* w_att; const unsigned* spa_bits; const float* w_spa;
  const unsigned* con_bits; const float* w_con; const float* w_obj;
  const float* both_w;          /* mode 'both': f32[R] = 1 / (windows the token appears in), 0 for tokens without a window */
  /* work_sorted != 0: the three work lists are ordered long-first and n_*_long = their items of segments longer than 16 rows */
  int work_sorted, n_local_long, n_glob_long, n_cls_long;
  /* union_rows == 3: zero-suppressed rows with 12-bit stored values: union_feat = low bytes u8 [nnz], union_hx = 4-bit codes
   * (two per byte; high byte = union_base[row] + code), union_base = u8 [R*49] */
  const void* union_hx; const unsigned char* union_base;
  /* values outside their row's 16-step window: element index (row * 2048 + channel) and bf16 bits, written over the decoded rows */
  const unsigned* union_exc_pos; const unsigned short* union_exc_val; int n_union_exc;
} nlv_batch;

typedef struct nlv_outputs {
  float* obj_logits;      /* [N,37] (NULL in predcls) */
  float* logits26;        /* [R,26] attention logits | spatial logits | contacting logits */
  float* att; float* spa; float* con;   /* [R,3] logits, [R,6] / [R,17] sigmoid (NLV_RUN_ACTIVATIONS) */
  float* loss;            /* [1] (NLV_RUN_LOSS) */
  float* masks;           /* [R,2,27,27] spatial masks used (the caller's, or the rasterised ones) */
  float* rel_tokens;      /* [R,1936] pair tokens (lib/sttran.py:399) */
  float* rel_out;         /* [R,1936] transformer output (lib/sttran.py:401) */
  float* d26; float* dobj;/* loss gradients w.r.t. logits26 / obj_logits (NLV_RUN_LOSS) */
} nlv_outputs;

#define NLV_RUN_CTX 1          /* keep activations for nlv_session_backward */
#define NLV_RUN_LOSS 2         /* fused CE / CE / BCE / BCE loss from the batch labels (+ its gradients when NLV_RUN_CTX) */
#define NLV_RUN_BACKWARD 4     /* (plan only) size the workspace for a backward pass too */
#define NLV_RUN_ACTIVATIONS 8  /* att / spa / con outputs (lib/sttran.py:404-409) */
#define NLV_RUN_OBJECT_ONLY 16 /* stop after the object classifier head (obj_logits only): first stage of the sgcls test branch */

/* out[0..3] = sizeof(nlv_model), sizeof(nlv_batch), sizeof(nlv_outputs), sizeof(nlv_gemm_args); returns 4 */
int nlv_struct_sizes(int* out, int n);

/* diagnostic per-call timing of the sequencer (bench.py roofline legs): nlv_profile(1) records CUDA events around every
 * kernel entry; nlv_profile_read synchronises, writes "name\tms\tflops\tunits\tm\tn\tk\tdtype\n" lines into buf (NULL: size only),
 * clears the records and returns the bytes needed */
int nlv_profile(int on);
long long nlv_profile_read(char* buf, long long cap);

typedef struct nlv_session nlv_session;
nlv_session* nlv_session_create(void);
void nlv_session_destroy(nlv_session* s);
/* dry run: workspace bytes needed by forward (+ backward with NLV_RUN_BACKWARD) for this model / batch; < 0 on error */
long long nlv_session_plan(nlv_session* s, const nlv_model* model, const nlv_batch* batch, int flags);
int nlv_session_forward(nlv_session* s, const nlv_model* model, const nlv_batch* batch, void* workspace,
                        long long workspace_bytes, int flags, nlv_outputs* out, void* stream);
/* where the next backward writes (replaces model->grad_base / grad_elems / grad_offset given at forward time) */
int nlv_session_set_gradients(nlv_session* s, float* grad_base, long long grad_elems, const long long* grad_offset, int n_slots);
/* gradients of every parameter into model->grad_base; d26 / dobj NULL -> the loss's own gradients (NLV_RUN_LOSS) */
int nlv_session_backward(nlv_session* s, const float* d26, const float* dobj, void* stream);
/* fn(user) is called from inside nlv_session_backward, on the calling thread, right after the last kernel that writes a gradient
 * of the temporal-decoder (STTran) / global (DSG-DETR) layers has been enqueued — those layers are the tail of the gradient
 * buffer, so a data-parallel caller can start their all-reduce on another stream while the rest of the backward pass runs
 * (bucketed and overlapped).  NULL removes the hook. */
int nlv_session_set_tail_hook(nlv_session* s, void (*fn)(void*), void* user);
/* standalone spatio-temporal transformer (lib/transformer_wk.py:130-217 forward(features, im_idx)) over the same
 * session machinery: only the layer slots and NLV_P_POS of `model` are read; x f32[R,1936] -> *out f32[R,1936] (in the
 * workspace).  workspace == NULL: dry run, returns the workspace bytes needed (with NLV_RUN_BACKWARD: backward included) */
long long nlv_session_transformer_forward(nlv_session* s, const nlv_model* model, const nlv_batch* batch, const float* x,
                                          void* workspace, long long workspace_bytes, int flags, float** out, void* stream);
int nlv_session_transformer_backward(nlv_session* s, const float* dout, float** dx, void* stream);
/* d26[r, 0:3|3:9|9:26] = datt | dspa * spa(1-spa) | dcon * con(1-con): chain rule of lib/sttran.py:404-409 for callers that
 * differentiate the activated outputs (any of datt / dspa / dcon may be NULL = zero) */
int nlv_heads_activation_bwd(const float* datt, const float* dspa, const float* dcon, const float* spa, const float* con,
                             long long r, float* d26, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NLV_B200_H_ */
