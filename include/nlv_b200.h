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
} nlv_gemm_args;

int nlv_gemm(const nlv_gemm_args* args, void* stream);

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
/* dst[i,:] = src[idx[i],:] (+ add[add_idx[i],:]); idx null = identity, idx<0 = zero row; optional 2nd output.
 * Builds the sliding-window token stream + frame position embedding (lib/transformer_wk.py:163-171) and the
 * DSG-DETR class sequences + sinusoidal encoding (lib/dsg_detr.py:545-559). */
int nlv_gather_rows(const void* src, int src_dtype, int lds, const int* idx, const float* add, const int* add_idx,
                    int ld_add, long long n_out, int cols, void* dst, int dst_dtype, int ldd, void* dst2,
                    int dst2_dtype, int ldd2, void* stream);
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
int nlv_bn_apply(const void* x, int x_dtype, int ldx, const int* row_seg, const float* mean, const float* var,
                 const float* w, const float* b, float eps, int relu, long long rows, int c, void* y, int y_dtype, int ldy,
                 void* y2, int y2_dtype, int ldy2, void* stream);
int nlv_bn_bwd(const void* dy, int dy_dtype, int lddy, const void* x, int x_dtype, int ldx, const void* yout, int y_dtype, int ldy,
               const int* seg, const int* row_seg, int nseg, const float* mean, const float* var, const float* w, float eps,
               int use_batch_stats, int gate_by_x, long long rows, int c, double* sums_ws, void* dx, int dx_dtype, int lddx,
               float* dw, float* db, void* stream);

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

#ifdef __cplusplus
}
#endif
#endif /* NLV_B200_H_ */
