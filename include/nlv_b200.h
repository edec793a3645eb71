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

#ifdef __cplusplus
}
#endif
#endif /* NLV_B200_H_ */
