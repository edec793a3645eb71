// Box / mask kernels: union-box mask rasteriser (bit-exact), float64 IoU.
#include "common.cuh"

namespace nlv {
namespace {

__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.f), 1.f); }

// One CTA per pair, threads over the 2*ps*ps cells.  Arithmetic is written with explicit
// round-to-nearest intrinsics (no FMA contraction) in the reference's operation order
// (lib/draw_rectangles/draw_rectangles.pyx:45-65) so the result is bit-identical to it.
template <bool GATHER>
__global__ void union_mask_kernel(const float* __restrict__ src, const int64_t* __restrict__ pair_idx, int r, int ps,
                                  float offset, float* __restrict__ out) {
  const int n = blockIdx.x;
  if (n >= r) return;
  __shared__ float b[8];
  if (threadIdx.x < 8) {
    if (GATHER) {
      const int which = threadIdx.x >> 2;
      const int64_t row = pair_idx[2 * (size_t)n + which];
      b[threadIdx.x] = src[row * 5 + 1 + (threadIdx.x & 3)];
    } else {
      b[threadIdx.x] = src[(size_t)n * 8 + threadIdx.x];
    }
  }
  __syncthreads();
  const float x1u = fminf(b[0], b[4]), y1u = fminf(b[1], b[5]);
  const float x2u = fmaxf(b[2], b[6]), y2u = fmaxf(b[3], b[7]);
  const float w = __fsub_rn(x2u, x1u), h = __fsub_rn(y2u, y1u);
  const float fps = (float)ps;
  const int cells = 2 * ps * ps;
  for (int c = threadIdx.x; c < cells; c += blockDim.x) {
    const int i = c / (ps * ps);
    const int rem = c - i * ps * ps;
    const int j = rem / ps, k = rem - j * ps;
    const float x1 = __fdiv_rn(__fmul_rn(__fsub_rn(b[0 + 4 * i], x1u), fps), w);
    const float y1 = __fdiv_rn(__fmul_rn(__fsub_rn(b[1 + 4 * i], y1u), fps), h);
    const float x2 = __fdiv_rn(__fmul_rn(__fsub_rn(b[2 + 4 * i], x1u), fps), w);
    const float y2 = __fdiv_rn(__fmul_rn(__fsub_rn(b[3 + 4 * i], y1u), fps), h);
    const float yc = __fmul_rn(clamp01(__fsub_rn((float)(j + 1), y1)), clamp01(__fsub_rn(y2, (float)j)));
    const float xc = __fmul_rn(clamp01(__fsub_rn((float)(k + 1), x1)), clamp01(__fsub_rn(x2, (float)k)));
    out[(size_t)n * cells + c] = __fadd_rn(__fmul_rn(xc, yc), offset);
  }
}

__global__ void bbox_overlaps_f64_kernel(const double* __restrict__ boxes, int n, const double* __restrict__ query, int k,
                                         double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * k) return;
  const int i = (int)(idx / k), q = (int)(idx % k);
  const double* bb = boxes + 4 * (size_t)i;
  const double* qb = query + 4 * (size_t)q;
  double v = 0.0;
  const double qa = __dmul_rn(__dadd_rn(__dsub_rn(qb[2], qb[0]), 1.0), __dadd_rn(__dsub_rn(qb[3], qb[1]), 1.0));
  const double iw = __dadd_rn(__dsub_rn(fmin(bb[2], qb[2]), fmax(bb[0], qb[0])), 1.0);
  if (iw > 0) {
    const double ih = __dadd_rn(__dsub_rn(fmin(bb[3], qb[3]), fmax(bb[1], qb[1])), 1.0);
    if (ih > 0) {
      const double ba = __dmul_rn(__dadd_rn(__dsub_rn(bb[2], bb[0]), 1.0), __dadd_rn(__dsub_rn(bb[3], bb[1]), 1.0));
      const double inter = __dmul_rn(iw, ih);
      const double ua = __dsub_rn(__dadd_rn(ba, qa), inter);
      v = __ddiv_rn(inter, ua);
    }
  }
  out[idx] = v;
}

}  // namespace
}  // namespace nlv

extern "C" {

int nlv_draw_union_boxes(const float* box_pairs, int r, int pooling_size, float offset, float* out, void* stream) {
  using namespace nlv;
  NLV_CHECK_ARG(r >= 0 && pooling_size > 0, "draw_union_boxes: bad sizes r=%d ps=%d", r, pooling_size);
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(box_pairs && out, "draw_union_boxes: null pointer");
  union_mask_kernel<false><<<r, 256, 0, (cudaStream_t)stream>>>(box_pairs, nullptr, r, pooling_size, offset, out);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_union_mask_pairs(const float* boxes, const int64_t* pair_idx, int r, int pooling_size, float offset, float* out,
                         void* stream) {
  using namespace nlv;
  NLV_CHECK_ARG(r >= 0 && pooling_size > 0, "union_mask_pairs: bad sizes r=%d ps=%d", r, pooling_size);
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(boxes && pair_idx && out, "union_mask_pairs: null pointer");
  union_mask_kernel<true><<<r, 256, 0, (cudaStream_t)stream>>>(boxes, pair_idx, r, pooling_size, offset, out);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_bbox_overlaps_f64(const double* boxes, int n, const double* query, int k, double* out, void* stream) {
  using namespace nlv;
  NLV_CHECK_ARG(n >= 0 && k >= 0, "bbox_overlaps: bad sizes");
  if (n == 0 || k == 0) return NLV_OK;
  NLV_CHECK_ARG(boxes && query && out, "bbox_overlaps: null pointer");
  const long long total = (long long)n * k;
  bbox_overlaps_f64_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(boxes, n, query, k, out);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
}
