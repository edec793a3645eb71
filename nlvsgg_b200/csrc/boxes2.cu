// RoIAlign forward/backward, NMS, and the DSG-DETR tracking cost matrix.
//
// RoIAlign : fasterRCNN/lib/model/csrc/cuda/ROIAlign_cuda.cu:65-122 (forward), :178-254 (backward) — aligned=False,
//            malformed ROIs forced to 1x1, adaptive sampling grid when sampling_ratio <= 0.  The forward is written with
//            explicit round-to-nearest intrinsics in the operation order of the reference CPU kernel
//            (cpu/ROIAlign_cpu.cpp:114-219), so it is bit-identical to it (and to torchvision.ops.roi_align(aligned=False)).
// NMS      : fasterRCNN/lib/model/csrc/cuda/nms.cu:23-131 (suppress when IoU > thr; `strict=0` gives the CPU rule >=,
//            cpu/nms_cpu.cpp:60); +1 pixel areas; result = flags per ORIGINAL index (callers take nonzero() -> ascending).
// Cost     : lib/matcher.py:102-150 HungarianMatcher cost  C = wc*(1-cos(dist)) + wf*(1-cos(feat)) + wb*L1(cxcywh) - wg*GIoU
#include "common.cuh"

namespace nlv {
namespace {

// ---------------------------------------------------------------------------------------------------------
struct Bilinear { int p1, p2, p3, p4; float w1, w2, w3, w4; bool valid; };

__device__ __forceinline__ Bilinear bilinear(float y, float x, int H, int W) {
  Bilinear b;
  b.valid = !(y < -1.0f || y > (float)H || x < -1.0f || x > (float)W);
  if (!b.valid) { b.p1 = b.p2 = b.p3 = b.p4 = 0; b.w1 = b.w2 = b.w3 = b.w4 = 0.f; return b; }
  if (y <= 0) y = 0;
  if (x <= 0) x = 0;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  const float ly = __fsub_rn(y, (float)yl), lx = __fsub_rn(x, (float)xl);
  const float hy = __fsub_rn(1.f, ly), hx = __fsub_rn(1.f, lx);
  b.w1 = __fmul_rn(hy, hx); b.w2 = __fmul_rn(hy, lx); b.w3 = __fmul_rn(ly, hx); b.w4 = __fmul_rn(ly, lx);
  b.p1 = yl * W + xl; b.p2 = yl * W + xh; b.p3 = yh * W + xl; b.p4 = yh * W + xh;
  return b;
}

struct RoiGeom { int batch, gh, gw; float sh, sw, bh, bw, count; };

__device__ __forceinline__ RoiGeom roi_geom(const float* roi, float scale, int PH, int PW, int sampling_ratio) {
  RoiGeom g;
  g.batch = (int)roi[0];
  g.sw = __fmul_rn(roi[1], scale); g.sh = __fmul_rn(roi[2], scale);
  const float ew = __fmul_rn(roi[3], scale), eh = __fmul_rn(roi[4], scale);
  const float rw = fmaxf(__fsub_rn(ew, g.sw), 1.f), rh = fmaxf(__fsub_rn(eh, g.sh), 1.f);
  g.bh = __fdiv_rn(rh, (float)PH); g.bw = __fdiv_rn(rw, (float)PW);
  g.gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rh, (float)PH));
  g.gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(__fdiv_rn(rw, (float)PW));
  g.count = (float)(g.gh * g.gw);
  return g;
}

__global__ void roi_align_fwd_kernel(const float* __restrict__ in, int C, int H, int W, const float* __restrict__ rois,
                                     long long total, float scale, int PH, int PW, int sampling_ratio, float* __restrict__ out) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % PW), ph = (int)((idx / PW) % PH);
    const int c = (int)((idx / ((long long)PW * PH)) % C);
    const long long n = idx / ((long long)PW * PH * C);
    const RoiGeom g = roi_geom(rois + 5 * n, scale, PH, PW, sampling_ratio);
    const float* src = in + ((size_t)g.batch * C + c) * H * W;
    float acc = 0.f;
    for (int iy = 0; iy < g.gh; ++iy) {
      const float yy = __fadd_rn(__fadd_rn(g.sh, __fmul_rn((float)ph, g.bh)), __fdiv_rn(__fmul_rn((float)iy + .5f, g.bh), (float)g.gh));
      for (int ix = 0; ix < g.gw; ++ix) {
        const float xx = __fadd_rn(__fadd_rn(g.sw, __fmul_rn((float)pw, g.bw)), __fdiv_rn(__fmul_rn((float)ix + .5f, g.bw), (float)g.gw));
        const Bilinear b = bilinear(yy, xx, H, W);
        if (!b.valid) continue;
        const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(b.w1, src[b.p1]), __fmul_rn(b.w2, src[b.p2])),
                                            __fmul_rn(b.w3, src[b.p3])), __fmul_rn(b.w4, src[b.p4]));
        acc = __fadd_rn(acc, s);
      }
    }
    out[idx] = __fdiv_rn(acc, g.count);
  }
}

__global__ void roi_align_bwd_kernel(const float* __restrict__ grad, int C, int H, int W, const float* __restrict__ rois,
                                     long long total, float scale, int PH, int PW, int sampling_ratio, float* __restrict__ din) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % PW), ph = (int)((idx / PW) % PH);
    const int c = (int)((idx / ((long long)PW * PH)) % C);
    const long long n = idx / ((long long)PW * PH * C);
    const RoiGeom g = roi_geom(rois + 5 * n, scale, PH, PW, sampling_ratio);
    float* dst = din + ((size_t)g.batch * C + c) * H * W;
    const float gv = grad[idx] / g.count;
    for (int iy = 0; iy < g.gh; ++iy) {
      const float yy = __fadd_rn(__fadd_rn(g.sh, __fmul_rn((float)ph, g.bh)), __fdiv_rn(__fmul_rn((float)iy + .5f, g.bh), (float)g.gh));
      for (int ix = 0; ix < g.gw; ++ix) {
        const float xx = __fadd_rn(__fadd_rn(g.sw, __fmul_rn((float)pw, g.bw)), __fdiv_rn(__fmul_rn((float)ix + .5f, g.bw), (float)g.gw));
        const Bilinear b = bilinear(yy, xx, H, W);
        if (!b.valid) continue;
        atomicAdd(dst + b.p1, gv * b.w1); atomicAdd(dst + b.p2, gv * b.w2);
        atomicAdd(dst + b.p3, gv * b.w3); atomicAdd(dst + b.p4, gv * b.w4);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float iou_plus1(const float* a, const float* b) {
  const float xx1 = fmaxf(a[0], b[0]), yy1 = fmaxf(a[1], b[1]);
  const float xx2 = fminf(a[2], b[2]), yy2 = fminf(a[3], b[3]);
  const float w = fmaxf(__fadd_rn(__fsub_rn(xx2, xx1), 1.f), 0.f), h = fmaxf(__fadd_rn(__fsub_rn(yy2, yy1), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  const float sa = __fmul_rn(__fadd_rn(__fsub_rn(a[2], a[0]), 1.f), __fadd_rn(__fsub_rn(a[3], a[1]), 1.f));
  const float sb = __fmul_rn(__fadd_rn(__fsub_rn(b[2], b[0]), 1.f), __fadd_rn(__fsub_rn(b[3], b[1]), 1.f));
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
}

// mask[i, w] bit j: sorted box (64*w + j) is suppressed by sorted box i (j > i).  grid (words, n), block 64
__global__ void nms_mask_kernel(const float* __restrict__ dets, const long long* __restrict__ order, int n, float thr, int strict,
                                unsigned long long* __restrict__ mask, int words) {
  const int i = blockIdx.y, wd = blockIdx.x;
  const int j = wd * 64 + threadIdx.x;
  __shared__ unsigned long long bits;
  if (threadIdx.x == 0) bits = 0ull;
  __syncthreads();
  if (j < n && j > i) {
    const float ov = iou_plus1(dets + 4 * order[i], dets + 4 * order[j]);
    if (strict ? (ov > thr) : (ov >= thr)) atomicOr(&bits, 1ull << threadIdx.x);
  }
  __syncthreads();
  if (threadIdx.x == 0) mask[(size_t)i * words + wd] = bits;
}

// sequential sweep over the score-sorted boxes (one warp; lane w owns removed-word w, w+32, ...)
__global__ void nms_sweep_kernel(const unsigned long long* __restrict__ mask, const long long* __restrict__ order, int n, int words,
                                 unsigned char* __restrict__ keep_flags) {
  extern __shared__ unsigned long long removed[];
  for (int w = threadIdx.x; w < words; w += 32) removed[w] = 0ull;
  __syncwarp();
  for (int i = 0; i < n; ++i) {
    const bool dead = (removed[i >> 6] >> (i & 63)) & 1ull;
    __syncwarp();
    if (!dead) {
      if (threadIdx.x == 0) keep_flags[order[i]] = 1;
      for (int w = threadIdx.x; w < words; w += 32) removed[w] |= mask[(size_t)i * words + w];
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------
// tracking cost: one CTA per (detection i, track j).  boxes are xywh (normalised); feats [*,F]; dists [*,D].
__global__ void track_cost_kernel(const float* __restrict__ det_box, const float* __restrict__ trk_box,
                                  const float* __restrict__ det_feat, const float* __restrict__ trk_feat, int F,
                                  const float* __restrict__ det_dist, const float* __restrict__ trk_dist, int D, int n_trk,
                                  float wc, float wf, float wb, float wg, float* __restrict__ cost,
                                  float* __restrict__ cost_dist, float* __restrict__ cost_feat) {
  const int i = blockIdx.y, j = blockIdx.x;
  __shared__ float red[5][32];
  float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // |x|^2, |y|^2 (feat); then dist sums; dot products computed in pass 2
  const float* x = det_feat + (size_t)i * F;
  const float* y = trk_feat + (size_t)j * F;
  const float* p = det_dist + (size_t)i * D;
  const float* q = trk_dist + (size_t)j * D;
  for (int k = threadIdx.x; k < F; k += blockDim.x) { s[0] += x[k] * x[k]; s[1] += y[k] * y[k]; }
  for (int k = threadIdx.x; k < D; k += blockDim.x) { s[2] += p[k] * p[k]; s[3] += q[k] * q[k]; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
  for (int t = 0; t < 4; ++t) { s[t] = warp_sum(s[t]); if (lane == 0) red[t][warp] = s[t]; }
  __syncthreads();
  float nrm[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) { float v = 0.f; for (int w = 0; w < nw; ++w) v += red[t][w]; nrm[t] = sqrtf(v) + 1e-12f; }
  __syncthreads();
  float df = 0.f, dd = 0.f;  // cos = sum (x/|x|)(y/|y|)   (matcher.py:70-78 divides first, then multiplies)
  for (int k = threadIdx.x; k < F; k += blockDim.x) df += (x[k] / nrm[0]) * (y[k] / nrm[1]);
  for (int k = threadIdx.x; k < D; k += blockDim.x) dd += (p[k] / nrm[2]) * (q[k] / nrm[3]);
  df = warp_sum(df); dd = warp_sum(dd);
  if (lane == 0) { red[0][warp] = df; red[1][warp] = dd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float cf = 0.f, cd = 0.f;
    for (int w = 0; w < nw; ++w) { cf += red[0][w]; cd += red[1][w]; }
    cf = 1.f - cf; cd = 1.f - cd;
    // xywh -> cxcywh (matcher.py:27-31), L1 distance (:140), xyxy (:9-13), GIoU (:49-68)
    const float* a = det_box + 4 * i; const float* b = trk_box + 4 * j;
    const float ac[4] = {a[0] + a[2] / 2, a[1] + a[3] / 2, a[2], a[3]};
    const float bc[4] = {b[0] + b[2] / 2, b[1] + b[3] / 2, b[2], b[3]};
    const float l1 = fabsf(ac[0] - bc[0]) + fabsf(ac[1] - bc[1]) + fabsf(ac[2] - bc[2]) + fabsf(ac[3] - bc[3]);
    const float ax[4] = {ac[0] - 0.5f * ac[2], ac[1] - 0.5f * ac[3], ac[0] + 0.5f * ac[2], ac[1] + 0.5f * ac[3]};
    const float bx[4] = {bc[0] - 0.5f * bc[2], bc[1] - 0.5f * bc[3], bc[0] + 0.5f * bc[2], bc[1] + 0.5f * bc[3]};
    const float area1 = (ax[2] - ax[0]) * (ax[3] - ax[1]), area2 = (bx[2] - bx[0]) * (bx[3] - bx[1]);
    const float iw = fmaxf(fminf(ax[2], bx[2]) - fmaxf(ax[0], bx[0]), 0.f), ih = fmaxf(fminf(ax[3], bx[3]) - fmaxf(ax[1], bx[1]), 0.f);
    const float inter = iw * ih;
    const float uni = area1 + area2 - inter;
    const float iou = inter / uni;
    const float ew = fmaxf(fmaxf(ax[2], bx[2]) - fminf(ax[0], bx[0]), 0.f), eh = fmaxf(fmaxf(ax[3], bx[3]) - fminf(ax[1], bx[1]), 0.f);
    const float earea = ew * eh;
    const float giou = iou - (earea - uni) / earea;
    const size_t o = (size_t)i * n_trk + j;
    cost[o] = wc * cd + wf * cf + wb * l1 + wg * (-giou);
    cost_dist[o] = cd;
    cost_feat[o] = cf;
  }
}

}  // namespace
}  // namespace nlv

using namespace nlv;
#define STREAM ((cudaStream_t)stream)

extern "C" {

int nlv_roi_align_fwd(const float* input, int b, int c, int h, int w, const float* rois, int r, float spatial_scale, int ph,
                      int pw, int sampling_ratio, float* out, void* stream) {
  NLV_CHECK_ARG(b >= 0 && c > 0 && h > 0 && w > 0 && r >= 0 && ph > 0 && pw > 0, "roi_align_fwd: bad sizes");
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(input && rois && out, "roi_align_fwd: null pointer");
  const long long total = (long long)r * c * ph * pw;
  int grid = cdiv(total, 256);
  const int cap = 32 * sm_count();
  if (grid > cap) grid = cap;
  roi_align_fwd_kernel<<<grid, 256, 0, STREAM>>>(input, c, h, w, rois, total, spatial_scale, ph, pw, sampling_ratio, out);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* dinput must be zero-initialised by the caller (gradients are accumulated with atomics). */
int nlv_roi_align_bwd(const float* grad, const float* rois, int r, float spatial_scale, int ph, int pw, int b, int c, int h, int w,
                      int sampling_ratio, float* dinput, void* stream) {
  NLV_CHECK_ARG(b >= 0 && c > 0 && h > 0 && w > 0 && r >= 0 && ph > 0 && pw > 0, "roi_align_bwd: bad sizes");
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(grad && rois && dinput, "roi_align_bwd: null pointer");
  const long long total = (long long)r * c * ph * pw;
  int grid = cdiv(total, 256);
  const int cap = 32 * sm_count();
  if (grid > cap) grid = cap;
  roi_align_bwd_kernel<<<grid, 256, 0, STREAM>>>(grad, c, h, w, rois, total, spatial_scale, ph, pw, sampling_ratio, dinput);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* order: indices sorted by descending score (int64[n]); mask_ws: u64[n * ceil(n/64)] workspace; keep_flags: u8[n], zeroed by the
 * caller, set to 1 for kept ORIGINAL indices.  strict=1: suppress when IoU > thr (CUDA reference); 0: >= (CPU reference). */
int nlv_nms(const float* dets, const long long* order, int n, float thr, int strict, unsigned long long* mask_ws,
            unsigned char* keep_flags, void* stream) {
  NLV_CHECK_ARG(n >= 0, "nms: bad size");
  if (n == 0) return NLV_OK;
  NLV_CHECK_ARG(dets && order && mask_ws && keep_flags, "nms: null pointer");
  const int words = cdiv(n, 64);
  NLV_CHECK_ARG(n <= 65535 && words * 8 <= 48 * 1024, "nms: n=%d too large", n);
  nms_mask_kernel<<<dim3(words, n), 64, 0, STREAM>>>(dets, order, n, thr, strict, mask_ws, words);
  NLV_CHECK_LAUNCH();
  nms_sweep_kernel<<<1, 32, words * sizeof(unsigned long long), STREAM>>>(mask_ws, order, n, words, keep_flags);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_track_cost(const float* det_box_xywh, const float* trk_box_xywh, const float* det_feat, const float* trk_feat, int feat_dim,
                   const float* det_dist, const float* trk_dist, int dist_dim, int n_det, int n_trk, float w_class, float w_feat,
                   float w_bbox, float w_giou, float* cost, float* cost_dist, float* cost_feat, void* stream) {
  NLV_CHECK_ARG(n_det >= 0 && n_trk >= 0 && feat_dim > 0 && dist_dim > 0, "track_cost: bad sizes");
  if (n_det == 0 || n_trk == 0) return NLV_OK;
  NLV_CHECK_ARG(det_box_xywh && trk_box_xywh && det_feat && trk_feat && det_dist && trk_dist && cost && cost_dist && cost_feat,
                "track_cost: null pointer");
  NLV_CHECK_ARG(n_det <= 65535, "track_cost: too many detections");
  track_cost_kernel<<<dim3(n_trk, n_det), 256, 0, STREAM>>>(det_box_xywh, trk_box_xywh, det_feat, trk_feat, feat_dim, det_dist,
                                                          trk_dist, dist_dim, n_trk, w_class, w_feat, w_bbox, w_giou, cost,
                                                          cost_dist, cost_feat);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
}
