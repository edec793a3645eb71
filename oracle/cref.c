/* TEST INFRASTRUCTURE — plain-C restatement of the reference's native (Cython / C++) helpers.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 * Each function cites the reference source it restates (paths relative to /root/reference).
 * Pinned against the reference's own Cython build by oracle/validate_oracle.py and against
 * tests/golden/native_*.npz.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, so float results are
 * the ones the reference's x86-64 build produces).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float clamp01f(float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); }

/* lib/draw_rectangles/draw_rectangles.pyx:26-66  draw_union_boxes_c
 * box_pairs f32[n,8] (x1,y1,x2,y2 of box A, then of box B) -> out f32[n,2,ps,ps]. */
void oracle_draw_union_boxes(const float* box_pairs, int n, int ps, float* out) {
  for (int r = 0; r < n; ++r) {
    const float* b = box_pairs + 8 * r;
    float x1u = b[0] < b[4] ? b[0] : b[4];
    float y1u = b[1] < b[5] ? b[1] : b[5];
    float x2u = b[2] > b[6] ? b[2] : b[6];
    float y2u = b[3] > b[7] ? b[3] : b[7];
    float w = x2u - x1u, h = y2u - y1u;
    for (int i = 0; i < 2; ++i) {
      float x1 = (b[0 + 4 * i] - x1u) * (float)ps / w;
      float y1 = (b[1 + 4 * i] - y1u) * (float)ps / h;
      float x2 = (b[2 + 4 * i] - x1u) * (float)ps / w;
      float y2 = (b[3 + 4 * i] - y1u) * (float)ps / h;
      float* o = out + ((size_t)r * 2 + i) * ps * ps;
      for (int j = 0; j < ps; ++j) {
        float yc = clamp01f((float)(j + 1) - y1) * clamp01f(y2 - (float)j);
        for (int k = 0; k < ps; ++k) {
          float xc = clamp01f((float)(k + 1) - x1) * clamp01f(x2 - (float)k);
          o[j * ps + k] = xc * yc;
        }
      }
    }
  }
}

/* lib/fpn/box_intersections_cpu/bbox.pyx:21-61  bbox_overlaps_c  (float64, +1 pixel convention)
 * boxes f64[n,4], query f64[k,4] -> out f64[n,k]. */
void oracle_bbox_overlaps(const double* boxes, int n, const double* query, int k, double* out) {
  memset(out, 0, sizeof(double) * (size_t)n * k);
  for (int q = 0; q < k; ++q) {
    const double* qb = query + 4 * q;
    double qa = (qb[2] - qb[0] + 1) * (qb[3] - qb[1] + 1);
    for (int i = 0; i < n; ++i) {
      const double* bb = boxes + 4 * i;
      double iw = (bb[2] < qb[2] ? bb[2] : qb[2]) - (bb[0] > qb[0] ? bb[0] : qb[0]) + 1;
      if (iw > 0) {
        double ih = (bb[3] < qb[3] ? bb[3] : qb[3]) - (bb[1] > qb[1] ? bb[1] : qb[1]) + 1;
        if (ih > 0) {
          double ua = (bb[2] - bb[0] + 1) * (bb[3] - bb[1] + 1) + qa - iw * ih;
          out[(size_t)i * k + q] = iw * ih / ua;
        }
      }
    }
  }
}

/* fasterRCNN/lib/model/csrc/cpu/nms_cpu.cpp:5-65 (strict=0: suppress when IoU >= thr) and
 * fasterRCNN/lib/model/csrc/cuda/nms.cu:23-131   (strict=1: suppress when IoU >  thr).
 * `order` = indices sorted by descending score (the caller sorts, as scores.sort(0, true) does);
 * keep[] receives the kept ORIGINAL indices in ascending order; returns their count. */
int oracle_nms(const float* dets, const int64_t* order, int n, float thr, int strict, int64_t* keep) {
  uint8_t* sup = (uint8_t*)calloc((size_t)n + 1, 1);
  for (int _i = 0; _i < n; ++_i) {
    int64_t i = order[_i];
    if (sup[i]) continue;
    const float* a = dets + 4 * i;
    float iarea = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
    for (int _j = _i + 1; _j < n; ++_j) {
      int64_t j = order[_j];
      if (sup[j]) continue;
      const float* b = dets + 4 * j;
      float xx1 = a[0] > b[0] ? a[0] : b[0], yy1 = a[1] > b[1] ? a[1] : b[1];
      float xx2 = a[2] < b[2] ? a[2] : b[2], yy2 = a[3] < b[3] ? a[3] : b[3];
      float w = xx2 - xx1 + 1, h = yy2 - yy1 + 1;
      if (w < 0) w = 0;
      if (h < 0) h = 0;
      float inter = w * h;
      float barea = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
      float ovr = inter / (iarea + barea - inter);
      if (strict ? (ovr > thr) : (ovr >= thr)) sup[j] = 1;
    }
  }
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (!sup[i]) keep[m++] = i;
  free(sup);
  return m;
}

/* fasterRCNN/lib/model/csrc/cpu/ROIAlign_cpu.cpp:17-219 (forward, aligned=False, min roi size 1,
 * adaptive sampling grid when sampling_ratio<=0).  input f32[B,C,H,W], rois f32[R,5] -> out f32[R,C,ph,pw]. */
void oracle_roi_align_fwd(const float* in, int B, int C, int H, int W, const float* rois, int R,
                          float scale, int PH, int PW, int sampling_ratio, float* out) {
  (void)B;
  for (int n = 0; n < R; ++n) {
    const float* roi = rois + 5 * n;
    int bi = (int)roi[0];
    float sw = roi[1] * scale, sh = roi[2] * scale, ew = roi[3] * scale, eh = roi[4] * scale;
    float rw = ew - sw, rh = eh - sh;
    if (rw < 1.f) rw = 1.f;
    if (rh < 1.f) rh = 1.f;
    float bh = rh / (float)PH, bw = rw / (float)PW;
    int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / PH);
    int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / PW);
    float count = (float)(gh * gw);
    for (int c = 0; c < C; ++c) {
      const float* src = in + ((size_t)bi * C + c) * H * W;
      for (int ph = 0; ph < PH; ++ph)
        for (int pw = 0; pw < PW; ++pw) {
          float acc = 0.f;
          for (int iy = 0; iy < gh; ++iy) {
            float yy = sh + ph * bh + (float)(iy + .5f) * bh / (float)gh;
            for (int ix = 0; ix < gw; ++ix) {
              float xx = sw + pw * bw + (float)(ix + .5f) * bw / (float)gw;
              float x = xx, y = yy;
              if (y < -1.0 || y > H || x < -1.0 || x > W) continue; /* zero weights */
              if (y <= 0) y = 0;
              if (x <= 0) x = 0;
              int yl = (int)y, xl = (int)x, yh, xh;
              if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
              if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
              float ly = y - yl, lx = x - xl;
              float hy = 1.f - ly, hx = 1.f - lx;
              float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
              acc += w1 * src[yl * W + xl] + w2 * src[yl * W + xh] + w3 * src[yh * W + xl] + w4 * src[yh * W + xh];
            }
          }
          out[(((size_t)n * C + c) * PH + ph) * PW + pw] = acc / count;
        }
    }
  }
}
