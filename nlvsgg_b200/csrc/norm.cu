// LayerNorm (row-wise, one warp per row) and segmented BatchNorm (per-video batch statistics over a
// rows x channels matrix) — forward and backward.
//
// LayerNorm  : lib/transformer.py:15-16,45 (nn.LayerNorm(1936), eps 1e-5)
// BatchNorm  : lib/sttran.py:43,49,340,344 (BatchNorm1d(4, m=0.001), BatchNorm1d(1024), BatchNorm2d(128/256, m=0.01)).
//              The reference runs one video per step, so batch statistics are per video; `seg` holds the row
//              offsets of each video inside a multi-video batch (nseg+1 entries).
#include <stdlib.h>

#include "common.cuh"

namespace nlv {
namespace {

constexpr int LN_MAX_PER_LANE = 64;  // supports up to 2048 columns

// ------------------------------------------------------------------------------------------
// LayerNorm forward: y = (x - mean) * rstd * w + b ; saves mean/rstd; optional second (operand) output
// ------------------------------------------------------------------------------------------
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, long long rows, int cols, const float* __restrict__ w,
                                     const float* __restrict__ b, float eps, float* __restrict__ y,
                                     void* __restrict__ y2, int y2dt, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * cols;
  float v[LN_MAX_PER_LANE];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + 32 * i;
    v[i] = c < cols ? xr[c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + 32 * i;
    const float d = c < cols ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + 32 * i;
    if (c < cols) {
      const float o = (v[i] - mean) * rstd * w[c] + b[c];
      if (y) y[(size_t)row * cols + c] = o;
      if (y2) st_from_float(y2, y2dt, (size_t)row * cols + c, o);
    }
  }
}

// LayerNorm backward, input gradient.  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w
// Optional dx2: operand copy of dx in another dtype.
__global__ void layernorm_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ w, long long rows, int cols, float* __restrict__ dx,
                                        void* __restrict__ dx2, int dx2dt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const float m = mean[row], rs = rstd[row];
  const float* dyr = dy + (size_t)row * cols;
  const float* xr = x + (size_t)row * cols;
  float g[LN_MAX_PER_LANE], xh[LN_MAX_PER_LANE];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + 32 * i;
    if (c < cols) {
      xh[i] = (xr[c] - m) * rs;
      g[i] = dyr[c] * w[c];
      s1 += g[i];
      s2 += g[i] * xh[i];
    } else { g[i] = 0.f; xh[i] = 0.f; }
  }
  s1 = warp_sum(s1) / (float)cols;
  s2 = warp_sum(s2) / (float)cols;
#pragma unroll
  for (int i = 0; i < LN_MAX_PER_LANE; ++i) {
    const int c = lane + 32 * i;
    if (c < cols) {
      const float o = rs * (g[i] - s1 - xh[i] * s2);
      if (dx) dx[(size_t)row * cols + c] = o;
      if (dx2) st_from_float(dx2, dx2dt, (size_t)row * cols + c, o);
    }
  }
}

// LayerNorm backward, parameter gradients: dw[c] += sum_r dy*xhat, db[c] += sum_r dy.
// Each thread owns 4 consecutive columns (float4); grid (ceil(cols/128), row_splits), block 32 x 8.
__global__ void layernorm_bwd_param_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                           const float* __restrict__ mean, const float* __restrict__ rstd, long long rows,
                                           int cols, float* __restrict__ dw, float* __restrict__ db) {
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * per, r1 = min(rows, r0 + per);
  float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < cols)   // cols % 4 == 0 is checked by the launcher
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      const float4 d = *reinterpret_cast<const float4*>(dy + (size_t)r * cols + c);
      const float4 xv = *reinterpret_cast<const float4*>(x + (size_t)r * cols + c);
      const float m = mean[r], rs = rstd[r];
      a[0] += d.x * (xv.x - m) * rs; a[1] += d.y * (xv.y - m) * rs; a[2] += d.z * (xv.z - m) * rs; a[3] += d.w * (xv.w - m) * rs;
      b[0] += d.x; b[1] += d.y; b[2] += d.z; b[3] += d.w;
    }
  __shared__ float red[2][8][32 * 4 + 4];
#pragma unroll
  for (int q = 0; q < 4; ++q) { red[0][threadIdx.y][threadIdx.x * 4 + q] = a[q]; red[1][threadIdx.y][threadIdx.x * 4 + q] = b[q]; }
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { t1 += red[0][j][threadIdx.x * 4 + q]; t2 += red[1][j][threadIdx.x * 4 + q]; }
      atomicAdd(dw + c + q, t1);
      atomicAdd(db + c + q, t2);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Segmented BatchNorm
// ------------------------------------------------------------------------------------------
// pass 1: per (segment, channel) sum and sum of squares in double.  grid (ceil(C/32), nseg, splits), block 32x8
__global__ void bn_stats_kernel(const void* __restrict__ x, int xdt, int ld, const int* __restrict__ seg, int C,
                                double* __restrict__ sums /*[nseg,2,C]*/) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int s = blockIdx.y;
  const long long a = seg[s], e = seg[s + 1];
  const long long per = (e - a + gridDim.z - 1) / gridDim.z;
  const long long r0 = a + (long long)blockIdx.z * per, r1 = min(e, r0 + per);
  double s1 = 0.0, s2 = 0.0;
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      const double v = (double)ld_as_float(x, xdt, (size_t)r * ld + c);
      s1 += v; s2 += v * v;
    }
  __shared__ double red[2][8][33];
  red[0][threadIdx.y][threadIdx.x] = s1;
  red[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && c < C && r0 < r1) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { t1 += red[0][j][threadIdx.x]; t2 += red[1][j][threadIdx.x]; }
    atomicAdd(sums + ((size_t)s * 2 + 0) * C + c, t1);
    atomicAdd(sums + ((size_t)s * 2 + 1) * C + c, t2);
  }
}

// pass 2: mean / biased variance per (segment, channel); running statistics updated sequentially over the
// segments (= the order in which the reference would have seen the videos), unbiased variance, momentum m.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const int* __restrict__ seg, int nseg, int C,
                                   float momentum, float* __restrict__ mean, float* __restrict__ var,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
  for (int s = 0; s < nseg; ++s) {
    const double n = (double)(seg[s + 1] - seg[s]);
    if (n <= 0) { mean[(size_t)s * C + c] = 0.f; var[(size_t)s * C + c] = 0.f; continue; }
    const double m = sums[((size_t)s * 2 + 0) * C + c] / n;
    double v = sums[((size_t)s * 2 + 1) * C + c] / n - m * m;
    if (v < 0) v = 0;
    mean[(size_t)s * C + c] = (float)m;
    var[(size_t)s * C + c] = (float)v;
    const float unb = (float)(n > 1 ? v * n / (n - 1) : v);
    rm = (1.f - momentum) * rm + momentum * (float)m;
    rv = (1.f - momentum) * rv + momentum * unb;
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
}

// The same with the (segment, channel) statistics computed by 8 segment lanes per channel in parallel — the sequential kernel
// above pays one dependent global double load per segment (70 us for 64 videos, four times per step); only the running-statistics
// recursion (same order, same arithmetic) stays sequential, over shared memory.  nseg <= FIN_MAX_SEG.
constexpr int FIN_MAX_SEG = 128;
__global__ void __launch_bounds__(256)
bn_finalize_par_kernel(const double* __restrict__ sums, const int* __restrict__ seg, int nseg, int C, float momentum,
                       float* __restrict__ mean, float* __restrict__ var, float* __restrict__ running_mean, float* __restrict__ running_var) {
  __shared__ float sm_m[FIN_MAX_SEG][32], sm_u[FIN_MAX_SEG][32];
  __shared__ unsigned char sm_ok[FIN_MAX_SEG];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int s = threadIdx.y; s < nseg; s += 8) {
    const double n = (double)(seg[s + 1] - seg[s]);
    if (threadIdx.x == 0) sm_ok[s] = n > 0 ? 1 : 0;
    if (c >= C) continue;
    if (n <= 0) { mean[(size_t)s * C + c] = 0.f; var[(size_t)s * C + c] = 0.f; continue; }
    const double m = sums[((size_t)s * 2 + 0) * C + c] / n;
    double v = sums[((size_t)s * 2 + 1) * C + c] / n - m * m;
    if (v < 0) v = 0;
    mean[(size_t)s * C + c] = (float)m;
    var[(size_t)s * C + c] = (float)v;
    sm_m[s][threadIdx.x] = (float)m;
    sm_u[s][threadIdx.x] = (float)(n > 1 ? v * n / (n - 1) : v);
  }
  __syncthreads();
  if (threadIdx.y != 0 || c >= C) return;
  float rm = running_mean ? running_mean[c] : 0.f, rv = running_var ? running_var[c] : 0.f;
  for (int s = 0; s < nseg; ++s) {
    if (!sm_ok[s]) continue;
    rm = (1.f - momentum) * rm + momentum * sm_m[s][threadIdx.x];
    rv = (1.f - momentum) * rv + momentum * sm_u[s][threadIdx.x];
  }
  if (running_mean) running_mean[c] = rm;
  if (running_var) running_var[c] = rv;
}

// apply: y = (x - mean[s,c]) * rsqrt(var[s,c] + eps) * w[c] + b[c]  (optional ReLU); up to two outputs
__global__ void bn_apply_kernel(const void* __restrict__ x, int xdt, int ldx, const int* __restrict__ row_seg, int row_div,
                                const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ w,
                                const float* __restrict__ b, float eps, int relu, long long rows, int C,
                                void* __restrict__ y, int ydt, int ldy, void* __restrict__ y2, int y2dt, int ldy2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  const int s = row_seg ? row_seg[r / row_div] : 0;
  const float m = mean[(size_t)s * C + c], v = var[(size_t)s * C + c];
  float o = (ld_as_float(x, xdt, (size_t)r * ldx + c) - m) * rsqrtf(v + eps) * w[c] + b[c];
  if (relu) o = fmaxf(o, 0.f);
  if (y) st_from_float(y, ydt, (size_t)r * ldy + c, o);
  if (y2) st_from_float(y2, y2dt, (size_t)r * ldy2 + c, o);
}

// backward pass 1: per (segment, channel) sum(dy) and sum(dy * xhat), with the optional ReLU mask (y_out > 0)
__global__ void bn_bwd_stats_kernel(const void* __restrict__ dy, int dydt, int lddy, const void* __restrict__ x, int xdt, int ldx,
                                    const void* __restrict__ yout, int ydt, int ldy, const int* __restrict__ seg,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps, int C,
                                    double* __restrict__ sums /*[nseg,2,C]*/) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int s = blockIdx.y;
  const long long a = seg[s], e = seg[s + 1];
  const long long per = (e - a + gridDim.z - 1) / gridDim.z;
  const long long r0 = a + (long long)blockIdx.z * per, r1 = min(e, r0 + per);
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    const float m = mean[(size_t)s * C + c], rs = rsqrtf(var[(size_t)s * C + c] + eps);
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
      float d = ld_as_float(dy, dydt, (size_t)r * lddy + c);
      if (yout != nullptr && !(ld_as_float(yout, ydt, (size_t)r * ldy + c) > 0.f)) d = 0.f;
      const float xh = (ld_as_float(x, xdt, (size_t)r * ldx + c) - m) * rs;
      s1 += (double)d; s2 += (double)d * (double)xh;
    }
  }
  __shared__ double red[2][8][33];
  red[0][threadIdx.y][threadIdx.x] = s1;
  red[1][threadIdx.y][threadIdx.x] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && c < C && r0 < r1) {
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { t1 += red[0][j][threadIdx.x]; t2 += red[1][j][threadIdx.x]; }
    atomicAdd(sums + ((size_t)s * 2 + 0) * C + c, t1);
    atomicAdd(sums + ((size_t)s * 2 + 1) * C + c, t2);
  }
}

// backward pass 2 (training statistics): dx = w*rstd*(dy - sum_dy/n - xhat*sum_dy_xhat/n)
// eval statistics (use_batch_stats=0): dx = w*rstd*dy.  Output dtype selectable (operand for the next GEMM).
__global__ void bn_bwd_apply_kernel(const void* __restrict__ dy, int dydt, int lddy, const void* __restrict__ x, int xdt, int ldx,
                                    const void* __restrict__ yout, int ydt, int ldy, const int* __restrict__ row_seg, int row_div,
                                    const int* __restrict__ seg, const float* __restrict__ mean,
                                    const float* __restrict__ var, const float* __restrict__ w, float eps,
                                    const double* __restrict__ sums, int use_batch_stats, int gate_by_x, long long rows, int C,
                                    void* __restrict__ dx, int dxdt, int lddx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long long r = i / C;
  const int c = (int)(i - r * C);
  const int s = row_seg ? row_seg[r / row_div] : 0;
  const float m = mean[(size_t)s * C + c], rs = rsqrtf(var[(size_t)s * C + c] + eps);
  float d = ld_as_float(dy, dydt, (size_t)r * lddy + c);
  if (yout != nullptr && !(ld_as_float(yout, ydt, (size_t)r * ldy + c) > 0.f)) d = 0.f;
  float o;
  if (use_batch_stats) {
    const float n = (float)(seg[s + 1] - seg[s]);
    const float xh = (ld_as_float(x, xdt, (size_t)r * ldx + c) - m) * rs;
    const float s1 = (float)(sums[((size_t)s * 2 + 0) * C + c]) / n;
    const float s2 = (float)(sums[((size_t)s * 2 + 1) * C + c]) / n;
    o = w[c] * rs * (d - s1 - xh * s2);
  } else {
    o = w[c] * rs * d;
  }
  if (gate_by_x && !(ld_as_float(x, xdt, (size_t)r * ldx + c) > 0.f)) o = 0.f;   // ReLU that precedes the BN
  st_from_float(dx, dxdt, (size_t)r * lddx + c, o);
}

// dw[c] += sum_s sums[s,1,c] ; db[c] += sum_s sums[s,0,c]
__global__ void bn_bwd_param_kernel(const double* __restrict__ sums, int nseg, int C, float* __restrict__ dw,
                                    float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int s = 0; s < nseg; ++s) { b += sums[((size_t)s * 2 + 0) * C + c]; a += sums[((size_t)s * 2 + 1) * C + c]; }
  dw[c] += (float)a;
  db[c] += (float)b;
}

int bn_splits(const int* /*seg device*/, long long rows, int nseg) {
  long long per_seg = nseg > 0 ? rows / nseg : rows;
  int s = (int)((per_seg + 511) / 512);
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}

}  // namespace

// 16-byte vectorised variants (norm_vec.cu), used when C % 8 == 0 and everything is 16-byte aligned
int launch_bn_sums_fwd_v8(const void* x, int xdt, int ld, const int* seg, int nseg, long long rows, int C, double* sums, cudaStream_t s);
int launch_bn_sums_bwd_v8(const void* dy, int dydt, int lddy, const void* x, int xdt, int ldx, const void* yout, int ydt, int ldy, const int* seg,
                          int nseg, const float* mean, const float* var, float eps, long long rows, int C, double* sums, cudaStream_t s);
int launch_bn_apply_v8(const void* x, int xdt, int ldx, const int* row_seg, int row_div, const float* mean, const float* var, const float* w,
                       const float* b, float eps, int relu, long long rows, int C, void* y, int ydt, int ldy, void* y2, int y2dt,
                       int ldy2, cudaStream_t s);
int launch_bn_bwd_apply_v8(const void* dy, int dydt, int lddy, const void* x, int xdt, int ldx, const void* yout, int ydt, int ldy,
                           const int* row_seg, int row_div, const int* seg, const float* mean, const float* var, const float* w, float eps,
                           const double* sums, int use_batch_stats, int gate_by_x, long long rows, int C, void* dx, int dxdt, int lddx,
                           float* dx_colsum, cudaStream_t s);
int launch_layernorm_fwd_v4(const float* x, long long rows, int cols, const float* w, const float* b, float eps, float* y, void* y2,
                            int y2dt, float* mean, float* rstd, cudaStream_t s);
int launch_layernorm_bwd_dx_v4(const float* dy, const float* x, const float* mean, const float* rstd, const float* w, long long rows,
                               int cols, float* dx, void* dx2, int dx2dt, const nlv_dropout* drop, cudaStream_t s);
int launch_layernorm_bwd_fused(const float* dy, const float* x, const float* mean, const float* rstd, const float* w, long long rows,
                               int cols, float* dx, void* dx2, int dx2dt, float* dw, float* db, float* dprev, const nlv_dropout* dr,
                               cudaStream_t s);
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
}  // namespace nlv

namespace nlv {
static int run_bn_finalize(const double* sums, const int* seg, int nseg, int c, float momentum, float* mean, float* var, float* running_mean,
                           float* running_var, cudaStream_t s) {
  if (nseg <= FIN_MAX_SEG) bn_finalize_par_kernel<<<cdiv(c, 32), dim3(32, 8), 0, s>>>(sums, seg, nseg, c, momentum, mean, var, running_mean, running_var);
  else bn_finalize_kernel<<<cdiv(c, 128), 128, 0, s>>>(sums, seg, nseg, c, momentum, mean, var, running_mean, running_var);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_bn_bwd_param(const double* sums, int nseg, int c, float* dw, float* db, cudaStream_t s) {
  bn_bwd_param_kernel<<<cdiv(c, 128), 128, 0, s>>>(sums, nseg, c, dw, db);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
// the finalize pass alone, for producers that accumulate the (segment, channel) sums themselves (maskconv.cu)
int launch_bn_finalize(const double* sums, const int* seg, int nseg, int c, float momentum, float* mean, float* var, float* running_mean,
                       float* running_var, cudaStream_t s) {
  return run_bn_finalize(sums, seg, nseg, c, momentum, mean, var, running_mean, running_var, s);
}
}  // namespace nlv

using namespace nlv;
#define STREAM ((cudaStream_t)stream)

extern "C" {

int nlv_layernorm_fwd(const float* x, long long rows, int cols, const float* w, const float* b, float eps, float* y,
                      void* y2, int y2_dtype, float* mean, float* rstd, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols > 0 && cols <= 32 * LN_MAX_PER_LANE, "layernorm_fwd: cols=%d unsupported", cols);
  if (rows == 0) return NLV_OK;
  NLV_CHECK_ARG(x && w && b && (y || y2), "layernorm_fwd: null pointer");
  if ((cols & 3) == 0 && al16(x) && al16(w) && al16(b) && al16(y) && (y2 == nullptr || (reinterpret_cast<uintptr_t>(y2) & 7) == 0))
    return launch_layernorm_fwd_v4(x, rows, cols, w, b, eps, y, y2, y2_dtype, mean, rstd, STREAM);
  const int warps = 4;
  layernorm_fwd_kernel<<<cdiv(rows, warps), warps * 32, 0, STREAM>>>(x, rows, cols, w, b, eps, y, y2, y2_dtype, mean, rstd);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* dw, db are ACCUMULATED into (callers zero them). */
int nlv_layernorm_bwd_drop(const float* dy, const float* x, const float* mean, const float* rstd, const float* w,
                           long long rows, int cols, float* dx, void* dx2, int dx2_dtype, float* dw, float* db,
                           const nlv_dropout* drop, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols > 0 && cols <= 32 * LN_MAX_PER_LANE, "layernorm_bwd: cols=%d unsupported", cols);
  if (rows == 0) return NLV_OK;
  NLV_CHECK_ARG(dy && x && mean && rstd && w && dw && db && (dx || dx2), "layernorm_bwd: null pointer");
  const bool dropping = drop != nullptr && drop->thr16 != 0u;
  if ((cols & 3) == 0 && al16(dy) && al16(x) && al16(w) && al16(dx) && (dx2 == nullptr || (reinterpret_cast<uintptr_t>(dx2) & 15) == 0)) {
    const int rc = launch_layernorm_bwd_dx_v4(dy, x, mean, rstd, w, rows, cols, dx, dx2, dx2_dtype, drop, STREAM);
    if (rc != NLV_OK) return rc;
  } else {
    if (dropping) { nlv::set_error("layernorm_bwd: the dropout variant needs 16-byte aligned rows (cols %% 4 == 0)"); return NLV_ERR_UNSUPPORTED; }
    layernorm_bwd_dx_kernel<<<cdiv(rows, 4), 128, 0, STREAM>>>(dy, x, mean, rstd, w, rows, cols, dx, dx2, dx2_dtype);
    NLV_CHECK_LAUNCH();
  }
  NLV_CHECK_ARG((cols & 3) == 0, "layernorm_bwd: cols=%d must be a multiple of 4", cols);
  int splits = (int)((rows + 255) / 256);
  if (splits > 1024) splits = 1024;
  dim3 grid(cdiv(cols, 128), splits), block(32, 8);
  layernorm_bwd_param_kernel<<<grid, block, 0, STREAM>>>(dy, x, mean, rstd, rows, cols, dw, db);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* one pass: additionally accumulates colsum(dx2 as stored, before rounding) into dprev (nullable) — the bias gradient of the
 * Linear layer in front of the residual sum.  cols % 8 == 0, cols <= 2048, 16-byte aligned rows. */
int nlv_layernorm_bwd_fused(const float* dy, const float* x, const float* mean, const float* rstd, const float* w,
                            long long rows, int cols, float* dx, void* dx2, int dx2_dtype, float* dw, float* db, float* dprev,
                            const nlv_dropout* drop, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols > 0 && cols <= 2048 && (cols & 7) == 0, "layernorm_bwd_fused: cols=%d unsupported", cols);
  if (rows == 0) return NLV_OK;
  NLV_CHECK_ARG(dy && x && mean && rstd && w && dw && db && (dx || dx2), "layernorm_bwd_fused: null pointer");
  NLV_CHECK_ARG(al16(dy) && al16(x) && (dx == nullptr || al16(dx)) && (dx2 == nullptr || al16(dx2)), "layernorm_bwd_fused: 16-byte alignment");
  static const bool two_pass = [] { const char* e = getenv("NLV_LN_BWD_FUSED"); return e != nullptr && e[0] == '0'; }();
  if (two_pass) {   // debugging switch: the two-kernel form + a separate column sum
    int rc = nlv_layernorm_bwd_drop(dy, x, mean, rstd, w, rows, cols, dx, dx2, dx2_dtype, dw, db, drop, stream);
    if (rc != NLV_OK || dprev == nullptr) return rc;
    const bool masked = drop != nullptr && drop->thr16 != 0u && dx2 != nullptr;
    return nlv_colsum(masked ? dx2 : (const void*)dx, masked ? dx2_dtype : NLV_F32, cols, rows, cols, nullptr, 1, dprev, stream);
  }
  return launch_layernorm_bwd_fused(dy, x, mean, rstd, w, rows, cols, dx, dx2, dx2_dtype, dw, db, dprev, drop, STREAM);
}

int nlv_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* w,
                      long long rows, int cols, float* dx, void* dx2, int dx2_dtype, float* dw, float* db, void* stream) {
  return nlv_layernorm_bwd_drop(dy, x, mean, rstd, w, rows, cols, dx, dx2, dx2_dtype, dw, db, nullptr, stream);
}

/* Training-mode statistics.  seg: int[nseg+1] row offsets (device).  sums_ws: double[nseg*2*C] workspace (zeroed here).
 * mean/var: float[nseg*C] outputs.  running_mean/var updated in place (may be null). */
int nlv_bn_stats(const void* x, int x_dtype, int ld, const int* seg, int nseg, long long rows, int c, float momentum,
                 double* sums_ws, float* mean, float* var, float* running_mean, float* running_var, void* stream) {
  NLV_CHECK_ARG(nseg >= 1 && c > 0 && rows >= 0, "bn_stats: bad sizes");
  NLV_CHECK_ARG(x && seg && sums_ws && mean && var, "bn_stats: null pointer");
  { int zrc = zero_fill(reinterpret_cast<float*>(sums_ws), 1, 4 * nseg * c, 4 * nseg * c, STREAM); if (zrc != NLV_OK) return zrc; }
  if (rows > 0) {
    NLV_CHECK_ARG(nseg <= 65535, "bn_stats: too many segments");
    if ((c & 7) == 0 && (ld & 7) == 0 && al16(x)) {
      int rc = launch_bn_sums_fwd_v8(x, x_dtype, ld, seg, nseg, rows, c, sums_ws, STREAM);
      if (rc != NLV_OK) return rc;
    } else {
      dim3 grid(cdiv(c, 32), nseg, bn_splits(seg, rows, nseg)), block(32, 8);
      bn_stats_kernel<<<grid, block, 0, STREAM>>>(x, x_dtype, ld, seg, c, sums_ws);
      NLV_CHECK_LAUNCH();
    }
  }
  return run_bn_finalize(sums_ws, seg, nseg, c, momentum, mean, var, running_mean, running_var, STREAM);
}

/* mean/var are [nseg,C] (training: from nlv_bn_stats with row_seg; eval: running stats with row_seg = null). */
int nlv_bn_apply(const void* x, int x_dtype, int ldx, const int* row_seg, int row_div, const float* mean, const float* var,
                 const float* w, const float* b, float eps, int relu, long long rows, int c, void* y, int y_dtype, int ldy,
                 void* y2, int y2_dtype, int ldy2, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && c > 0, "bn_apply: bad sizes");
  if (rows == 0) return NLV_OK;
  NLV_CHECK_ARG(x && mean && var && w && b && (y || y2), "bn_apply: null pointer");
  if (row_div < 1) row_div = 1;
  if ((c & 7) == 0 && (ldx & 7) == 0 && (ldy & 7) == 0 && (ldy2 & 7) == 0 && al16(x) && al16(y) && al16(y2) && al16(mean) && al16(var) &&
      al16(w) && al16(b))
    return launch_bn_apply_v8(x, x_dtype, ldx, row_seg, row_div, mean, var, w, b, eps, relu, rows, c, y, y_dtype, ldy, y2, y2_dtype, ldy2, STREAM);
  bn_apply_kernel<<<cdiv(rows * c, 256), 256, 0, STREAM>>>(x, x_dtype, ldx, row_seg, row_div, mean, var, w, b, eps, relu, rows, c, y,
                                                          y_dtype, ldy, y2, y2_dtype, ldy2);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* Backward.  dy: fp32 or bf16.  yout (optional) = the ReLU'd forward output, masks dy (Linear -> BN -> ReLU).  gate_by_x: zero dx where x <= 0
 * (conv -> ReLU -> BN: x is the ReLU output, so this is the ReLU backward fused in).  dw/db accumulated.  sums_ws as in bn_stats. */
int nlv_bn_bwd(const void* dy, int dy_dtype, int lddy, const void* x, int x_dtype, int ldx, const void* yout, int y_dtype, int ldy,
               const int* seg, const int* row_seg, int row_div, int nseg, const float* mean, const float* var, const float* w, float eps,
               int use_batch_stats, int gate_by_x, long long rows, int c, double* sums_ws, void* dx, int dx_dtype, int lddx, float* dw,
               float* db, void* stream) {
  return nlv_bn_bwd_colsum(dy, dy_dtype, lddy, x, x_dtype, ldx, yout, y_dtype, ldy, seg, row_seg, row_div, nseg, mean, var, w, eps, use_batch_stats,
                           gate_by_x, rows, c, sums_ws, dx, dx_dtype, lddx, dw, db, nullptr, stream);
}

/* dx_colsum (nullable): += column sums of dx, accumulated by the kernel that writes dx (the bias gradient of the conv / linear
 * layer in front of the BatchNorm, otherwise a separate pass over dx) */
int nlv_bn_bwd_colsum(const void* dy, int dy_dtype, int lddy, const void* x, int x_dtype, int ldx, const void* yout, int y_dtype, int ldy,
                      const int* seg, const int* row_seg, int row_div, int nseg, const float* mean, const float* var, const float* w, float eps,
                      int use_batch_stats, int gate_by_x, long long rows, int c, double* sums_ws, void* dx, int dx_dtype, int lddx, float* dw,
                      float* db, float* dx_colsum, void* stream) {
  NLV_CHECK_ARG(nseg >= 1 && c > 0 && rows >= 0, "bn_bwd: bad sizes");
  NLV_CHECK_ARG(dy && x && seg && mean && var && w && sums_ws && dx && dw && db, "bn_bwd: null pointer");
  if (row_div < 1) row_div = 1;
  { int zrc = zero_fill(reinterpret_cast<float*>(sums_ws), 1, 4 * nseg * c, 4 * nseg * c, STREAM); if (zrc != NLV_OK) return zrc; }
  if (rows == 0) return NLV_OK;
  NLV_CHECK_ARG(nseg <= 65535, "bn_bwd: too many segments");
  const bool vec = (c & 7) == 0 && (lddy & 7) == 0 && (ldx & 7) == 0 && (ldy & 7) == 0 && (lddx & 7) == 0 && al16(dy) && al16(x) &&
                   al16(yout) && al16(dx) && al16(mean) && al16(var) && al16(w);
  if (vec) {
    int rc = launch_bn_sums_bwd_v8(dy, dy_dtype, lddy, x, x_dtype, ldx, yout, y_dtype, ldy, seg, nseg, mean, var, eps, rows, c, sums_ws, STREAM);
    if (rc != NLV_OK) return rc;
    rc = launch_bn_bwd_apply_v8(dy, dy_dtype, lddy, x, x_dtype, ldx, yout, y_dtype, ldy, row_seg, row_div, seg, mean, var, w, eps, sums_ws, use_batch_stats,
                                gate_by_x, rows, c, dx, dx_dtype, lddx, dx_colsum, STREAM);
    if (rc != NLV_OK) return rc;
  } else {
    dim3 grid(cdiv(c, 32), nseg, bn_splits(seg, rows, nseg)), block(32, 8);
    bn_bwd_stats_kernel<<<grid, block, 0, STREAM>>>(dy, dy_dtype, lddy, x, x_dtype, ldx, yout, y_dtype, ldy, seg, mean, var, eps, c, sums_ws);
    NLV_CHECK_LAUNCH();
    bn_bwd_apply_kernel<<<cdiv(rows * c, 256), 256, 0, STREAM>>>(dy, dy_dtype, lddy, x, x_dtype, ldx, yout, y_dtype, ldy, row_seg, row_div, seg, mean,
                                                                var, w, eps, sums_ws, use_batch_stats, gate_by_x, rows, c, dx, dx_dtype,
                                                                lddx);
    NLV_CHECK_LAUNCH();
  }
  bn_bwd_param_kernel<<<cdiv(c, 128), 128, 0, STREAM>>>(sums_ws, nseg, c, dw, db);
  NLV_CHECK_LAUNCH();
  if (!vec && dx_colsum != nullptr) return nlv_colsum(dx, dx_dtype, lddx, rows, c, nullptr, 1, dx_colsum, stream);
  return NLV_OK;
}
}
