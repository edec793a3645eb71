// 16-byte vectorised variants of the segmented BatchNorm kernels (channels-last rows x C, C % 8 == 0).
#include "common.cuh"

namespace nlv {

struct V8 { float v[8]; };
__device__ __forceinline__ V8 nv_ld8(const void* p, int dt, size_t i) {
  V8 r;
  if (dt == NLV_BF16) {
    const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float2 f = __bfloat1622float2(h[q]); r.v[2 * q] = f.x; r.v[2 * q + 1] = f.y; }
  } else {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i + 4);
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  }
  return r;
}
__device__ __forceinline__ void nv_st8(void* p, int dt, size_t i, const V8& r) {
  if (dt == NLV_BF16) {
    uint4 t;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(r.v[2 * q], r.v[2 * q + 1]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p) + i) = t;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
  }
}

namespace {

// per (segment, channel) sums: block (bx = min(C8,32) channel groups, by rows); grid (ceil(C8/bx), nseg, splits).
// MODE 0: sum x, sum x^2 (forward statistics).  MODE 1: sum dy, sum dy*xhat (backward), optional ReLU mask by yout.
template <int MODE>
__global__ void bn_sums_v8_kernel(const void* __restrict__ x, int xdt, int ldx, const float* __restrict__ dy, int lddy,
                                  const void* __restrict__ yout, int ydt, int ldy, const int* __restrict__ seg,
                                  const float* __restrict__ mean, const float* __restrict__ var, float eps, int C,
                                  double* __restrict__ sums) {
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  const int C8 = C >> 3;
  const int s = blockIdx.y;
  const long long a = seg[s], e = seg[s + 1];
  const long long per = (e - a + gridDim.z - 1) / gridDim.z;
  const long long r0 = a + (long long)blockIdx.z * per, r1 = min(e, r0 + per);
  const bool active = c8 < C8 && r0 < r1;
  float m[8], rs[8];
  if (MODE == 1 && active) {
#pragma unroll
    for (int q = 0; q < 8; ++q) { m[q] = mean[(size_t)s * C + c8 * 8 + q]; rs[q] = rsqrtf(var[(size_t)s * C + c8 * 8 + q] + eps); }
  }
  // fp32 partials over short runs of rows, folded into double every 64 rows
  double d1[8], d2[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { d1[q] = 0.0; d2[q] = 0.0; }
  float f1[8], f2[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { f1[q] = 0.f; f2[q] = 0.f; }
  int run = 0;
  for (long long r = r0 + threadIdx.y; active && r < r1; r += blockDim.y) {
    const V8 xv = nv_ld8(x, xdt, (size_t)r * ldx + (size_t)c8 * 8);
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { f1[q] += xv.v[q]; f2[q] = fmaf(xv.v[q], xv.v[q], f2[q]); }
    } else {
      V8 g = nv_ld8(dy, NLV_F32, (size_t)r * lddy + (size_t)c8 * 8);
      if (yout != nullptr) {
        const V8 yo = nv_ld8(yout, ydt, (size_t)r * ldy + (size_t)c8 * 8);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (!(yo.v[q] > 0.f)) g.v[q] = 0.f;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) { f1[q] += g.v[q]; f2[q] = fmaf(g.v[q], (xv.v[q] - m[q]) * rs[q], f2[q]); }
    }
    if (++run == 16) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { d1[q] += (double)f1[q]; d2[q] += (double)f2[q]; f1[q] = 0.f; f2[q] = 0.f; }
      run = 0;
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) { d1[q] += (double)f1[q]; d2[q] += (double)f2[q]; }
  // reduce over threadIdx.y in shared memory, then one double atomic per (segment, channel) per block
  __shared__ double red[256 * 8];
  for (int pass = 0; pass < 2; ++pass) {
    double* mine = red + (threadIdx.y * blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) mine[q] = pass == 0 ? d1[q] : d2[q];
    __syncthreads();
    if (threadIdx.y == 0 && active) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        double t = 0.0;
        for (int yy = 0; yy < blockDim.y; ++yy) t += red[(yy * blockDim.x + threadIdx.x) * 8 + q];
        atomicAdd(sums + ((size_t)s * 2 + pass) * C + c8 * 8 + q, t);
      }
    }
    __syncthreads();
  }
}

__global__ void bn_apply_v8_kernel(const void* __restrict__ x, int xdt, int ldx, const int* __restrict__ row_seg,
                                   const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ w,
                                   const float* __restrict__ b, float eps, int relu, long long rows, int C,
                                   void* __restrict__ y, int ydt, int ldy, void* __restrict__ y2, int y2dt, int ldy2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int C8 = C >> 3;
  if (i >= rows * C8) return;
  const long long r = i / C8;
  const int c0 = (int)(i - r * C8) * 8;
  const int s = row_seg ? row_seg[r] : 0;
  const V8 xv = nv_ld8(x, xdt, (size_t)r * ldx + c0);
  const V8 m = nv_ld8(mean, NLV_F32, (size_t)s * C + c0), v = nv_ld8(var, NLV_F32, (size_t)s * C + c0);
  const V8 ww = nv_ld8(w, NLV_F32, c0), bb = nv_ld8(b, NLV_F32, c0);
  V8 o;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    o.v[q] = (xv.v[q] - m.v[q]) * rsqrtf(v.v[q] + eps) * ww.v[q] + bb.v[q];
    if (relu) o.v[q] = fmaxf(o.v[q], 0.f);
  }
  if (y) nv_st8(y, ydt, (size_t)r * ldy + c0, o);
  if (y2) nv_st8(y2, y2dt, (size_t)r * ldy2 + c0, o);
}

// dx = w*rstd*(dy - sum_dy/n - xhat*sum_dy_xhat/n) [batch stats] or w*rstd*dy [running stats]; optional ReLU mask by
// yout on dy, optional ReLU gate by `x > 0` on the RESULT (ReLU that precedes the BN: conv -> ReLU -> BN).
__global__ void bn_bwd_apply_v8_kernel(const float* __restrict__ dy, int lddy, const void* __restrict__ x, int xdt, int ldx,
                                       const void* __restrict__ yout, int ydt, int ldy, const int* __restrict__ row_seg,
                                       const int* __restrict__ seg, const float* __restrict__ mean, const float* __restrict__ var,
                                       const float* __restrict__ w, float eps, const double* __restrict__ sums, int use_batch_stats,
                                       int gate_by_x, long long rows, int C, void* __restrict__ dx, int dxdt, int lddx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int C8 = C >> 3;
  if (i >= rows * C8) return;
  const long long r = i / C8;
  const int c0 = (int)(i - r * C8) * 8;
  const int s = row_seg ? row_seg[r] : 0;
  const V8 xv = nv_ld8(x, xdt, (size_t)r * ldx + c0);
  V8 g = nv_ld8(dy, NLV_F32, (size_t)r * lddy + c0);
  if (yout != nullptr) {
    const V8 yo = nv_ld8(yout, ydt, (size_t)r * ldy + c0);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (!(yo.v[q] > 0.f)) g.v[q] = 0.f;
  }
  const V8 m = nv_ld8(mean, NLV_F32, (size_t)s * C + c0), v = nv_ld8(var, NLV_F32, (size_t)s * C + c0), ww = nv_ld8(w, NLV_F32, c0);
  const float n = use_batch_stats ? (float)(seg[s + 1] - seg[s]) : 1.f;
  V8 o;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float rs = rsqrtf(v.v[q] + eps);
    if (use_batch_stats) {
      const float xh = (xv.v[q] - m.v[q]) * rs;
      const float s1 = (float)(sums[((size_t)s * 2 + 0) * C + c0 + q]) / n;
      const float s2 = (float)(sums[((size_t)s * 2 + 1) * C + c0 + q]) / n;
      o.v[q] = ww.v[q] * rs * (g.v[q] - s1 - xh * s2);
    } else {
      o.v[q] = ww.v[q] * rs * g.v[q];
    }
    if (gate_by_x && !(xv.v[q] > 0.f)) o.v[q] = 0.f;
  }
  nv_st8(dx, dxdt, (size_t)r * lddx + c0, o);
}

void sums_geometry(int C, long long rows, int nseg, dim3& grid, dim3& block) {
  const int C8 = C >> 3;
  int bx = 1;
  while (bx < C8 && bx < 32) bx <<= 1;
  const int by = 256 / bx;
  long long per_seg = nseg > 0 ? rows / nseg : rows;
  int splits = (int)((per_seg + 32LL * by - 1) / (32LL * by));
  if (splits < 1) splits = 1;
  if (splits > 256) splits = 256;
  grid = dim3(cdiv(C8, bx), nseg, splits);
  block = dim3(bx, by);
}

}  // namespace

int launch_bn_sums_fwd_v8(const void* x, int xdt, int ld, const int* seg, int nseg, long long rows, int C, double* sums, cudaStream_t s) {
  dim3 grid, block;
  sums_geometry(C, rows, nseg, grid, block);
  bn_sums_v8_kernel<0><<<grid, block, 0, s>>>(x, xdt, ld, nullptr, 0, nullptr, 0, 0, seg, nullptr, nullptr, 0.f, C, sums);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_bn_sums_bwd_v8(const float* dy, int lddy, const void* x, int xdt, int ldx, const void* yout, int ydt, int ldy, const int* seg,
                          int nseg, const float* mean, const float* var, float eps, long long rows, int C, double* sums, cudaStream_t s) {
  dim3 grid, block;
  sums_geometry(C, rows, nseg, grid, block);
  bn_sums_v8_kernel<1><<<grid, block, 0, s>>>(x, xdt, ldx, dy, lddy, yout, ydt, ldy, seg, mean, var, eps, C, sums);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_bn_apply_v8(const void* x, int xdt, int ldx, const int* row_seg, const float* mean, const float* var, const float* w,
                       const float* b, float eps, int relu, long long rows, int C, void* y, int ydt, int ldy, void* y2, int y2dt,
                       int ldy2, cudaStream_t s) {
  bn_apply_v8_kernel<<<cdiv(rows * (C / 8), 256), 256, 0, s>>>(x, xdt, ldx, row_seg, mean, var, w, b, eps, relu, rows, C, y, ydt, ldy,
                                                              y2, y2dt, ldy2);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_bn_bwd_apply_v8(const float* dy, int lddy, const void* x, int xdt, int ldx, const void* yout, int ydt, int ldy,
                           const int* row_seg, const int* seg, const float* mean, const float* var, const float* w, float eps,
                           const double* sums, int use_batch_stats, int gate_by_x, long long rows, int C, void* dx, int dxdt, int lddx,
                           cudaStream_t s) {
  bn_bwd_apply_v8_kernel<<<cdiv(rows * (C / 8), 256), 256, 0, s>>>(dy, lddy, x, xdt, ldx, yout, ydt, ldy, row_seg, seg, mean, var, w, eps,
                                                                  sums, use_batch_stats, gate_by_x, rows, C, dx, dxdt, lddx);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

}  // namespace nlv
