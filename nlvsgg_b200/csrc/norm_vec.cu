// 16-byte vectorised variants of the segmented BatchNorm kernels (channels-last rows x C, C % 8 == 0).
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"

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

// raw (still packed) 8-element load: 16 bytes for bf16, 32 for fp32; unpacked at the point of use so that several rows
// can be in flight per thread without holding their fp32 expansions in registers
struct R8 { uint4 a, b; };
__device__ __forceinline__ R8 nv_ld8_raw(const void* p, int dt, size_t i) {
  R8 r;
  if (dt == NLV_BF16) {
    r.a = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
    r.b = make_uint4(0u, 0u, 0u, 0u);
  } else {
    r.a = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p) + i);
    r.b = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p) + i + 4);
  }
  return r;
}
__device__ __forceinline__ V8 nv_unpack(const R8& r, int dt) {
  V8 o;
  if (dt == NLV_BF16) {
    const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) { o.v[2 * q] = __uint_as_float(w[q] << 16); o.v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u); }
  } else {
    o.v[0] = __uint_as_float(r.a.x); o.v[1] = __uint_as_float(r.a.y); o.v[2] = __uint_as_float(r.a.z); o.v[3] = __uint_as_float(r.a.w);
    o.v[4] = __uint_as_float(r.b.x); o.v[5] = __uint_as_float(r.b.y); o.v[6] = __uint_as_float(r.b.z); o.v[7] = __uint_as_float(r.b.w);
  }
  return o;
}

namespace {

// per (segment, channel) sums: block (bx = min(C8,32) channel groups, by rows); grid (ceil(C8/bx), nseg, splits).
// MODE 0: sum x, sum x^2 (forward statistics).  MODE 1: sum dy, sum dy*xhat (backward), optional ReLU mask by yout.
// BF: every input is bf16 (compile time) -> a row in flight is one uint4 per tensor; U rows are requested before the first
// is unpacked.  A block walks a LONG row range (geometry below: ~4 blocks per SM in total), so the shared-memory reduction
// and the double atomics at its end are amortised over thousands of rows; per-thread partials are fp32 over 16 x U rows,
// folded into double.
template <int MODE, bool BF, int U>
__global__ void __launch_bounds__(256, 2)
bn_sums_v8_kernel(const void* __restrict__ x, int xdt, int ldx, const void* __restrict__ dy, int dydt, int lddy,
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
  double d1[8], d2[8];
  float f1[8], f2[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { d1[q] = 0.0; d2[q] = 0.0; f1[q] = 0.f; f2[q] = 0.f; }
  int run = 0;
  const bool relu_mask = MODE == 1 && yout != nullptr;
  const int xd = BF ? NLV_BF16 : xdt, gd = BF ? NLV_BF16 : dydt, yd = BF ? NLV_BF16 : ydt;
  for (long long r = r0 + threadIdx.y; active && r < r1; r += (long long)blockDim.y * U) {
    R8 xr[U], gr[U], yr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * blockDim.y;
      if (rr < r1) {
        if (BF) {
          xr[u].a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + (size_t)rr * ldx + (size_t)c8 * 8));
          if (MODE == 1) gr[u].a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)rr * lddy + (size_t)c8 * 8));
          if (relu_mask) yr[u].a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(yout) + (size_t)rr * ldy + (size_t)c8 * 8));
        } else {
          xr[u] = nv_ld8_raw(x, xdt, (size_t)rr * ldx + (size_t)c8 * 8);
          if (MODE == 1) gr[u] = nv_ld8_raw(dy, dydt, (size_t)rr * lddy + (size_t)c8 * 8);
          if (relu_mask) yr[u] = nv_ld8_raw(yout, ydt, (size_t)rr * ldy + (size_t)c8 * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * blockDim.y;
      if (rr < r1) {
        const V8 xv = nv_unpack(xr[u], xd);
        if (MODE == 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) { f1[q] += xv.v[q]; f2[q] = fmaf(xv.v[q], xv.v[q], f2[q]); }
        } else {
          V8 g = nv_unpack(gr[u], gd);
          if (relu_mask) {
            const V8 yo = nv_unpack(yr[u], yd);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (!(yo.v[q] > 0.f)) g.v[q] = 0.f;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) { f1[q] += g.v[q]; f2[q] = fmaf(g.v[q], xv.v[q], f2[q]); }   // sum dy * x: centred below
        }
      }
    }
    if (++run == 16) {
#pragma unroll
      for (int q = 0; q < 8; ++q) { d1[q] += (double)f1[q]; d2[q] += (double)f2[q]; f1[q] = 0.f; f2[q] = 0.f; }
      run = 0;
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) { d1[q] += (double)f1[q]; d2[q] += (double)f2[q]; }
  if (MODE == 1 && active) {   // sum dy * xhat = rstd * (sum dy * x - mean * sum dy), in double
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const double mq = (double)mean[(size_t)s * C + c8 * 8 + q];
      const double rq = (double)rsqrtf(var[(size_t)s * C + c8 * 8 + q] + eps);
      d2[q] = rq * (d2[q] - mq * d1[q]);
    }
  }
  // reduce over threadIdx.y in shared memory, then one double atomic per (segment, channel) per block
  __shared__ double red[256 * 8];
  for (int pass = 0; pass < 2; ++pass) {
    double* mine = red + (threadIdx.y * blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) mine[q] = pass == 0 ? d1[q] : d2[q];
    __syncthreads();
    // thread (x, y) adds up channel q = y % 8 (+ 8 per further y group) of column group x over all rows of the tile
    for (int q = threadIdx.y; q < 8; q += blockDim.y) {
      if (c8 < C8) {
        double t = 0.0;
        for (int yy = 0; yy < blockDim.y; ++yy) t += red[(yy * blockDim.x + threadIdx.x) * 8 + q];
        atomicAdd(sums + ((size_t)s * 2 + pass) * C + c8 * 8 + q, t);
      }
    }
    __syncthreads();
  }
}

// Apply kernels: a thread owns 8 consecutive channels and walks down the rows of its block's row range, so the
// per-(segment, channel) coefficients live in registers and are rebuilt only when the segment (video) changes.
// block (bx channel groups, by rows), grid (ceil(C8/bx), row ranges); U rows are in flight per thread.
struct ApplyGeom { dim3 grid, block; long long rows_per_block; };
ApplyGeom apply_geometry(int C, long long rows, int unroll) {
  const int C8 = C >> 3;
  int bx = 1;
  while (bx < C8 && bx < 32) bx <<= 1;
  const int by = 128 / bx;
  const int gx = cdiv(C8, bx);
  long long want = (long long)sm_count() * 12 / gx;           // ~12 blocks of 4 warps per SM in total
  if (want < 1) want = 1;
  long long per = cdiv(rows, want);
  const long long quantum = (long long)by * unroll;            // whole unrolled sweeps
  per = cdiv(per, quantum) * quantum;
  ApplyGeom g;
  g.grid = dim3(gx, (unsigned)cdiv(rows, per));
  g.block = dim3(bx, by);
  g.rows_per_block = per;
  return g;
}

template <int U, bool BF>   // BF: x is bf16 (known at compile time: the raw row buffers shrink to 16 bytes)
__global__ void __launch_bounds__(128)
bn_apply_v8_kernel(const void* __restrict__ x, int xdt_rt, int ldx, const int* __restrict__ row_seg, int row_div,
                   const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ w,
                   const float* __restrict__ b, float eps, int relu, long long rows, long long rows_per_block, int C,
                   void* __restrict__ y, int ydt, int ldy, void* __restrict__ y2, int y2dt, int ldy2) {
  const int xdt = BF ? NLV_BF16 : xdt_rt;
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  if (c8 >= (C >> 3)) return;
  const int c0 = c8 * 8;
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  const V8 ww = nv_ld8(w, NLV_F32, c0), bb = nv_ld8(b, NLV_F32, c0);
  int cur = -1;
  float m[8], k[8];
  for (long long r = r0 + threadIdx.y; r < r1; r += (long long)blockDim.y * U) {
    R8 xr[U];
    int sg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * blockDim.y;
      if (rr < r1) {
        xr[u] = nv_ld8_raw(x, xdt, (size_t)rr * ldx + c0);
        sg[u] = row_seg ? row_seg[rr / row_div] : 0;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * blockDim.y;
      if (rr >= r1) break;
      if (sg[u] != cur) {
        cur = sg[u];
        const V8 mm = nv_ld8(mean, NLV_F32, (size_t)cur * C + c0), vv = nv_ld8(var, NLV_F32, (size_t)cur * C + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) { m[q] = mm.v[q]; k[q] = rsqrtf(vv.v[q] + eps) * ww.v[q]; }
      }
      const V8 xv1 = nv_unpack(xr[u], xdt);
      V8 o;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        o.v[q] = fmaf(xv1.v[q] - m[q], k[q], bb.v[q]);
        if (relu) o.v[q] = fmaxf(o.v[q], 0.f);
      }
      if (y) nv_st8(y, ydt, (size_t)rr * ldy + c0, o);
      if (y2) nv_st8(y2, y2dt, (size_t)rr * ldy2 + c0, o);
    }
  }
}

// dx = w*rstd*(dy - sum_dy/n - xhat*sum_dy_xhat/n) [batch stats] or w*rstd*dy [running stats]; optional ReLU mask by
// yout on dy, optional ReLU gate by `x > 0` on the RESULT (ReLU that precedes the BN: conv -> ReLU -> BN).
template <int U, bool BF>   // BF: dy, x and yout are all bf16
__global__ void __launch_bounds__(128)
bn_bwd_apply_v8_kernel(const void* __restrict__ dy, int dydt_rt, int lddy, const void* __restrict__ x, int xdt_rt, int ldx,
                       const void* __restrict__ yout, int ydt_rt, int ldy, const int* __restrict__ row_seg, int row_div,
                       const int* __restrict__ seg, const float* __restrict__ mean, const float* __restrict__ var,
                       const float* __restrict__ w, float eps, const double* __restrict__ sums, int use_batch_stats,
                       int gate_by_x, long long rows, long long rows_per_block, int C, void* __restrict__ dx, int dxdt, int lddx,
                       float* __restrict__ dx_colsum) {
  const int dydt = BF ? NLV_BF16 : dydt_rt, xdt = BF ? NLV_BF16 : xdt_rt, ydt = BF ? NLV_BF16 : ydt_rt;
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = c8 < (C >> 3);
  if (!active && dx_colsum == nullptr) return;
  const int c0 = active ? c8 * 8 : 0;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};     // column sums of dx: the bias gradient of the layer in front of this BatchNorm
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  const V8 ww = nv_ld8(w, NLV_F32, c0);
  int cur = -1;
  float m[8], k[8], s1[8], t2[8];   // k = w*rstd, s1 = sum_dy/n, t2 = rstd*sum_dy_xhat/n
  for (long long r = r0 + threadIdx.y; active && r < r1; r += (long long)blockDim.y * U) {
    R8 xr[U], gr[U], yr[U];
    int sg[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * blockDim.y;
      if (rr < r1) {
        xr[u] = nv_ld8_raw(x, xdt, (size_t)rr * ldx + c0);
        gr[u] = nv_ld8_raw(dy, dydt, (size_t)rr * lddy + c0);
        if (yout != nullptr) yr[u] = nv_ld8_raw(yout, ydt, (size_t)rr * ldy + c0);
        sg[u] = row_seg ? row_seg[rr / row_div] : 0;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * blockDim.y;
      if (rr >= r1) break;
      if (sg[u] != cur) {
        cur = sg[u];
        const V8 mm = nv_ld8(mean, NLV_F32, (size_t)cur * C + c0), vv = nv_ld8(var, NLV_F32, (size_t)cur * C + c0);
        const float n = use_batch_stats ? (float)(seg[cur + 1] - seg[cur]) : 1.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float rs = rsqrtf(vv.v[q] + eps);
          m[q] = mm.v[q];
          k[q] = ww.v[q] * rs;
          if (use_batch_stats) {
            s1[q] = (float)(sums[((size_t)cur * 2 + 0) * C + c0 + q]) / n;
            t2[q] = rs * ((float)(sums[((size_t)cur * 2 + 1) * C + c0 + q]) / n);
          } else { s1[q] = 0.f; t2[q] = 0.f; }
        }
      }
      const V8 xv1 = nv_unpack(xr[u], xdt), g1 = nv_unpack(gr[u], dydt);
      V8 yo1;
      if (yout != nullptr) yo1 = nv_unpack(yr[u], ydt);
      V8 o;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float gq = g1.v[q];
        if (yout != nullptr && !(yo1.v[q] > 0.f)) gq = 0.f;
        o.v[q] = k[q] * (gq - s1[q] - (xv1.v[q] - m[q]) * t2[q]);
        if (gate_by_x && !(xv1.v[q] > 0.f)) o.v[q] = 0.f;
        cs[q] += o.v[q];
      }
      nv_st8(dx, dxdt, (size_t)rr * lddx + c0, o);
    }
  }
  if (dx_colsum != nullptr) {      // reduce over the block's row lanes, then one atomic per channel per block
    __shared__ float red[128 * 8];
    float* mine = red + (threadIdx.y * blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) mine[q] = cs[q];
    __syncthreads();
    if (threadIdx.y == 0 && active) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float t = 0.f;
        for (int yy = 0; yy < (int)blockDim.y; ++yy) t += red[(yy * blockDim.x + threadIdx.x) * 8 + q];
        atomicAdd(dx_colsum + c0 + q, t);
      }
    }
  }
}

void sums_geometry(int C, long long rows, int nseg, int unroll, dim3& grid, dim3& block) {
  const int C8 = C >> 3;
  int bx = 1;
  while (bx < C8 && bx < 32) bx <<= 1;
  const int by = 256 / bx;
  const int gx = cdiv(C8, bx);
  const long long per_seg = nseg > 0 ? rows / nseg : rows;
  // ~4 blocks per SM over the whole grid; a block covers at least one unrolled sweep of its tile
  long long splits = cdiv(4LL * sm_count(), (long long)gx * (nseg > 0 ? nseg : 1));
  const long long max_splits = cdiv(per_seg, (long long)by * unroll);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 256) splits = 256;
  grid = dim3(gx, nseg, (unsigned)splits);
  block = dim3(bx, by);
}

}  // namespace

int launch_bn_sums_fwd_v8(const void* x, int xdt, int ld, const int* seg, int nseg, long long rows, int C, double* sums, cudaStream_t s) {
  dim3 grid, block;
  sums_geometry(C, rows, nseg, 8, grid, block);
  if (xdt == NLV_BF16) bn_sums_v8_kernel<0, true, 8><<<grid, block, 0, s>>>(x, xdt, ld, nullptr, 0, 0, nullptr, 0, 0, seg, nullptr, nullptr, 0.f, C, sums);
  else bn_sums_v8_kernel<0, false, 4><<<grid, block, 0, s>>>(x, xdt, ld, nullptr, 0, 0, nullptr, 0, 0, seg, nullptr, nullptr, 0.f, C, sums);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_bn_sums_bwd_v8(const void* dy, int dydt, int lddy, const void* x, int xdt, int ldx, const void* yout, int ydt, int ldy, const int* seg,
                          int nseg, const float* mean, const float* var, float eps, long long rows, int C, double* sums, cudaStream_t s) {
  dim3 grid, block;
  sums_geometry(C, rows, nseg, 4, grid, block);
  if (xdt == NLV_BF16 && dydt == NLV_BF16 && (yout == nullptr || ydt == NLV_BF16))
    bn_sums_v8_kernel<1, true, 4><<<grid, block, 0, s>>>(x, xdt, ldx, dy, dydt, lddy, yout, ydt, ldy, seg, mean, var, eps, C, sums);
  else
    bn_sums_v8_kernel<1, false, 2><<<grid, block, 0, s>>>(x, xdt, ldx, dy, dydt, lddy, yout, ydt, ldy, seg, mean, var, eps, C, sums);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_bn_apply_v8(const void* x, int xdt, int ldx, const int* row_seg, int row_div, const float* mean, const float* var, const float* w,
                       const float* b, float eps, int relu, long long rows, int C, void* y, int ydt, int ldy, void* y2, int y2dt,
                       int ldy2, cudaStream_t s) {
  if (xdt == NLV_BF16) {
    const ApplyGeom g = apply_geometry(C, rows, 4);
    bn_apply_v8_kernel<4, true><<<g.grid, g.block, 0, s>>>(x, xdt, ldx, row_seg, row_div, mean, var, w, b, eps, relu, rows, g.rows_per_block, C, y, ydt,
                                                          ldy, y2, y2dt, ldy2);
  } else {
    const ApplyGeom g = apply_geometry(C, rows, 2);
    bn_apply_v8_kernel<2, false><<<g.grid, g.block, 0, s>>>(x, xdt, ldx, row_seg, row_div, mean, var, w, b, eps, relu, rows, g.rows_per_block, C, y, ydt,
                                                           ldy, y2, y2dt, ldy2);
  }
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_bn_bwd_apply_v8(const void* dy, int dydt, int lddy, const void* x, int xdt, int ldx, const void* yout, int ydt, int ldy,
                           const int* row_seg, int row_div, const int* seg, const float* mean, const float* var, const float* w, float eps,
                           const double* sums, int use_batch_stats, int gate_by_x, long long rows, int C, void* dx, int dxdt, int lddx,
                           float* dx_colsum, cudaStream_t s) {
  if (dydt == NLV_BF16 && xdt == NLV_BF16 && (yout == nullptr || ydt == NLV_BF16)) {
    const ApplyGeom g = apply_geometry(C, rows, 4);
    bn_bwd_apply_v8_kernel<4, true><<<g.grid, g.block, 0, s>>>(dy, dydt, lddy, x, xdt, ldx, yout, ydt, ldy, row_seg, row_div, seg, mean, var, w, eps, sums,
                                                              use_batch_stats, gate_by_x, rows, g.rows_per_block, C, dx, dxdt, lddx, dx_colsum);
  } else {
    const ApplyGeom g = apply_geometry(C, rows, 2);
    bn_bwd_apply_v8_kernel<2, false><<<g.grid, g.block, 0, s>>>(dy, dydt, lddy, x, xdt, ldx, yout, ydt, ldy, row_seg, row_div, seg, mean, var, w, eps, sums,
                                                               use_batch_stats, gate_by_x, rows, g.rows_per_block, C, dx, dxdt, lddx, dx_colsum);
  }
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}


// ------------------------------------------------------------------------------------------
// LayerNorm, 16-byte vectorised: one warp per row, lane owns the float4 words lane + 32 j (cols % 4 == 0, cols <= 128 NV).
// All loads of a row are issued back to back (volatile asm keeps the compiler from threading each load through its
// consumer), so a warp pays one memory latency per row with 16-32 x 512 B in flight.
// ------------------------------------------------------------------------------------------
namespace {

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_row4(void* base, int dt, size_t elem, float4 o) {
  if (dt == NLV_BF16) {
    uint2 t;
    *reinterpret_cast<__nv_bfloat162*>(&t.x) = __floats2bfloat162_rn(o.x, o.y);
    *reinterpret_cast<__nv_bfloat162*>(&t.y) = __floats2bfloat162_rn(o.z, o.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + elem) = t;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + elem) = o;
  }
}

template <int NV>
__global__ void __launch_bounds__(128)
layernorm_fwd_v4_kernel(const float* __restrict__ x, long long rows, int cols, const float* __restrict__ w,
                        const float* __restrict__ b, float eps, float* __restrict__ y, void* __restrict__ y2, int y2dt,
                        float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const int nv = cols >> 2;
  const float* xr = x + (size_t)row * cols;
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    v[j] = ldg_f4(xr + 4 * (i < nv ? i : 0));
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (lane + 32 * j >= nv) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(s) / (float)cols;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (lane + 32 * j < nv) {
      const float a = v[j].x - mean, bq = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += (a * a + bq * bq) + (c * c + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    if (i < nv) {
      const float4 ww = *reinterpret_cast<const float4*>(w + 4 * i), bb = *reinterpret_cast<const float4*>(b + 4 * i);
      float4 o;
      o.x = (v[j].x - mean) * rstd * ww.x + bb.x; o.y = (v[j].y - mean) * rstd * ww.y + bb.y;
      o.z = (v[j].z - mean) * rstd * ww.z + bb.z; o.w = (v[j].w - mean) * rstd * ww.w + bb.w;
      if (y) *reinterpret_cast<float4*>(y + (size_t)row * cols + 4 * i) = o;
      if (y2) st_row4(y2, y2dt, (size_t)row * cols + 4 * i, o);
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(128)
layernorm_bwd_dx_v4_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                           const float* __restrict__ rstd, const float* __restrict__ w, long long rows, int cols,
                           float* __restrict__ dx, void* __restrict__ dx2, int dx2dt, const DropCfg drop) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const int nv = cols >> 2;
  const float* dyr = dy + (size_t)row * cols;
  const float* xr = x + (size_t)row * cols;
  float4 g[NV], xh[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j, ic = 4 * (i < nv ? i : 0);
    g[j] = ldg_f4(dyr + ic);
    xh[j] = ldg_f4(xr + ic);
  }
  const float m = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    if (i < nv) {
      const float4 ww = *reinterpret_cast<const float4*>(w + 4 * i);
      xh[j].x = (xh[j].x - m) * rs; xh[j].y = (xh[j].y - m) * rs; xh[j].z = (xh[j].z - m) * rs; xh[j].w = (xh[j].w - m) * rs;
      g[j].x *= ww.x; g[j].y *= ww.y; g[j].z *= ww.z; g[j].w *= ww.w;
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
      s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
    }
  }
  s1 = warp_sum(s1) / (float)cols;
  s2 = warp_sum(s2) / (float)cols;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = lane + 32 * j;
    if (i < nv) {
      float4 o;
      o.x = rs * (g[j].x - s1 - xh[j].x * s2); o.y = rs * (g[j].y - s1 - xh[j].y * s2);
      o.z = rs * (g[j].z - s1 - xh[j].z * s2); o.w = rs * (g[j].w - s1 - xh[j].w * s2);
      if (dx) *reinterpret_cast<float4*>(dx + (size_t)row * cols + 4 * i) = o;
      if (dx2) {
        if (drop.thr16 != 0u) {   // second output = mask * dx / (1 - p): columns 4 i .. 4 i + 3 are one nibble of group i / 2
          const uint32_t keep = keep8_matrix(drop, row, i >> 1, (cols + 7) >> 3) >> ((i & 1) * 4);
          o.x = (keep & 1u) ? o.x * drop.scale : 0.f; o.y = (keep & 2u) ? o.y * drop.scale : 0.f;
          o.z = (keep & 4u) ? o.z * drop.scale : 0.f; o.w = (keep & 8u) ? o.w * drop.scale : 0.f;
        }
        st_row4(dx2, dx2dt, (size_t)row * cols + 4 * i, o);
      }
    }
  }
}


// LayerNorm backward in ONE pass over dy and x (the two-kernel form read both twice: 22 bytes per element instead of 14):
// dx (fp32), its operand copy dx2 (bf16 / fp32, optionally dropout-masked), dw, db AND the column sums of dx2 — the bias
// gradient of the Linear layer that produced the normalised sum, which used to be a separate pass over dx.
// Column-owner layout: thread t owns columns [8 t, 8 t + 8) for every row of the block's row range, so the three per-column
// accumulators live in registers; the two per-row statistics (sum g, sum g * xhat) are block reductions, done for U rows at
// a time behind one barrier (double-buffered scratch).  U rows x 64 bytes are in flight per thread.
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB)
layernorm_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                           const float* __restrict__ rstd, const float* __restrict__ w, long long rows, int cols,
                           long long rows_per_block, float* __restrict__ dx, void* __restrict__ dx2, int dx2dt,
                           float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dprev, const DropCfg drop) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int c0 = t * 8;
  const bool active = c0 < cols;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  __shared__ float red[2][8][2 * U];
  float wv[8], aw[8], ab[8], ap[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { wv[q] = active ? w[c0 + q] : 0.f; aw[q] = 0.f; ab[q] = 0.f; ap[q] = 0.f; }
  const float inv_cols = 1.f / (float)cols;
  const int gpr = (cols + 7) >> 3;
  int buf = 0;
  for (long long r = r0; r < r1; r += U, buf ^= 1) {
    float4 g0[U], g1[U], x0[U], x1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u;
      if (active && rr < r1) {
        const float4* gp = reinterpret_cast<const float4*>(dy + (size_t)rr * cols + c0);
        const float4* xp = reinterpret_cast<const float4*>(x + (size_t)rr * cols + c0);
        g0[u] = __ldg(gp); g1[u] = __ldg(gp + 1); x0[u] = __ldg(xp); x1[u] = __ldg(xp + 1);
      } else {
        g0[u] = g1[u] = x0[u] = x1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float mu[U], rs[U], s[2 * U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u < r1 ? r + u : r1 - 1;
      mu[u] = __ldg(mean + rr); rs[u] = __ldg(rstd + rr);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float* xe = reinterpret_cast<float*>(&x0[u]);      // x0 / x1 become xhat in place
      float* xf = reinterpret_cast<float*>(&x1[u]);
      const float* ge = reinterpret_cast<const float*>(&g0[u]);
      const float* gf = reinterpret_cast<const float*>(&g1[u]);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        xe[q] = (xe[q] - mu[u]) * rs[u]; xf[q] = (xf[q] - mu[u]) * rs[u];
        const float ga = ge[q] * wv[q], gb = gf[q] * wv[4 + q];
        s1 += ga + gb;
        s2 += ga * xe[q] + gb * xf[q];
      }
      if (!active) { s1 = 0.f; s2 = 0.f; }       // xhat of the zero-filled idle columns is not zero
      s[2 * u] = s1; s[2 * u + 1] = s2;
    }
#pragma unroll
    for (int k = 0; k < 2 * U; ++k) s[k] = warp_sum(s[k]);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < 2 * U; ++k) red[buf][warp][k] = s[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2 * U; ++k) {
      float tot = 0.f;
#pragma unroll
      for (int wq = 0; wq < 8; ++wq) tot += red[buf][wq][k];
      s[k] = tot * inv_cols;
    }
    if (!active) continue;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + u;
      if (rr >= r1) break;
      const float* xe = reinterpret_cast<const float*>(&x0[u]);
      const float* xf = reinterpret_cast<const float*>(&x1[u]);
      const float* ge = reinterpret_cast<const float*>(&g0[u]);
      const float* gf = reinterpret_cast<const float*>(&g1[u]);
      float o[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        o[q] = rs[u] * (ge[q] * wv[q] - s[2 * u] - xe[q] * s[2 * u + 1]);
        o[4 + q] = rs[u] * (gf[q] * wv[4 + q] - s[2 * u] - xf[q] * s[2 * u + 1]);
        aw[q] = fmaf(ge[q], xe[q], aw[q]); aw[4 + q] = fmaf(gf[q], xf[q], aw[4 + q]);
        ab[q] += ge[q]; ab[4 + q] += gf[q];
      }
      if (dx != nullptr) {
        float4* op = reinterpret_cast<float4*>(dx + (size_t)rr * cols + c0);
        op[0] = make_float4(o[0], o[1], o[2], o[3]); op[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (drop.thr16 != 0u) {
        const uint32_t keep = keep8_matrix(drop, rr, t, gpr);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = ((keep >> q) & 1u) ? o[q] * drop.scale : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) ap[q] += o[q];
      if (dx2 != nullptr) {
        if (dx2dt == NLV_BF16) {
          uint4 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(o[2 * q], o[2 * q + 1]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dx2) + (size_t)rr * cols + c0) = pk;
        } else {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(dx2) + (size_t)rr * cols + c0);
          op[0] = make_float4(o[0], o[1], o[2], o[3]); op[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      atomicAdd(dw + c0 + q, aw[q]);
      atomicAdd(db + c0 + q, ab[q]);
      if (dprev != nullptr) atomicAdd(dprev + c0 + q, ap[q]);
    }
  }
}

}  // namespace

int launch_layernorm_fwd_v4(const float* x, long long rows, int cols, const float* w, const float* b, float eps, float* y, void* y2,
                            int y2dt, float* mean, float* rstd, cudaStream_t s) {
  const unsigned grid = (unsigned)cdiv(rows, 4);
  if (cols <= 512) layernorm_fwd_v4_kernel<4><<<grid, 128, 0, s>>>(x, rows, cols, w, b, eps, y, y2, y2dt, mean, rstd);
  else if (cols <= 1024) layernorm_fwd_v4_kernel<8><<<grid, 128, 0, s>>>(x, rows, cols, w, b, eps, y, y2, y2dt, mean, rstd);
  else layernorm_fwd_v4_kernel<16><<<grid, 128, 0, s>>>(x, rows, cols, w, b, eps, y, y2, y2dt, mean, rstd);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_layernorm_bwd_dx_v4(const float* dy, const float* x, const float* mean, const float* rstd, const float* w, long long rows,
                               int cols, float* dx, void* dx2, int dx2dt, const nlv_dropout* dr, cudaStream_t s) {
  const unsigned grid = (unsigned)cdiv(rows, 4);
  DropCfg d = drop_off();
  if (dr != nullptr && dr->thr16 != 0u) { d.thr16 = dr->thr16; d.scale = dr->scale; d.seed_lo = dr->seed_lo; d.seed_hi = dr->seed_hi; d.stream = dr->stream; }
  if (cols <= 512) layernorm_bwd_dx_v4_kernel<4><<<grid, 128, 0, s>>>(dy, x, mean, rstd, w, rows, cols, dx, dx2, dx2dt, d);
  else if (cols <= 1024) layernorm_bwd_dx_v4_kernel<8><<<grid, 128, 0, s>>>(dy, x, mean, rstd, w, rows, cols, dx, dx2, dx2dt, d);
  else layernorm_bwd_dx_v4_kernel<16><<<grid, 128, 0, s>>>(dy, x, mean, rstd, w, rows, cols, dx, dx2, dx2dt, d);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int launch_layernorm_bwd_fused(const float* dy, const float* x, const float* mean, const float* rstd, const float* w, long long rows,
                               int cols, float* dx, void* dx2, int dx2dt, float* dw, float* db, float* dprev, const nlv_dropout* dr,
                               cudaStream_t s) {
  DropCfg d = drop_off();
  if (dr != nullptr && dr->thr16 != 0u) { d.thr16 = dr->thr16; d.scale = dr->scale; d.seed_lo = dr->seed_lo; d.seed_hi = dr->seed_hi; d.stream = dr->stream; }
  // two blocks of U = 2 rows per SM (NLV_LN_BWD_U=4: one block of 4 rows; measured alternative)
  static const int u4 = [] { const char* e = getenv("NLV_LN_BWD_U"); return e != nullptr && atoi(e) == 4; }();
  const int U = u4 ? 4 : 2;
  long long blocks = (u4 ? 1LL : 2LL) * sm_count();
  long long per = (rows + blocks - 1) / blocks;
  per = (per + U - 1) / U * U;
  blocks = (rows + per - 1) / per;
  if (u4) layernorm_bwd_fused_kernel<4, 1><<<(unsigned)blocks, 256, 0, s>>>(dy, x, mean, rstd, w, rows, cols, per, dx, dx2, dx2dt, dw, db, dprev, d);
  else layernorm_bwd_fused_kernel<2, 2><<<(unsigned)blocks, 256, 0, s>>>(dy, x, mean, rstd, w, rows, cols, per, dx, dx2, dx2dt, dw, db, dprev, d);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

}  // namespace nlv
