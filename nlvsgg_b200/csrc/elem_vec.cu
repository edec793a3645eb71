// Vectorised (16-byte) variants of the channels-last layout kernels used by the pair-feature conv stack and of the
// flat conversion / column-sum kernels.  Each thread owns 8 consecutive channels (one uint4 of bf16, or two float4),
// so every warp access is a run of full 128-byte lines; index arithmetic is done once per 8 elements.
// Dispatch (elem.cu / norm.cu) falls back to the scalar kernels when C % 8 != 0 or a pointer is not 16-byte aligned.
#include "common.cuh"

namespace nlv {

struct V8 { float v[8]; };

__device__ __forceinline__ V8 ld8(const void* p, int dt, size_t i) {  // i = element index, multiple of 8
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
__device__ __forceinline__ void st8(void* p, int dt, size_t i, const V8& r) {
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
constexpr int TPB = 256;

__global__ void convert8_kernel(const void* __restrict__ src, int sdt, void* __restrict__ dst, int ddt, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  st8(dst, ddt, (size_t)i * 8, ld8(src, sdt, (size_t)i * 8));
}

// im2col 3x3/s1/p1 over NHWC with column order (ky, kx, c): out[row, (ky*3+kx)*C + c]
__global__ void im2col_3x3_v8_kernel(const void* __restrict__ x, int xdt, int H, int W, int C8, long long total,
                                     void* __restrict__ dst, int ddt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % C8);
  const int tap = (int)((i / C8) % 9);
  const long long row = i / ((long long)C8 * 9);
  const int ox = (int)(row % W), oy = (int)((row / W) % H);
  const long long r = row / ((long long)H * W);
  const int iy = oy - 1 + tap / 3, ix = ox - 1 + tap % 3;
  V8 v;
#pragma unroll
  for (int q = 0; q < 8; ++q) v.v[q] = 0.f;
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = ld8(x, xdt, ((size_t)((r * H + iy) * W + ix) * C8 + c8) * 8);
  st8(dst, ddt, (size_t)i * 8, v);
}

// adjoint: dx[r,y,x,c] = sum_taps dcol[(r, y+1-ky, x+1-kx), tap*C + c]
template <typename IT>
__global__ void col2im_3x3_v8_kernel(const void* __restrict__ dcol, int cdt, int H, int W, int C8, long long total,
                                     float* __restrict__ dx) {
  const IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (IT)total) return;
  const int c8 = (int)(i % (IT)C8);
  const IT pos = i / (IT)C8;
  const int x = (int)(pos % (IT)W), y = (int)((pos / (IT)W) % (IT)H);
  const IT r = pos / (IT)(H * W);
  V8 acc;
#pragma unroll
  for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int oy = y + 1 - tap / 3, ox = x + 1 - tap % 3;
    if (oy >= 0 && oy < H && ox >= 0 && ox < W) {
      const V8 t = ld8(dcol, cdt, (((size_t)((r * H + oy) * W + ox) * 9 + tap) * C8 + c8) * 8);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc.v[q] += t.v[q];
    }
  }
  st8(dx, NLV_F32, (size_t)i * 8, acc);
}

template <typename IT>
__global__ void maxpool_fwd_v8_kernel(const void* __restrict__ x, int xdt, int C8, long long total, void* __restrict__ y, int ydt,
                                      uint8_t* __restrict__ arg) {
  const IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (IT)total) return;
  const int c8 = (int)(i % (IT)C8);
  const IT pos = i / (IT)C8;
  const int ox = (int)(pos % 7), oy = (int)((pos / 7) % 7);
  const IT r = pos / 49;
  V8 best;
  int bi[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { best.v[q] = -INFINITY; bi[q] = 0; }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = oy * 2 - 1 + tap / 3, ix = ox * 2 - 1 + tap % 3;
    if (iy >= 0 && iy < 14 && ix >= 0 && ix < 14) {
      const V8 v = ld8(x, xdt, ((size_t)((r * 14 + iy) * 14 + ix) * C8 + c8) * 8);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (v.v[q] > best.v[q] || (v.v[q] != v.v[q])) { best.v[q] = v.v[q]; bi[q] = tap; }
    }
  }
  st8(y, ydt, (size_t)i * 8, best);
  uint2 packed;
  packed.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
  packed.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
  *reinterpret_cast<uint2*>(arg + (size_t)i * 8) = packed;
}

template <typename IT>
__global__ void maxpool_bwd_v8_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ arg, int C8, long long total,
                                      void* __restrict__ dx, int dxdt) {
  const IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (IT)total) return;
  const int c8 = (int)(i % (IT)C8);
  const IT pos = i / (IT)C8;
  const int ix = (int)(pos % 14), iy = (int)((pos / 14) % 14);
  const IT r = pos / 196;
  V8 acc;
#pragma unroll
  for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int t = iy + 1 - ky;
    if (t < 0 || (t & 1) || (t >> 1) >= 7) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int u = ix + 1 - kx;
      if (u < 0 || (u & 1) || (u >> 1) >= 7) continue;
      const size_t o = ((size_t)((r * 7 + (t >> 1)) * 7 + (u >> 1)) * C8 + c8) * 8;
      const uint2 a = *reinterpret_cast<const uint2*>(arg + o);
      const V8 g = ld8(dy, NLV_F32, o);
      const int tap = ky * 3 + kx;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int sel = ((q < 4 ? a.x : a.y) >> (8 * (q & 3))) & 0xff;
        if (sel == tap) acc.v[q] += g.v[q];
      }
    }
  }
  st8(dx, dxdt, (size_t)i * 8, acc);
}

// im2col of the 2x27x27 masks, 8 output columns per thread (ld % 8 == 0)
__global__ void im2col_mask_v8_kernel(const float* __restrict__ m, long long total, void* __restrict__ dst, int ddt, int ld8n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cb = (int)(i % ld8n) * 8;
  const long long row = i / ld8n;
  const int ox = (int)(row % 14), oy = (int)((row / 14) % 14);
  const long long r = row / 196;
  V8 v;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int col = cb + q;
    float t = 0.f;
    if (col < 98) {
      const int c = col / 49, ky = (col % 49) / 7, kx = col % 7;
      const int iy = oy * 2 - 3 + ky, ix = ox * 2 - 3 + kx;
      if (iy >= 0 && iy < 27 && ix >= 0 && ix < 27) t = m[((r * 2 + c) * 27 + iy) * 27 + ix];
    }
    v.v[q] = t;
  }
  st8(dst, ddt, (size_t)i * 8, v);
}

// Per-pair variants of the two im2col kernels for the shapes of the mask conv stack (bf16 output): the pair's input map is
// parked in shared memory, every thread then writes consecutive 16-byte pieces of the pair's output rows, with 32-bit
// constant-divisor index arithmetic only (the generic kernels above spend their time in 64-bit divisions).
__global__ void __launch_bounds__(256)
im2col_3x3_pair_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ dst) {
  constexpr int HW = 49, C16 = 16, ROW16 = 9 * C16;            // 128 channels = 16 uint4; 1152 columns = 144 uint4
  __shared__ uint4 tile[HW * C16];
  const long long r = blockIdx.x;
  const uint4* src = reinterpret_cast<const uint4*>(x) + r * (HW * C16);
  for (int i = threadIdx.x; i < HW * C16; i += 256) tile[i] = src[i];
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(dst) + r * (HW * ROW16);
  for (int e = threadIdx.x; e < HW * ROW16; e += 256) {
    const int row = e / ROW16, q = e - row * ROW16;
    const int tap = q >> 4, c16 = q & 15;
    const int oy = row / 7, ox = row - oy * 7;
    const int iy = oy - 1 + tap / 3, ix = ox - 1 + tap % 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < 7 && ix >= 0 && ix < 7) v = tile[(iy * 7 + ix) * C16 + c16];
    out[e] = v;
  }
}

__global__ void __launch_bounds__(256)
im2col_mask_pair_kernel(const float* __restrict__ m, __nv_bfloat16* __restrict__ dst) {
  constexpr int LD8 = 13;                                      // 104 columns = 13 uint4 (98 used, 6 zero)
  __shared__ float tile[2 * 27 * 27];
  const long long r = blockIdx.x;
  const float* src = m + r * (2 * 27 * 27);
  for (int i = threadIdx.x; i < 2 * 27 * 27; i += 256) tile[i] = src[i];
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(dst) + r * (196 * LD8);
  for (int e = threadIdx.x; e < 196 * LD8; e += 256) {
    const int row = e / LD8, cb = (e - row * LD8) * 8;
    const int oy = row / 14, ox = row - oy * 14;
    float t[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int col = cb + q;
      t[q] = 0.f;
      if (col < 98) {
        const int c = col / 49, rem = col - c * 49, ky = rem / 7, kx = rem - ky * 7;
        const int iy = oy * 2 - 3 + ky, ix = ox * 2 - 3 + kx;
        if (iy >= 0 && iy < 27 && ix >= 0 && ix < 27) t[q] = tile[(c * 27 + iy) * 27 + ix];
      }
    }
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(t[2 * q], t[2 * q + 1]);
    out[e] = v;
  }
}

// column sums with 8 channels per thread: block (C8x, 256/C8x) ; grid (ceil(C8/bx), row splits).
// NC = 1 or 2 row classes accumulated in ONE pass (the decoder's per-slot sums).  BF = the input is bf16, known at compile
// time: the U rows a thread has in flight are held as RAW 16 / 32-byte registers and unpacked only after all U loads were
// issued (with a run-time dtype branch around load + unpack the compiler serialised the bf16 loads: one in flight per thread,
// ~30% of the HBM rate).  U x 256 threads x 16-32 B x 4 blocks = 64-128 KB outstanding per SM.
template <int NC, bool BF>
__global__ void __launch_bounds__(256)
colsum_v8_kernel(const void* __restrict__ x, int ld, long long rows, int C8,
                 const int* __restrict__ row_class, float* __restrict__ out, int cols) {
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * per, r1 = min(rows, r0 + per);
  float acc[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[c][q] = 0.f;
  constexpr int U = BF ? 8 : 4;
  if (c8 < C8) {
    const char* base = reinterpret_cast<const char*>(x) + (size_t)c8 * (BF ? 16 : 32);
    const size_t row_bytes = (size_t)ld * (BF ? 2 : 4);
    for (long long r = r0 + threadIdx.y; r < r1; r += (long long)blockDim.y * U) {
      uint4 ra[U], rb[U];
      int cl[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long rr = r + (long long)u * blockDim.y;
        cl[u] = -1;
        ra[u] = make_uint4(0u, 0u, 0u, 0u);
        rb[u] = ra[u];
        if (rr < r1) {
          const uint4* q = reinterpret_cast<const uint4*>(base + (size_t)rr * row_bytes);
          ra[u] = __ldg(q);
          if (!BF) rb[u] = __ldg(q + 1);
          cl[u] = (NC > 1 && row_class != nullptr) ? __ldg(row_class + rr) : 0;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float f[8];
        if (BF) {
          const uint32_t w[4] = {ra[u].x, ra[u].y, ra[u].z, ra[u].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { f[2 * q] = __uint_as_float(w[q] << 16); f[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u); }
        } else {
          f[0] = __uint_as_float(ra[u].x); f[1] = __uint_as_float(ra[u].y); f[2] = __uint_as_float(ra[u].z); f[3] = __uint_as_float(ra[u].w);
          f[4] = __uint_as_float(rb[u].x); f[5] = __uint_as_float(rb[u].y); f[6] = __uint_as_float(rb[u].z); f[7] = __uint_as_float(rb[u].w);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (cl[u] == c) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[c][q] += f[q];
          }
      }
    }
  }
  // reduce over threadIdx.y in shared memory, then one atomic per column per block
  __shared__ float red[256 * 8];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float* mine = red + (threadIdx.y * blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) mine[q] = acc[c][q];
    __syncthreads();
    if (threadIdx.y == 0 && c8 < C8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float t = 0.f;
        for (int yy = 0; yy < blockDim.y; ++yy) t += red[(yy * blockDim.x + threadIdx.x) * 8 + q];
        atomicAdd(out + (size_t)c * cols + c8 * 8 + q, t);
      }
    }
    __syncthreads();
  }
}

// any number of classes: one pass per class
__global__ void colsum_v8_multi_kernel(const void* __restrict__ x, int xdt, int ld, long long rows, int C8,
                                       const int* __restrict__ row_class, int n_class, float* __restrict__ out, int cols) {
  const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
  const long long per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * per, r1 = min(rows, r0 + per);
  for (int cls = 0; cls < n_class; ++cls) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (c8 < C8)
      for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y)
        if (row_class == nullptr || row_class[r] == cls) {
          const V8 v = ld8(x, xdt, (size_t)r * ld + (size_t)c8 * 8);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] += v.v[q];
        }
    __shared__ float red[256 * 8];
    float* mine = red + (threadIdx.y * blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) mine[q] = acc[q];
    __syncthreads();
    if (threadIdx.y == 0 && c8 < C8) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float t = 0.f;
        for (int yy = 0; yy < blockDim.y; ++yy) t += red[(yy * blockDim.x + threadIdx.x) * 8 + q];
        atomicAdd(out + (size_t)cls * cols + c8 * 8 + q, t);
      }
    }
    __syncthreads();
  }
}

// row gather with optional add, 4 floats per thread (cols % 4 == 0), fp32 source
__global__ void gather_rows_v4_kernel(const float* __restrict__ src, int lds, const int* __restrict__ idx,
                                      const float* __restrict__ add, const int* __restrict__ add_idx, int ld_add, long long n_out,
                                      int cols4, float* __restrict__ dst, int ldd, void* __restrict__ dst2, int d2dt, int ldd2) {
  // one row per CTA (no index division); the row's source indices are read once
  const long long r = blockIdx.x;
  const int s = idx ? idx[r] : (int)r;
  const int sa = add != nullptr ? (add_idx ? add_idx[r] : (int)r) : 0;
  for (int c = threadIdx.x * 4; c < cols4 * 4; c += blockDim.x * 4) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s >= 0) {
    v = *reinterpret_cast<const float4*>(src + (size_t)s * lds + c);
    if (add != nullptr) {
      const float4 a = *reinterpret_cast<const float4*>(add + (size_t)sa * ld_add + c);
      v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    }
  }
  if (dst) *reinterpret_cast<float4*>(dst + (size_t)r * ldd + c) = v;
  if (dst2) {
    if (d2dt == NLV_BF16) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
      uint2 t; t.x = *reinterpret_cast<uint32_t*>(&h0); t.y = *reinterpret_cast<uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dst2) + (size_t)r * ldd2 + c) = t;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst2) + (size_t)r * ldd2 + c) = v;
    }
  }
  }
}

}  // namespace

// ---- launchers used by elem.cu ----
#define GRIDV(total) cdiv((total), TPB), TPB, 0, s
int launch_convert8(const void* src, int sdt, void* dst, int ddt, long long n, cudaStream_t s) {
  convert8_kernel<<<GRIDV(n / 8)>>>(src, sdt, dst, ddt, n / 8);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_im2col_3x3_v8(const void* x, int xdt, int r, int h, int w, int c, void* dst, int ddt, cudaStream_t s) {
  if (xdt == NLV_BF16 && ddt == NLV_BF16 && h == 7 && w == 7 && c == 128) {
    im2col_3x3_pair_kernel<<<r, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)dst);
    NLV_CHECK_LAUNCH();
    return NLV_OK;
  }
  const long long total = (long long)r * h * w * 9 * (c / 8);
  im2col_3x3_v8_kernel<<<GRIDV(total)>>>(x, xdt, h, w, c / 8, total, dst, ddt);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_col2im_3x3_v8(const void* dcol, int cdt, int r, int h, int w, int c, float* dx, cudaStream_t s) {
  const long long total = (long long)r * h * w * (c / 8);
  if (total < (1ll << 31)) col2im_3x3_v8_kernel<unsigned><<<GRIDV(total)>>>(dcol, cdt, h, w, c / 8, total, dx);
  else col2im_3x3_v8_kernel<long long><<<GRIDV(total)>>>(dcol, cdt, h, w, c / 8, total, dx);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_maxpool_fwd_v8(const void* x, int xdt, int r, int c, void* y, int ydt, uint8_t* arg, cudaStream_t s) {
  const long long total = (long long)r * 49 * (c / 8);
  if (total < (1ll << 31)) maxpool_fwd_v8_kernel<unsigned><<<GRIDV(total)>>>(x, xdt, c / 8, total, y, ydt, arg);
  else maxpool_fwd_v8_kernel<long long><<<GRIDV(total)>>>(x, xdt, c / 8, total, y, ydt, arg);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_maxpool_bwd_v8(const float* dy, const uint8_t* arg, int r, int c, void* dx, int dxdt, cudaStream_t s) {
  const long long total = (long long)r * 196 * (c / 8);
  if (total < (1ll << 31)) maxpool_bwd_v8_kernel<unsigned><<<GRIDV(total)>>>(dy, arg, c / 8, total, dx, dxdt);
  else maxpool_bwd_v8_kernel<long long><<<GRIDV(total)>>>(dy, arg, c / 8, total, dx, dxdt);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_im2col_mask_v8(const float* m, int r, void* dst, int ddt, int ld, cudaStream_t s) {
  if (ddt == NLV_BF16 && ld == 104) {
    im2col_mask_pair_kernel<<<r, 256, 0, s>>>(m, (__nv_bfloat16*)dst);
    NLV_CHECK_LAUNCH();
    return NLV_OK;
  }
  const long long total = (long long)r * 196 * (ld / 8);
  im2col_mask_v8_kernel<<<GRIDV(total)>>>(m, total, dst, ddt, ld / 8);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_colsum_v8(const void* x, int xdt, int ld, long long rows, int cols, const int* row_class, int n_class, float* out,
                     cudaStream_t s) {
  const int C8 = cols / 8;
  int bx = 1;
  while (bx < C8 && bx < 32) bx <<= 1;
  const int by = 256 / bx;
  const int gx = cdiv(C8, bx);
  // ~4 blocks per SM, but at least 4 * by rows per block (one unrolled iteration per thread)
  long long splits = (4ll * sm_count() + gx - 1) / gx;
  const long long max_splits = (rows + 4 * by - 1) / (4 * by);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  dim3 grid(gx, (unsigned)splits), block(bx, by);
  const bool bf = xdt == NLV_BF16;
  if (n_class == 1 || row_class == nullptr) {
    if (bf) colsum_v8_kernel<1, true><<<grid, block, 0, s>>>(x, ld, rows, C8, nullptr, out, cols);
    else colsum_v8_kernel<1, false><<<grid, block, 0, s>>>(x, ld, rows, C8, nullptr, out, cols);
  } else if (n_class == 2) {
    if (bf) colsum_v8_kernel<2, true><<<grid, block, 0, s>>>(x, ld, rows, C8, row_class, out, cols);
    else colsum_v8_kernel<2, false><<<grid, block, 0, s>>>(x, ld, rows, C8, row_class, out, cols);
  }
  else colsum_v8_multi_kernel<<<grid, block, 0, s>>>(x, xdt, ld, rows, C8, row_class, n_class, out, cols);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
int launch_gather_rows_v4(const float* src, int lds, const int* idx, const float* add, const int* add_idx, int ld_add,
                          long long n_out, int cols, float* dst, int ldd, void* dst2, int d2dt, int ldd2, cudaStream_t s) {
  gather_rows_v4_kernel<<<(unsigned)n_out, 128, 0, s>>>(src, lds, idx, add, add_idx, ld_add, n_out, cols / 4, dst, ldd, dst2, d2dt, ldd2);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

}  // namespace nlv
