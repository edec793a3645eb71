// Layout, conversion, gather/scatter and small reduction kernels (all HBM-bound, coalesced along
// the contiguous dimension, grid sized from the element count).
#include "common.cuh"

namespace nlv {
// 16-byte vectorised variants (elem_vec.cu)
int launch_convert8(const void* src, int sdt, void* dst, int ddt, long long n, cudaStream_t s);
int launch_im2col_3x3_v8(const void* x, int xdt, int r, int h, int w, int c, void* dst, int ddt, cudaStream_t s);
int launch_col2im_3x3_v8(const void* dcol, int cdt, int r, int h, int w, int c, float* dx, cudaStream_t s);
int launch_maxpool_fwd_v8(const void* x, int xdt, int r, int c, void* y, int ydt, uint8_t* arg, cudaStream_t s);
int launch_maxpool_bwd_v8(const float* dy, const uint8_t* arg, int r, int c, void* dx, int dxdt, cudaStream_t s);
int launch_im2col_mask_v8(const float* m, int r, void* dst, int ddt, int ld, cudaStream_t s);
int launch_colsum_v8(const void* x, int xdt, int ld, long long rows, int cols, const int* row_class, int n_class, float* out,
                     cudaStream_t s);
int launch_gather_rows_v4(const float* src, int lds, const int* idx, const float* add, const int* add_idx, int ld_add,
                          long long n_out, int cols, float* dst, int ldd, void* dst2, int d2dt, int ldd2, cudaStream_t s);
inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

namespace {

constexpr int TPB = 256;

// ------------------------------------------------------------------------------------------
// 2-D strided convert / copy (f32 <-> bf16)
// ------------------------------------------------------------------------------------------
__global__ void convert_kernel(const void* __restrict__ src, int sdt, int lds, void* __restrict__ dst, int ddt, int ldd,
                               long long rows, int cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  st_from_float(dst, ddt, (size_t)r * ldd + c, ld_as_float(src, sdt, (size_t)r * lds + c));
}

// bf16x3 split: x = hi + lo with hi = bf16(x), lo = bf16(x - hi).  pattern 0 (A operand): hi,hi,lo;
// pattern 1 (B operand): hi,lo,hi, so that sum over the 3 blocks of A'_k * B'_k = hi*hi + hi*lo + lo*hi.
__global__ void split3_kernel(const float* __restrict__ src, int lds, long long rows, int cols, __nv_bfloat16* __restrict__ dst,
                              int ldd, int block_dim, int pattern) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  const float x = src[(size_t)r * lds + c];
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
  const __nv_bfloat16 v1 = pattern == 0 ? hi : lo;
  const __nv_bfloat16 v2 = pattern == 0 ? lo : hi;
  if (block_dim == 1) {
    __nv_bfloat16* o = dst + (size_t)r * ldd + c;
    o[0] = hi; o[cols] = v1; o[2 * (size_t)cols] = v2;
  } else {
    dst[(size_t)r * ldd + c] = hi;
    dst[(size_t)(r + rows) * ldd + c] = v1;
    dst[(size_t)(r + 2 * rows) * ldd + c] = v2;
  }
}

// ------------------------------------------------------------------------------------------
// union_feat [R, C, HW] (NCHW, fp32) -> rows [R*HW, C] (channels contiguous), via a smem transpose.
// One CTA per (pair, 64-channel slab): reads 64 x HW floats contiguously, writes HW rows of 64.
// ------------------------------------------------------------------------------------------
template <int HW>
__global__ void nchw_to_rows_kernel(const void* __restrict__ src, int sdt, int C, void* __restrict__ dst, int ddt) {
  __shared__ float tile[64][HW + 1];
  const int r = blockIdx.y, c0 = blockIdx.x * 64;
  const size_t s0 = ((size_t)r * C + c0) * HW;
  const int nch = min(64, C - c0);
  for (int i = threadIdx.x; i < nch * HW; i += blockDim.x) tile[i / HW][i % HW] = ld_as_float(src, sdt, s0 + i);
  __syncthreads();
  for (int i = threadIdx.x; i < HW * 64; i += blockDim.x) {
    const int hw = i >> 6, c = i & 63;
    if (c < nch) st_from_float(dst, ddt, ((size_t)r * HW + hw) * C + c0 + c, tile[c][hw]);
  }
}

// bf16 output, C % 128 == 0: one CTA per (pair, 128-channel slab).  The slab is 128 x 49 contiguous floats: read with
// 16-byte loads (all of a thread's loads in flight at once), parked in shared memory in the same linear order, and
// written back as bf16 channel pairs: a warp-store is 128 contiguous bytes of one output row.
__global__ void __launch_bounds__(256)
nchw_to_rows_bf16_kernel(const float* __restrict__ src, int C, __nv_bfloat16* __restrict__ dst) {
  constexpr int HW = 49, CS = 128, NF4 = CS * HW / 4;   // 1568 float4 per slab
  __shared__ __align__(16) float tile[CS * HW];
  const int r = blockIdx.y, c0 = blockIdx.x * CS;
  const float4* s4 = reinterpret_cast<const float4*>(src + ((size_t)r * C + c0) * HW);
  float4 v[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int i = threadIdx.x + 256 * k;
    if (i < NF4) v[k] = __ldcs(s4 + i);     // streamed once: do not keep in L2
  }
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int i = threadIdx.x + 256 * k;
    if (i < NF4) reinterpret_cast<float4*>(tile)[i] = v[k];
  }
  __syncthreads();
  __nv_bfloat16* d = dst + (size_t)r * HW * C + c0;
#pragma unroll 5
  for (int e = threadIdx.x; e < HW * (CS / 2); e += 256) {
    const int hw = e >> 6, c = (e & 63) * 2;
    *reinterpret_cast<__nv_bfloat162*>(d + (size_t)hw * C + c) = __floats2bfloat162_rn(tile[c * HW + hw], tile[(c + 1) * HW + hw]);
  }
}

// same for a bf16 source (packed per-video feature files, SURVEY 8f-2: the loader hands bf16, halving the host -> device bytes)
__global__ void __launch_bounds__(256)
nchw_to_rows_bf16in_kernel(const __nv_bfloat16* __restrict__ src, int C, __nv_bfloat16* __restrict__ dst) {
  constexpr int HW = 49, CS = 128, NU4 = CS * HW / 8;   // 784 uint4 per slab
  __shared__ __align__(16) __nv_bfloat16 tile[CS * HW];
  const int r = blockIdx.y, c0 = blockIdx.x * CS;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + ((size_t)r * C + c0) * HW);
  uint4 v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + 256 * k;
    if (i < NU4) v[k] = __ldcs(s4 + i);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = threadIdx.x + 256 * k;
    if (i < NU4) reinterpret_cast<uint4*>(tile)[i] = v[k];
  }
  __syncthreads();
  __nv_bfloat16* d = dst + (size_t)r * HW * C + c0;
#pragma unroll 5
  for (int e = threadIdx.x; e < HW * (CS / 2); e += 256) {
    const int hw = e >> 6, c = (e & 63) * 2;
    __nv_bfloat162 o;
    o.x = tile[c * HW + hw];
    o.y = tile[(c + 1) * HW + hw];
    *reinterpret_cast<__nv_bfloat162*>(d + (size_t)hw * C + c) = o;
  }
}

// ------------------------------------------------------------------------------------------
// im2col for Conv2d(2,128,k7,s2,p3) on spatial masks [R,2,27,27] -> [R*14*14, ld] (98 cols, zero pad)
// column order (c, ky, kx) = the flattening of conv.0.weight [128, 2,7,7]  (lib/sttran.py:338)
// ------------------------------------------------------------------------------------------
__global__ void im2col_mask_kernel(const float* __restrict__ m, long long total, void* __restrict__ dst, int ddt, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int col = (int)(i % ld);
  const long long row = i / ld;
  float v = 0.f;
  if (col < 98) {
    const int ox = (int)(row % 14), oy = (int)((row / 14) % 14);
    const long long r = row / 196;
    const int c = col / 49, ky = (col % 49) / 7, kx = col % 7;
    const int iy = oy * 2 - 3 + ky, ix = ox * 2 - 3 + kx;
    if (iy >= 0 && iy < 27 && ix >= 0 && ix < 27) v = m[((r * 2 + c) * 27 + iy) * 27 + ix];
  }
  st_from_float(dst, ddt, (size_t)i, v);
}

// im2col for a 3x3/s1/p1 conv over NHWC [R,H,W,C] -> [R*H*W, 9*C], column = (ky*3 + kx)*C + c
__global__ void im2col_3x3_kernel(const void* __restrict__ x, int xdt, int H, int W, int C, long long total,
                                  void* __restrict__ dst, int ddt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int K = C * 9;
  const int col = (int)(i % K);
  const long long row = i / K;
  const int c = col % C, ky = (col / C) / 3, kx = (col / C) % 3;
  const int ox = (int)(row % W), oy = (int)((row / W) % H);
  const long long r = row / ((long long)H * W);
  const int iy = oy - 1 + ky, ix = ox - 1 + kx;
  float v = 0.f;
  if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = ld_as_float(x, xdt, (size_t)((r * H + iy) * W + ix) * C + c);
  st_from_float(dst, ddt, (size_t)i, v);
}

// transpose of the above (gather form, no atomics): dx[r,y,x,c] = sum_{ky,kx} dcol[(r,y+1-ky,x+1-kx), (ky*3+kx)*C+c]
__global__ void col2im_3x3_kernel(const void* __restrict__ dcol, int cdt, int H, int W, int C, long long total,
                                  float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long pos = i / C;
  const int x = (int)(pos % W), y = (int)((pos / W) % H);
  const long long r = pos / ((long long)H * W);
  const int K = C * 9;
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int oy = y + 1 - ky, ox = x + 1 - kx;
      if (oy >= 0 && oy < H && ox >= 0 && ox < W)
        acc += ld_as_float(dcol, cdt, (size_t)((r * H + oy) * W + ox) * K + (ky * 3 + kx) * C + c);
    }
  dx[i] = acc;
}

// MaxPool2d(3, stride 2, pad 1) on NHWC [R,14,14,C] -> [R,7,7,C]; argmax stored as ky*3+kx (first max wins,
// scanning ky then kx, as ATen's max_pool2d does).
__global__ void maxpool_fwd_kernel(const void* __restrict__ x, int xdt, int C, long long total, void* __restrict__ y,
                                   int ydt, uint8_t* __restrict__ arg) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long pos = i / C;
  const int ox = (int)(pos % 7), oy = (int)((pos / 7) % 7);
  const long long r = pos / 49;
  float best = -INFINITY;
  int bi = 0;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int iy = oy * 2 - 1 + ky, ix = ox * 2 - 1 + kx;
      if (iy >= 0 && iy < 14 && ix >= 0 && ix < 14) {
        const float v = ld_as_float(x, xdt, (size_t)((r * 14 + iy) * 14 + ix) * C + c);
        if (v > best || (v != v)) { best = v; bi = ky * 3 + kx; }
      }
    }
  st_from_float(y, ydt, (size_t)i, best);
  arg[i] = (uint8_t)bi;
}

// dx[r,iy,ix,c] = sum over the (<=4) pooling windows that contain (iy,ix) and selected it
__global__ void maxpool_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ arg, int C, long long total,
                                   void* __restrict__ dx, int dxdt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const long long pos = i / C;
  const int ix = (int)(pos % 14), iy = (int)((pos / 14) % 14);
  const long long r = pos / 196;
  float acc = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int t = iy + 1 - ky;
    if (t < 0 || (t & 1)) continue;
    const int oy = t >> 1;
    if (oy >= 7) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int u = ix + 1 - kx;
      if (u < 0 || (u & 1)) continue;
      const int ox = u >> 1;
      if (ox >= 7) continue;
      const size_t o = (size_t)((r * 7 + oy) * 7 + ox) * C + c;
      if (arg[o] == ky * 3 + kx) acc += dy[o];
    }
  }
  st_from_float(dx, dxdt, (size_t)i, acc);
}

// ------------------------------------------------------------------------------------------
// Row gather (optionally adding a per-row class vector, e.g. a positional encoding) and row gather-sum
// ------------------------------------------------------------------------------------------
// dst[i,:] = src[idx[i],:] (+ add[add_idx[i],:]);  idx[i] < 0 -> zeros.  Second optional output dst2 (bf16/f32 copy).
__global__ void gather_rows_kernel(const void* __restrict__ src, int sdt, int lds, const int* __restrict__ idx,
                                   const float* __restrict__ add, const int* __restrict__ add_idx, int ld_add,
                                   long long n_out, int cols, void* __restrict__ dst, int ddt, int ldd,
                                   void* __restrict__ dst2, int d2dt, int ldd2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  const int s = idx ? idx[r] : (int)r;
  float v = s >= 0 ? ld_as_float(src, sdt, (size_t)s * lds + c) : 0.f;
  if (add != nullptr && s >= 0) v += add[(size_t)(add_idx ? add_idx[r] : (int)r) * ld_add + c];
  if (dst) st_from_float(dst, ddt, (size_t)r * ldd + c, v);
  if (dst2) st_from_float(dst2, d2dt, (size_t)r * ldd2 + c, v);
}

// dst[i,:] = sum_{j<fan} src[idx[i*fan+j],:] over idx >= 0   (adjoint of a gather with bounded fan-out)
__global__ void gather_sum_rows_kernel(const float* __restrict__ src, int lds, const int* __restrict__ idx, int fan,
                                       long long n_out, int cols, float* __restrict__ dst, int ldd, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  float acc = accumulate ? dst[(size_t)r * ldd + c] : 0.f;
  for (int j = 0; j < fan; ++j) {
    const int s = idx[r * fan + j];
    if (s >= 0) acc += src[(size_t)s * lds + c];
  }
  dst[(size_t)r * ldd + c] = acc;
}

__global__ void scale_rows_kernel(const float* __restrict__ src, int lds, const float* __restrict__ w, long long rows, int cols,
                                  float* __restrict__ dst, int ldd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  dst[(size_t)r * ldd + c] = src[(size_t)r * lds + c] * w[r];
}

// scatter-add rows with atomics: dst[idx[i],:] += src[i,:]  (embedding / feature-row gradients)
__global__ void scatter_add_rows_kernel(const float* __restrict__ src, int lds, const long long* __restrict__ idx,
                                        int idx_stride, long long n, int cols, float* __restrict__ dst, int ldd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  atomicAdd(dst + (size_t)idx[r * idx_stride] * ldd + c, src[(size_t)r * lds + c]);
}

// 1936-d relation token assembly (lib/sttran.py:381-399): subject/object projections are gathered from the
// per-box projection table fo[N,1024], class embeddings from the two [37,200] tables; the 512 visual-relation
// columns [1024,1536) are written in place by the vr_fc GEMM.
__global__ void assemble_tokens_kernel(const float* __restrict__ fo, const long long* __restrict__ pair_idx,
                                       const long long* __restrict__ labels, const float* __restrict__ e1,
                                       const float* __restrict__ e2, long long R, float* __restrict__ rel) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 1424) return;
  const long long r = i / 1424;
  const int c = (int)(i - r * 1424);
  const long long s = pair_idx[2 * r], o = pair_idx[2 * r + 1];
  float v;
  int col;
  if (c < 512) { v = fo[s * 1024 + c]; col = c; }
  else if (c < 1024) { v = fo[o * 1024 + c]; col = c; }
  else if (c < 1224) { v = e1[labels[s] * 200 + (c - 1024)]; col = 1536 + (c - 1024); }
  else { v = e2[labels[o] * 200 + (c - 1224)]; col = 1736 + (c - 1224); }
  rel[r * 1936 + col] = v;
}

__global__ void assemble_tokens_bwd_kernel(const float* __restrict__ drel, const long long* __restrict__ pair_idx,
                                           const long long* __restrict__ labels, long long R, float* __restrict__ dfo,
                                           float* __restrict__ de1, float* __restrict__ de2) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 1424) return;
  const long long r = i / 1424;
  const int c = (int)(i - r * 1424);
  const long long s = pair_idx[2 * r], o = pair_idx[2 * r + 1];
  if (c < 512) atomicAdd(dfo + s * 1024 + c, drel[r * 1936 + c]);
  else if (c < 1024) atomicAdd(dfo + o * 1024 + c, drel[r * 1936 + c]);
  else if (c < 1224) atomicAdd(de1 + labels[s] * 200 + (c - 1024), drel[r * 1936 + 1536 + (c - 1024)]);
  else atomicAdd(de2 + labels[o] * 200 + (c - 1224), drel[r * 1936 + 1736 + (c - 1224)]);
}

// (cx, cy, w, h) with +1 — lib/fpn/box_utils.py:51-63 — from boxes[:,1:5]
__global__ void center_size_kernel(const float* __restrict__ boxes, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x1 = boxes[i * 5 + 1], y1 = boxes[i * 5 + 2], x2 = boxes[i * 5 + 3], y2 = boxes[i * 5 + 4];
  const float w = x2 - x1 + 1.0f, h = y2 - y1 + 1.0f;
  out[i * 4 + 0] = x1 + 0.5f * w;
  out[i * 4 + 1] = y1 + 0.5f * h;
  out[i * 4 + 2] = w;
  out[i * 4 + 3] = h;
}

// ------------------------------------------------------------------------------------------
// Column sums (bias gradients), optionally per row class: out[cls, c] (+)= sum_{rows of class} x[row, c]
// grid (ceil(cols/32), row_splits), block 32 x 8
// ------------------------------------------------------------------------------------------
__global__ void colsum_kernel(const void* __restrict__ x, int xdt, int ld, long long rows, int cols,
                              const int* __restrict__ row_class, int n_class, float* __restrict__ out) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long rows_per = (rows + gridDim.y - 1) / gridDim.y;
  const long long r0 = (long long)blockIdx.y * rows_per;
  const long long r1 = min(rows, r0 + rows_per);
  __shared__ float red[8][33];
  for (int cls = 0; cls < n_class; ++cls) {
    float acc = 0.f;
    if (c < cols)
      for (long long r = r0 + threadIdx.y; r < r1; r += 8)
        if (row_class == nullptr || row_class[r] == cls) acc += ld_as_float(x, xdt, (size_t)r * ld + c);
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
      atomicAdd(out + (size_t)cls * cols + c, t);
    }
    __syncthreads();
  }
}

// y = x * (gate > 0) elementwise (ReLU backward); in-place allowed; 2-D strided
__global__ void relu_mask_kernel(const void* __restrict__ x, int xdt, int ldx, const void* __restrict__ gate, int gdt,
                                 int ldg, long long rows, int cols, void* __restrict__ y, int ydt, int ldy) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  const float g = ld_as_float(gate, gdt, (size_t)r * ldg + c);
  const float v = ld_as_float(x, xdt, (size_t)r * ldx + c);
  st_from_float(y, ydt, (size_t)r * ldy + c, g > 0.f ? v : 0.f);
}

// y = a + b (fp32), 1-D
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

}  // namespace
}  // namespace nlv

using namespace nlv;
#define STREAM ((cudaStream_t)stream)
#define GRID1D(total) cdiv((total), TPB), TPB, 0, STREAM

extern "C" {

int nlv_convert(const void* src, int src_dtype, int lds, void* dst, int dst_dtype, int ldd, long long rows, int cols,
                void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols >= 0, "convert: bad sizes");
  if (rows * cols == 0) return NLV_OK;
  NLV_CHECK_ARG(src && dst, "convert: null pointer");
  if ((rows == 1 || (lds == cols && ldd == cols)) && ((rows * cols) & 7) == 0 && al16(src) && al16(dst))
    return launch_convert8(src, src_dtype, dst, dst_dtype, rows * cols, STREAM);
  convert_kernel<<<GRID1D(rows * cols)>>>(src, src_dtype, lds, dst, dst_dtype, ldd, rows, cols);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_split3(const float* src, int lds, long long rows, int cols, void* dst, int ldd, int block_dim, int pattern,
               void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols >= 0 && (block_dim == 0 || block_dim == 1) && (pattern == 0 || pattern == 1),
                "split3: bad arguments");
  if (rows * cols == 0) return NLV_OK;
  NLV_CHECK_ARG(src && dst, "split3: null pointer");
  split3_kernel<<<GRID1D(rows * cols)>>>(src, lds, rows, cols, (__nv_bfloat16*)dst, ldd, block_dim, pattern);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_nchw_to_rows(const void* src, int src_dtype, int r, int c, int hw, void* dst, int dst_dtype, void* stream) {
  NLV_CHECK_ARG(r >= 0 && c > 0, "nchw_to_rows: bad sizes");
  NLV_CHECK_ARG(hw == 49, "nchw_to_rows: only 7x7 maps are supported (hw=%d)", hw);
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(src && dst, "nchw_to_rows: null pointer");
  NLV_CHECK_ARG(r <= 65535, "nchw_to_rows: r=%d exceeds the grid limit; split the call", r);
  if (dst_dtype == NLV_BF16 && (c & 127) == 0 && al16(src) && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
    dim3 grid(c / 128, r);
    if (src_dtype == NLV_BF16) nchw_to_rows_bf16in_kernel<<<grid, 256, 0, STREAM>>>((const __nv_bfloat16*)src, c, (__nv_bfloat16*)dst);
    else nchw_to_rows_bf16_kernel<<<grid, 256, 0, STREAM>>>((const float*)src, c, (__nv_bfloat16*)dst);
    NLV_CHECK_LAUNCH();
    return NLV_OK;
  }
  dim3 grid(cdiv(c, 64), r);
  nchw_to_rows_kernel<49><<<grid, 256, 0, STREAM>>>(src, src_dtype, c, dst, dst_dtype);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_im2col_mask(const float* masks, int r, void* dst, int dst_dtype, int ldd, void* stream) {
  NLV_CHECK_ARG(r >= 0 && ldd >= 98, "im2col_mask: bad sizes");
  if (r == 0) return NLV_OK;
  if ((ldd & 7) == 0 && al16(dst)) return launch_im2col_mask_v8(masks, r, dst, dst_dtype, ldd, STREAM);
  const long long total = (long long)r * 196 * ldd;
  im2col_mask_kernel<<<GRID1D(total)>>>(masks, total, dst, dst_dtype, ldd);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_im2col_3x3(const void* x, int x_dtype, int r, int h, int w, int c, void* dst, int dst_dtype, void* stream) {
  NLV_CHECK_ARG(r >= 0 && h > 0 && w > 0 && c > 0, "im2col_3x3: bad sizes");
  if (r == 0) return NLV_OK;
  if ((c & 7) == 0 && al16(x) && al16(dst)) return launch_im2col_3x3_v8(x, x_dtype, r, h, w, c, dst, dst_dtype, STREAM);
  const long long total = (long long)r * h * w * c * 9;
  im2col_3x3_kernel<<<GRID1D(total)>>>(x, x_dtype, h, w, c, total, dst, dst_dtype);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_col2im_3x3(const void* dcol, int dtype, int r, int h, int w, int c, float* dx, void* stream) {
  NLV_CHECK_ARG(r >= 0 && h > 0 && w > 0 && c > 0, "col2im_3x3: bad sizes");
  if (r == 0) return NLV_OK;
  if ((c & 7) == 0 && al16(dcol) && al16(dx)) return launch_col2im_3x3_v8(dcol, dtype, r, h, w, c, dx, STREAM);
  const long long total = (long long)r * h * w * c;
  col2im_3x3_kernel<<<GRID1D(total)>>>(dcol, dtype, h, w, c, total, dx);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_maxpool_fwd(const void* x, int x_dtype, int r, int c, void* y, int y_dtype, uint8_t* argmax, void* stream) {
  NLV_CHECK_ARG(r >= 0 && c > 0, "maxpool_fwd: bad sizes");
  if (r == 0) return NLV_OK;
  if ((c & 7) == 0 && al16(x) && al16(y) && al16(argmax)) return launch_maxpool_fwd_v8(x, x_dtype, r, c, y, y_dtype, argmax, STREAM);
  const long long total = (long long)r * 49 * c;
  maxpool_fwd_kernel<<<GRID1D(total)>>>(x, x_dtype, c, total, y, y_dtype, argmax);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_maxpool_bwd(const float* dy, const uint8_t* argmax, int r, int c, void* dx, int dx_dtype, void* stream) {
  NLV_CHECK_ARG(r >= 0 && c > 0, "maxpool_bwd: bad sizes");
  if (r == 0) return NLV_OK;
  if ((c & 7) == 0 && al16(dy) && al16(dx) && al16(argmax)) return launch_maxpool_bwd_v8(dy, argmax, r, c, dx, dx_dtype, STREAM);
  const long long total = (long long)r * 196 * c;
  maxpool_bwd_kernel<<<GRID1D(total)>>>(dy, argmax, c, total, dx, dx_dtype);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_gather_rows(const void* src, int src_dtype, int lds, const int* idx, const float* add, const int* add_idx,
                    int ld_add, long long n_out, int cols, void* dst, int dst_dtype, int ldd, void* dst2,
                    int dst2_dtype, int ldd2, void* stream) {
  NLV_CHECK_ARG(n_out >= 0 && cols >= 0, "gather_rows: bad sizes");
  if (n_out * cols == 0) return NLV_OK;
  NLV_CHECK_ARG(src && (dst || dst2), "gather_rows: null pointer");
  if (src_dtype == NLV_F32 && (dst == nullptr || dst_dtype == NLV_F32) && (cols & 3) == 0 && (lds & 3) == 0 && (ldd & 3) == 0 &&
      (ldd2 & 3) == 0 && (ld_add & 3) == 0 && al16(src) && al16(dst) && al16(dst2) && al16(add))
    return launch_gather_rows_v4((const float*)src, lds, idx, add, add_idx, ld_add, n_out, cols, (float*)dst, ldd, dst2, dst2_dtype,
                                 ldd2, STREAM);
  gather_rows_kernel<<<GRID1D(n_out * cols)>>>(src, src_dtype, lds, idx, add, add_idx, ld_add, n_out, cols, dst,
                                              dst_dtype, ldd, dst2, dst2_dtype, ldd2);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_gather_sum_rows(const float* src, int lds, const int* idx, int fan, long long n_out, int cols, float* dst,
                        int ldd, int accumulate, void* stream) {
  NLV_CHECK_ARG(n_out >= 0 && cols >= 0 && fan >= 1, "gather_sum_rows: bad sizes");
  if (n_out * cols == 0) return NLV_OK;
  NLV_CHECK_ARG(src && idx && dst, "gather_sum_rows: null pointer");
  gather_sum_rows_kernel<<<GRID1D(n_out * cols)>>>(src, lds, idx, fan, n_out, cols, dst, ldd, accumulate);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_scale_rows(const float* src, int lds, const float* row_scale, long long rows, int cols, float* dst, int ldd, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols >= 0, "scale_rows: bad sizes");
  if (rows * cols == 0) return NLV_OK;
  NLV_CHECK_ARG(src && row_scale && dst, "scale_rows: null pointer");
  scale_rows_kernel<<<GRID1D(rows * cols)>>>(src, lds, row_scale, rows, cols, dst, ldd);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_scatter_add_rows(const float* src, int lds, const long long* idx, int idx_stride, long long n, int cols,
                         float* dst, int ldd, void* stream) {
  NLV_CHECK_ARG(n >= 0 && cols >= 0, "scatter_add_rows: bad sizes");
  if (n * cols == 0) return NLV_OK;
  NLV_CHECK_ARG(src && idx && dst, "scatter_add_rows: null pointer");
  scatter_add_rows_kernel<<<GRID1D(n * cols)>>>(src, lds, idx, idx_stride, n, cols, dst, ldd);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_assemble_tokens(const float* fo, const long long* pair_idx, const long long* labels, const float* e1,
                        const float* e2, long long r, float* rel, void* stream) {
  NLV_CHECK_ARG(r >= 0, "assemble_tokens: bad sizes");
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(fo && pair_idx && labels && e1 && e2 && rel, "assemble_tokens: null pointer");
  assemble_tokens_kernel<<<GRID1D(r * 1424)>>>(fo, pair_idx, labels, e1, e2, r, rel);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_assemble_tokens_bwd(const float* drel, const long long* pair_idx, const long long* labels, long long r,
                            float* dfo, float* de1, float* de2, void* stream) {
  NLV_CHECK_ARG(r >= 0, "assemble_tokens_bwd: bad sizes");
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(drel && pair_idx && labels && dfo && de1 && de2, "assemble_tokens_bwd: null pointer");
  assemble_tokens_bwd_kernel<<<GRID1D(r * 1424)>>>(drel, pair_idx, labels, r, dfo, de1, de2);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_center_size(const float* boxes, long long n, float* out, void* stream) {
  NLV_CHECK_ARG(n >= 0, "center_size: bad sizes");
  if (n == 0) return NLV_OK;
  center_size_kernel<<<GRID1D(n)>>>(boxes, n, out);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_colsum(const void* x, int x_dtype, int ld, long long rows, int cols, const int* row_class, int n_class,
               float* out, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols >= 0 && n_class >= 1, "colsum: bad sizes");
  if (rows == 0 || cols == 0) return NLV_OK;
  NLV_CHECK_ARG(x && out, "colsum: null pointer");
  if ((cols & 7) == 0 && (ld & 7) == 0 && al16(x)) return launch_colsum_v8(x, x_dtype, ld, rows, cols, row_class, n_class, out, STREAM);
  int splits = (int)((rows + 255) / 256);
  if (splits > 1024) splits = 1024;
  dim3 grid(cdiv(cols, 32), splits), block(32, 8);
  colsum_kernel<<<grid, block, 0, STREAM>>>(x, x_dtype, ld, rows, cols, row_class, n_class, out);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_relu_mask(const void* x, int x_dtype, int ldx, const void* gate, int gate_dtype, int ldg, long long rows,
                  int cols, void* y, int y_dtype, int ldy, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols >= 0, "relu_mask: bad sizes");
  if (rows * cols == 0) return NLV_OK;
  relu_mask_kernel<<<GRID1D(rows * cols)>>>(x, x_dtype, ldx, gate, gate_dtype, ldg, rows, cols, y, y_dtype, ldy);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_add(const float* a, const float* b, long long n, float* y, void* stream) {
  NLV_CHECK_ARG(n >= 0, "add: bad sizes");
  if (n == 0) return NLV_OK;
  add_kernel<<<GRID1D(n)>>>(a, b, n, y);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
}
