// Small helper kernels of the model sequencer (step.cu): multi-segment convert / copy, weight layout permutation,
// byte fill, the chain rule of the activated heads.
#include "common.cuh"

namespace nlv {
namespace {

constexpr int kMaxSeg = 32;
struct MultiSeg {
  const void* src[kMaxSeg];
  void* dst[kMaxSeg];
  long long start[kMaxSeg + 1];   // prefix sums of 8-element chunks
  long long n[kMaxSeg];
  int count;
};

// one thread per 8-element chunk; 16-byte accesses when both sides of the segment are aligned
__global__ void convert_multi_kernel(const MultiSeg s, int sdt, int ddt) {
  const long long total = s.start[s.count];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= s.start[k + 1]) ++k;
    const long long e0 = (i - s.start[k]) * 8;
    const long long left = s.n[k] - e0;
    const void* src = s.src[k];
    void* dst = s.dst[k];
    const int ssz = sdt == NLV_BF16 ? 2 : 4, dsz = ddt == NLV_BF16 ? 2 : 4;
    const bool al = (((uintptr_t)src + e0 * ssz) & 15) == 0 && (((uintptr_t)dst + e0 * dsz) & 15) == 0;
    if (left >= 8 && al) {
      float v[8];
      if (sdt == NLV_BF16) {
        const uint4 t = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(src) + e0);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float2 f2 = __bfloat1622float2(h[q]); v[2 * q] = f2.x; v[2 * q + 1] = f2.y; }
      } else {
        const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + e0);
        const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(src) + e0 + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      }
      if (ddt == NLV_BF16) {
        uint4 t;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
        t.x = *reinterpret_cast<uint32_t*>(&h0); t.y = *reinterpret_cast<uint32_t*>(&h1);
        t.z = *reinterpret_cast<uint32_t*>(&h2); t.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dst) + e0) = t;
      } else {
        float* o = reinterpret_cast<float*>(dst) + e0;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    } else {
      const int m = left < 8 ? (int)left : 8;
      for (int q = 0; q < m; ++q) st_from_float(dst, ddt, (size_t)(e0 + q), ld_as_float(src, sdt, (size_t)(e0 + q)));
    }
  }
}

// dst[a, c, b] = src[a, b, c] through a 32 x 33 shared tile
__global__ void permute_021_kernel(const void* __restrict__ src, int sdt, int Bn, int Cn, void* __restrict__ dst, int ddt) {
  __shared__ float tile[32][33];
  const int a = blockIdx.z;
  const int c0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const size_t base = (size_t)a * Bn * Cn;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int b = b0 + j, c = c0 + threadIdx.x;
    if (b < Bn && c < Cn) tile[j][threadIdx.x] = ld_as_float(src, sdt, base + (size_t)b * Cn + c);
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, b = b0 + threadIdx.x;
    if (b < Bn && c < Cn) st_from_float(dst, ddt, base + (size_t)c * Bn + b, tile[threadIdx.x][j]);
  }
}

__global__ void zero16_kernel(uint4* p, long long n16) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void zero1_kernel(unsigned char* p, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = 0;
}

__global__ void heads_activation_bwd_kernel(const float* __restrict__ datt, const float* __restrict__ dspa, const float* __restrict__ dcon,
                                            const float* __restrict__ spa, const float* __restrict__ con, long long r, float* __restrict__ d26) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r * 26) return;
  const long long row = i / 26;
  const int c = (int)(i - row * 26);
  float v = 0.f;
  if (c < 3) { if (datt) v = datt[row * 3 + c]; }
  else if (c < 9) { if (dspa) { const float s = spa[row * 6 + c - 3]; v = dspa[row * 6 + c - 3] * s * (1.f - s); } }
  else { if (dcon) { const float s = con[row * 17 + c - 9]; v = dcon[row * 17 + c - 9] * s * (1.f - s); } }
  d26[i] = v;
}

__global__ void set_seg_kernel(int* p, int rows) { p[0] = 0; p[1] = rows; }

}  // namespace

int set_seg(int* p, int rows, cudaStream_t s) {
  set_seg_kernel<<<1, 1, 0, s>>>(p, rows);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

}  // namespace nlv

using namespace nlv;
#define STREAM ((cudaStream_t)stream)

extern "C" {

int nlv_convert_multi(const void* const* src, void* const* dst, const long long* n, int count, int src_dtype, int dst_dtype, void* stream) {
  NLV_CHECK_ARG(count >= 0 && (count == 0 || (src && dst && n)), "convert_multi: bad arguments");
  for (int k0 = 0; k0 < count; k0 += kMaxSeg) {
    MultiSeg s;
    s.count = 0;
    s.start[0] = 0;
    for (int k = k0; k < count && s.count < kMaxSeg; ++k) {
      if (n[k] <= 0) continue;
      NLV_CHECK_ARG(src[k] && dst[k], "convert_multi: null segment %d", k);
      s.src[s.count] = src[k]; s.dst[s.count] = dst[k]; s.n[s.count] = n[k];
      s.start[s.count + 1] = s.start[s.count] + (n[k] + 7) / 8;
      ++s.count;
    }
    const long long total = s.start[s.count];
    if (total == 0) continue;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    convert_multi_kernel<<<(unsigned)blocks, 256, 0, STREAM>>>(s, src_dtype, dst_dtype);
    NLV_CHECK_LAUNCH();
  }
  return NLV_OK;
}

int nlv_permute_021(const void* src, int src_dtype, int a, int b, int c, void* dst, int dst_dtype, void* stream) {
  NLV_CHECK_ARG(a >= 0 && b >= 0 && c >= 0 && a <= 65535, "permute_021: bad sizes");
  if ((long long)a * b * c == 0) return NLV_OK;
  NLV_CHECK_ARG(src && dst, "permute_021: null pointer");
  dim3 grid(cdiv(c, 32), cdiv(b, 32), a), block(32, 8);
  permute_021_kernel<<<grid, block, 0, STREAM>>>(src, src_dtype, b, c, dst, dst_dtype);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_zero_bytes(void* p, long long nbytes, void* stream) {
  NLV_CHECK_ARG(nbytes >= 0, "zero_bytes: bad size");
  if (nbytes == 0) return NLV_OK;
  NLV_CHECK_ARG(p != nullptr, "zero_bytes: null pointer");
  unsigned char* q = reinterpret_cast<unsigned char*>(p);
  const long long head = (16 - (reinterpret_cast<uintptr_t>(q) & 15)) & 15;
  const long long h = head < nbytes ? head : nbytes;
  if (h > 0) { zero1_kernel<<<1, 32, 0, STREAM>>>(q, h); NLV_CHECK_LAUNCH(); }
  const long long n16 = (nbytes - h) / 16;
  if (n16 > 0) {
    long long blocks = (n16 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    zero16_kernel<<<(unsigned)blocks, 256, 0, STREAM>>>(reinterpret_cast<uint4*>(q + h), n16);
    NLV_CHECK_LAUNCH();
  }
  const long long tail = nbytes - h - n16 * 16;
  if (tail > 0) { zero1_kernel<<<1, 32, 0, STREAM>>>(q + h + n16 * 16, tail); NLV_CHECK_LAUNCH(); }
  return NLV_OK;
}

int nlv_heads_activation_bwd(const float* datt, const float* dspa, const float* dcon, const float* spa, const float* con, long long r,
                             float* d26, void* stream) {
  NLV_CHECK_ARG(r >= 0, "heads_activation_bwd: bad size");
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(d26 && (!dspa || spa) && (!dcon || con), "heads_activation_bwd: null pointer");
  heads_activation_bwd_kernel<<<cdiv(r * 26, 256), 256, 0, STREAM>>>(datt, dspa, dcon, spa, con, r, d26);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
}
