// Small helper kernels of the model sequencer (step.cu): multi-segment convert / copy, weight layout permutation,
// byte fill, the chain rule of the activated heads.
#include "common.cuh"
#include "philox.cuh"

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

// dst = keep ? src * scale : 0 ; one thread per group of 8 columns (one Philox call)
__global__ void dropout_apply_kernel(const void* __restrict__ src, int sdt, int lds, void* dst, int ddt, int ldd, long long rows, int cols,
                                     const DropCfg drop) {
  const int gpr = (cols + 7) >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * gpr) return;
  const long long r = i / gpr;
  const int g = (int)(i - r * gpr);
  const uint32_t keep = keep8_matrix(drop, r, g, gpr);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = g * 8 + q;
    if (c < cols) {
      const float v = ld_as_float(src, sdt, (size_t)r * lds + c);
      st_from_float(dst, ddt, (size_t)r * ldd + c, ((keep >> q) & 1u) ? v * drop.scale : 0.f);
    }
  }
}
__global__ void dropout_mask_kernel(long long rows, int cols, const DropCfg drop, unsigned char* __restrict__ out) {
  const int gpr = (cols + 7) >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * gpr) return;
  const long long r = i / gpr;
  const int g = (int)(i - r * gpr);
  const uint32_t keep = keep8_matrix(drop, r, g, gpr);
  for (int q = 0; q < 8; ++q)
    if (g * 8 + q < cols) out[(size_t)r * cols + g * 8 + q] = (keep >> q) & 1u;
}
__global__ void dropout_mask_attn_kernel(long long rows, int heads, int nkeys, const DropCfg drop, unsigned char* __restrict__ out) {
  const int kgs = (nkeys + 7) >> 3;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * heads * kgs) return;
  const int kg = (int)(i % kgs);
  const long long rh = i / kgs;
  const int h = (int)(rh % heads);
  const long long r = rh / heads;
  const uint32_t keep = keep8_attn(drop, r, heads, h, kg);
  for (int q = 0; q < 8; ++q)
    if (kg * 8 + q < nkeys) out[((size_t)r * heads + h) * nkeys + kg * 8 + q] = (keep >> q) & 1u;
}

// Zero-suppressed union features (packed feature files, nlvsgg_b200/featfile.py) -> dense bf16 rows [rows, 2048].
// One warp per row.  The row's stored values (channel order, vals[off[row] .. off[row+1])) are staged in shared memory with
// 16-byte loads (from the 16-byte boundary below off[row]; the value array is padded).  Lane l holds occupancy word l
// (channels 64 l .. 64 l + 63); a warp scan of the word popcounts gives every word's first value.  Output is written in
// 16-byte chunks of 8 channels, chunk j = 32 g + lane in pass g, so every store instruction covers 512 contiguous bytes;
// channel i of a chunk is the popc(occupancy bits below it)-th stored value — independent predicated loads, no serial chain.
__global__ void __launch_bounds__(256)
union_unpack_kernel(const unsigned long long* __restrict__ bitmap, const unsigned* __restrict__ off,
                    const unsigned short* __restrict__ vals, long long rows, unsigned short* __restrict__ dst) {
  __shared__ __align__(16) unsigned short sv[8][2048 + 8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const unsigned long long bits = bitmap[r * 32 + lane];
    const size_t b0 = off[r];
    const int total = (int)(off[r + 1] - off[r]);
    const size_t a0 = b0 & ~(size_t)7;
    const int mis = (int)(b0 - a0);
    const uint4* src = reinterpret_cast<const uint4*>(vals + a0);
    uint4* stg = reinterpret_cast<uint4*>(sv[w]);
    for (int c = lane; c * 8 < mis + total; c += 32) stg[c] = src[c];
    const int cnt = __popcll(bits);
    int pre = cnt;                                  // inclusive prefix sum over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const int excl = pre - cnt;                     // first stored value of word `lane`
    __syncwarp();
    const unsigned short* v = sv[w] + mis;
    uint4* out = reinterpret_cast<uint4*>(dst + (size_t)r * 2048);
    const int byte = lane & 7;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int word = (lane >> 3) + 4 * g;         // chunk j = 32 g + lane lies in word j / 8, byte j % 8
      const unsigned long long wb = __shfl_sync(0xffffffffu, bits, word);
      const int base = __shfl_sync(0xffffffffu, excl, word) + __popcll(wb & ((1ull << (8 * byte)) - 1ull));
      const unsigned m = (unsigned)(wb >> (8 * byte)) & 0xffu;
      unsigned wd[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const int i = 2 * h;
        const int i0 = base + __popc(m & ((1u << i) - 1u));
        const bool s0 = (m >> i) & 1u, s1 = (m >> (i + 1)) & 1u;
        const unsigned lo = s0 ? v[i0] : 0u;
        const unsigned hi = s1 ? v[i0 + (s0 ? 1 : 0)] : 0u;
        wd[h] = lo | (hi << 16);
      }
      out[32 * g + lane] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
    __syncwarp();
  }
}

// The same with 12-bit stored values (featfile.py "sparse12"): value i of the batch = (base[row] + code_i) << 8 | lo_i, lo = one byte per
// value, code = 4 bits per value (value i in nibble i & 1 of byte i >> 1).  A quarter less to move over PCIe; the decode stages
// the row's two planes in shared memory with 16-byte aligned-superset copies.
__global__ void __launch_bounds__(256)
union_unpack12_kernel(const unsigned long long* __restrict__ bitmap, const unsigned* __restrict__ off, const unsigned char* __restrict__ lo,
                      const unsigned char* __restrict__ hx, const unsigned char* __restrict__ base, long long rows,
                      unsigned short* __restrict__ dst) {
  __shared__ __align__(16) unsigned char slo[8][2048 + 32];
  __shared__ __align__(16) unsigned char shx[8][1024 + 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += nwarps) {
    const unsigned long long bits = bitmap[r * 32 + lane];
    const size_t b0 = off[r];
    const int total = (int)(off[r + 1] - off[r]);
    const unsigned bb = base[r];
    const size_t a0 = b0 & ~(size_t)15;              // low-byte plane: aligned superset of [b0, b0 + total)
    const int mis = (int)(b0 - a0);
    const size_t h0 = b0 >> 1, ha0 = h0 & ~(size_t)15;   // code plane: bytes [b0 / 2, (b0 + total + 1) / 2)
    const int hmis = (int)(h0 - ha0), hbytes = (int)(((b0 + total + 1) >> 1) - ha0);
    {
      const uint4* src = reinterpret_cast<const uint4*>(lo + a0);
      uint4* stg = reinterpret_cast<uint4*>(slo[w]);
      for (int c = lane; c * 16 < mis + total; c += 32) stg[c] = src[c];
      const uint4* srch = reinterpret_cast<const uint4*>(hx + ha0);
      uint4* stgh = reinterpret_cast<uint4*>(shx[w]);
      for (int c = lane; c * 16 < hbytes; c += 32) stgh[c] = srch[c];
    }
    const int cnt = __popcll(bits);
    int pre = cnt;                                  // inclusive prefix sum over the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const int excl = pre - cnt;                     // first stored value of word `lane`
    __syncwarp();
    // a chunk's stored values are consecutive in the stream: its (at most) 8 low bytes and 8 codes are fetched as aligned words and
    // funnel-shifted into place — 5 word loads per chunk instead of two byte loads per value — then consumed from the low end
    const uint32_t* wl = reinterpret_cast<const uint32_t*>(slo[w]);
    const uint32_t* wh = reinterpret_cast<const uint32_t*>(shx[w]);
    const unsigned par = (unsigned)(b0 & 1);
    uint4* out = reinterpret_cast<uint4*>(dst + (size_t)r * 2048);
    const int byte = lane & 7;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int word = (lane >> 3) + 4 * g;         // chunk j = 32 g + lane lies in word j / 8, byte j % 8
      const unsigned long long wb = __shfl_sync(0xffffffffu, bits, word);
      const int first = __shfl_sync(0xffffffffu, excl, word) + __popcll(wb & ((1ull << (8 * byte)) - 1ull));
      const unsigned m = (unsigned)(wb >> (8 * byte)) & 0xffu;
      const unsigned bo = (unsigned)(mis + first);                  // byte offset of the chunk's first low byte
      const unsigned no = 2u * (unsigned)hmis + par + (unsigned)first;   // nibble offset of its first code
      const uint32_t a0 = wl[bo >> 2], a1 = wl[(bo >> 2) + 1], a2 = wl[(bo >> 2) + 2];
      const uint32_t c0 = wh[no >> 3], c1 = wh[(no >> 3) + 1];
      uint32_t l0 = __funnelshift_r(a0, a1, 8u * (bo & 3u)), l1 = __funnelshift_r(a1, a2, 8u * (bo & 3u));
      uint32_t nib = __funnelshift_r(c0, c1, 4u * (no & 7u));
      unsigned wd[4];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        unsigned v = 0u;
        if ((m >> c) & 1u) {
          v = ((bb + (nib & 15u)) << 8) | (l0 & 0xffu);
          nib >>= 4;
          l0 = __funnelshift_r(l0, l1, 8);
          l1 >>= 8;
        }
        if (c & 1) wd[c >> 1] |= v << 16; else wd[c >> 1] = v;
      }
      out[32 * g + lane] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
    __syncwarp();
  }
}

__global__ void union_patch_kernel(unsigned short* __restrict__ dst, const unsigned* __restrict__ pos, const unsigned short* __restrict__ val, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[pos[i]] = val[i];
}

// lib/assign_pseudo_label.py:934-938 create_dis: d = zeros(36); d[idx] = conf; d[d == 0] = (1 - conf) / 35
// `other`: the value of the 35 remaining entries as the producer computed it (python double or fp32 tensor arithmetic,
// depending on the caller); NULL -> (1 - conf) / 35 in fp32
__global__ void create_dis_kernel(const float* __restrict__ conf, const float* __restrict__ other, const int* __restrict__ idx,
                                  long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 36) return;
  const long long r = i / 36;
  const int c = (int)(i - r * 36);
  const float cf = conf[r];
  float v = (c == idx[r]) ? cf : 0.f;
  if (v == 0.f) v = other != nullptr ? other[r] : __fdiv_rn(__fsub_rn(1.f, cf), 35.f);
  out[i] = v;
}

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

int nlv_union_unpack(const void* bitmap, const unsigned* off, const void* vals, long long rows, void* dst_bf16, void* stream) {
  NLV_CHECK_ARG(rows >= 0, "union_unpack: bad size");
  if (rows == 0) return NLV_OK;
  NLV_CHECK_ARG(bitmap && off && vals && dst_bf16, "union_unpack: null pointer");
  NLV_CHECK_ARG((reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(bitmap) & 7) == 0, "union_unpack: misaligned buffer");
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  union_unpack_kernel<<<(unsigned)blocks, 256, 0, STREAM>>>(reinterpret_cast<const unsigned long long*>(bitmap), off,
                                                           reinterpret_cast<const unsigned short*>(vals), rows,
                                                           reinterpret_cast<unsigned short*>(dst_bf16));
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_union_unpack12(const void* bitmap, const unsigned* off, const void* lo, const void* hx, const unsigned char* base, long long rows,
                       void* dst_bf16, void* stream) {
  NLV_CHECK_ARG(rows >= 0, "union_unpack12: bad size");
  if (rows == 0) return NLV_OK;
  NLV_CHECK_ARG(bitmap && off && lo && hx && base && dst_bf16, "union_unpack12: null pointer");
  NLV_CHECK_ARG((reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(bitmap) & 7) == 0 &&
                (reinterpret_cast<uintptr_t>(lo) & 15) == 0 && (reinterpret_cast<uintptr_t>(hx) & 15) == 0, "union_unpack12: misaligned buffer");
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  union_unpack12_kernel<<<(unsigned)blocks, 256, 0, STREAM>>>(reinterpret_cast<const unsigned long long*>(bitmap), off,
                                                             reinterpret_cast<const unsigned char*>(lo), reinterpret_cast<const unsigned char*>(hx),
                                                             base, rows, reinterpret_cast<unsigned short*>(dst_bf16));
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_union_patch(void* dst_bf16, const unsigned* pos, const unsigned short* val, int n, void* stream) {
  NLV_CHECK_ARG(n >= 0, "union_patch: bad size");
  if (n == 0) return NLV_OK;
  NLV_CHECK_ARG(dst_bf16 && pos && val, "union_patch: null pointer");
  union_patch_kernel<<<cdiv(n, 256), 256, 0, STREAM>>>(reinterpret_cast<unsigned short*>(dst_bf16), pos, val, n);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_create_dis(const float* conf, const float* other, const int* idx, long long n, float* out, void* stream) {
  NLV_CHECK_ARG(n >= 0, "create_dis: bad size");
  if (n == 0) return NLV_OK;
  NLV_CHECK_ARG(conf && idx && out, "create_dis: null pointer");
  create_dis_kernel<<<cdiv(n * 36, 256), 256, 0, STREAM>>>(conf, other, idx, n, out);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

static DropCfg to_cfg(const nlv_dropout* d) {
  DropCfg c = drop_off();
  if (d != nullptr && d->thr16 != 0u) { c.thr16 = d->thr16; c.scale = d->scale; c.seed_lo = d->seed_lo; c.seed_hi = d->seed_hi; c.stream = d->stream; }
  return c;
}

int nlv_dropout_apply(const void* src, int src_dtype, int lds, void* dst, int dst_dtype, int ldd, long long rows, int cols,
                      const nlv_dropout* drop, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols >= 0, "dropout_apply: bad sizes");
  if (rows * cols == 0) return NLV_OK;
  NLV_CHECK_ARG(src && dst, "dropout_apply: null pointer");
  const DropCfg c = to_cfg(drop);
  if (c.thr16 == 0u) return nlv_convert(src, src_dtype, lds, dst, dst_dtype, ldd, rows, cols, stream);
  const long long n = rows * ((cols + 7) / 8);
  dropout_apply_kernel<<<cdiv(n, 256), 256, 0, STREAM>>>(src, src_dtype, lds, dst, dst_dtype, ldd, rows, cols, c);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_dropout_mask(long long rows, int cols, const nlv_dropout* drop, unsigned char* out, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && cols >= 0 && out, "dropout_mask: bad arguments");
  if (rows * cols == 0) return NLV_OK;
  const long long n = rows * ((cols + 7) / 8);
  dropout_mask_kernel<<<cdiv(n, 256), 256, 0, STREAM>>>(rows, cols, to_cfg(drop), out);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_dropout_mask_attn(long long rows, int heads, int nkeys, const nlv_dropout* drop, unsigned char* out, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && heads > 0 && nkeys > 0 && out, "dropout_mask_attn: bad arguments");
  if (rows == 0) return NLV_OK;
  const long long n = rows * heads * ((nkeys + 7) / 8);
  dropout_mask_attn_kernel<<<cdiv(n, 256), 256, 0, STREAM>>>(rows, heads, nkeys, to_cfg(drop), out);
  NLV_CHECK_LAUNCH();
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
