// First stage of the spatial-mask branch, lib/sttran.py:337-341 (Conv2d(2,128,k7,s2,p3) -> ReLU -> BatchNorm2d(128) ->
// MaxPool2d(3,2,1) over the 2x27x27 box masks of a pair), bf16 path, without the im2col matrix.
//
// The generic route (im2col -> tcgen05 GEMM [R*196 x 104] x [104 x 128] -> statistics pass -> BN pass -> pooling pass) moves a
// 483 MB column matrix and three 595 MB maps for 62 GFLOP: the GEMM ran at 0.1 of the tensor peak, everything else at
// DRAM speed.  Here a pair's two masks (5.8 KB) sit zero-padded in shared memory and the products run on mma.sync.m16n8k16:
//
//   nlv_mask_conv1_fwd     A fragments are 4-byte shared-memory loads straight from the padded mask: with the reduction index
//                          ordered (c, ky, kx padded to 8) a k-step of 16 is two mask rows, a thread's element pair (k = 2t, 2t+1)
//                          two adjacent mask pixels.  The weight fragments live in registers for the whole kernel.  Epilogue:
//                          bias + ReLU, bf16 store of the activation map (BatchNorm's input, kept for the backward pass) and the
//                          per-(video, channel) sum / sum of squares of the ROUNDED values — the statistics pass is gone.
//   nlv_bn_apply_maxpool   BatchNorm apply + 3x3/2 max pooling in one pass: reads the activation map once, writes the 7x7 map and
//                          the argmax taps (the values of nlv_bn_apply followed by nlv_maxpool_fwd; the window scan runs on
//                          the activations — BatchNorm is monotone per channel — and BatchNorm once per output).
//   nlv_pool_bn_bwd        backward of MaxPool + BatchNorm + ReLU in one sweep: the dense 14x14 gradient map is never written — a
//                          position gathers the pooled gradients of the (at most four) windows whose argmax it is; the reduction
//                          pass runs over the 7x7 cells only (pooled gradient x the activation at the argmax, kept by the forward).
//   nlv_mask_conv1_dw      weight gradient dW[co, (c,ky,kx)] = sum over pairs and positions of dY[pos, co] * mask patch: dY tiles
//                          arrive by cp.async (double-buffered), the transposed operand by ldmatrix.trans, the patch operand again
//                          straight from (de-interleaved copies of) the padded mask; per-CTA partial sums, one reduction kernel.
#include "common.cuh"

namespace nlv {
int launch_bn_finalize(const double* sums, const int* seg, int nseg, int c, float momentum, float* mean, float* var, float* running_mean,
                       float* running_var, cudaStream_t s);

int launch_bn_sums_bwd_v8(const void* dy, int dydt, int lddy, const void* x, int xdt, int ldx, const void* yout, int ydt, int ldy, const int* seg,
                          int nseg, const float* mean, const float* var, float eps, long long rows, int C, double* sums, cudaStream_t s);
int launch_bn_bwd_param(const double* sums, int nseg, int c, float* dw, float* db, cudaStream_t s);

namespace {

typedef __nv_bfloat16 bf16;

constexpr int MR = 33;             // padded mask rows:   y' = 2 oy + ky       in [0, 32]
constexpr int MP = 46;             // padded mask pitch:  x' = 2 ox + kx       in [0, 33] (kx = 7 is the zero-weight pad column); 23 words:
                                   // two mask rows = 46 words = 14 (mod 32), so the 8 consecutive positions of an A-fragment load
                                   // (word = 46 oy + ox + t) fall on consecutive banks also across the end of an output row
constexpr int MPW = MP / 2;        // pitch in 32-bit words
constexpr int MPLANE_W = MR * MPW; // words per mask channel
constexpr int MASK_W = 2 * MPLANE_W;
constexpr int NPIX = 2 * 27 * 27;  // mask elements of one pair
constexpr int PF = (NPIX + 255) / 256;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v));
}

// ---- bulk-copy pipeline helpers (cp.async.bulk = TMA 1-D: one instruction moves a pair's whole contiguous tile) ----
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// bounded wait (a protocol bug must trap, not hang the GPU)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}

// a[j] of lane t (t = lane & 3) <- a[t] of lane j of the same quad
__device__ __forceinline__ void quad_transpose(uint32_t (&a)[4], int t) {
  const bool lo = (t & 1) == 0, hi = (t & 2) == 0;
  uint32_t r0 = __shfl_xor_sync(0xffffffffu, lo ? a[1] : a[0], 1), r1 = __shfl_xor_sync(0xffffffffu, lo ? a[3] : a[2], 1);
  if (lo) { a[1] = r0; a[3] = r1; } else { a[0] = r0; a[2] = r1; }
  r0 = __shfl_xor_sync(0xffffffffu, hi ? a[2] : a[0], 2); r1 = __shfl_xor_sync(0xffffffffu, hi ? a[3] : a[1], 2);
  if (hi) { a[2] = r0; a[3] = r1; } else { a[0] = r0; a[1] = r1; }
}

// mask pixel (c, y, x) of the fp32 source -> bf16 slot (c, y + 3, x + 3) of the padded tile
__device__ __forceinline__ void prefetch_mask(const float* __restrict__ m, long long pair, float (&v)[PF]) {
  const float* src = m + pair * NPIX;
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const int e = threadIdx.x + i * 256;
    v[i] = e < NPIX ? src[e] : 0.f;
  }
}
__device__ __forceinline__ void park_mask(bf16* tile, const float (&v)[PF]) {
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const int e = threadIdx.x + i * 256;
    if (e < NPIX) {
      const int c = e / 729, rem = e - c * 729, y = rem / 27, x = rem - y * 27;
      tile[(c * MR + y + 3) * MP + x + 3] = __float2bfloat16_rn(v[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// forward: CTA = 8 warps; warp = 32 output channels (4 n-tiles) x half of the 13 position tiles of a pair
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
mask_conv1_fwd_kernel(const float* __restrict__ masks, const float* __restrict__ w /*[128,98] (c,ky,kx)*/, const float* __restrict__ bias,
                      const int* __restrict__ pair_video, long long R, bf16* __restrict__ out /*[R*196,128]*/,
                      double* __restrict__ sums /*[nv,2,128]*/) {
  __shared__ __align__(16) uint32_t msw[2][MASK_W];
  __shared__ double st[8][2][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int cg = warp & 3, mh = warp >> 2;
  const long long p0 = R * blockIdx.x / gridDim.x, p1 = R * (blockIdx.x + 1) / gridDim.x;
  if (p0 >= p1) return;

  for (int i = threadIdx.x; i < 2 * MASK_W; i += 256) (&msw[0][0])[i] = 0u;
  for (int i = threadIdx.x; i < 8 * 2 * 32; i += 256) (&st[0][0][0])[i] = 0.0;

  // weight fragments: breg[j][s] = {W'(co, k' = 16 s + 2t, +1), W'(co, k' = 16 s + 8 + 2t, +1)}, co = cg*32 + j*8 + g,
  // k' = (c*7 + ky) * 8 + kx
  uint32_t breg[4][7][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float* wr = w + (size_t)(cg * 32 + j * 8 + g) * 98;
#pragma unroll
    for (int s = 0; s < 7; ++s)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * s + hh;                    // = c*7 + ky
        const int k0 = h * 7 + 2 * t;                // (c*49 + ky*7) + kx, kx = 2t
        const float lo = wr[k0];
        const float hi = (2 * t + 1 < 7) ? wr[k0 + 1] : 0.f;
        breg[j][s][hh] = pack2(lo, hi);
      }
  }
  float bs[4][2];
#pragma unroll
  for (int j = 0; j < 4; ++j) { bs[j][0] = bias[cg * 32 + j * 8 + 2 * t]; bs[j][1] = bias[cg * 32 + j * 8 + 2 * t + 1]; }

  float pf[PF];
  prefetch_mask(masks, p0, pf);
  __syncthreads();                                   // zero fill done
  park_mask(reinterpret_cast<bf16*>(msw[0]), pf);
  __syncthreads();

  int cur = 0, cur_vid = -1;
  const int mt0 = mh * 7, mt1 = mh ? 13 : 7;
  for (long long p = p0; p < p1; ++p) {
    const int vid = pair_video ? pair_video[p] : 0;
    if (vid != cur_vid) {
      if (cur_vid >= 0) {                            // block-uniform
        const int c = threadIdx.x & 127, k = threadIdx.x >> 7, q = c >> 5;
        const double v = st[q][k][c & 31] + st[q + 4][k][c & 31];
        if (v != 0.0) atomicAdd(sums + ((size_t)cur_vid * 2 + k) * 128 + c, v);
        __syncthreads();
        for (int i = threadIdx.x; i < 8 * 2 * 32; i += 256) (&st[0][0][0])[i] = 0.0;
        __syncthreads();
      }
      cur_vid = vid;
    }
    if (p + 1 < p1) prefetch_mask(masks, p + 1, pf);
    const uint32_t* mw = msw[cur];
    float ssum[4][2], ssq[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { ssum[j][0] = ssum[j][1] = ssq[j][0] = ssq[j][1] = 0.f; }
    for (int mt = mt0; mt < mt1; ++mt) {
      const int pos0 = mt * 16 + g, pos1 = pos0 + 8;
      const int q0 = min(pos0, 195), q1 = min(pos1, 195);
      const int oy0 = q0 / 14, ox0 = q0 - oy0 * 14, oy1 = q1 / 14, ox1 = q1 - oy1 * 14;
      const uint32_t* r0 = mw + (2 * oy0) * MPW + ox0 + t;
      const uint32_t* r1 = mw + (2 * oy1) * MPW + ox1 + t;
      float acc[4][4];                                 // accumulators start at the bias
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc[j][0] = acc[j][2] = bs[j][0]; acc[j][1] = acc[j][3] = bs[j][1]; }
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const int h0 = 2 * s, h1 = 2 * s + 1;
        const int o0 = ((h0 / 7) * MR + (h0 % 7)) * MPW, o1 = ((h1 / 7) * MR + (h1 % 7)) * MPW;
        const uint32_t a[4] = {r0[o0], r1[o0], r0[o1], r1[o1]};
#pragma unroll
        for (int j = 0; j < 4; ++j) mma16816(acc[j], a, breg[j][s][0], breg[j][s][1]);
      }
      const bool ok0 = pos0 < 196, ok1 = pos1 < 196;
      uint32_t u0[4], u1[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        u0[j] = ok0 ? pack2(fmaxf(acc[j][0], 0.f), fmaxf(acc[j][1], 0.f)) : 0u;
        u1[j] = ok1 ? pack2(fmaxf(acc[j][2], 0.f), fmaxf(acc[j][3], 0.f)) : 0u;
        const float2 f0 = unpack2(u0[j]), f1 = unpack2(u1[j]);       // statistics of the stored (rounded) values; 0 for padding rows
        ssum[j][0] += f0.x + f1.x; ssum[j][1] += f0.y + f1.y;
        ssq[j][0] = fmaf(f0.x, f0.x, fmaf(f1.x, f1.x, ssq[j][0])); ssq[j][1] = fmaf(f0.y, f0.y, fmaf(f1.y, f1.y, ssq[j][1]));
      }
      // 4 x 4 transposition inside the quad: lane t ends up with n-tile t of all four lanes = 8 consecutive channels (16 bytes)
      quad_transpose(u0, t);
      quad_transpose(u1, t);
      if (ok0) *reinterpret_cast<uint4*>(out + ((size_t)p * 196 + q0) * 128 + cg * 32 + t * 8) = make_uint4(u0[0], u0[1], u0[2], u0[3]);
      if (ok1) *reinterpret_cast<uint4*>(out + ((size_t)p * 196 + q1) * 128 + cg * 32 + t * 8) = make_uint4(u1[0], u1[1], u1[2], u1[3]);
    }
    // this pair's column sums: over the 8 row groups of the warp, then into the warp's double slots
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float a = ssum[j][e], b = ssq[j][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        if (g == 0) { st[warp][0][j * 8 + 2 * t + e] += (double)a; st[warp][1][j * 8 + 2 * t + e] += (double)b; }
      }
    if (p + 1 < p1) park_mask(reinterpret_cast<bf16*>(msw[cur ^ 1]), pf);
    __syncthreads();
    cur ^= 1;
  }
  {
    const int c = threadIdx.x & 127, k = threadIdx.x >> 7, q = c >> 5;
    const double v = st[q][k][c & 31] + st[q + 4][k][c & 31];
    if (v != 0.0) atomicAdd(sums + ((size_t)cur_vid * 2 + k) * 128 + c, v);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// BatchNorm apply + MaxPool2d(3, 2, 1): block = one pair, thread = (output position, 8 channels)
// ------------------------------------------------------------------------------------------------------------------
constexpr int XT_BYTES = 196 * 128 * 2;          // a pair's activation map: 50176 contiguous bytes
constexpr int F2_SMEM = 2 * XT_BYTES + 64;      // two stages + four mbarriers

// Persistent blocks; the activation maps arrive by bulk copy, double-buffered: the next pair's 50 KB are in flight while this
// one is pooled out of shared memory (the kernel is a pure stream: 595 MB in, 340 MB out).
__global__ void __launch_bounds__(256, 2)
bn_apply_maxpool_kernel(const bf16* __restrict__ x /*[R*196,128]*/, const int* __restrict__ pair_video, const float* __restrict__ mean,
                        const float* __restrict__ var, const float* __restrict__ w, const float* __restrict__ b, float eps, long long R,
                        bf16* __restrict__ y /*[R*49,128]*/, uint8_t* __restrict__ arg, bf16* __restrict__ xmax) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t buf0 = smem_addr(smem), bar0 = buf0 + 2 * XT_BYTES;
  const int c8 = threadIdx.x & 15, c0 = c8 * 8;
  // full[s] (bar0 + 8 s): the stage's bytes have landed; empty[s] (bar0 + 16 + 8 s): all 8 warps are done reading it.  No block-wide
  // barrier in the loop: warps drift, thread 0 refills a stage one pair ahead once its readers have arrived.
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(bar0 + 16, 8); mbar_init(bar0 + 24, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long first = blockIdx.x, stride = gridDim.x;
  if (threadIdx.x == 0 && first < R) { mbar_expect_tx(bar0, XT_BYTES); bulk_load(buf0, x + first * (196 * 128), XT_BYTES, bar0); }
  float m[8], k[8], bb[8];
  uint32_t sgn[4] = {0u, 0u, 0u, 0u};           // sign bits of the scales, on the packed bf16 pairs
  int cur_vid = -1, it = 0;
  for (long long p = first; p < R; p += stride, ++it) {
    const int s = it & 1;
    const int vid = pair_video ? pair_video[p] : 0;
    if (vid != cur_vid) {
      cur_vid = vid;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        m[q] = mean[(size_t)vid * 128 + c0 + q];
        k[q] = rsqrtf(var[(size_t)vid * 128 + c0 + q] + eps) * w[c0 + q];
        bb[q] = b[c0 + q];
      }
#pragma unroll
      for (int h = 0; h < 4; ++h) sgn[h] = (k[2 * h] < 0.f ? 0x00008000u : 0u) | (k[2 * h + 1] < 0.f ? 0x80000000u : 0u);
    }
    if (threadIdx.x == 0 && p + stride < R) {        // refill the other stage for the next pair
      if (it >= 1) mbar_wait(bar0 + 16 + 8 * (s ^ 1), ((it - 1) >> 1) & 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar0 + 8 * (s ^ 1), XT_BYTES);
      bulk_load(buf0 + (s ^ 1) * XT_BYTES, x + (p + stride) * (196 * 128), XT_BYTES, bar0 + 8 * (s ^ 1));
    }
    mbar_wait(bar0 + 8 * s, (it >> 1) & 1);
    const uint4* tile = reinterpret_cast<const uint4*>(smem + s * XT_BYTES) + c8;
    for (int pos = threadIdx.x >> 4; pos < 49; pos += 16) {
      const int oy = pos / 7, ox = pos - oy * 7;
      // BatchNorm (rounded to bf16, as the separate pass stores it) is monotone per channel: increasing for a positive scale,
      // decreasing for a negative one.  So the window maximum of its output is its value at the maximum (minimum) ACTIVATION:
      // the scan compares activations with the sign of the scale folded in, and BatchNorm runs once per output.  (Taps whose
      // outputs only tie after rounding: the larger activation wins — the fp32 model's argmax.)
      float best[8];
      int bi[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { best[q] = -INFINITY; bi[q] = 0; }
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int iy = oy * 2 - 1 + tap / 3, ix = ox * 2 - 1 + tap % 3;
        if (iy >= 0 && iy < 14 && ix >= 0 && ix < 14) {
          const uint4 raw = tile[(iy * 14 + ix) * 16];
          const uint32_t* rw = reinterpret_cast<const uint32_t*>(&raw);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 f = unpack2(rw[h] ^ sgn[h]);
            if (f.x > best[2 * h] || (f.x != f.x)) { best[2 * h] = f.x; bi[2 * h] = tap; }
            if (f.y > best[2 * h + 1] || (f.y != f.y)) { best[2 * h + 1] = f.y; bi[2 * h + 1] = tap; }
          }
        }
      }
      uint4 o, xm;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
      uint32_t* xw = reinterpret_cast<uint32_t*>(&xm);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        xw[h] = pack2(best[2 * h], best[2 * h + 1]) ^ sgn[h];          // exact: the keys are bf16 values
        const float2 f = unpack2(xw[h]);
        ow[h] = pack2(fmaf(f.x - m[2 * h], k[2 * h], bb[2 * h]), fmaf(f.y - m[2 * h + 1], k[2 * h + 1], bb[2 * h + 1]));
      }
      const size_t e = ((size_t)p * 49 + pos) * 16 + c8;
      reinterpret_cast<uint4*>(y)[e] = o;
      uint2 pk;
      pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      reinterpret_cast<uint2*>(arg)[e] = pk;
      if (xmax != nullptr) reinterpret_cast<uint4*>(xmax)[e] = xm;
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar0 + 16 + 8 * s);     // this warp is done reading stage s
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward of MaxPool2d(3,2,1) + BatchNorm (+ the ReLU in front of it): dx[pair, iy, ix, c] from the pooled gradient
// ------------------------------------------------------------------------------------------------------------------
constexpr int DP_BYTES = 49 * 128 * 4, ARG_BYTES = 49 * 128;
constexpr int BW_STAGE = XT_BYTES + DP_BYTES + ARG_BYTES;        // 81536 bytes per pair: activation map, pooled gradient, taps
constexpr int BW_SMEM = 2 * BW_STAGE + 64 + 4 * 128 * 4 + 16 * 128 * 4;

// One persistent block of 16 warps per SM; a pair's three inputs arrive by bulk copy, double-buffered.
__global__ void __launch_bounds__(512, 1)
pool_bn_bwd_apply_kernel(const float* __restrict__ dp /*[R*49,128]*/, const uint8_t* __restrict__ arg, const bf16* __restrict__ x /*[R*196,128]*/,
                         const int* __restrict__ pair_video, const int* __restrict__ seg196, const float* __restrict__ mean,
                         const float* __restrict__ var, const float* __restrict__ w, float eps, const double* __restrict__ sums,
                         int use_batch_stats, long long R, bf16* __restrict__ dx, float* __restrict__ dx_colsum) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t buf0 = smem_addr(smem), bar0 = buf0 + 2 * BW_STAGE;
  float (*prm)[128] = reinterpret_cast<float (*)[128]>(smem + 2 * BW_STAGE + 64);   // mean, w * rstd, sum_dy / n, rstd * sum_dy_xhat / n
  float (*red)[128] = reinterpret_cast<float (*)[128]>(smem + 2 * BW_STAGE + 64 + 4 * 128 * 4);
  const int c8 = threadIdx.x & 15, c0 = c8 * 8, lane_pos = threadIdx.x >> 4;        // 32 positions per sweep
  const int hsel = (c8 >> 2) & 1;
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); mbar_init(bar0 + 16, 16); mbar_init(bar0 + 24, 16);   // full[2], empty[2] (16 warps)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long first = blockIdx.x, stride = gridDim.x;
  auto issue = [&](int s, long long p) {
    const uint32_t d = buf0 + s * BW_STAGE, bar = bar0 + 8 * s;
    mbar_expect_tx(bar, BW_STAGE);
    bulk_load(d, x + p * (196 * 128), XT_BYTES, bar);
    bulk_load(d + XT_BYTES, dp + p * (49 * 128), DP_BYTES, bar);
    bulk_load(d + XT_BYTES + DP_BYTES, arg + p * (49 * 128), ARG_BYTES, bar);
  };
  if (threadIdx.x == 0 && first < R) issue(0, first);
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float m[8], k[8], s1[8], t2[8];
  int cur = -1, it = 0;
  for (long long p = first; p < R; p += stride, ++it) {
    const int s = it & 1;
    const int vid = pair_video ? pair_video[p] : 0;
    if (vid != cur) {                                // block-uniform
      __syncthreads();
      cur = vid;
      if (threadIdx.x < 128) {
        const int c = threadIdx.x;
        const float rs = rsqrtf(var[(size_t)vid * 128 + c] + eps);
        const float n = use_batch_stats ? (float)(seg196[vid + 1] - seg196[vid]) : 1.f;
        prm[0][c] = mean[(size_t)vid * 128 + c];
        prm[1][c] = w[c] * rs;
        prm[2][c] = use_batch_stats ? (float)(sums[((size_t)vid * 2 + 0) * 128 + c]) / n : 0.f;
        prm[3][c] = use_batch_stats ? rs * ((float)(sums[((size_t)vid * 2 + 1) * 128 + c]) / n) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 8; ++q) { m[q] = prm[0][c0 + q]; k[q] = prm[1][c0 + q]; s1[q] = prm[2][c0 + q]; t2[q] = prm[3][c0 + q]; }
    }
    if (threadIdx.x == 0 && p + stride < R) {        // refill the other stage for the next pair once its 16 readers have arrived
      if (it >= 1) mbar_wait(bar0 + 16 + 8 * (s ^ 1), ((it - 1) >> 1) & 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(s ^ 1, p + stride);
    }
    mbar_wait(bar0 + 8 * s, (it >> 1) & 1);
    const uint8_t* st = smem + s * BW_STAGE;
    const uint4* xs = reinterpret_cast<const uint4*>(st) + c8;
    const float4* dps = reinterpret_cast<const float4*>(st + XT_BYTES) + c8 * 2;
    const uint2* as = reinterpret_cast<const uint2*>(st + XT_BYTES + DP_BYTES) + c8;
    // 196 positions over 32 slots = 6.125 sweeps: the slots rotate with the pair so that the extra sweep does not always fall to the
    // same warps (there is no block barrier to hide behind)
    for (int pos = (lane_pos + 4 * it) & 31; pos < 196; pos += 32) {
      const int iy = pos / 14, ix = pos - iy * 14;
      const uint4 xr = xs[pos * 16];
      float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      // windows (oy, ox) covering (iy, ix): oy in {iy/2, (iy+1)/2} (one window for even iy, two for odd), same along x
      const int oy0 = iy >> 1, ny = (iy & 1) + 1, ox0 = ix >> 1, nx = (ix & 1) + 1;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int oy = oy0 + a, ox = ox0 + b;
          if (a < ny && b < nx && oy < 7 && ox < 7) {
            const int cell = oy * 7 + ox;
            const unsigned tap = (unsigned)((iy - (2 * oy - 1)) * 3 + (ix - (2 * ox - 1)));
            const uint2 t = as[cell * 16];
            // a lane's 8 gradients are two 16-byte pieces 32 bytes apart from its neighbour's: lanes 4..7 of a quarter warp take the
            // pieces in the other order, so that one instruction touches every bank once
            const float4 da = dps[cell * 32 + hsel], db = dps[cell * 32 + (hsel ^ 1)];
            const float4 d0 = hsel ? db : da, d1 = hsel ? da : db;
            if (((t.x) & 0xffu) == tap) g[0] += d0.x;
            if (((t.x >> 8) & 0xffu) == tap) g[1] += d0.y;
            if (((t.x >> 16) & 0xffu) == tap) g[2] += d0.z;
            if (((t.x >> 24) & 0xffu) == tap) g[3] += d0.w;
            if (((t.y) & 0xffu) == tap) g[4] += d1.x;
            if (((t.y >> 8) & 0xffu) == tap) g[5] += d1.y;
            if (((t.y >> 16) & 0xffu) == tap) g[6] += d1.z;
            if (((t.y >> 24) & 0xffu) == tap) g[7] += d1.w;
          }
        }
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(&xr);
      uint32_t ow[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 xv = unpack2(xw[h]);
        float o0 = k[2 * h] * (g[2 * h] - s1[2 * h] - (xv.x - m[2 * h]) * t2[2 * h]);
        float o1 = k[2 * h + 1] * (g[2 * h + 1] - s1[2 * h + 1] - (xv.y - m[2 * h + 1]) * t2[2 * h + 1]);
        if (!(xv.x > 0.f)) o0 = 0.f;               // the ReLU in front of the BatchNorm
        if (!(xv.y > 0.f)) o1 = 0.f;
        cs[2 * h] += o0; cs[2 * h + 1] += o1;
        ow[h] = pack2(o0, o1);
      }
      reinterpret_cast<uint4*>(dx)[(p * 196 + pos) * 16 + c8] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar0 + 16 + 8 * s);     // this warp is done reading stage s
  }
  if (dx_colsum != nullptr) {                        // = the bias gradient of the 7x7 conv
    __syncthreads();
    for (int half = 0; half < 2; ++half) {
      if ((lane_pos >> 4) == half) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (half == 0) red[lane_pos & 15][c0 + q] = cs[q]; else red[lane_pos & 15][c0 + q] += cs[q];
        }
      }
      __syncthreads();
    }
    if (threadIdx.x < 128) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) t += red[i][threadIdx.x];
      atomicAdd(dx_colsum + threadIdx.x, t);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient: CTA = 8 warps; warp = 32 output channels x the 7 ky groups of one mask channel; reduction over the 196 positions
// ------------------------------------------------------------------------------------------------------------------
constexpr int DYP = 136;                 // dY tile pitch (bf16): 272-byte rows keep ldmatrix row addresses on distinct banks
constexpr int DYROWS = 208;              // 13 k-tiles of 16 positions; rows 196.. stay zero
constexpr int DW_ARR = 2 * MR * 8 + 8;   // words of one de-interleaved mask copy (+8: the four copies start 8 banks apart)
constexpr int DW_SMEM = DYROWS * DYP * 2 + 4 * DW_ARR * 4;

__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// The patch operand B[pos][(c, ky, kx)] = M[c][2 oy + ky][2 ox + kx] (M = padded mask) needs, per thread, the element PAIR of
// positions (pos, pos + 1) = (ox, ox + 1), ox even: two mask pixels 2 apart.  Copy a = (kx & 3) of the mask holds
// D_a[c][y][i] = M[c][y][2 (i + (a >> 1)) + (a & 1)], so that pair is the aligned word i = ox + 2 (kx >> 2) of row y.
__device__ __forceinline__ void park_mask_split(bf16* d, const float (&v)[PF]) {
#pragma unroll
  for (int i = 0; i < PF; ++i) {
    const int e = threadIdx.x + i * 256;
    if (e < NPIX) {
      const int c = e / 729, rem = e - c * 729, y = rem / 27, x = rem - y * 27;
      const int yp = y + 3, xp = x + 3, par = xp & 1, half = xp >> 1;
      const bf16 b = __float2bfloat16_rn(v[i]);
      d[(par * DW_ARR) * 2 + (c * MR + yp) * 16 + half] = b;                        // a = par      (shift 0)
      if (half >= 1) d[((par + 2) * DW_ARR) * 2 + (c * MR + yp) * 16 + half - 1] = b;   // a = par + 2  (shift 1)
    }
  }
}

__global__ void __launch_bounds__(256, 2)
mask_conv1_dw_kernel(const bf16* __restrict__ dy /*[R*196,128]*/, const float* __restrict__ masks, long long R,
                     float* __restrict__ partial /*[grid,128*98]*/) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* dys = reinterpret_cast<bf16*>(smem);
  uint32_t* dw = reinterpret_cast<uint32_t*>(smem + DYROWS * DYP * 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long p0 = R * blockIdx.x / gridDim.x, p1 = R * (blockIdx.x + 1) / gridDim.x;

  for (int i = threadIdx.x; i < DW_SMEM / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  // warp = 32 output channels (2 m-tiles) x the 7 (ky) column groups of one mask channel
  const int cgrp = warp & 3, nh = warp >> 2;
  float acc[2][7][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 7; ++n) { acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f; }
  float pf[PF];
  if (p0 < p1) prefetch_mask(masks, p0, pf);
  __syncthreads();

  const uint32_t dys_a = smem_addr(dys);
  // ldmatrix.trans row address of this lane: matrix mi = lane >> 3 -> positions +8 for mi >= 2, channels +8 for odd mi
  const uint32_t a_lane = dys_a + (((lane & 7) + ((lane >> 4) << 3)) * DYP + cgrp * 32 + ((lane >> 3) & 1) * 8) * 2;
  const uint32_t* b_lane = dw + (g & 3) * DW_ARR + (g >> 2) + nh * MR * 8;
  for (long long p = p0; p < p1; ++p) {
    // dY tile of the pair: 196 rows x 16 chunks of 16 bytes
    const uint4* src = reinterpret_cast<const uint4*>(dy) + p * (196 * 16);
    for (int e = threadIdx.x; e < 196 * 16; e += 256) {
      const int row = e >> 4, ch = e & 15;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dys_a + (row * DYP + ch * 8) * 2), "l"(src + e) : "memory");
    }
    park_mask_split(reinterpret_cast<bf16*>(dw), pf);
    if (p + 1 < p1) prefetch_mask(masks, p + 1, pf);
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll 1
    for (int kt = 0; kt < 13; ++kt) {
      uint32_t a0[4], a1[4];
      ldsm_x4_t(a0, a_lane + kt * 16 * DYP * 2);
      ldsm_x4_t(a1, a_lane + kt * 16 * DYP * 2 + 32);
      const int q0 = min(kt * 16 + 2 * t, 194), q1 = min(kt * 16 + 8 + 2 * t, 194);
      const int oy0 = q0 / 14, ox0 = q0 - oy0 * 14, oy1 = q1 / 14, ox1 = q1 - oy1 * 14;
      const uint32_t* r0 = b_lane + (2 * oy0) * 8 + (ox0 >> 1);
      const uint32_t* r1 = b_lane + (2 * oy1) * 8 + (ox1 >> 1);
#pragma unroll
      for (int n = 0; n < 7; ++n) {
        const uint32_t b0 = r0[n * 8], b1 = r1[n * 8];
        mma16816(acc[0][n], a0, b0, b1);
        mma16816(acc[1][n], a1, b0, b1);
      }
    }
    __syncthreads();
  }
  float* out = partial + (size_t)blockIdx.x * (128 * 98);
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int n = 0; n < 7; ++n) {
      const int co = cgrp * 32 + m * 16 + g, col = (nh * 7 + n) * 7 + 2 * t;
      out[co * 98 + col] = acc[m][n][0];
      out[(co + 8) * 98 + col] = acc[m][n][2];
      if (2 * t + 1 < 7) {
        out[co * 98 + col + 1] = acc[m][n][1];
        out[(co + 8) * 98 + col + 1] = acc[m][n][3];
      }
    }
}

__global__ void sum_partials_kernel(const float* __restrict__ partial, int n_part, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < n_part; ++k) s += partial[(size_t)k * n + i];
  out[i] = s;
}

}  // namespace
}  // namespace nlv

using namespace nlv;
#define STREAM ((cudaStream_t)stream)

/* Conv2d(2,128,k7,s2,p3) + bias + ReLU over the pair masks [r,2,27,27] -> out bf16 [r*196,128] (NHWC), and the training-mode
 * BatchNorm statistics of that map per video (seg196 / pair_video as for nlv_bn_stats / nlv_bn_apply with row_div = 196):
 * mean/var [nv,128], running statistics updated.  sums_ws: double[nv*2*128].  mean == NULL: no statistics (eval mode). */
int nlv_mask_conv1_fwd(const float* masks, const float* w, const float* bias, long long r, const int* pair_video, const int* seg196,
                       int nv, void* out, float momentum, double* sums_ws, float* mean, float* var, float* running_mean,
                       float* running_var, void* stream) {
  NLV_CHECK_ARG(r >= 0 && nv >= 1, "mask_conv1_fwd: bad sizes");
  NLV_CHECK_ARG(masks && w && bias && out && sums_ws, "mask_conv1_fwd: null pointer");
  { int zrc = zero_fill(reinterpret_cast<float*>(sums_ws), 1, 4 * nv * 128, 4 * nv * 128, STREAM); if (zrc != NLV_OK) return zrc; }
  if (r > 0) {
    const long long want = 2ll * sm_count();
    const int grid = (int)(r < want ? r : want);
    mask_conv1_fwd_kernel<<<grid, 256, 0, STREAM>>>(masks, w, bias, pair_video, r, reinterpret_cast<bf16*>(out), sums_ws);
    NLV_CHECK_LAUNCH();
  }
  if (mean != nullptr) {
    NLV_CHECK_ARG(var && seg196, "mask_conv1_fwd: null pointer");
    return launch_bn_finalize(sums_ws, seg196, nv, 128, momentum, mean, var, running_mean, running_var, STREAM);
  }
  return NLV_OK;
}

/* y = maxpool3x3/2(BatchNorm(x)) for x bf16 [r*196,128] -> y bf16 [r*49,128], argmax tap u8 [r*49,128]; mean/var [nv,128] indexed by
 * pair_video[pair] (NULL: segment 0).  Same values as nlv_bn_apply (bf16 output) followed by nlv_maxpool_fwd; same taps except
 * where two BatchNorm outputs tie only after their rounding to bf16 (then the larger activation, the fp32 argmax, is taken).
 * xmax (nullable): bf16 [r*49,128], the activation x at every argmax (input of nlv_pool_bn_bwd). */
int nlv_bn_apply_maxpool(const void* x, const int* pair_video, const float* mean, const float* var, const float* w, const float* b,
                         float eps, long long r, void* y, uint8_t* argmax, void* xmax, void* stream) {
  NLV_CHECK_ARG(r >= 0 && r < 0x7fffffffll, "bn_apply_maxpool: bad sizes");
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(x && mean && var && w && b && y && argmax, "bn_apply_maxpool: null pointer");
  static bool attr_set = false;
  if (!attr_set) {
    NLV_CHECK_CUDA(cudaFuncSetAttribute(bn_apply_maxpool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM));
    attr_set = true;
  }
  const long long want = 2ll * sm_count();
  bn_apply_maxpool_kernel<<<(unsigned)(r < want ? r : want), 256, F2_SMEM, STREAM>>>(reinterpret_cast<const bf16*>(x), pair_video, mean, var, w, b,
                                                                                   eps, r, reinterpret_cast<bf16*>(y), argmax,
                                                                                   reinterpret_cast<bf16*>(xmax));
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* Weight gradient of the 7x7 mask conv: dw[128,98] (assigned) = sum over pairs / positions of dy[r*196,128] (bf16) x mask patches.
 * ws: float[nlv_mask_conv1_dw_ws_floats()] workspace of per-CTA partial sums. */
long long nlv_mask_conv1_dw_ws_floats(void) { return 2ll * sm_count() * 128 * 98; }
int nlv_mask_conv1_dw(const void* dy, const float* masks, long long r, float* ws, float* dw, void* stream) {
  NLV_CHECK_ARG(r >= 0, "mask_conv1_dw: bad sizes");
  NLV_CHECK_ARG(dy && masks && ws && dw, "mask_conv1_dw: null pointer");
  static bool attr_set = false;
  if (!attr_set) {
    NLV_CHECK_CUDA(cudaFuncSetAttribute(mask_conv1_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM));
    attr_set = true;
  }
  const long long want = 2ll * sm_count();
  const int grid = (int)(r < want ? (r > 0 ? r : 1) : want);
  mask_conv1_dw_kernel<<<grid, 256, DW_SMEM, STREAM>>>(reinterpret_cast<const bf16*>(dy), masks, r, ws);
  NLV_CHECK_LAUNCH();
  sum_partials_kernel<<<cdiv(128 * 98, 256), 256, 0, STREAM>>>(ws, grid, 128 * 98, dw);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* Backward of conv-ReLU -> BatchNorm -> MaxPool2d(3,2,1) from the pooled gradient dp fp32 [r*49,128] (+ argmax taps, the activation map
 * x bf16 [r*196,128] and xmax from nlv_bn_apply_maxpool): dx bf16 [r*196,128] = gradient at the conv output (ReLU applied), dw / db
 * (BatchNorm weight / bias gradients, accumulated), dx_colsum (column sums of dx = conv bias gradient, accumulated; nullable).
 * Same result as nlv_maxpool_bwd + nlv_bn_bwd_colsum(gate_by_x = 1) without the dense pooled-gradient map.  sums_ws: double[nv*2*128]. */
int nlv_pool_bn_bwd(const float* dp, const uint8_t* argmax, const void* x, const void* xmax, const int* pair_video, const int* seg196,
                    const int* seg49, int nv, const float* mean, const float* var, const float* w, float eps, int use_batch_stats,
                    long long r, double* sums_ws, void* dx, float* dw, float* db, float* dx_colsum, void* stream) {
  NLV_CHECK_ARG(r >= 0 && nv >= 1, "pool_bn_bwd: bad sizes");
  NLV_CHECK_ARG(dp && argmax && x && xmax && seg196 && seg49 && mean && var && w && sums_ws && dx && dw && db, "pool_bn_bwd: null pointer");
  { int zrc = zero_fill(reinterpret_cast<float*>(sums_ws), 1, 4 * nv * 128, 4 * nv * 128, STREAM); if (zrc != NLV_OK) return zrc; }
  if (r == 0) return NLV_OK;
  // sum_dy and sum_dy_xhat per (video, channel): every non-zero of the dense gradient is one pooled cell
  int rc = launch_bn_sums_bwd_v8(dp, NLV_F32, 128, xmax, NLV_BF16, 128, nullptr, 0, 0, seg49, nv, mean, var, eps, r * 49, 128, sums_ws, STREAM);
  if (rc != NLV_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    NLV_CHECK_CUDA(cudaFuncSetAttribute(pool_bn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BW_SMEM));
    attr_set = true;
  }
  const long long want = sm_count();
  pool_bn_bwd_apply_kernel<<<(unsigned)(r < want ? r : want), 512, BW_SMEM, STREAM>>>(dp, argmax, reinterpret_cast<const bf16*>(x), pair_video,
                                                                                    seg196, mean, var, w, eps, sums_ws, use_batch_stats, r,
                                                                                    reinterpret_cast<bf16*>(dx), dx_colsum);
  NLV_CHECK_LAUNCH();
  return launch_bn_bwd_param(sums_ws, nv, 128, dw, db, STREAM);
}
