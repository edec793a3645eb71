// Fused variable-length multi-head attention over contiguous token segments (forward + backward).
//
// One launch covers every segment of the batch: a segment is a frame (spatial encoder,
// lib/transformer.py:20-22), a 2-frame sliding window (temporal decoder, :49-52, lib/transformer_wk.py:163-171)
// or a per-class sequence (lib/dsg_detr.py:545-559).  Segments are unpadded, so no key_padding_mask exists:
// a query only ever sees the keys of its own segment ("bool masking" semantics).
//
// Work decomposition: a work item = 16 consecutive rows of one segment; grid = (work items, heads); 4 warps per CTA,
// 4 rows per warp, every warp self-contained.  A warp keeps its own 4 rows (Q, and dO / K,V in the backward) in
// REGISTERS, sliced over the lanes: lane owns the 2-element words lane + 32 j of the head (j < 4 -> head_dim <= 256),
// so every global access is a contiguous 128/256-byte run of one row.  The other side of the product streams by row:
// a row is loaded once into registers, used for the per-lane partial dot products (all-reduced across the warp with
// a halving/doubling shuffle schedule) and reused for the accumulation.  No shared memory, no barriers; softmax is
// online in the forward (one rescale per chunk of 4 keys) and recomputed from the stored log-sum-exp in the backward.
// fp32 math; I/O fp32 or bf16; head_dim even (242 here).
#include <stdlib.h>

#include "common.cuh"

namespace nlv {
namespace {

constexpr int NW = 4;      // words (element pairs) owned per lane: word lane + 32*i  -> head_dim <= 256
constexpr int THREADS = 128;

struct AttnArgs {
  const void *q, *k, *v;   // [rows, *] with row strides ldq/ldk/ldv (elements), head h at column h*hd
  int ldq, ldk, ldv;
  int hd, heads;
  float scale;
  const int4* work;        // {segment first row, segment length, first row of this item relative to segment, padded keys}
  // torch-1.10.1 semantics of the int key_padding_mask of lib/transformer_wk.py:154: a frame padded to the longest frame of its
  // video keeps its work.w padded keys in the softmax with +1 added to their logits; a padded row is all-zero, so its key /
  // value are the K / V slices of in_proj_bias (fp32 [heads*hd] each).  NULL: true masking (segments are simply unpadded).
  const float* kpad;
  const float* vpad;
};

template <typename T> __device__ __forceinline__ void st2(T* row, int w, float a, float b);
template <> __device__ __forceinline__ void st2<float>(float* row, int w, float a, float b) {
  *reinterpret_cast<float2*>(row + 2 * w) = make_float2(a, b);
}
template <> __device__ __forceinline__ void st2<__nv_bfloat16>(__nv_bfloat16* row, int w, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(row + 2 * w) = __floats2bfloat162_rn(a, b);
}

// Raw (still packed) words of one row, sliced over the warp: lane owns words lane + 32 j.  Loads are unconditional
// (out-of-range lanes / rows re-read a valid word and are zeroed at unpack time) and go through volatile asm so that
// the whole batch of a chunk is in flight before the first use — the compiler otherwise interleaves each load with
// its consumer and the warp pays one DRAM latency per word instead of one per chunk.
template <typename T> struct Raw;
template <> struct Raw<float> {
  typedef float2 W;
  static __device__ __forceinline__ W ldg(const float* p) {
    W v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
  }
  static __device__ __forceinline__ float2 unpack(W v) { return v; }
};
template <> struct Raw<__nv_bfloat16> {
  typedef unsigned int W;
  static __device__ __forceinline__ W ldg(const __nv_bfloat16* p) {
    W v;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
  }
  static __device__ __forceinline__ float2 unpack(W v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
};

template <typename T>
__device__ __forceinline__ void load_raw(const T* row, int nwords, typename Raw<T>::W (&r)[NW]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NW; ++j) {
    const int wd = lane + 32 * j;
    r[j] = Raw<T>::ldg(row + 2 * (wd < nwords ? wd : 0));
  }
}
template <typename T>
__device__ __forceinline__ void unpack_row(const typename Raw<T>::W (&r)[NW], int nwords, bool valid, float mul, float (&x)[2 * NW]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NW; ++j) {
    const float2 t = Raw<T>::unpack(r[j]);
    const bool ok = valid && lane + 32 * j < nwords;
    x[2 * j] = ok ? t.x * mul : 0.f; x[2 * j + 1] = ok ? t.y * mul : 0.f;
  }
}
template <typename T>
__device__ __forceinline__ void store_row(T* row, int nwords, const float (&x)[2 * NW], float mul) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NW; ++j) {
    const int wd = lane + 32 * j;
    if (wd < nwords) st2<T>(row, wd, x[2 * j] * mul, x[2 * j + 1] * mul);
  }
}
__device__ __forceinline__ float dot8(const float (&a)[2 * NW], const float (&b)[2 * NW]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * NW; ++i) s = fmaf(a[i], b[i], s);
  return s;
}
__device__ __forceinline__ void axpy8(float c, const float (&x)[2 * NW], float (&acc)[2 * NW]) {
#pragma unroll
  for (int i = 0; i < 2 * NW; ++i) acc[i] = fmaf(c, x[i], acc[i]);
}

// All-reduce N (power of two, <= 16) per-lane partial sums over the warp; every lane ends with all N totals.
// Reduce-scatter by recursive halving (each exchange carries half of what is left), plain butterflies for the
// remaining lane bits, then the mirror-image all-gather: 2(N-1) + log2(32/N) shuffles instead of 5 N.
template <int N>
__device__ __forceinline__ void allreduce(float (&v)[N]) {
  static_assert(N >= 2 && N <= 16 && (N & (N - 1)) == 0, "N must be 2, 4, 8 or 16");
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int half = N / 2, o = 16; half >= 1; half >>= 1, o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      if (i < half) {
        const float send = up ? v[i] : v[i + half];
        const float keep = up ? v[i + half] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
      }
    }
  }
  constexpr int rem = 32 / N;
#pragma unroll
  for (int o = rem / 2; o >= 1; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
#pragma unroll
  for (int cnt = 1, o = rem; cnt < N; cnt <<= 1, o <<= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      if (i < cnt) {
        const float other = __shfl_xor_sync(0xffffffffu, v[i], o);
        v[i + cnt] = up ? v[i] : other;
        v[i] = up ? other : v[i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward: keys in chunks of 4 (16 scores all-reduced at once, one softmax rescale per chunk); the K and V rows of
// the next chunk are requested before the current chunk is computed.
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(THREADS)
attn_fwd_kernel(AttnArgs a, TO* __restrict__ o, int ldo, float* __restrict__ lse) {
  constexpr int C = 4;
  typedef typename Raw<TI>::W W;
  const int hd = a.hd, nwords = hd >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int4 w = a.work[blockIdx.x / a.heads];            // heads of one item are neighbouring CTAs: together they
  const int h = blockIdx.x % a.heads, col0 = h * hd;       // read whole contiguous rows (DRAM page / TLB locality)
  const long long seg0 = w.x;
  const int L = w.y, q0 = w.z + warp * 4;       // first row of this warp inside the segment
  const int nq = min(4, L - q0);
  if (nq <= 0) return;                          // warps are independent: no barrier anywhere
  const TI* Q = reinterpret_cast<const TI*>(a.q) + col0;
  const TI* K = reinterpret_cast<const TI*>(a.k) + col0;
  const TI* V = reinterpret_cast<const TI*>(a.v) + col0;

  W rk[C][NW], rv[C][NW];
  float q[4][2 * NW], acc[4][2 * NW], m[4], l[4];
  {
    W rq[4][NW];
#pragma unroll
    for (int i = 0; i < 4; ++i) load_raw<TI>(Q + (size_t)(seg0 + min(q0 + i, L - 1)) * a.ldq, nwords, rq[i]);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const size_t row = seg0 + min(c, L - 1);
      load_raw<TI>(K + row * a.ldk, nwords, rk[c]);
      load_raw<TI>(V + row * a.ldv, nwords, rv[c]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      unpack_row<TI>(rq[i], nwords, i < nq, a.scale, q[i]);
      m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
      for (int e = 0; e < 2 * NW; ++e) acc[i][e] = 0.f;
    }
  }
  for (int k0 = 0; k0 < L; k0 += C) {
    W nk[C][NW], nv[C][NW];
    const bool more = k0 + C < L;
    if (more) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const size_t row = seg0 + min(k0 + C + c, L - 1);
        load_raw<TI>(K + row * a.ldk, nwords, nk[c]);
        load_raw<TI>(V + row * a.ldv, nwords, nv[c]);
      }
    }
    float s[C * 4];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float kr[2 * NW];
      unpack_row<TI>(rk[c], nwords, k0 + c < L, 1.f, kr);
#pragma unroll
      for (int i = 0; i < 4; ++i) s[c * 4 + i] = dot8(q[i], kr);
    }
    allreduce<C * 4>(s);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = m[i];
#pragma unroll
      for (int c = 0; c < C; ++c) if (k0 + c < L) mx = fmaxf(mx, s[c * 4 + i]);
      const float corr = __expf(m[i] - mx);     // m = -inf on the first chunk -> 0
      m[i] = mx;
      float ps = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float p = k0 + c < L ? __expf(s[c * 4 + i] - mx) : 0.f;
        s[c * 4 + i] = p;
        ps += p;
      }
      l[i] = l[i] * corr + ps;
#pragma unroll
      for (int e = 0; e < 2 * NW; ++e) acc[i][e] *= corr;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float vr[2 * NW];
      unpack_row<TI>(rv[c], nwords, k0 + c < L, 1.f, vr);   // p = 0 for the rows beyond L anyway
#pragma unroll
      for (int i = 0; i < 4; ++i) axpy8(s[c * 4 + i], vr, acc[i]);
    }
    if (more) {
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < NW; ++j) { rk[c][j] = nk[c][j]; rv[c][j] = nv[c][j]; }
    }
  }
  if (a.kpad != nullptr && w.w > 0) {
    // the padded keys: one extra logit q . b_k + 1 per row, counted w.w times, carrying b_v
    float kr[2 * NW], vr[2 * NW], z[4];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const int wd = lane + 32 * j;
      const bool ok = wd < nwords;
      const float2 kk = ok ? *reinterpret_cast<const float2*>(a.kpad + col0 + 2 * wd) : make_float2(0.f, 0.f);
      const float2 vv = ok ? *reinterpret_cast<const float2*>(a.vpad + col0 + 2 * wd) : make_float2(0.f, 0.f);
      kr[2 * j] = kk.x; kr[2 * j + 1] = kk.y; vr[2 * j] = vv.x; vr[2 * j + 1] = vv.y;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i] = dot8(q[i], kr);
    allreduce<4>(z);
    const float npad = (float)w.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float zi = z[i] + 1.f;
      const float mx = fmaxf(m[i], zi);
      const float corr = __expf(m[i] - mx), p = npad * __expf(zi - mx);
      m[i] = mx;
      l[i] = l[i] * corr + p;
#pragma unroll
      for (int e = 0; e < 2 * NW; ++e) acc[i][e] = fmaf(p, vr[e], acc[i][e] * corr);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i >= nq) continue;
    const long long row = seg0 + q0 + i;
    store_row<TO>(o + (size_t)row * ldo + col0, nwords, acc[i], 1.f / l[i]);
    if (lane == 0 && lse != nullptr) lse[row * a.heads + h] = m[i] + __logf(l[i]);
  }
}

// ------------------------------------------------------------------------------------------
// backward, query side: dQ (and delta = rowsum(dO * O), stored for the key side).  Keys two at a time, next pair
// prefetched; the key row loaded for the score is reused from registers for the dQ accumulation.
// ------------------------------------------------------------------------------------------
template <typename TI, typename TG>
__global__ void __launch_bounds__(THREADS)
attn_bwd_dq_kernel(AttnArgs a, const TI* __restrict__ o, int ldo, const TG* __restrict__ dout, int lddo,
                   const float* __restrict__ lse, float* __restrict__ delta, TI* __restrict__ dq, int lddq) {
  constexpr int C = 2;
  typedef typename Raw<TI>::W W;
  typedef typename Raw<TG>::W WG;
  const int hd = a.hd, nwords = hd >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int4 w = a.work[blockIdx.x / a.heads];
  const int h = blockIdx.x % a.heads, col0 = h * hd;
  const long long seg0 = w.x;
  const int L = w.y, q0 = w.z + warp * 4;
  const int nq = min(4, L - q0);
  if (nq <= 0) return;
  const TI* Q = reinterpret_cast<const TI*>(a.q) + col0;
  const TI* K = reinterpret_cast<const TI*>(a.k) + col0;
  const TI* V = reinterpret_cast<const TI*>(a.v) + col0;

  W rk[C][NW], rv[C][NW];
  float q[4][2 * NW], g[4][2 * NW], acc[4][2 * NW], dl[4], ls[4];
  {
    W rq[4][NW], ro[4][NW];
    WG rg[4][NW];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t row = seg0 + min(q0 + i, L - 1);
      load_raw<TI>(Q + row * a.ldq, nwords, rq[i]);
      load_raw<TG>(dout + row * lddo + col0, nwords, rg[i]);
      load_raw<TI>(o + row * ldo + col0, nwords, ro[i]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const size_t row = seg0 + min(c, L - 1);
      load_raw<TI>(K + row * a.ldk, nwords, rk[c]);
      load_raw<TI>(V + row * a.ldv, nwords, rv[c]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float orow[2 * NW];
      unpack_row<TI>(rq[i], nwords, i < nq, a.scale, q[i]);
      unpack_row<TG>(rg[i], nwords, i < nq, 1.f, g[i]);
      unpack_row<TI>(ro[i], nwords, i < nq, 1.f, orow);
      dl[i] = dot8(g[i], orow);
      ls[i] = i < nq ? lse[(seg0 + q0 + i) * a.heads + h] : 0.f;
#pragma unroll
      for (int e = 0; e < 2 * NW; ++e) acc[i][e] = 0.f;
    }
  }
  allreduce<4>(dl);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (lane == 0 && i < nq) delta[(seg0 + q0 + i) * a.heads + h] = dl[i];

  for (int k0 = 0; k0 < L; k0 += C) {
    W nk[C][NW], nv[C][NW];
    const bool more = k0 + C < L;
    if (more) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const size_t row = seg0 + min(k0 + C + c, L - 1);
        load_raw<TI>(K + row * a.ldk, nwords, nk[c]);
        load_raw<TI>(V + row * a.ldv, nwords, nv[c]);
      }
    }
    float kr[C][2 * NW], sv[C * 8];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float vr[2 * NW];
      unpack_row<TI>(rk[c], nwords, k0 + c < L, 1.f, kr[c]);
      unpack_row<TI>(rv[c], nwords, k0 + c < L, 1.f, vr);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sv[c * 8 + i] = dot8(q[i], kr[c]);
        sv[c * 8 + 4 + i] = dot8(g[i], vr);
      }
    }
    allreduce<C * 8>(sv);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (k0 + c < L) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p = __expf(sv[c * 8 + i] - ls[i]);
          axpy8(p * (sv[c * 8 + 4 + i] - dl[i]), kr[c], acc[i]);
        }
      }
    }
    if (more) {
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int j = 0; j < NW; ++j) { rk[c][j] = nk[c][j]; rv[c][j] = nv[c][j]; }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < nq) store_row<TI>(dq + (size_t)(seg0 + q0 + i) * lddq + col0, nwords, acc[i], a.scale);
}

// ------------------------------------------------------------------------------------------
// backward, key side: dK, dV.  The warp's 4 rows are KEYS (K and V rows live in registers); queries stream by,
// the next query's Q / dO rows prefetched while the current one is processed.
// ------------------------------------------------------------------------------------------
template <typename TI, typename TG>
__global__ void __launch_bounds__(THREADS)
attn_bwd_dkv_kernel(AttnArgs a, const TG* __restrict__ dout, int lddo, const float* __restrict__ lse,
                    const float* __restrict__ delta, TI* __restrict__ dk, int lddk, TI* __restrict__ dv, int lddv) {
  typedef typename Raw<TI>::W W;
  typedef typename Raw<TG>::W WG;
  const int hd = a.hd, nwords = hd >> 1;
  const int warp = threadIdx.x >> 5;
  const int4 w = a.work[blockIdx.x / a.heads];
  const int h = blockIdx.x % a.heads, col0 = h * hd;
  const long long seg0 = w.x;
  const int L = w.y, k0 = w.z + warp * 4;
  const int nkeys = min(4, L - k0);
  if (nkeys <= 0) return;
  const TI* Q = reinterpret_cast<const TI*>(a.q) + col0;
  const TI* K = reinterpret_cast<const TI*>(a.k) + col0;
  const TI* V = reinterpret_cast<const TI*>(a.v) + col0;
  const TG* G = dout + col0;

  W rq[NW];
  WG rg[NW];
  float kk[4][2 * NW], vv[4][2 * NW], accK[4][2 * NW], accV[4][2 * NW];
  float lq, dr;
  {
    W rk[4][NW], rv[4][NW];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const size_t row = seg0 + min(k0 + i, L - 1);
      load_raw<TI>(K + row * a.ldk, nwords, rk[i]);
      load_raw<TI>(V + row * a.ldv, nwords, rv[i]);
    }
    load_raw<TI>(Q + (size_t)seg0 * a.ldq, nwords, rq);
    load_raw<TG>(G + (size_t)seg0 * lddo, nwords, rg);
    lq = lse[seg0 * a.heads + h];
    dr = delta[seg0 * a.heads + h];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      unpack_row<TI>(rk[i], nwords, i < nkeys, a.scale, kk[i]);
      unpack_row<TI>(rv[i], nwords, i < nkeys, 1.f, vv[i]);
#pragma unroll
      for (int e = 0; e < 2 * NW; ++e) { accK[i][e] = 0.f; accV[i][e] = 0.f; }
    }
  }
  for (int r = 0; r < L; ++r) {
    float qr[2 * NW], gr[2 * NW], sv[8];
    unpack_row<TI>(rq, nwords, true, 1.f, qr);
    unpack_row<TG>(rg, nwords, true, 1.f, gr);
    const float lq_c = lq, dr_c = dr;
    if (r + 1 < L) {
      const long long row = seg0 + r + 1;
      load_raw<TI>(Q + (size_t)row * a.ldq, nwords, rq);
      load_raw<TG>(G + (size_t)row * lddo, nwords, rg);
      lq = lse[row * a.heads + h];
      dr = delta[row * a.heads + h];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sv[i] = dot8(kk[i], qr);        // already times scale
      sv[4 + i] = dot8(vv[i], gr);
    }
    allreduce<8>(sv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float p = i < nkeys ? __expf(sv[i] - lq_c) : 0.f;
      axpy8(p, gr, accV[i]);
      axpy8(p * (sv[4 + i] - dr_c), qr, accK[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (i >= nkeys) continue;
    const long long row = seg0 + k0 + i;
    store_row<TI>(dk + (size_t)row * lddk + col0, nwords, accK[i], a.scale);
    store_row<TI>(dv + (size_t)row * lddv + col0, nwords, accV[i], 1.f);
  }
}

int check_common(int hd, int heads, int n_work, int ld_all_even) {
  NLV_CHECK_ARG(hd > 0 && hd <= 64 * NW && (hd & 1) == 0, "attention: head_dim=%d unsupported (even, max %d)", hd, 64 * NW);
  NLV_CHECK_ARG(heads > 0 && n_work >= 0 && (long long)heads * n_work < (1ll << 31), "attention: bad sizes");
  NLV_CHECK_ARG(ld_all_even, "attention: row strides must be even (vector access)");
  return NLV_OK;
}

}  // namespace
}  // namespace nlv

// tensor-core path for bf16 I/O (attn_mma.cu); NLV_ATTN_SIMT=1 keeps the SIMT kernels (A/B runs, unit tests of both)
namespace nlv {
bool attn_mma_supported(int hd, int heads, int ld_or);
int launch_attn_fwd_mma(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int hd, int heads, float scale,
                        const void* work, int n_work, void* o, int ldo, float* lse, const nlv_dropout* drop, cudaStream_t s);
int launch_attn_bwd_mma(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int hd, int heads, float scale,
                        const void* work, int n_work, int n_long, const void* dout, int lddo, const float* lse, float* delta, void* dq,
                        int lddq, void* dk, int lddk, void* dv, int lddv, const nlv_dropout* drop, cudaStream_t s);
static bool use_mma() {
  const char* e = getenv("NLV_ATTN_SIMT");
  return !(e != nullptr && e[0] == '1');
}
}  // namespace nlv

using namespace nlv;
#define STREAM ((cudaStream_t)stream)
typedef __nv_bfloat16 bf16;

extern "C" {

/* work: int4[n_work] = {segment first row, segment length, item first row within the segment, 0}; every segment is
 * covered by ceil(len/16) items.  lse: float[rows*heads] (optional, needed for backward).
 * q/k/v share in_dtype; base pointers must be 8-byte (f32) / 4-byte (bf16) aligned and row strides even. */
int nlv_attn_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                 float scale, const void* work, int n_work, void* o, int ldo, int o_dtype, float* lse, void* stream) {
  return nlv_attn_fwd_drop(q, ldq, k, ldk, v, ldv, in_dtype, hd, heads, scale, work, n_work, o, ldo, o_dtype, lse, nullptr, stream);
}

int nlv_attn_fwd_drop(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                      float scale, const void* work, int n_work, void* o, int ldo, int o_dtype, float* lse, const nlv_dropout* drop,
                      void* stream) {
  return nlv_attn_fwd_padkeys(q, ldq, k, ldk, v, ldv, in_dtype, hd, heads, scale, work, n_work, o, ldo, o_dtype, lse, drop, nullptr, nullptr,
                              stream);
}

int nlv_attn_fwd_padkeys(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                         float scale, const void* work, int n_work, void* o, int ldo, int o_dtype, float* lse, const nlv_dropout* drop,
                         const float* kpad, const float* vpad, void* stream) {
  int rc = check_common(hd, heads, n_work, ((ldq | ldk | ldv | ldo) & 1) == 0);
  if (rc != NLV_OK) return rc;
  if (n_work == 0) return NLV_OK;
  NLV_CHECK_ARG(q && k && v && work && o, "attn_fwd: null pointer");
  NLV_CHECK_ARG((kpad == nullptr) == (vpad == nullptr), "attn_fwd: kpad and vpad go together");
  NLV_CHECK_ARG(kpad == nullptr || ((((uintptr_t)kpad | (uintptr_t)vpad) & 7) == 0), "attn_fwd: kpad / vpad must be 8-byte aligned");
  if (kpad == nullptr && in_dtype == NLV_BF16 && o_dtype == NLV_BF16 && use_mma() && attn_mma_supported(hd, heads, ldq | ldk | ldv | ldo))
    return launch_attn_fwd_mma(q, ldq, k, ldk, v, ldv, hd, heads, scale, work, n_work, o, ldo, lse, drop, STREAM);
  if (drop != nullptr && drop->thr16 != 0u) {
    nlv::set_error("attn_fwd: attention-weight dropout is implemented on the bf16 tensor-core path only");
    return NLV_ERR_UNSUPPORTED;
  }
  AttnArgs a{q, k, v, ldq, ldk, ldv, hd, heads, scale, (const int4*)work, kpad, vpad};
  const dim3 grid((unsigned)n_work * (unsigned)heads);
#define FWD(TI, TO) attn_fwd_kernel<TI, TO><<<grid, THREADS, 0, STREAM>>>(a, (TO*)o, ldo, lse)
  if (in_dtype == NLV_BF16 && o_dtype == NLV_BF16) FWD(bf16, bf16);
  else if (in_dtype == NLV_BF16) FWD(bf16, float);
  else if (o_dtype == NLV_BF16) FWD(float, bf16);
  else FWD(float, float);
#undef FWD
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* delta: float[rows*heads] workspace written by the query-side kernel and read by the key-side kernel.
 * o and dq/dk/dv share q's dtype (in_dtype == o_dtype == dqkv_dtype); dout has do_dtype. */
int nlv_attn_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                 float scale, const void* work, int n_work, const void* o, int ldo, int o_dtype, const void* dout, int lddo,
                 int do_dtype, const float* lse, float* delta, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv,
                 int dqkv_dtype, void* stream) {
  return nlv_attn_bwd_drop(q, ldq, k, ldk, v, ldv, in_dtype, hd, heads, scale, work, n_work, o, ldo, o_dtype, dout, lddo, do_dtype, lse, delta,
                           dq, lddq, dk, lddk, dv, lddv, dqkv_dtype, nullptr, stream);
}

int nlv_attn_bwd_drop(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                      float scale, const void* work, int n_work, const void* o, int ldo, int o_dtype, const void* dout, int lddo,
                      int do_dtype, const float* lse, float* delta, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv,
                      int dqkv_dtype, const nlv_dropout* drop, void* stream) {
  return nlv_attn_bwd_sorted(q, ldq, k, ldk, v, ldv, in_dtype, hd, heads, scale, work, n_work, -1, o, ldo, o_dtype, dout, lddo, do_dtype, lse,
                             delta, dq, lddq, dk, lddk, dv, lddv, dqkv_dtype, drop, stream);
}

int nlv_attn_bwd_sorted(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                        float scale, const void* work, int n_work, int n_long_work, const void* o, int ldo, int o_dtype, const void* dout,
                        int lddo, int do_dtype, const float* lse, float* delta, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv,
                        int dqkv_dtype, const nlv_dropout* drop, void* stream) {
  int rc = check_common(hd, heads, n_work, ((ldq | ldk | ldv | ldo | lddo | lddq | lddk | lddv) & 1) == 0);
  if (rc != NLV_OK) return rc;
  if (n_work == 0) return NLV_OK;
  NLV_CHECK_ARG(q && k && v && work && o && dout && lse && delta && dq && dk && dv, "attn_bwd: null pointer");
  NLV_CHECK_ARG(in_dtype == o_dtype && in_dtype == dqkv_dtype, "attn_bwd: q/k/v, o and dq/dk/dv must share one dtype");
  if (in_dtype == NLV_BF16 && do_dtype == NLV_BF16 && use_mma() &&
      attn_mma_supported(hd, heads, ldq | ldk | ldv | ldo | lddo | lddq | lddk | lddv))
    return launch_attn_bwd_mma(q, ldq, k, ldk, v, ldv, hd, heads, scale, work, n_work, n_long_work, dout, lddo, lse, delta, dq, lddq, dk, lddk,
                               dv, lddv, drop, STREAM);
  if (drop != nullptr && drop->thr16 != 0u) {
    nlv::set_error("attn_bwd: attention-weight dropout is implemented on the bf16 tensor-core path only");
    return NLV_ERR_UNSUPPORTED;
  }
  AttnArgs a{q, k, v, ldq, ldk, ldv, hd, heads, scale, (const int4*)work, nullptr, nullptr};
  const dim3 grid((unsigned)n_work * (unsigned)heads);
#define BWD(TI, TG)                                                                                                         \
  do {                                                                                                                      \
    attn_bwd_dq_kernel<TI, TG><<<grid, THREADS, 0, STREAM>>>(a, (const TI*)o, ldo, (const TG*)dout, lddo, lse, delta,       \
                                                            (TI*)dq, lddq);                                                 \
    NLV_CHECK_LAUNCH();                                                                                                     \
    attn_bwd_dkv_kernel<TI, TG><<<grid, THREADS, 0, STREAM>>>(a, (const TG*)dout, lddo, lse, delta, (TI*)dk, lddk,          \
                                                             (TI*)dv, lddv);                                                \
  } while (0)
  if (in_dtype == NLV_BF16 && do_dtype == NLV_BF16) BWD(bf16, bf16);
  else if (in_dtype == NLV_BF16) BWD(bf16, float);
  else if (do_dtype == NLV_BF16) BWD(float, bf16);
  else BWD(float, float);
#undef BWD
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
}
