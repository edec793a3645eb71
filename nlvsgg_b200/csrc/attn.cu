// Fused variable-length multi-head attention over contiguous token segments (forward + backward).
//
// One launch covers every segment of the batch: a segment is a frame (spatial encoder,
// lib/transformer.py:20-22), a 2-frame sliding window (temporal decoder, :49-52, lib/transformer_wk.py:163-171)
// or a per-class sequence (lib/dsg_detr.py:545-559).  Segments are unpadded, so no key_padding_mask exists:
// a query only ever sees the keys of its own segment ("bool masking" semantics).
//
// Work decomposition: a work item = 16 consecutive rows of one segment; grid = (work items, heads); 4 warps per CTA,
// 4 rows per warp.  Only the item's own 16 rows are staged (transposed, fp32) in shared memory; the other side
// streams straight from global/L2 in tiles of 32 rows — one row per lane for the dot products (each lane walks its
// row in 4/8-byte words; a segment-head is <= 15 KB and stays in L1), one 128-byte line per warp for the
// accumulations.  Softmax is online (running max / sum in registers) with warp-shuffle reductions.
// head_dim even and <= 256 (242 here); fp32 math; I/O fp32 or bf16.
#include "common.cuh"

namespace nlv {
namespace {

constexpr int QB = 16;     // rows per work item
constexpr int KT = 32;     // rows per streamed tile (one per lane)
constexpr int NW = 4;      // words (element pairs) owned per lane: word lane + 32*i  -> head_dim <= 256
constexpr int THREADS = 128;

struct AttnArgs {
  const void *q, *k, *v;   // [rows, *] with row strides ldq/ldk/ldv (elements), head h at column h*hd
  int ldq, ldk, ldv;
  int hd, heads;
  float scale;
  const int4* work;        // {segment first row, segment length, first row of this item relative to segment, unused}
};

// two consecutive elements as float2 (word index w within a row that starts at `row`)
template <typename T> __device__ __forceinline__ float2 ld2(const T* row, int w);
template <> __device__ __forceinline__ float2 ld2<float>(const float* row, int w) {
  return *reinterpret_cast<const float2*>(row + 2 * w);
}
template <> __device__ __forceinline__ float2 ld2<__nv_bfloat16>(const __nv_bfloat16* row, int w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + 2 * w));
}
template <typename T> __device__ __forceinline__ void st2(T* row, int w, float a, float b);
template <> __device__ __forceinline__ void st2<float>(float* row, int w, float a, float b) {
  *reinterpret_cast<float2*>(row + 2 * w) = make_float2(a, b);
}
template <> __device__ __forceinline__ void st2<__nv_bfloat16>(__nv_bfloat16* row, int w, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(row + 2 * w) = __floats2bfloat162_rn(a, b);
}

// stage `nrows` rows of one head transposed into t[d][QB] (fp32, times mul); missing rows are zero
template <typename T>
__device__ __forceinline__ void stage_t(const T* base, int ld, long long row0, int nrows, int col0, int hd, float mul, float* t) {
  // lane = (row r, element e of the pair): shared-memory writes of a warp are 32 consecutive floats (conflict-free);
  // the 16 rows' lines stay in L1 across the hd/2 iterations
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwords = hd >> 1;
  const int r = lane & (QB - 1), e = lane >> 4;
  const T* row = base + (size_t)(row0 + r) * ld + col0 + e;
  for (int w = warp; w < nwords; w += THREADS / 32) {
    float x = 0.f;
    if (r < nrows) x = (float)row[2 * w];
    t[(2 * w + e) * QB + r] = x * mul;
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(THREADS)
attn_fwd_kernel(AttnArgs a, TO* __restrict__ o, int ldo, float* __restrict__ lse) {
  extern __shared__ float sm[];
  const int hd = a.hd, nwords = hd >> 1;
  float* Qs = sm;                    // [hd][QB]   (pre-scaled)
  float* Ps = Qs + hd * QB;          // [KT][QB]
  const int4 w = a.work[blockIdx.x];
  const int h = blockIdx.y, col0 = h * hd;
  const long long seg0 = w.x;
  const int L = w.y, q0 = w.z;
  const int nq = min(QB, L - q0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TI* Q = reinterpret_cast<const TI*>(a.q);
  const TI* K = reinterpret_cast<const TI*>(a.k);
  const TI* V = reinterpret_cast<const TI*>(a.v);

  stage_t<TI>(Q, a.ldq, seg0 + q0, nq, col0, hd, a.scale, Qs);
  __syncthreads();
  if (warp * 4 >= nq) return;   // this warp's four rows are all beyond the segment (no block-wide sync follows)
  float m[4], l[4], acc[4][2 * NW];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    m[qi] = -INFINITY; l[qi] = 0.f;
#pragma unroll
    for (int i = 0; i < 2 * NW; ++i) acc[qi][i] = 0.f;
  }
  for (int k0 = 0; k0 < L; k0 += KT) {
    const int nk = min(KT, L - k0);
    // scores: G = 1, 2 or 4 lanes per key (short segments would otherwise leave most lanes idle); each lane walks every
    // G-th word of its key's row, partial dot products are combined with log2(G) shuffles
    const int lg = nk <= 8 ? 2 : (nk <= 16 ? 1 : 0);
    const int G = 1 << lg, key = lane >> lg, part = lane & (G - 1);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (key < nk) {
      const TI* kr = K + (size_t)(seg0 + k0 + key) * a.ldk + col0;
#pragma unroll 4
      for (int wd = part; wd < nwords; wd += G) {
        const float2 kv = ld2<TI>(kr, wd);
        const float4 qa = *reinterpret_cast<const float4*>(Qs + (2 * wd) * QB + warp * 4);
        const float4 qb = *reinterpret_cast<const float4*>(Qs + (2 * wd + 1) * QB + warp * 4);
        s[0] = fmaf(qa.x, kv.x, s[0]); s[1] = fmaf(qa.y, kv.x, s[1]); s[2] = fmaf(qa.z, kv.x, s[2]); s[3] = fmaf(qa.w, kv.x, s[3]);
        s[0] = fmaf(qb.x, kv.y, s[0]); s[1] = fmaf(qb.y, kv.y, s[1]); s[2] = fmaf(qb.z, kv.y, s[2]); s[3] = fmaf(qb.w, kv.y, s[3]);
      }
    }
    for (int o = 1; o < G; o <<= 1) {
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) s[qi] += __shfl_xor_sync(0xffffffffu, s[qi], o);
    }
    const bool valid = key < nk;
    float p[4];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const float sv = valid ? s[qi] : -INFINITY;
      const float mn = fmaxf(m[qi], warp_max(sv));
      const float corr = __expf(m[qi] - mn);  // m = -inf on the first tile -> 0
      p[qi] = valid ? __expf(sv - mn) : 0.f;
      l[qi] = l[qi] * corr + warp_sum(part == 0 ? p[qi] : 0.f);
      m[qi] = mn;
#pragma unroll
      for (int i = 0; i < 2 * NW; ++i) acc[qi][i] *= corr;
    }
    __syncwarp();
    if (valid && part == 0) *reinterpret_cast<float4*>(Ps + key * QB + warp * 4) = make_float4(p[0], p[1], p[2], p[3]);
    __syncwarp();
    // PV: lane owns words lane + 32*i of the head; V rows stream from global, one line per warp
    for (int j = 0; j < nk; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(Ps + j * QB + warp * 4);
      const TI* vr = V + (size_t)(seg0 + k0 + j) * a.ldv + col0;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const int wd = lane + 32 * i;
        if (wd < nwords) {
          const float2 vv = ld2<TI>(vr, wd);
          acc[0][2 * i] = fmaf(p4.x, vv.x, acc[0][2 * i]); acc[0][2 * i + 1] = fmaf(p4.x, vv.y, acc[0][2 * i + 1]);
          acc[1][2 * i] = fmaf(p4.y, vv.x, acc[1][2 * i]); acc[1][2 * i + 1] = fmaf(p4.y, vv.y, acc[1][2 * i + 1]);
          acc[2][2 * i] = fmaf(p4.z, vv.x, acc[2][2 * i]); acc[2][2 * i + 1] = fmaf(p4.z, vv.y, acc[2][2 * i + 1]);
          acc[3][2 * i] = fmaf(p4.w, vv.x, acc[3][2 * i]); acc[3][2 * i + 1] = fmaf(p4.w, vv.y, acc[3][2 * i + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int qr = warp * 4 + qi;
    if (qr >= nq) continue;
    const long long row = seg0 + q0 + qr;
    const float inv = 1.f / l[qi];
    TO* orow = o + (size_t)row * ldo + col0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int wd = lane + 32 * i;
      if (wd < nwords) st2<TO>(orow, wd, acc[qi][2 * i] * inv, acc[qi][2 * i + 1] * inv);
    }
    if (lane == 0 && lse != nullptr) lse[row * a.heads + h] = m[qi] + __logf(l[qi]);
  }
}

// ------------------------------------------------------------------------------------------
// backward, query side: dQ (and delta = rowsum(dO * O), stored for the key side)
// ------------------------------------------------------------------------------------------
template <typename TI, typename TG>
__global__ void __launch_bounds__(THREADS)
attn_bwd_dq_kernel(AttnArgs a, const TI* __restrict__ o, int ldo, const TG* __restrict__ dout, int lddo,
                   const float* __restrict__ lse, float* __restrict__ delta, TI* __restrict__ dq, int lddq) {
  extern __shared__ float sm[];
  const int hd = a.hd, nwords = hd >> 1;
  float* Qs = sm;                    // [hd][QB] pre-scaled
  float* dOs = Qs + hd * QB;         // [hd][QB]
  float* Ss = dOs + hd * QB;         // [KT][QB]  dS
  const int4 w = a.work[blockIdx.x];
  const int h = blockIdx.y, col0 = h * hd;
  const long long seg0 = w.x;
  const int L = w.y, q0 = w.z;
  const int nq = min(QB, L - q0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TI* Q = reinterpret_cast<const TI*>(a.q);
  const TI* K = reinterpret_cast<const TI*>(a.k);
  const TI* V = reinterpret_cast<const TI*>(a.v);

  stage_t<TI>(Q, a.ldq, seg0 + q0, nq, col0, hd, a.scale, Qs);
  stage_t<TG>(dout, lddo, seg0 + q0, nq, col0, hd, 1.f, dOs);
  float dl[4], ls[4];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int qr = warp * 4 + qi;
    float t = 0.f;
    if (qr < nq) {
      const long long row = seg0 + q0 + qr;
      const TG* gr = dout + (size_t)row * lddo + col0;
      const TI* orow = o + (size_t)row * ldo + col0;
      for (int wd = lane; wd < nwords; wd += 32) {
        const float2 g = ld2<TG>(gr, wd), ov = ld2<TI>(orow, wd);
        t = fmaf(g.x, ov.x, fmaf(g.y, ov.y, t));
      }
    }
    dl[qi] = warp_sum(t);
    ls[qi] = qr < nq ? lse[(seg0 + q0 + qr) * a.heads + h] : 0.f;
    if (lane == 0 && qr < nq) delta[(seg0 + q0 + qr) * a.heads + h] = dl[qi];
  }
  __syncthreads();
  if (warp * 4 >= nq) return;
  float acc[4][2 * NW];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi)
#pragma unroll
    for (int i = 0; i < 2 * NW; ++i) acc[qi][i] = 0.f;
  for (int k0 = 0; k0 < L; k0 += KT) {
    const int nk = min(KT, L - k0);
    const int lg = nk <= 8 ? 2 : (nk <= 16 ? 1 : 0);
    const int G = 1 << lg, key = lane >> lg, part = lane & (G - 1);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
    if (key < nk) {
      const TI* kr = K + (size_t)(seg0 + k0 + key) * a.ldk + col0;
      const TI* vr = V + (size_t)(seg0 + k0 + key) * a.ldv + col0;
#pragma unroll 2
      for (int wd = part; wd < nwords; wd += G) {
        const float2 kv = ld2<TI>(kr, wd), vv = ld2<TI>(vr, wd);
        const float4 qa = *reinterpret_cast<const float4*>(Qs + (2 * wd) * QB + warp * 4);
        const float4 qb = *reinterpret_cast<const float4*>(Qs + (2 * wd + 1) * QB + warp * 4);
        const float4 ga = *reinterpret_cast<const float4*>(dOs + (2 * wd) * QB + warp * 4);
        const float4 gb = *reinterpret_cast<const float4*>(dOs + (2 * wd + 1) * QB + warp * 4);
        s[0] = fmaf(qa.x, kv.x, s[0]); s[1] = fmaf(qa.y, kv.x, s[1]); s[2] = fmaf(qa.z, kv.x, s[2]); s[3] = fmaf(qa.w, kv.x, s[3]);
        s[0] = fmaf(qb.x, kv.y, s[0]); s[1] = fmaf(qb.y, kv.y, s[1]); s[2] = fmaf(qb.z, kv.y, s[2]); s[3] = fmaf(qb.w, kv.y, s[3]);
        dp[0] = fmaf(ga.x, vv.x, dp[0]); dp[1] = fmaf(ga.y, vv.x, dp[1]); dp[2] = fmaf(ga.z, vv.x, dp[2]); dp[3] = fmaf(ga.w, vv.x, dp[3]);
        dp[0] = fmaf(gb.x, vv.y, dp[0]); dp[1] = fmaf(gb.y, vv.y, dp[1]); dp[2] = fmaf(gb.z, vv.y, dp[2]); dp[3] = fmaf(gb.w, vv.y, dp[3]);
      }
    }
    for (int o = 1; o < G; o <<= 1) {
#pragma unroll
      for (int qi = 0; qi < 4; ++qi) {
        s[qi] += __shfl_xor_sync(0xffffffffu, s[qi], o);
        dp[qi] += __shfl_xor_sync(0xffffffffu, dp[qi], o);
      }
    }
    float ds[4];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const float p = key < nk ? __expf(s[qi] - ls[qi]) : 0.f;
      ds[qi] = p * (dp[qi] - dl[qi]);
    }
    __syncwarp();
    if (key < nk && part == 0) *reinterpret_cast<float4*>(Ss + key * QB + warp * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    __syncwarp();
    for (int j = 0; j < nk; ++j) {
      const float4 s4 = *reinterpret_cast<const float4*>(Ss + j * QB + warp * 4);
      const TI* kj = K + (size_t)(seg0 + k0 + j) * a.ldk + col0;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const int wd = lane + 32 * i;
        if (wd < nwords) {
          const float2 kv = ld2<TI>(kj, wd);
          acc[0][2 * i] = fmaf(s4.x, kv.x, acc[0][2 * i]); acc[0][2 * i + 1] = fmaf(s4.x, kv.y, acc[0][2 * i + 1]);
          acc[1][2 * i] = fmaf(s4.y, kv.x, acc[1][2 * i]); acc[1][2 * i + 1] = fmaf(s4.y, kv.y, acc[1][2 * i + 1]);
          acc[2][2 * i] = fmaf(s4.z, kv.x, acc[2][2 * i]); acc[2][2 * i + 1] = fmaf(s4.z, kv.y, acc[2][2 * i + 1]);
          acc[3][2 * i] = fmaf(s4.w, kv.x, acc[3][2 * i]); acc[3][2 * i + 1] = fmaf(s4.w, kv.y, acc[3][2 * i + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int qr = warp * 4 + qi;
    if (qr >= nq) continue;
    TI* drow = dq + (size_t)(seg0 + q0 + qr) * lddq + col0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int wd = lane + 32 * i;
      if (wd < nwords) st2<TI>(drow, wd, acc[qi][2 * i] * a.scale, acc[qi][2 * i + 1] * a.scale);
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward, key side: dK, dV.  The work item's 16 rows are the KEYS; queries stream in tiles of 32.
// ------------------------------------------------------------------------------------------
template <typename TI, typename TG>
__global__ void __launch_bounds__(THREADS)
attn_bwd_dkv_kernel(AttnArgs a, const TG* __restrict__ dout, int lddo, const float* __restrict__ lse,
                    const float* __restrict__ delta, TI* __restrict__ dk, int lddk, TI* __restrict__ dv, int lddv) {
  extern __shared__ float sm[];
  const int hd = a.hd, nwords = hd >> 1;
  float* Kt = sm;                    // [hd][QB]  this item's keys, transposed
  float* Vt = Kt + hd * QB;          // [hd][QB]
  float* Ps = Vt + hd * QB;          // [KT][QB]
  float* Ss = Ps + KT * QB;          // [KT][QB]
  const int4 w = a.work[blockIdx.x];
  const int h = blockIdx.y, col0 = h * hd;
  const long long seg0 = w.x;
  const int L = w.y, k0 = w.z;
  const int nkeys = min(QB, L - k0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TI* Q = reinterpret_cast<const TI*>(a.q);
  const TI* K = reinterpret_cast<const TI*>(a.k);
  const TI* V = reinterpret_cast<const TI*>(a.v);

  stage_t<TI>(K, a.ldk, seg0 + k0, nkeys, col0, hd, 1.f, Kt);
  stage_t<TI>(V, a.ldv, seg0 + k0, nkeys, col0, hd, 1.f, Vt);
  __syncthreads();
  if (warp * 4 >= nkeys) return;
  float accK[4][2 * NW], accV[4][2 * NW];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
#pragma unroll
    for (int i = 0; i < 2 * NW; ++i) { accK[kk][i] = 0.f; accV[kk][i] = 0.f; }
  for (int q0 = 0; q0 < L; q0 += KT) {
    const int nq = min(KT, L - q0);
    // lane = query of the tile; 4 keys of this warp
    const int lg = nq <= 8 ? 2 : (nq <= 16 ? 1 : 0);
    const int G = 1 << lg, qrow = lane >> lg, part = lane & (G - 1);
    float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
    float lq = 0.f, dq_ = 0.f;
    if (qrow < nq) {
      const long long row = seg0 + q0 + qrow;
      const TI* qr = Q + (size_t)row * a.ldq + col0;
      const TG* gr = dout + (size_t)row * lddo + col0;
      lq = lse[row * a.heads + h];
      dq_ = delta[row * a.heads + h];
#pragma unroll 2
      for (int wd = part; wd < nwords; wd += G) {
        const float2 qv = ld2<TI>(qr, wd), gv = ld2<TG>(gr, wd);
        const float4 ka = *reinterpret_cast<const float4*>(Kt + (2 * wd) * QB + warp * 4);
        const float4 kb = *reinterpret_cast<const float4*>(Kt + (2 * wd + 1) * QB + warp * 4);
        const float4 va = *reinterpret_cast<const float4*>(Vt + (2 * wd) * QB + warp * 4);
        const float4 vb = *reinterpret_cast<const float4*>(Vt + (2 * wd + 1) * QB + warp * 4);
        s[0] = fmaf(ka.x, qv.x, s[0]); s[1] = fmaf(ka.y, qv.x, s[1]); s[2] = fmaf(ka.z, qv.x, s[2]); s[3] = fmaf(ka.w, qv.x, s[3]);
        s[0] = fmaf(kb.x, qv.y, s[0]); s[1] = fmaf(kb.y, qv.y, s[1]); s[2] = fmaf(kb.z, qv.y, s[2]); s[3] = fmaf(kb.w, qv.y, s[3]);
        dp[0] = fmaf(va.x, gv.x, dp[0]); dp[1] = fmaf(va.y, gv.x, dp[1]); dp[2] = fmaf(va.z, gv.x, dp[2]); dp[3] = fmaf(va.w, gv.x, dp[3]);
        dp[0] = fmaf(vb.x, gv.y, dp[0]); dp[1] = fmaf(vb.y, gv.y, dp[1]); dp[2] = fmaf(vb.z, gv.y, dp[2]); dp[3] = fmaf(vb.w, gv.y, dp[3]);
      }
    }
    for (int o = 1; o < G; o <<= 1) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        s[kk] += __shfl_xor_sync(0xffffffffu, s[kk], o);
        dp[kk] += __shfl_xor_sync(0xffffffffu, dp[kk], o);
      }
    }
    float p[4], ds[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const bool ok = qrow < nq && (warp * 4 + kk) < nkeys;
      p[kk] = ok ? __expf(s[kk] * a.scale - lq) : 0.f;
      ds[kk] = p[kk] * (dp[kk] - dq_);
    }
    __syncwarp();
    if (qrow < nq && part == 0) {
      *reinterpret_cast<float4*>(Ps + qrow * QB + warp * 4) = make_float4(p[0], p[1], p[2], p[3]);
      *reinterpret_cast<float4*>(Ss + qrow * QB + warp * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    }
    __syncwarp();
    for (int j = 0; j < nq; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(Ps + j * QB + warp * 4);
      const float4 s4 = *reinterpret_cast<const float4*>(Ss + j * QB + warp * 4);
      const long long row = seg0 + q0 + j;
      const TG* gj = dout + (size_t)row * lddo + col0;
      const TI* qj = Q + (size_t)row * a.ldq + col0;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const int wd = lane + 32 * i;
        if (wd < nwords) {
          const float2 gv = ld2<TG>(gj, wd), qv = ld2<TI>(qj, wd);
          accV[0][2 * i] = fmaf(p4.x, gv.x, accV[0][2 * i]); accV[0][2 * i + 1] = fmaf(p4.x, gv.y, accV[0][2 * i + 1]);
          accV[1][2 * i] = fmaf(p4.y, gv.x, accV[1][2 * i]); accV[1][2 * i + 1] = fmaf(p4.y, gv.y, accV[1][2 * i + 1]);
          accV[2][2 * i] = fmaf(p4.z, gv.x, accV[2][2 * i]); accV[2][2 * i + 1] = fmaf(p4.z, gv.y, accV[2][2 * i + 1]);
          accV[3][2 * i] = fmaf(p4.w, gv.x, accV[3][2 * i]); accV[3][2 * i + 1] = fmaf(p4.w, gv.y, accV[3][2 * i + 1]);
          accK[0][2 * i] = fmaf(s4.x, qv.x, accK[0][2 * i]); accK[0][2 * i + 1] = fmaf(s4.x, qv.y, accK[0][2 * i + 1]);
          accK[1][2 * i] = fmaf(s4.y, qv.x, accK[1][2 * i]); accK[1][2 * i + 1] = fmaf(s4.y, qv.y, accK[1][2 * i + 1]);
          accK[2][2 * i] = fmaf(s4.z, qv.x, accK[2][2 * i]); accK[2][2 * i + 1] = fmaf(s4.z, qv.y, accK[2][2 * i + 1]);
          accK[3][2 * i] = fmaf(s4.w, qv.x, accK[3][2 * i]); accK[3][2 * i + 1] = fmaf(s4.w, qv.y, accK[3][2 * i + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int kr = warp * 4 + kk;
    if (kr >= nkeys) continue;
    const long long row = seg0 + k0 + kr;
    TI* dkr = dk + (size_t)row * lddk + col0;
    TI* dvr = dv + (size_t)row * lddv + col0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int wd = lane + 32 * i;
      if (wd < nwords) {
        st2<TI>(dkr, wd, accK[kk][2 * i] * a.scale, accK[kk][2 * i + 1] * a.scale);
        st2<TI>(dvr, wd, accV[kk][2 * i], accV[kk][2 * i + 1]);
      }
    }
  }
}

size_t fwd_smem(int hd) { return sizeof(float) * ((size_t)hd * QB + KT * QB); }
size_t dq_smem(int hd) { return sizeof(float) * (2 * (size_t)hd * QB + KT * QB); }
size_t dkv_smem(int hd) { return sizeof(float) * (2 * (size_t)hd * QB + 2 * KT * QB); }

int check_common(int hd, int heads, int n_work, int ld_all_even) {
  NLV_CHECK_ARG(hd > 0 && hd <= 64 * NW && (hd & 1) == 0, "attention: head_dim=%d unsupported (even, max %d)", hd, 64 * NW);
  NLV_CHECK_ARG(heads > 0 && heads <= 65535 && n_work >= 0, "attention: bad sizes");
  NLV_CHECK_ARG(ld_all_even, "attention: row strides must be even (vector access)");
  return NLV_OK;
}

template <typename K> int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) NLV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return NLV_OK;
}

}  // namespace
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
  int rc = check_common(hd, heads, n_work, ((ldq | ldk | ldv | ldo) & 1) == 0);
  if (rc != NLV_OK) return rc;
  if (n_work == 0) return NLV_OK;
  NLV_CHECK_ARG(q && k && v && work && o, "attn_fwd: null pointer");
  AttnArgs a{q, k, v, ldq, ldk, ldv, hd, heads, scale, (const int4*)work};
  const size_t smem = fwd_smem(hd);
  const dim3 grid(n_work, heads);
#define FWD(TI, TO)                                                                    \
  do {                                                                                 \
    rc = set_smem(attn_fwd_kernel<TI, TO>, smem);                                      \
    if (rc != NLV_OK) return rc;                                                       \
    attn_fwd_kernel<TI, TO><<<grid, THREADS, smem, STREAM>>>(a, (TO*)o, ldo, lse);     \
  } while (0)
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
  int rc = check_common(hd, heads, n_work, ((ldq | ldk | ldv | ldo | lddo | lddq | lddk | lddv) & 1) == 0);
  if (rc != NLV_OK) return rc;
  if (n_work == 0) return NLV_OK;
  NLV_CHECK_ARG(q && k && v && work && o && dout && lse && delta && dq && dk && dv, "attn_bwd: null pointer");
  NLV_CHECK_ARG(in_dtype == o_dtype && in_dtype == dqkv_dtype, "attn_bwd: q/k/v, o and dq/dk/dv must share one dtype");
  AttnArgs a{q, k, v, ldq, ldk, ldv, hd, heads, scale, (const int4*)work};
  const size_t s1 = dq_smem(hd), s2 = dkv_smem(hd);
  const dim3 grid(n_work, heads);
#define BWD(TI, TG)                                                                                                         \
  do {                                                                                                                      \
    rc = set_smem(attn_bwd_dq_kernel<TI, TG>, s1);                                                                          \
    if (rc != NLV_OK) return rc;                                                                                            \
    rc = set_smem(attn_bwd_dkv_kernel<TI, TG>, s2);                                                                         \
    if (rc != NLV_OK) return rc;                                                                                            \
    attn_bwd_dq_kernel<TI, TG><<<grid, THREADS, s1, STREAM>>>(a, (const TI*)o, ldo, (const TG*)dout, lddo, lse, delta,      \
                                                             (TI*)dq, lddq);                                                \
    NLV_CHECK_LAUNCH();                                                                                                     \
    attn_bwd_dkv_kernel<TI, TG><<<grid, THREADS, s2, STREAM>>>(a, (const TG*)dout, lddo, lse, delta, (TI*)dk, lddk,         \
                                                              (TI*)dv, lddv);                                               \
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
