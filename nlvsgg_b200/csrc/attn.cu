// Fused variable-length multi-head attention over contiguous token segments (forward + backward).
//
// One launch covers every segment of the batch: a segment is a frame (spatial encoder,
// lib/transformer.py:20-22), a 2-frame sliding window (temporal decoder, :49-52, lib/transformer_wk.py:163-171)
// or a per-class sequence (lib/dsg_detr.py:545-559).  Segments are unpadded, so no key_padding_mask exists:
// a query only ever sees the keys of its own segment ("bool masking" semantics).
//
// Work decomposition: a work item = 16 consecutive rows of one segment; grid = (work items, heads).
// 4 warps per CTA, 4 rows per warp.  Keys/values stream through shared memory in tiles of 32 rows with an
// online (running max / running sum) softmax held in registers; scores are reduced with warp shuffles.
// head_dim <= 256 (242 here); fp32 math; I/O fp32 or bf16.
#include "common.cuh"

namespace nlv {
namespace {

constexpr int QB = 16;     // rows per work item
constexpr int KT = 32;     // rows per streamed tile (one per lane)
constexpr int NI = 8;      // head-dim columns owned per lane (lane + 32*i)
constexpr int THREADS = 128;

struct AttnArgs {
  const void *q, *k, *v;   // [rows, *] with row strides ldq/ldk/ldv, head h at column h*hd
  int ldq, ldk, ldv, in_dtype;
  int hd, heads;
  float scale;
  const int4* work;        // {segment first row, segment length, first row of this item relative to segment, unused}
};

// load `nrows` rows (starting at global row `row0`) of one head into smem tile[r][hdp], zero-fill missing rows
__device__ __forceinline__ void load_tile(const void* src, int dt, int ld, long long row0, int nrows, int col0, int hd,
                                          int hdp, float mul, float* tile, int tile_rows) {
  for (int i = threadIdx.x; i < tile_rows * hd; i += THREADS) {
    const int r = i / hd, d = i - r * hd;
    tile[r * hdp + d] = r < nrows ? mul * ld_as_float(src, dt, (size_t)(row0 + r) * ld + col0 + d) : 0.f;
  }
}
// load rows transposed: t[d][QB]
__device__ __forceinline__ void load_tile_t(const void* src, int dt, int ld, long long row0, int nrows, int col0, int hd,
                                            float mul, float* t) {
  for (int i = threadIdx.x; i < QB * hd; i += THREADS) {
    const int r = i / hd, d = i - r * hd;
    t[d * QB + r] = r < nrows ? mul * ld_as_float(src, dt, (size_t)(row0 + r) * ld + col0 + d) : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
attn_fwd_kernel(AttnArgs a, void* __restrict__ o, int ldo, int o_dtype, float* __restrict__ lse) {
  extern __shared__ float sm[];
  const int hd = a.hd, hdp = hd | 1;
  float* Qs = sm;                    // [hd][QB]   (pre-scaled)
  float* Ks = Qs + hd * QB;          // [KT][hdp]
  float* Vs = Ks + KT * hdp;         // [KT][hdp]
  float* Ps = Vs + KT * hdp;         // [KT][QB]
  const int4 w = a.work[blockIdx.x];
  const int h = blockIdx.y, col0 = h * hd;
  const long long seg0 = w.x;
  const int L = w.y, q0 = w.z;
  const int nq = min(QB, L - q0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_tile_t(a.q, a.in_dtype, a.ldq, seg0 + q0, nq, col0, hd, a.scale, Qs);
  float m[4], l[4], acc[4][NI];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    m[qi] = -INFINITY; l[qi] = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[qi][i] = 0.f;
  }
  for (int k0 = 0; k0 < L; k0 += KT) {
    const int nk = min(KT, L - k0);
    __syncthreads();  // previous tile fully consumed (and Qs visible on the first pass)
    load_tile(a.k, a.in_dtype, a.ldk, seg0 + k0, nk, col0, hd, hdp, 1.f, Ks, KT);
    load_tile(a.v, a.in_dtype, a.ldv, seg0 + k0, nk, col0, hd, hdp, 1.f, Vs, KT);
    __syncthreads();
    // scores: lane = key, 4 queries of this warp
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const float* kr = Ks + lane * hdp;
    for (int d = 0; d < hd; ++d) {
      const float kv = kr[d];
      const float4 q4 = *reinterpret_cast<const float4*>(Qs + d * QB + warp * 4);
      s[0] = fmaf(q4.x, kv, s[0]); s[1] = fmaf(q4.y, kv, s[1]); s[2] = fmaf(q4.z, kv, s[2]); s[3] = fmaf(q4.w, kv, s[3]);
    }
    float p[4];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const float sv = lane < nk ? s[qi] : -INFINITY;
      const float mn = fmaxf(m[qi], warp_max(sv));
      const float corr = __expf(m[qi] - mn);  // m = -inf on the first tile -> 0
      p[qi] = lane < nk ? __expf(sv - mn) : 0.f;
      l[qi] = l[qi] * corr + warp_sum(p[qi]);
      m[qi] = mn;
#pragma unroll
      for (int i = 0; i < NI; ++i) acc[qi][i] *= corr;
    }
    *reinterpret_cast<float4*>(Ps + lane * QB + warp * 4) = make_float4(p[0], p[1], p[2], p[3]);
    __syncwarp();
    // PV: lane owns columns lane + 32*i
    for (int j = 0; j < nk; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(Ps + j * QB + warp * 4);
      const float* vr = Vs + j * hdp;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int d = lane + 32 * i;
        const float vv = d < hd ? vr[d] : 0.f;
        acc[0][i] = fmaf(p4.x, vv, acc[0][i]); acc[1][i] = fmaf(p4.y, vv, acc[1][i]);
        acc[2][i] = fmaf(p4.z, vv, acc[2][i]); acc[3][i] = fmaf(p4.w, vv, acc[3][i]);
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int qr = warp * 4 + qi;
    if (qr >= nq) continue;
    const long long row = seg0 + q0 + qr;
    const float inv = 1.f / l[qi];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int d = lane + 32 * i;
      if (d < hd) st_from_float(o, o_dtype, (size_t)row * ldo + col0 + d, acc[qi][i] * inv);
    }
    if (lane == 0 && lse != nullptr) lse[row * a.heads + h] = m[qi] + __logf(l[qi]);
  }
}

// ------------------------------------------------------------------------------------------
// backward, query side: dQ (and delta = rowsum(dO * O), stored for the key side)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
attn_bwd_dq_kernel(AttnArgs a, const void* __restrict__ o, int ldo, int o_dtype, const void* __restrict__ dout, int lddo,
                   int do_dtype, const float* __restrict__ lse, float* __restrict__ delta, void* __restrict__ dq, int lddq,
                   int dq_dtype) {
  extern __shared__ float sm[];
  const int hd = a.hd, hdp = hd | 1;
  float* Qs = sm;                    // [hd][QB] pre-scaled
  float* dOs = Qs + hd * QB;         // [hd][QB]
  float* Ks = dOs + hd * QB;         // [KT][hdp]
  float* Vs = Ks + KT * hdp;         // [KT][hdp]
  float* Ss = Vs + KT * hdp;         // [KT][QB]  dS
  const int4 w = a.work[blockIdx.x];
  const int h = blockIdx.y, col0 = h * hd;
  const long long seg0 = w.x;
  const int L = w.y, q0 = w.z;
  const int nq = min(QB, L - q0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_tile_t(a.q, a.in_dtype, a.ldq, seg0 + q0, nq, col0, hd, a.scale, Qs);
  load_tile_t(dout, do_dtype, lddo, seg0 + q0, nq, col0, hd, 1.f, dOs);
  // delta and lse for this warp's 4 rows
  float dl[4], ls[4];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int qr = warp * 4 + qi;
    float t = 0.f;
    if (qr < nq) {
      const long long row = seg0 + q0 + qr;
      for (int d = lane; d < hd; d += 32)
        t += ld_as_float(dout, do_dtype, (size_t)row * lddo + col0 + d) * ld_as_float(o, o_dtype, (size_t)row * ldo + col0 + d);
    }
    dl[qi] = warp_sum(t);
    ls[qi] = qr < nq ? lse[(seg0 + q0 + qr) * a.heads + h] : 0.f;
    if (lane == 0 && qr < nq) delta[(seg0 + q0 + qr) * a.heads + h] = dl[qi];
  }
  float acc[4][NI];
#pragma unroll
  for (int qi = 0; qi < 4; ++qi)
#pragma unroll
    for (int i = 0; i < NI; ++i) acc[qi][i] = 0.f;
  for (int k0 = 0; k0 < L; k0 += KT) {
    const int nk = min(KT, L - k0);
    __syncthreads();
    load_tile(a.k, a.in_dtype, a.ldk, seg0 + k0, nk, col0, hd, hdp, 1.f, Ks, KT);
    load_tile(a.v, a.in_dtype, a.ldv, seg0 + k0, nk, col0, hd, hdp, 1.f, Vs, KT);
    __syncthreads();
    float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
    const float* kr = Ks + lane * hdp;
    const float* vr = Vs + lane * hdp;
    for (int d = 0; d < hd; ++d) {
      const float kv = kr[d], vv = vr[d];
      const float4 q4 = *reinterpret_cast<const float4*>(Qs + d * QB + warp * 4);
      const float4 g4 = *reinterpret_cast<const float4*>(dOs + d * QB + warp * 4);
      s[0] = fmaf(q4.x, kv, s[0]); s[1] = fmaf(q4.y, kv, s[1]); s[2] = fmaf(q4.z, kv, s[2]); s[3] = fmaf(q4.w, kv, s[3]);
      dp[0] = fmaf(g4.x, vv, dp[0]); dp[1] = fmaf(g4.y, vv, dp[1]); dp[2] = fmaf(g4.z, vv, dp[2]); dp[3] = fmaf(g4.w, vv, dp[3]);
    }
    float ds[4];
#pragma unroll
    for (int qi = 0; qi < 4; ++qi) {
      const float p = lane < nk ? __expf(s[qi] - ls[qi]) : 0.f;
      ds[qi] = p * (dp[qi] - dl[qi]);
    }
    *reinterpret_cast<float4*>(Ss + lane * QB + warp * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    __syncwarp();
    for (int j = 0; j < nk; ++j) {
      const float4 s4 = *reinterpret_cast<const float4*>(Ss + j * QB + warp * 4);
      const float* kj = Ks + j * hdp;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int d = lane + 32 * i;
        const float kv = d < hd ? kj[d] : 0.f;
        acc[0][i] = fmaf(s4.x, kv, acc[0][i]); acc[1][i] = fmaf(s4.y, kv, acc[1][i]);
        acc[2][i] = fmaf(s4.z, kv, acc[2][i]); acc[3][i] = fmaf(s4.w, kv, acc[3][i]);
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < 4; ++qi) {
    const int qr = warp * 4 + qi;
    if (qr >= nq) continue;
    const long long row = seg0 + q0 + qr;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int d = lane + 32 * i;
      if (d < hd) st_from_float(dq, dq_dtype, (size_t)row * lddq + col0 + d, acc[qi][i] * a.scale);
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward, key side: dK, dV.  The work item's 16 rows are the KEYS; queries stream in tiles of 32.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS)
attn_bwd_dkv_kernel(AttnArgs a, const void* __restrict__ dout, int lddo, int do_dtype, const float* __restrict__ lse,
                    const float* __restrict__ delta, void* __restrict__ dk, int lddk, void* __restrict__ dv, int lddv,
                    int dkv_dtype) {
  extern __shared__ float sm[];
  const int hd = a.hd, hdp = hd | 1;
  float* Kt = sm;                    // [hd][QB]  this item's keys, transposed
  float* Vt = Kt + hd * QB;          // [hd][QB]
  float* Qs = Vt + hd * QB;          // [KT][hdp] query tile (pre-scaled)
  float* dOs = Qs + KT * hdp;        // [KT][hdp]
  float* Ps = dOs + KT * hdp;        // [KT][QB]
  float* Ss = Ps + KT * QB;          // [KT][QB]
  float* stat = Ss + KT * QB;        // [2][KT] lse, delta of the query tile
  const int4 w = a.work[blockIdx.x];
  const int h = blockIdx.y, col0 = h * hd;
  const long long seg0 = w.x;
  const int L = w.y, k0 = w.z;
  const int nkeys = min(QB, L - k0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  load_tile_t(a.k, a.in_dtype, a.ldk, seg0 + k0, nkeys, col0, hd, 1.f, Kt);
  load_tile_t(a.v, a.in_dtype, a.ldv, seg0 + k0, nkeys, col0, hd, 1.f, Vt);
  float accK[4][NI], accV[4][NI];
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
#pragma unroll
    for (int i = 0; i < NI; ++i) { accK[kk][i] = 0.f; accV[kk][i] = 0.f; }
  for (int q0 = 0; q0 < L; q0 += KT) {
    const int nq = min(KT, L - q0);
    __syncthreads();
    load_tile(a.q, a.in_dtype, a.ldq, seg0 + q0, nq, col0, hd, hdp, a.scale, Qs, KT);
    load_tile(dout, do_dtype, lddo, seg0 + q0, nq, col0, hd, hdp, 1.f, dOs, KT);
    if (threadIdx.x < KT) {
      const bool ok = threadIdx.x < nq;
      stat[threadIdx.x] = ok ? lse[(seg0 + q0 + threadIdx.x) * a.heads + h] : 0.f;
      stat[KT + threadIdx.x] = ok ? delta[(seg0 + q0 + threadIdx.x) * a.heads + h] : 0.f;
    }
    __syncthreads();
    // lane = query of the tile; 4 keys of this warp
    float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
    const float* qr = Qs + lane * hdp;
    const float* gr = dOs + lane * hdp;
    for (int d = 0; d < hd; ++d) {
      const float qv = qr[d], gv = gr[d];
      const float4 k4 = *reinterpret_cast<const float4*>(Kt + d * QB + warp * 4);
      const float4 v4 = *reinterpret_cast<const float4*>(Vt + d * QB + warp * 4);
      s[0] = fmaf(k4.x, qv, s[0]); s[1] = fmaf(k4.y, qv, s[1]); s[2] = fmaf(k4.z, qv, s[2]); s[3] = fmaf(k4.w, qv, s[3]);
      dp[0] = fmaf(v4.x, gv, dp[0]); dp[1] = fmaf(v4.y, gv, dp[1]); dp[2] = fmaf(v4.z, gv, dp[2]); dp[3] = fmaf(v4.w, gv, dp[3]);
    }
    float p[4], ds[4];
    const float lq = stat[lane], dq_ = stat[KT + lane];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const bool ok = lane < nq && (warp * 4 + kk) < nkeys;
      p[kk] = ok ? __expf(s[kk] - lq) : 0.f;
      ds[kk] = p[kk] * (dp[kk] - dq_);
    }
    *reinterpret_cast<float4*>(Ps + lane * QB + warp * 4) = make_float4(p[0], p[1], p[2], p[3]);
    *reinterpret_cast<float4*>(Ss + lane * QB + warp * 4) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    __syncwarp();
    for (int j = 0; j < nq; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(Ps + j * QB + warp * 4);
      const float4 s4 = *reinterpret_cast<const float4*>(Ss + j * QB + warp * 4);
      const float* gj = dOs + j * hdp;
      const float* qj = Qs + j * hdp;
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int d = lane + 32 * i;
        const float gv = d < hd ? gj[d] : 0.f;
        const float qv = d < hd ? qj[d] : 0.f;
        accV[0][i] = fmaf(p4.x, gv, accV[0][i]); accV[1][i] = fmaf(p4.y, gv, accV[1][i]);
        accV[2][i] = fmaf(p4.z, gv, accV[2][i]); accV[3][i] = fmaf(p4.w, gv, accV[3][i]);
        accK[0][i] = fmaf(s4.x, qv, accK[0][i]); accK[1][i] = fmaf(s4.y, qv, accK[1][i]);
        accK[2][i] = fmaf(s4.z, qv, accK[2][i]); accK[3][i] = fmaf(s4.w, qv, accK[3][i]);
      }
    }
  }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int kr = warp * 4 + kk;
    if (kr >= nkeys) continue;
    const long long row = seg0 + k0 + kr;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int d = lane + 32 * i;
      if (d < hd) {
        st_from_float(dk, dkv_dtype, (size_t)row * lddk + col0 + d, accK[kk][i]);  // Qs carries the 1/sqrt(hd) scale
        st_from_float(dv, dkv_dtype, (size_t)row * lddv + col0 + d, accV[kk][i]);
      }
    }
  }
}

size_t fwd_smem(int hd) { const int hdp = hd | 1; return sizeof(float) * ((size_t)hd * QB + 2 * KT * hdp + KT * QB); }
size_t dq_smem(int hd) { const int hdp = hd | 1; return sizeof(float) * (2 * (size_t)hd * QB + 2 * KT * hdp + KT * QB); }
size_t dkv_smem(int hd) { const int hdp = hd | 1; return sizeof(float) * (2 * (size_t)hd * QB + 2 * KT * hdp + 2 * KT * QB + 2 * KT); }

int check_common(int hd, int heads, int n_work) {
  NLV_CHECK_ARG(hd > 0 && hd <= 32 * NI, "attention: head_dim=%d unsupported (max %d)", hd, 32 * NI);
  NLV_CHECK_ARG(heads > 0 && heads <= 65535 && n_work >= 0, "attention: bad sizes");
  return NLV_OK;
}

}  // namespace
}  // namespace nlv

using namespace nlv;
#define STREAM ((cudaStream_t)stream)

extern "C" {

/* work: int4[n_work] = {segment first row, segment length, item first row within the segment, 0}; every segment is
 * covered by ceil(len/16) items.  lse: float[rows*heads] (optional, needed for backward). */
int nlv_attn_fwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                 float scale, const void* work, int n_work, void* o, int ldo, int o_dtype, float* lse, void* stream) {
  int rc = check_common(hd, heads, n_work);
  if (rc != NLV_OK) return rc;
  if (n_work == 0) return NLV_OK;
  NLV_CHECK_ARG(q && k && v && work && o, "attn_fwd: null pointer");
  AttnArgs a{q, k, v, ldq, ldk, ldv, in_dtype, hd, heads, scale, (const int4*)work};
  const size_t smem = fwd_smem(hd);
  static size_t set = 0;
  if (smem > set) { NLV_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = smem; }
  attn_fwd_kernel<<<dim3(n_work, heads), THREADS, smem, STREAM>>>(a, o, ldo, o_dtype, lse);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* delta: float[rows*heads] workspace written by the query-side kernel and read by the key-side kernel. */
int nlv_attn_bwd(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int in_dtype, int hd, int heads,
                 float scale, const void* work, int n_work, const void* o, int ldo, int o_dtype, const void* dout, int lddo,
                 int do_dtype, const float* lse, float* delta, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv,
                 int dqkv_dtype, void* stream) {
  int rc = check_common(hd, heads, n_work);
  if (rc != NLV_OK) return rc;
  if (n_work == 0) return NLV_OK;
  NLV_CHECK_ARG(q && k && v && work && o && dout && lse && delta && dq && dk && dv, "attn_bwd: null pointer");
  AttnArgs a{q, k, v, ldq, ldk, ldv, in_dtype, hd, heads, scale, (const int4*)work};
  const size_t s1 = dq_smem(hd), s2 = dkv_smem(hd);
  static size_t set1 = 0, set2 = 0;
  if (s1 > set1) { NLV_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1)); set1 = s1; }
  if (s2 > set2) { NLV_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2)); set2 = s2; }
  attn_bwd_dq_kernel<<<dim3(n_work, heads), THREADS, s1, STREAM>>>(a, o, ldo, o_dtype, dout, lddo, do_dtype, lse, delta, dq, lddq, dqkv_dtype);
  NLV_CHECK_LAUNCH();
  attn_bwd_dkv_kernel<<<dim3(n_work, heads), THREADS, s2, STREAM>>>(a, dout, lddo, do_dtype, lse, delta, dk, lddk, dv, lddv, dqkv_dtype);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
}
