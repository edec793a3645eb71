// nlv_gemm dispatcher + the exact-fp32 SIMT GEMM (parity mode, tiny / odd shapes).
#include "common.cuh"
#include "philox.cuh"

namespace nlv {

int gemm_tc(const nlv_gemm_args& g, cudaStream_t stream);  // gemm_tc.cu
int gemm_tc_conv3x3(const void* x, long long pairs, int c, const void* b, int ldb, int n, int flip, void* d, int d_dtype, int ldd,
                    const float* bias, int relu, cudaStream_t stream);

namespace {

// 64x64 output tile, 16-deep k slices, 256 threads, 4x4 outputs per thread.  fp32 throughout:
// products are accumulated in k order with FMA, which is what a CPU sgemm does up to reassociation.
constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float ld_elem(const void* p, int dtype, size_t i) { return ld_as_float(p, dtype, i); }

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const void* __restrict__ A, const void* __restrict__ B, void* D, const float* __restrict__ bias,
                 const void* residual, int M, int N, int K, int lda, int ldb, int ldd, int ldr, int ab_dtype,
                 int d_dtype, int r_dtype, int relu, int k_per_split, const void* gate, int ldg, int gate_dtype, float gate_scale,
                 const DropCfg drop) {
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4x4 (strided by 16)
  float acc[4][4] = {};
  // split-K: blockIdx.z owns k in [kz0, kz1); partial sums are atomically added into a zeroed fp32 D
  const int kz0 = blockIdx.z * k_per_split;
  const int kz1 = min(K, kz0 + k_per_split);
  const bool split = gridDim.z > 1;
  for (int k0 = kz0; k0 < kz1; k0 += TK) {
    // load A tile: TM x TK
    for (int i = tid; i < TM * TK; i += 256) {
      int mm, kk;
      if (A_MN) { mm = i % TM; kk = i / TM; } else { kk = i % TK; mm = i / TK; }
      const int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < M && gk < kz1) v = A_MN ? ld_elem(A, ab_dtype, (size_t)gk * lda + gm) : ld_elem(A, ab_dtype, (size_t)gm * lda + gk);
      As[kk][mm] = v;
    }
    for (int i = tid; i < TN * TK; i += 256) {
      int nn, kk;
      if (B_MN) { nn = i % TN; kk = i / TN; } else { kk = i % TK; nn = i / TK; }
      const int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < N && gk < kz1) v = B_MN ? ld_elem(B, ab_dtype, (size_t)gk * ldb + gn) : ld_elem(B, ab_dtype, (size_t)gn * ldb + gk);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty + 16 * i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx + 16 * j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (split) { atomicAdd(reinterpret_cast<float*>(D) + (size_t)gm * ldd + gn, v); continue; }
      if (bias != nullptr) v += bias[gn];
      if (relu) v = fmaxf(v, 0.f);
      if (drop.thr16 != 0u) v = ((keep8_matrix(drop, gm, gn >> 3, (N + 7) >> 3) >> (gn & 7)) & 1u) ? v * drop.scale : 0.f;
      if (gate != nullptr) v = (ld_as_float(gate, gate_dtype, (size_t)gm * ldg + gn) > 0.f) ? v * gate_scale : 0.f;
      if (residual != nullptr) v += ld_as_float(residual, r_dtype, (size_t)gm * ldr + gn);
      st_from_float(D, d_dtype, (size_t)gm * ldd + gn, v);
    }
  }
}

int gemm_simt(const nlv_gemm_args& g, cudaStream_t s) {
  dim3 grid(cdiv(g.n, TN), cdiv(g.m, TM), 1);
  NLV_CHECK_ARG(grid.y <= 65535, "gemm(simt): m=%d too large", g.m);
  int k_per_split = g.k;
  const int tiles = (int)(grid.x * grid.y);
  if (tiles < 2 * sm_count() && g.k >= 2048 && g.d_dtype == NLV_F32 && !g.bias && !g.residual && !g.relu && !g.gate && g.drop.thr16 == 0u) {
    int want = cdiv(4 * sm_count(), tiles);
    if (want > g.k / 256) want = g.k / 256;
    if (want > 1) {
      k_per_split = cdiv(cdiv(g.k, want), TK) * TK;
      grid.z = cdiv(g.k, k_per_split);
      { int zrc = zero_fill(reinterpret_cast<float*>(g.d), g.m, g.n, g.ldd, s); if (zrc != NLV_OK) return zrc; }
    }
  }
  DropCfg dc;
  dc.thr16 = g.drop.thr16; dc.scale = g.drop.scale; dc.seed_lo = g.drop.seed_lo; dc.seed_hi = g.drop.seed_hi; dc.stream = g.drop.stream;
#define LAUNCH(AM, BM)                                                                                          \
  gemm_simt_kernel<AM, BM><<<grid, 256, 0, s>>>(g.a, g.b, g.d, g.bias, g.residual, g.m, g.n, g.k, g.lda, g.ldb, \
                                                g.ldd, g.ldr, g.ab_dtype, g.d_dtype, g.r_dtype, g.relu, k_per_split, g.gate, g.ldg, g.gate_dtype, \
                                                g.gate_scale == 0.f ? 1.f : g.gate_scale, dc)
  if (g.a_major == NLV_MAJOR_K && g.b_major == NLV_MAJOR_K) LAUNCH(false, false);
  else if (g.a_major == NLV_MAJOR_K) LAUNCH(false, true);
  else if (g.b_major == NLV_MAJOR_K) LAUNCH(true, false);
  else LAUNCH(true, true);
#undef LAUNCH
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

}  // namespace
}  // namespace nlv

extern "C" int nlv_gemm(const nlv_gemm_args* g, void* stream) {
  using namespace nlv;
  NLV_CHECK_ARG(g != nullptr, "gemm: null args");
  NLV_CHECK_ARG(g->m >= 0 && g->n >= 0 && g->k >= 1, "gemm: bad shape m=%d n=%d k=%d", g->m, g->n, g->k);
  NLV_CHECK_ARG(g->a && g->b && g->d, "gemm: null operand");
  NLV_CHECK_ARG(g->a_major == NLV_MAJOR_K || g->a_major == NLV_MAJOR_MN, "gemm: bad a_major");
  NLV_CHECK_ARG(g->b_major == NLV_MAJOR_K || g->b_major == NLV_MAJOR_MN, "gemm: bad b_major");
  if (g->m == 0 || g->n == 0) return NLV_OK;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // backend: bit 8 of ab_dtype forces the SIMT kernel on bf16 operands (used by tests)
  const int dt = g->ab_dtype & 0xff;
  const bool force_simt = (g->ab_dtype & 0x100) != 0;
  if (dt == NLV_BF16 && !force_simt) return gemm_tc(*g, s);
  nlv_gemm_args t = *g;
  t.ab_dtype = dt;
  return gemm_simt(t, s);
}

/* Data gradient of the 3x3 / pad 1 convolution of the mask branch (lib/sttran.py:342) as ONE implicit GEMM on the tcgen05 kernel:
 * dx[r*49, c_in] = sum over taps, c_out of dy[r, y + 1 - ky, x + 1 - kx, c_out] * w[c_out, c_in, ky, kx], with
 * dy bf16 NHWC [r,7,7,c_out] and wt bf16 [c_in, 9 * c_out] (k = (ky*3 + kx) * c_out + channel: the [c_out, c_in*9] weight
 * transposed).  Replaces the [r*49, 9*c_in] column-gradient product + nlv_col2im_3x3. */
extern "C" int nlv_conv3x3_dgrad(const void* dy, long long r, int c_out, const void* wt, int c_in, void* dx, int dx_dtype, void* stream) {
  NLV_CHECK_ARG(dy && wt && dx, "conv3x3_dgrad: null pointer");
  return nlv::gemm_tc_conv3x3(dy, r, c_out, wt, 9 * c_out, c_in, 1, dx, dx_dtype, c_in, nullptr, 0, (cudaStream_t)stream);
}
