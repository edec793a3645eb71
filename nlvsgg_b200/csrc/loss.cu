// Relation-head activations, fused loss forward+backward, and the fused AdamW step.
//
// heads   : lib/sttran.py:404-409 (attention logits stay raw, spatial/contacting get a sigmoid)
// losses  : tools/train_STTran.py:169-189 — CrossEntropyLoss (objects, attention) and BCELoss (spatial, contacting),
//           each mean-reduced per video; per-row weights carry 1/(rows in that video's loss) (x 1/videos in a batch)
// AdamW   : lib/AdamW.py:52-114 — decoupled decay applied BEFORE the moment update (:69), bias-corrected step,
//           preceded by clip_grad_norm_(max_norm=5) (tools/train_STTran.py:193)
#include "common.cuh"

namespace nlv {
namespace {

__global__ void heads_activation_kernel(const float* __restrict__ logits, long long R, float* __restrict__ att,
                                        float* __restrict__ spa, float* __restrict__ con) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * 26) return;
  const long long r = i / 26;
  const int c = (int)(i - r * 26);
  const float z = logits[i];
  if (c < 3) att[r * 3 + c] = z;
  else if (c < 9) spa[r * 6 + (c - 3)] = 1.f / (1.f + expf(-z));
  else con[r * 17 + (c - 9)] = 1.f / (1.f + expf(-z));
}

// one warp per row; C <= 64 classes.  label < 0 or weight == 0 -> row ignored (zero gradient).
__global__ void ce_loss_kernel(const float* __restrict__ logits, int ld, int C, const long long* __restrict__ labels,
                               const float* __restrict__ wrow, long long rows, float* __restrict__ loss,
                               float* __restrict__ dlogits, int ldd) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (r >= rows) return;
  const long long lab = labels[r];
  const float w = wrow[r];
  const float z0 = lane < C ? logits[r * ld + lane] : -INFINITY;
  const float z1 = lane + 32 < C ? logits[r * ld + lane + 32] : -INFINITY;
  const float m = warp_max(fmaxf(z0, z1));
  const float e0 = lane < C ? expf(z0 - m) : 0.f, e1 = lane + 32 < C ? expf(z1 - m) : 0.f;
  const float s = warp_sum(e0 + e1);
  const bool active = lab >= 0 && w != 0.f;
  if (dlogits) {
    if (lane < C) dlogits[r * ldd + lane] = active ? w * (e0 / s - (lane == lab ? 1.f : 0.f)) : 0.f;
    if (lane + 32 < C) dlogits[r * ldd + lane + 32] = active ? w * (e1 / s - (lane + 32 == lab ? 1.f : 0.f)) : 0.f;
  }
  if (active && lane == 0) atomicAdd(loss, w * (m + logf(s) - logits[r * ld + lab]));
}

// BCELoss(sigmoid(z), y): y given as a bit mask per row.  weight = 1/(rows*C) for the video (0 -> ignored row)
__global__ void bce_loss_kernel(const float* __restrict__ logits, int ld, int C, const unsigned* __restrict__ bits,
                                const float* __restrict__ wrow, long long rows, float* __restrict__ loss,
                                float* __restrict__ dlogits, int ldd) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float contrib = 0.f;
  if (i < rows * C) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float w = wrow[r];
    const float z = logits[r * ld + c];
    const float p = 1.f / (1.f + expf(-z));
    const float y = (bits[r] >> c) & 1u ? 1.f : 0.f;
    if (w != 0.f) {
      const float lp = fmaxf(logf(p), -100.f), lq = fmaxf(logf(1.f - p), -100.f);
      contrib = -w * (y * lp + (1.f - y) * lq);
    }
    if (dlogits) {
      // ATen: grad_p = (p - y) / max((1-p)*p, 1e-12); then sigmoid': p*(1-p)
      const float gp = (p - y) / fmaxf((1.f - p) * p, 1e-12f);
      dlogits[r * ldd + c] = w != 0.f ? w * gp * (p * (1.f - p)) : 0.f;
    }
  }
  contrib = warp_sum(contrib);
  if ((threadIdx.x & 31) == 0 && contrib != 0.f) atomicAdd(loss, contrib);
}

__global__ void sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    acc += v * v;
  }
  acc = warp_sum(acc);
  __shared__ float red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) atomicAdd(out, t);
  }
}

// Fused clip + AdamW over one flat parameter buffer.  total_sq: device scalar holding sum of squared gradients of
// ALL parameters (clip coefficient = min(1, max_norm / (sqrt(total_sq) + 1e-6)), as clip_grad_norm_ computes it).
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                             const float* __restrict__ total_sq, float max_norm, __nv_bfloat16* __restrict__ p_bf16) {
  float coef = 1.f;
  if (total_sq != nullptr && max_norm > 0.f) coef = fminf(1.f, max_norm / (sqrtf(*total_sq) + 1e-6f));
  const float step_size = lr * bc2_sqrt / bc1;
  auto update = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= coef;
    pi *= (1.f - lr * wd);
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) + eps;
    pi += -step_size * (mi / denom);
  };
  // 16-byte path: four elements per thread and the four streams' loads issued together (the kernel is a pure 28-byte-per-element
  // stream); same arithmetic per element as the scalar tail
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(p_bf16) & 7) == 0;
  const long long n4 = vec ? n >> 2 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<const float4*>(p)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<const float4*>(m)[i], vv = reinterpret_cast<const float4*>(v)[i];
    update(pp.x, gg.x, mm.x, vv.x); update(pp.y, gg.y, mm.y, vv.y); update(pp.z, gg.z, mm.z, vv.z); update(pp.w, gg.w, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    if (p_bf16) {
      uint2 o;
      *reinterpret_cast<__nv_bfloat162*>(&o.x) = __floats2bfloat162_rn(pp.x, pp.y);
      *reinterpret_cast<__nv_bfloat162*>(&o.y) = __floats2bfloat162_rn(pp.z, pp.w);
      reinterpret_cast<uint2*>(p_bf16)[i] = o;
    }
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    update(pi, g[i], mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pi);
  }
}

// Same update driven by device-side state, so a step can be skipped without a host sync (tools/train_STTran.py:191
// check_valid_iter: NaN loss / NaN outputs -> `continue` before backward and optimizer.step()):
//   state[0] = optimiser steps applied so far (bias corrections use state[0] + 1), state[1] = steps skipped;
//   skipped when *skip_flag != 0 or the gradient norm is not finite; grad_scale folds the 1/world of the gradient
//   all-reduce (total_sq is the squared norm of the SUMMED gradients).
__global__ void adamw_state_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                   long long n, float lr, float b1, float b2, float eps, float wd, const int* __restrict__ state,
                                   const float* __restrict__ total_sq, const int* __restrict__ skip_flag, float grad_scale,
                                   float max_norm, __nv_bfloat16* __restrict__ p_bf16) {
  const float tsq = total_sq != nullptr ? *total_sq : 0.f;
  if ((skip_flag != nullptr && *skip_flag != 0) || !isfinite(tsq)) return;
  float coef = grad_scale;
  if (total_sq != nullptr && max_norm > 0.f) coef *= fminf(1.f, max_norm / (sqrtf(tsq) * grad_scale + 1e-6f));
  const int step = state[0] + 1;
  const float bc1 = (float)(1.0 - pow((double)b1, (double)step));
  const float bc2s = (float)sqrt(1.0 - pow((double)b2, (double)step));
  const float step_size = lr * bc2s / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float denom = sqrtf(vi) + eps;
    pi += -step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (p_bf16) p_bf16[i] = __float2bfloat16_rn(pi);
  }
}

__global__ void adamw_finish_kernel(int* state, const float* total_sq, const int* skip_flag) {
  const float tsq = total_sq != nullptr ? *total_sq : 0.f;
  if ((skip_flag != nullptr && *skip_flag != 0) || !isfinite(tsq)) state[1] += 1;
  else state[0] += 1;
}

__global__ void flag_nonfinite_kernel(const float* x, int n, int* flag) {
  int bad = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) bad |= !isfinite(x[i]);
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) flag[0] = bad ? 1 : 0;
}

}  // namespace
}  // namespace nlv

using namespace nlv;
#define STREAM ((cudaStream_t)stream)

extern "C" {

int nlv_heads_activation(const float* logits, long long r, float* att, float* spa, float* con, void* stream) {
  NLV_CHECK_ARG(r >= 0, "heads_activation: bad sizes");
  if (r == 0) return NLV_OK;
  NLV_CHECK_ARG(logits && att && spa && con, "heads_activation: null pointer");
  heads_activation_kernel<<<cdiv(r * 26, 256), 256, 0, STREAM>>>(logits, r, att, spa, con);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* loss (device scalar) is ACCUMULATED into.  dlogits may be null (loss only). */
int nlv_ce_loss(const float* logits, int ld, int c, const long long* labels, const float* row_weight, long long rows,
                float* loss, float* dlogits, int ldd, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && c >= 1 && c <= 64, "ce_loss: C=%d unsupported", c);
  if (rows == 0) return NLV_OK;
  NLV_CHECK_ARG(logits && labels && row_weight && loss, "ce_loss: null pointer");
  ce_loss_kernel<<<cdiv(rows, 4), 128, 0, STREAM>>>(logits, ld, c, labels, row_weight, rows, loss, dlogits, ldd);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_bce_sigmoid_loss(const float* logits, int ld, int c, const unsigned* label_bits, const float* row_weight,
                         long long rows, float* loss, float* dlogits, int ldd, void* stream) {
  NLV_CHECK_ARG(rows >= 0 && c >= 1 && c <= 32, "bce_loss: C=%d unsupported", c);
  if (rows == 0) return NLV_OK;
  NLV_CHECK_ARG(logits && label_bits && row_weight && loss, "bce_loss: null pointer");
  bce_loss_kernel<<<cdiv(rows * c, 256), 256, 0, STREAM>>>(logits, ld, c, label_bits, row_weight, rows, loss, dlogits, ldd);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

/* out (device scalar) += sum x^2 */
int nlv_sumsq(const float* x, long long n, float* out, void* stream) {
  NLV_CHECK_ARG(n >= 0, "sumsq: bad size");
  if (n == 0) return NLV_OK;
  int blocks = cdiv(n, 256 * 8);
  const int cap = 8 * sm_count();
  if (blocks > cap) blocks = cap;
  sumsq_kernel<<<blocks, 256, 0, STREAM>>>(x, n, out);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, const float* total_sq, float max_norm, void* p_bf16, void* stream) {
  NLV_CHECK_ARG(n >= 0 && step >= 1, "adamw: bad arguments");
  if (n == 0) return NLV_OK;
  NLV_CHECK_ARG(p && g && m && v, "adamw: null pointer");
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  int blocks = cdiv(n, 256 * 4);
  const int cap = 16 * sm_count();
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<blocks, 256, 0, STREAM>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, total_sq, max_norm,
                                          (__nv_bfloat16*)p_bf16);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_adamw_step_state(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                         float weight_decay, const int* state, const float* total_sq, const int* skip_flag, float grad_scale,
                         float max_norm, void* p_bf16, void* stream) {
  NLV_CHECK_ARG(n >= 0, "adamw_state: bad arguments");
  if (n == 0) return NLV_OK;
  NLV_CHECK_ARG(p && g && m && v && state, "adamw_state: null pointer");
  int blocks = cdiv(n, 256 * 4);
  const int cap = 16 * sm_count();
  if (blocks > cap) blocks = cap;
  adamw_state_kernel<<<blocks, 256, 0, STREAM>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, state, total_sq, skip_flag,
                                                grad_scale, max_norm, (__nv_bfloat16*)p_bf16);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_adamw_finish(int* state, const float* total_sq, const int* skip_flag, void* stream) {
  NLV_CHECK_ARG(state != nullptr, "adamw_finish: null state");
  adamw_finish_kernel<<<1, 1, 0, STREAM>>>(state, total_sq, skip_flag);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_flag_nonfinite(const float* x, int n, int* flag, void* stream) {
  NLV_CHECK_ARG(x && flag && n >= 0, "flag_nonfinite: bad arguments");
  flag_nonfinite_kernel<<<1, 128, 0, STREAM>>>(x, n, flag);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}
}
