// Shared helpers for the nlv_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nlv_b200.h"

namespace nlv {

// thread-local last error text, returned by nlv_last_error()
void set_error(const char* fmt, ...);

#define NLV_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      nlv::set_error(__VA_ARGS__);          \
      return NLV_ERR_INVALID_ARGUMENT;      \
    }                                       \
  } while (0)

#define NLV_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      nlv::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return NLV_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

void count_launch(int n);
// every kernel launch site ends with this: checks the launch and feeds nlv_launch_count()
#define NLV_CHECK_LAUNCH()               \
  do {                                   \
    NLV_CHECK_CUDA(cudaGetLastError());  \
    nlv::count_launch(1);                \
  } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();
int zero_fill(float* p, long long rows, int cols, int ld, cudaStream_t s);

// dtype-tagged load/store used by the elementwise kernels (dtype: NLV_F32 / NLV_BF16)
__device__ __forceinline__ float ld_as_float(const void* p, int dtype, size_t i) {
  return dtype == NLV_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                           : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_from_float(void* p, int dtype, size_t i, float v) {
  if (dtype == NLV_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace nlv
