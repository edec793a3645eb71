// Error text, launch counter and device queries shared by every translation unit.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace nlv {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

__global__ void zero_fill_kernel(float* p, long long rows, int cols, int ld) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  p[r * ld + (i - r * cols)] = 0.f;
}

// Zero a strided fp32 matrix with a kernel (cudaMemset*Async may be routed to a copy engine, where it would queue
// behind a multi-GB host-to-device prefetch running on another stream).
int zero_fill(float* p, long long rows, int cols, int ld, cudaStream_t s) {
  if (rows * cols == 0) return NLV_OK;
  zero_fill_kernel<<<cdiv(rows * cols, 256), 256, 0, s>>>(p, rows, cols, ld);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace nlv

extern "C" {
const char* nlv_last_error(void) { return nlv::g_err; }
int nlv_version(void) { return 100; }
long long nlv_launch_count(void) { return nlv::g_launches.load(); }
}
