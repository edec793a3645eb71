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
