// Version, error string and architecture guard of the C ABI (include/mae_clip_b200.h).
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace mc {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int arch_check() {
  static int cached[64];  // 0 unknown, 1 ok, 2 bad
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice failed: %s (no CUDA device? this library has no CPU path)",
              cudaGetErrorString(e));
    return MC_ERR_CUDA;
  }
  if (dev >= 0 && dev < 64 && cached[dev] == 1) return MC_OK;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    set_error("device %d is sm_%d%d; mae_clip_b200 kernels are built for sm_100a only", dev, major,
              minor);
    if (dev >= 0 && dev < 64) cached[dev] = 2;
    return MC_ERR_ARCH;
  }
  if (dev >= 0 && dev < 64) cached[dev] = 1;
  return MC_OK;
}

}  // namespace mc

extern "C" {

int mc_version(void) { return 0 * 10000 + 1 * 100 + 0; }

const char* mc_last_error_string(void) { return mc::g_err; }

unsigned long long mc_kernel_launch_count(void) {
  return mc::g_launches.load(std::memory_order_relaxed);
}

int mc_device_supported(int device) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

}  // extern "C"
