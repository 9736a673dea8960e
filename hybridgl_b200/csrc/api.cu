// libhgl: error plumbing and device checks (host side of the C ABI, include/hgl.h).
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "hgl_common.cuh"

namespace hgl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return HGL_ECUDA;
  }
  return HGL_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
    cached_dev = dev;
  }
  return cached;
}

int ensure_dyn_smem(const void* kernel, size_t bytes, const char* what) {
  if (bytes <= 48 * 1024) return HGL_OK;                      // the default limit needs no opt-in
  struct Entry { const void* k; int dev; size_t granted; };
  static std::mutex mu;
  static std::vector<Entry> table;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  std::lock_guard<std::mutex> lock(mu);
  Entry* hit = nullptr;
  for (auto& e : table)
    if (e.k == kernel && e.dev == dev) { hit = &e; break; }
  if (hit && hit->granted >= bytes) return HGL_OK;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%zu B of shared memory): %s", what, bytes, cudaGetErrorString(e));
    return HGL_ECUDA;
  }
  if (hit) hit->granted = bytes; else table.push_back({kernel, dev, bytes});
  return HGL_OK;
}

}  // namespace hgl

extern "C" const char* hgl_last_error(void) { return hgl::g_err; }

extern "C" int hgl_version(void) { return 100; }

extern "C" int hgl_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    hgl::set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return HGL_ECUDA;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    hgl::set_error("libhgl is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return HGL_EARCH;
  }
  return HGL_OK;
}
