// Diagnostics half of the C ABI: version, thread-local error string, device probe.
#include <cstdarg>
#include <cstdio>

#include "xa_common.cuh"

namespace xa {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return XA_OK;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return static_cast<int>(e);
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    cached_dev = dev;
    cached = n;
  }
  return cached;
}

}  // namespace xa

extern "C" {

int xa_version(void) { return XA_VERSION; }

const char* xa_last_error(void) { return xa::g_error; }

int xa_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    xa::set_error("xa_device_info: no CUDA device (%s); this library has no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return e == cudaSuccess ? XA_EINVAL : static_cast<int>(e);
  }
  XA_REQUIRE(device >= 0 && device < n, XA_EINVAL, "xa_device_info: device %d out of range [0,%d)", device, n);
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) {
    xa::set_error("xa_device_info: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return XA_OK;
}

}  // extern "C"
