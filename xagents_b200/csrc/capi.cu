// Diagnostics half of the C ABI: version, thread-local error string, device probe.
#include <cstdarg>
#include <cstdio>

#include <stdlib.h>

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

static int g_pdl_level = -1;   // -1: not read yet (XA_PDL in the environment, else 1)

int pdl_level() {
  if (g_pdl_level < 0) {
    const char* e = getenv("XA_PDL");
    g_pdl_level = e != nullptr && e[0] >= '0' && e[0] <= '3' ? e[0] - '0' : 1;
  }
  return g_pdl_level;
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

int xa_set_chained_launches(int level) {
  const int before = xa::pdl_level();
  if (level >= 0 && level <= 3) xa::g_pdl_level = level;
  return before;
}

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
