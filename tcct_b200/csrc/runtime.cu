// Error reporting and device queries behind the C ABI (include/tcct_b200.h).
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void tcct_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* tcct_last_error() { return g_err; }

int tcct_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

static unsigned long long g_launches = 0;
void tcct_count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
// Number of kernels this library has launched (or recorded into a CUDA graph) in this process.
extern "C" long long tcct_launch_count() { return (long long)g_launches; }

// Launches per tensor-core route (tests assert that the tcgen05 kernels, not a fallback, served a given shape).
static unsigned long long g_routes[TCCT_ROUTE_COUNT] = {0};
void tcct_count_route(int id) { if (id >= 0 && id < TCCT_ROUTE_COUNT) __atomic_fetch_add(&g_routes[id], 1ull, __ATOMIC_RELAXED); }
extern "C" long long tcct_route_count(int id) { return (id >= 0 && id < TCCT_ROUTE_COUNT) ? (long long)g_routes[id] : -1; }

extern "C" int tcct_abi_version() { return 2; }

// Compute capability of the current device as major*10+minor, or -1 without a usable device.
extern "C" int tcct_device_arch() {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}

// cuTensorMapEncodeTiled resolved through the runtime (the library does not link libcuda); null when unavailable.
#include "tma.cuh"
tcct_encode_tiled_fn tcct_tensor_map_encoder() {
  static const tcct_encode_tiled_fn fn = [] {      // initialised once, thread-safe (C++11 static)
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      return (tcct_encode_tiled_fn)p;
    return (tcct_encode_tiled_fn) nullptr;
  }();
  return fn;
}
