// Shared device/host helpers for the tcct_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define TCCT_OK 0
#define TCCT_ERR_ARG 1
#define TCCT_ERR_CUDA 2

extern "C" const char* tcct_last_error();
void tcct_set_error(const char* fmt, ...);

#define TCCT_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      tcct_set_error(__VA_ARGS__);                \
      return TCCT_ERR_ARG;                        \
    }                                             \
  } while (0)

void tcct_count_launch();
// tcct_route_count ids (include/tcct_b200.h)
enum { TCCT_ROUTE_CONV_TMA = 0, TCCT_ROUTE_WGRAD_TMA = 1, TCCT_ROUTE_GEMM_TMA = 2, TCCT_ROUTE_WGRAD_GEMM_TMA = 3, TCCT_ROUTE_COUNT = 8 };
void tcct_count_route(int id);

#define TCCT_CHECK_LAUNCH(name)                                          \
  do {                                                                   \
    tcct_count_launch();                                                 \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) {                                            \
      tcct_set_error("%s: %s", name, cudaGetErrorString(e__));           \
      return TCCT_ERR_CUDA;                                              \
    }                                                                    \
  } while (0)

int tcct_num_sms();

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- activations (codes shared with include/tcct_b200.h) -------------------
enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_HSWISH = 2, ACT_GELU = 3 };

// Standard normal cdf and pdf at z for the exact (erf) GELU of the reference (F.gelu, tcct.py:35,826).
// erfc(|z|/sqrt2) comes from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 on erf, i.e. fp32 round-off level) and shares
// its exp(-z^2/2) with the density, so value and derivative cost one exponential and no erff call; the lower tail is
// formed without cancellation.
__device__ __forceinline__ void gelu_cdf_pdf(float z, float& cdf, float& pdf) {
  const float x = fabsf(z) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, x, 1.f));
  const float e = __expf(-x * x);
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float h = 0.5f * p * t * e;          // 0.5 * erfc(|z| / sqrt 2)
  cdf = z >= 0.f ? 1.f - h : h;
  pdf = 0.3989422804014327f * e;
}

__device__ __forceinline__ float act_fwd(int act, float z) {
  switch (act) {
    case ACT_LRELU: return z > 0.f ? z : 0.01f * z;
    case ACT_HSWISH: return z * fminf(fmaxf(z + 3.f, 0.f), 6.f) * (1.f / 6.f);
    case ACT_GELU: { float c, d; gelu_cdf_pdf(z, c, d); return z * c; }
    default: return z;
  }
}
// Activation inside a statistics epilogue.  The kind is a run-time value there; written as `act_fwd(kind, v)` in the element loop the
// compiler evaluates every kind (GELU's exponential and divide included) and selects -- measured 2x on the 1x1-conv GEMM with
// statistics.  The two kinds the networks use are inline, the rest sit behind a call that cannot be if-converted.
static __device__ __noinline__ float act_fwd_rare(int act, float z) { return act_fwd(act, z); }
__device__ __forceinline__ float stat_act(int act, float z) {
  if (act == ACT_NONE) return z;
  if (act == ACT_LRELU) return z > 0.f ? z : 0.01f * z;
  return act_fwd_rare(act, z);
}
// derivative with respect to the pre-activation z
__device__ __forceinline__ float act_bwd(int act, float z) {
  switch (act) {
    case ACT_LRELU: return z > 0.f ? 1.f : 0.01f;
    case ACT_HSWISH: return z < -3.f ? 0.f : (z <= 3.f ? z * (1.f / 3.f) + 0.5f : 1.f);
    case ACT_GELU: { float c, d; gelu_cdf_pdf(z, c, d); return c + z * d; }
    default: return 1.f;
  }
}

// Linear index -> (fastest, middle, slowest) coordinates with 32-bit arithmetic.  64-bit integer division is a ~100-instruction
// sequence on the GPU; three of them per element made the strided depthwise-conv gradients 3-5x slower than their memory traffic
// (8x64x64x64 dy -> 128x128 dx: 29 us for 50 MB).  Every launcher checks that the element count stays below 2^31.
__device__ __forceinline__ void split3(long long i, int d0, int d1, int& i0, int& i1, int& i2) {
  const unsigned int u = (unsigned int)i;
  const unsigned int q = u / (unsigned int)d0;
  i0 = (int)(u - q * (unsigned int)d0);
  const unsigned int q2 = q / (unsigned int)d1;
  i1 = (int)(q - q2 * (unsigned int)d1);
  i2 = (int)q2;
}
__device__ __forceinline__ void split4(long long i, int d0, int d1, int d2, int& i0, int& i1, int& i2, int& i3) {
  const unsigned int u = (unsigned int)i;
  const unsigned int q = u / (unsigned int)d0;
  i0 = (int)(u - q * (unsigned int)d0);
  split3((long long)q, d1, d2, i1, i2, i3);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// v[0..31] per lane -> v[0] on lane c = sum over the warp's lanes of v[c]  (31 shuffles: at step s the lanes with bit s set
// keep the upper half of their values and hand the lower half to their partner, and vice versa)
__device__ __forceinline__ void warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; k++) {
      const float keep = up ? v[k + s] : v[k];
      const float send = up ? v[k] : v[k + s];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- TF32 tensor-core primitives (legacy warp-level path) -------------------
__device__ __forceinline__ uint32_t f2tf32(float f) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(f));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// 16-byte async copy global->shared; src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
