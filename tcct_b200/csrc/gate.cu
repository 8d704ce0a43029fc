// GateFusion of the gtc_* models (tcct.py:916-932): out = x1 * alpha + x2 * (1 - alpha).  In training alpha is a random field:
// torch.rand(B, C, max(3, H/32), max(3, W/32)) up-sampled to (H, W) with F.interpolate(mode='bicubic') and clamped to [0, 1]; in
// eval mode it is the constant 0.5.  The reference materialises the up-sampled field (a full-size tensor made on the CPU and copied
// to the device every call); here the bicubic interpolation of the small field is evaluated in registers per output element, so the
// pass moves the two operands and the result only.  NHWC fp32 activations; the small field keeps torch.rand's [B, C, hs, ws] layout.
#include "common.cuh"

// Cubic-convolution coefficients of ATen's upsample_bicubic2d (A = -0.75) for the fractional offset t in [0, 1).
__device__ __forceinline__ void cubic_coeffs(float t, float (&w)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// align_corners = False: src = scale * (dst + 0.5) - 0.5 (not clamped for the cubic kernel), taps clamped to the border.
struct CubicTap { int i[4]; float w[4]; };
__device__ __forceinline__ CubicTap cubic_tap(int dst, int in, int out) {
  const float scale = (float)in / (float)out;
  const float src = scale * ((float)dst + 0.5f) - 0.5f;
  const float fl = floorf(src);
  CubicTap t;
  cubic_coeffs(src - fl, t.w);
  const int i0 = (int)fl;
#pragma unroll
  for (int k = 0; k < 4; k++) t.i[k] = min(max(i0 - 1 + k, 0), in - 1);
  return t;
}

__device__ __forceinline__ float gate_alpha(const float* __restrict__ plane, int ws, const CubicTap& ty, const CubicTap& tx) {
  float acc = 0.f;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const float* row = plane + ty.i[r] * ws;
    const float v = tx.w[0] * __ldg(row + tx.i[0]) + tx.w[1] * __ldg(row + tx.i[1]) + tx.w[2] * __ldg(row + tx.i[2]) +
                    tx.w[3] * __ldg(row + tx.i[3]);
    acc += ty.w[r] * v;
  }
  return fminf(fmaxf(acc, 0.f), 1.f);
}

// BWD = false: out = x1 * a + x2 * (1 - a);  BWD = true (x1 = dy): o1 = dy * a, o2 = dy * (1 - a)
template <bool BWD>
__global__ void __launch_bounds__(256) gate_fuse_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                                        const float* __restrict__ alpha, float* __restrict__ o1, float* __restrict__ o2,
                                                        int B, int H, int W, int C, int hs, int ws) {
  const int c4 = C >> 2;
  const long long n = (long long)B * H * W * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int cg, x, y, b;
    split4(i, c4, W, H, cg, x, y, b);
    float a[4] = {0.5f, 0.5f, 0.5f, 0.5f};
    if (alpha) {
      const CubicTap ty = cubic_tap(y, hs, H), tx = cubic_tap(x, ws, W);
#pragma unroll
      for (int k = 0; k < 4; k++) a[k] = gate_alpha(alpha + ((size_t)b * C + cg * 4 + k) * hs * ws, ws, ty, tx);
    }
    const float4 u = reinterpret_cast<const float4*>(x1)[i];
    if (!BWD) {
      const float4 v = reinterpret_cast<const float4*>(x2)[i];
      float4 r;
      r.x = u.x * a[0] + v.x * (1.f - a[0]); r.y = u.y * a[1] + v.y * (1.f - a[1]);
      r.z = u.z * a[2] + v.z * (1.f - a[2]); r.w = u.w * a[3] + v.w * (1.f - a[3]);
      reinterpret_cast<float4*>(o1)[i] = r;
    } else {
      reinterpret_cast<float4*>(o1)[i] = make_float4(u.x * a[0], u.y * a[1], u.z * a[2], u.w * a[3]);
      reinterpret_cast<float4*>(o2)[i] = make_float4(u.x * (1.f - a[0]), u.y * (1.f - a[1]), u.z * (1.f - a[2]), u.w * (1.f - a[3]));
    }
  }
}

static int gate_grid(long long n) {
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)tcct_num_sms() * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// alpha: [B, C, hs, ws] uniform draws (training) or null (eval: 0.5)
extern "C" int tcct_gate_fuse_fwd(const float* x1, const float* x2, const float* alpha, float* out, int B, int H, int W, int C,
                                  int hs, int ws, void* stream) {
  TCCT_CHECK_ARG(C % 4 == 0 && B > 0 && H > 0 && W > 0, "gate_fuse: C must be a multiple of 4 (got %d)", C);
  TCCT_CHECK_ARG(!alpha || (hs >= 1 && ws >= 1), "gate_fuse: empty gate field");
  const long long n = (long long)B * H * W * (C / 4);
  TCCT_CHECK_ARG(n < (1ll << 31), "gate_fuse: tensor too large for 32-bit indices");
  gate_fuse_kernel<false><<<gate_grid(n), 256, 0, (cudaStream_t)stream>>>(x1, x2, alpha, out, nullptr, B, H, W, C, hs, ws);
  TCCT_CHECK_LAUNCH("gate_fuse_fwd");
  return TCCT_OK;
}

extern "C" int tcct_gate_fuse_bwd(const float* dy, const float* alpha, float* d1, float* d2, int B, int H, int W, int C, int hs,
                                  int ws, void* stream) {
  TCCT_CHECK_ARG(C % 4 == 0 && B > 0 && H > 0 && W > 0, "gate_fuse: C must be a multiple of 4 (got %d)", C);
  const long long n = (long long)B * H * W * (C / 4);
  TCCT_CHECK_ARG(n < (1ll << 31), "gate_fuse: tensor too large for 32-bit indices");
  gate_fuse_kernel<true><<<gate_grid(n), 256, 0, (cudaStream_t)stream>>>(dy, nullptr, alpha, d1, d2, B, H, W, C, hs, ws);
  TCCT_CHECK_LAUNCH("gate_fuse_bwd");
  return TCCT_OK;
}
