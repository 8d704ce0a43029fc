// clip_grad_norm_(12) + AdamW over the flat parameter / gradient buffers
// (task1/kite/loop_seg.py:128-130, task1/kite/loopback.py:126-128: AdamW(lr, weight_decay=2e-4)).
// Everything the host would have to read (norm, step count, learning rate) stays on the device so the
// whole train step is capturable in one CUDA graph.
#include "common.cuh"

__global__ void sqnorm_kernel(const float* __restrict__ g, long long n, double* out) {
  __shared__ float sw[32];
  float s = 0.f;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(g)[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sw[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, (double)s);
  }
}

// state: [0] step count (as float), [1] lr, [2] last total grad norm (output)
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, const double* sqnorm, float* state, float max_norm, float beta1, float beta2,
                             float eps, float wd, float grad_scale) {
  // the step constants once per CTA (a double-precision square root and two powf per THREAD were most of this kernel's 9 us)
  __shared__ float s_k[4];
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(*sqnorm) * grad_scale;
    float c = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
    if (c > 1.f) c = 1.f;
    const float step = state[0] + 1.f;
    const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
    s_k[0] = c * grad_scale; s_k[1] = state[1]; s_k[2] = state[1] / bc1; s_k[3] = rsqrtf(bc2);
  }
  __syncthreads();
  const float clip = s_k[0], lr = s_k[1], step_size = s_k[2], inv_sqrt_bc2 = s_k[3];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * clip;
    float pi = p[i];
    pi *= 1.f - lr * wd;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pi - step_size * mi / denom;
  }
}
__global__ void adamw_tick_kernel(float* state, const double* sqnorm, float grad_scale) {
  state[0] += 1.f;
  state[2] = (float)sqrt(*sqnorm) * grad_scale;
}

extern "C" int tcct_sqnorm(const float* g, long long n, double* out, void* stream) {
  int blocks = ceil_div(n / 4 + 1, 256);
  const int cap = tcct_num_sms() * 4;
  if (blocks > cap) blocks = cap;
  sqnorm_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out);
  TCCT_CHECK_LAUNCH("sqnorm");
  return TCCT_OK;
}

// sqnorm: device double holding sum(g^2) over ALL clipped gradients (tcct_sqnorm, zeroed by the caller before);
// grad_scale multiplies every gradient first (1/world_size after a sum all-reduce).
extern "C" int tcct_adamw_step(float* p, const float* g, float* m, float* v, long long n, const double* sqnorm,
                               float* state, float max_norm, float beta1, float beta2, float eps, float wd,
                               float grad_scale, void* stream) {
  int blocks = ceil_div(n, 256);
  const int cap = tcct_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, sqnorm, state, max_norm, beta1, beta2, eps, wd, grad_scale);
  TCCT_CHECK_LAUNCH("adamw");
  adamw_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, sqnorm, grad_scale);
  TCCT_CHECK_LAUNCH("adamw_tick");
  return TCCT_OK;
}
