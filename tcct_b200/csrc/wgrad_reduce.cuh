// Second phase of the tcgen05 weight-gradient kernels: every CTA of the main kernel leaves a partial-sum slab in a workspace, the
// bodies below fold the slabs into dW.  They run either right behind the main kernel (one small launch per layer:
// wgrad_line_reduce_kernel, wgrad_gemm_reduce_kernel) or, for a whole backward pass at once, from wgrad_reduce_batch_kernel
// (csrc/wgrad_reduce.cu): the per-layer launches are pure latency (7-12 us for a few MB) and sat, 48 of them, on the
// weight-gradient streams that finish the step.
#pragma once
#include "tma.cuh"

// conv (line) layout: partials [part][S][4096] (lane (j, co), column (group, ci) of the accumulators) -> dW[co][ci][tap].
// Four threads share an element (partials q, q+4, ...; four loads in flight each), then two shuffles.  256 threads per block.
__device__ __forceinline__ void wgrad_line_reduce_body(const float* __restrict__ ws, int nparts, int S, int KA, int KL, float* dw, int block) {
  const int T = KA * KL;
  const int total = S * 4096;
  const int gt = block * 256 + threadIdx.x;
  const int e = gt >> 2, q = gt & 3;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (e < total) {
    const float* p = ws + e;
    int c = q;
    for (; c + 12 < nparts; c += 16) {
      s0 += __ldcg(p + (size_t)c * total); s1 += __ldcg(p + (size_t)(c + 4) * total);
      s2 += __ldcg(p + (size_t)(c + 8) * total); s3 += __ldcg(p + (size_t)(c + 12) * total);
    }
    for (; c < nparts; c += 4) s0 += __ldcg(p + (size_t)c * total);
  }
  float sum = (s0 + s1) + (s2 + s3);
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  if (e < total && q == 0) {
    const int g = e >> 12, j = (e >> 10) & 3, co = (e >> 5) & 31, ci = e & 31;
    int tap;
    if (KA == 3) { const int kx = 3 - j; tap = (kx >= 0 && kx < 3) ? g * 3 + kx : -1; }
    else { const int kl = 4 * g + 3 - j; tap = kl < KL ? kl : -1; }
    if (tap >= 0) dw[(size_t)co * 32 * T + (size_t)ci * T + tap] += sum;
  }
}
static inline int wgrad_line_reduce_blocks(int S) { return (S * 4096 * 4 + 255) / 256; }

// 1x1 / linear layout: partials [part][128][K] -> dW[n][k] (row stride ld); only rows n < N are real.  NW warps per block.
template <int NW>
__device__ __forceinline__ void wgrad_gemm_reduce_body(const float* __restrict__ ws, int nparts, int N, int K, int ld, float* dw, int block,
                                                       int nblocks, float* s_part) {
  const int total = N * K;
  const int per = (((total + nblocks - 1) / nblocks) + 31) & ~31;
  const int e0 = block * per, e1 = min(e0 + per, total);
  reduce_partials<NW>(ws, 128 * K, (unsigned int)nparts, e0, e1, s_part, [&](int e, float sum) {
    const int n = e / K, k = e - n * K;
    dw[(size_t)n * ld + k] += sum;
  });
}
static inline int wgrad_gemm_reduce_blocks(int N, int K, int sms) {
  int rg = (N * K + 31) / 32;           // one 32-element slice per CTA while they last: the reduction is pure load latency
  return rg > 2 * sms ? 2 * sms : rg;
}
