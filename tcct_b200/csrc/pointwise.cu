// Bandwidth-bound kernels of the stc_tt path (NHWC fp32 activations unless stated):
// batch-norm statistics/finalise, the fused (BN|identity)+(BN|identity) -> activation family and its
// backward, max-pool, depthwise 3x3, LayerNorm, MetaPool, bilinear resampling, L2 normalisation,
// the 3-channel stem convs and the 32->C logit heads.
// Reference semantics: task1/nets/tcct.py (line ranges cited per kernel).
#include "common.cuh"
#include <stdlib.h>

// Thread <-> data mapping shared by the channel-reducing kernels: a block is PPB pixels x (C/4) channel
// groups; thread (prow, cg) owns channels [4cg, 4cg+4) of pixels prow, prow+PPB*grid, ... so per-channel
// partial sums live in registers.
struct CgMap {
  int cgs, ppb, threads;
};
static CgMap cg_map(int C) {
  CgMap m;
  m.cgs = C / 4;
  m.ppb = 256 / m.cgs;
  if (m.ppb < 1) m.ppb = 1;
  m.threads = m.ppb * m.cgs;
  return m;
}
static int grid_for(long long npix, int ppb, int per_sm = 8) {
  long long blocks = (npix + ppb - 1) / ppb;
  long long cap = (long long)tcct_num_sms() * per_sm;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// ----------------------------------------------------------------------------------------------
// Per-channel sum / sum of squares of an NHWC tensor (double accumulators in global memory).
// ----------------------------------------------------------------------------------------------
__global__ void stats_nhwc_kernel(const float* __restrict__ x, long long npix, int C, int ppb, double* stats) {
  extern __shared__ float sred[];     // [2*C]
  const int cgs = C >> 2;
  const int cg = threadIdx.x % cgs, prow = threadIdx.x / cgs;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  for (long long p = (long long)blockIdx.x * ppb + prow; p < npix; p += (long long)gridDim.x * ppb) {
    const float4 v = *reinterpret_cast<const float4*>(x + p * C + cg * 4);
    s[0] += v.x; q[0] += v.x * v.x; s[1] += v.y; q[1] += v.y * v.y;
    s[2] += v.z; q[2] += v.z * v.z; s[3] += v.w; q[3] += v.w * v.w;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    atomicAdd(&sred[cg * 4 + i], s[i]);
    atomicAdd(&sred[C + cg * 4 + i], q[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(stats + i, (double)sred[i]);
}

extern "C" int tcct_stats_nhwc(const float* x, long long npix, int C, double* stats, void* stream) {
  TCCT_CHECK_ARG(C % 4 == 0 && C <= 1024, "stats_nhwc: C must be a multiple of 4 (got %d)", C);
  const CgMap m = cg_map(C);
  stats_nhwc_kernel<<<grid_for(npix, m.ppb, 4), m.threads, 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      x, npix, C, m.ppb, stats);
  TCCT_CHECK_LAUNCH("stats_nhwc");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// BatchNorm finalise (nn.BatchNorm2d train mode: biased variance normalises, unbiased variance feeds
// the running estimate; momentum 0.1).  coef = [scale | shift | mean | invstd], each [C].
// stats == null -> eval mode: coefficients from the running statistics.
// ----------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* stats, double count, const float* gamma, const float* beta,
                                   float eps, float momentum, float* running_mean, float* running_var,
                                   long long* num_batches, int update_running, float* coef, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && stats && update_running && num_batches) *num_batches += 1;
  if (c >= C) return;
  float mean, invstd;
  if (stats) {
    const double m = stats[c] / count;
    double var = stats[C + c] / count - m * m;
    if (var < 0) var = 0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (update_running) {
      const double unb = count > 1 ? var * count / (count - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
  } else {
    mean = running_mean[c];
    invstd = rsqrtf(running_var[c] + eps);
  }
  const float sc = gamma[c] * invstd;
  coef[c] = sc;
  coef[C + c] = beta[c] - mean * sc;
  coef[2 * C + c] = mean;
  coef[3 * C + c] = invstd;
}

extern "C" int tcct_bn_finalize(const double* stats, double count, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, long long* num_batches,
                                int update_running, float* coef, int C, void* stream) {
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(stats, count, gamma, beta, eps, momentum,
                                                                         running_mean, running_var, num_batches,
                                                                         update_running, coef, C);
  TCCT_CHECK_LAUNCH("bn_finalize");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// out = post( opA(a) + opB(b) ),  op(v) = scale[c]*pre(v) + shift[c]   (coef null -> identity op,
// b null -> single operand).  Covers BN+act, BN+BN fusion adds, GELU(BN(a)+BN(b)) of
// CrossCNNBlock.forward (tcct.py:825-828), x + BN(conv2) of ResBlock (tcct.py:562-571), plain adds.
// ----------------------------------------------------------------------------------------------
struct Bn2Args {
  const float* a; const float* coefA; int preA;
  const float* b; const float* coefB; int preB;
  int post;
  long long npix; int C;
};

// Activation codes as template parameters (ACT_DYN = read the runtime code): the hot CrossCNNBlock combination
// (LeakyReLU, LeakyReLU, GELU) is compiled branch-free, everything else shares one generic instantiation.
#define ACT_DYN (-1)
template <int A> __device__ __forceinline__ float actf(int rt, float z) { return act_fwd(A == ACT_DYN ? rt : A, z); }
template <int A> __device__ __forceinline__ float actb(int rt, float z) { return act_bwd(A == ACT_DYN ? rt : A, z); }

// Thread <-> data mapping of the whole family (CgMap): thread (prow, cg) owns channels [4cg, 4cg+4) of pixels
// prow, prow + ppb*grid, ...  Its per-channel constants (BN scale/shift/mean/invstd, reduction means) live in
// registers, so the inner loop is loads, arithmetic and one store.
struct ChanCoef { float sc[4], sh[4], mu[4], is[4]; };
__device__ __forceinline__ ChanCoef load_coef(const float* coef, int C, int c0) {
  ChanCoef k;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    k.sc[i] = coef ? coef[c0 + i] : 1.f; k.sh[i] = coef ? coef[C + c0 + i] : 0.f;
    k.mu[i] = coef ? coef[2 * C + c0 + i] : 0.f; k.is[i] = coef ? coef[3 * C + c0 + i] : 0.f;
  }
  return k;
}

// One BatchNorm operand of the fused forward: batch sums in (train) or running statistics (eval), coefficients out.
// Mirrors `tcct_bn_src` of include/tcct_b200.h.
struct BnSrc {
  const double* stats;      // [2C] sum | sum of squares of the operand (null: use the running statistics)
  double count;             // elements per channel behind `stats`
  const float* gamma; const float* beta;
  float eps, momentum;
  float* running_mean; float* running_var; long long* num_batches;
  int update_running;
  float* coef;              // out [4C]: scale | shift | mean | invstd (what the backward needs); null: no BatchNorm
};
// The finalisation of bn_finalize_kernel inside a CTA: thread c finalises channel c (double-precision mean / variance / 1/sqrt ONCE per
// channel and CTA) into shared memory [sc | sh | mu | is][C].  An earlier version let every thread finalise its own four channels: 256
// threads x 4 double divisions and square roots in the prologue of every CTA cost 8.5 us per launch whatever the tensor size -- more
// than the whole pass on the small maps (scripts/time_bnfwd.py).
__device__ __forceinline__ void make_coef_cta(const BnSrc& b, int C, float* s, bool writer_cta) {
  if (!b.coef) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) { s[c] = 1.f; s[C + c] = 0.f; s[2 * C + c] = 0.f; s[3 * C + c] = 0.f; }
    return;
  }
  const double inv_count = 1.0 / b.count;        // one double division per thread; the rest is multiplies and one rsqrt
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, invstd;
    if (b.stats) {
      const double m = b.stats[c] * inv_count;
      double var = b.stats[C + c] * inv_count - m * m;
      if (var < 0) var = 0;
      mean = (float)m;
      invstd = (float)rsqrt(var + (double)b.eps);
      if (writer_cta && b.update_running) {
        const double unb = b.count > 1 ? var * b.count / (b.count - 1) : var;
        b.running_mean[c] = (1.f - b.momentum) * b.running_mean[c] + b.momentum * (float)m;
        b.running_var[c] = (1.f - b.momentum) * b.running_var[c] + b.momentum * (float)unb;
      }
    } else {
      mean = b.running_mean[c];
      invstd = rsqrtf(b.running_var[c] + b.eps);
    }
    const float sc = b.gamma[c] * invstd, sh = b.beta[c] - mean * sc;
    s[c] = sc; s[C + c] = sh; s[2 * C + c] = mean; s[3 * C + c] = invstd;
    if (writer_cta) { b.coef[c] = sc; b.coef[C + c] = sh; b.coef[2 * C + c] = mean; b.coef[3 * C + c] = invstd; }
  }
  if (writer_cta && threadIdx.x == 0 && b.stats && b.update_running && b.num_batches) *b.num_batches += 1;
}
__device__ __forceinline__ ChanCoef smem_coef(const float* s, int C, int c0) {
  ChanCoef k;
#pragma unroll
  for (int i = 0; i < 4; i++) { k.sc[i] = s[c0 + i]; k.sh[i] = s[C + c0 + i]; k.mu[i] = s[2 * C + c0 + i]; k.is[i] = s[3 * C + c0 + i]; }
  return k;
}

#define BN_U 4
#define BN_SLOTS 8
// FUSED: the BatchNorm finalisation (batch sums -> coefficients, running statistics) happens in this kernel's prologue
template <int PA, int PB, int PO, bool FUSED>
__global__ void __launch_bounds__(256) bn_act2_fwd_kernel(const Bn2Args g, const BnSrc sa, const BnSrc sb, int ppb, float* __restrict__ out) {
  extern __shared__ float s_coef[];      // FUSED: [A: sc sh mu is][C] [B: ...][C]
  const int C = g.C, cgs = C >> 2;
  const int cg = threadIdx.x % cgs, prow = threadIdx.x / cgs;
  ChanCoef ka, kb;
  if (FUSED) {
    make_coef_cta(sa, C, s_coef, blockIdx.x == 0);
    make_coef_cta(sb, C, s_coef + 4 * C, blockIdx.x == 0);
    __syncthreads();
    ka = smem_coef(s_coef, C, cg * 4);
    kb = smem_coef(s_coef + 4 * C, C, cg * 4);
  } else {
    ka = load_coef(g.coefA, C, cg * 4);
    kb = load_coef(g.coefB, C, cg * 4);
  }
  const bool has_b = g.b != nullptr;
  // BN_U pixels per iteration: all loads are issued before the arithmetic (bytes in flight, not occupancy, feed HBM).  (Issuing the
  // first iteration's loads ahead of the coefficient prologue was measured and is slower: 9.3 vs 9.0 us at 16.8 MB.)
  const long long stride = (long long)gridDim.x * ppb;
  for (long long p = (long long)blockIdx.x * ppb + prow; p < g.npix; p += BN_U * stride) {
    long long off[BN_U];
    bool ok[BN_U];
    float4 a4[BN_U], b4[BN_U];
#pragma unroll
    for (int u = 0; u < BN_U; u++) {
      off[u] = (p + u * stride) * C + cg * 4;
      ok[u] = p + u * stride < g.npix;
      a4[u] = b4[u] = make_float4(0, 0, 0, 0);
      if (ok[u]) {
        a4[u] = __ldcs(reinterpret_cast<const float4*>(g.a + off[u]));
        if (has_b) b4[u] = __ldcs(reinterpret_cast<const float4*>(g.b + off[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < BN_U; u++) {
      if (!ok[u]) continue;
      const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w}, bv[4] = {b4[u].x, b4[u].y, b4[u].z, b4[u].w};
      float o[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float z = actf<PA>(g.preA, av[i]) * ka.sc[i] + ka.sh[i];
        if (has_b) z += actf<PB>(g.preB, bv[i]) * kb.sc[i] + kb.sh[i];
        o[i] = actf<PO>(g.post, z);
      }
      *reinterpret_cast<float4*>(out + off[u]) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

// The activation kinds are template arguments for the combinations the networks use: with run-time kinds the compiler evaluates
// every activation (GELU's exponential and divide included) per element and selects, which made the pass compute-bound (50 us for
// the 134 MB of a full-resolution 32-channel tensor against 21 us at copy speed; same finding as in the backward below).
template <bool FUSED>
static void bn_act2_fwd_launch(const Bn2Args& g, const BnSrc& sa, const BnSrc& sb, int grid, int threads, size_t smem, cudaStream_t st,
                               int ppb, float* out) {
#define BN_FWD(PA, PB, PO) bn_act2_fwd_kernel<PA, PB, PO, FUSED><<<grid, threads, smem, st>>>(g, sa, sb, ppb, out)
  const bool b_plain = !g.b || g.preB == ACT_NONE;
  if (g.b && g.preA == ACT_LRELU && g.preB == ACT_LRELU && g.post == ACT_GELU) BN_FWD(ACT_LRELU, ACT_LRELU, ACT_GELU);
  else if (g.preA == ACT_LRELU && b_plain && g.post == ACT_NONE) BN_FWD(ACT_LRELU, ACT_NONE, ACT_NONE);
  else if (g.preA == ACT_NONE && b_plain && g.post == ACT_HSWISH) BN_FWD(ACT_NONE, ACT_NONE, ACT_HSWISH);
  else if (g.preA == ACT_NONE && b_plain && g.post == ACT_LRELU) BN_FWD(ACT_NONE, ACT_NONE, ACT_LRELU);
  else if (g.preA == ACT_NONE && b_plain && g.post == ACT_NONE) BN_FWD(ACT_NONE, ACT_NONE, ACT_NONE);
  else if (g.preA == ACT_NONE && b_plain && g.post == ACT_GELU) BN_FWD(ACT_NONE, ACT_NONE, ACT_GELU);
  else BN_FWD(ACT_DYN, ACT_DYN, ACT_DYN);
#undef BN_FWD
}

extern "C" int tcct_bn_act2_fwd(const float* a, const float* coefA, int preA, const float* b, const float* coefB,
                                int preB, int post, float* out, long long npix, int C, void* stream) {
  TCCT_CHECK_ARG(C % 4 == 0 && C <= 1024, "bn_act2: C must be a multiple of 4, <= 1024 (got %d)", C);
  Bn2Args g{a, coefA, preA, b, coefB, preB, post, npix, C};
  const CgMap m = cg_map(C);
  const int grid = grid_for(npix, m.ppb, 8);
  const BnSrc none{};
  bn_act2_fwd_launch<false>(g, none, none, grid, m.threads, 0, (cudaStream_t)stream, m.ppb, out);
  TCCT_CHECK_LAUNCH("bn_act2_fwd");
  return TCCT_OK;
}

// The same with the BatchNorm finalisation fused in: bnA / bnB are HOST pointers to tcct_bn_src records (null: the
// operand is not normalised); the records are read at call time, their coef arrays are written by the kernel.
extern "C" int tcct_bn_act2_fwd_bn(const float* a, const BnSrc* bnA, int preA, const float* b, const BnSrc* bnB, int preB,
                                   int post, float* out, long long npix, int C, void* stream) {
  TCCT_CHECK_ARG(C % 4 == 0 && C <= 1024, "bn_act2: C must be a multiple of 4, <= 1024 (got %d)", C);
  const BnSrc none{};
  const BnSrc sa = bnA ? *bnA : none, sb = (bnB && b) ? *bnB : none;
  Bn2Args g{a, sa.coef, preA, b, sb.coef, preB, post, npix, C};
  const CgMap m = cg_map(C);
  const int grid = grid_for(npix, m.ppb, 8);
  const size_t smem = (size_t)8 * C * sizeof(float);
  bn_act2_fwd_launch<true>(g, sa, sb, grid, m.threads, smem, (cudaStream_t)stream, m.ppb, out);
  TCCT_CHECK_LAUNCH("bn_act2_fwd_bn");
  return TCCT_OK;
}

// Backward, pass 1: per channel  S1 = sum dz,  S2a = sum dz*xhat_a,  S2b = sum dz*xhat_b,  dz = dout*post'(z)
template <int PA, int PB, int PO>
__global__ void __launch_bounds__(256) bn_act2_bwd_reduce_kernel(const Bn2Args g, const float* __restrict__ dout, int ppb, double* sums) {
  extern __shared__ float sred[];     // [3*C]
  const int C = g.C, cgs = C >> 2;
  const int cg = threadIdx.x % cgs, prow = threadIdx.x / cgs;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  const ChanCoef ka = load_coef(g.coefA, C, cg * 4), kb = load_coef(g.coefB, C, cg * 4);
  const bool has_b = g.b != nullptr;
  float s1[4] = {0, 0, 0, 0}, sa[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
  const long long stride = (long long)gridDim.x * ppb;
  for (long long p = (long long)blockIdx.x * ppb + prow; p < g.npix; p += 2 * stride) {
    const long long off[2] = {p * C + cg * 4, (p + stride) * C + cg * 4};
    const bool ok1 = p + stride < g.npix;
    float4 a4[2], b4[2], d4[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      a4[u] = b4[u] = d4[u] = make_float4(0, 0, 0, 0);
      if (u == 0 || ok1) {
        a4[u] = *reinterpret_cast<const float4*>(g.a + off[u]);
        if (has_b) b4[u] = *reinterpret_cast<const float4*>(g.b + off[u]);
        d4[u] = *reinterpret_cast<const float4*>(dout + off[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      if (u == 1 && !ok1) break;
      const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w}, bv[4] = {b4[u].x, b4[u].y, b4[u].z, b4[u].w},
                  dv[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float pa = actf<PA>(g.preA, av[i]);
        float z = pa * ka.sc[i] + ka.sh[i];
        float pb = 0.f;
        if (has_b) { pb = actf<PB>(g.preB, bv[i]); z += pb * kb.sc[i] + kb.sh[i]; }
        const float dz = dv[i] * actb<PO>(g.post, z);
        s1[i] += dz;
        sa[i] += dz * (pa - ka.mu[i]) * ka.is[i];
        if (has_b) sb[i] += dz * (pb - kb.mu[i]) * kb.is[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    atomicAdd(&sred[cg * 4 + i], s1[i]);
    atomicAdd(&sred[C + cg * 4 + i], sa[i]);
    atomicAdd(&sred[2 * C + cg * 4 + i], sb[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) atomicAdd(sums + i, (double)sred[i]);
}

// Backward, pass 2: da, db (+ dgamma/dbeta accumulated into the parameter gradients by block 0)
template <int PA, int PB, int PO>
__global__ void __launch_bounds__(256) bn_act2_bwd_apply_kernel(const Bn2Args g, const float* __restrict__ dout, const double* sums,
                                                                const float* gammaA, const float* gammaB, int ppb,
                                                                float* __restrict__ da, float* __restrict__ db, float* dgammaA,
                                                                float* dbetaA, float* dgammaB, float* dbetaB) {
  const int C = g.C, cgs = C >> 2;
  if (blockIdx.x == 0 && sums) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (g.coefA && dgammaA) { dgammaA[c] += (float)sums[C + c]; dbetaA[c] += (float)sums[c]; }
      if (g.coefB && dgammaB) { dgammaB[c] += (float)sums[2 * C + c]; dbetaB[c] += (float)sums[c]; }
    }
  }
  const int cg = threadIdx.x % cgs, prow = threadIdx.x / cgs;
  const ChanCoef ka = load_coef(g.coefA, C, cg * 4), kb = load_coef(g.coefB, C, cg * 4);
  const bool has_b = g.b != nullptr, bn_a = g.coefA != nullptr, bn_b = g.coefB != nullptr;
  const float inv_n = 1.f / (float)g.npix;
  // da = gi * (dz - m1 - xhat * m2) with gi = gamma * invstd and the batch means m1, m2 (0 in eval mode)
  float gia[4], gib[4], m1[4], m2a[4], m2b[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = cg * 4 + i;
    gia[i] = bn_a ? gammaA[c] * ka.is[i] : 1.f;
    gib[i] = bn_b ? gammaB[c] * kb.is[i] : 1.f;
    m1[i] = sums ? (float)sums[c] * inv_n : 0.f;
    m2a[i] = sums ? (float)sums[C + c] * inv_n : 0.f;
    m2b[i] = sums ? (float)sums[2 * C + c] * inv_n : 0.f;
  }
  const long long stride = (long long)gridDim.x * ppb;
  for (long long p = (long long)blockIdx.x * ppb + prow; p < g.npix; p += 2 * stride) {
    const long long off[2] = {p * C + cg * 4, (p + stride) * C + cg * 4};
    const bool ok1 = p + stride < g.npix;
    float4 a4[2], b4[2], d4[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      a4[u] = b4[u] = d4[u] = make_float4(0, 0, 0, 0);
      if (u == 0 || ok1) {
        a4[u] = *reinterpret_cast<const float4*>(g.a + off[u]);
        if (has_b) b4[u] = *reinterpret_cast<const float4*>(g.b + off[u]);
        d4[u] = *reinterpret_cast<const float4*>(dout + off[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      if (u == 1 && !ok1) break;
      const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w}, bv[4] = {b4[u].x, b4[u].y, b4[u].z, b4[u].w},
                  dv[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
      float ra[4], rb[4];
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float pa = actf<PA>(g.preA, av[i]);
        float z = pa * ka.sc[i] + ka.sh[i];
        float pb = 0.f;
        if (has_b) { pb = actf<PB>(g.preB, bv[i]); z += pb * kb.sc[i] + kb.sh[i]; }
        const float dz = dv[i] * actb<PO>(g.post, z);
        const float ga = bn_a ? gia[i] * (dz - m1[i] - (pa - ka.mu[i]) * ka.is[i] * m2a[i]) : dz;
        ra[i] = ga * actb<PA>(g.preA, av[i]);
        if (has_b) {
          const float gb = bn_b ? gib[i] * (dz - m1[i] - (pb - kb.mu[i]) * kb.is[i] * m2b[i]) : dz;
          rb[i] = gb * actb<PB>(g.preB, bv[i]);
        }
      }
      *reinterpret_cast<float4*>(da + off[u]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
      if (has_b && db) *reinterpret_cast<float4*>(db + off[u]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
  }
}


// Backward in two launches of the same persistent grid (train mode).  Each CTA owns a contiguous pixel range: launch 1
// reduces it front to back, launch 2 walks the SAME range BACK to front, so the lines it re-reads are the ones touched
// last and are largely still in the 126 MB L2 (two grid-strided launches stream every operand from HBM twice).  An
// earlier single-launch version joined the passes with a software grid barrier: that needs either co-residency luck
// (two such kernels on parallel graph branches can dead-lock) or a cooperative launch (which waits for an empty GPU and
// cost 2 % of the step); the kernel boundary is the barrier that is always safe.  Per-channel constants live in shared memory (float4 per channel group) so that four pixels
// of loads per thread fit the register budget of two CTAs per SM.
struct BnBwdArgs {
  Bn2Args g;
  const float* dout;
  double* sums;               // zeroed: BN_SLOTS replicas of [S1 | S2a | S2b] (3C each; CTA i adds into replica i % BN_SLOTS so that
                              // the same-address atomics of ~300 CTAs do not serialise), then the barrier counter (32-bit)
  const float* gammaA; const float* gammaB;
  float* da; float* db;
  float* dgammaA; float* dbetaA; float* dgammaB; float* dbetaB;
  long long chunk;            // pixels per CTA (multiple of ppb)
  int ppb;
  int phase;                  // 1: reduce pass (front to back), 2: apply pass (back to front)
};
template <int PA, int PB, int PO, int UU = 4, int MINB = 2>
__global__ void __launch_bounds__(256, MINB) bn_act2_bwd_fused_kernel(const BnBwdArgs q) {
  extern __shared__ __align__(16) float sm[];
  const Bn2Args& g = q.g;
  const int C = g.C, cgs = C >> 2;
  // shared layout: red [3C] | kA: sc sh mu is [4C] | kB [4C] | gi_a gi_b m1 m2a m2b [5C] | part [pixel rows][3C]
  float* red = sm;
  float* kA = sm + 3 * C;
  float* kB = kA + 4 * C;
  float* kk = kB + 4 * C;
  const int cg = threadIdx.x % cgs, prow = threadIdx.x / cgs;
  const bool has_b = g.b != nullptr, bn_a = g.coefA != nullptr, bn_b = g.coefB != nullptr;
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) red[i] = 0.f;
  for (int i = threadIdx.x; i < 4 * C; i += blockDim.x) {
    const int which = i / C;      // 0 sc, 1 sh, 2 mu, 3 is
    kA[i] = bn_a ? g.coefA[i] : (which == 0 ? 1.f : 0.f);
    kB[i] = bn_b ? g.coefB[i] : (which == 0 ? 1.f : 0.f);
  }
  __syncthreads();
  const long long p_begin = (long long)blockIdx.x * q.chunk;
  const long long p_end = min(p_begin + q.chunk, g.npix);
  const int ppb = q.ppb;
  const int c0 = cg * 4;
  // ---- pass 1
  if (q.phase == 1) {
    const float4 sca = *reinterpret_cast<const float4*>(kA + c0), sha = *reinterpret_cast<const float4*>(kA + C + c0);
    const float4 mua = *reinterpret_cast<const float4*>(kA + 2 * C + c0), isa = *reinterpret_cast<const float4*>(kA + 3 * C + c0);
    const float4 scb = *reinterpret_cast<const float4*>(kB + c0), shb = *reinterpret_cast<const float4*>(kB + C + c0);
    const float4 mub = *reinterpret_cast<const float4*>(kB + 2 * C + c0), isb = *reinterpret_cast<const float4*>(kB + 3 * C + c0);
    const float ksca[4] = {sca.x, sca.y, sca.z, sca.w}, ksha[4] = {sha.x, sha.y, sha.z, sha.w};
    const float kmua[4] = {mua.x, mua.y, mua.z, mua.w}, kisa[4] = {isa.x, isa.y, isa.z, isa.w};
    const float kscb[4] = {scb.x, scb.y, scb.z, scb.w}, kshb[4] = {shb.x, shb.y, shb.z, shb.w};
    const float kmub[4] = {mub.x, mub.y, mub.z, mub.w}, kisb[4] = {isb.x, isb.y, isb.z, isb.w};
    float s1[4] = {0, 0, 0, 0}, sa[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
    for (long long p = p_begin + prow; p < p_end; p += UU * ppb) {
      float4 a4[UU], b4[UU], d4[UU];
      bool ok[UU];
#pragma unroll
      for (int u = 0; u < UU; u++) {
        const long long pp = p + u * ppb;
        ok[u] = pp < p_end;
        a4[u] = b4[u] = d4[u] = make_float4(0, 0, 0, 0);
        if (ok[u]) {
          const long long off = pp * C + c0;
          a4[u] = *reinterpret_cast<const float4*>(g.a + off);
          if (has_b) b4[u] = *reinterpret_cast<const float4*>(g.b + off);
          d4[u] = *reinterpret_cast<const float4*>(q.dout + off);
        }
      }
#pragma unroll
      for (int u = 0; u < UU; u++) {
        if (!ok[u]) continue;
        const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w}, bv[4] = {b4[u].x, b4[u].y, b4[u].z, b4[u].w},
                    dv[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float pa = actf<PA>(g.preA, av[i]);
          float z = pa * ksca[i] + ksha[i];
          float pb = 0.f;
          if (has_b) { pb = actf<PB>(g.preB, bv[i]); z += pb * kscb[i] + kshb[i]; }
          const float dz = dv[i] * actb<PO>(g.post, z);
          s1[i] += dz;
          sa[i] += dz * (pa - kmua[i]) * kisa[i];
          if (has_b) sb[i] += dz * (pb - kmub[i]) * kisb[i];
        }
      }
    }
    // per-thread partials -> shared [pixel row][3C] -> column sums (float atomics on shared memory are CAS loops on this
    // architecture: 32 pixel rows contending for one address cost more than the whole reduction pass on small maps)
    float* part = kk + 5 * C;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      part[prow * 3 * C + c0 + i] = s1[i];
      part[prow * 3 * C + C + c0 + i] = sa[i];
      part[prow * 3 * C + 2 * C + c0 + i] = sb[i];
    }
    __syncthreads();
    const int rows = blockDim.x / cgs;
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
      float t = 0.f;
      for (int r = 0; r < rows; r++) t += part[r * 3 * C + i];
      atomicAdd(q.sums + (size_t)(blockIdx.x % BN_SLOTS) * 3 * C + i, (double)t);
    }
    return;
  }
  // ---- pass 2 constants: gi = gamma * invstd, batch means of dz and dz * xhat
  const float inv_n = 1.f / (float)g.npix;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double t1 = 0, t2a = 0, t2b = 0;
#pragma unroll
    for (int sl = 0; sl < BN_SLOTS; sl++) {
      const double* sp = q.sums + (size_t)sl * 3 * C;
      t1 += __ldcg(sp + c); t2a += __ldcg(sp + C + c); t2b += __ldcg(sp + 2 * C + c);
    }
    const float S1 = (float)t1, S2a = (float)t2a, S2b = (float)t2b;
    kk[c] = bn_a ? q.gammaA[c] * kA[3 * C + c] : 1.f;
    kk[C + c] = bn_b ? q.gammaB[c] * kB[3 * C + c] : 1.f;
    kk[2 * C + c] = S1 * inv_n;
    kk[3 * C + c] = S2a * inv_n;
    kk[4 * C + c] = S2b * inv_n;
    if (blockIdx.x == 0) {
      if (bn_a && q.dgammaA) { q.dgammaA[c] += S2a; q.dbetaA[c] += S1; }
      if (bn_b && q.dgammaB) { q.dgammaB[c] += S2b; q.dbetaB[c] += S1; }
    }
  }
  __syncthreads();
  {
    const long long span = p_end - p_begin - prow;
    const long long step = (long long)UU * ppb;
    long long iters = span > 0 ? (span + step - 1) / step : 0;
    for (long long it = iters - 1; it >= 0; it--) {
      const long long p = p_begin + prow + it * step;
      float4 a4[UU], b4[UU], d4[UU];
      bool ok[UU];
#pragma unroll
      for (int u = 0; u < UU; u++) {
        const long long pp = p + u * ppb;
        ok[u] = pp < p_end;
        a4[u] = b4[u] = d4[u] = make_float4(0, 0, 0, 0);
        if (ok[u]) {
          const long long off = pp * C + c0;
          a4[u] = __ldcs(reinterpret_cast<const float4*>(g.a + off));
          if (has_b) b4[u] = __ldcs(reinterpret_cast<const float4*>(g.b + off));
          d4[u] = __ldcs(reinterpret_cast<const float4*>(q.dout + off));
        }
      }
      const float4 sca = *reinterpret_cast<const float4*>(kA + c0), sha = *reinterpret_cast<const float4*>(kA + C + c0);
      const float4 mua = *reinterpret_cast<const float4*>(kA + 2 * C + c0), isa = *reinterpret_cast<const float4*>(kA + 3 * C + c0);
      const float4 gia4 = *reinterpret_cast<const float4*>(kk + c0), m14 = *reinterpret_cast<const float4*>(kk + 2 * C + c0);
      const float4 m2a4 = *reinterpret_cast<const float4*>(kk + 3 * C + c0);
      const float ksca[4] = {sca.x, sca.y, sca.z, sca.w}, ksha[4] = {sha.x, sha.y, sha.z, sha.w};
      const float kmua[4] = {mua.x, mua.y, mua.z, mua.w}, kisa[4] = {isa.x, isa.y, isa.z, isa.w};
      const float gia[4] = {gia4.x, gia4.y, gia4.z, gia4.w}, m1[4] = {m14.x, m14.y, m14.z, m14.w};
      const float m2a[4] = {m2a4.x, m2a4.y, m2a4.z, m2a4.w};
      float kscb[4] = {1, 1, 1, 1}, kshb[4] = {0, 0, 0, 0}, kmub[4] = {0, 0, 0, 0}, kisb[4] = {0, 0, 0, 0}, gib[4] = {1, 1, 1, 1},
            m2b[4] = {0, 0, 0, 0};
      if (has_b) {
        const float4 scb = *reinterpret_cast<const float4*>(kB + c0), shb = *reinterpret_cast<const float4*>(kB + C + c0);
        const float4 mub = *reinterpret_cast<const float4*>(kB + 2 * C + c0), isb = *reinterpret_cast<const float4*>(kB + 3 * C + c0);
        const float4 gib4 = *reinterpret_cast<const float4*>(kk + C + c0), m2b4 = *reinterpret_cast<const float4*>(kk + 4 * C + c0);
        kscb[0] = scb.x; kscb[1] = scb.y; kscb[2] = scb.z; kscb[3] = scb.w;
        kshb[0] = shb.x; kshb[1] = shb.y; kshb[2] = shb.z; kshb[3] = shb.w;
        kmub[0] = mub.x; kmub[1] = mub.y; kmub[2] = mub.z; kmub[3] = mub.w;
        kisb[0] = isb.x; kisb[1] = isb.y; kisb[2] = isb.z; kisb[3] = isb.w;
        gib[0] = gib4.x; gib[1] = gib4.y; gib[2] = gib4.z; gib[3] = gib4.w;
        m2b[0] = m2b4.x; m2b[1] = m2b4.y; m2b[2] = m2b4.z; m2b[3] = m2b4.w;
      }
#pragma unroll
      for (int u = 0; u < UU; u++) {
        if (!ok[u]) continue;
        const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w}, bv[4] = {b4[u].x, b4[u].y, b4[u].z, b4[u].w},
                    dv[4] = {d4[u].x, d4[u].y, d4[u].z, d4[u].w};
        float ra[4], rb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float pa = actf<PA>(g.preA, av[i]);
          float z = pa * ksca[i] + ksha[i];
          float pb = 0.f;
          if (has_b) { pb = actf<PB>(g.preB, bv[i]); z += pb * kscb[i] + kshb[i]; }
          const float dz = dv[i] * actb<PO>(g.post, z);
          const float ga = bn_a ? gia[i] * (dz - m1[i] - (pa - kmua[i]) * kisa[i] * m2a[i]) : dz;
          ra[i] = ga * actb<PA>(g.preA, av[i]);
          if (has_b) {
            const float gb = bn_b ? gib[i] * (dz - m1[i] - (pb - kmub[i]) * kisb[i] * m2b[i]) : dz;
            rb[i] = gb * actb<PB>(g.preB, bv[i]);
          }
        }
        const long long off = (p + u * ppb) * C + c0;
        *reinterpret_cast<float4*>(q.da + off) = make_float4(ra[0], ra[1], ra[2], ra[3]);
        if (has_b && q.db) *reinterpret_cast<float4*>(q.db + off) = make_float4(rb[0], rb[1], rb[2], rb[3]);
      }
    }
  }
}

template <int PA, int PB, int PO, int UU = 4, int MINB = 2>
static int launch_bn_bwd_fused(BnBwdArgs& q, const CgMap& m, cudaStream_t st) {
  const size_t smem = ((size_t)16 * q.g.C + (size_t)(m.threads / (q.g.C / 4)) * 3 * q.g.C) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(bn_act2_bwd_fused_kernel<PA, PB, PO, UU, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_act2_bwd_fused_kernel<PA, PB, PO, UU, MINB>, m.threads, smem);
  if (per_sm < 1) per_sm = 1;
  long long blocks = (q.g.npix + m.ppb - 1) / m.ppb;
  const long long cap = (long long)tcct_num_sms() * per_sm;
  int grid = (int)(blocks < cap ? blocks : cap);
  if (grid < 1) grid = 1;
  long long chunk = (q.g.npix + grid - 1) / grid;
  chunk = (chunk + m.ppb - 1) / m.ppb * m.ppb;
  grid = (int)((q.g.npix + chunk - 1) / chunk);
  q.chunk = chunk; q.ppb = m.ppb;
  static const int only = [] { const char* e = getenv("TCCT_BN_BWD_PHASE"); return e ? atoi(e) : 0; }();     // timing experiments only
  q.phase = 1;
  if (only != 2) bn_act2_bwd_fused_kernel<PA, PB, PO, UU, MINB><<<grid, m.threads, smem, st>>>(q);
  q.phase = 2;
  if (only != 1) bn_act2_bwd_fused_kernel<PA, PB, PO, UU, MINB><<<grid, m.threads, smem, st>>>(q);
  tcct_count_launch();
  return grid;
}

// sums: zeroed double[8*3*C + 1] workspace (8 replicas of the batch sums, then the grid-barrier counter; ignored when neither operand is
// batch-normalised in train mode;
// pass sums = null for eval-mode BN: statistics are constants, no correction terms).
extern "C" int tcct_bn_act2_bwd(const float* a, const float* coefA, int preA, const float* gammaA, const float* b,
                                const float* coefB, int preB, const float* gammaB, int post, const float* dout,
                                double* sums, float* da, float* db, float* dgammaA, float* dbetaA, float* dgammaB,
                                float* dbetaB, long long npix, int C, void* stream) {
  TCCT_CHECK_ARG(C % 4 == 0 && C <= 1024, "bn_act2_bwd: C must be a multiple of 4, <= 1024 (got %d)", C);
  Bn2Args g{a, coefA, preA, b, coefB, preB, post, npix, C};
  const CgMap m = cg_map(C);
  const bool hot = preA == ACT_LRELU && preB == ACT_LRELU && post == ACT_GELU && b;
  cudaStream_t st = (cudaStream_t)stream;
  if (sums && (coefA || coefB)) {
    BnBwdArgs q{g, dout, sums, gammaA, gammaB, da, db, dgammaA, dbetaA, dgammaB, dbetaB, 0, 0, 0};
    // The activation kinds are template parameters for every combination the networks use: with run-time kinds (ACT_DYN) the compiler
    // evaluates all four activations -- GELU's exponential and divide included -- for every element and selects, which makes the
    // pass compute-bound (8x128x128x64: 51-60 us against 30 us; scripts/time_bncfg.py).  ACT_DYN remains for anything else.
    const bool pre_none = preA == ACT_NONE && (!b || preB == ACT_NONE);
    if (hot) launch_bn_bwd_fused<ACT_LRELU, ACT_LRELU, ACT_GELU>(q, m, st);
    else if (preA == ACT_LRELU && !b && post == ACT_NONE) launch_bn_bwd_fused<ACT_LRELU, ACT_NONE, ACT_NONE>(q, m, st);
    else if (pre_none && post == ACT_HSWISH) launch_bn_bwd_fused<ACT_NONE, ACT_NONE, ACT_HSWISH>(q, m, st);      // MPViT Conv2d_BN / DWConv2d_BN
    else if (pre_none && post == ACT_LRELU) launch_bn_bwd_fused<ACT_NONE, ACT_NONE, ACT_LRELU>(q, m, st);        // MPUpBlock.prep, FTC.head
    else if (pre_none && post == ACT_NONE) launch_bn_bwd_fused<ACT_NONE, ACT_NONE, ACT_NONE>(q, m, st);          // stems, tran_*, ResBlock.conv2 + residual
    else launch_bn_bwd_fused<ACT_DYN, ACT_DYN, ACT_DYN>(q, m, st);
    TCCT_CHECK_LAUNCH("bn_act2_bwd_fused");
    return TCCT_OK;
  }
  sums = nullptr;
  const int grid = grid_for(npix, m.ppb, 8);
  if (hot)
    bn_act2_bwd_apply_kernel<ACT_LRELU, ACT_LRELU, ACT_GELU><<<grid, m.threads, 0, st>>>(g, dout, sums, gammaA, gammaB, m.ppb, da, db, dgammaA,
                                                                                         dbetaA, dgammaB, dbetaB);
  else if (preA == ACT_NONE && (!b || preB == ACT_NONE) && post == ACT_GELU)       // the MLP's activation pass (unfused path)
    bn_act2_bwd_apply_kernel<ACT_NONE, ACT_NONE, ACT_GELU><<<grid, m.threads, 0, st>>>(g, dout, sums, gammaA, gammaB, m.ppb, da, db, dgammaA,
                                                                                       dbetaA, dgammaB, dbetaB);
  else
    bn_act2_bwd_apply_kernel<ACT_DYN, ACT_DYN, ACT_DYN><<<grid, m.threads, 0, st>>>(g, dout, sums, gammaA, gammaB, m.ppb, da, db, dgammaA, dbetaA,
                                                                                    dgammaB, dbetaB);
  TCCT_CHECK_LAUNCH("bn_act2_bwd_apply");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// MaxPool2d(2) (tcct.py:867,883).  Backward routes to the first maximum in window scan order.
// ----------------------------------------------------------------------------------------------
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const int Ho = H >> 1, Wo = W >> 1, c4 = C >> 2;
  const long long n = (long long)B * Ho * Wo * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int c, ox, oy, b;
    split4(i, c4, Wo, Ho, c, ox, oy, b);
    const float4* r0 = reinterpret_cast<const float4*>(x + (((size_t)b * H + 2 * oy) * W + 2 * ox) * C) + c;
    const float4* r1 = reinterpret_cast<const float4*>(x + (((size_t)b * H + 2 * oy + 1) * W + 2 * ox) * C) + c;
    const float4 v0 = r0[0], v1 = r0[c4], v2 = r1[0], v3 = r1[c4];
    float4 o;
    o.x = fmaxf(fmaxf(v0.x, v1.x), fmaxf(v2.x, v3.x));
    o.y = fmaxf(fmaxf(v0.y, v1.y), fmaxf(v2.y, v3.y));
    o.z = fmaxf(fmaxf(v0.z, v1.z), fmaxf(v2.z, v3.z));
    o.w = fmaxf(fmaxf(v0.w, v1.w), fmaxf(v2.w, v3.w));
    reinterpret_cast<float4*>(y)[i] = o;
  }
}

__device__ __forceinline__ void route4(float v0, float v1, float v2, float v3, float d, float& o0, float& o1,
                                       float& o2, float& o3) {
  int k = 0; float m = v0;
  if (v1 > m) { m = v1; k = 1; }
  if (v2 > m) { m = v2; k = 2; }
  if (v3 > m) { m = v3; k = 3; }
  o0 = k == 0 ? d : 0.f; o1 = k == 1 ? d : 0.f; o2 = k == 2 ? d : 0.f; o3 = k == 3 ? d : 0.f;
}

__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                    int B, int H, int W, int C) {
  const int Ho = H >> 1, Wo = W >> 1, c4 = C >> 2;
  const long long n = (long long)B * Ho * Wo * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int c, ox, oy, b;
    split4(i, c4, Wo, Ho, c, ox, oy, b);
    const size_t o0 = (((size_t)b * H + 2 * oy) * W + 2 * ox) * C, o1 = o0 + (size_t)W * C;
    const float4* r0 = reinterpret_cast<const float4*>(x + o0) + c;
    const float4* r1 = reinterpret_cast<const float4*>(x + o1) + c;
    const float4 v0 = r0[0], v1 = r0[c4], v2 = r1[0], v3 = r1[c4];
    const float4 d = reinterpret_cast<const float4*>(dy)[i];
    float4 g0, g1, g2, g3;
    route4(v0.x, v1.x, v2.x, v3.x, d.x, g0.x, g1.x, g2.x, g3.x);
    route4(v0.y, v1.y, v2.y, v3.y, d.y, g0.y, g1.y, g2.y, g3.y);
    route4(v0.z, v1.z, v2.z, v3.z, d.z, g0.z, g1.z, g2.z, g3.z);
    route4(v0.w, v1.w, v2.w, v3.w, d.w, g0.w, g1.w, g2.w, g3.w);
    float4* w0 = reinterpret_cast<float4*>(dx + o0) + c;
    float4* w1 = reinterpret_cast<float4*>(dx + o1) + c;
    w0[0] = g0; w0[c4] = g1; w1[0] = g2; w1[c4] = g3;
  }
}

extern "C" int tcct_maxpool2_fwd(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  TCCT_CHECK_ARG((long long)B * H * W * C < (1ll << 33), "maxpool2: tensor too large for 32-bit pixel indices");
  TCCT_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "maxpool2: H, W must be even and C a multiple of 4");
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 4);
  maxpool2_fwd_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, y, B, H, W, C);
  TCCT_CHECK_LAUNCH("maxpool2_fwd");
  return TCCT_OK;
}
extern "C" int tcct_maxpool2_bwd(const float* x, const float* dy, float* dx, int B, int H, int W, int C, void* stream) {
  TCCT_CHECK_ARG((long long)B * H * W * C < (1ll << 33), "maxpool2: tensor too large for 32-bit pixel indices");
  TCCT_CHECK_ARG(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "maxpool2: H, W must be even and C a multiple of 4");
  const long long n = (long long)B * (H / 2) * (W / 2) * (C / 4);
  maxpool2_bwd_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, B, H, W, C);
  TCCT_CHECK_LAUNCH("maxpool2_bwd");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// Depthwise 3x3, pad 1, stride 1|2 (DWConv2d_BN tcct.py:99-147, ResBlock dwconv 535-543,
// ConvPosEnc 197-217 with add_input: y = dw(x) + bias + x).  w: [C][9] (PyTorch [C,1,3,3]).
// ----------------------------------------------------------------------------------------------
__global__ void dwconv3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                   float* __restrict__ y, int B, int H, int W, int C, int stride, int add_input,
                                   int ppb, double* stats) {
  extern __shared__ float sred[];
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int cgs = C >> 2, cg = threadIdx.x % cgs, prow = threadIdx.x / cgs;
  if (stats) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();
  }
  float wr[4][9];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 9; k++) wr[i][k] = w[(cg * 4 + i) * 9 + k];
  float bs[4] = {0, 0, 0, 0};
  if (bias) { bs[0] = bias[cg * 4]; bs[1] = bias[cg * 4 + 1]; bs[2] = bias[cg * 4 + 2]; bs[3] = bias[cg * 4 + 3]; }
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  const long long npix = (long long)B * Ho * Wo;
  for (long long p = (long long)blockIdx.x * ppb + prow; p < npix; p += (long long)gridDim.x * ppb) {
    int ox, oy, b;
    split3(p, Wo, Ho, ox, oy, b);
    float o[4] = {bs[0], bs[1], bs[2], bs[3]};
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      const int iy = oy * stride + ky - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int ix = ox * stride + kx - 1;
        if (ix < 0 || ix >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(x + (((size_t)b * H + iy) * W + ix) * C + cg * 4);
        o[0] += v.x * wr[0][ky * 3 + kx]; o[1] += v.y * wr[1][ky * 3 + kx];
        o[2] += v.z * wr[2][ky * 3 + kx]; o[3] += v.w * wr[3][ky * 3 + kx];
        if (add_input && ky == 1 && kx == 1) { o[0] += v.x; o[1] += v.y; o[2] += v.z; o[3] += v.w; }
      }
    }
    *reinterpret_cast<float4*>(y + p * C + cg * 4) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
    for (int i = 0; i < 4; i++) { s[i] += o[i]; q[i] += o[i] * o[i]; }
  }
  if (stats) {      // per-thread partials -> shared [pixel row][2C] -> column sums (no shared-memory float atomics)
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
      sred[(size_t)prow * 2 * C + cg * 4 + i] = s[i];
      sred[(size_t)prow * 2 * C + C + cg * 4 + i] = q[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      float t = 0.f;
      for (int r = 0; r < ppb; r++) t += sred[(size_t)r * 2 * C + i];
      atomicAdd(stats + i, (double)t);
    }
  }
}

// dx[iy][ix] = sum_{ky,kx : (iy+1-ky)%s==0, ...} w[ky][kx] * dy[(iy+1-ky)/s][(ix+1-kx)/s]   (+ dy if add_input)
// Thread (pixel row, channel group) keeps the 4 x 9 weights of its channels in registers; S is a compile-time stride so
// that the "which outputs read this input" tests are bit operations instead of integer divisions.
template <int S>
__global__ void __launch_bounds__(256) dwconv3_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                               float* __restrict__ dx, int B, int H, int W, int C, int add_input,
                                                               int ppb) {
  const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1;
  const int cgs = C >> 2, cg = threadIdx.x % cgs, prow = threadIdx.x / cgs;
  float wr[4][9];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 9; k++) wr[i][k] = w[(cg * 4 + i) * 9 + k];
  if (add_input) { wr[0][4] += 1.f; wr[1][4] += 1.f; wr[2][4] += 1.f; wr[3][4] += 1.f; }
  const long long npix = (long long)B * H * W;
  for (long long p = (long long)blockIdx.x * ppb + prow; p < npix; p += (long long)gridDim.x * ppb) {
    int ix, iy, b;
    split3(p, W, H, ix, iy, b);
    float o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      const int ty = iy + 1 - ky;
      if (ty < 0 || (ty % S)) continue;
      const int oy = ty / S;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int tx = ix + 1 - kx;
        if (tx < 0 || (tx % S)) continue;
        const int ox = tx / S;
        if (ox >= Wo) continue;
        const float4 d = *reinterpret_cast<const float4*>(dy + (((size_t)b * Ho + oy) * Wo + ox) * C + cg * 4);
        const int k = ky * 3 + kx;
        o[0] += d.x * wr[0][k]; o[1] += d.y * wr[1][k]; o[2] += d.z * wr[2][k]; o[3] += d.w * wr[3][k];
      }
    }
    *reinterpret_cast<float4*>(dx + p * C + cg * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// dw[c][k] += sum_p dy[p][c] * x[p*s + k - 1][c];  dbias[c] += sum_p dy[p][c]
__global__ void dwconv3_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* dw,
                                          float* dbias, int B, int H, int W, int C, int stride, int ppb) {
  extern __shared__ float sred[];     // [10*C]
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int cgs = C >> 2, cg = threadIdx.x % cgs, prow = threadIdx.x / cgs;
  for (int i = threadIdx.x; i < 10 * C; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float acc[4][10];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 10; k++) acc[i][k] = 0.f;
  const long long npix = (long long)B * Ho * Wo;
  for (long long p = (long long)blockIdx.x * ppb + prow; p < npix; p += (long long)gridDim.x * ppb) {
    int ox, oy, b;
    split3(p, Wo, Ho, ox, oy, b);
    const float4 d4 = *reinterpret_cast<const float4*>(dy + p * C + cg * 4);
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
    for (int i = 0; i < 4; i++) acc[i][9] += d[i];
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      const int iy = oy * stride + ky - 1;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int ix = ox * stride + kx - 1;
        if (ix < 0 || ix >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(x + (((size_t)b * H + iy) * W + ix) * C + cg * 4);
        acc[0][ky * 3 + kx] += d[0] * v.x; acc[1][ky * 3 + kx] += d[1] * v.y;
        acc[2][ky * 3 + kx] += d[2] * v.z; acc[3][ky * 3 + kx] += d[3] * v.w;
      }
    }
  }
  // per-thread partials -> shared [pixel row][10C] -> column sums (no shared-memory float atomics: they are CAS loops)
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 10; k++) sred[(size_t)prow * 10 * C + (cg * 4 + i) * 10 + k] = acc[i][k];
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * C; i += blockDim.x) {
    float v = 0.f;
    for (int r = 0; r < ppb; r++) v += sred[(size_t)r * 10 * C + i];
    const int c = i / 10, k = i % 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, v);
    else if (dbias) atomicAdd(dbias + c, v);
  }
}

// ---- stride-1 fast path: a block owns a TX-wide, DW_ROWS-tall strip of one image; thread (tx, cg) walks down its column
// keeping the 3x3 input window of its 4 channels in registers (3 new float4 loads per output; the horizontal
// neighbours are L1 hits of the same block).  FLIP selects the transposed stencil, i.e. the data gradient.
template <bool FLIP, int DW_ROWS>
__global__ void __launch_bounds__(256, 2) dwconv3_s1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, int H, int W,
                                                         int C, int TX, int add_input, double* stats) {
  extern __shared__ float sred[];
  const int cgs = C >> 2, cg = threadIdx.x % cgs, tx = threadIdx.x / cgs;
  const bool live = tx < TX;
  const int ox = blockIdx.x * TX + tx, y0 = blockIdx.y * DW_ROWS, b = blockIdx.z;
  if (stats) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
    __syncthreads();
  }
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (live && ox < W) {
    float wr[4][9];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int k = 0; k < 9; k++) wr[i][k] = w[(cg * 4 + i) * 9 + (FLIP ? 8 - k : k)];
    float bs[4] = {0, 0, 0, 0};
    if (bias) { bs[0] = bias[cg * 4]; bs[1] = bias[cg * 4 + 1]; bs[2] = bias[cg * 4 + 2]; bs[3] = bias[cg * 4 + 3]; }
    const float4 z4 = make_float4(0, 0, 0, 0);
    const float* xb = x + (size_t)b * H * W * C + cg * 4;
    auto load_row = [&](int iy, float4& l, float4& m, float4& r) {
      if (iy < 0 || iy >= H) { l = m = r = z4; return; }
      const float* row = xb + (size_t)iy * W * C;
      m = *reinterpret_cast<const float4*>(row + (size_t)ox * C);
      l = ox > 0 ? *reinterpret_cast<const float4*>(row + (size_t)(ox - 1) * C) : z4;
      r = ox + 1 < W ? *reinterpret_cast<const float4*>(row + (size_t)(ox + 1) * C) : z4;
    };
    float4 a0, a1, a2, b0, b1, b2, c0, c1, c2, n0, n1, n2;      // rows oy-1, oy, oy+1 and the prefetched oy+2
    const int y1 = min(y0 + DW_ROWS, H);
    load_row(y0 - 1, a0, a1, a2);
    load_row(y0, b0, b1, b2);
    load_row(y0 + 1, c0, c1, c2);
    for (int oy = y0; oy < y1; oy++) {
      load_row(oy + 2 <= y1 ? oy + 2 : H, n0, n1, n2);          // one row ahead of the one this iteration consumes
      float o[4] = {bs[0], bs[1], bs[2], bs[3]};
#define DW_TAP(v, k) o[0] += v.x * wr[0][k]; o[1] += v.y * wr[1][k]; o[2] += v.z * wr[2][k]; o[3] += v.w * wr[3][k];
      DW_TAP(a0, 0) DW_TAP(a1, 1) DW_TAP(a2, 2) DW_TAP(b0, 3) DW_TAP(b1, 4) DW_TAP(b2, 5) DW_TAP(c0, 6) DW_TAP(c1, 7) DW_TAP(c2, 8)
#undef DW_TAP
      if (add_input) { o[0] += b1.x; o[1] += b1.y; o[2] += b1.z; o[3] += b1.w; }
      *reinterpret_cast<float4*>(y + (((size_t)b * H + oy) * W + ox) * C + cg * 4) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
      for (int i = 0; i < 4; i++) { s[i] += o[i]; q[i] += o[i] * o[i]; }
      a0 = b0; a1 = b1; a2 = b2; b0 = c0; b1 = c1; b2 = c2; c0 = n0; c1 = n1; c2 = n2;
    }
  }
  if (stats) {      // per-thread partials (zero for idle threads) -> shared [pixel column][2C] -> column sums
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) {
      sred[(size_t)tx * 2 * C + cg * 4 + i] = s[i];
      sred[(size_t)tx * 2 * C + C + cg * 4 + i] = q[i];
    }
    __syncthreads();
    const int rows = blockDim.x / cgs;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
      float t = 0.f;
      for (int r = 0; r < rows; r++) t += sred[(size_t)r * 2 * C + i];
      atomicAdd(stats + i, (double)t);
    }
  }
}

// weight gradient, stride 1: same strip walk; per-thread accumulators for 4 channels x (9 taps + bias)
template <int DW_ROWS>
__global__ void __launch_bounds__(256, 2) dwconv3_s1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* dw,
                                                               float* dbias, int H, int W, int C, int TX) {
  extern __shared__ float sred[];     // [pixel columns of the block][10*C]: per-thread partials, summed without atomics
  const int cgs = C >> 2, cg = threadIdx.x % cgs, tx = threadIdx.x / cgs;
  const int ox = blockIdx.x * TX + tx, y0 = blockIdx.y * DW_ROWS, b = blockIdx.z;
  float acc[4][10];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 10; k++) acc[i][k] = 0.f;
  if (tx < TX && ox < W) {
    const float4 z4 = make_float4(0, 0, 0, 0);
    const float* xb = x + (size_t)b * H * W * C + cg * 4;
    auto load_row = [&](int iy, float4& l, float4& m, float4& r) {
      if (iy < 0 || iy >= H) { l = m = r = z4; return; }
      const float* row = xb + (size_t)iy * W * C;
      m = *reinterpret_cast<const float4*>(row + (size_t)ox * C);
      l = ox > 0 ? *reinterpret_cast<const float4*>(row + (size_t)(ox - 1) * C) : z4;
      r = ox + 1 < W ? *reinterpret_cast<const float4*>(row + (size_t)(ox + 1) * C) : z4;
    };
    // software pipeline: the loads of input row oy+2 and of dy row oy+1 are in flight while row oy is accumulated
    float4 a0, a1, a2, b0, b1, b2, c0, c1, c2, n0, n1, n2;
    const int y1 = min(y0 + DW_ROWS, H);
    const float* dyb = dy + ((size_t)b * H * W + ox) * C + cg * 4;
    load_row(y0 - 1, a0, a1, a2);
    load_row(y0, b0, b1, b2);
    load_row(y0 + 1, c0, c1, c2);
    float4 d = *reinterpret_cast<const float4*>(dyb + (size_t)y0 * W * C), dn = z4;
    for (int oy = y0; oy < y1; oy++) {
      load_row(oy + 2 <= y1 ? oy + 2 : H, n0, n1, n2);
      if (oy + 1 < y1) dn = *reinterpret_cast<const float4*>(dyb + (size_t)(oy + 1) * W * C);
      acc[0][9] += d.x; acc[1][9] += d.y; acc[2][9] += d.z; acc[3][9] += d.w;
#define DW_ACC(v, k) acc[0][k] += d.x * v.x; acc[1][k] += d.y * v.y; acc[2][k] += d.z * v.z; acc[3][k] += d.w * v.w;
      DW_ACC(a0, 0) DW_ACC(a1, 1) DW_ACC(a2, 2) DW_ACC(b0, 3) DW_ACC(b1, 4) DW_ACC(b2, 5) DW_ACC(c0, 6) DW_ACC(c1, 7) DW_ACC(c2, 8)
#undef DW_ACC
      a0 = b0; a1 = b1; a2 = b2; b0 = c0; b1 = c1; b2 = c2; c0 = n0; c1 = n1; c2 = n2; d = dn;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int k = 0; k < 10; k++) sred[(size_t)tx * 10 * C + (cg * 4 + i) * 10 + k] = acc[i][k];
  __syncthreads();
  const int rows = blockDim.x / cgs;
  for (int i = threadIdx.x; i < 10 * C; i += blockDim.x) {
    float v = 0.f;
    for (int r = 0; r < rows; r++) v += sred[(size_t)r * 10 * C + i];
    const int c = i / 10, k = i % 10;
    if (k < 9) atomicAdd(dw + c * 9 + k, v);
    else if (dbias) atomicAdd(dbias + c, v);
  }
}


struct DwTile { int tx, threads, rows; dim3 grid; };
static DwTile dw_tile(int B, int H, int W, int C) {
  DwTile t;
  const int cgs = C / 4;
  t.tx = 256 / cgs;
  if (t.tx > W) t.tx = W;
  t.threads = t.tx * cgs;
  // 16-row strips amortise the two halo rows; on the small maps shorter strips make enough blocks to occupy the GPU
  // (8x32x32x128: 7.6 -> 5.6 us with 8 rows, 8x16x16x160: 7.4 -> 4.2 us with 4)
  t.rows = 16;
  while (t.rows > 4 && ceil_div(W, t.tx) * ceil_div(H, t.rows) * B < 128) t.rows >>= 1;
  t.grid = dim3(ceil_div(W, t.tx), ceil_div(H, t.rows), B);
  return t;
}
#define DW_DISPATCH(rows, LAUNCH) \
  switch (rows) { case 4: { constexpr int R = 4; LAUNCH; } break; case 8: { constexpr int R = 8; LAUNCH; } break; \
                  default: { constexpr int R = 16; LAUNCH; } }

extern "C" int tcct_dwconv3_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                                int C, int stride, int add_input, double* stats, void* stream) {
  TCCT_CHECK_ARG((long long)B * H * W < (1ll << 31), "dwconv3: too many pixels for 32-bit indices");
  TCCT_CHECK_ARG(C % 4 == 0 && C <= 1024 && (stride == 1 || stride == 2), "dwconv3: C %% 4 == 0 and stride 1|2 expected");
  if (stride == 1) {
    const DwTile t = dw_tile(B, H, W, C);
    DW_DISPATCH(t.rows, (dwconv3_s1_kernel<false, R><<<t.grid, t.threads, (size_t)8 * t.threads * sizeof(float), (cudaStream_t)stream>>>(
                            x, w, bias, y, H, W, C, t.tx, add_input, stats)));
    TCCT_CHECK_LAUNCH("dwconv3_s1_fwd");
    return TCCT_OK;
  }
  const CgMap m = cg_map(C);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  dwconv3_fwd_kernel<<<grid_for((long long)B * Ho * Wo, m.ppb, 8), m.threads, (size_t)8 * m.threads * sizeof(float),
                       (cudaStream_t)stream>>>(x, w, bias, y, B, H, W, C, stride, add_input, m.ppb, stats);
  TCCT_CHECK_LAUNCH("dwconv3_fwd");
  return TCCT_OK;
}
extern "C" int tcct_dwconv3_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias,
                                int B, int H, int W, int C, int stride, int add_input, void* stream) {
  TCCT_CHECK_ARG((long long)B * H * W < (1ll << 31), "dwconv3: too many pixels for 32-bit indices");
  TCCT_CHECK_ARG(C % 4 == 0 && C <= 1024 && (stride == 1 || stride == 2), "dwconv3: C %% 4 == 0 and stride 1|2 expected");
  const CgMap m = cg_map(C);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  if (stride == 1) {
    const DwTile t = dw_tile(B, H, W, C);
    if (dx) {     // the data gradient of a stride-1 'same' correlation is the correlation with the flipped stencil
      DW_DISPATCH(t.rows, (dwconv3_s1_kernel<true, R><<<t.grid, t.threads, (size_t)8 * t.threads * sizeof(float), (cudaStream_t)stream>>>(
                              dy, w, nullptr, dx, H, W, C, t.tx, add_input, nullptr)));
      TCCT_CHECK_LAUNCH("dwconv3_s1_bwd_data");
    }
    if (dw) {
      DW_DISPATCH(t.rows, (dwconv3_s1_wgrad_kernel<R><<<t.grid, t.threads, (size_t)40 * t.threads * sizeof(float), (cudaStream_t)stream>>>(
                              x, dy, dw, dbias, H, W, C, t.tx)));
      TCCT_CHECK_LAUNCH("dwconv3_s1_wgrad");
    }
    return TCCT_OK;
  }
  if (dx) {
    const int grid = grid_for((long long)B * H * W, m.ppb, 8);
    if (stride == 2) dwconv3_bwd_data_kernel<2><<<grid, m.threads, 0, (cudaStream_t)stream>>>(dy, w, dx, B, H, W, C, add_input, m.ppb);
    else dwconv3_bwd_data_kernel<1><<<grid, m.threads, 0, (cudaStream_t)stream>>>(dy, w, dx, B, H, W, C, add_input, m.ppb);
    TCCT_CHECK_LAUNCH("dwconv3_bwd_data");
  }
  if (dw) {
    dwconv3_bwd_weight_kernel<<<grid_for((long long)B * Ho * Wo, m.ppb, 4), m.threads, (size_t)40 * m.threads * sizeof(float),
                                (cudaStream_t)stream>>>(x, dy, dw, dbias, B, H, W, C, stride, m.ppb);
    TCCT_CHECK_LAUNCH("dwconv3_bwd_weight");
  }
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// LayerNorm over C (eps 1e-6, tcct.py:427,454-455): one warp per token, lanes own channels lane+32i.
// ----------------------------------------------------------------------------------------------
#define LN_MAXI 8   // C <= 256
template <int LN_NI>
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean_rstd,
                                     long long ntok, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long tk = warp0; tk < ntok; tk += nwarp) {
    float v[LN_NI];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_NI; i++) {
      const int c = lane + 32 * i;
      v[i] = c < C ? x[tk * C + c] : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_NI; i++) {
      const int c = lane + 32 * i;
      const float d = c < C ? v[i] - mean : 0.f;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / C + eps);
#pragma unroll
    for (int i = 0; i < LN_NI; i++) {
      const int c = lane + 32 * i;
      if (c < C) y[tk * C + c] = (v[i] - mean) * rstd * gamma[c] + beta[c];
    }
    if (lane == 0) { mean_rstd[2 * tk] = mean; mean_rstd[2 * tk + 1] = rstd; }
  }
}

template <int LN_NI>
__global__ void layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ mean_rstd, const float* __restrict__ dy,
                                     float* __restrict__ dx, float* dgamma, float* dbeta, long long ntok, int C) {
  extern __shared__ float sred[];     // [2*C]
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float dg[LN_NI], dbt[LN_NI], gm[LN_NI];
#pragma unroll
  for (int i = 0; i < LN_NI; i++) {
    dg[i] = dbt[i] = 0.f;
    const int c = lane + 32 * i;
    gm[i] = c < C ? gamma[c] : 0.f;
  }
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long tk = warp0; tk < ntok; tk += nwarp) {
    const float mean = mean_rstd[2 * tk], rstd = mean_rstd[2 * tk + 1];
    float xh[LN_NI], d[LN_NI];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_NI; i++) {
      const int c = lane + 32 * i;
      if (c < C) {
        xh[i] = (x[tk * C + c] - mean) * rstd;
        const float g = dy[tk * C + c];
        dg[i] += g * xh[i]; dbt[i] += g;
        d[i] = g * gm[i];
        s1 += d[i]; s2 += d[i] * xh[i];
      } else { xh[i] = 0.f; d[i] = 0.f; }
    }
    s1 = warp_sum(s1) / C; s2 = warp_sum(s2) / C;
#pragma unroll
    for (int i = 0; i < LN_NI; i++) {
      const int c = lane + 32 * i;
      if (c < C) dx[tk * C + c] = rstd * (d[i] - s1 - xh[i] * s2);
    }
  }
#pragma unroll
  for (int i = 0; i < LN_NI; i++) {
    const int c = lane + 32 * i;
    if (c < C) { atomicAdd(&sred[c], dg[i]); atomicAdd(&sred[C + c], dbt[i]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, sred[i]);
    atomicAdd(dbeta + i, sred[C + i]);
  }
}

extern "C" int tcct_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean_rstd,
                                  long long ntok, int C, float eps, void* stream) {
  TCCT_CHECK_ARG(C <= 32 * LN_MAXI, "layernorm: C too large (%d)", C);
  const int ni = (C + 31) / 32;
  const dim3 grid(grid_for(ntok, 8, 8));
#define LN_FWD(NI) layernorm_fwd_kernel<NI><<<grid, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean_rstd, ntok, C, eps)
  switch (ni) {
    case 1: LN_FWD(1); break; case 2: LN_FWD(2); break; case 3: LN_FWD(3); break; case 4: LN_FWD(4); break;
    case 5: LN_FWD(5); break; case 6: LN_FWD(6); break; case 7: LN_FWD(7); break; default: LN_FWD(8); break;
  }
#undef LN_FWD
  TCCT_CHECK_LAUNCH("layernorm_fwd");
  return TCCT_OK;
}
extern "C" int tcct_layernorm_bwd(const float* x, const float* gamma, const float* mean_rstd, const float* dy,
                                  float* dx, float* dgamma, float* dbeta, long long ntok, int C, void* stream) {
  TCCT_CHECK_ARG(C <= 32 * LN_MAXI, "layernorm: C too large (%d)", C);
  const int ni = (C + 31) / 32;
  const dim3 grid(grid_for(ntok, 8, 8));
#define LN_BWD(NI) layernorm_bwd_kernel<NI><<<grid, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>(x, gamma, mean_rstd, dy, dx, dgamma, dbeta, ntok, C)
  switch (ni) {
    case 1: LN_BWD(1); break; case 2: LN_BWD(2); break; case 3: LN_BWD(3); break; case 4: LN_BWD(4); break;
    case 5: LN_BWD(5); break; case 6: LN_BWD(6); break; case 7: LN_BWD(7); break; default: LN_BWD(8); break;
  }
#undef LN_BWD
  TCCT_CHECK_LAUNCH("layernorm_bwd");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// MetaPool token mixer with residual (MHCABlock.forward tcct.py:457-469, MetaPool 405-415):
//   out[b,n,c] = t[b,n,c] + s[b] * ( avg3x3_{(n,c) plane, valid count}(cur)[n,c] - cur[n,c] )
// The 3x3 window spans neighbouring TOKENS and CHANNELS (AvgPool2d applied to a 3-D [B,N,C] tensor).
// ----------------------------------------------------------------------------------------------
// One thread per (token, 4-channel group): three coalesced float4 row loads give the vertical sums of its own
// channels; the two neighbouring channels come from the adjacent lanes (shuffles), or from three scalar loads at
// warp edges.  BWD = false: plain sums (forward); BWD = true: every term pre-divided by its own window count.
template <bool BWD>
__device__ __forceinline__ float4 metapool_window(const float* __restrict__ src, long long row0, int tk, int N, int C, int cg,
                                                  int lane, float4& centre) {
  const float4 z = make_float4(0, 0, 0, 0);
  const int c4 = cg * 4;
  const float4 r1 = *reinterpret_cast<const float4*>(src + (row0 + tk) * C + c4);
  const float4 r0 = tk > 0 ? *reinterpret_cast<const float4*>(src + (row0 + tk - 1) * C + c4) : z;
  const float4 r2 = tk + 1 < N ? *reinterpret_cast<const float4*>(src + (row0 + tk + 1) * C + c4) : z;
  centre = r1;
  // row weights: forward 1; backward 1/rn(n') with rn = valid rows of the window centred on n'
  float w0 = 1.f, w1 = 1.f, w2 = 1.f;
  if (BWD) {
    const auto rn = [N](int n) { return (float)(min(n + 1, N - 1) - max(n - 1, 0) + 1); };
    w0 = tk > 0 ? 1.f / rn(tk - 1) : 0.f; w1 = 1.f / rn(tk); w2 = tk + 1 < N ? 1.f / rn(tk + 1) : 0.f;
  }
  float4 v = make_float4(w0 * r0.x + w1 * r1.x + w2 * r2.x, w0 * r0.y + w1 * r1.y + w2 * r2.y,
                         w0 * r0.z + w1 * r1.z + w2 * r2.z, w0 * r0.w + w1 * r1.w + w2 * r2.w);
  if (BWD) {   // column weights 1/rc(c')
    const auto rc = [C](int c) { return 1.f / (float)(min(c + 1, C - 1) - max(c - 1, 0) + 1); };
    v.x *= rc(c4); v.y *= rc(c4 + 1); v.z *= rc(c4 + 2); v.w *= rc(c4 + 3);
  }
  float vl = __shfl_up_sync(0xffffffffu, v.w, 1), vr = __shfl_down_sync(0xffffffffu, v.x, 1);
  const int cgs = C >> 2;
  auto edge = [&](int c) {        // vertical (weighted) sum of a single channel, straight from memory
    float e = w1 * src[(row0 + tk) * C + c];
    if (tk > 0) e += w0 * src[(row0 + tk - 1) * C + c];
    if (tk + 1 < N) e += w2 * src[(row0 + tk + 1) * C + c];
    if (BWD) e *= 1.f / (float)(min(c + 1, C - 1) - max(c - 1, 0) + 1);
    return e;
  };
  if (cg == 0) vl = 0.f; else if (lane == 0) vl = edge(c4 - 1);
  if (cg == cgs - 1) vr = 0.f; else if (lane == 31) vr = edge(c4 + 4);
  return make_float4(vl + v.x + v.y, v.x + v.y + v.z, v.y + v.z + v.w, v.z + v.w + vr);
}

__global__ void __launch_bounds__(256) metapool_fwd_kernel(const float* __restrict__ t, const float* __restrict__ cur,
                                                           const float* __restrict__ scale, float* __restrict__ out, int B, int N, int C) {
  const int cgs = C >> 2, lane = threadIdx.x & 31;
  const long long n4 = (long long)B * N * cgs;
  const long long n4_up = (n4 + 31) & ~31ll;               // whole warps stay in the loop (shuffles)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_up; i += (long long)gridDim.x * blockDim.x) {
    const long long j = i < n4 ? i : n4 - 1;
    const unsigned int tok = (unsigned int)j / (unsigned int)cgs;
    const int cg = (int)((unsigned int)j - tok * (unsigned int)cgs);
    const int b = (int)(tok / (unsigned int)N), tk = (int)(tok - (unsigned int)b * (unsigned int)N);
    float4 c;
    const float4 s = metapool_window<false>(cur, (long long)b * N, tk, N, C, cg, lane, c);
    if (i >= n4) continue;
    const float rn = (float)(min(tk + 1, N - 1) - max(tk - 1, 0) + 1);
    const int c4 = cg * 4;
    const float i0 = 1.f / (rn * (float)(min(c4 + 1, C - 1) - max(c4 - 1, 0) + 1));
    const float i1 = 1.f / (rn * 3.f);
    const float i3 = 1.f / (rn * (float)(min(c4 + 4, C - 1) - (c4 + 2) + 1));
    const float sc = scale ? scale[b] : 1.f;
    const float4 tv = reinterpret_cast<const float4*>(t)[j];
    reinterpret_cast<float4*>(out)[j] = make_float4(tv.x + sc * (s.x * i0 - c.x), tv.y + sc * (s.y * i1 - c.y),
                                                    tv.z + sc * (s.z * i1 - c.z), tv.w + sc * (s.w * i3 - c.w));
  }
}
// dcur[n,c] = s[b] * ( sum_{(n',c') in window(n,c)} dy[n',c'] / cnt(n',c')  -  dy[n,c] )
__global__ void __launch_bounds__(256) metapool_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ scale,
                                                           float* __restrict__ dcur, int B, int N, int C) {
  const int cgs = C >> 2, lane = threadIdx.x & 31;
  const long long n4 = (long long)B * N * cgs;
  const long long n4_up = (n4 + 31) & ~31ll;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4_up; i += (long long)gridDim.x * blockDim.x) {
    const long long j = i < n4 ? i : n4 - 1;
    const unsigned int tok = (unsigned int)j / (unsigned int)cgs;
    const int cg = (int)((unsigned int)j - tok * (unsigned int)cgs);
    const int b = (int)(tok / (unsigned int)N), tk = (int)(tok - (unsigned int)b * (unsigned int)N);
    float4 c;
    const float4 s = metapool_window<true>(dy, (long long)b * N, tk, N, C, cg, lane, c);
    if (i >= n4) continue;
    const float sc = scale ? scale[b] : 1.f;
    reinterpret_cast<float4*>(dcur)[j] = make_float4(sc * (s.x - c.x), sc * (s.y - c.y), sc * (s.z - c.z), sc * (s.w - c.w));
  }
}
extern "C" int tcct_metapool_fwd(const float* t, const float* cur, const float* scale, float* out, int B, int N, int C,
                                 void* stream) {
  TCCT_CHECK_ARG((long long)B * N * (C / 4) < (1ll << 31), "metapool: tensor too large for 32-bit indices");
  TCCT_CHECK_ARG(C % 4 == 0 && C >= 8, "metapool: C must be a multiple of 4, >= 8 (got %d)", C);
  metapool_fwd_kernel<<<grid_for((long long)B * N * (C / 4), 256, 8), 256, 0, (cudaStream_t)stream>>>(t, cur, scale, out, B, N, C);
  TCCT_CHECK_LAUNCH("metapool_fwd");
  return TCCT_OK;
}
extern "C" int tcct_metapool_bwd(const float* dy, const float* scale, float* dcur, int B, int N, int C, void* stream) {
  TCCT_CHECK_ARG((long long)B * N * (C / 4) < (1ll << 31), "metapool: tensor too large for 32-bit indices");
  TCCT_CHECK_ARG(C % 4 == 0 && C >= 8, "metapool: C must be a multiple of 4, >= 8 (got %d)", C);
  metapool_bwd_kernel<<<grid_for((long long)B * N * (C / 4), 256, 8), 256, 0, (cudaStream_t)stream>>>(dy, scale, dcur, B, N, C);
  TCCT_CHECK_LAUNCH("metapool_bwd");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// Bilinear resampling (ATen upsample_bilinear2d index rules).
//   align_corners=True  (nn.Upsample in MPUpBlock, tcct.py:890): src = dst*(in-1)/(out-1)
//   align_corners=False (F.interpolate, tcct.py:941,1042-1044): src = max((dst+.5)*in/out - .5, 0)
// ----------------------------------------------------------------------------------------------
struct Lin1 { int i0, i1; float w1; };
__device__ __forceinline__ Lin1 src_index(int dst, int in, int out, int align) {
  float src;
  if (align) {
    const float sc = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    src = sc * dst;
  } else {
    const float sc = (float)in / (float)out;
    src = sc * (dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
  }
  Lin1 r;
  r.i0 = min((int)src, in - 1);
  r.i1 = min(r.i0 + 1, in - 1);
  r.w1 = src - (float)r.i0;
  return r;
}
// the same with the scale factor (in-1)/(out-1) or in/out hoisted out of the call
__device__ __forceinline__ float rs_scale(int in, int out, int align) {
  return align ? (out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f) : (float)in / (float)out;
}
__device__ __forceinline__ Lin1 src_index_s(int dst, int in, float sc, int align) {
  float src = align ? sc * dst : fmaxf(sc * (dst + 0.5f) - 0.5f, 0.f);
  Lin1 r;
  r.i0 = min((int)src, in - 1);
  r.i1 = min(r.i0 + 1, in - 1);
  r.w1 = src - (float)r.i0;
  return r;
}
__device__ __forceinline__ float adj_weight_s(int j, int i, int in, float sc, int align) {
  const Lin1 l = src_index_s(j, in, sc, align);
  float w = 0.f;
  if (l.i0 == i) w += 1.f - l.w1;
  if (l.i1 == i) w += l.w1;
  return w;
}
// weight with which dst position j reads source index i
__device__ __forceinline__ float adj_weight(int j, int i, int in, int out, int align) {
  const Lin1 l = src_index(j, in, out, align);
  float w = 0.f;
  if (l.i0 == i) w += 1.f - l.w1;
  if (l.i1 == i) w += l.w1;
  return w;
}

// out[b,oy,ox,:] = alpha * bilinear(x)[...] (+ add[b,oy,ox,:])      NHWC
__global__ void resize_nhwc_fwd_kernel(const float* __restrict__ x, const float* __restrict__ add, float* __restrict__ out,
                                       int B, int h, int w, int H, int W, int C, int align, float alpha, int accumulate) {
  const int c4 = C >> 2;
  const long long n = (long long)B * H * W * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int cg, ox, oy, b;
    split4(i, c4, W, H, cg, ox, oy, b);
    const Lin1 ly = src_index(oy, h, H, align), lx = src_index(ox, w, W, align);
    const float* base = x + (size_t)b * h * w * C + cg * 4;
    const float4 v00 = *reinterpret_cast<const float4*>(base + ((size_t)ly.i0 * w + lx.i0) * C);
    const float4 v01 = *reinterpret_cast<const float4*>(base + ((size_t)ly.i0 * w + lx.i1) * C);
    const float4 v10 = *reinterpret_cast<const float4*>(base + ((size_t)ly.i1 * w + lx.i0) * C);
    const float4 v11 = *reinterpret_cast<const float4*>(base + ((size_t)ly.i1 * w + lx.i1) * C);
    const float w00 = (1.f - ly.w1) * (1.f - lx.w1), w01 = (1.f - ly.w1) * lx.w1, w10 = ly.w1 * (1.f - lx.w1),
                w11 = ly.w1 * lx.w1;
    float4 o;
    o.x = alpha * (w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x);
    o.y = alpha * (w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y);
    o.z = alpha * (w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z);
    o.w = alpha * (w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w);
    if (add) {
      const float4 a = reinterpret_cast<const float4*>(add)[i];
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    if (accumulate) {
      const float4 a = reinterpret_cast<const float4*>(out)[i];
      o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    reinterpret_cast<float4*>(out)[i] = o;
  }
}

// dx[b,iy,ix,:] = alpha * sum_{oy,ox} wy(oy,iy) wx(ox,ix) dout[b,oy,ox,:]      (gather form of the adjoint)
#define RS_MAXW 10     // output positions that can read one input position per axis: 2*factor + 2, factor <= 4
template <int WIN>     // WIN >= 2*factor + 2
__global__ void resize_nhwc_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int B, int h, int w,
                                       int H, int W, int C, int align, float alpha) {
  const int c4 = C >> 2;
  const int fy = (H + h - 1) / h, fx = (W + w - 1) / w;
  const float scy = rs_scale(h, H, align), scx = rs_scale(w, W, align);
  const long long n = (long long)B * h * w * c4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // 32-bit index arithmetic (B*h*w*c4 < 2^31 is checked by the launcher); the scale factors are hoisted
    const unsigned int iu = (unsigned int)i;
    const int cg = (int)(iu % (unsigned int)c4);
    unsigned int p = iu / (unsigned int)c4;
    const int ix = (int)(p % (unsigned int)w); p /= (unsigned int)w;
    const int iy = (int)(p % (unsigned int)h);
    const int b = (int)(p / (unsigned int)h);
    int oy0 = max(0, fy * (iy - 1) - 1), ox0 = max(0, fx * (ix - 1) - 1);
    const int oy1 = min(H - 1, fy * (iy + 2) + 1), ox1 = min(W - 1, fx * (ix + 2) + 1);
    // the outputs that read this input form one run of at most 2f+2 positions per axis: skip to its start
    while (oy0 < oy1 && adj_weight_s(oy0, iy, h, scy, align) == 0.f) oy0++;
    while (ox0 < ox1 && adj_weight_s(ox0, ix, w, scx, align) == 0.f) ox0++;
    // separable adjoint: the per-axis weights are evaluated once, not once per (oy, ox) pair
    float wy[WIN], wx[WIN];
#pragma unroll
    for (int k = 0; k < WIN; k++) {
      wy[k] = oy0 + k <= oy1 ? adj_weight_s(oy0 + k, iy, h, scy, align) : 0.f;
      wx[k] = ox0 + k <= ox1 ? adj_weight_s(ox0 + k, ix, w, scx, align) : 0.f;
    }
    float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int ky = 0; ky < WIN; ky++) {
      if (wy[ky] == 0.f) continue;
      const float* row = dout + (((size_t)b * H + oy0 + ky) * W + ox0) * C + cg * 4;
#pragma unroll
      for (int kx = 0; kx < WIN; kx++) {
        if (wx[kx] == 0.f) continue;
        const float4 d = *reinterpret_cast<const float4*>(row + (size_t)kx * C);
        const float ww = wy[ky] * wx[kx];
        acc.x += ww * d.x; acc.y += ww * d.y; acc.z += ww * d.z; acc.w += ww * d.w;
      }
    }
    acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha;
    reinterpret_cast<float4*>(dx)[i] = acc;
  }
}

extern "C" int tcct_resize_nhwc_fwd(const float* x, const float* add, float* out, int B, int h, int w, int H, int W,
                                    int C, int align, float alpha, int accumulate, void* stream) {
  TCCT_CHECK_ARG((long long)B * H * W * (C / 4) < (1ll << 31), "resize_nhwc: tensor too large for 32-bit indices");
  TCCT_CHECK_ARG(C % 4 == 0, "resize_nhwc: C must be a multiple of 4");
  const long long n = (long long)B * H * W * (C / 4);
  resize_nhwc_fwd_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, add, out, B, h, w, H, W, C, align,
                                                                               alpha, accumulate);
  TCCT_CHECK_LAUNCH("resize_nhwc_fwd");
  return TCCT_OK;
}
extern "C" int tcct_resize_nhwc_bwd(const float* dout, float* dx, int B, int h, int w, int H, int W, int C, int align,
                                    float alpha, void* stream) {
  TCCT_CHECK_ARG(C % 4 == 0, "resize_nhwc: C must be a multiple of 4");
  TCCT_CHECK_ARG(2 * ((H + h - 1) / h) + 2 <= RS_MAXW && 2 * ((W + w - 1) / w) + 2 <= RS_MAXW, "resize_nhwc_bwd: scale factor above 4");
  const long long n = (long long)B * h * w * (C / 4);
  TCCT_CHECK_ARG(n < (1ll << 31), "resize_nhwc_bwd: too many elements");
  const int win = 2 * (((H + h - 1) / h) > ((W + w - 1) / w) ? ((H + h - 1) / h) : ((W + w - 1) / w)) + 2;
  if (win <= 4) resize_nhwc_bwd_kernel<4><<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(dout, dx, B, h, w, H, W, C, align, alpha);
  else if (win <= 6) resize_nhwc_bwd_kernel<6><<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(dout, dx, B, h, w, H, W, C, align, alpha);
  else resize_nhwc_bwd_kernel<RS_MAXW><<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(dout, dx, B, h, w, H, W, C, align, alpha);
  TCCT_CHECK_LAUNCH("resize_nhwc_bwd");
  return TCCT_OK;
}

// NCHW planes (logit heads, tcct.py:1042-1044), align_corners=False
__global__ void resize_nchw_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int planes, int h, int w,
                                       int H, int W) {
  const long long n = (long long)planes * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int ox, oy, pl;
    split3(i, W, H, ox, oy, pl);
    const Lin1 ly = src_index(oy, h, H, 0), lx = src_index(ox, w, W, 0);
    const float* base = x + (size_t)pl * h * w;
    const float v00 = __ldg(base + ly.i0 * w + lx.i0), v01 = __ldg(base + ly.i0 * w + lx.i1);
    const float v10 = __ldg(base + ly.i1 * w + lx.i0), v11 = __ldg(base + ly.i1 * w + lx.i1);
    out[i] = (1.f - ly.w1) * ((1.f - lx.w1) * v00 + lx.w1 * v01) + ly.w1 * ((1.f - lx.w1) * v10 + lx.w1 * v11);
  }
}
// gather form of the adjoint; the per-axis weights of the (at most 2*factor + 2) outputs that read an input position are
// evaluated once per thread, the double loop is loads and FMAs only
template <int WIN>
__global__ void resize_nchw_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int planes, int h, int w,
                                       int H, int W) {
  const int fy = (H + h - 1) / h, fx = (W + w - 1) / w;
  const long long n = (long long)planes * h * w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int ix, iy, pl;
    split3(i, w, h, ix, iy, pl);
    int oy0 = max(0, fy * (iy - 1) - 1), ox0 = max(0, fx * (ix - 1) - 1);
    const int oy1 = min(H - 1, fy * (iy + 2) + 1), ox1 = min(W - 1, fx * (ix + 2) + 1);
    while (oy0 < oy1 && adj_weight(oy0, iy, h, H, 0) == 0.f) oy0++;
    while (ox0 < ox1 && adj_weight(ox0, ix, w, W, 0) == 0.f) ox0++;
    float wy[WIN], wx[WIN];
#pragma unroll
    for (int k = 0; k < WIN; k++) {
      wy[k] = oy0 + k <= oy1 ? adj_weight(oy0 + k, iy, h, H, 0) : 0.f;
      wx[k] = ox0 + k <= ox1 ? adj_weight(ox0 + k, ix, w, W, 0) : 0.f;
    }
    const float* base = dout + (size_t)pl * H * W + (size_t)oy0 * W + ox0;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < WIN; ky++) {
      if (wy[ky] == 0.f) continue;
      float row = 0.f;
#pragma unroll
      for (int kx = 0; kx < WIN; kx++)
        if (wx[kx] != 0.f) row += wx[kx] * __ldg(base + (size_t)ky * W + kx);
      acc += wy[ky] * row;
    }
    dx[i] = acc;
  }
}
extern "C" int tcct_resize_nchw_fwd(const float* x, float* out, int planes, int h, int w, int H, int W, void* stream) {
  TCCT_CHECK_ARG((long long)planes * H * W < (1ll << 31), "resize_nchw: tensor too large for 32-bit indices");
  resize_nchw_fwd_kernel<<<grid_for((long long)planes * H * W, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, out, planes, h, w, H, W);
  TCCT_CHECK_LAUNCH("resize_nchw_fwd");
  return TCCT_OK;
}
extern "C" int tcct_resize_nchw_bwd(const float* dout, float* dx, int planes, int h, int w, int H, int W, void* stream) {
  TCCT_CHECK_ARG((long long)planes * H * W < (1ll << 31), "resize_nchw: tensor too large for 32-bit indices");
  const int fy = (H + h - 1) / h, fx = (W + w - 1) / w;
  const int win = 2 * (fy > fx ? fy : fx) + 2;
  TCCT_CHECK_ARG(win <= 18, "resize_nchw_bwd: scale factor above 8");
  const int grid = grid_for((long long)planes * h * w, 128, 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (win <= 6) resize_nchw_bwd_kernel<6><<<grid, 128, 0, st>>>(dout, dx, planes, h, w, H, W);
  else if (win <= 10) resize_nchw_bwd_kernel<10><<<grid, 128, 0, st>>>(dout, dx, planes, h, w, H, W);
  else resize_nchw_bwd_kernel<18><<<grid, 128, 0, st>>>(dout, dx, planes, h, w, H, W);
  TCCT_CHECK_LAUNCH("resize_nchw_bwd");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// F.normalize(x, dim=channel, p=2, eps=1e-12) on NHWC with C = 32 (norm_add, tcct.py:937-942):
// 8 lanes per pixel (float4 each).
// ----------------------------------------------------------------------------------------------
__global__ void l2norm32_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long npix, float alpha) {
  const long long n = npix * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 31) & ~31ll); i += (long long)gridDim.x * blockDim.x) {
    float4 v = make_float4(0, 0, 0, 0);
    if (i < n) v = reinterpret_cast<const float4*>(x)[i];
    float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    const float inv = alpha / fmaxf(sqrtf(s), 1e-12f);
    if (i < n) reinterpret_cast<float4*>(y)[i] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
}
// dx = alpha * (dy - n * <n, dy>) / max(|x|, eps)
__global__ void l2norm32_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                    long long npix, float alpha) {
  const long long n = npix * 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ((n + 31) & ~31ll); i += (long long)gridDim.x * blockDim.x) {
    float4 v = make_float4(0, 0, 0, 0), d = make_float4(0, 0, 0, 0);
    if (i < n) { v = reinterpret_cast<const float4*>(x)[i]; d = reinterpret_cast<const float4*>(dy)[i]; }
    d.x *= alpha; d.y *= alpha; d.z *= alpha; d.w *= alpha;
    float s = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    float dt = v.x * d.x + v.y * d.y + v.z * d.z + v.w * d.w;
    s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
    dt += __shfl_xor_sync(0xffffffffu, dt, 1); dt += __shfl_xor_sync(0xffffffffu, dt, 2); dt += __shfl_xor_sync(0xffffffffu, dt, 4);
    const float nrm = sqrtf(s);
    float4 o;
    if (nrm > 1e-12f) {
      const float inv = 1.f / nrm, k = dt * inv * inv;     // <n,dy>/|x| = <x,dy>/|x|^2
      o = make_float4((d.x - v.x * k) * inv, (d.y - v.y * k) * inv, (d.z - v.z * k) * inv, (d.w - v.w * k) * inv);
    } else {
      o = make_float4(d.x * 1e12f, d.y * 1e12f, d.z * 1e12f, d.w * 1e12f);
    }
    if (i < n) reinterpret_cast<float4*>(dx)[i] = o;
  }
}
extern "C" int tcct_l2norm32_fwd(const float* x, float* y, long long npix, float alpha, void* stream) {
  l2norm32_fwd_kernel<<<grid_for(npix * 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, y, npix, alpha);
  TCCT_CHECK_LAUNCH("l2norm32_fwd");
  return TCCT_OK;
}
extern "C" int tcct_l2norm32_bwd(const float* x, const float* dy, float* dx, long long npix, float alpha, void* stream) {
  l2norm32_bwd_kernel<<<grid_for(npix * 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, npix, alpha);
  TCCT_CHECK_LAUNCH("l2norm32_bwd");
  return TCCT_OK;
}


// norm_add (tcct.py:937-942) in one pass: out = alpha * ( x0/|x0| + up(n1) + up(n2) ), n1 / n2 already L2-normalised maps of
// lower resolution, bilinear align_corners=False; 8 lanes per output pixel (32 channels).  n2 may be null.
__global__ void norm_add3_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ n1, const float* __restrict__ n2,
                                     float* __restrict__ out, int B, int H, int W, int h1, int w1, int h2, int w2, float alpha) {
  const long long n = (long long)B * H * W * 8, nr = (n + 31) & ~31ll;
  // the rest of one element once its full-resolution operand v and the sum of squares over its pixel's 8 lanes are there
  auto finish = [&](long long i, const float4& v, float s) {
    const float inv = alpha / fmaxf(sqrtf(s), 1e-12f);
    float4 o = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    const int cg = (int)(i & 7);
    int ox, oy, b;
    split3(i >> 3, W, H, ox, oy, b);
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const float* src = k == 0 ? n1 : n2;
      if (!src) continue;
      const int h = k == 0 ? h1 : h2, w = k == 0 ? w1 : w2;
      const Lin1 ly = src_index(oy, h, H, 0), lx = src_index(ox, w, W, 0);
      const float* base = src + (size_t)b * h * w * 32 + cg * 4;
      const float4 v00 = *reinterpret_cast<const float4*>(base + ((size_t)ly.i0 * w + lx.i0) * 32);
      const float4 v01 = *reinterpret_cast<const float4*>(base + ((size_t)ly.i0 * w + lx.i1) * 32);
      const float4 v10 = *reinterpret_cast<const float4*>(base + ((size_t)ly.i1 * w + lx.i0) * 32);
      const float4 v11 = *reinterpret_cast<const float4*>(base + ((size_t)ly.i1 * w + lx.i1) * 32);
      const float w00 = (1.f - ly.w1) * (1.f - lx.w1), w01 = (1.f - ly.w1) * lx.w1, w10 = ly.w1 * (1.f - lx.w1), w11 = ly.w1 * lx.w1;
      o.x += alpha * (w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x);
      o.y += alpha * (w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y);
      o.z += alpha * (w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z);
      o.w += alpha * (w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w);
    }
    reinterpret_cast<float4*>(out)[i] = o;
  };
  // two elements per iteration, both full-resolution loads issued first (a warp's 32 consecutive elements are all below nr or all not)
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += 2 * stride) {
    const long long i1 = i + stride;
    const bool ok0 = i < n, ok1 = i1 < n, second = i1 < nr;
    float4 v0 = make_float4(0, 0, 0, 0), v1 = v0;
    if (ok0) v0 = __ldcs(reinterpret_cast<const float4*>(x0) + i);
    if (ok1) v1 = __ldcs(reinterpret_cast<const float4*>(x0) + i1);
    float s0 = v0.x * v0.x + v0.y * v0.y + v0.z * v0.z + v0.w * v0.w;
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s0 += __shfl_xor_sync(0xffffffffu, s0, 4);
    float s1 = 0.f;
    if (second) {
      s1 = v1.x * v1.x + v1.y * v1.y + v1.z * v1.z + v1.w * v1.w;
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 4);
    }
    if (ok0) finish(i, v0, s0);
    if (ok1) finish(i1, v1, s1);
  }
}
extern "C" int tcct_norm_add3_fwd(const float* x0, const float* n1, const float* n2, float* out, int B, int H, int W, int h1,
                                  int w1, int h2, int w2, float alpha, void* stream) {
  TCCT_CHECK_ARG((long long)B * H * W * 8 < (1ll << 31), "norm_add3: tensor too large for 32-bit indices");
  TCCT_CHECK_ARG(n1 != nullptr, "norm_add3: n1 is required");
  norm_add3_fwd_kernel<<<grid_for((long long)B * H * W * 8, 256, 8), 256, 0, (cudaStream_t)stream>>>(x0, n1, n2, out, B, H, W, h1, w1,
                                                                                                     h2, w2, alpha);
  TCCT_CHECK_LAUNCH("norm_add3_fwd");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// Stem convs: 3x3, 3 -> 32 channels, pad 1, stride 1|2, NCHW fp32 image in, NHWC out
// (CrossResNet.cnn tcct.py:873, MPViT.stem[0] 673-681).  8 threads per output pixel (4 channels each).
// ----------------------------------------------------------------------------------------------
// A block owns a 32-wide, ST_ROWS-tall tile of output pixels: the 3-channel image patch (with halo) is staged in shared
// memory.  Thread (pixel pair j, channel group cg, row half) keeps the 27 x 4 weights of its 4 output channels in registers and
// walks down 8 rows of TWO neighbouring output pixels: their 3x3 windows overlap, so one 64/128-bit shared-memory load per
// (input channel, kernel row) feeds 24 FMAs (the one-pixel form issued 27 32-bit loads per 108 FMAs and ran at ~1/4 of the FMA
// issue rate: 67 us for the full-resolution stem against a ~15 us FMA floor).
#define ST_ROWS 16
// CTAs of a persistent kernel: as many as are co-resident (occupancy x SMs), at most one per work item
template <class K>
static int persistent_grid(K kernel, int threads, size_t smem, int items) {
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int cap = per_sm * tcct_num_sms();
  return items < cap ? items : cap;
}
template <int STRIDE> struct StemGeom {
  static constexpr int PW = 31 * STRIDE + 3, PH = (ST_ROWS - 1) * STRIDE + 3;
  static constexpr int PP = (PW + 3) / 4 * 4;          // row pitch: 36 (stride 1), 68 (stride 2) -> aligned vector loads
  static constexpr int NV = STRIDE + 3;                // input columns under two neighbouring output pixels: 4 or 5
};
// the NV input values of one (channel, row) under the pixel pair j
template <int STRIDE>
__device__ __forceinline__ void stem_window(const float* row, int j, float (&v)[5]) {
  if (STRIDE == 1) {
    const float2 a = *reinterpret_cast<const float2*>(row + 2 * j), b = *reinterpret_cast<const float2*>(row + 2 * j + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = 0.f;
  } else {
    const float4 a = *reinterpret_cast<const float4*>(row + 4 * j);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = row[4 * j + 4];
  }
}
// The image patch of a tile, fetched with 4-byte cp.async (zero-filled outside the image): the persistent kernels below stage tile t + 1 into the
// other buffer while tile t is computed -- with one 256-thread CTA per SM (163-226 registers) nothing else hides that round trip.
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(saddr), "l"(g), "r"(src_bytes));
}
template <int STRIDE>
__device__ __forceinline__ void stem_stage_async(const float* __restrict__ img, int t, int tiles_x, int tiles_y, int H, int W,
                                                 float* simg) {
  typedef StemGeom<STRIDE> G;
  const int b = t / (tiles_x * tiles_y), r0 = t - b * tiles_x * tiles_y;
  const int ix0 = (r0 % tiles_x) * 32 * STRIDE - 1, iy0 = (r0 / tiles_x) * ST_ROWS * STRIDE - 1;
  const uint32_t sbase = smem_u32(simg);
  for (int i = threadIdx.x; i < 3 * G::PH * G::PP; i += 256) {
    const int px = i % G::PP, py = (i / G::PP) % G::PH, ci = i / (G::PP * G::PH);
    const int iy = iy0 + py, ix = ix0 + px;
    const bool ok = px < G::PW && iy >= 0 && iy < H && ix >= 0 && ix < W;
    cp_async4(sbase + 4u * i, ok ? img + (((size_t)b * 3 + ci) * H + iy) * W + ix : img, ok ? 4 : 0);
  }
  cp_async_commit();
}

// Persistent: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... and keeps its BatchNorm partial sums in registers (one set
// of double atomics per CTA).
template <int STRIDE>
__global__ void __launch_bounds__(256) stem_conv_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W,
                                                         int Ho, int Wo, double* stats) {
  typedef StemGeom<STRIDE> G;
  extern __shared__ __align__(16) float sbuf[];     // 2 x [3][PH][PP]
  __shared__ float sred[64];
  constexpr int PATCH = 3 * G::PH * G::PP;
  const int cg = threadIdx.x & 7, j = (threadIdx.x >> 3) & 15, half = threadIdx.x >> 7;
  const int tiles_x = (Wo + 31) / 32, tiles_y = (Ho + ST_ROWS - 1) / ST_ROWS;
  const int ntiles = tiles_x * tiles_y * B;
  if (threadIdx.x < 64) sred[threadIdx.x] = 0.f;
  float wr[27][4];
#pragma unroll
  for (int k = 0; k < 27; k++)
#pragma unroll
    for (int i = 0; i < 4; i++) wr[k][i] = w[(cg * 4 + i) * 27 + k];
  float bs[4];
#pragma unroll
  for (int i = 0; i < 4; i++) bs[i] = bias ? bias[cg * 4 + i] : 0.f;
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (blockIdx.x < ntiles) stem_stage_async<STRIDE>(img, blockIdx.x, tiles_x, tiles_y, H, W, sbuf);
  int cur = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, cur ^= 1) {
    const int b = t / (tiles_x * tiles_y), r0 = t - b * tiles_x * tiles_y;
    const int ox0 = (r0 % tiles_x) * 32, oy0 = (r0 / tiles_x) * ST_ROWS;
    const float* simg = sbuf + cur * PATCH;
    __syncthreads();                                  // every thread is done with the buffer the prefetch overwrites
    if (t + gridDim.x < ntiles) {
      stem_stage_async<STRIDE>(img, t + gridDim.x, tiles_x, tiles_y, H, W, sbuf + (cur ^ 1) * PATCH);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int ox = ox0 + 2 * j;
    const int rows = min(ST_ROWS, Ho - oy0);
    for (int r = half * (ST_ROWS / 2); r < min(rows, (half + 1) * (ST_ROWS / 2)); r++) {
      float o[2][4];
#pragma unroll
      for (int i = 0; i < 4; i++) o[0][i] = o[1][i] = bs[i];
#pragma unroll
      for (int ci = 0; ci < 3; ci++)
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
          float v[5];
          stem_window<STRIDE>(simg + (ci * G::PH + r * STRIDE + ky) * G::PP, j, v);
#pragma unroll
          for (int kx = 0; kx < 3; kx++) {
            const int k = ci * 9 + ky * 3 + kx;
#pragma unroll
            for (int i = 0; i < 4; i++) { o[0][i] += v[kx] * wr[k][i]; o[1][i] += v[kx + STRIDE] * wr[k][i]; }
          }
        }
      float* dst = y + (((size_t)b * Ho + oy0 + r) * Wo + ox) * 32 + cg * 4;
#pragma unroll
      for (int p = 0; p < 2; p++)
        if (ox + p < Wo) {
          __stcs(reinterpret_cast<float4*>(dst + p * 32), make_float4(o[p][0], o[p][1], o[p][2], o[p][3]));
#pragma unroll
          for (int i = 0; i < 4; i++) { s[i] += o[p][i]; q[i] += o[p][i] * o[p][i]; }
        }
    }
  }
  if (stats) {
#pragma unroll
    for (int i = 0; i < 4; i++) {      // lanes l, l+8, l+16, l+24 share a channel group
      s[i] += __shfl_xor_sync(0xffffffffu, s[i], 8); s[i] += __shfl_xor_sync(0xffffffffu, s[i], 16);
      q[i] += __shfl_xor_sync(0xffffffffu, q[i], 8); q[i] += __shfl_xor_sync(0xffffffffu, q[i], 16);
    }
    __syncthreads();
    if ((threadIdx.x & 31) < 8) {
#pragma unroll
      for (int i = 0; i < 4; i++) { atomicAdd(&sred[cg * 4 + i], s[i]); atomicAdd(&sred[32 + cg * 4 + i], q[i]); }
    }
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(stats + threadIdx.x, (double)sred[threadIdx.x]);
  }
}

// dw[co][ci*9+tap] += sum_p dy[p][co] * img[...];  dbias[co] += sum_p dy[p][co]    (same tiling and pixel pairing as the forward).
// Persistent: a CTA walks its tiles with the 28 x 4 partial sums in registers and touches global memory with atomics once at the end.
template <int STRIDE>
__global__ void __launch_bounds__(256) stem_conv_wgrad_kernel(const float* __restrict__ img, const float* __restrict__ dy, float* dw,
                                                              float* dbias, int B, int H, int W, int Ho, int Wo) {
  typedef StemGeom<STRIDE> G;
  extern __shared__ __align__(16) float sbuf[];     // 2 x [3][PH][PP] | per-warp partial sums [8][28*32] (no shared-memory float atomics)
  constexpr int PATCH = 3 * G::PH * G::PP;
  float* sred = sbuf + 2 * PATCH;
  const int cg = threadIdx.x & 7, j = (threadIdx.x >> 3) & 15, half = threadIdx.x >> 7;
  const int tiles_x = (Wo + 31) / 32, tiles_y = (Ho + ST_ROWS - 1) / ST_ROWS;
  const int ntiles = tiles_x * tiles_y * B;
  float acc[28][4];
#pragma unroll
  for (int k = 0; k < 28; k++)
#pragma unroll
    for (int i = 0; i < 4; i++) acc[k][i] = 0.f;
  if (blockIdx.x < ntiles) stem_stage_async<STRIDE>(img, blockIdx.x, tiles_x, tiles_y, H, W, sbuf);
  int cur = 0;
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x, cur ^= 1) {
    const int b = t / (tiles_x * tiles_y), r0 = t - b * tiles_x * tiles_y;
    const int ox0 = (r0 % tiles_x) * 32, oy0 = (r0 / tiles_x) * ST_ROWS;
    const float* simg = sbuf + cur * PATCH;
    const int ox = ox0 + 2 * j;
    const int rows = min(ST_ROWS, Ho - oy0);
    const bool ok0 = ox < Wo, ok1 = ox + 1 < Wo;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int r_begin = half * (ST_ROWS / 2), r_end = min(rows, (half + 1) * (ST_ROWS / 2));
    // the gradient rows are read one iteration ahead (the first one before waiting for the patch)
    const float* dyp = dy + (((size_t)b * Ho + oy0 + r_begin) * Wo + ox) * 32 + cg * 4;
    float4 a4 = z4, b4 = z4;
    if (r_begin < r_end) {
      a4 = ok0 ? __ldcs(reinterpret_cast<const float4*>(dyp)) : z4;
      b4 = ok1 ? __ldcs(reinterpret_cast<const float4*>(dyp + 32)) : z4;
    }
    __syncthreads();                                  // every thread is done with the buffer the prefetch overwrites
    if (t + gridDim.x < ntiles) {
      stem_stage_async<STRIDE>(img, t + gridDim.x, tiles_x, tiles_y, H, W, sbuf + (cur ^ 1) * PATCH);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    for (int r = r_begin; r < r_end; r++) {
      const float d0[4] = {a4.x, a4.y, a4.z, a4.w}, d1[4] = {b4.x, b4.y, b4.z, b4.w};
      if (r + 1 < r_end) {
        dyp += (size_t)Wo * 32;
        a4 = ok0 ? __ldcs(reinterpret_cast<const float4*>(dyp)) : z4;
        b4 = ok1 ? __ldcs(reinterpret_cast<const float4*>(dyp + 32)) : z4;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) acc[27][i] += d0[i] + d1[i];
#pragma unroll
      for (int ci = 0; ci < 3; ci++)
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
          float v[5];
          stem_window<STRIDE>(simg + (ci * G::PH + r * STRIDE + ky) * G::PP, j, v);
#pragma unroll
          for (int kx = 0; kx < 3; kx++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[ci * 9 + ky * 3 + kx][i] += d0[i] * v[kx] + d1[i] * v[kx + STRIDE];
        }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 28; k++)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float v = acc[k][i];
      v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 31) < 8) sred[(threadIdx.x >> 5) * 28 * 32 + k * 32 + cg * 4 + i] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 28 * 32; i += 256) {
    const int k = i >> 5, co = i & 31;
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) v += sred[wv * 28 * 32 + i];
    if (k < 27) atomicAdd(dw + co * 27 + k, v);
    else if (dbias) atomicAdd(dbias + co, v);
  }
}

extern "C" int tcct_stem_conv_fwd(const float* img, const float* w, const float* bias, float* y, int B, int H, int W,
                                  int stride, double* stats, void* stream) {
  TCCT_CHECK_ARG(stride == 1 || stride == 2, "stem_conv: stride 1|2 expected");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int ntiles = ceil_div(Wo, 32) * ceil_div(Ho, ST_ROWS) * B;
  if (stride == 1) {
    const size_t smem = (size_t)2 * 3 * StemGeom<1>::PH * StemGeom<1>::PP * sizeof(float);
    const int ctas = persistent_grid(stem_conv_fwd_kernel<1>, 256, smem, ntiles);
    stem_conv_fwd_kernel<1><<<ctas, 256, smem, (cudaStream_t)stream>>>(img, w, bias, y, B, H, W, Ho, Wo, stats);
  } else {
    const size_t smem = (size_t)2 * 3 * StemGeom<2>::PH * StemGeom<2>::PP * sizeof(float);
    cudaFuncSetAttribute(stem_conv_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ctas = persistent_grid(stem_conv_fwd_kernel<2>, 256, smem, ntiles);
    stem_conv_fwd_kernel<2><<<ctas, 256, smem, (cudaStream_t)stream>>>(img, w, bias, y, B, H, W, Ho, Wo, stats);
  }
  TCCT_CHECK_LAUNCH("stem_conv_fwd");
  return TCCT_OK;
}
extern "C" int tcct_stem_conv_wgrad(const float* img, const float* dy, float* dw, float* dbias, int B, int H, int W,
                                    int stride, void* stream) {
  TCCT_CHECK_ARG(stride == 1 || stride == 2, "stem_conv: stride 1|2 expected");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int ntiles = ceil_div(Wo, 32) * ceil_div(Ho, ST_ROWS) * B;
  if (stride == 1) {
    const size_t smem = ((size_t)2 * 3 * StemGeom<1>::PH * StemGeom<1>::PP + 8 * 28 * 32) * sizeof(float);
    cudaFuncSetAttribute(stem_conv_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ctas = persistent_grid(stem_conv_wgrad_kernel<1>, 256, smem, ntiles);
    stem_conv_wgrad_kernel<1><<<ctas, 256, smem, (cudaStream_t)stream>>>(img, dy, dw, dbias, B, H, W, Ho, Wo);
  } else {
    const size_t smem = ((size_t)2 * 3 * StemGeom<2>::PH * StemGeom<2>::PP + 8 * 28 * 32) * sizeof(float);
    cudaFuncSetAttribute(stem_conv_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ctas = persistent_grid(stem_conv_wgrad_kernel<2>, 256, smem, ntiles);
    stem_conv_wgrad_kernel<2><<<ctas, 256, smem, (cudaStream_t)stream>>>(img, dy, dw, dbias, B, H, W, Ho, Wo);
  }
  TCCT_CHECK_LAUNCH("stem_conv_wgrad");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// Logit heads: 1x1 conv 32 -> Cc (aux0/1/2/4, tcct.py:994-997,1041), NHWC in, NCHW logits out.
// ----------------------------------------------------------------------------------------------
#define HEAD_MAXC 16
__global__ void head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                float* __restrict__ out, int B, int HW, int Cc) {
  __shared__ float sw[HEAD_MAXC * 32 + HEAD_MAXC];
  for (int i = threadIdx.x; i < Cc * 32; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cc; i += blockDim.x) sw[HEAD_MAXC * 32 + i] = bias[i];
  __syncthreads();
  const long long npix = (long long)B * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    float v[32];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const float4 t = reinterpret_cast<const float4*>(x + p * 32)[k];
      v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
    const int b = (int)((unsigned int)p / (unsigned int)HW);      // 32-bit: the launcher checks B * HW < 2^31
    const int q = (int)((unsigned int)p - (unsigned int)b * (unsigned int)HW);
    for (int c = 0; c < Cc; c++) {
      float s = sw[HEAD_MAXC * 32 + c];
#pragma unroll
      for (int k = 0; k < 32; k++) s += v[k] * sw[c * 32 + k];
      out[((size_t)b * Cc + c) * HW + q] = s;
    }
  }
}

// dx[p][k] = sum_c dl[b,c,q] w[c][k];  dw[c][k] += sum_p dl*x;  db[c] += sum_p dl.
// Eight lanes per pixel (one float4 of the 128-byte feature row each: coalesced x reads and dx writes); persistent CTAs keep
// the Cc x 4 partial weight gradients of their lane in registers and touch global memory once, at the end.
template <int CC>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ dl, float* __restrict__ dx,
                                                       float* dw, float* db, int B, int HW, int Cc) {
  __shared__ float sw[HEAD_MAXC * 32];
  __shared__ float sacc[HEAD_MAXC * 33];
  const int tid = threadIdx.x, sub = tid & 7;
  for (int i = tid; i < Cc * 32; i += 256) sw[i] = w[i];
  for (int i = tid; i < HEAD_MAXC * 33; i += 256) sacc[i] = 0.f;
  __syncthreads();
  const long long npix = (long long)B * HW;
  float acc[CC][4], dsum[CC];
#pragma unroll
  for (int c = 0; c < CC; c++) { acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f; dsum[c] = 0.f; }
  for (long long p = (long long)blockIdx.x * 32 + (tid >> 3); p < npix; p += (long long)gridDim.x * 32) {
    const int b = (int)((unsigned int)p / (unsigned int)HW);      // a 64-bit division here costs more than the pixel's arithmetic
    const int q = (int)((unsigned int)p - (unsigned int)b * (unsigned int)HW);
    const float4 xv = __ldcs(reinterpret_cast<const float4*>(x + p * 32) + sub);
    float d[CC];
#pragma unroll
    for (int c = 0; c < CC; c++) d[c] = c < Cc ? __ldg(dl + ((size_t)b * Cc + c) * HW + q) : 0.f;
    float4 o = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int c = 0; c < CC; c++) {
      const float4 wv = *reinterpret_cast<const float4*>(&sw[c * 32 + sub * 4]);
      o.x += d[c] * wv.x; o.y += d[c] * wv.y; o.z += d[c] * wv.z; o.w += d[c] * wv.w;
      acc[c][0] += d[c] * xv.x; acc[c][1] += d[c] * xv.y; acc[c][2] += d[c] * xv.z; acc[c][3] += d[c] * xv.w;
      dsum[c] += d[c];
    }
    if (dx) reinterpret_cast<float4*>(dx + p * 32)[sub] = o;
  }
  // lanes l, l+8, l+16, l+24 hold the same channels
#pragma unroll
  for (int c = 0; c < CC; c++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float v = acc[c][i];
      v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((tid & 31) < 8 && c < Cc) atomicAdd(&sacc[c * 33 + sub * 4 + i], v);
    }
    float ds = dsum[c];
    ds += __shfl_xor_sync(0xffffffffu, ds, 8); ds += __shfl_xor_sync(0xffffffffu, ds, 16);
    if ((tid & 31) == 0 && c < Cc) atomicAdd(&sacc[c * 33 + 32], ds);
  }
  __syncthreads();
  for (int e = tid; e < Cc * 33; e += 256) {
    const int c = e / 33, k = e - c * 33;
    if (k < 32) atomicAdd(dw + c * 32 + k, sacc[e]);
    else atomicAdd(db + c, sacc[e]);
  }
}

extern "C" int tcct_head_fwd(const float* x, const float* w, const float* bias, float* out, int B, int HW, int Cc,
                             void* stream) {
  TCCT_CHECK_ARG(Cc >= 1 && Cc <= HEAD_MAXC, "head: 1 <= classes <= %d expected (got %d)", HEAD_MAXC, Cc);
  TCCT_CHECK_ARG((long long)B * HW < (1ll << 31), "head: too many pixels for 32-bit indices");
  head_fwd_kernel<<<grid_for((long long)B * HW, 128, 8), 128, 0, (cudaStream_t)stream>>>(x, w, bias, out, B, HW, Cc);
  TCCT_CHECK_LAUNCH("head_fwd");
  return TCCT_OK;
}
extern "C" int tcct_head_bwd(const float* x, const float* w, const float* dl, float* dx, float* dw, float* db, int B,
                             int HW, int Cc, void* stream) {
  TCCT_CHECK_ARG(Cc >= 1 && Cc <= HEAD_MAXC, "head: 1 <= classes <= %d expected (got %d)", HEAD_MAXC, Cc);
  TCCT_CHECK_ARG((long long)B * HW < (1ll << 31), "head: too many pixels for 32-bit indices");
  const int grid = grid_for((long long)B * HW, 32, 4);
  if (Cc <= 5) head_bwd_kernel<5><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, dl, dx, dw, db, B, HW, Cc);
  else if (Cc <= 9) head_bwd_kernel<9><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, dl, dx, dw, db, B, HW, Cc);
  else head_bwd_kernel<HEAD_MAXC><<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, dl, dx, dw, db, B, HW, Cc);
  TCCT_CHECK_LAUNCH("head_bwd");
  return TCCT_OK;
}

// y[p][c] = x[p][c] * scale[p / px_per_sample]   (DropPath branch gradient, timm DropPath semantics)
__global__ void scale_per_sample_kernel(const float* __restrict__ x, const float* __restrict__ scale, float* __restrict__ y,
                                        long long n4, int row4_per_sample) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float s = scale[(unsigned int)i / (unsigned int)row4_per_sample];      // n4 < 2^31 (launcher)
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<float4*>(y)[i] = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
  }
}
// n = total elements, per_sample = elements per batch sample (both multiples of 4)
extern "C" int tcct_scale_per_sample(const float* x, const float* scale, float* y, long long n, int per_sample, void* stream) {
  TCCT_CHECK_ARG(n % 4 == 0 && per_sample % 4 == 0, "scale_per_sample: sizes must be multiples of 4");
  TCCT_CHECK_ARG(n / 4 < (1ll << 31), "scale_per_sample: tensor too large for 32-bit indices");
  scale_per_sample_kernel<<<grid_for(n / 4, 256, 8), 256, 0, (cudaStream_t)stream>>>(x, scale, y, n / 4, per_sample / 4);
  TCCT_CHECK_LAUNCH("scale_per_sample");
  return TCCT_OK;
}
