// MHCABlock token mixer in one pass (task1/nets/tcct.py:457-469 with MetaPool 405-415, LayerNorm eps 1e-6 at 427,454-455):
//     cur  = LayerNorm1(t)
//     t2   = t + s[b] * ( AvgPool3x3_{(token, channel) plane, count_include_pad=False}(cur) - cur )      (DropPath scale s[b])
//     cur2 = LayerNorm2(t2)                                  -> the input of Mlp.fc1; t2 is the residual of Mlp.fc2
// SURVEY 8(b) `ln_metapool_{fwd,bwd}`.  The unfused path ran LayerNorm, MetaPool and LayerNorm as three kernels (7 tensor passes
// forward, 14 backward with autograd's two gradient adds); here t is read once and t2, cur2 are written once (3 passes), and the
// backward reads t, t2, d(t2), d(cur2) and writes d(t) (5 passes).
//
// A warp walks a run of consecutive tokens of one sample; lanes own channels lane + 32 i.  The 3x3 window spans the neighbouring
// TOKENS (rows of the [N, C] plane) and CHANNELS: channel neighbours come from the adjacent lanes by shuffle, token neighbours
// from a rolling window of the per-token horizontal sums held in registers (the two tokens bordering a run are recomputed).
#include "common.cuh"

#define LM_MAXI 8          // C <= 256
#define LM_WARPS 8

// horizontal 3-sum over channels of a per-lane channel vector (v[i] = channel lane + 32 i; entries beyond C hold 0)
template <int NI>
__device__ __forceinline__ void hsum3(const float (&v)[NI], float (&h)[NI], int lane) {
#pragma unroll
  for (int i = 0; i < NI; i++) {
    const float up = __shfl_up_sync(0xffffffffu, v[i], 1), dn = __shfl_down_sync(0xffffffffu, v[i], 1);
    const float wrapl = i > 0 ? __shfl_sync(0xffffffffu, v[i > 0 ? i - 1 : 0], 31) : 0.f;
    const float wrapr = i + 1 < NI ? __shfl_sync(0xffffffffu, v[i + 1 < NI ? i + 1 : i], 0) : 0.f;
    h[i] = (lane == 0 ? wrapl : up) + v[i] + (lane == 31 ? wrapr : dn);
  }
}

struct LmArgs {
  const float* t; const float* g1; const float* b1; const float* g2; const float* b2; const float* scale;
  float* t2; float* cur2; float* stats;      // stats: [B*N][4] = mean1, rstd1, mean2, rstd2
  int B, N, C, run; float eps;
};

// LayerNorm statistics of a channel vector (two-pass, as the reference's ATen kernel and csrc/pointwise.cu do)
template <int NI>
__device__ __forceinline__ void ln_stats(const float (&x)[NI], int C, int lane, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NI; i++) s += x[i];
  mean = warp_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NI; i++) {
    const float d = (lane + 32 * i < C) ? x[i] - mean : 0.f;
    q += d * d;
  }
  rstd = rsqrtf(warp_sum(q) / C + eps);
}

template <int NI>
__global__ void __launch_bounds__(32 * LM_WARPS) ln_metapool_fwd_kernel(const LmArgs a) {
  const int lane = threadIdx.x & 31;
  const int C = a.C, N = a.N;
  const int runs_per_b = (N + a.run - 1) / a.run;
  const long long nruns = (long long)a.B * runs_per_b;
  float g1[NI], b1[NI], g2[NI], b2[NI], irc[NI];
#pragma unroll
  for (int i = 0; i < NI; i++) {
    const int c = lane + 32 * i;
    const bool ok = c < C;
    g1[i] = ok ? a.g1[c] : 0.f; b1[i] = ok ? a.b1[c] : 0.f; g2[i] = ok ? a.g2[c] : 0.f; b2[i] = ok ? a.b2[c] : 0.f;
    irc[i] = ok ? 1.f / (float)(min(c + 1, C - 1) - max(c - 1, 0) + 1) : 0.f;      // 1 / valid channels of the window centred on c
  }
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nruns; r += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int b = (int)(r / runs_per_b);
    const int n0 = (int)(r - (long long)b * runs_per_b) * a.run, n1 = min(n0 + a.run, N);
    const float* tb = a.t + (size_t)b * N * C;
    const float sc = a.scale ? a.scale[b] : 1.f;
    // raw values of a token (zeros outside the sample), issued one iteration ahead of their use
    auto load = [&](int n, float (&x)[NI]) {
#pragma unroll
      for (int i = 0; i < NI; i++) x[i] = (n >= 0 && n < N && lane + 32 * i < C) ? __ldg(tb + (size_t)n * C + lane + 32 * i) : 0.f;
    };
    // LayerNorm1 of a loaded token and its horizontal sums
    auto process = [&](int n, const float (&x)[NI], float (&y)[NI], float (&h)[NI], float& mean, float& rstd) {
      if (n < 0 || n >= N) {
#pragma unroll
        for (int i = 0; i < NI; i++) y[i] = h[i] = 0.f;
        mean = 0.f; rstd = 0.f;
        return;
      }
      ln_stats<NI>(x, C, lane, a.eps, mean, rstd);
#pragma unroll
      for (int i = 0; i < NI; i++) y[i] = (lane + 32 * i < C) ? (x[i] - mean) * rstd * g1[i] + b1[i] : 0.f;
      hsum3<NI>(y, h, lane);
    };
    float xm[NI], ym[NI], hm[NI], xc[NI], yc[NI], hc[NI], xn[NI], yn[NI], hn[NI], xp[NI];
    float mm, rm, mc, rc, mn, rn_;
    load(n0 - 1, xm); load(n0, xc); load(n0 + 1, xn);
    process(n0 - 1, xm, ym, hm, mm, rm);
    process(n0, xc, yc, hc, mc, rc);
    for (int n = n0; n < n1; n++) {
      load(n + 2, xp);                                  // in flight while token n + 1 is normalised and token n is finished
      process(n + 1, xn, yn, hn, mn, rn_);
      const float irn = 1.f / (float)(min(n + 1, N - 1) - max(n - 1, 0) + 1);      // 1 / valid tokens of the window
      float o[NI];
#pragma unroll
      for (int i = 0; i < NI; i++) o[i] = (lane + 32 * i < C) ? xc[i] + sc * ((hm[i] + hc[i] + hn[i]) * irn * irc[i] - yc[i]) : 0.f;
      float m2, r2;
      ln_stats<NI>(o, C, lane, a.eps, m2, r2);
      const size_t off = ((size_t)b * N + n) * C;
#pragma unroll
      for (int i = 0; i < NI; i++) {
        const int c = lane + 32 * i;
        if (c < C) { __stcs(a.t2 + off + c, o[i]); __stcs(a.cur2 + off + c, (o[i] - m2) * r2 * g2[i] + b2[i]); }
      }
      if (lane == 0) *reinterpret_cast<float4*>(a.stats + ((size_t)b * N + n) * 4) = make_float4(mc, rc, m2, r2);
#pragma unroll
      for (int i = 0; i < NI; i++) { hm[i] = hc[i]; xc[i] = xn[i]; yc[i] = yn[i]; hc[i] = hn[i]; xn[i] = xp[i]; }
      mc = mn; rc = rn_;
    }
  }
}

struct LmBwdArgs {
  const float* t; const float* t2; const float* stats; const float* g1; const float* g2; const float* scale;
  const float* dt2; const float* dcur2;      // either may be null (no gradient from that consumer)
  float* dt; float* dg1; float* db1; float* dg2; float* db2;     // parameter gradients are accumulated
  int B, N, C, run;
};

template <int NI>
__global__ void __launch_bounds__(32 * LM_WARPS) ln_metapool_bwd_kernel(const LmBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int C = a.C, N = a.N;
  const int runs_per_b = (N + a.run - 1) / a.run;
  const long long nruns = (long long)a.B * runs_per_b;
  float g1[NI], g2[NI], irc[NI], dg1[NI], db1[NI], dg2[NI], db2[NI];
#pragma unroll
  for (int i = 0; i < NI; i++) {
    const int c = lane + 32 * i;
    const bool ok = c < C;
    g1[i] = ok ? a.g1[c] : 0.f; g2[i] = ok ? a.g2[c] : 0.f;
    irc[i] = ok ? 1.f / (float)(min(c + 1, C - 1) - max(c - 1, 0) + 1) : 0.f;
    dg1[i] = db1[i] = dg2[i] = db2[i] = 0.f;
  }
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nruns; r += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int b = (int)(r / runs_per_b);
    const int n0 = (int)(r - (long long)b * runs_per_b) * a.run, n1 = min(n0 + a.run, N);
    const size_t base = (size_t)b * N * C;
    const float sc = a.scale ? a.scale[b] : 1.f;
    // raw operands of a token, issued one iteration ahead: d(t2), d(cur2), t2 and the saved statistics
    auto load = [&](int n, float (&G)[NI], float (&dc)[NI], float (&v2)[NI], float4& st) {
      const bool in = n >= 0 && n < N;
      const size_t off = base + (size_t)(in ? n : 0) * C;
      st = in ? *reinterpret_cast<const float4*>(a.stats + ((size_t)b * N + n) * 4) : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int i = 0; i < NI; i++) {
        const int c = lane + 32 * i;
        const bool ok = in && c < C;
        G[i] = (ok && a.dt2) ? __ldg(a.dt2 + off + c) : 0.f;
        dc[i] = (ok && a.dcur2) ? __ldg(a.dcur2 + off + c) : 0.f;
        v2[i] = (ok && a.dcur2) ? __ldg(a.t2 + off + c) : 0.f;
      }
    };
    // G <- total gradient with respect to t2 (d(t2) + LayerNorm2 backward of d(cur2)); hw = horizontal sums of G / window count;
    // own: the token belongs to this run (its LayerNorm2 parameter gradients are accumulated here, halo tokens' by their own run)
    auto process = [&](int n, bool own, float (&G)[NI], const float (&dc)[NI], const float (&v2)[NI], const float4& st, float (&hw)[NI]) {
      if (n < 0 || n >= N) {
#pragma unroll
        for (int i = 0; i < NI; i++) G[i] = hw[i] = 0.f;
        return;
      }
      if (a.dcur2) {
        float xh[NI], d[NI];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NI; i++) {
          xh[i] = (lane + 32 * i < C) ? (v2[i] - st.z) * st.w : 0.f;
          if (own) { dg2[i] += dc[i] * xh[i]; db2[i] += dc[i]; }
          d[i] = dc[i] * g2[i];
          s1 += d[i]; s2 += d[i] * xh[i];
        }
        s1 = warp_sum(s1) / C; s2 = warp_sum(s2) / C;
#pragma unroll
        for (int i = 0; i < NI; i++)
          if (lane + 32 * i < C) G[i] += st.w * (d[i] - s1 - xh[i] * s2);
      }
      const float irn = 1.f / (float)(min(n + 1, N - 1) - max(n - 1, 0) + 1);
      float gw[NI];
#pragma unroll
      for (int i = 0; i < NI; i++) gw[i] = G[i] * irn * irc[i];
      hsum3<NI>(gw, hw, lane);
    };
    float Gm[NI], hm[NI], Gc[NI], hc[NI], Gn[NI], hn[NI], dcn[NI], v2n[NI], Gp[NI], dcp[NI], v2p[NI], tx[NI], txn[NI];
    float4 stn, stp, stc, stcn;
    load(n0 - 1, Gm, dcn, v2n, stn);
    process(n0 - 1, false, Gm, dcn, v2n, stn, hm);
    load(n0, Gc, dcn, v2n, stc);
    process(n0, true, Gc, dcn, v2n, stc, hc);
    load(n0 + 1, Gn, dcn, v2n, stn);
#pragma unroll
    for (int i = 0; i < NI; i++) tx[i] = (lane + 32 * i < C) ? __ldg(a.t + base + (size_t)n0 * C + lane + 32 * i) : 0.f;
    for (int n = n0; n < n1; n++) {
      load(n + 2, Gp, dcp, v2p, stp);                   // in flight while token n + 1 and the output of token n are computed
#pragma unroll
      for (int i = 0; i < NI; i++) txn[i] = (n + 1 < n1 && lane + 32 * i < C) ? __ldg(a.t + base + (size_t)(n + 1) * C + lane + 32 * i) : 0.f;
      stcn = stn;
      process(n + 1, n + 1 < n1, Gn, dcn, v2n, stn, hn);
      // d(cur) = s * (pool^T(G) - G), then LayerNorm1 backward
      const size_t off = base + (size_t)n * C;
      float xh[NI], d[NI];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NI; i++) {
        xh[i] = d[i] = 0.f;
        if (lane + 32 * i < C) {
          xh[i] = (tx[i] - stc.x) * stc.y;
          const float dy = sc * (hm[i] + hc[i] + hn[i] - Gc[i]);
          dg1[i] += dy * xh[i]; db1[i] += dy;
          d[i] = dy * g1[i];
          s1 += d[i]; s2 += d[i] * xh[i];
        }
      }
      s1 = warp_sum(s1) / C; s2 = warp_sum(s2) / C;
#pragma unroll
      for (int i = 0; i < NI; i++) {
        const int c = lane + 32 * i;
        if (c < C) __stcs(a.dt + off + c, Gc[i] + stc.y * (d[i] - s1 - xh[i] * s2));
      }
#pragma unroll
      for (int i = 0; i < NI; i++) {
        hm[i] = hc[i]; Gc[i] = Gn[i]; hc[i] = hn[i]; Gn[i] = Gp[i]; dcn[i] = dcp[i]; v2n[i] = v2p[i]; tx[i] = txn[i];
      }
      stc = stcn; stn = stp;
    }
  }
  // parameter gradients: per-lane partials -> per-warp rows of shared memory (8 warps, no float atomics) -> global adds
  __shared__ float part[LM_WARPS][4][32 * LM_MAXI];
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NI; i++) {
    part[warp][0][lane + 32 * i] = dg1[i]; part[warp][1][lane + 32 * i] = db1[i];
    part[warp][2][lane + 32 * i] = dg2[i]; part[warp][3][lane + 32 * i] = db2[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 4 * C; e += blockDim.x) {
    const int k = e / C, c = e - k * C;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LM_WARPS; w++) s += part[w][k][c];
    float* dst = k == 0 ? a.dg1 : (k == 1 ? a.db1 : (k == 2 ? a.dg2 : a.db2));
    if (dst && s != 0.f) atomicAdd(dst + c, s);
  }
}

// ================================================================================================================
// Half-warp layout for C = 16 * CPL (64, 96, 128 channels: MPViT stages 0-2).  The kernels above spend most of their issue slots
// on the two 5-stage warp reductions per LayerNorm and on the channel-neighbour shuffles, which cost the same for 64 channels as
// for 256 (~400 issue slots per token; fetching further ahead changes nothing: 48.6 us on 8x128x128x64 at any depth).  Here 16
// lanes own one token, CPL CONTIGUOUS channels each: the channel neighbours are in the lane's own registers except at the two
// edges (2 shuffles), a reduction is 4 stages and serves two tokens at once, and a token is read and written with 128- / 64-bit
// accesses.  The two halves of a warp walk two different runs in lock step (uniform trip count, results of the padding
// iterations discarded), so every shuffle runs with the full mask.
template <int CPL>
__device__ __forceinline__ void ld_tok(const float* __restrict__ p, float (&x)[CPL]) {
  if constexpr (CPL % 4 == 0) {
#pragma unroll
    for (int k = 0; k < CPL / 4; k++) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + k);
      x[4 * k] = v.x; x[4 * k + 1] = v.y; x[4 * k + 2] = v.z; x[4 * k + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < CPL / 2; k++) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(p) + k);
      x[2 * k] = v.x; x[2 * k + 1] = v.y;
    }
  }
}
template <int CPL>
__device__ __forceinline__ void st_tok(float* __restrict__ p, const float (&x)[CPL]) {
  if constexpr (CPL % 4 == 0) {
#pragma unroll
    for (int k = 0; k < CPL / 4; k++) __stcs(reinterpret_cast<float4*>(p) + k, make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]));
  } else {
#pragma unroll
    for (int k = 0; k < CPL / 2; k++) __stcs(reinterpret_cast<float2*>(p) + k, make_float2(x[2 * k], x[2 * k + 1]));
  }
}
template <int CPL>
__device__ __forceinline__ void zero_tok(float (&x)[CPL]) {
#pragma unroll
  for (int i = 0; i < CPL; i++) x[i] = 0.f;
}
__device__ __forceinline__ float half_sum(float v) {      // over the 16 lanes of a token
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// 3-sum over neighbouring channels, channels j*CPL .. j*CPL+CPL-1 in lane j of the half-warp
template <int CPL>
__device__ __forceinline__ void hsum3c(const float (&v)[CPL], float (&h)[CPL], int j) {
  float left = __shfl_up_sync(0xffffffffu, v[CPL - 1], 1), right = __shfl_down_sync(0xffffffffu, v[0], 1);
  if (j == 0) left = 0.f;
  if (j == 15) right = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; i++) h[i] = (i > 0 ? v[i > 0 ? i - 1 : 0] : left) + v[i] + (i + 1 < CPL ? v[i + 1 < CPL ? i + 1 : i] : right);
}
template <int CPL>
__device__ __forceinline__ void ln_stats16(const float (&x)[CPL], int C, float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; i++) s += x[i];
  mean = half_sum(s) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CPL; i++) { const float d = x[i] - mean; q += d * d; }
  rstd = rsqrtf(half_sum(q) / C + eps);
}

template <int CPL>
__global__ void __launch_bounds__(32 * LM_WARPS) ln_metapool_fwd16_kernel(const LmArgs a) {
  const int lane = threadIdx.x & 31, j = lane & 15, half = lane >> 4;
  const int C = a.C, N = a.N, c0 = j * CPL;
  const int runs_per_b = (N + a.run - 1) / a.run;
  const long long nruns = (long long)a.B * runs_per_b;
  float g1[CPL], b1[CPL], g2[CPL], b2[CPL], irc[CPL];
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    const int c = c0 + i;
    g1[i] = a.g1[c]; b1[i] = a.b1[c]; g2[i] = a.g2[c]; b2[i] = a.b2[c];
    irc[i] = 1.f / (float)(min(c + 1, C - 1) - max(c - 1, 0) + 1);
  }
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long rw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; 2 * rw < nruns; rw += nwarps) {
    const bool active = 2 * rw + half < nruns;
    const long long r = active ? 2 * rw + half : nruns - 1;      // an odd tail: the idle half shadows the last run, nothing stored
    const int b = (int)(r / runs_per_b);
    const int n0 = (int)(r - (long long)b * runs_per_b) * a.run, n1 = min(n0 + a.run, N);
    const float* tb = a.t + (size_t)b * N * C + c0;
    const float sc = a.scale ? a.scale[b] : 1.f;
    auto load = [&](int n, float (&x)[CPL]) {
      if (n >= 0 && n < N && n <= n1) ld_tok<CPL>(tb + (size_t)n * C, x);
      else zero_tok<CPL>(x);
    };
    // LayerNorm1 of a loaded token and its channel sums; a token outside the sample contributes zeros (branch-free: both halves of
    // the warp run the shuffles)
    auto process = [&](int n, const float (&x)[CPL], float (&y)[CPL], float (&h)[CPL], float& mean, float& rstd) {
      const bool in = n >= 0 && n < N;
      ln_stats16<CPL>(x, C, a.eps, mean, rstd);
#pragma unroll
      for (int i = 0; i < CPL; i++) y[i] = in ? (x[i] - mean) * rstd * g1[i] + b1[i] : 0.f;
      hsum3c<CPL>(y, h, j);
    };
    float xm[CPL], ym[CPL], hm[CPL], xc[CPL], yc[CPL], hc[CPL], xn[CPL], yn[CPL], hn[CPL], xp[CPL];
    float mm, rm, mc, rc, mn, rn_;
    load(n0 - 1, xm); load(n0, xc); load(n0 + 1, xn);
    process(n0 - 1, xm, ym, hm, mm, rm);
    process(n0, xc, yc, hc, mc, rc);
    for (int k = 0; k < a.run; k++) {
      const int n = n0 + k;
      const bool live = active && n < n1;
      load(n + 2, xp);                                  // fetching further ahead does not pay (registers: 89.8 -> 118 us fwd + bwd at depth 2)
      process(n + 1, xn, yn, hn, mn, rn_);
      const float irn = 1.f / (float)(min(n + 1, N - 1) - max(n - 1, 0) + 1);
      float o[CPL];
#pragma unroll
      for (int i = 0; i < CPL; i++) o[i] = xc[i] + sc * ((hm[i] + hc[i] + hn[i]) * irn * irc[i] - yc[i]);
      float m2, r2;
      ln_stats16<CPL>(o, C, a.eps, m2, r2);
      if (live) {
        const size_t off = ((size_t)b * N + n) * C + c0;
        st_tok<CPL>(a.t2 + off, o);
        float o2[CPL];
#pragma unroll
        for (int i = 0; i < CPL; i++) o2[i] = (o[i] - m2) * r2 * g2[i] + b2[i];
        st_tok<CPL>(a.cur2 + off, o2);
        if (j == 0) *reinterpret_cast<float4*>(a.stats + ((size_t)b * N + n) * 4) = make_float4(mc, rc, m2, r2);
      }
#pragma unroll
      for (int i = 0; i < CPL; i++) { hm[i] = hc[i]; xc[i] = xn[i]; yc[i] = yn[i]; hc[i] = hn[i]; xn[i] = xp[i]; }
      mc = mn; rc = rn_;
    }
  }
}

template <int CPL>
__global__ void __launch_bounds__(32 * LM_WARPS) ln_metapool_bwd16_kernel(const LmBwdArgs a) {
  const int lane = threadIdx.x & 31, j = lane & 15, half = lane >> 4;
  const int C = a.C, N = a.N, c0 = j * CPL;
  const int runs_per_b = (N + a.run - 1) / a.run;
  const long long nruns = (long long)a.B * runs_per_b;
  float g1[CPL], g2[CPL], irc[CPL], dg1[CPL], db1[CPL], dg2[CPL], db2[CPL];
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    const int c = c0 + i;
    g1[i] = a.g1[c]; g2[i] = a.g2[c];
    irc[i] = 1.f / (float)(min(c + 1, C - 1) - max(c - 1, 0) + 1);
    dg1[i] = db1[i] = dg2[i] = db2[i] = 0.f;
  }
  const bool has_dc = a.dcur2 != nullptr, has_dt2 = a.dt2 != nullptr;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long rw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; 2 * rw < nruns; rw += nwarps) {
    const bool active = 2 * rw + half < nruns;
    const long long r = active ? 2 * rw + half : nruns - 1;
    const int b = (int)(r / runs_per_b);
    const int n0 = (int)(r - (long long)b * runs_per_b) * a.run, n1 = min(n0 + a.run, N);
    const size_t base = (size_t)b * N * C + c0;
    const float sc = a.scale ? a.scale[b] : 1.f;
    auto load = [&](int n, float (&G)[CPL], float (&dc)[CPL], float (&v2)[CPL], float4& st) {
      const bool in = n >= 0 && n < N && n <= n1;
      const size_t off = base + (size_t)(in ? n : 0) * C;
      st = in ? *reinterpret_cast<const float4*>(a.stats + ((size_t)b * N + n) * 4) : make_float4(0, 0, 0, 0);
      if (in && has_dt2) ld_tok<CPL>(a.dt2 + off, G); else zero_tok<CPL>(G);
      if (in && has_dc) { ld_tok<CPL>(a.dcur2 + off, dc); ld_tok<CPL>(a.t2 + off, v2); } else { zero_tok<CPL>(dc); zero_tok<CPL>(v2); }
    };
    // G <- total gradient with respect to t2; hw = channel sums of G / window count.  Branch-free in n: a token outside the
    // sample arrives as zeros (G, dc, v2, st) and leaves as zeros.
    auto process = [&](int n, bool own, float (&G)[CPL], const float (&dc)[CPL], const float (&v2)[CPL], const float4& st, float (&hw)[CPL]) {
      if (has_dc) {
        float xh[CPL], d[CPL];
        float s1 = 0.f, s2 = 0.f;
        const float ow = own ? 1.f : 0.f;
#pragma unroll
        for (int i = 0; i < CPL; i++) {
          xh[i] = (v2[i] - st.z) * st.w;
          dg2[i] += ow * dc[i] * xh[i]; db2[i] += ow * dc[i];
          d[i] = dc[i] * g2[i];
          s1 += d[i]; s2 += d[i] * xh[i];
        }
        s1 = half_sum(s1) / C; s2 = half_sum(s2) / C;
#pragma unroll
        for (int i = 0; i < CPL; i++) G[i] += st.w * (d[i] - s1 - xh[i] * s2);
      }
      const int nc = min(max(n, 0), N - 1);
      const float irn = 1.f / (float)(min(nc + 1, N - 1) - max(nc - 1, 0) + 1);
      float gw[CPL];
#pragma unroll
      for (int i = 0; i < CPL; i++) gw[i] = G[i] * irn * irc[i];
      hsum3c<CPL>(gw, hw, j);
    };
    float Gm[CPL], hm[CPL], Gc[CPL], hc[CPL], Gn[CPL], hn[CPL], dcn[CPL], v2n[CPL], Gp[CPL], dcp[CPL], v2p[CPL], tx[CPL], txn[CPL];
    float4 stn, stp, stc, stcn;
    load(n0 - 1, Gm, dcn, v2n, stn);
    process(n0 - 1, false, Gm, dcn, v2n, stn, hm);
    load(n0, Gc, dcn, v2n, stc);
    process(n0, active, Gc, dcn, v2n, stc, hc);
    load(n0 + 1, Gn, dcn, v2n, stn);
    ld_tok<CPL>(a.t + base + (size_t)n0 * C, tx);
    for (int k = 0; k < a.run; k++) {
      const int n = n0 + k;
      const bool live = active && n < n1;
      load(n + 2, Gp, dcp, v2p, stp);
      if (n + 1 < n1) ld_tok<CPL>(a.t + base + (size_t)(n + 1) * C, txn); else zero_tok<CPL>(txn);
      stcn = stn;
      process(n + 1, active && n + 1 < n1, Gn, dcn, v2n, stn, hn);
      // d(cur) = s * (pool^T(G) - G), then LayerNorm1 backward
      float xh[CPL], d[CPL];
      float s1 = 0.f, s2 = 0.f;
      const float lv = live ? 1.f : 0.f;
#pragma unroll
      for (int i = 0; i < CPL; i++) {
        xh[i] = (tx[i] - stc.x) * stc.y;
        const float dy = sc * (hm[i] + hc[i] + hn[i] - Gc[i]);
        dg1[i] += lv * dy * xh[i]; db1[i] += lv * dy;
        d[i] = dy * g1[i];
        s1 += d[i]; s2 += d[i] * xh[i];
      }
      s1 = half_sum(s1) / C; s2 = half_sum(s2) / C;
      if (live) {
        float o[CPL];
#pragma unroll
        for (int i = 0; i < CPL; i++) o[i] = Gc[i] + stc.y * (d[i] - s1 - xh[i] * s2);
        st_tok<CPL>(a.dt + base + (size_t)n * C, o);
      }
#pragma unroll
      for (int i = 0; i < CPL; i++) {
        hm[i] = hc[i]; Gc[i] = Gn[i]; hc[i] = hn[i]; Gn[i] = Gp[i]; dcn[i] = dcp[i]; v2n[i] = v2p[i]; tx[i] = txn[i];
      }
      stc = stcn; stn = stp;
    }
  }
  // parameter gradients: the two halves of a warp hold partials of the same channels
  __shared__ float part[LM_WARPS][4][32 * LM_MAXI];
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < CPL; i++) {
    const float v0 = dg1[i] + __shfl_xor_sync(0xffffffffu, dg1[i], 16), v1 = db1[i] + __shfl_xor_sync(0xffffffffu, db1[i], 16);
    const float v2 = dg2[i] + __shfl_xor_sync(0xffffffffu, dg2[i], 16), v3 = db2[i] + __shfl_xor_sync(0xffffffffu, db2[i], 16);
    if (half == 0) { part[warp][0][c0 + i] = v0; part[warp][1][c0 + i] = v1; part[warp][2][c0 + i] = v2; part[warp][3][c0 + i] = v3; }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 4 * C; e += blockDim.x) {
    const int k = e / C, c = e - k * C;
    float sum = 0.f;
#pragma unroll
    for (int w = 0; w < LM_WARPS; w++) sum += part[w][k][c];
    float* dst = k == 0 ? a.dg1 : (k == 1 ? a.db1 : (k == 2 ? a.dg2 : a.db2));
    if (dst && sum != 0.f) atomicAdd(dst + c, sum);
  }
}

// channels per lane of the half-warp layout, or 0 when the shape takes the warp-per-token kernels
// (measured, forward / forward + backward in us, warp-per-token -> half-warp: 8x16384x64 48.6 / 134.9 -> 33.5 / 89.8, 8x4096x96
// 19.6 / 49.3 -> 12.5 / 43.2, 8x1024x128 11.2 / 25.8 -> 11.0 / 27.1: the 128-channel backward (201 registers) stays on the old kernel)
static int lm_cpl(int C, bool backward) {
  if (C % 16 != 0) return 0;
  const int cpl = C / 16;
  return (cpl == 4 || cpl == 6 || (cpl == 8 && !backward)) ? cpl : 0;
}
#define LM16_DISPATCH(CPL_EXPR, CALL)                                                                       \
  switch (CPL_EXPR) {                                                                                        \
    case 4: { constexpr int CPL = 4; CALL; } break; case 6: { constexpr int CPL = 6; CALL; } break;         \
    default: { constexpr int CPL = 8; CALL; } break;                                                        \
  }

static int lm_run(int B, int N, bool half_warp) {
  // long runs amortise the two recomputed border tokens; short ones keep every SM busy on the small maps (a warp of the half-warp
  // layout takes two runs, so it needs twice as many of them)
  const long long tokens = (long long)B * N;
  if (half_warp) return tokens >= 131072 ? 16 : (tokens >= 16384 ? 8 : 4);
  return tokens >= 32768 ? 16 : (tokens >= 4096 ? 8 : 4);
}

#define LM_DISPATCH(NI_EXPR, CALL)                                                                         \
  switch (NI_EXPR) {                                                                                        \
    case 1: { constexpr int NI = 1; CALL; } break; case 2: { constexpr int NI = 2; CALL; } break;         \
    case 3: { constexpr int NI = 3; CALL; } break; case 4: { constexpr int NI = 4; CALL; } break;         \
    case 5: { constexpr int NI = 5; CALL; } break; case 6: { constexpr int NI = 6; CALL; } break;         \
    case 7: { constexpr int NI = 7; CALL; } break; default: { constexpr int NI = 8; CALL; } break;        \
  }

// t [B,N,C] tokens; LayerNorm1 (g1, b1), LayerNorm2 (g2, b2), eps; scale [B] DropPath factor (mask / keep) or null;
// out: t2, cur2 [B,N,C]; stats [B*N*4] (mean1, rstd1, mean2, rstd2: what the backward needs)
extern "C" int tcct_ln_metapool_fwd(const float* t, const float* g1, const float* b1, const float* g2, const float* b2, const float* scale,
                                    float* t2, float* cur2, float* stats, int B, int N, int C, float eps, void* stream) {
  TCCT_CHECK_ARG(C >= 2 && C <= 32 * LM_MAXI, "ln_metapool: 2 <= C <= %d expected (got %d)", 32 * LM_MAXI, C);
  LmArgs a{t, g1, b1, g2, b2, scale, t2, cur2, stats, B, N, C, lm_run(B, N, lm_cpl(C, false) != 0), eps};
  const long long nruns = (long long)B * ((N + a.run - 1) / a.run);
  if (const int cpl = lm_cpl(C, false)) {
    int grid = (int)(((nruns + 1) / 2 + LM_WARPS - 1) / LM_WARPS);
    if (grid > tcct_num_sms() * 16) grid = tcct_num_sms() * 16;
    LM16_DISPATCH(cpl, (ln_metapool_fwd16_kernel<CPL><<<grid, 32 * LM_WARPS, 0, (cudaStream_t)stream>>>(a)));
    TCCT_CHECK_LAUNCH("ln_metapool_fwd16");
    return TCCT_OK;
  }
  int grid = (int)((nruns + LM_WARPS - 1) / LM_WARPS);
  if (grid > tcct_num_sms() * 16) grid = tcct_num_sms() * 16;
  LM_DISPATCH((C + 31) / 32, (ln_metapool_fwd_kernel<NI><<<grid, 32 * LM_WARPS, 0, (cudaStream_t)stream>>>(a)));
  TCCT_CHECK_LAUNCH("ln_metapool_fwd");
  return TCCT_OK;
}

// dt2 / dcur2: gradients of the two outputs (either may be null); dt [B,N,C] written; dg1, db1, dg2, db2 [C] accumulated
extern "C" int tcct_ln_metapool_bwd(const float* t, const float* t2, const float* stats, const float* g1, const float* g2, const float* scale,
                                    const float* dt2, const float* dcur2, float* dt, float* dg1, float* db1, float* dg2, float* db2,
                                    int B, int N, int C, void* stream) {
  TCCT_CHECK_ARG(C >= 2 && C <= 32 * LM_MAXI, "ln_metapool: 2 <= C <= %d expected (got %d)", 32 * LM_MAXI, C);
  LmBwdArgs a{t, t2, stats, g1, g2, scale, dt2, dcur2, dt, dg1, db1, dg2, db2, B, N, C, lm_run(B, N, lm_cpl(C, true) != 0)};
  const long long nruns = (long long)B * ((N + a.run - 1) / a.run);
  if (const int cpl = lm_cpl(C, true)) {
    int grid = (int)(((nruns + 1) / 2 + LM_WARPS - 1) / LM_WARPS);
    if (grid > tcct_num_sms() * 8) grid = tcct_num_sms() * 8;
    LM16_DISPATCH(cpl, (ln_metapool_bwd16_kernel<CPL><<<grid, 32 * LM_WARPS, 0, (cudaStream_t)stream>>>(a)));
    TCCT_CHECK_LAUNCH("ln_metapool_bwd16");
    return TCCT_OK;
  }
  int grid = (int)((nruns + LM_WARPS - 1) / LM_WARPS);
  if (grid > tcct_num_sms() * 8) grid = tcct_num_sms() * 8;
  LM_DISPATCH((C + 31) / 32, (ln_metapool_bwd_kernel<NI><<<grid, 32 * LM_WARPS, 0, (cudaStream_t)stream>>>(a)));
  TCCT_CHECK_LAUNCH("ln_metapool_bwd");
  return TCCT_OK;
}
