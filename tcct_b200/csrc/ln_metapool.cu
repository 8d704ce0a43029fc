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

static int lm_run(int B, int N) {
  // long runs amortise the two recomputed border tokens; short ones keep every SM busy on the small maps
  const long long tokens = (long long)B * N;
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
  LmArgs a{t, g1, b1, g2, b2, scale, t2, cur2, stats, B, N, C, lm_run(B, N), eps};
  const long long nruns = (long long)B * ((N + a.run - 1) / a.run);
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
  LmBwdArgs a{t, t2, stats, g1, g2, scale, dt2, dcur2, dt, dg1, db1, dg2, db2, B, N, C, lm_run(B, N)};
  const long long nruns = (long long)B * ((N + a.run - 1) / a.run);
  int grid = (int)((nruns + LM_WARPS - 1) / LM_WARPS);
  if (grid > tcct_num_sms() * 8) grid = tcct_num_sms() * 8;
  LM_DISPATCH((C + 31) / 32, (ln_metapool_bwd_kernel<NI><<<grid, 32 * LM_WARPS, 0, (cudaStream_t)stream>>>(a)));
  TCCT_CHECK_LAUNCH("ln_metapool_bwd");
  return TCCT_OK;
}
