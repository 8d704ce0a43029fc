// Boundary regression, RegNet.regular_reg (task1/nets/reg.py:109-156; modules at 64-77):
//   x_pred = logits[:,1:], x_true = onehot[:,1:]                                   [B,Cm,H,W], Cm = C-1
//   a      = | dw3x3(dw3x3(x, w0)+b0, w1)+b1 |                                     lap_reg (65-67,115-116)
//   s      = softmax_H( a - log(-log eps)/2 ) / (1e-6 + sum_H)                     sampling_softmax (118-126)
//   m      = sum_c s ;  t1 = conv3x3(m)+b ; t2 = BN_{eps=1}(t1) ; ps = sigmoid(conv3x3(t2)+b)   lap_map (71-76,128-129)
//   edge   = sum_H ps * (h + jitter - .5) / H                                      (146-150)
//   loss   = MSE(edge_p, sg edge_t) + MSE(sg edge_p, edge_t) + MSE(prob_true, softmax_H ps_t) + MSE(prob_true, softmax_H ps_p)
//   prob_true[h] = (label[h] != label[h-1]), row 0 = 0                             (113-114)
// Both branches (0 = pred, 1 = true) run in the same launches (blockIdx.z / leading dim 2).
// Column reductions over H stage an [H x TW] strip in shared memory (coalesced row segments in, strided
// column walks on-chip); forward reads 4(C-1) logit bytes + noise per pixel, the rest is single-channel maps.
#include "common.cuh"

#define BR_THREADS 256

struct BregDims { int B, C, H, W; };

__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// partial-over-rows -> per-column total, all threads of a column get the result.  red: [nr][nc]
__device__ __forceinline__ float col_reduce(float part, bool is_max, float* red, int col, int rg, int nc, int nr, bool active) {
  __syncthreads();
  if (active) red[rg * nc + col] = part;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
  if (active)
    for (int i = 0; i < nr; i++) r = is_max ? fmaxf(r, red[i * nc + col]) : r + red[i * nc + col];
  return r;
}

struct LapArgs {
  const float* logits;            // [B,C,H,W]
  const unsigned char* lab;       // [B,H,W]
  const float* eps;               // [2][B,Cm,H,W]  (pred, true)
  const float* w0; const float* b0; const float* w1; const float* b1;   // lap_reg: [Cm,9],[Cm],[Cm,9],[Cm]
  float* m;                       // [2][B,H,W]   (forward: atomically accumulated over channels; zeroed by caller)
  const float* dm;                // backward: [2][B,H,W]
  float* dlogits;                 // backward: [B,C,H,W] (channels 1.. written; channel 0 untouched)
  float* dpar;                    // backward: float[Cm*20] accumulators: per channel dw0[9], db0, dw1[9], db1
  BregDims d;
  int TW;
};

// shared layout helper
struct LapSmem {
  float *xs, *u1s, *gs, *du1s, *red;
  signed char* sg;
};
__device__ __forceinline__ LapSmem lap_carve(float* base, int H, int NCX, int NCU, int NCG, bool bwd) {
  LapSmem s;
  s.xs = base;
  s.u1s = s.xs + (H + 4) * NCX;
  s.gs = s.u1s + (H + 2) * NCU;
  s.red = s.gs + H * NCG;
  float* nxt = s.red + BR_THREADS;
  s.du1s = nxt;
  s.sg = reinterpret_cast<signed char*>(bwd ? nxt + (H + 2) * NCU : nxt);
  return s;
}
static size_t lap_smem_bytes(int H, int TW, bool bwd) {
  const int E = bwd ? 2 : 0;          // extra columns each side that are recomputed for the backward stencils
  const int NCG = TW + 2 * E, NCU = NCG + 2, NCX = NCG + 4;
  size_t fl = (size_t)(H + 4) * NCX + (size_t)(H + 2) * NCU + (size_t)H * NCG + BR_THREADS;
  if (bwd) fl += (size_t)(H + 2) * NCU;
  size_t bytes = fl * 4;
  if (bwd) bytes += (size_t)H * NCG;
  return (bytes + 15) & ~(size_t)15;
}

// Stage x, compute u1, u2 (sign kept in sg when BWD) and g = |u2| - log(-log eps)/2 into gs.
// Column window: g columns [c0-E, c0+TW+E).  Rows [0,H).
template <bool BWD>
__device__ __forceinline__ void lap_forward_tile(const LapArgs& a, const LapSmem& s, int branch, int b, int c, int c0) {
  const int H = a.d.H, W = a.d.W, Cm = a.d.C - 1;
  constexpr int E = BWD ? 2 : 0;
  const int NCG = a.TW + 2 * E, NCU = NCG + 2, NCX = NCG + 4;
  const int tid = threadIdx.x;
  // x strip: rows [-2,H+2), cols [c0-E-2, c0-E-2+NCX)
  for (int i = tid; i < (H + 4) * NCX; i += BR_THREADS) {
    const int r = i / NCX - 2, col = c0 - E - 2 + i % NCX;
    float v = 0.f;
    if (r >= 0 && r < H && col >= 0 && col < W) {
      if (branch == 0) v = a.logits[(((size_t)b * a.d.C + c + 1) * H + r) * W + col];
      else v = a.lab[((size_t)b * H + r) * W + col] == c + 1 ? 1.f : 0.f;
    }
    s.xs[i] = v;
  }
  float w0[9], w1[9];
#pragma unroll
  for (int k = 0; k < 9; k++) { w0[k] = a.w0[c * 9 + k]; w1[k] = a.w1[c * 9 + k]; }
  const float b0 = a.b0[c], b1 = a.b1[c];
  __syncthreads();
  // u1: rows [-1,H+1), cols [c0-E-1, ...+NCU); zero outside the image (conv zero padding of the 2nd layer)
  for (int i = tid; i < (H + 2) * NCU; i += BR_THREADS) {
    const int ri = i / NCU, ci = i % NCU;
    const int r = ri - 1, col = c0 - E - 1 + ci;
    float v = 0.f;
    if (r >= 0 && r < H && col >= 0 && col < W) {
      v = b0;
#pragma unroll
      for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) v += w0[ky * 3 + kx] * s.xs[(ri + ky) * NCX + ci + kx];
    }
    s.u1s[i] = v;
  }
  __syncthreads();
  const float* ep = a.eps + ((size_t)branch * a.d.B * Cm + (size_t)b * Cm + c) * H * W;
  for (int i = tid; i < H * NCG; i += BR_THREADS) {
    const int r = i / NCG, ci = i % NCG;
    const int col = c0 - E + ci;
    float g = -INFINITY;
    signed char sgv = 0;
    if (col >= 0 && col < W) {
      float v = b1;
#pragma unroll
      for (int ky = 0; ky < 3; ky++)
#pragma unroll
        for (int kx = 0; kx < 3; kx++) v += w1[ky * 3 + kx] * s.u1s[(r + ky) * NCU + ci + kx];
      sgv = (signed char)sgnf(v);
      g = fabsf(v) - 0.5f * logf(-logf(ep[(size_t)r * W + col]));
    }
    s.gs[i] = g;
    if (BWD) s.sg[i] = sgv;
  }
  __syncthreads();      // the column passes below read gs entries written by other threads
}

// softmax over the rows of gs per column; returns S = sum_h s (of the normalised softmax) and leaves s in gs.
__device__ __forceinline__ float lap_col_softmax(const LapSmem& s, int H, int NCG, int col, int rg, int nr, bool active) {
  float mx = -INFINITY;
  if (active) for (int r = rg; r < H; r += nr) mx = fmaxf(mx, s.gs[r * NCG + col]);
  mx = col_reduce(mx, true, s.red, col, rg, NCG, nr, active);
  float sum = 0.f;
  if (active && mx > -INFINITY)
    for (int r = rg; r < H; r += nr) { const float e = expf(s.gs[r * NCG + col] - mx); s.gs[r * NCG + col] = e; sum += e; }
  sum = col_reduce(sum, false, s.red, col, rg, NCG, nr, active);
  float S = 0.f;
  if (active && mx > -INFINITY) {
    const float inv = 1.f / sum;
    for (int r = rg; r < H; r += nr) { const float v = s.gs[r * NCG + col] * inv; s.gs[r * NCG + col] = v; S += v; }
  }
  S = col_reduce(S, false, s.red, col, rg, NCG, nr, active);
  return S;
}

__global__ void __launch_bounds__(BR_THREADS) breg_lap_fwd_kernel(const LapArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.d.H, W = a.d.W, Cm = a.d.C - 1, TW = a.TW;
  const int branch = blockIdx.z, b = blockIdx.y / Cm, c = blockIdx.y % Cm, c0 = blockIdx.x * TW;
  const LapSmem s = lap_carve(smem, H, TW + 4, TW + 2, TW, false);
  lap_forward_tile<false>(a, s, branch, b, c, c0);
  const int nr = BR_THREADS / TW;
  const int col = threadIdx.x % TW, rg = threadIdx.x / TW;
  const bool active = rg < nr && c0 + col < W;
  const float S = lap_col_softmax(s, H, TW, col, rg, nr, active);
  if (active) {
    const float invZ = 1.f / (1e-6f + S);
    float* mp = a.m + ((size_t)branch * a.d.B + b) * H * W;
    for (int r = rg; r < H; r += nr) atomicAdd(mp + (size_t)r * W + c0 + col, s.gs[r * TW + col] * invZ);
  }
}

__global__ void __launch_bounds__(BR_THREADS) breg_lap_bwd_kernel(const LapArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float spar[20];
  const int H = a.d.H, W = a.d.W, Cm = a.d.C - 1, TW = a.TW;
  const int NCG = TW + 4, NCU = NCG + 2, NCX = NCG + 4;
  const int branch = blockIdx.z, b = blockIdx.y / Cm, c = blockIdx.y % Cm, c0 = blockIdx.x * TW;
  const int tid = threadIdx.x;
  const LapSmem s = lap_carve(smem, H, NCX, NCU, NCG, true);
  if (tid < 20) spar[tid] = 0.f;
  lap_forward_tile<true>(a, s, branch, b, c, c0);
  const int nr = BR_THREADS / NCG;
  const int col = tid % NCG, rg = tid / NCG;
  const int gcol = c0 - 2 + col;
  const bool active = rg < nr && gcol >= 0 && gcol < W;
  const float S = lap_col_softmax(s, H, NCG, col, rg, nr, active);
  // ds = dm/Z - R1/Z^2 ;  dg = s * (ds - sum_k s_k ds_k) ;  du2 = dg * sign(u2)
  const float* dmp = a.dm + ((size_t)branch * a.d.B + b) * H * W;
  float r1 = 0.f;
  if (active) for (int r = rg; r < H; r += nr) r1 += dmp[(size_t)r * W + gcol] * s.gs[r * NCG + col];
  r1 = col_reduce(r1, false, s.red, col, rg, NCG, nr, active);
  if (active) {
    const float Z = 1e-6f + S, invZ = 1.f / Z;
    const float sds = r1 * invZ - S * r1 * invZ * invZ;        // sum_k s_k ds_k
    for (int r = rg; r < H; r += nr) {
      const float sv = s.gs[r * NCG + col];
      const float ds = dmp[(size_t)r * W + gcol] * invZ - r1 * invZ * invZ;
      s.gs[r * NCG + col] = sv * (ds - sds) * (float)s.sg[r * NCG + col];
    }
  } else if (rg < nr) {
    for (int r = rg; r < H; r += nr) s.gs[r * NCG + col] = 0.f;     // columns outside the image
  }
  __syncthreads();
  float w0[9], w1[9];
#pragma unroll
  for (int k = 0; k < 9; k++) { w0[k] = a.w0[c * 9 + k]; w1[k] = a.w1[c * 9 + k]; }
  // du1: rows [-1,H+1) x cols [c0-1, c0+TW+1)  (index space of u1s shifted by one column: u1s col ci <-> c0-3+ci)
  // du1[y][x] = sum_k w1[ky][kx] * du2[y-ky+1][x-kx+1]
  const int NCD = TW + 2;
  for (int i = tid; i < (H + 2) * NCD; i += BR_THREADS) {
    const int r = i / NCD - 1, ci = i % NCD;
    const int colg = c0 - 1 + ci;
    float v = 0.f;
    if (r >= 0 && r < H && colg >= 0 && colg < W) {
#pragma unroll
      for (int ky = 0; ky < 3; ky++) {
        const int yy = r - ky + 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; kx++) {
          const int gc = ci + 2 - kx;        // gs column index of global col (colg - kx + 1): (colg-kx+1) - (c0-2)
          v += w1[ky * 3 + kx] * s.gs[yy * NCG + gc];
        }
      }
    }
    s.du1s[i] = v;
  }
  __syncthreads();
  // owned positions: parameter gradients and dx
  float acc[20];
#pragma unroll
  for (int k = 0; k < 20; k++) acc[k] = 0.f;
  float* dl = a.dlogits ? a.dlogits + (((size_t)b * a.d.C + c + 1) * H) * W : nullptr;
  for (int i = tid; i < H * TW; i += BR_THREADS) {
    const int r = i / TW, ci = i % TW;
    const int colg = c0 + ci;
    if (colg >= W) continue;
    const float du2 = s.gs[r * NCG + ci + 2];
    const float du1 = s.du1s[(r + 1) * NCD + ci + 1];
    acc[9] += du1; acc[19] += du2;
    float dx = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ky++)
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        // dw1[k] += du2[p] * u1[p+k-1] ; u1s index: row (r+ky-1)+1, col (colg+kx-1) - (c0-3)
        acc[10 + ky * 3 + kx] += du2 * s.u1s[(r + ky) * NCU + ci + kx + 2];
        // dw0[k] += du1[p] * x[p+k-1] ; xs index: row (r+ky-1)+2, col (colg+kx-1) - (c0-4)
        acc[ky * 3 + kx] += du1 * s.xs[(r + ky + 1) * NCX + ci + kx + 3];
        // dx[p] = sum_k w0[k] * du1[p-k+1]
        const int yy = r - ky + 1;
        if (yy >= 0 && yy < H) dx += w0[ky * 3 + kx] * s.du1s[(yy + 1) * NCD + ci + 2 - kx];
      }
    if (branch == 0 && dl) dl[(size_t)r * W + colg] = dx;
  }
#pragma unroll
  for (int k = 0; k < 20; k++) {
    const float v = warp_sum(acc[k]);
    if ((tid & 31) == 0) atomicAdd(&spar[k], v);
  }
  __syncthreads();
  if (tid < 20) atomicAdd(a.dpar + c * 20 + tid, spar[tid]);
}

// ---------------------------------------------------------------------------------------------- lap_map
// t1 = conv3x3(m, wm0) + bm0  and per-branch BN statistics
__global__ void breg_map1_kernel(const float* __restrict__ m, const float* wm0, const float* bm0, float* __restrict__ t1,
                                 int B, int H, int W, double* stats /*[2][2]*/) {
  const int branch = blockIdx.y;
  const long long n = (long long)B * H * W;
  const float* mp = m + (size_t)branch * n;
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; k++) w[k] = wm0[k];
  const float bias = bm0[0];
  float s = 0.f, q = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long base = i - (long long)y * W - x;
    float v = bias;
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W) continue;
        v += w[ky * 3 + kx] * __ldg(mp + base + (long long)yy * W + xx);
      }
    }
    t1[(size_t)branch * n + i] = v;
    s += v; q += v * v;
  }
  s = warp_sum(s); q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) { atomicAdd(stats + branch * 2, (double)s); atomicAdd(stats + branch * 2 + 1, (double)q); }
}

// BN(1, eps=1) coefficients per branch + running-statistics update (pred first, then true: reg.py:128-129)
// coef[branch] = {scale, shift, mean, invstd}
__global__ void breg_bn_kernel(const double* stats, double count, const float* gamma, const float* beta, float eps,
                               float momentum, float* rmean, float* rvar, long long* nbt, int training, float* coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int br = 0; br < 2; br++) {
    float mean, invstd;
    if (training) {
      const double mu = stats[br * 2] / count;
      double var = stats[br * 2 + 1] / count - mu * mu;
      if (var < 0) var = 0;
      mean = (float)mu; invstd = (float)(1.0 / sqrt(var + (double)eps));
      const double unb = count > 1 ? var * count / (count - 1) : var;
      rmean[0] = (1.f - momentum) * rmean[0] + momentum * (float)mu;
      rvar[0] = (1.f - momentum) * rvar[0] + momentum * (float)unb;
      nbt[0] += 1;
    } else {
      mean = rmean[0]; invstd = rsqrtf(rvar[0] + eps);
    }
    const float sc = gamma[0] * invstd;
    coef[br * 4] = sc; coef[br * 4 + 1] = beta[0] - mean * sc; coef[br * 4 + 2] = mean; coef[br * 4 + 3] = invstd;
  }
}

// ps = sigmoid( conv3x3( BN(t1), wm2 ) + bm2 )
__global__ void breg_map2_kernel(const float* __restrict__ t1, const float* coef, const float* wm2, const float* bm2,
                                 float* __restrict__ ps, int B, int H, int W) {
  const int branch = blockIdx.y;
  const long long n = (long long)B * H * W;
  const float* tp = t1 + (size_t)branch * n;
  const float sc = coef[branch * 4], sh = coef[branch * 4 + 1];
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; k++) w[k] = wm2[k];
  const float bias = bm2[0];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long base = i - (long long)y * W - x;
    float v = bias;
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= W) continue;
        v += w[ky * 3 + kx] * (__ldg(tp + base + (long long)yy * W + xx) * sc + sh);
      }
    }
    ps[(size_t)branch * n + i] = 1.f / (1.f + expf(-v));
  }
}

// Column pass over ps: edge[b,w], softmax_H(ps) vs prob_true squared error (forward) / d ps -> d t3 (backward).
struct ColArgs {
  const float* ps; const unsigned char* lab; const float* jit;   // jit: [2][H]  (pred, true)
  float* edge;                 // [2][B][W]
  double* acc;                 // [0],[1]: sum (sm - pt)^2 per branch
  const float* gout; const float* dedge;   // backward: upstream scalar grad, d loss / d edge [2][B][W] (unscaled)
  float* dt3;                  // backward out [2][B][H][W]
  int B, H, W, TW;
};
template <bool BWD>
__global__ void __launch_bounds__(BR_THREADS) breg_cols_kernel(const ColArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, W = a.W, TW = a.TW;
  float* tile = smem;                    // [H][TW]  ps, then softmax
  float* red = tile + H * TW;            // [BR_THREADS]
  float* ptile = red + BR_THREADS;       // [H][TW]  copy of ps (backward only)
  const int branch = blockIdx.z, b = blockIdx.y, c0 = blockIdx.x * TW;
  const int tid = threadIdx.x;
  const size_t plane = ((size_t)branch * a.B + b) * H * W;
  for (int i = tid; i < H * TW; i += BR_THREADS) {
    const int r = i / TW, col = c0 + i % TW;
    const float v = col < W ? a.ps[plane + (size_t)r * W + col] : 0.f;
    tile[i] = v;
    if (BWD) ptile[i] = v;
  }
  const int nr = BR_THREADS / TW, col = tid % TW, rg = tid / TW;
  const bool active = rg < nr && c0 + col < W;
  const float* jit = a.jit + branch * H;
  const float invH = 1.f / (float)H;
  float e = 0.f, mx = -INFINITY;
  __syncthreads();
  if (active)
    for (int r = rg; r < H; r += nr) {
      const float v = tile[r * TW + col];
      e += v * ((float)r + jit[r] - 0.5f);
      mx = fmaxf(mx, v);
    }
  if (!BWD) {
    e = col_reduce(e, false, red, col, rg, TW, nr, active);
    if (active && rg == 0) a.edge[((size_t)branch * a.B + b) * W + c0 + col] = e * invH;
  }
  mx = col_reduce(mx, true, red, col, rg, TW, nr, active);
  float sum = 0.f;
  if (active)
    for (int r = rg; r < H; r += nr) { const float ex = expf(tile[r * TW + col] - mx); tile[r * TW + col] = ex; sum += ex; }
  sum = col_reduce(sum, false, red, col, rg, TW, nr, active);
  const unsigned char* lp = a.lab + (size_t)b * H * W + c0 + col;
  const float inv = active ? 1.f / sum : 0.f;
  if (!BWD) {
    float sq = 0.f;
    if (active)
      for (int r = rg; r < H; r += nr) {
        const float pt = (r > 0 && lp[(size_t)r * W] != lp[(size_t)(r - 1) * W]) ? 1.f : 0.f;
        const float d = tile[r * TW + col] * inv - pt;
        sq += d * d;
      }
    sq = warp_sum(sq);
    if ((tid & 31) == 0) atomicAdd(a.acc + branch, (double)sq);
  } else {
    // dL/dsm = 2 (sm - pt) / (B*H*W);  d ps = sm * (dL/dsm - sum_k sm_k dL/dsm_k) + dedge*(h+jit-.5)/H;  dt3 = d ps * ps (1-ps)
    const float kk = 2.f / ((float)a.B * (float)H * (float)W);
    float dot = 0.f;
    if (active)
      for (int r = rg; r < H; r += nr) {
        const float pt = (r > 0 && lp[(size_t)r * W] != lp[(size_t)(r - 1) * W]) ? 1.f : 0.f;
        const float sm = tile[r * TW + col] * inv;
        dot += sm * kk * (sm - pt);
      }
    dot = col_reduce(dot, false, red, col, rg, TW, nr, active);
    if (active) {
      const float g = a.gout[0];
      const float de = a.dedge[((size_t)branch * a.B + b) * W + c0 + col];
      for (int r = rg; r < H; r += nr) {
        const float pt = (r > 0 && lp[(size_t)r * W] != lp[(size_t)(r - 1) * W]) ? 1.f : 0.f;
        const float sm = tile[r * TW + col] * inv;
        const float p = ptile[r * TW + col];
        const float dps = sm * (kk * (sm - pt) - dot) + de * ((float)r + jit[r] - 0.5f) * invH;
        a.dt3[plane + (size_t)r * W + c0 + col] = g * dps * p * (1.f - p);
      }
    }
  }
}

// loss and d loss / d edge
__global__ void breg_final_kernel(const float* edge, const double* acc, int B, int H, int W, float* loss, float* dedge) {
  __shared__ float red[32];
  const int n = B * W;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = edge[i] - edge[n + i];      // pred - true
    s += d * d;
    dedge[i] = 2.f * d / (float)n;
    dedge[n + i] = -2.f * d / (float)n;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      const double npx = (double)B * H * W;
      *loss = (float)(2.0 * (double)s / (double)n + (acc[0] + acc[1]) / npx);
    }
  }
}

// Backward through conv3x3(wm2): dt2 = conv^T(dt3), dwm2, dbm2, and the BN reduction sums S1 = sum dt2, S2 = sum dt2*xhat
// bsum (double): per branch [S1, S2];  gpar (float): [dwm2[9], dbm2]
__global__ void breg_map2_bwd_kernel(const float* __restrict__ dt3, const float* __restrict__ t1, const float* coef,
                                     const float* wm2, float* __restrict__ dt2, int B, int H, int W, double* bsum,
                                     float* gpar) {
  __shared__ float sp[10];
  if (threadIdx.x < 10) sp[threadIdx.x] = 0.f;
  __syncthreads();
  const int branch = blockIdx.y;
  const long long n = (long long)B * H * W;
  const float* dp = dt3 + (size_t)branch * n;
  const float* tp = t1 + (size_t)branch * n;
  const float sc = coef[branch * 4], sh = coef[branch * 4 + 1], mean = coef[branch * 4 + 2], invstd = coef[branch * 4 + 3];
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; k++) w[k] = wm2[k];
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; k++) acc[k] = 0.f;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long base = i - (long long)y * W - x;
    const float d3 = dp[i];
    acc[9] += d3;
    float v = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        // forward: t3[p] += w[k] * t2[p+k-1]  =>  dt2[q] += w[k] * dt3[q-k+1] ;  dw[k] += dt3[p] * t2[p+k-1]
        const int yb = y - ky + 1, xb = x - kx + 1;
        if (yb >= 0 && yb < H && xb >= 0 && xb < W) v += w[ky * 3 + kx] * __ldg(dp + base + (long long)yb * W + xb);
        const int yf = y + ky - 1, xf = x + kx - 1;
        if (yf >= 0 && yf < H && xf >= 0 && xf < W) acc[ky * 3 + kx] += d3 * (__ldg(tp + base + (long long)yf * W + xf) * sc + sh);
      }
    }
    dt2[(size_t)branch * n + i] = v;
    s1 += v;
    s2 += v * (tp[i] - mean) * invstd;
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) { atomicAdd(bsum + branch * 2, (double)s1); atomicAdd(bsum + branch * 2 + 1, (double)s2); }
#pragma unroll
  for (int k = 0; k < 10; k++) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sp[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 10) atomicAdd(gpar + threadIdx.x, sp[threadIdx.x]);
}

// Backward through BN and conv3x3(wm0): dt1 = gamma*invstd*(dt2 - S1/n - xhat*S2/n); dm = conv^T(dt1); dwm0, dbm0.
// gpar (float): [dwm0[9], dbm0]
__global__ void breg_map1_bwd_kernel(const float* __restrict__ dt2, const float* __restrict__ t1, const float* __restrict__ m,
                                     const float* coef, const float* gamma, const double* bsum, int training,
                                     const float* wm0, float* __restrict__ dm, int B, int H, int W, float* gpar) {
  __shared__ float sp[10];
  if (threadIdx.x < 10) sp[threadIdx.x] = 0.f;
  __syncthreads();
  const int branch = blockIdx.y;
  const long long n = (long long)B * H * W;
  const float* dp = dt2 + (size_t)branch * n;
  const float* tp = t1 + (size_t)branch * n;
  const float* mp = m + (size_t)branch * n;
  const float mean = coef[branch * 4 + 2], invstd = coef[branch * 4 + 3];
  const float gi = gamma[0] * invstd;
  const float m1 = training ? (float)(bsum[branch * 2] / (double)n) : 0.f;
  const float m2 = training ? (float)(bsum[branch * 2 + 1] / (double)n) : 0.f;
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; k++) w[k] = wm0[k];
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; k++) acc[k] = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long base = i - (long long)y * W - x;
    const float d1 = gi * (dp[i] - m1 - (tp[i] - mean) * invstd * m2);
    acc[9] += d1;
    float v = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
#pragma unroll
      for (int kx = 0; kx < 3; kx++) {
        const int yb = y - ky + 1, xb = x - kx + 1;
        if (yb >= 0 && yb < H && xb >= 0 && xb < W) {
          const long long j = base + (long long)yb * W + xb;
          v += w[ky * 3 + kx] * gi * (__ldg(dp + j) - m1 - (__ldg(tp + j) - mean) * invstd * m2);
        }
        const int yf = y + ky - 1, xf = x + kx - 1;
        if (yf >= 0 && yf < H && xf >= 0 && xf < W) acc[ky * 3 + kx] += d1 * __ldg(mp + base + (long long)yf * W + xf);
      }
    }
    dm[(size_t)branch * n + i] = v;
  }
#pragma unroll
  for (int k = 0; k < 10; k++) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sp[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 10) atomicAdd(gpar + threadIdx.x, sp[threadIdx.x]);
}

// scatter the accumulated parameter gradients into the parameter gradient buffers (accumulate)
__global__ void breg_param_grads_kernel(const float* dpar_lap, const float* gpar2, const float* gpar0, const double* bsum,
                                        int Cm, float* dw0, float* db0, float* dw1, float* db1, float* dwm0, float* dbm0,
                                        float* dgamma, float* dbeta, float* dwm2, float* dbm2, int training) {
  const int t = threadIdx.x;
  for (int i = t; i < Cm * 20; i += blockDim.x) {
    const int c = i / 20, k = i % 20;
    const float v = dpar_lap[i];
    if (k < 9) dw0[c * 9 + k] += v;
    else if (k == 9) db0[c] += v;
    else if (k < 19) dw1[c * 9 + k - 10] += v;
    else db1[c] += v;
  }
  if (t < 9) { dwm2[t] += gpar2[t]; dwm0[t] += gpar0[t]; }
  if (t == 9) { dbm2[0] += gpar2[9]; dbm0[0] += gpar0[9]; }
  if (t == 10 && training) {
    dgamma[0] += (float)(bsum[1] + bsum[3]);
    dbeta[0] += (float)(bsum[0] + bsum[2]);
  }
}

static int pick_tw(int H, bool bwd) {
  const int cands[3] = {32, 16, 8};
  for (int i = 0; i < 3; i++)
    if (lap_smem_bytes(H, cands[i], bwd) <= 220 * 1024) return cands[i];
  return 0;
}
static int ew_blocks(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)tcct_num_sms() * 8;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// Workspace (floats), zeroed by the caller before the forward:
//   m [2*B*H*W] | t1 [2*B*H*W] | ps [2*B*H*W] | edge [2*B*W] | dedge [2*B*W] | coef [8]
// dws (doubles, zeroed): stats[4] | acc[2] | bsum[4]
struct BregWs {
  float *m, *t1, *ps, *edge, *dedge, *coef;
  double *stats, *acc, *bsum;
};
static BregWs breg_ws(float* ws, double* dws, int B, int H, int W) {
  BregWs r;
  const size_t n2 = (size_t)2 * B * H * W;
  r.m = ws; r.t1 = r.m + n2; r.ps = r.t1 + n2; r.edge = r.ps + n2; r.dedge = r.edge + 2 * B * W; r.coef = r.dedge + 2 * B * W;
  r.stats = dws; r.acc = dws + 4; r.bsum = dws + 6;
  return r;
}
extern "C" long long tcct_breg_ws_floats(int B, int H, int W) { return (long long)6 * B * H * W + 4ll * B * W + 8; }
extern "C" long long tcct_breg_bwd_ws_floats(int B, int C, int H, int W) {
  return (long long)6 * B * H * W + (long long)(C - 1) * 20 + 20;     // dt3 | dt2 | dm | dpar_lap | gpar2 | gpar0
}

extern "C" int tcct_breg_forward(const float* logits, const unsigned char* lab, const float* eps, const float* jit,
                                 const float* w0, const float* b0, const float* w1, const float* b1, const float* wm0,
                                 const float* bm0, const float* gamma, const float* beta, const float* wm2,
                                 const float* bm2, float* rmean, float* rvar, long long* nbt, int training, int B, int C,
                                 int H, int W, float* ws, double* dws, float* loss, void* stream) {
  TCCT_CHECK_ARG(C >= 2 && C <= 17, "breg: 2 <= classes <= 17 expected (got %d)", C);
  cudaStream_t st = (cudaStream_t)stream;
  const int TW = pick_tw(H, false);
  TCCT_CHECK_ARG(TW > 0, "breg: H = %d is too tall for the shared-memory column strip", H);
  BregWs w = breg_ws(ws, dws, B, H, W);
  LapArgs a;
  a.logits = logits; a.lab = lab; a.eps = eps; a.w0 = w0; a.b0 = b0; a.w1 = w1; a.b1 = b1;
  a.m = w.m; a.dm = nullptr; a.dlogits = nullptr; a.dpar = nullptr;
  a.d = BregDims{B, C, H, W}; a.TW = TW;
  const size_t smem = lap_smem_bytes(H, TW, false);
  cudaFuncSetAttribute(breg_lap_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  breg_lap_fwd_kernel<<<dim3(ceil_div(W, TW), B * (C - 1), 2), BR_THREADS, smem, st>>>(a);
  TCCT_CHECK_LAUNCH("breg_lap_fwd");
  const long long n = (long long)B * H * W;
  breg_map1_kernel<<<dim3(ew_blocks(n), 2), 256, 0, st>>>(w.m, wm0, bm0, w.t1, B, H, W, w.stats);
  TCCT_CHECK_LAUNCH("breg_map1");
  breg_bn_kernel<<<1, 32, 0, st>>>(w.stats, (double)n, gamma, beta, 1.0f, 0.1f, rmean, rvar, nbt, training, w.coef);
  TCCT_CHECK_LAUNCH("breg_bn");
  breg_map2_kernel<<<dim3(ew_blocks(n), 2), 256, 0, st>>>(w.t1, w.coef, wm2, bm2, w.ps, B, H, W);
  TCCT_CHECK_LAUNCH("breg_map2");
  ColArgs c;
  c.ps = w.ps; c.lab = lab; c.jit = jit; c.edge = w.edge; c.acc = w.acc; c.gout = nullptr; c.dedge = nullptr; c.dt3 = nullptr;
  c.B = B; c.H = H; c.W = W; c.TW = 32;
  const size_t csm = ((size_t)H * 32 + BR_THREADS) * 4;
  TCCT_CHECK_ARG(csm <= 220 * 1024, "breg: H too large");
  cudaFuncSetAttribute(breg_cols_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm);
  breg_cols_kernel<false><<<dim3(ceil_div(W, 32), B, 2), BR_THREADS, csm, st>>>(c);
  TCCT_CHECK_LAUNCH("breg_cols");
  breg_final_kernel<<<1, 256, 0, st>>>(w.edge, w.acc, B, H, W, loss, w.dedge);
  TCCT_CHECK_LAUNCH("breg_final");
  return TCCT_OK;
}

// ws/dws: the forward workspaces (unchanged since the forward); bws: zeroed float workspace of
// tcct_breg_bwd_ws_floats; gout: device scalar upstream gradient; dlogits: [B,C,H,W] zero-initialised by the caller.
extern "C" int tcct_breg_backward(const float* logits, const unsigned char* lab, const float* eps, const float* jit,
                                  const float* w0, const float* b0, const float* w1, const float* b1, const float* wm0,
                                  const float* gamma, const float* wm2, int training, int B, int C, int H, int W,
                                  float* ws, double* dws, float* bws, const float* gout, float* dlogits, float* dw0,
                                  float* db0, float* dw1, float* db1, float* dwm0, float* dbm0, float* dgamma,
                                  float* dbeta, float* dwm2, float* dbm2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int TW = pick_tw(H, true);
  TCCT_CHECK_ARG(TW > 0, "breg: H = %d is too tall for the shared-memory column strip", H);
  BregWs w = breg_ws(ws, dws, B, H, W);
  const long long n = (long long)B * H * W;
  float* dt3 = bws; float* dt2 = dt3 + 2 * n; float* dm = dt2 + 2 * n;
  float* dpar = dm + 2 * n; float* gpar2 = dpar + (C - 1) * 20; float* gpar0 = gpar2 + 10;
  ColArgs c;
  c.ps = w.ps; c.lab = lab; c.jit = jit; c.edge = w.edge; c.acc = w.acc; c.gout = gout; c.dedge = w.dedge; c.dt3 = dt3;
  c.B = B; c.H = H; c.W = W; c.TW = 32;
  const size_t csm = ((size_t)2 * H * 32 + BR_THREADS) * 4;
  TCCT_CHECK_ARG(csm <= 220 * 1024, "breg: H too large");
  cudaFuncSetAttribute(breg_cols_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm);
  breg_cols_kernel<true><<<dim3(ceil_div(W, 32), B, 2), BR_THREADS, csm, st>>>(c);
  TCCT_CHECK_LAUNCH("breg_cols_bwd");
  breg_map2_bwd_kernel<<<dim3(ew_blocks(n), 2), 256, 0, st>>>(dt3, w.t1, w.coef, wm2, dt2, B, H, W, w.bsum, gpar2);
  TCCT_CHECK_LAUNCH("breg_map2_bwd");
  breg_map1_bwd_kernel<<<dim3(ew_blocks(n), 2), 256, 0, st>>>(dt2, w.t1, w.m, w.coef, gamma, w.bsum, training, wm0, dm, B, H, W, gpar0);
  TCCT_CHECK_LAUNCH("breg_map1_bwd");
  LapArgs a;
  a.logits = logits; a.lab = lab; a.eps = eps; a.w0 = w0; a.b0 = b0; a.w1 = w1; a.b1 = b1;
  a.m = nullptr; a.dm = dm; a.dlogits = dlogits; a.dpar = dpar;
  a.d = BregDims{B, C, H, W}; a.TW = TW;
  const size_t smem = lap_smem_bytes(H, TW, true);
  cudaFuncSetAttribute(breg_lap_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  breg_lap_bwd_kernel<<<dim3(ceil_div(W, TW), B * (C - 1), 2), BR_THREADS, smem, st>>>(a);
  TCCT_CHECK_LAUNCH("breg_lap_bwd");
  breg_param_grads_kernel<<<1, 256, 0, st>>>(dpar, gpar2, gpar0, w.bsum, C - 1, dw0, db0, dw1, db1, dwm0, dbm0, dgamma,
                                             dbeta, dwm2, dbm2, training);
  TCCT_CHECK_LAUNCH("breg_param_grads");
  return TCCT_OK;
}
