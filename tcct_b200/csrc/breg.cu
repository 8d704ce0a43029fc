// Boundary regression, RegNet.regular_reg (task1/nets/reg.py:109-156; modules at 64-77):
//   x_pred = logits[:,1:], x_true = onehot[:,1:]                                   [B,Cm,H,W], Cm = C-1
//   a      = | dw3x3(dw3x3(x, w0)+b0, w1)+b1 |                                     lap_reg (65-67,115-116)
//   s      = softmax_H( a - log(-log eps)/2 ) / (1e-6 + sum_H)                     sampling_softmax (118-126)
//   m      = sum_c s ;  t1 = conv3x3(m)+b ; t2 = BN_{eps=1}(t1) ; ps = sigmoid(conv3x3(t2)+b)   lap_map (71-76,128-129)
//   edge   = sum_H ps * (h + jitter - .5) / H                                      (146-150)
//   loss   = MSE(edge_p, sg edge_t) + MSE(sg edge_p, edge_t) + MSE(prob_true, softmax_H ps_t) + MSE(prob_true, softmax_H ps_p)
//   prob_true[h] = (label[h] != label[h-1]), row 0 = 0                             (113-114)
// Both branches (0 = pred, 1 = true) run in the same launches.
//
// Every stage is a fully parallel kernel over 32x32 image tiles (stencils, halos recomputed in shared memory) or over
// groups of 8 columns (the reductions over H: 32 row groups x 8 columns per CTA, one 32-byte sector per row), and the
// single-channel intermediates (g, m, t1, ps and their gradients) stay L2-resident between them.  Statistics are
// reduced per CTA before they touch global memory (one double atomic per CTA and quantity).
// sum_H of a normalised softmax is 1 up to rounding: the reference's 1/(1e-6 + sum_H s) is applied as the constant
// BR_INVZ (relative deviation ~1e-7, far inside the fp32 round-off of the column sums themselves).
#include "common.cuh"

#define BR_THREADS 256
#define BT 32            // tile edge
#define XS (BT + 4)      // tile + halo 2
#define XP (XS + 1)      // row pitch of halo-2 tiles
#define US (BT + 2)      // tile + halo 1
#define UP (US + 1)
#define CG 8             // columns per CTA in the column passes
#define RG (BR_THREADS / CG)
#define BR_INVZ (1.f / (1e-6f + 1.f))

__device__ __forceinline__ float sgnf(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }

// partial-over-row-groups -> per-column total, every thread of a column gets the result.  red: [RG][CG]
__device__ __forceinline__ float col_reduce(float part, bool is_max, float* red) {
  const int col = threadIdx.x & (CG - 1);
  __syncthreads();
  red[threadIdx.x] = part;
  __syncthreads();
  float r = is_max ? -INFINITY : 0.f;
#pragma unroll 8
  for (int i = 0; i < RG; i++) r = is_max ? fmaxf(r, red[i * CG + col]) : r + red[i * CG + col];
  return r;
}

// sum over the CTA; result valid in thread 0.  red: [8]
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x == 0)
    for (int i = 0; i < BR_THREADS / 32; i++) r += red[i];
  return r;
}

struct LapArgs {
  const float* logits;            // [B,C,H,W]
  const unsigned char* lab;       // [B,H,W]
  const float* eps;               // [2][B,Cm,H,W]  (pred, true)
  const float* w0; const float* b0; const float* w1; const float* b1;   // lap_reg: [Cm,9],[Cm],[Cm,9],[Cm]
  float* G;                       // [2][B,Cm,H,W]  g = |u2| - log(-log eps)/2
  signed char* sg;                // [2][B,Cm,H,W]  sign(u2)
  float* cmax; float* cinv;       // [2][B,Cm,W]    column max of g, 1 / column sum of exp(g - max)
  const float* dm;                // backward: [2][B,H,W]
  float* r1;                      // backward: [2][B,Cm,W]  sum_h dm * s
  float* dlogits;                 // backward: [B,C,H,W] (channels 1.. written; channel 0 untouched)
  float* dpar;                    // backward: float[Cm*20] accumulators: per channel dw0[9], db0, dw1[9], db1
  int B, C, H, W;
  int tx, ty, ntiles, per_cta;    // tiles per plane (x, y), total, per CTA (contiguous ranges)
};

// Stage the x tile (halo 2) and u1 = dw3x3(x, w0) + b0 (halo 1, zero outside the image: the padding of the 2nd conv).
__device__ __forceinline__ void lap_stage(const LapArgs& a, int branch, int b, int c, int y0, int x0, float* xs, float* u1s,
                                          const float (&w0)[9], float b0) {
  const int H = a.H, W = a.W;
  const float* lg = a.logits + ((size_t)b * a.C + c + 1) * H * W;
  const unsigned char* lb = a.lab + (size_t)b * H * W;
  {
    int ry = threadIdx.x / XS, rx = threadIdx.x - ry * XS;
    for (int i = threadIdx.x; i < XS * XS; i += BR_THREADS) {
      const int y = y0 - 2 + ry, x = x0 - 2 + rx;
      float v = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        if (branch == 0) v = lg[y * W + x];
        else v = lb[y * W + x] == c + 1 ? 1.f : 0.f;
      }
      xs[ry * XP + rx] = v;
      ry += BR_THREADS / XS; rx += BR_THREADS % XS;
      if (rx >= XS) { rx -= XS; ry++; }
    }
  }
  __syncthreads();
  {
    int ry = threadIdx.x / US, rx = threadIdx.x - ry * US;
    for (int i = threadIdx.x; i < US * US; i += BR_THREADS) {
      const int y = y0 - 1 + ry, x = x0 - 1 + rx;
      float v = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        v = b0;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
          for (int kx = 0; kx < 3; kx++) v += w0[ky * 3 + kx] * xs[(ry + ky) * XP + rx + kx];
      }
      u1s[ry * UP + rx] = v;
      ry += BR_THREADS / US; rx += BR_THREADS % US;
      if (rx >= US) { rx -= US; ry++; }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void lap_tile_coords(const LapArgs& a, int t, int& plane, int& branch, int& b, int& c, int& y0, int& x0) {
  const int per_plane = a.tx * a.ty, Cm = a.C - 1;
  plane = t / per_plane;
  const int r = t - plane * per_plane;
  y0 = (r / a.tx) * BT; x0 = (r % a.tx) * BT;
  branch = plane / (a.B * Cm);
  const int q = plane - branch * a.B * Cm;
  b = q / Cm; c = q - b * Cm;
}

// g = |u2| - log(-log eps)/2 and sign(u2) for every pixel of every (branch, image, channel) plane
__global__ void __launch_bounds__(BR_THREADS) breg_lap_g_kernel(const LapArgs a) {
  __shared__ float xs[XS * XP], u1s[US * UP];
  const int t0 = blockIdx.x * a.per_cta, t1 = min(t0 + a.per_cta, a.ntiles);
  const int lx = threadIdx.x & 31, lyb = threadIdx.x >> 5;
  for (int t = t0; t < t1; t++) {
    int plane, branch, b, c, y0, x0;
    lap_tile_coords(a, t, plane, branch, b, c, y0, x0);
    float w0[9], w1[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { w0[k] = a.w0[c * 9 + k]; w1[k] = a.w1[c * 9 + k]; }
    const float b1 = a.b1[c];
    lap_stage(a, branch, b, c, y0, x0, xs, u1s, w0, a.b0[c]);
    const size_t pbase = (size_t)plane * a.H * a.W;
    const int x = x0 + lx;
#pragma unroll
    for (int k = 0; k < BT / 8; k++) {
      const int ly = lyb + 8 * k, y = y0 + ly;
      if (y < a.H && x < a.W) {
        float v = b1;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
          for (int kx = 0; kx < 3; kx++) v += w1[ky * 3 + kx] * u1s[(ly + ky) * UP + lx + kx];
        const size_t o = pbase + (size_t)y * a.W + x;
        a.G[o] = fabsf(v) - 0.5f * __logf(-logf(a.eps[o]));      // the inner log needs full accuracy near eps = 1
        a.sg[o] = (signed char)sgnf(v);
      }
    }
    __syncthreads();
  }
}

// per (plane, column): max_h g and 1 / sum_h exp(g - max)
__global__ void __launch_bounds__(BR_THREADS) breg_colstat_kernel(const float* __restrict__ G, float* cmax, float* cinv, int H, int W) {
  __shared__ float red[BR_THREADS];
  const int col = threadIdx.x & (CG - 1), rg = threadIdx.x / CG;
  const int gcol = min(blockIdx.x * CG + col, W - 1);
  const float* gp = G + (size_t)blockIdx.y * H * W + gcol;
  float mx = -INFINITY;
  for (int r = rg; r < H; r += RG) mx = fmaxf(mx, gp[(size_t)r * W]);
  mx = col_reduce(mx, true, red);
  float sum = 0.f;
  for (int r = rg; r < H; r += RG) sum += expf(gp[(size_t)r * W] - mx);
  sum = col_reduce(sum, false, red);
  if (rg == 0 && blockIdx.x * CG + col < W) {
    cmax[(size_t)blockIdx.y * W + gcol] = mx;
    cinv[(size_t)blockIdx.y * W + gcol] = 1.f / sum;
  }
}

// ---------------------------------------------------------------------------------------------- lap_map
// m = sum_c softmax_H(g_c) / Z  (halo 1 recomputed), t1 = conv3x3(m, wm0) + bm0, per-branch BN statistics of t1
struct Map1Args {
  const float* G; const float* cmax; const float* cinv;
  const float* wm0; const float* bm0;
  float* m; float* t1; double* stats;   // stats: [2 branches][sum, sum of squares]
  int B, Cm, H, W, tx, ty, ntiles, per_cta;
};
__global__ void __launch_bounds__(BR_THREADS) breg_map1_kernel(const Map1Args a) {
  __shared__ float ms[US * UP];
  __shared__ float red[8];
  const int H = a.H, W = a.W;
  const int t0 = blockIdx.x * a.per_cta, t1e = min(t0 + a.per_cta, a.ntiles);
  const int lx = threadIdx.x & 31, lyb = threadIdx.x >> 5;
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; k++) w[k] = a.wm0[k];
  const float bias = a.bm0[0];
  float s[2] = {0.f, 0.f}, q[2] = {0.f, 0.f};
  const int per_plane = a.tx * a.ty;
  for (int t = t0; t < t1e; t++) {
    const int plane2 = t / per_plane;                 // branch * B + b
    const int r = t - plane2 * per_plane;
    const int y0 = (r / a.tx) * BT, x0 = (r % a.tx) * BT;
    const int branch = plane2 / a.B;
    const size_t obase = (size_t)plane2 * H * W;
    for (int i = threadIdx.x; i < US * US; i += BR_THREADS) {
      const int ry = i / US, rx = i - ry * US;
      const int y = y0 - 1 + ry, x = x0 - 1 + rx;
      float v = 0.f;
      if (y >= 0 && y < H && x >= 0 && x < W) {
        for (int c = 0; c < a.Cm; c++) {
          const size_t pl = (size_t)plane2 * a.Cm + c;
          v += expf(a.G[(pl * H + y) * W + x] - a.cmax[pl * W + x]) * a.cinv[pl * W + x];
        }
        v *= BR_INVZ;
        if (ry >= 1 && ry <= BT && rx >= 1 && rx <= BT) a.m[obase + (size_t)y * W + x] = v;
      }
      ms[ry * UP + rx] = v;
    }
    __syncthreads();
    const int x = x0 + lx;
#pragma unroll
    for (int k = 0; k < BT / 8; k++) {
      const int ly = lyb + 8 * k, y = y0 + ly;
      if (y < H && x < W) {
        float v = bias;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
          for (int kx = 0; kx < 3; kx++) v += w[ky * 3 + kx] * ms[(ly + ky) * UP + lx + kx];
        a.t1[obase + (size_t)y * W + x] = v;
        s[branch] += v; q[branch] += v * v;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int br = 0; br < 2; br++) {
    const float ss = block_sum(s[br], red), qq = block_sum(q[br], red);
    if (threadIdx.x == 0 && (ss != 0.f || qq != 0.f)) { atomicAdd(a.stats + br * 2, (double)ss); atomicAdd(a.stats + br * 2 + 1, (double)qq); }
  }
}

// BN(1, eps=1) coefficients of one branch from the batch sums (train) or the running statistics (eval)
__device__ __forceinline__ void breg_bn_coef(const double* stats, double count, float gamma, float beta, float eps, const float* rmean,
                                             const float* rvar, int training, int br, float& sc, float& sh, float& mean, float& invstd) {
  if (training) {
    const double mu = stats[br * 2] / count;
    double var = stats[br * 2 + 1] / count - mu * mu;
    if (var < 0) var = 0;
    mean = (float)mu; invstd = (float)(1.0 / sqrt(var + (double)eps));
  } else {
    mean = rmean[0]; invstd = rsqrtf(rvar[0] + eps);
  }
  sc = gamma * invstd; sh = beta - mean * sc;
}

// ps = sigmoid( conv3x3( BN(t1), wm2 ) + bm2 ).  The BatchNorm finalisation is this kernel's prologue: every thread derives
// its branch's coefficients; block (0,0) also writes coef[branch] = {scale, shift, mean, invstd} for the backward and
// updates the running statistics (pred first, then true: reg.py:128-129).
struct Map2Args {
  const float* t1; const double* stats; double count;
  const float* gamma; const float* beta; float eps, momentum;
  float* rmean; float* rvar; long long* nbt; int training;
  float* coef; const float* wm2; const float* bm2; float* ps;
  int B, H, W;
};
__global__ void __launch_bounds__(BR_THREADS) breg_map2_kernel(const Map2Args a) {
  const int branch = blockIdx.y;
  const int H = a.H, W = a.W;
  const long long n = (long long)a.B * H * W;
  float sc, sh, mean, invstd;
  // eval mode reads the running statistics, which nobody writes then; train mode reads only the batch sums.  One thread per CTA does
  // the double-precision divisions and the square root (every thread doing them costs microseconds per launch on this fp64 rate).
  __shared__ float s_coef[4];
  if (threadIdx.x == 0) {
    breg_bn_coef(a.stats, a.count, a.gamma[0], a.beta[0], a.eps, a.rmean, a.rvar, a.training, branch, sc, sh, mean, invstd);
    s_coef[0] = sc; s_coef[1] = sh; s_coef[2] = mean; s_coef[3] = invstd;
  }
  __syncthreads();
  sc = s_coef[0]; sh = s_coef[1]; mean = s_coef[2]; invstd = s_coef[3];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.coef[branch * 4] = sc; a.coef[branch * 4 + 1] = sh; a.coef[branch * 4 + 2] = mean; a.coef[branch * 4 + 3] = invstd;
    if (branch == 0 && a.training) {
      float rm = a.rmean[0], rv = a.rvar[0];
      for (int br = 0; br < 2; br++) {
        const double mu = a.stats[br * 2] / a.count;
        double var = a.stats[br * 2 + 1] / a.count - mu * mu;
        if (var < 0) var = 0;
        const double unb = a.count > 1 ? var * a.count / (a.count - 1) : var;
        rm = (1.f - a.momentum) * rm + a.momentum * (float)mu;
        rv = (1.f - a.momentum) * rv + a.momentum * (float)unb;
      }
      a.rmean[0] = rm; a.rvar[0] = rv; a.nbt[0] += 2;
    }
  }
  const float* tp = a.t1 + (size_t)branch * n;
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; k++) w[k] = a.wm2[k];
  const float bias = a.bm2[0];
  // image rows are dealt to the CTAs, a thread walks columns threadIdx.x, threadIdx.x + 256, ...: no per-pixel integer divisions
  for (int r = blockIdx.x; r < a.B * H; r += gridDim.x) {
    const int y = r % H;
    const long long base = (long long)(r - y) * W;        // first pixel of the image
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const long long i = base + (long long)y * W + x;
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
    a.ps[(size_t)branch * n + i] = 1.f / (1.f + expf(-v));
    }
  }
}

// Column pass over ps: edge[b,w], softmax_H(ps) vs prob_true squared error (forward) / d ps -> d t3 (backward).
struct ColArgs {
  const float* ps; const unsigned char* lab; const float* jit;   // jit: [2][H]  (pred, true)
  float* edge;                 // [2][B][W]
  double* acc;                 // [0],[1]: sum (sm - pt)^2 per branch
  const float* gout; const float* dedge;   // backward: upstream scalar grad, d loss / d edge [2][B][W] (unscaled)
  float* dt3;                  // backward out [2][B][H][W]
  int B, H, W;
};
template <bool BWD>
__global__ void __launch_bounds__(BR_THREADS) breg_cols_kernel(const ColArgs a) {
  __shared__ float red[BR_THREADS];
  __shared__ float red8[8];
  const int H = a.H, W = a.W;
  const int branch = blockIdx.z, b = blockIdx.y;
  const int col = threadIdx.x & (CG - 1), rg = threadIdx.x / CG;
  const bool incol = blockIdx.x * CG + col < W;
  const int gcol = min(blockIdx.x * CG + col, W - 1);
  const size_t plane = ((size_t)branch * a.B + b) * H * W;
  const float* pp = a.ps + plane + gcol;
  const float* jit = a.jit + branch * H;
  const float invH = 1.f / (float)H;
  float e = 0.f, mx = -INFINITY;
  for (int r = rg; r < H; r += RG) {
    const float v = pp[(size_t)r * W];
    e += v * ((float)r + jit[r] - 0.5f);
    mx = fmaxf(mx, v);
  }
  if (!BWD) {
    e = col_reduce(e, false, red);
    if (rg == 0 && incol) a.edge[((size_t)branch * a.B + b) * W + gcol] = e * invH;
  }
  mx = col_reduce(mx, true, red);
  float sum = 0.f;
  for (int r = rg; r < H; r += RG) sum += expf(pp[(size_t)r * W] - mx);
  sum = col_reduce(sum, false, red);
  const unsigned char* lp = a.lab + (size_t)b * H * W + gcol;
  const float inv = 1.f / sum;
  if (!BWD) {
    float sq = 0.f;
    if (incol)
      for (int r = rg; r < H; r += RG) {
        const float pt = (r > 0 && lp[(size_t)r * W] != lp[(size_t)(r - 1) * W]) ? 1.f : 0.f;
        const float d = expf(pp[(size_t)r * W] - mx) * inv - pt;
        sq += d * d;
      }
    sq = block_sum(sq, red8);
    if (threadIdx.x == 0) atomicAdd(a.acc + branch, (double)sq);
  } else {
    // dL/dsm = 2 (sm - pt) / (B*H*W);  d ps = sm * (dL/dsm - sum_k sm_k dL/dsm_k) + dedge*(h+jit-.5)/H;  dt3 = d ps * ps (1-ps)
    const float kk = 2.f / ((float)a.B * (float)H * (float)W);
    float dot = 0.f;
    for (int r = rg; r < H; r += RG) {
      const float pt = (r > 0 && lp[(size_t)r * W] != lp[(size_t)(r - 1) * W]) ? 1.f : 0.f;
      const float sm = expf(pp[(size_t)r * W] - mx) * inv;
      dot += sm * kk * (sm - pt);
    }
    dot = col_reduce(dot, false, red);
    if (incol) {
      const float g = a.gout[0];
      const float de = a.dedge[((size_t)branch * a.B + b) * W + gcol];
      for (int r = rg; r < H; r += RG) {
        const float pt = (r > 0 && lp[(size_t)r * W] != lp[(size_t)(r - 1) * W]) ? 1.f : 0.f;
        const float p = pp[(size_t)r * W];
        const float sm = expf(p - mx) * inv;
        const float dps = sm * (kk * (sm - pt) - dot) + de * ((float)r + jit[r] - 0.5f) * invH;
        a.dt3[plane + (size_t)r * W + gcol] = g * dps * p * (1.f - p);
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
__global__ void __launch_bounds__(BR_THREADS) breg_map2_bwd_kernel(const float* __restrict__ dt3, const float* __restrict__ t1,
                                                                   const float* coef, const float* wm2, float* __restrict__ dt2,
                                                                   int B, int H, int W, double* bsum, float* gpar) {
  __shared__ float red[8];
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
  for (int r = blockIdx.x; r < B * H; r += gridDim.x) {
    const int y = r % H;
    const long long base = (long long)(r - y) * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const long long i = base + (long long)y * W + x;
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
  }
  s1 = block_sum(s1, red); s2 = block_sum(s2, red);
  if (threadIdx.x == 0) { atomicAdd(bsum + branch * 2, (double)s1); atomicAdd(bsum + branch * 2 + 1, (double)s2); }
#pragma unroll
  for (int k = 0; k < 10; k++) {
    const float v = block_sum(acc[k], red);
    if (threadIdx.x == 0) atomicAdd(gpar + k, v);
  }
}

// Backward through BN and conv3x3(wm0): dt1 = gamma*invstd*(dt2 - S1/n - xhat*S2/n); dm = conv^T(dt1); dwm0, dbm0.
// gpar (float): [dwm0[9], dbm0]
__global__ void __launch_bounds__(BR_THREADS) breg_map1_bwd_kernel(const float* __restrict__ dt2, const float* __restrict__ t1,
                                                                   const float* __restrict__ m, const float* coef, const float* gamma,
                                                                   const double* bsum, int training, const float* wm0,
                                                                   float* __restrict__ dm, int B, int H, int W, float* gpar) {
  __shared__ float red[8];
  const int branch = blockIdx.y;
  const long long n = (long long)B * H * W;
  const float* dp = dt2 + (size_t)branch * n;
  const float* tp = t1 + (size_t)branch * n;
  const float* mp = m + (size_t)branch * n;
  const float mean = coef[branch * 4 + 2], invstd = coef[branch * 4 + 3];
  const float gi = gamma[0] * invstd;
  __shared__ float s_m[2];
  if (threadIdx.x == 0) {      // one double division pair per CTA, not per thread
    s_m[0] = training ? (float)(bsum[branch * 2] / (double)n) : 0.f;
    s_m[1] = training ? (float)(bsum[branch * 2 + 1] / (double)n) : 0.f;
  }
  __syncthreads();
  const float m1 = s_m[0], m2 = s_m[1];
  float w[9];
#pragma unroll
  for (int k = 0; k < 9; k++) w[k] = wm0[k];
  float acc[10];
#pragma unroll
  for (int k = 0; k < 10; k++) acc[k] = 0.f;
  for (int r = blockIdx.x; r < B * H; r += gridDim.x) {
    const int y = r % H;
    const long long base = (long long)(r - y) * W;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const long long i = base + (long long)y * W + x;
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
  }
#pragma unroll
  for (int k = 0; k < 10; k++) {
    const float v = block_sum(acc[k], red);
    if (threadIdx.x == 0) atomicAdd(gpar + k, v);
  }
}

// per (plane, column): r1 = sum_h dm[h] * s[h],  s = exp(g - max) / sum
__global__ void __launch_bounds__(BR_THREADS) breg_lap_r1_kernel(const float* __restrict__ G, const float* __restrict__ cmax,
                                                                 const float* __restrict__ cinv, const float* __restrict__ dm,
                                                                 float* r1, int Cm, int H, int W) {
  __shared__ float red[BR_THREADS];
  const int col = threadIdx.x & (CG - 1), rg = threadIdx.x / CG;
  const int gcol = min(blockIdx.x * CG + col, W - 1);
  const int plane = blockIdx.y;
  const float* gp = G + (size_t)plane * H * W + gcol;
  const float* dp = dm + (size_t)(plane / Cm) * H * W + gcol;
  const float mx = cmax[(size_t)plane * W + gcol], ci = cinv[(size_t)plane * W + gcol];
  float acc = 0.f;
  for (int r = rg; r < H; r += RG) acc += dp[(size_t)r * W] * expf(gp[(size_t)r * W] - mx) * ci;
  acc = col_reduce(acc, false, red);
  if (rg == 0 && blockIdx.x * CG + col < W) r1[(size_t)plane * W + gcol] = acc;
}

// du2 = s * (dm - r1) / Z * sign(u2)  ->  du1 = conv^T(du2, w1)  ->  dx = conv^T(du1, w0); lap_reg parameter gradients
__global__ void __launch_bounds__(BR_THREADS) breg_lap_bwd_kernel(const LapArgs a) {
  __shared__ float xs[XS * XP], u1s[US * UP], du2s[XS * XP], du1s[US * UP];
  __shared__ float spar[20];
  const int H = a.H, W = a.W, Cm = a.C - 1;
  const int t0 = blockIdx.x * a.per_cta, t1 = min(t0 + a.per_cta, a.ntiles);
  const int lx = threadIdx.x & 31, lyb = threadIdx.x >> 5;
  if (threadIdx.x < 20) spar[threadIdx.x] = 0.f;
  float acc[20];
#pragma unroll
  for (int k = 0; k < 20; k++) acc[k] = 0.f;
  int cur_c = -1;
  for (int t = t0; t <= t1; t++) {
    int plane = 0, branch = 0, b = 0, c = -1, y0 = 0, x0 = 0;
    if (t < t1) lap_tile_coords(a, t, plane, branch, b, c, y0, x0);
    if (c != cur_c) {                        // channel change (or end of range): flush the parameter-gradient accumulators
      if (cur_c >= 0) {
#pragma unroll
        for (int k = 0; k < 20; k++) {
          const float v = warp_sum(acc[k]);
          if ((threadIdx.x & 31) == 0) atomicAdd(&spar[k], v);
          acc[k] = 0.f;
        }
        __syncthreads();
        if (threadIdx.x < 20) { atomicAdd(a.dpar + cur_c * 20 + threadIdx.x, spar[threadIdx.x]); spar[threadIdx.x] = 0.f; }
        __syncthreads();
      }
      cur_c = c;
    }
    if (t == t1) break;
    float w0[9], w1[9];
#pragma unroll
    for (int k = 0; k < 9; k++) { w0[k] = a.w0[c * 9 + k]; w1[k] = a.w1[c * 9 + k]; }
    lap_stage(a, branch, b, c, y0, x0, xs, u1s, w0, a.b0[c]);
    const size_t pbase = (size_t)plane * H * W;
    const float* dmp = a.dm + (size_t)(plane / Cm) * H * W;
    {
      const float* Gp = a.G + pbase;
      const signed char* sp = a.sg + pbase;
      const float* cm = a.cmax + (size_t)plane * W;
      const float* cv = a.cinv + (size_t)plane * W;
      const float* rp = a.r1 + (size_t)plane * W;
      int ry = threadIdx.x / XS, rx = threadIdx.x - ry * XS;
      for (int i = threadIdx.x; i < XS * XS; i += BR_THREADS) {
        const int y = y0 - 2 + ry, x = x0 - 2 + rx;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) {
          const int o = y * W + x;
          const float s = __expf(Gp[o] - cm[x]) * cv[x];
          v = s * BR_INVZ * (dmp[o] - rp[x]) * (float)sp[o];
        }
        du2s[ry * XP + rx] = v;
        ry += BR_THREADS / XS; rx += BR_THREADS % XS;
        if (rx >= XS) { rx -= XS; ry++; }
      }
    }
    __syncthreads();
    // du1[y][x] = sum_k w1[ky][kx] * du2[y-ky+1][x-kx+1]   (zero outside the image)
    {
      int ry = threadIdx.x / US, rx = threadIdx.x - ry * US;
      for (int i = threadIdx.x; i < US * US; i += BR_THREADS) {
        const int y = y0 - 1 + ry, x = x0 - 1 + rx;
        float v = 0.f;
        if (y >= 0 && y < H && x >= 0 && x < W) {
#pragma unroll
          for (int ky = 0; ky < 3; ky++)
#pragma unroll
            for (int kx = 0; kx < 3; kx++) v += w1[ky * 3 + kx] * du2s[(ry + 2 - ky) * XP + rx + 2 - kx];
        }
        du1s[ry * UP + rx] = v;
        ry += BR_THREADS / US; rx += BR_THREADS % US;
        if (rx >= US) { rx -= US; ry++; }
      }
    }
    __syncthreads();
    const int x = x0 + lx;
    float* dl = (branch == 0 && a.dlogits) ? a.dlogits + (((size_t)b * a.C + c + 1) * H) * W : nullptr;
#pragma unroll
    for (int k4 = 0; k4 < BT / 8; k4++) {
      const int ly = lyb + 8 * k4, y = y0 + ly;
      if (y < H && x < W) {
        const float du2 = du2s[(ly + 2) * XP + lx + 2];
        const float du1 = du1s[(ly + 1) * UP + lx + 1];
        acc[9] += du1; acc[19] += du2;
        float dx = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ky++)
#pragma unroll
          for (int kx = 0; kx < 3; kx++) {
            acc[10 + ky * 3 + kx] += du2 * u1s[(ly + ky) * UP + lx + kx];          // dw1[k] += du2[p] * u1[p+k-1]
            acc[ky * 3 + kx] += du1 * xs[(ly + ky + 1) * XP + lx + kx + 1];        // dw0[k] += du1[p] * x[p+k-1]
            dx += w0[ky * 3 + kx] * du1s[(ly + 2 - ky) * UP + lx + 2 - kx];        // dx[p] = sum_k w0[k] * du1[p-k+1]
          }
        if (dl) dl[(size_t)y * W + x] = dx;
      }
    }
    __syncthreads();
  }
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

static void tile_grid(int planes, int H, int W, int& tx, int& ty, int& ntiles, int& per_cta, int& ctas) {
  tx = ceil_div(W, BT); ty = ceil_div(H, BT);
  ntiles = planes * tx * ty;
  const int cap = tcct_num_sms() * 6;
  per_cta = ceil_div(ntiles, cap);
  ctas = ceil_div(ntiles, per_cta);
}

// Workspaces.  n = B*H*W, P = 2*B*(C-1) planes.
//   ws (floats): m [2n] | t1 [2n] | ps [2n] | edge [2BW] | dedge [2BW] | coef [8] | G [P*H*W] | cmax [P*W] | cinv [P*W] |
//                sign(u2) bytes [P*H*W]                                   -- nothing needs zeroing
//   dws (doubles, zeroed): stats[4] | acc[2] | bsum[4]
//   bws (floats): dt3 [2n] | dt2 [2n] | dm [2n] | r1 [P*W] | dpar [(C-1)*20] | gpar2 [10] | gpar0 [10]   -- from 6n on zeroed
struct BregWs {
  float *m, *t1, *ps, *edge, *dedge, *coef, *G, *cmax, *cinv;
  signed char* sg;
  double *stats, *acc, *bsum;
};
static BregWs breg_ws(float* ws, double* dws, int B, int C, int H, int W) {
  BregWs r;
  const size_t n2 = (size_t)2 * B * H * W, P = (size_t)2 * B * (C - 1);
  r.m = ws; r.t1 = r.m + n2; r.ps = r.t1 + n2; r.edge = r.ps + n2; r.dedge = r.edge + 2 * B * W; r.coef = r.dedge + 2 * B * W;
  r.G = r.coef + 8; r.cmax = r.G + P * H * W; r.cinv = r.cmax + P * W;
  r.sg = reinterpret_cast<signed char*>(r.cinv + P * W);
  r.stats = dws; r.acc = dws + 4; r.bsum = dws + 6;
  return r;
}
extern "C" long long tcct_breg_ws_floats(int B, int C, int H, int W) {
  const long long n = (long long)B * H * W, P = 2ll * B * (C - 1);
  return 6 * n + 4ll * B * W + 8 + P * H * W + 2 * P * W + (P * H * W + 3) / 4;
}
extern "C" long long tcct_breg_bwd_ws_floats(int B, int C, int H, int W) {
  return (long long)6 * B * H * W + 2ll * B * (C - 1) * W + (long long)(C - 1) * 20 + 20;
}

extern "C" int tcct_breg_forward(const float* logits, const unsigned char* lab, const float* eps, const float* jit,
                                 const float* w0, const float* b0, const float* w1, const float* b1, const float* wm0,
                                 const float* bm0, const float* gamma, const float* beta, const float* wm2,
                                 const float* bm2, float* rmean, float* rvar, long long* nbt, int training, int B, int C,
                                 int H, int W, float* ws, double* dws, float* loss, void* stream) {
  TCCT_CHECK_ARG(C >= 2 && C <= 17, "breg: 2 <= classes <= 17 expected (got %d)", C);
  TCCT_CHECK_ARG(B > 0 && H > 0 && W > 0, "breg: empty input (%d x %d x %d)", B, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  BregWs w = breg_ws(ws, dws, B, C, H, W);
  const int Cm = C - 1, P = 2 * B * Cm;
  LapArgs a{};
  a.logits = logits; a.lab = lab; a.eps = eps; a.w0 = w0; a.b0 = b0; a.w1 = w1; a.b1 = b1;
  a.G = w.G; a.sg = w.sg; a.cmax = w.cmax; a.cinv = w.cinv;
  a.B = B; a.C = C; a.H = H; a.W = W;
  int ctas;
  tile_grid(P, H, W, a.tx, a.ty, a.ntiles, a.per_cta, ctas);
  breg_lap_g_kernel<<<ctas, BR_THREADS, 0, st>>>(a);
  TCCT_CHECK_LAUNCH("breg_lap_g");
  breg_colstat_kernel<<<dim3(ceil_div(W, CG), P), BR_THREADS, 0, st>>>(w.G, w.cmax, w.cinv, H, W);
  TCCT_CHECK_LAUNCH("breg_colstat");
  Map1Args m1{};
  m1.G = w.G; m1.cmax = w.cmax; m1.cinv = w.cinv; m1.wm0 = wm0; m1.bm0 = bm0; m1.m = w.m; m1.t1 = w.t1; m1.stats = w.stats;
  m1.B = B; m1.Cm = Cm; m1.H = H; m1.W = W;
  tile_grid(2 * B, H, W, m1.tx, m1.ty, m1.ntiles, m1.per_cta, ctas);
  breg_map1_kernel<<<ctas, BR_THREADS, 0, st>>>(m1);
  TCCT_CHECK_LAUNCH("breg_map1");
  const long long n = (long long)B * H * W;
  Map2Args m2{};
  m2.t1 = w.t1; m2.stats = w.stats; m2.count = (double)n; m2.gamma = gamma; m2.beta = beta; m2.eps = 1.0f; m2.momentum = 0.1f;
  m2.rmean = rmean; m2.rvar = rvar; m2.nbt = nbt; m2.training = training; m2.coef = w.coef; m2.wm2 = wm2; m2.bm2 = bm2; m2.ps = w.ps;
  m2.B = B; m2.H = H; m2.W = W;
  breg_map2_kernel<<<dim3(B * H < 4 * tcct_num_sms() ? B * H : 4 * tcct_num_sms(), 2), BR_THREADS, 0, st>>>(m2);
  TCCT_CHECK_LAUNCH("breg_map2");
  ColArgs c{};
  c.ps = w.ps; c.lab = lab; c.jit = jit; c.edge = w.edge; c.acc = w.acc;
  c.B = B; c.H = H; c.W = W;
  breg_cols_kernel<false><<<dim3(ceil_div(W, CG), B, 2), BR_THREADS, 0, st>>>(c);
  TCCT_CHECK_LAUNCH("breg_cols");
  breg_final_kernel<<<1, 256, 0, st>>>(w.edge, w.acc, B, H, W, loss, w.dedge);
  TCCT_CHECK_LAUNCH("breg_final");
  return TCCT_OK;
}

// ws/dws: the forward workspaces (unchanged since the forward); bws: float workspace of tcct_breg_bwd_ws_floats, zeroed
// from float 6*B*H*W on; gout: device scalar upstream gradient; dlogits: [B,C,H,W] zero-initialised by the caller.
extern "C" int tcct_breg_backward(const float* logits, const unsigned char* lab, const float* eps, const float* jit,
                                  const float* w0, const float* b0, const float* w1, const float* b1, const float* wm0,
                                  const float* gamma, const float* wm2, int training, int B, int C, int H, int W,
                                  float* ws, double* dws, float* bws, const float* gout, float* dlogits, float* dw0,
                                  float* db0, float* dw1, float* db1, float* dwm0, float* dbm0, float* dgamma,
                                  float* dbeta, float* dwm2, float* dbm2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  BregWs w = breg_ws(ws, dws, B, C, H, W);
  const long long n = (long long)B * H * W;
  const int Cm = C - 1, P = 2 * B * Cm;
  float* dt3 = bws; float* dt2 = dt3 + 2 * n; float* dm = dt2 + 2 * n;
  float* r1 = dm + 2 * n;
  float* dpar = r1 + (size_t)P * W; float* gpar2 = dpar + Cm * 20; float* gpar0 = gpar2 + 10;
  ColArgs c{};
  c.ps = w.ps; c.lab = lab; c.jit = jit; c.edge = w.edge; c.acc = w.acc; c.gout = gout; c.dedge = w.dedge; c.dt3 = dt3;
  c.B = B; c.H = H; c.W = W;
  breg_cols_kernel<true><<<dim3(ceil_div(W, CG), B, 2), BR_THREADS, 0, st>>>(c);
  TCCT_CHECK_LAUNCH("breg_cols_bwd");
  breg_map2_bwd_kernel<<<dim3(B * H < 4 * tcct_num_sms() ? B * H : 4 * tcct_num_sms(), 2), BR_THREADS, 0, st>>>(dt3, w.t1, w.coef, wm2, dt2, B, H, W, w.bsum, gpar2);
  TCCT_CHECK_LAUNCH("breg_map2_bwd");
  breg_map1_bwd_kernel<<<dim3(B * H < 4 * tcct_num_sms() ? B * H : 4 * tcct_num_sms(), 2), BR_THREADS, 0, st>>>(dt2, w.t1, w.m, w.coef, gamma, w.bsum, training, wm0, dm, B, H, W, gpar0);
  TCCT_CHECK_LAUNCH("breg_map1_bwd");
  breg_lap_r1_kernel<<<dim3(ceil_div(W, CG), P), BR_THREADS, 0, st>>>(w.G, w.cmax, w.cinv, dm, r1, Cm, H, W);
  TCCT_CHECK_LAUNCH("breg_lap_r1");
  LapArgs a{};
  a.logits = logits; a.lab = lab; a.eps = eps; a.w0 = w0; a.b0 = b0; a.w1 = w1; a.b1 = b1;
  a.G = w.G; a.sg = w.sg; a.cmax = w.cmax; a.cinv = w.cinv;
  a.dm = dm; a.r1 = r1; a.dlogits = dlogits; a.dpar = dpar;
  a.B = B; a.C = C; a.H = H; a.W = W;
  int ctas;
  tile_grid(P, H, W, a.tx, a.ty, a.ntiles, a.per_cta, ctas);
  breg_lap_bwd_kernel<<<ctas, BR_THREADS, 0, st>>>(a);
  TCCT_CHECK_LAUNCH("breg_lap_bwd");
  breg_param_grads_kernel<<<1, 256, 0, st>>>(dpar, gpar2, gpar0, w.bsum, Cm, dw0, db0, dw1, db1, dwm0, dbm0, dgamma,
                                             dbeta, dwm2, dbm2, training);
  TCCT_CHECK_LAUNCH("breg_param_grads");
  return TCCT_OK;
}
