// Deep-supervision Dice in ONE pass: KiteBack.grad_calc (task1/kite/loopback.py:62-73) calls MultiLoss(DiceLoss)
// (task1/kite/losses/loss.py:83-99, dice 28-32) on the four logit maps FTC.forward returns (task1/nets/tcct.py:1041-1044), three
// of which are bilinear up-samplings (F.interpolate, align_corners=False) of low-resolution auxiliary logits.  Here the
// auxiliary logits are read at their NATIVE resolution and up-sampled in registers, so the three full-resolution maps are never
// written or re-read: forward 4C(1 + 1/4 + 1/16 + 1/64) + 1 bytes per pixel instead of 16C + 4, one launch instead
// of 3 resizes + 4 x (Dice + finalize); the backward writes the head-0 gradient and scatters the auxiliary ones through the
// adjoint of the interpolation into the low-resolution gradient maps (shared-memory tile, then one global add per element).
#include "common.cuh"

#define DM_HEADS 4
#define DM_TW 32
#define DM_TH 8
#define DM_THREADS 256
#define DM_MAXLX (DM_TW + 2)      // low-resolution columns a tile can touch (factor >= 1)

struct DmArgs {
  const float* z[DM_HEADS];      // logits: head 0 [B,C,H,W]; head k [B,C,h_k,w_k]
  int h[DM_HEADS], w[DM_HEADS];
  const unsigned char* lab;      // [B,H,W]
  int B, H, W;
  float weight[DM_HEADS];        // deep-supervision weights (1, coff_ds, coff_ds, coff_ds)
  float sy[DM_HEADS], sx[DM_HEADS];   // h_k / H, w_k / W
  double* sums;                  // [DM_HEADS][2C] inter | sum p, then [C] label counts, then [1] finished-CTA ticket (all zeroed)
  float* loss;                   // out [DM_HEADS + 1]: per-head Dice sums, then the weighted total
  float* coef;                   // out [DM_HEADS][2C]: dL_h/dp_c = ca[c] * g_c + cb[c]
};

struct Tap { int i0, i1; float w1; };
__device__ __forceinline__ Tap dm_tap(int dst, int in, float sc) {     // F.interpolate(bilinear, align_corners=False)
  const float src = fmaxf(sc * ((float)dst + 0.5f) - 0.5f, 0.f);
  Tap t;
  t.i0 = min((int)src, in - 1);
  t.i1 = min(t.i0 + 1, in - 1);
  t.w1 = src - (float)t.i0;
  return t;
}

// softmax over C of the (interpolated) logits of head hd at pixel (b, y, x); v <- probabilities
template <int C>
__device__ __forceinline__ void dm_probs(const DmArgs& a, int hd, int b, int y, int x, float (&v)[C]) {
  const int h = a.h[hd], w = a.w[hd];
  const float* base = a.z[hd] + (size_t)b * C * h * w;
  if (hd == 0) {
#pragma unroll
    for (int c = 0; c < C; c++) v[c] = __ldcs(base + (size_t)c * h * w + (size_t)y * w + x);
  } else {
    const Tap ty = dm_tap(y, h, a.sy[hd]), tx = dm_tap(x, w, a.sx[hd]);
    const float w00 = (1.f - ty.w1) * (1.f - tx.w1), w01 = (1.f - ty.w1) * tx.w1, w10 = ty.w1 * (1.f - tx.w1), w11 = ty.w1 * tx.w1;
    const int o00 = ty.i0 * w + tx.i0, o01 = ty.i0 * w + tx.i1, o10 = ty.i1 * w + tx.i0, o11 = ty.i1 * w + tx.i1;
#pragma unroll
    for (int c = 0; c < C; c++) {
      const float* p = base + (size_t)c * h * w;
      // same association as resize_nchw_fwd_kernel: rows first, then the blend of the two rows
      v[c] = (1.f - ty.w1) * ((1.f - tx.w1) * __ldg(p + o00) + tx.w1 * __ldg(p + o01)) +
             ty.w1 * ((1.f - tx.w1) * __ldg(p + o10) + tx.w1 * __ldg(p + o11));
    }
    (void)w00; (void)w01; (void)w10; (void)w11;
  }
  float m = v[0];
#pragma unroll
  for (int c = 1; c < C; c++) m = fmaxf(m, v[c]);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; c++) { v[c] = __expf(v[c] - m); s += v[c]; }      // |rel err| <= 2^-21: far inside the 1e-3 loss tolerance
  const float inv = __fdividef(1.f, s);
#pragma unroll
  for (int c = 0; c < C; c++) v[c] *= inv;
}

template <int C>
__global__ void __launch_bounds__(DM_THREADS) dice_multi_fwd_kernel(const DmArgs a, int tiles_x, int tiles_y, int ntiles) {
  __shared__ float sred[8][DM_HEADS * 2 * C + C];
  // one head at a time (2C live accumulators instead of 8C: the kernel is bound by the latency of the interpolation gathers,
  // so occupancy matters more than the 4 extra label-byte reads per pixel)
  const int tx = threadIdx.x & (DM_TW - 1), ty = threadIdx.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int hd = 0; hd < DM_HEADS; hd++) {
    float inter[C], psum[C], gsum[C];
#pragma unroll
    for (int c = 0; c < C; c++) inter[c] = psum[c] = gsum[c] = 0.f;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int b = t / (tiles_x * tiles_y);
      const int r = t - b * tiles_x * tiles_y;
      const int y = (r / tiles_x) * DM_TH + ty, x = (r % tiles_x) * DM_TW + tx;
      if (y >= a.H || x >= a.W) continue;
      const int l = a.lab[((size_t)b * a.H + y) * a.W + x];
      float v[C];
      dm_probs<C>(a, hd, b, y, x, v);
#pragma unroll
      for (int c = 0; c < C; c++) {
        psum[c] += v[c]; inter[c] += (c == l) ? v[c] : 0.f;
        if (hd == 0) gsum[c] += (c == l) ? 1.f : 0.f;
      }
    }
    // CTA reduction: warp shuffles, one shared slot per warp, then one double atomic per value and CTA
#pragma unroll
    for (int c = 0; c < C; c++) {
      const float i1 = warp_sum(inter[c]), p1 = warp_sum(psum[c]);
      if (lane == 0) { sred[warp][hd * 2 * C + c] = i1; sred[warp][hd * 2 * C + C + c] = p1; }
      if (hd == 0) {
        const float g1 = warp_sum(gsum[c]);
        if (lane == 0) sred[warp][DM_HEADS * 2 * C + c] = g1;
      }
    }
  }
  __syncthreads();
  constexpr int NV = DM_HEADS * 2 * C + C;
  if (threadIdx.x < NV) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) s += sred[wv][threadIdx.x];
    atomicAdd(a.sums + threadIdx.x, (double)s);
  }
  // the last CTA to finish turns the batch sums into the four losses and the backward coefficients
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(reinterpret_cast<unsigned long long*>(a.sums + NV), 1ull) == (unsigned long long)gridDim.x - 1ull;
  __syncthreads();
  if (last) {      // block-uniform
    // one thread per (head, class): 20-36 threads each fetch three sums and divide (one thread doing all of it serialised 60 L2 round
    // trips and 60 double divisions at the end of the kernel); thread 0 then adds the terms in the same order as before
    __shared__ double s_term[DM_HEADS * C];
    __threadfence();
    if (threadIdx.x < DM_HEADS * C) {
      const int hd = threadIdx.x / C, c = threadIdx.x - hd * C;
      const double I = __ldcg(a.sums + hd * 2 * C + c), U = __ldcg(a.sums + hd * 2 * C + C + c) + __ldcg(a.sums + DM_HEADS * 2 * C + c);
      s_term[threadIdx.x] = 1.0 - (1.0 + 2.0 * I) / (1.0 + U);
      a.coef[hd * 2 * C + c] = (float)(-2.0 / (1.0 + U));
      a.coef[hd * 2 * C + C + c] = (float)((1.0 + 2.0 * I) / ((1.0 + U) * (1.0 + U)));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double total = 0;
      for (int hd = 0; hd < DM_HEADS; hd++) {
        double l = 0;
        for (int c = 0; c < C; c++) l += s_term[hd * C + c];
        a.loss[hd] = (float)l;
        total += (double)a.weight[hd] * (double)(float)l;
      }
      a.loss[DM_HEADS] = (float)total;
    }
  }
}

struct DmBwdArgs {
  DmArgs f;
  const float* gscale;           // dL/d(total), device scalar
  float* dz[DM_HEADS];           // head 0: written; heads 1..3: ZEROED low-resolution maps, accumulated with atomics
};

// Backward: dlogit_j = g * w_h * p_j * (t_j - sum_c p_c t_c), t_c = ca[c] * [c == label] + cb[c].  Head 0 is written in place;
// for an auxiliary head the per-pixel gradient of the UP-SAMPLED logits goes to a shared-memory tile, and the threads then
// gather, for every low-resolution logit the tile touches, the adjoint-interpolation sum over the tile's pixels and add it
// to the low-resolution gradient map (one atomic per element and tile; tiles that share a low-resolution logit each add
// their part).
template <int C>
__global__ void __launch_bounds__(DM_THREADS) dice_multi_bwd_kernel(const DmBwdArgs q, int tiles_x, int tiles_y, int ntiles) {
  extern __shared__ float s_dz[];          // [C][DM_TH][DM_TW] gradient of one head's up-sampled logits over the tile
  float* s_r = s_dz + C * DM_TH * DM_TW;   // [C][DM_TH][DM_MAXLX] row-wise adjoint sums
  __shared__ int s_tx[2 * DM_TW], s_ty[2 * DM_TH];
  __shared__ float s_wx[DM_TW], s_wy[DM_TH];
  const DmArgs& a = q.f;
  const int tx = threadIdx.x & (DM_TW - 1), ty = threadIdx.x >> 5;
  const float g = q.gscale[0];
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int b = t / (tiles_x * tiles_y);
    const int r = t - b * tiles_x * tiles_y;
    const int y0 = (r / tiles_x) * DM_TH, x0 = (r % tiles_x) * DM_TW;
    const int y = y0 + ty, x = x0 + tx;
    const bool in = y < a.H && x < a.W;
    const int l = in ? a.lab[((size_t)b * a.H + y) * a.W + x] : 0;
#pragma unroll
    for (int hd = 0; hd < DM_HEADS; hd++) {
      float d[C];
      if (in) {
        float v[C];
        dm_probs<C>(a, hd, b, y, x, v);
        float tv[C], dot = 0.f;
#pragma unroll
        for (int c = 0; c < C; c++) {
          tv[c] = a.coef[hd * 2 * C + c] * ((c == l) ? 1.f : 0.f) + a.coef[hd * 2 * C + C + c];
          dot += v[c] * tv[c];
        }
        const float gs = g * a.weight[hd];
#pragma unroll
        for (int c = 0; c < C; c++) d[c] = gs * v[c] * (tv[c] - dot);
      } else {
#pragma unroll
        for (int c = 0; c < C; c++) d[c] = 0.f;
      }
      if (hd == 0) {
        if (in) {
          float* o = q.dz[0] + (size_t)b * C * a.H * a.W + (size_t)y * a.W + x;
#pragma unroll
          for (int c = 0; c < C; c++) __stcs(o + (size_t)c * a.H * a.W, d[c]);
        }
        continue;
      }
      __syncthreads();                       // the previous head's gather has finished with the tile
#pragma unroll
      for (int c = 0; c < C; c++) s_dz[(c * DM_TH + ty) * DM_TW + tx] = d[c];
      // interpolation taps of the tile's columns / rows for this head, evaluated once
      const int h = a.h[hd], w = a.w[hd];
      const float sy = a.sy[hd], sx = a.sx[hd];
      if (threadIdx.x < DM_TW) {
        const Tap t2 = dm_tap(min(x0 + (int)threadIdx.x, a.W - 1), w, sx);
        s_tx[threadIdx.x] = t2.i0; s_tx[DM_TW + threadIdx.x] = t2.i1; s_wx[threadIdx.x] = t2.w1;
      } else if (threadIdx.x < DM_TW + DM_TH) {
        const int k = threadIdx.x - DM_TW;
        const Tap t1 = dm_tap(min(y0 + k, a.H - 1), h, sy);
        s_ty[k] = t1.i0; s_ty[DM_TH + k] = t1.i1; s_wy[k] = t1.w1;
      }
      __syncthreads();
      // low-resolution logits touched by the tile: rows [ly0, ly1], columns [lx0, lx1]
      const int ny = min(DM_TH, a.H - y0), nx = min(DM_TW, a.W - x0);
      const int ly0 = s_ty[0], ly1 = s_ty[DM_TH + ny - 1], lx0 = s_tx[0], lx1 = s_tx[DM_TW + nx - 1];
      const int nly = ly1 - ly0 + 1, nlx = lx1 - lx0 + 1;
      const int fx = (a.W + w - 1) / w;
      // stage X: R[c][row][qx] = sum over the row's pixels of wx(pixel -> qx) * dz      (separable adjoint; thread = (row, qx))
      if (ty < ny && tx < nlx) {
        const int qx = lx0 + tx;
        const int xa = max(0, fx * (qx - 1) - x0), xb = min(nx - 1, fx * (qx + 2) - 1 - x0);
        float acc[C];
#pragma unroll
        for (int c = 0; c < C; c++) acc[c] = 0.f;
        for (int xx = xa; xx <= xb; xx++) {
          const float w1 = s_wx[xx];
          const float wgt = (s_tx[xx] == qx ? 1.f - w1 : 0.f) + (s_tx[DM_TW + xx] == qx ? w1 : 0.f);
#pragma unroll
          for (int c = 0; c < C; c++) acc[c] += wgt * s_dz[(c * DM_TH + ty) * DM_TW + xx];
        }
#pragma unroll
        for (int c = 0; c < C; c++) s_r[(c * DM_TH + ty) * DM_MAXLX + tx] = acc[c];
      }
      __syncthreads();
      // stage Y: sum over the tile's rows (thread = (qy, qx)), then one global add per low-resolution logit
      for (int qq = ty; qq < nly; qq += DM_THREADS / DM_TW) {
        if (tx < nlx) {
          const int qy = ly0 + qq;
          float acc[C];
#pragma unroll
          for (int c = 0; c < C; c++) acc[c] = 0.f;
          for (int row = 0; row < ny; row++) {
            const float w1 = s_wy[row];
            const float wgt = (s_ty[row] == qy ? 1.f - w1 : 0.f) + (s_ty[DM_TH + row] == qy ? w1 : 0.f);
            if (wgt != 0.f) {
#pragma unroll
              for (int c = 0; c < C; c++) acc[c] += wgt * s_r[(c * DM_TH + row) * DM_MAXLX + tx];
            }
          }
#pragma unroll
          for (int c = 0; c < C; c++) atomicAdd(q.dz[hd] + ((size_t)b * C + c) * h * w + (size_t)qy * w + (lx0 + tx), acc[c]);
        }
      }
    }
    __syncthreads();
  }
}

#define DM_DISPATCH(C, CALL)                                     \
  switch (C) {                                                   \
    case 2: { constexpr int CC = 2; CALL; } break;               \
    case 3: { constexpr int CC = 3; CALL; } break;               \
    case 4: { constexpr int CC = 4; CALL; } break;               \
    case 5: { constexpr int CC = 5; CALL; } break;               \
    case 6: { constexpr int CC = 6; CALL; } break;               \
    case 7: { constexpr int CC = 7; CALL; } break;               \
    case 8: { constexpr int CC = 8; CALL; } break;               \
    case 9: { constexpr int CC = 9; CALL; } break;               \
    case 10: { constexpr int CC = 10; CALL; } break;             \
    case 11: { constexpr int CC = 11; CALL; } break;             \
    case 12: { constexpr int CC = 12; CALL; } break;             \
    default: tcct_set_error("dice_multi: 2 <= classes <= 12 supported (got %d)", C); return TCCT_ERR_ARG; \
  }

static int dm_fill(DmArgs& a, const float* z0, const float* z1, const float* z2, const float* z3, const int* hs, const int* ws,
                   const unsigned char* lab, int B, int C, int H, int W, const float* weights, double* sums, float* loss, float* coef) {
  a.z[0] = z0; a.z[1] = z1; a.z[2] = z2; a.z[3] = z3;
  a.h[0] = H; a.w[0] = W;
  for (int k = 1; k < DM_HEADS; k++) {
    a.h[k] = hs[k - 1]; a.w[k] = ws[k - 1];
    TCCT_CHECK_ARG(a.h[k] >= 1 && a.w[k] >= 1 && a.h[k] <= H && a.w[k] <= W, "dice_multi: auxiliary head %d is %dx%d for a %dx%d map", k,
                   a.h[k], a.w[k], H, W);
  }
  for (int k = 0; k < DM_HEADS; k++) { a.weight[k] = weights[k]; a.sy[k] = (float)a.h[k] / (float)H; a.sx[k] = (float)a.w[k] / (float)W; }
  a.lab = lab; a.B = B; a.H = H; a.W = W; a.sums = sums; a.loss = loss; a.coef = coef;
  return TCCT_OK;
}

extern "C" long long tcct_dice_multi_sums_doubles(int C) { return (long long)DM_HEADS * 2 * C + C + 1; }

// z0 [B,C,H,W]; z1..z3 [B,C,hs[k],ws[k]] low-resolution auxiliary logits; lab uint8 [B,H,W]; weights[4] (host);
// sums: ZEROED double[tcct_dice_multi_sums_doubles(C)]; loss: float[5] (four Dice sums, weighted total); coef: float[4*2C]
extern "C" int tcct_dice_multi_fwd(const float* z0, const float* z1, const float* z2, const float* z3, const int* hs, const int* ws,
                                   const unsigned char* lab, int B, int C, int H, int W, const float* weights, double* sums,
                                   float* loss, float* coef, void* stream) {
  DmArgs a;
  const int rc = dm_fill(a, z0, z1, z2, z3, hs, ws, lab, B, C, H, W, weights, sums, loss, coef);
  if (rc != TCCT_OK) return rc;
  const int tiles_x = ceil_div(W, DM_TW), tiles_y = ceil_div(H, DM_TH), ntiles = B * tiles_x * tiles_y;
  int grid = tcct_num_sms() * 4;
  if (grid > ntiles) grid = ntiles;
  DM_DISPATCH(C, (dice_multi_fwd_kernel<CC><<<grid, DM_THREADS, 0, (cudaStream_t)stream>>>(a, tiles_x, tiles_y, ntiles)));
  TCCT_CHECK_LAUNCH("dice_multi_fwd");
  return TCCT_OK;
}

// d0: [B,C,H,W] written; d1..d3: ZEROED [B,C,hs[k],ws[k]], accumulated; gscale: device scalar dL/d(total)
extern "C" int tcct_dice_multi_bwd(const float* z0, const float* z1, const float* z2, const float* z3, const int* hs, const int* ws,
                                   const unsigned char* lab, int B, int C, int H, int W, const float* weights, const float* coef,
                                   const float* gscale, float* d0, float* d1, float* d2, float* d3, void* stream) {
  DmBwdArgs q;
  const int rc = dm_fill(q.f, z0, z1, z2, z3, hs, ws, lab, B, C, H, W, weights, nullptr, nullptr, const_cast<float*>(coef));
  if (rc != TCCT_OK) return rc;
  q.gscale = gscale;
  q.dz[0] = d0; q.dz[1] = d1; q.dz[2] = d2; q.dz[3] = d3;
  const int tiles_x = ceil_div(W, DM_TW), tiles_y = ceil_div(H, DM_TH), ntiles = B * tiles_x * tiles_y;
  int grid = tcct_num_sms() * 4;
  if (grid > ntiles) grid = ntiles;
  const size_t smem = (size_t)C * DM_TH * (DM_TW + DM_MAXLX) * sizeof(float);
  DM_DISPATCH(C, (dice_multi_bwd_kernel<CC><<<grid, DM_THREADS, smem, (cudaStream_t)stream>>>(q, tiles_x, tiles_y, ntiles)));
  TCCT_CHECK_LAUNCH("dice_multi_bwd");
  return TCCT_OK;
}
