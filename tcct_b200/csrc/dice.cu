// MultiLoss(DiceLoss) / MultiLoss(MSELoss) of the `--los` registry (task1/kite/losses/loss.py:70-110,
// dice at 28-32): softmax over classes, then per class 1-(1+2*sum(p*g))/(1+sum(p)+sum(g)) with sums over the
// WHOLE batch, summed over classes.  Logits NCHW fp32, labels as a uint8 index map (one byte per pixel).
// Bandwidth-bound: forward reads C logits + 1 label byte per pixel, backward re-reads them and writes C grads.
#include "common.cuh"

#define DICE_MAXC 16

__global__ void onehot_to_index_kernel(const long long* __restrict__ onehot, unsigned char* __restrict__ lab, int B,
                                       int C, int HW) {
  const long long n = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int q = (int)(i - (long long)b * HW);
    int best = 0;
    long long bv = onehot[((size_t)b * C) * HW + q];
    for (int c = 1; c < C; c++) {
      const long long v = onehot[((size_t)b * C + c) * HW + q];
      if (v > bv) { bv = v; best = c; }
    }
    lab[i] = (unsigned char)best;
  }
}
__global__ void index64_to_u8_kernel(const long long* __restrict__ idx, unsigned char* __restrict__ lab, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    lab[i] = (unsigned char)idx[i];
}

extern "C" int tcct_onehot_to_index(const long long* onehot, unsigned char* lab, int B, int C, int HW, void* stream) {
  const long long n = (long long)B * HW;
  int blocks = ceil_div(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  onehot_to_index_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(onehot, lab, B, C, HW);
  TCCT_CHECK_LAUNCH("onehot_to_index");
  return TCCT_OK;
}
extern "C" int tcct_index64_to_u8(const long long* idx, unsigned char* lab, long long n, void* stream) {
  int blocks = ceil_div(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  index64_to_u8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(idx, lab, n);
  TCCT_CHECK_LAUNCH("index64_to_u8");
  return TCCT_OK;
}

// sums layout (double): [0,C) inter, [C,2C) sum p, [2C,3C) sum g, [3C] sum of squared error (mse mode)
template <int C>
__global__ void __launch_bounds__(256) dice_fwd_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ lab,
                                                       int B, int HW, double* sums) {
  __shared__ float sred[3 * DICE_MAXC + 1];
  if (threadIdx.x < 3 * DICE_MAXC + 1) sred[threadIdx.x] = 0.f;
  __syncthreads();
  float inter[C], psum[C], gsum[C];
  float sq = 0.f;
#pragma unroll
  for (int c = 0; c < C; c++) inter[c] = psum[c] = gsum[c] = 0.f;
  const long long n = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int q = (int)(i - (long long)b * HW);
    const float* lp = logits + ((size_t)b * C) * HW + q;
    float v[C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; c++) { v[c] = lp[(size_t)c * HW]; m = fmaxf(m, v[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) { v[c] = expf(v[c] - m); s += v[c]; }
    const float inv = 1.f / s;
    const int l = lab[i];
#pragma unroll
    for (int c = 0; c < C; c++) {
      const float p = v[c] * inv;
      const float g = (c == l) ? 1.f : 0.f;
      psum[c] += p; inter[c] += p * g; gsum[c] += g;
      sq += (p - g) * (p - g);
    }
  }
#pragma unroll
  for (int c = 0; c < C; c++) {
    const float a = warp_sum(inter[c]), b2 = warp_sum(psum[c]), g2 = warp_sum(gsum[c]);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&sred[c], a); atomicAdd(&sred[DICE_MAXC + c], b2); atomicAdd(&sred[2 * DICE_MAXC + c], g2);
    }
  }
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) atomicAdd(&sred[3 * DICE_MAXC], sq);
  __syncthreads();
  if (threadIdx.x < C) {
    atomicAdd(sums + threadIdx.x, (double)sred[threadIdx.x]);
    atomicAdd(sums + C + threadIdx.x, (double)sred[DICE_MAXC + threadIdx.x]);
    atomicAdd(sums + 2 * C + threadIdx.x, (double)sred[2 * DICE_MAXC + threadIdx.x]);
  }
  if (threadIdx.x == 0) atomicAdd(sums + 3 * C, (double)sred[3 * DICE_MAXC]);
}

// loss (scalar) and backward coefficients:  dL/dp_c(px) = ca[c]*g_c(px) + cb[c]   (dice)
//                                           dL/dp_c(px) = cm*(p_c - g_c)           (mse; cm in coef[2C])
__global__ void dice_finalize_kernel(const double* sums, int C, double npix, int mode, float* loss, float* coef) {
  if (threadIdx.x == 0) {
    double l = 0;
    if (mode == 0) {
      for (int c = 0; c < C; c++) {
        const double I = sums[c], U = sums[C + c] + sums[2 * C + c];
        l += 1.0 - (1.0 + 2.0 * I) / (1.0 + U);
        coef[c] = (float)(-2.0 / (1.0 + U));
        coef[C + c] = (float)((1.0 + 2.0 * I) / ((1.0 + U) * (1.0 + U)));
      }
      coef[2 * C] = 0.f;
    } else {
      l = sums[3 * C] / npix;
      for (int c = 0; c < 2 * C; c++) coef[c] = 0.f;
      coef[2 * C] = (float)(2.0 / npix);
    }
    *loss = (float)l;
  }
}

// dlogit_j = gscale * p_j * (t_j - sum_c p_c t_c),  t_c = dL/dp_c
template <int C>
__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ lab,
                                                       const float* __restrict__ coef, const float* __restrict__ gscale,
                                                       float weight, float* __restrict__ dlogits, int B, int HW,
                                                       int accumulate) {
  float ca[C], cb[C];
#pragma unroll
  for (int c = 0; c < C; c++) { ca[c] = coef[c]; cb[c] = coef[C + c]; }
  const float cm = coef[2 * C];
  const float gs = gscale[0] * weight;
  const long long n = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int q = (int)(i - (long long)b * HW);
    const size_t base = ((size_t)b * C) * HW + q;
    float v[C];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; c++) { v[c] = logits[base + (size_t)c * HW]; m = fmaxf(m, v[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) { v[c] = expf(v[c] - m); s += v[c]; }
    const float inv = 1.f / s;
    const int l = lab[i];
    float tv[C];
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; c++) {
      v[c] *= inv;
      const float g = (c == l) ? 1.f : 0.f;
      tv[c] = ca[c] * g + cb[c] + cm * (v[c] - g);
      dot += v[c] * tv[c];
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
      const float d = gs * v[c] * (tv[c] - dot);
      if (accumulate) dlogits[base + (size_t)c * HW] += d;
      else dlogits[base + (size_t)c * HW] = d;
    }
  }
}

template <int C>
static void launch_fwd(const float* logits, const unsigned char* lab, int B, int HW, double* sums, cudaStream_t st) {
  int blocks = ceil_div((long long)B * HW, 256);
  const int cap = tcct_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  dice_fwd_kernel<C><<<blocks, 256, 0, st>>>(logits, lab, B, HW, sums);
}
template <int C>
static void launch_bwd(const float* logits, const unsigned char* lab, const float* coef, const float* gscale, float weight,
                       float* dlogits, int B, int HW, int accumulate, cudaStream_t st) {
  int blocks = ceil_div((long long)B * HW, 256);
  const int cap = tcct_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  dice_bwd_kernel<C><<<blocks, 256, 0, st>>>(logits, lab, coef, gscale, weight, dlogits, B, HW, accumulate);
}

#define DICE_DISPATCH(C, CALL)                                   \
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
    default: tcct_set_error("dice: 2 <= classes <= 12 supported (got %d)", C); return TCCT_ERR_ARG; \
  }

// sums: zeroed double[3C+1];  loss: float[1];  coef: float[2C+1]
extern "C" int tcct_dice_fwd(const float* logits, const unsigned char* lab, int B, int C, int HW, int mode, double* sums,
                             float* loss, float* coef, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DICE_DISPATCH(C, launch_fwd<CC>(logits, lab, B, HW, sums, st));
  TCCT_CHECK_LAUNCH("dice_fwd");
  dice_finalize_kernel<<<1, 32, 0, st>>>(sums, C, (double)B * HW * 1.0, mode, loss, coef);
  TCCT_CHECK_LAUNCH("dice_finalize");
  return TCCT_OK;
}
extern "C" int tcct_dice_bwd(const float* logits, const unsigned char* lab, int B, int C, int HW, const float* coef,
                             const float* gscale, float weight, float* dlogits, int accumulate, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DICE_DISPATCH(C, launch_bwd<CC>(logits, lab, coef, gscale, weight, dlogits, B, HW, accumulate, st));
  TCCT_CHECK_LAUNCH("dice_bwd");
  return TCCT_OK;
}

// KiteSeg.predict (task1/kite/loop_seg.py:21-33): label = argmax_c softmax(logits) = argmax_c logits,
// first maximum wins (torch.argmax).  NCHW logits -> uint8 [B,H,W].
__global__ void argmax_nchw_kernel(const float* __restrict__ logits, unsigned char* __restrict__ lab, int B, int C, int HW) {
  const long long n = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const int q = (int)(i - (long long)b * HW);
    const float* lp = logits + ((size_t)b * C) * HW + q;
    int best = 0;
    float bv = lp[0];
    for (int c = 1; c < C; c++) {
      const float v = lp[(size_t)c * HW];
      if (v > bv) { bv = v; best = c; }
    }
    lab[i] = (unsigned char)best;
  }
}
extern "C" int tcct_argmax_nchw(const float* logits, unsigned char* lab, int B, int C, int HW, void* stream) {
  TCCT_CHECK_ARG(C >= 1 && C <= 255, "argmax_nchw: 1 <= C <= 255 expected");
  int blocks = ceil_div((long long)B * HW, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  argmax_nchw_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, lab, B, C, HW);
  TCCT_CHECK_LAUNCH("argmax_nchw");
  return TCCT_OK;
}

// soft_argmax (task1/nets/reg.py:27-35): out[b,0,p] = sum_c c * softmax_C(beta * logits)[c]  (a differentiable label index;
// unused by the reference's training path, part of the inference contract of SURVEY 8a I2).
__global__ void soft_argmax_kernel(const float* __restrict__ logits, float* __restrict__ out, int B, int C, int HW, float beta) {
  const long long n = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const float* lp = logits + ((size_t)b * C) * HW + (i - (long long)b * HW);
    float m = -INFINITY;
    for (int c = 0; c < C; c++) m = fmaxf(m, beta * lp[(size_t)c * HW]);
    float s = 0.f, t = 0.f;
    for (int c = 0; c < C; c++) {
      const float e = expf(beta * lp[(size_t)c * HW] - m);
      s += e; t += (float)c * e;
    }
    out[i] = t / s;
  }
}
extern "C" int tcct_soft_argmax(const float* logits, float* out, int B, int C, int HW, float beta, void* stream) {
  TCCT_CHECK_ARG(C >= 1 && C <= 255, "soft_argmax: 1 <= C <= 255 expected");
  const long long n = (long long)B * HW;
  int blocks = (int)((n + 255) / 256);
  if (blocks > tcct_num_sms() * 8) blocks = tcct_num_sms() * 8;
  soft_argmax_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, out, B, C, HW, beta);
  TCCT_CHECK_LAUNCH("soft_argmax");
  return TCCT_OK;
}

// Soft-argmax boundary extraction (inference; the reference has no implementation beyond soft_argmax and step (6)
// of regular_reg, so the definition below is this repository's, pinned only by its own CPU restatement in oracle/):
//   p = softmax_C(logits);  d_c[h] = |p_c[h] - p_c[h-1]| (d_c[0] = 0);  pos[b,c-1,w] = sum_h h * softmax_H(beta * d_c)[h]   for c >= 1
// Eight columns x 32 row blocks per CTA (one 32-byte sector per row and class plane), two passes over the column (max, then the
// normalised sums): the logits of a batch fit the L2, the second pass does not touch HBM.
#define BP_CG 8
#define BP_RG 32
#define BP_MAXC 16
// softmax_C at one pixel, all classes
template <int CC>
__device__ __forceinline__ void bp_probs(const float* lp, int C, size_t HW, float (&p)[CC]) {
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < CC; k++) { p[k] = k < C ? lp[(size_t)k * HW] : -INFINITY; m = fmaxf(m, p[k]); }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CC; k++) { p[k] = k < C ? expf(p[k] - m) : 0.f; s += p[k]; }
  const float inv = 1.f / s;
#pragma unroll
  for (int k = 0; k < CC; k++) p[k] *= inv;
}
// A thread owns one column and a contiguous block of rows, for ALL classes: one softmax_C per pixel and pass; the row above is
// carried over.  red: [CC][BP_RG][BP_CG]
template <int CC>
__global__ void __launch_bounds__(BP_CG * BP_RG) boundary_pos_kernel(const float* __restrict__ logits, float* __restrict__ pos, int B, int C,
                                                                     int H, int W, float beta) {
  __shared__ float red[CC * BP_CG * BP_RG];
  const int col = threadIdx.x & (BP_CG - 1), rg = threadIdx.x / BP_CG;
  const int b = blockIdx.y;
  const int gcol = min(blockIdx.x * BP_CG + col, W - 1);
  const size_t HW = (size_t)H * W;
  const float* lp = logits + (size_t)b * C * HW + gcol;
  const int R = (H + BP_RG - 1) / BP_RG;
  const int h0 = rg * R, h1 = min(h0 + R, H);
  float prev[CC], cur[CC], mx[CC], s[CC], t[CC];
#pragma unroll
  for (int k = 0; k < CC; k++) { mx[k] = -INFINITY; s[k] = t[k] = 0.f; prev[k] = 0.f; }
  // per-class column reduction over the 32 row blocks; every thread of a column gets the result
  auto colred = [&](float (&v)[CC], bool is_max) {
    __syncthreads();
#pragma unroll
    for (int k = 1; k < CC; k++) red[(k * BP_RG + rg) * BP_CG + col] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 1; k < CC; k++) {
      float r = is_max ? -INFINITY : 0.f;
      for (int i = 0; i < BP_RG; i++) { const float u = red[(k * BP_RG + i) * BP_CG + col]; r = is_max ? fmaxf(r, u) : r + u; }
      v[k] = r;
    }
  };
  if (h0 > 0 && h0 < H) bp_probs<CC>(lp + (size_t)(h0 - 1) * W, C, HW, prev);
  for (int h = h0; h < h1; h++) {
    bp_probs<CC>(lp + (size_t)h * W, C, HW, cur);
#pragma unroll
    for (int k = 1; k < CC; k++) { mx[k] = fmaxf(mx[k], h > 0 ? beta * fabsf(cur[k] - prev[k]) : 0.f); prev[k] = cur[k]; }
  }
  colred(mx, true);
  if (h0 > 0 && h0 < H) bp_probs<CC>(lp + (size_t)(h0 - 1) * W, C, HW, prev);
  for (int h = h0; h < h1; h++) {
    bp_probs<CC>(lp + (size_t)h * W, C, HW, cur);
#pragma unroll
    for (int k = 1; k < CC; k++) {
      const float e = expf((h > 0 ? beta * fabsf(cur[k] - prev[k]) : 0.f) - mx[k]);
      s[k] += e; t[k] += (float)h * e;
      prev[k] = cur[k];
    }
  }
  colred(s, false);
  colred(t, false);
  if (rg == 0 && blockIdx.x * BP_CG + col < W) {
#pragma unroll
    for (int k = 1; k < CC; k++)
      if (k < C) pos[((size_t)b * (C - 1) + (k - 1)) * W + gcol] = t[k] / s[k];
  }
}
extern "C" int tcct_boundary_positions(const float* logits, float* pos, int B, int C, int H, int W, float beta, void* stream) {
  TCCT_CHECK_ARG(C >= 2 && C <= BP_MAXC, "boundary_positions: 2 <= C <= %d expected (got %d)", BP_MAXC, C);
  TCCT_CHECK_ARG(B > 0 && H > 0 && W > 0, "boundary_positions: empty input");
  const dim3 grid((W + BP_CG - 1) / BP_CG, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 5) boundary_pos_kernel<5><<<grid, BP_CG * BP_RG, 0, st>>>(logits, pos, B, C, H, W, beta);
  else if (C <= 9) boundary_pos_kernel<9><<<grid, BP_CG * BP_RG, 0, st>>>(logits, pos, B, C, H, W, beta);
  else boundary_pos_kernel<BP_MAXC><<<grid, BP_CG * BP_RG, 0, st>>>(logits, pos, B, C, H, W, beta);
  TCCT_CHECK_LAUNCH("boundary_positions");
  return TCCT_OK;
}

// Validation counts (MDiceLoss.score / MIouLoss.score, task1/kite/losses/miou.py:28-44,69-91, on one-hot argmax maps):
// counts[b][c] = { |pred==c & true==c|, |pred==c|, |true==c| }.   One block per (image, chunk).
__global__ void label_counts_kernel(const unsigned char* __restrict__ pred, const unsigned char* __restrict__ truth, int HW,
                                    int C, int* counts) {
  __shared__ int sc[3 * DICE_MAXC];
  const int b = blockIdx.y;
  if (threadIdx.x < 3 * DICE_MAXC) sc[threadIdx.x] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int p = pred[(size_t)b * HW + i], t = truth[(size_t)b * HW + i];
    if (p < C) atomicAdd(&sc[p * 3 + 1], 1);
    if (t < C) atomicAdd(&sc[t * 3 + 2], 1);
    if (p == t && p < C) atomicAdd(&sc[p * 3], 1);
  }
  __syncthreads();
  if (threadIdx.x < 3 * C && sc[threadIdx.x]) atomicAdd(counts + (size_t)b * C * 3 + threadIdx.x, sc[threadIdx.x]);
}
// counts: zeroed int32 [B][C][3]
extern "C" int tcct_label_counts(const unsigned char* pred, const unsigned char* truth, int B, int C, int HW, int* counts,
                                 void* stream) {
  TCCT_CHECK_ARG(C >= 1 && C <= DICE_MAXC, "label_counts: 1 <= C <= %d expected", DICE_MAXC);
  int bx = ceil_div(HW, 256 * 8);
  if (bx < 1) bx = 1;
  label_counts_kernel<<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(pred, truth, HW, C, counts);
  TCCT_CHECK_LAUNCH("label_counts");
  return TCCT_OK;
}

// Validation scores on arbitrary (soft or hard) maps -- MIouLoss.score / MDiceLoss.score of task1/kite/losses/miou.py:28-44,69-91 reduce
// sum(pr * gt), sum(pr), sum(gt) per image and class plane; out[b][c] = {inter, sum pr, sum gt} (double, zeroed by the caller).
// One CTA row per plane (blockIdx.y = b * C + c), float4 loads, one atomic triple per CTA.
template <typename GT>
__global__ void __launch_bounds__(256) score_sums_kernel(const float* __restrict__ pr, const GT* __restrict__ gt, int HW, double* out) {
  const size_t base = (size_t)blockIdx.y * HW;
  float si = 0.f, sp = 0.f, sg = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float p = pr[base + i], g = (float)gt[base + i];
    si += p * g; sp += p; sg += g;
  }
  __shared__ float red[3][8];
  si = warp_sum(si); sp = warp_sum(sp); sg = warp_sum(sg);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = si; red[1][threadIdx.x >> 5] = sp; red[2][threadIdx.x >> 5] = sg; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; w++) s += red[threadIdx.x][w];
    atomicAdd(out + (size_t)blockIdx.y * 3 + threadIdx.x, (double)s);
  }
}
// pr: float [B,C,H,W]; gt: float (gt_is_i64 = 0) or int64 (1) [B,C,H,W]; out: zeroed double [B*C*3]
extern "C" int tcct_score_sums(const float* pr, const void* gt, int gt_is_i64, int B, int C, int HW, double* out, void* stream) {
  TCCT_CHECK_ARG(B >= 1 && C >= 1 && (long long)B * C <= 65535, "score_sums: 1 <= B*C <= 65535 expected");
  int gx = ceil_div(HW, 256 * 8);
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  dim3 grid(gx, B * C);
  if (gt_is_i64) score_sums_kernel<long long><<<grid, 256, 0, (cudaStream_t)stream>>>(pr, (const long long*)gt, HW, out);
  else score_sums_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(pr, (const float*)gt, HW, out);
  TCCT_CHECK_LAUNCH("score_sums");
  return TCCT_OK;
}
