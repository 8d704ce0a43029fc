// Dense contraction kernels of the stc_tt path on the warp-level TF32 tensor-core path:
//   * conv_tile_kernel : spatial convs (3x3, 1xk, kx1; Cin 32|64 -> Cout 32-tiles), forward and dgrad
//   * gemm_px_kernel   : 1x1 convs / Linear over pixels (K = Cin up to 320), forward and dgrad
//   * wgrad_kernel     : weight/bias gradients for both (contraction over pixels)
//   * pack_weights     : one launch re-packs every weight of the model into MMA-fragment order
// Activations are NHWC fp32; accumulation is fp32; operands are rounded to TF32 (cvt.rna).
// Reference semantics: nn.Conv2d / nn.Linear as used by task1/nets/tcct.py:803-828 (CrossCNNBlock),
// 55-97 (Conv2d_BN), 29-53 (Mlp), 887-914 (MPUpBlock).
#include "common.cuh"
#include <stdlib.h>

// ----------------------------------------------------------------------------------------------
// Weight packing.  Logical operand per tap: Bmat[k][n] (k = contraction channel, n = output channel)
//   value = w[n*sn + k*sk + tapidx*st],  tapidx = flip ? T-1-tap : tap
// Packed order: [n_tile(32)][tap][k_step(8)][n8 tile(4)][lane(32)][2]  with
//   lane = g*4+t :  elem0 = Bmat[ks*8+t][nt*8+g], elem1 = Bmat[ks*8+t+4][nt*8+g]   (mma m16n8k8 B fragment)
// ----------------------------------------------------------------------------------------------
struct PackEntry {
  const float* w;
  float* out;
  int N, K, T;        // N padded up to a multiple of 32 in the packed buffer (zeros)
  int sn, sk, st;
  int flip;
  int first;          // prefix offset (in packed elements) of this entry
  int fmt;            // 0: mma.sync fragment order (below);
                      // 2: tcgen05 K-major 128-byte-swizzled rows [n_tile][tap][n 32][chunk ^ (n & 7)][4] (K = 32)
                      // 3: the same for 1x1 / Linear weights of any K: [slab k/32][n][chunk ^ (n & 7)][4]
  int pad_;
};

// The packed buffer holds two planes: [0,total) the TF32-rounded weights ("hi"), [total, 2*total) the TF32-rounded
// residuals w - hi ("lo") used by the error-compensated 3xTF32 mode.
__global__ void pack_weights_kernel(const PackEntry* __restrict__ tab, int n_entries, int total) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int lo = 0, hi = n_entries - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (tab[mid].first <= idx) lo = mid; else hi = mid - 1;
    }
    const PackEntry e = tab[lo];
    int r = idx - e.first;
    if (e.fmt == 3) {     // tcgen05 K-major SWIZZLE_128B rows, any K % 32 == 0: [slab k/32][n][chunk ^ (n & 7)][4]   (csrc/gemm_tma.cu)
      const int npad = (e.N + 31) & ~31;
      const int el = r & 3, cpos = (r >> 2) & 7;
      r >>= 5;
      const int n = r % npad, slab = r / npad;
      const int k = slab * 32 + ((cpos ^ (n & 7)) << 2) + el;
      float v = 0.f;
      if (n < e.N) v = e.w[(size_t)n * e.sn + (size_t)k * e.sk];
      const float vhi = __uint_as_float(f2tf32(v));
      e.out[idx - e.first] = vhi;
      e.out[idx - e.first + total] = __uint_as_float(f2tf32(v - vhi));
      continue;
    }
    if (e.fmt == 2) {     // tcgen05 K-major SWIZZLE_128B rows (K = 32): [n_tile][tap][n 32][chunk ^ (n & 7)][4]
      const int el = r & 3, cpos = (r >> 2) & 7, n32 = (r >> 5) & 31;
      r >>= 10;
      const int tap = r % e.T, cot = r / e.T;
      const int n = cot * 32 + n32, k = ((cpos ^ (n32 & 7)) << 2) + el;
      float v = 0.f;
      if (n < e.N) v = e.w[(size_t)n * e.sn + (size_t)k * e.sk + (size_t)(e.flip ? e.T - 1 - tap : tap) * e.st];
      const float vhi = __uint_as_float(f2tf32(v));
      e.out[idx - e.first] = vhi;
      e.out[idx - e.first + total] = __uint_as_float(f2tf32(v - vhi));
      continue;
    }
    const int j = r & 1; r >>= 1;
    const int lane = r & 31; r >>= 5;
    const int nt = r & 3; r >>= 2;
    const int KS = e.K >> 3;
    const int ks = r % KS; r /= KS;
    const int tap = r % e.T;
    const int cot = r / e.T;
    const int g = lane >> 2, t = lane & 3;
    const int n = cot * 32 + nt * 8 + g;
    const int k = ks * 8 + t + 4 * j;
    float v = 0.f;
    if (n < e.N) {
      const int tapidx = e.flip ? e.T - 1 - tap : tap;
      v = e.w[(size_t)n * e.sn + (size_t)k * e.sk + (size_t)tapidx * e.st];
    }
    const float vhi = __uint_as_float(f2tf32(v));
    e.out[idx - e.first] = vhi;
    e.out[idx - e.first + total] = __uint_as_float(f2tf32(v - vhi));
  }
}

extern "C" int tcct_pack_weights(const void* table_dev, int n_entries, int total_elems, void* stream) {
  if (n_entries == 0) return TCCT_OK;
  int blocks = ceil_div(total_elems, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const PackEntry*)table_dev, n_entries, total_elems);
  TCCT_CHECK_LAUNCH("pack_weights");
  return TCCT_OK;
}
extern "C" int tcct_pack_entry_size() { return (int)sizeof(PackEntry); }

// ----------------------------------------------------------------------------------------------
// Shared epilogue for conv_tile / gemm_px: a warp owns MT m16 tiles x 4 n8 tiles (32 output channels).
// ----------------------------------------------------------------------------------------------
struct Epilogue {
  const float* bias;        // [Cout] or null
  const float* res;         // [pixels, Cout] or null : out = res + res_scale[b] * (acc + bias)
  const float* res_scale;   // [B] or null (1)
  double* stats;            // [2*Cout] or null: sum, sum of squares of stats_act(out)
  int stats_act;
};

// ----------------------------------------------------------------------------------------------
// Spatial conv: 16x16 output tile per CTA iteration, input halo tile staged once in shared memory.
// 4 warps; warp w owns tile rows 4w..4w+3 (4 m16 tiles of 16 pixels) x 32 output channels.
// ----------------------------------------------------------------------------------------------
struct ConvArgs {
  const float* x;
  const float* wpk;
  const float* wpk_lo;     // residual plane (3xTF32 mode) or null
  float* y;
  Epilogue ep;
  int B, H, W, Cout, KH, KW;
  int tiles_x, tiles_y, n_tiles;
  int xs;                  // channels per pixel of the x tensor (>= CIN: the kernel reads the CIN channels x points at)
};

// X3: error-compensated 3xTF32 (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi): fp32-faithful products on the tensor cores.
// MT: m16 tiles (image rows) per warp -> the CTA tile is 16 x 4*MT pixels; small maps use MT < 4 to get more CTAs.
// WS: the CTA's weight fragments (all taps of its 32 output channels) are staged in shared memory together with the
// halo tile.  On small maps a CTA runs one short tile, and fetching the fragments of step s+1 from L2 during step s
// (a few dozen cycles of MMAs) exposes one L2 round trip per step: 36-52 dependent round trips per launch.
template <int CIN, bool X3, int MT, bool WS>
__global__ void __launch_bounds__(128) conv_tile_kernel(const ConvArgs a) {
  constexpr int S = CIN + 4;        // padded pixel stride (floats): ldmatrix rows hit distinct banks
  constexpr int KS = CIN / 8;
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_stats[64];
  constexpr int TH = 4 * MT;
  const int TWin = 16 + a.KW - 1, THin = TH + a.KH - 1;
  const int padH = a.KH >> 1, padW = a.KW >> 1;
  const int T = a.KH * a.KW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int cot = blockIdx.y, co0 = cot * 32;
  if (tid < 64) s_stats[tid] = 0.f;
  float st_sum[8], st_sq[8];
#pragma unroll
  for (int i = 0; i < 8; i++) st_sum[i] = st_sq[i] = 0.f;

  // ldmatrix lane addressing for the A operand (pixels x channels)
  const int lm = lane >> 3, lr = lane & 7;
  const int a_xoff = lr + 8 * (lm & 1), a_koff = 4 * (lm >> 1);
  const uint32_t halo_s = smem_u32(smem);
  const float2* wbase = reinterpret_cast<const float2*>(a.wpk) + (size_t)cot * T * KS * 4 * 32 + lane;
  const float2* wbase_lo = X3 ? reinterpret_cast<const float2*>(a.wpk_lo) + (size_t)cot * T * KS * 4 * 32 + lane : nullptr;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  // weight stage (WS): after the halo tile, 16-byte aligned
  const int halo_floats = ((THin * TWin * S + 3) / 4) * 4;
  const float2* wsm = reinterpret_cast<const float2*>(smem + halo_floats) + lane;
  if (WS) {
    const int chunks = T * KS * 64;        // 16-byte chunks: T*KS*4*32 float2
    const char* src = reinterpret_cast<const char*>(reinterpret_cast<const float2*>(a.wpk) + (size_t)cot * T * KS * 4 * 32);
    const uint32_t dst = halo_s + (uint32_t)halo_floats * 4u;
    for (int idx = tid; idx < chunks; idx += 128) cp_async16(dst + idx * 16u, src + (size_t)idx * 16, 16);
    // committed and awaited together with the first halo tile below
  }

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int trem = tile - b * tiles_per_img;
    const int y0 = (trem / a.tiles_x) * TH, x0 = (trem % a.tiles_x) * 16;
    __syncthreads();                       // previous iteration finished reading the halo tile
    {
      constexpr int CH = CIN / 4;
      const int total = THin * TWin * CH;
      for (int idx = tid; idx < total; idx += 128) {
        const int p = idx / CH, c = idx - p * CH;
        const int hy = p / TWin, hx = p - hy * TWin;
        const int gy = y0 + hy - padH, gx = x0 + hx - padW;
        const bool ok = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
        const float* src = ok ? a.x + (((size_t)b * a.H + gy) * a.W + gx) * a.xs + c * 4 : a.x;
        cp_async16(halo_s + (uint32_t)(p * S + c * 4) * 4u, src, ok ? 16 : 0);
      }
      cp_async_commit();
      cp_async_wait<0>();
    }
    __syncthreads();

    float acc[MT][4][4];
#pragma unroll
    for (int i = 0; i < MT; i++)
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[i][j][k] = 0.f;

    uint32_t a_base[MT];
#pragma unroll
    for (int mt = 0; mt < MT; mt++)
      a_base[mt] = halo_s + (uint32_t)(((warp * MT + mt) * TWin + a_xoff) * S + a_koff) * 4u;

    // weight fragments: from the shared-memory stage (WS), else straight from L2 with the loads of step s+1 issued
    // before the MMAs of step s
    const float2* wp = WS ? wsm : wbase;
    const float2* wpl = wbase_lo;
    float2 nbf[4], nbl[4];
    if (!WS) {
#pragma unroll
      for (int nt = 0; nt < 4; nt++) {
        nbf[nt] = __ldg(wp + nt * 32);
        if (X3) nbl[nt] = __ldg(wpl + nt * 32);
      }
    }
    for (int tap = 0; tap < T; tap++) {
      const int dy = tap / a.KW, dx = tap - dy * a.KW;
      const uint32_t tap_off = (uint32_t)((dy * TWin + dx) * S) * 4u;
#pragma unroll
      for (int ks = 0; ks < KS; ks++) {
        float2 bf[4], bl[4];
        if (WS) {
#pragma unroll
          for (int nt = 0; nt < 4; nt++) bf[nt] = wp[nt * 32];
          wp += 4 * 32;
        } else {
#pragma unroll
          for (int nt = 0; nt < 4; nt++) { bf[nt] = nbf[nt]; if (X3) bl[nt] = nbl[nt]; }
          wp += 4 * 32;
          if (X3) wpl += 4 * 32;
          if (tap * KS + ks + 1 < T * KS) {
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
              nbf[nt] = __ldg(wp + nt * 32);
              if (X3) nbl[nt] = __ldg(wpl + nt * 32);
            }
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
          uint32_t af[4], al[4];
          ldmatrix_x4(af, a_base[mt] + tap_off + ks * 32);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float v = __uint_as_float(af[i]);
            af[i] = f2tf32(v);
            if (X3) al[i] = f2tf32(v - __uint_as_float(af[i]));
          }
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            if (X3) {
              mma_tf32(acc[mt][nt], al, __float_as_uint(bf[nt].x), __float_as_uint(bf[nt].y));
              mma_tf32(acc[mt][nt], af, __float_as_uint(bl[nt].x), __float_as_uint(bl[nt].y));
            }
            mma_tf32(acc[mt][nt], af, __float_as_uint(bf[nt].x), __float_as_uint(bf[nt].y));
          }
        }
      }
    }

    // epilogue
    const float rs = (a.ep.res != nullptr && a.ep.res_scale != nullptr) ? a.ep.res_scale[b] : 1.f;
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      const int gy = y0 + warp * MT + mt;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int gx = x0 + g + 8 * h;
        if (gy < a.H && gx < a.W) {
          const size_t pix = ((size_t)b * a.H + gy) * a.W + gx;
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const int c = co0 + nt * 8 + 2 * t;
            float v0 = acc[mt][nt][2 * h], v1 = acc[mt][nt][2 * h + 1];
            if (a.ep.bias) { v0 += a.ep.bias[c]; v1 += a.ep.bias[c + 1]; }
            if (a.ep.res) {
              const float2 r = *reinterpret_cast<const float2*>(a.ep.res + pix * a.Cout + c);
              v0 = r.x + rs * v0; v1 = r.y + rs * v1;
            }
            *reinterpret_cast<float2*>(a.y + pix * a.Cout + c) = make_float2(v0, v1);
            if (a.ep.stats) {
              const float s0 = stat_act(a.ep.stats_act, v0), s1 = stat_act(a.ep.stats_act, v1);
              st_sum[nt * 2] += s0; st_sq[nt * 2] += s0 * s0;
              st_sum[nt * 2 + 1] += s1; st_sq[nt * 2 + 1] += s1 * s1;
            }
          }
        }
      }
    }
  }

  if (a.ep.stats) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      float s = st_sum[i], q = st_sq[i];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (g == 0) {
        const int c = (i >> 1) * 8 + 2 * t + (i & 1);
        atomicAdd(&s_stats[c], s);
        atomicAdd(&s_stats[32 + c], q);
      }
    }
    __syncthreads();
    if (tid < 32) {
      atomicAdd(a.ep.stats + co0 + tid, (double)s_stats[tid]);
      atomicAdd(a.ep.stats + a.Cout + co0 + tid, (double)s_stats[32 + tid]);
    }
  }
}

static int conv_smem_bytes(int cin, int kh, int kw, int th) { return (th + kh - 1) * (16 + kw - 1) * (cin + 4) * 4; }

template <int CIN, bool X3, int MT, bool WS>
static void launch_conv_mt(const ConvArgs& a, dim3 grid, int smem, cudaStream_t st) {
  cudaFuncSetAttribute(conv_tile_kernel<CIN, X3, MT, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  conv_tile_kernel<CIN, X3, MT, WS><<<grid, 128, smem, st>>>(a);
}
template <int CIN, bool X3>
static void launch_conv(const ConvArgs& a, int mt, bool ws, dim3 grid, int smem, cudaStream_t st) {
  if (!X3 && ws) {
    if (mt == 4) launch_conv_mt<CIN, false, 4, true>(a, grid, smem, st);
    else if (mt == 2) launch_conv_mt<CIN, false, 2, true>(a, grid, smem, st);
    else launch_conv_mt<CIN, false, 1, true>(a, grid, smem, st);
    return;
  }
  if (mt == 4) launch_conv_mt<CIN, X3, 4, false>(a, grid, smem, st);
  else if (mt == 2) launch_conv_mt<CIN, X3, 2, false>(a, grid, smem, st);
  else launch_conv_mt<CIN, X3, 1, false>(a, grid, smem, st);
}

// lo_off: 0 = plain TF32; otherwise the element offset from wpk to the residual plane (3xTF32 mode).
static int conv2d_nhwc_launch(const float* x, int x_ch, const float* wpk, long long lo_off, const float* bias, float* y, int B,
                              int H, int W, int Cin, int Cout, int KH, int KW, const float* res,
                              const float* res_scale, double* stats, int stats_act, void* stream) {
  TCCT_CHECK_ARG(Cin == 32 || Cin == 64, "conv2d_nhwc: Cin must be 32 or 64 (got %d)", Cin);
  TCCT_CHECK_ARG(x_ch >= Cin && x_ch % 4 == 0, "conv2d_nhwc: x carries %d channels per pixel, the slice needs %d", x_ch, Cin);
  TCCT_CHECK_ARG(Cout % 32 == 0 && Cout > 0, "conv2d_nhwc: Cout must be a multiple of 32 (got %d)", Cout);
  TCCT_CHECK_ARG((KH & 1) && (KW & 1) && KH * KW <= 25, "conv2d_nhwc: odd kernel with <= 25 taps expected (%dx%d)", KH, KW);
  TCCT_CHECK_ARG(B > 0 && H > 0 && W > 0, "conv2d_nhwc: empty input");
  ConvArgs a;
  a.x = x; a.wpk = wpk; a.wpk_lo = lo_off ? wpk + lo_off : nullptr; a.y = y;
  a.ep.bias = bias; a.ep.res = res; a.ep.res_scale = res_scale; a.ep.stats = stats; a.ep.stats_act = stats_act;
  a.B = B; a.H = H; a.W = W; a.Cout = Cout; a.KH = KH; a.KW = KW; a.xs = x_ch;
  // rows per warp: shrink the CTA tile on small maps until there is about one CTA per SM
  int mt = 4;
  while (mt > 1 && (long long)B * ceil_div(W, 16) * ceil_div(H, 4 * mt) * (Cout / 32) < tcct_num_sms()) mt >>= 1;
  a.tiles_x = ceil_div(W, 16); a.tiles_y = ceil_div(H, 4 * mt); a.n_tiles = B * a.tiles_x * a.tiles_y;
  int smem = conv_smem_bytes(Cin, KH, KW, 4 * mt);
  // few tiles per CTA: stage the weight fragments in shared memory (see conv_tile_kernel)
  const int wbytes = KH * KW * (Cin / 8) * 1024;
  const bool ws = !lo_off && (long long)a.n_tiles <= 4ll * tcct_num_sms() && ((smem + 15) / 16 * 16) + wbytes <= 200 * 1024;
  if (ws) smem = (smem + 15) / 16 * 16 + wbytes;
  int occ = 232448 / (smem + 1280);
  if (occ > 4) occ = 4;
  if (occ < 1) { tcct_set_error("conv2d_nhwc: tile does not fit in shared memory (%d B)", smem); return TCCT_ERR_ARG; }
  int gx = tcct_num_sms() * occ;
  if (gx > a.n_tiles) gx = a.n_tiles;
  dim3 grid(gx, Cout / 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cin == 32) { if (lo_off) launch_conv<32, true>(a, mt, false, grid, smem, st); else launch_conv<32, false>(a, mt, ws, grid, smem, st); }
  else { if (lo_off) launch_conv<64, true>(a, mt, false, grid, smem, st); else launch_conv<64, false>(a, mt, ws, grid, smem, st); }
  TCCT_CHECK_LAUNCH("conv2d_nhwc");
  return TCCT_OK;
}

extern "C" int tcct_conv2d_nhwc(const float* x, const float* wpk, long long lo_off, const float* bias, float* y, int B,
                                int H, int W, int Cin, int Cout, int KH, int KW, const float* res,
                                const float* res_scale, double* stats, int stats_act, void* stream) {
  return conv2d_nhwc_launch(x, Cin, wpk, lo_off, bias, y, B, H, W, Cin, Cout, KH, KW, res, res_scale, stats, stats_act, stream);
}
// One 32- or 64-channel slice of the reduction of a wider conv (the 64..256-channel CrossResNet of stc_tb / gtc_tb, tcct.py:861-864):
// x points at the first channel of the slice inside a [B,H,W,x_ch] tensor, wpk is the pack of that slice of the weight; the caller
// chains the slices through res = y (in place: each thread reads and writes its own elements), bias on the first, stats on the last.
extern "C" int tcct_conv2d_nhwc_slice(const float* x, int x_ch, const float* wpk, long long lo_off, const float* bias, float* y, int B,
                                      int H, int W, int Cin, int Cout, int KH, int KW, const float* res, double* stats,
                                      int stats_act, void* stream) {
  return conv2d_nhwc_launch(x, x_ch, wpk, lo_off, bias, y, B, H, W, Cin, Cout, KH, KW, res, nullptr, stats, stats_act, stream);
}

// ----------------------------------------------------------------------------------------------
// GEMM over pixels: y[M][N] = x[M][K] . Bpk   (1x1 conv / Linear), 128 px x 32 co per CTA,
// K streamed in 32-channel slabs through a 3-stage cp.async ring.
// ----------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* x;
  const float* wpk;
  const float* wpk_lo;
  float* y;
  Epilogue ep;
  int M, K, N;            // N = Cout (row stride of y and res)
  int px_per_sample;
};

template <bool X3>
__global__ void __launch_bounds__(128) gemm_px_kernel(const GemmArgs a) {
  constexpr int S = 36, STAGES = 3;
  extern __shared__ __align__(16) float sA_raw[];          // [STAGES][128 * S]
  float (*sA)[128 * S] = reinterpret_cast<float (*)[128 * S]>(sA_raw);
  __shared__ float s_stats[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * 128, cot = blockIdx.y, co0 = cot * 32;
  const int nslab = a.K >> 5;
  if (tid < 64) s_stats[tid] = 0.f;

  auto load_slab = [&](int slab, int stage) {
    const uint32_t sbase = smem_u32(&sA[stage][0]);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int idx = tid + i * 128;           // 1024 chunks of 16 B
      const int p = idx >> 3, c = idx & 7;
      const int m = m0 + p;
      const bool ok = m < a.M;
      const float* src = ok ? a.x + (size_t)m * a.K + slab * 32 + c * 4 : a.x;
      cp_async16(sbase + (uint32_t)(p * S + c * 4) * 4u, src, ok ? 16 : 0);
    }
  };
  // plain TF32: the CTA's weight fragments (K/8 steps x 4 n-tiles x 32 lanes of float2) are staged in shared memory behind the
  // A stages, in the first cp.async group; fetching step s+1 from L2 during step s exposed one L2 round trip per k-step
  const float2* wsm = reinterpret_cast<const float2*>(sA_raw + STAGES * 128 * S) + lane;
  if (!X3) {
    const int chunks = (a.K >> 3) * 64;          // 16-byte chunks
    const char* src = reinterpret_cast<const char*>(reinterpret_cast<const float2*>(a.wpk) + (size_t)cot * (a.K >> 3) * 4 * 32);
    const uint32_t dst = smem_u32(sA_raw + STAGES * 128 * S);
    for (int idx = tid; idx < chunks; idx += 128) cp_async16(dst + idx * 16u, src + (size_t)idx * 16, 16);
  }
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < nslab) load_slab(s, s);
    cp_async_commit();
  }

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int k = 0; k < 4; k++) acc[i][j][k] = 0.f;

  const int lm = lane >> 3, lr = lane & 7;
  const int a_row = warp * 32 + lr + 8 * (lm & 1), a_koff = 4 * (lm >> 1);
  const float2* wp = X3 ? reinterpret_cast<const float2*>(a.wpk) + (size_t)cot * (a.K >> 3) * 4 * 32 + lane : wsm;
  const float2* wpl = X3 ? reinterpret_cast<const float2*>(a.wpk_lo) + (size_t)cot * (a.K >> 3) * 4 * 32 + lane : nullptr;

  // 3xTF32 mode: weight fragments come straight from L2, the loads of step s+1 are issued before the MMAs of step s
  float2 nbf[4], nbl[4];
  if (X3) {
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      nbf[nt] = __ldg(wp + nt * 32);
      nbl[nt] = __ldg(wpl + nt * 32);
    }
  }
  for (int slab = 0; slab < nslab; slab++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {   // prefetch slab + STAGES-1 into the stage that was consumed in the previous iteration
      const int nxt = slab + STAGES - 1;
      if (nxt < nslab) load_slab(nxt, nxt % STAGES);
      cp_async_commit();
    }
    const uint32_t sbase = smem_u32(&sA[slab % STAGES][0]);
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      float2 bf[4], bl[4];
      if (X3) {
#pragma unroll
        for (int nt = 0; nt < 4; nt++) { bf[nt] = nbf[nt]; bl[nt] = nbl[nt]; }
        wp += 4 * 32;
        wpl += 4 * 32;
        if (slab * 4 + ks + 1 < nslab * 4) {
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            nbf[nt] = __ldg(wp + nt * 32);
            nbl[nt] = __ldg(wpl + nt * 32);
          }
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < 4; nt++) bf[nt] = wp[nt * 32];
        wp += 4 * 32;
      }
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        uint32_t af[4], al[4];
        ldmatrix_x4(af, sbase + (uint32_t)((a_row + mt * 16) * S + a_koff + ks * 8) * 4u);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float v = __uint_as_float(af[i]);
          af[i] = f2tf32(v);
          if (X3) al[i] = f2tf32(v - __uint_as_float(af[i]));
        }
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          if (X3) {
            mma_tf32(acc[mt][nt], al, __float_as_uint(bf[nt].x), __float_as_uint(bf[nt].y));
            mma_tf32(acc[mt][nt], af, __float_as_uint(bl[nt].x), __float_as_uint(bl[nt].y));
          }
          mma_tf32(acc[mt][nt], af, __float_as_uint(bf[nt].x), __float_as_uint(bf[nt].y));
        }
      }
    }
  }
  cp_async_wait<0>();

  float st_sum[8], st_sq[8];
#pragma unroll
  for (int i = 0; i < 8; i++) st_sum[i] = st_sq[i] = 0.f;
#pragma unroll
  for (int mt = 0; mt < 2; mt++) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int m = m0 + warp * 32 + mt * 16 + g + 8 * h;
      if (m < a.M) {
        float rs = 1.f;
        if (a.ep.res && a.ep.res_scale) rs = a.ep.res_scale[m / a.px_per_sample];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) {
          const int c = co0 + nt * 8 + 2 * t;
          if (c < a.N) {                       // N may be a partial 32-tile only for packed-zero padding
            float v0 = acc[mt][nt][2 * h], v1 = acc[mt][nt][2 * h + 1];
            if (a.ep.bias) { v0 += a.ep.bias[c]; v1 += a.ep.bias[c + 1]; }
            if (a.ep.res) {
              const float2 r = *reinterpret_cast<const float2*>(a.ep.res + (size_t)m * a.N + c);
              v0 = r.x + rs * v0; v1 = r.y + rs * v1;
            }
            *reinterpret_cast<float2*>(a.y + (size_t)m * a.N + c) = make_float2(v0, v1);
            if (a.ep.stats) {
              const float s0 = stat_act(a.ep.stats_act, v0), s1 = stat_act(a.ep.stats_act, v1);
              st_sum[nt * 2] += s0; st_sq[nt * 2] += s0 * s0;
              st_sum[nt * 2 + 1] += s1; st_sq[nt * 2 + 1] += s1 * s1;
            }
          }
        }
      }
    }
  }
  if (a.ep.stats) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      float s = st_sum[i], q = st_sq[i];
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
      }
      if (g == 0) {
        const int c = (i >> 1) * 8 + 2 * t + (i & 1);
        atomicAdd(&s_stats[c], s);
        atomicAdd(&s_stats[32 + c], q);
      }
    }
    __syncthreads();
    if (tid < 32 && co0 + tid < a.N) {
      atomicAdd(a.ep.stats + co0 + tid, (double)s_stats[tid]);
      atomicAdd(a.ep.stats + a.N + co0 + tid, (double)s_stats[32 + tid]);
    }
  }
}

extern "C" int tcct_gemm_px(const float* x, const float* wpk, long long lo_off, const float* bias, float* y, long long M, int K, int N,
                            const float* res, const float* res_scale, int px_per_sample, double* stats,
                            int stats_act, void* stream) {
  TCCT_CHECK_ARG(K % 32 == 0 && K > 0, "gemm_px: K must be a multiple of 32 (got %d)", K);
  TCCT_CHECK_ARG(N % 32 == 0 && N > 0, "gemm_px: N must be a multiple of 32 (got %d)", N);
  TCCT_CHECK_ARG(M > 0 && M < (1ll << 31), "gemm_px: bad M");
  GemmArgs a;
  a.x = x; a.wpk = wpk; a.wpk_lo = lo_off ? wpk + lo_off : nullptr; a.y = y;
  a.ep.bias = bias; a.ep.res = res; a.ep.res_scale = res_scale; a.ep.stats = stats; a.ep.stats_act = stats_act;
  a.M = (int)M; a.K = K; a.N = N; a.px_per_sample = px_per_sample > 0 ? px_per_sample : (int)M;
  dim3 grid(ceil_div(M, 128), N / 32);
  int smem = 3 * 128 * 36 * 4;
  if (!lo_off) smem += K * 128;          // weight-fragment stage: K/8 steps x 4 n-tiles x 32 lanes x 8 B
  TCCT_CHECK_ARG(smem <= 220 * 1024, "gemm_px: K = %d does not fit the shared-memory weight stage", K);
  if (lo_off) {
    cudaFuncSetAttribute(gemm_px_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    gemm_px_kernel<true><<<grid, 128, smem, (cudaStream_t)stream>>>(a);
  } else {
    cudaFuncSetAttribute(gemm_px_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    gemm_px_kernel<false><<<grid, 128, smem, (cudaStream_t)stream>>>(a);
  }
  TCCT_CHECK_LAUNCH("gemm_px");
  return TCCT_OK;
}

// ----------------------------------------------------------------------------------------------
// Weight gradient: dW[co][ci][tap] += sum_px dy[px][co] * x[px + tap][ci]   (+ dbias[co] += sum_px dy)
// "taps" are spatial offsets (spatial mode, Cin = 32) or 32-channel slabs of a wide x row (linear mode).
// Persistent CTAs (8 warps) loop over pixel tiles; work items (tap, k-part) are dealt to warps.
// ----------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* x;
  const float* dy;
  float* dw;
  float* dbias;
  int B, H, W;            // spatial mode: image dims; linear mode: H = 1, W = pixels, B = 1
  int xC, dyC;            // channels per pixel in x and dy
  int KH, KW, T;          // T = taps (spatial) or xC/32 (linear)
  int spatial;
  int TP;                 // pixels per tile (256 spatial; 64/128/256 linear)
  int n_tiles, tiles_x, tiles_y;
  int sco, sci, stp;      // dw strides
  int ksplit;
  int tz;                 // taps per blockIdx.z slice (small problems spread their taps over more CTAs)
};

template <bool X3>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int co0 = blockIdx.y * 32;
  const int SX = a.spatial ? 40 : a.xC + 8;   // == 8 (mod 32): conflict-free scalar fragment loads
  constexpr int SD = 40;
  const int TWin = a.spatial ? 16 + a.KW - 1 : 0, THin = a.spatial ? 16 + a.KH - 1 : 0;
  const int HP = a.spatial ? TWin * THin : a.TP;
  float* xs = smem;
  float* ds = smem + (size_t)HP * SX;
  const uint32_t xs_s = smem_u32(xs), ds_s = smem_u32(ds);
  const int padH = a.KH >> 1, padW = a.KW >> 1;

  // work items of this warp: (tap, k-part) pairs of this CTA's tap slice [tap0, tap0 + TL)
  const int tap0 = blockIdx.z * a.tz;
  const int TL = min(a.tz, a.T - tap0);
  const int n_items = TL * a.ksplit;
  int item[2] = {warp, warp + 8};
  int tap_i[2], kp_i[2], toff[2];
  bool have[2];
#pragma unroll
  for (int i = 0; i < 2; i++) {
    have[i] = item[i] < n_items;
    tap_i[i] = have[i] ? item[i] / a.ksplit : 0;
    kp_i[i] = have[i] ? item[i] % a.ksplit : 0;
    const int tg = tap0 + tap_i[i];
    toff[i] = a.spatial ? ((tg / a.KW) * TWin + (tg % a.KW)) * SX : tg * 32;
  }
  const int ksteps = a.TP >> 3;
  const int ks_per = ksteps / a.ksplit;

  float acc[2][2][4][4];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int m = 0; m < 2; m++)
#pragma unroll
      for (int n = 0; n < 4; n++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[i][m][n][k] = 0.f;
  float bsum = 0.f;
  const int tiles_per_img = a.tiles_x * a.tiles_y;
  const int npix_lin = a.W;

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    __syncthreads();
    if (a.spatial) {
      const int b = tile / tiles_per_img;
      const int trem = tile - b * tiles_per_img;
      const int y0 = (trem / a.tiles_x) * 16, x0 = (trem % a.tiles_x) * 16;
      for (int idx = tid; idx < HP * 8; idx += 256) {
        const int p = idx >> 3, c = idx & 7;
        const int hy = p / TWin, hx = p - hy * TWin;
        const int gy = y0 + hy - padH, gx = x0 + hx - padW;
        const bool ok = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
        const float* src = ok ? a.x + (((size_t)b * a.H + gy) * a.W + gx) * a.xC + c * 4 : a.x;
        cp_async16(xs_s + (uint32_t)(p * SX + c * 4) * 4u, src, ok ? 16 : 0);
      }
      for (int idx = tid; idx < 256 * 8; idx += 256) {
        const int p = idx >> 3, c = idx & 7;
        const int gy = y0 + (p >> 4), gx = x0 + (p & 15);
        const bool ok = gy < a.H && gx < a.W;
        const float* src = ok ? a.dy + (((size_t)b * a.H + gy) * a.W + gx) * a.dyC + co0 + c * 4 : a.dy;
        cp_async16(ds_s + (uint32_t)(p * SD + c * 4) * 4u, src, ok ? 16 : 0);
      }
    } else {
      const int p0 = tile * a.TP;
      const int CH = a.xC >> 2;
      for (int idx = tid; idx < a.TP * CH; idx += 256) {
        const int p = idx / CH, c = idx - p * CH;
        const bool ok = p0 + p < npix_lin;
        const float* src = ok ? a.x + (size_t)(p0 + p) * a.xC + c * 4 : a.x;
        cp_async16(xs_s + (uint32_t)(p * SX + c * 4) * 4u, src, ok ? 16 : 0);
      }
      for (int idx = tid; idx < a.TP * 8; idx += 256) {
        const int p = idx >> 3, c = idx & 7;
        const bool ok = p0 + p < npix_lin;
        const float* src = ok ? a.dy + (size_t)(p0 + p) * a.dyC + co0 + c * 4 : a.dy;
        cp_async16(ds_s + (uint32_t)(p * SD + c * 4) * 4u, src, ok ? 16 : 0);
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    if (a.dbias && blockIdx.z == 0) {
      const int c = tid & 31;
      for (int q = tid >> 5; q < a.TP; q += 8) bsum += ds[q * SD + c];
    }
    if (have[0]) {
      // both items of a warp share the k-part (ksplit == 1 whenever a warp owns two items)
      const int ks0 = kp_i[0] * ks_per;
      for (int ks = ks0; ks < ks0 + ks_per; ks++) {
        const int q0 = ks * 8 + t, q1 = q0 + 4;
        uint32_t af[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          const float v0 = ds[q0 * SD + mt * 16 + g], v1 = ds[q0 * SD + mt * 16 + g + 8];
          const float v2 = ds[q1 * SD + mt * 16 + g], v3 = ds[q1 * SD + mt * 16 + g + 8];
          af[mt][0] = f2tf32(v0); af[mt][1] = f2tf32(v1); af[mt][2] = f2tf32(v2); af[mt][3] = f2tf32(v3);
          if (X3) {
            al[mt][0] = f2tf32(v0 - __uint_as_float(af[mt][0])); al[mt][1] = f2tf32(v1 - __uint_as_float(af[mt][1]));
            al[mt][2] = f2tf32(v2 - __uint_as_float(af[mt][2])); al[mt][3] = f2tf32(v3 - __uint_as_float(af[mt][3]));
          }
        }
        int xb0, xb1;
        if (a.spatial) {
          xb0 = ((q0 >> 4) * TWin + (q0 & 15)) * SX;
          xb1 = ((q1 >> 4) * TWin + (q1 & 15)) * SX;
        } else {
          xb0 = q0 * SX; xb1 = q1 * SX;
        }
#pragma unroll
        for (int i = 0; i < 2; i++) {
          if (i == 1 && !have[1]) break;
#pragma unroll
          for (int nt = 0; nt < 4; nt++) {
            const float x0 = xs[xb0 + toff[i] + nt * 8 + g], x1 = xs[xb1 + toff[i] + nt * 8 + g];
            const uint32_t b0 = f2tf32(x0), b1 = f2tf32(x1);
            if (X3) {
              const uint32_t l0 = f2tf32(x0 - __uint_as_float(b0)), l1 = f2tf32(x1 - __uint_as_float(b1));
              mma_tf32(acc[i][0][nt], al[0], b0, b1);
              mma_tf32(acc[i][1][nt], al[1], b0, b1);
              mma_tf32(acc[i][0][nt], af[0], l0, l1);
              mma_tf32(acc[i][1][nt], af[1], l0, l1);
            }
            mma_tf32(acc[i][0][nt], af[0], b0, b1);
            mma_tf32(acc[i][1][nt], af[1], b0, b1);
          }
        }
      }
    }
  }

  // reduce the warps' partial results through shared memory (every (warp, item) pair owns a 32x32 slab: no shared-memory
  // float atomics, which are CAS loops), then one global atomic per element per CTA
  __syncthreads();
  float* part = smem;                    // [16 items][1024] (+32 for the bias)
#pragma unroll
  for (int i = 0; i < 2; i++) {
    if (!have[i]) continue;
    float* dst = part + (size_t)item[i] * 1024;
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
      for (int nt = 0; nt < 4; nt++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int co = mt * 16 + g + 8 * (k >> 1), ci = nt * 8 + 2 * t + (k & 1);
          dst[co * 32 + ci] = acc[i][mt][nt][k];
        }
  }
  float* sb = part + 16 * 1024;          // [8 warps][32] bias partials
  if (a.dbias && blockIdx.z == 0) sb[warp * 32 + lane] = bsum;
  __syncthreads();
  for (int i = tid; i < TL * 1024; i += 256) {
    const int tl = i >> 10, co = (i >> 5) & 31, ci = i & 31;
    float v = 0.f;
    for (int kp = 0; kp < a.ksplit; kp++) v += part[(size_t)(tl * a.ksplit + kp) * 1024 + (i & 1023)];
    const int tap = tap0 + tl;
    size_t off;
    if (a.spatial) off = (size_t)(co0 + co) * a.sco + (size_t)ci * a.sci + (size_t)tap * a.stp;
    else off = (size_t)(co0 + co) * a.sco + (size_t)(tap * 32 + ci) * a.sci;
    atomicAdd(a.dw + off, v);
  }
  if (a.dbias && blockIdx.z == 0 && tid < 32) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; w8++) v += sb[w8 * 32 + tid];
    atomicAdd(a.dbias + co0 + tid, v);
  }
}

// dw strides are given in elements: dw[co*sco + ci*sci + tap*stp]
// x3: 1 = error-compensated 3xTF32 products
static int wgrad_launch(const float* x, int x_ch, const float* dy, float* dw, float* dbias, int B, int H, int W, int Cin,
                        int Cout, int KH, int KW, int sco, int sci, int stp, int x3, void* stream) {
  TCCT_CHECK_ARG(Cout % 32 == 0 && Cin % 32 == 0, "wgrad: channels must be multiples of 32 (%d,%d)", Cin, Cout);
  TCCT_CHECK_ARG(x_ch >= Cin && x_ch % 4 == 0 && (x_ch == Cin || KH * KW > 1), "wgrad: x channel stride %d does not fit Cin %d", x_ch, Cin);
  WgradArgs a;
  a.x = x; a.dy = dy; a.dw = dw; a.dbias = dbias;
  a.xC = x_ch; a.dyC = Cout; a.KH = KH; a.KW = KW;
  a.sco = sco; a.sci = sci; a.stp = stp;
  a.spatial = (KH * KW > 1) ? 1 : 0;
  size_t smem;
  if (a.spatial) {
    TCCT_CHECK_ARG(Cin == 32, "wgrad: spatial mode needs Cin == 32 (got %d)", Cin);
    TCCT_CHECK_ARG((KH & 1) && (KW & 1) && KH * KW <= 16, "wgrad: unsupported kernel %dx%d", KH, KW);
    a.B = B; a.H = H; a.W = W; a.T = KH * KW; a.TP = 256;
    a.tiles_x = ceil_div(W, 16); a.tiles_y = ceil_div(H, 16); a.n_tiles = B * a.tiles_x * a.tiles_y;
    smem = ((size_t)(16 + KH - 1) * (16 + KW - 1) * 40 + 256 * 40) * 4;
  } else {
    TCCT_CHECK_ARG(Cin <= 512, "wgrad: Cin too large (%d)", Cin);
    const long long npix = (long long)B * H * W;
    a.B = 1; a.H = 1; a.W = (int)npix; a.T = Cin / 32;
    a.TP = Cin <= 64 ? 256 : (Cin <= 160 ? 128 : 64);
    a.tiles_x = a.tiles_y = 1; a.n_tiles = ceil_div(npix, a.TP);
    smem = ((size_t)a.TP * (Cin + 8) + (size_t)a.TP * 40) * 4;
  }
  TCCT_CHECK_ARG(a.T <= 16, "wgrad: too many taps/slabs (%d)", a.T);
  // few pixel tiles (small maps): give every CTA a slice of the taps so that about one CTA per SM is in flight
  int zs = 1;
  {
    const long long ctas = (long long)a.n_tiles * (Cout / 32);
    if (ctas < tcct_num_sms()) zs = (int)((tcct_num_sms() + ctas - 1) / ctas);
    if (zs > a.T) zs = a.T;
  }
  a.tz = ceil_div(a.T, zs);
  zs = ceil_div(a.T, a.tz);
  a.ksplit = a.tz >= 5 ? 1 : (a.tz >= 3 ? 2 : (a.tz == 2 ? 4 : 8));
  const size_t red = ((size_t)16 * 1024 + 8 * 32) * 4;      // the warps' partial slabs after the main loop
  if (smem < red) smem = red;
  int occ = (int)(232448 / (smem + 1024));
  if (occ > 2) occ = 2;
  TCCT_CHECK_ARG(occ >= 1, "wgrad: tile does not fit in shared memory");
  int gx = tcct_num_sms() * occ;
  if (gx > a.n_tiles) gx = a.n_tiles;
  dim3 grid(gx, Cout / 32, zs);
  if (x3) {
    cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wgrad_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  } else {
    cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wgrad_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(a);
  }
  TCCT_CHECK_LAUNCH("wgrad");
  return TCCT_OK;
}
extern "C" int tcct_wgrad(const float* x, const float* dy, float* dw, float* dbias, int B, int H, int W, int Cin,
                          int Cout, int KH, int KW, int sco, int sci, int stp, int x3, void* stream) {
  return wgrad_launch(x, Cin, dy, dw, dbias, B, H, W, Cin, Cout, KH, KW, sco, sci, stp, x3, stream);
}
// Spatial weight gradient of one 32-channel input slice of a wider conv: x points at the slice inside a [B,H,W,x_ch] tensor, dw at
// dW[0][slice][0] (strides as for tcct_wgrad); dbias on one slice only.
extern "C" int tcct_wgrad_slice(const float* x, int x_ch, const float* dy, float* dw, float* dbias, int B, int H, int W, int Cout,
                                int KH, int KW, int sco, int sci, int stp, int x3, void* stream) {
  TCCT_CHECK_ARG(KH * KW > 1, "wgrad_slice: spatial kernels only");
  return wgrad_launch(x, x_ch, dy, dw, dbias, B, H, W, 32, Cout, KH, KW, sco, sci, stp, x3, stream);
}
