// Feature polarisation, RegNet.regular_udh (task1/nets/reg.py:86-105) with FeatConSuper.select1 /
// points_selection_bins (task1/nets/fcs.py:25-50,82-96), cosinesim/foreach_loss (fcs.py:63-80) and
// FeatConPolar.choice (task1/nets/fcp.py:72-75):
//   key[px]  = softmax_C(logits)[label[px]]            (probability of the pixel's own class)
//   per class i: pixels sorted by key descending, n_i = count_i // 32, bin b = ranks [b n_i, (b+1) n_i),
//                pro_i[b] = mean of the 32-d feature rows of the bin (the count_i % 32 lowest-ranked rows are dropped)
//   loss = sum_i -( sum_b pro_i[b] . proto_i ) / (32*32)  +  mean( (pro_{C-1} - proto_{C-1})^2 )
// A class with fewer than 32 pixels gives NaN like the reference (fcs.py:36, mean of an empty bin).
// The sort is a stable LSD radix sort over the 34 significant bits of (class, ~bits(p)): p <= 1 leaves bits 30-31 of the
// key constant, so 4 passes of 9 + 9 + 8 + (4 key bits | 4 class bits) cover it; one warp per 512-element chunk
// (1024 warps at 8 x 256 x 256); the key kernel builds the first histogram.  Features are NHWC [B,H,W,32] so a pixel's
// row is one 128-byte line.
#include "common.cuh"

#define FP_CHUNK 512
#define FP_BINS 512
#define FP_PASSES 4
#define FP_MAXC 16

// ---- keys (+ histogram of the first radix pass) --------------------------------------------------------
__global__ void __launch_bounds__(128) fp_keys_kernel(const float* __restrict__ logits, const unsigned char* __restrict__ lab, int B, int C,
                                                      int HW, unsigned int* __restrict__ key, unsigned int* __restrict__ idx,
                                                      unsigned int* cls_count, int units, unsigned int* __restrict__ hist) {
  __shared__ unsigned int scnt[FP_MAXC];
  __shared__ unsigned int sh[4][FP_BINS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < FP_MAXC) scnt[threadIdx.x] = 0;
  for (int i = lane; i < FP_BINS; i += 32) sh[warp][i] = 0;
  __syncthreads();
  const long long n = (long long)B * HW;
  const int unit = blockIdx.x * 4 + warp;
  if (unit < units) {
    const long long j0 = (long long)unit * FP_CHUNK;
#pragma unroll 4
    for (int s = 0; s < FP_CHUNK; s += 32) {
      const long long i = j0 + s + lane;
      if (i < n) {
        const int b = (int)((unsigned int)i / (unsigned int)HW);      // n < 2^31 (tcct_fpolar_forward)
        const int q = (int)((unsigned int)i - (unsigned int)b * (unsigned int)HW);
        const float* lp = logits + ((size_t)b * C) * HW + q;
        const int l = lab[i];
        float m = -INFINITY;
        for (int c = 0; c < C; c++) m = fmaxf(m, lp[(size_t)c * HW]);
        float sum = 0.f, mine = 0.f;
        for (int c = 0; c < C; c++) {
          const float e = expf(lp[(size_t)c * HW] - m);
          sum += e;
          if (c == l) mine = e;
        }
        const float p = mine / sum;
        const unsigned int k = ~__float_as_uint(p);          // p >= 0: ascending order of ~bits == descending probability
        key[i] = k;
        idx[i] = (unsigned int)i;
        atomicAdd(&scnt[l < FP_MAXC ? l : FP_MAXC - 1], 1u);
        atomicAdd(&sh[warp][k & 511u], 1u);
      }
    }
    __syncwarp();
    for (int i = lane; i < FP_BINS; i += 32) hist[(size_t)i * units + unit] = sh[warp][i];
  }
  __syncthreads();
  if (threadIdx.x < FP_MAXC && scnt[threadIdx.x]) atomicAdd(cls_count + threadIdx.x, scnt[threadIdx.x]);
}

// ---- radix sort passes -------------------------------------------------------------------------------
// digit of element j in this pass: bits [0,9), [9,18), [18,26) of key[j], then bits [26,30) | class << 4
__device__ __forceinline__ unsigned int fp_digit(int pass, unsigned int k, unsigned int id, const unsigned char* lab) {
  if (pass == 0) return k & 511u;
  if (pass == 1) return (k >> 9) & 511u;
  if (pass == 2) return (k >> 18) & 255u;
  return ((k >> 26) & 15u) | ((unsigned int)lab[id] << 4);
}
__host__ __device__ __forceinline__ int fp_bins(int pass) { return pass < 2 ? 512 : 256; }

__global__ void __launch_bounds__(128) fp_hist_kernel(const unsigned int* __restrict__ key, const unsigned int* __restrict__ idx,
                                                      const unsigned char* __restrict__ lab, int pass, long long n, int units,
                                                      unsigned int* __restrict__ hist /*[bins][units]*/) {
  __shared__ unsigned int sh[4][FP_BINS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 4 + warp;
  const int bins = fp_bins(pass);
  for (int i = lane; i < bins; i += 32) sh[warp][i] = 0;
  __syncwarp();
  if (unit < units) {
    const long long j0 = (long long)unit * FP_CHUNK;
#pragma unroll 8
    for (int s = 0; s < FP_CHUNK; s += 32) {
      const long long j = j0 + s + lane;
      if (j < n) atomicAdd(&sh[warp][fp_digit(pass, key[j], idx[j], lab)], 1u);
    }
    __syncwarp();
    for (int i = lane; i < bins; i += 32) hist[(size_t)i * units + unit] = sh[warp][i];
  }
}

// Exclusive scan of hist in (digit-major, unit-minor) order, two launches of one block per digit:
// row totals first, then every block scans its own row on top of the totals of the lower digits.
__device__ __forceinline__ unsigned int fp_block_sum(unsigned int v, unsigned int* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  unsigned int t = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += red[i];
  return t;
}
__global__ void __launch_bounds__(256) fp_rowsum_kernel(const unsigned int* __restrict__ hist, int units, unsigned int* __restrict__ rowtot) {
  __shared__ unsigned int red[8];
  const unsigned int* row = hist + (size_t)blockIdx.x * units;
  unsigned int s = 0;
  for (int i = threadIdx.x; i < units; i += 256) s += row[i];
  s = fp_block_sum(s, red);
  if (threadIdx.x == 0) rowtot[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) fp_rowscan_kernel(unsigned int* __restrict__ hist, int units, const unsigned int* __restrict__ rowtot) {
  __shared__ unsigned int red[8];
  __shared__ unsigned int wsum[8];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  unsigned int below = 0;
  for (int d = t; d < (int)blockIdx.x; d += 256) below += rowtot[d];
  unsigned int base = fp_block_sum(below, red);       // digits below this one
  unsigned int* row = hist + (size_t)blockIdx.x * units;
  for (int i0 = 0; i0 < units; i0 += 256) {
    const int i = i0 + t;
    const unsigned int v = i < units ? row[i] : 0u;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    unsigned int woff = 0, tot = 0;
    for (int w = 0; w < 8; w++) { if (w < warp) woff += wsum[w]; tot += wsum[w]; }
    if (i < units) row[i] = base + woff + inc - v;
    base += tot;
  }
}

__global__ void __launch_bounds__(128) fp_scatter_kernel(const unsigned int* __restrict__ key, const unsigned int* __restrict__ idx,
                                                         const unsigned char* __restrict__ lab, int pass, long long n, int units,
                                                         const unsigned int* __restrict__ hist, unsigned int* __restrict__ key_out,
                                                         unsigned int* __restrict__ idx_out) {
  __shared__ unsigned int cnt[4][FP_BINS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 4 + warp;
  if (unit >= units) return;
  const int bins = fp_bins(pass);
  for (int i = lane; i < bins; i += 32) cnt[warp][i] = hist[(size_t)i * units + unit];
  __syncwarp();
  const long long j0 = (long long)unit * FP_CHUNK;
  const unsigned int lt = (1u << lane) - 1u;
  // the chunk's keys, indices and digits are fetched up front (the ranking loop below is a serial chain)
  constexpr int IT = FP_CHUNK / 32;
  unsigned int kk[IT], ids[IT], dd[IT];
#pragma unroll
  for (int it = 0; it < IT; it++) {
    const long long j = j0 + it * 32 + lane;
    kk[it] = 0; ids[it] = 0;
    if (j < n) { kk[it] = key[j]; ids[it] = idx[j]; }
  }
#pragma unroll
  for (int it = 0; it < IT; it++) {
    const long long j = j0 + it * 32 + lane;
    dd[it] = j < n ? fp_digit(pass, kk[it], ids[it], lab) : 1024u + lane;     // inactive lanes get unique digits -> no peers
  }
#pragma unroll
  for (int it = 0; it < IT; it++) {
    const bool ok = j0 + it * 32 + lane < n;
    const unsigned int d = dd[it];
    const unsigned int peers = __match_any_sync(0xffffffffu, d);
    if (ok) {
      const unsigned int pos = cnt[warp][d] + __popc(peers & lt);
      key_out[pos] = kk[it];
      idx_out[pos] = ids[it];
    }
    __syncwarp();
    if (ok && (peers & lt) == 0) cnt[warp][d] += __popc(peers);     // lowest lane of each digit group
    __syncwarp();
  }
}

// ---- bin accumulation over the sorted order ------------------------------------------------------------
// lin[i] += sum over selected pixels of feat[px].proto_i ;  binsum[b][c] += feat[px][c] for the last class.
// Eight lanes own FP_SEG consecutive ranks (one float4 of the 128-byte feature row each) and keep the running class /
// bin sums in registers: shared-memory atomics only when the class or the bin changes.  Rows are fetched eight at a
// time (rank -> pixel -> label -> feature row is a chain of dependent loads).
#define FP_SEG 32
__global__ void __launch_bounds__(256) fp_accum_kernel(const unsigned int* __restrict__ sidx, const unsigned char* __restrict__ lab,
                                                       const float* __restrict__ feat, const float* __restrict__ proto,
                                                       const unsigned int* __restrict__ cls_count, int C, long long n,
                                                       double* lin, float* binsum) {
  __shared__ float sproto[FP_MAXC * 32];
  __shared__ float slin[FP_MAXC];
  __shared__ float sbin[32 * 32];
  __shared__ unsigned int sstart[FP_MAXC + 1];
  __shared__ int stouched;
  const int tid = threadIdx.x;
  for (int i = tid; i < C * 32; i += 256) sproto[i] = proto[i];
  for (int i = tid; i < 1024; i += 256) sbin[i] = 0.f;
  if (tid < FP_MAXC) slin[tid] = 0.f;
  if (tid == 0) {
    unsigned int run = 0;
    for (int c = 0; c < C; c++) { sstart[c] = run; run += cls_count[c]; }
    sstart[C] = run;
    stouched = 0;
  }
  __syncthreads();
  const int sub = tid & 7;                 // 8 lanes per feature row (float4 each)
  const long long j0 = ((long long)blockIdx.x * 32 + (tid >> 3)) * FP_SEG;
  int cur_cls = -1, cur_bin = -1;
  float dsum = 0.f;
  float4 fs = make_float4(0, 0, 0, 0);
  bool touched = false;
  for (int s0 = 0; s0 < FP_SEG && j0 + s0 < n; s0 += 8) {
    unsigned int px[8];
    int cls[8];
    float4 f[8];
    int bin[8];                            // -2: row not selected, -1: selected, no bin (not the last class), >= 0: bin
#pragma unroll
    for (int u = 0; u < 8; u++) px[u] = j0 + s0 + u < n ? sidx[j0 + s0 + u] : 0u;
#pragma unroll
    for (int u = 0; u < 8; u++) cls[u] = lab[px[u]];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const long long j = j0 + s0 + u;
      bin[u] = -2;
      f[u] = make_float4(0, 0, 0, 0);
      if (j < n) {
        const unsigned int rank = (unsigned int)(j - sstart[cls[u]]);
        const unsigned int ni = cls_count[cls[u]] >> 5;
        if (ni > 0 && rank < 32u * ni) {
          bin[u] = cls[u] == C - 1 ? (int)(rank / ni) : -1;
          f[u] = *reinterpret_cast<const float4*>(feat + (size_t)px[u] * 32 + sub * 4);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      if (bin[u] == -2) continue;
      if (cls[u] != cur_cls) {
        if (cur_cls >= 0) atomicAdd(&slin[cur_cls], dsum);
        dsum = 0.f; cur_cls = cls[u];
      }
      if (bin[u] != cur_bin) {
        if (cur_bin >= 0) {
          float* bp = &sbin[cur_bin * 32 + sub * 4];
          atomicAdd(bp, fs.x); atomicAdd(bp + 1, fs.y); atomicAdd(bp + 2, fs.z); atomicAdd(bp + 3, fs.w);
          touched = true;
        }
        fs = make_float4(0, 0, 0, 0); cur_bin = bin[u];
      }
      const float* pr = &sproto[cls[u] * 32 + sub * 4];
      dsum += f[u].x * pr[0] + f[u].y * pr[1] + f[u].z * pr[2] + f[u].w * pr[3];
      if (bin[u] >= 0) { fs.x += f[u].x; fs.y += f[u].y; fs.z += f[u].z; fs.w += f[u].w; }
    }
  }
  if (cur_cls >= 0) atomicAdd(&slin[cur_cls], dsum);
  if (cur_bin >= 0) {
    float* bp = &sbin[cur_bin * 32 + sub * 4];
    atomicAdd(bp, fs.x); atomicAdd(bp + 1, fs.y); atomicAdd(bp + 2, fs.z); atomicAdd(bp + 3, fs.w);
    touched = true;
  }
  if (touched) stouched = 1;
  __syncthreads();
  if (tid < C && slin[tid] != 0.f) atomicAdd(lin + tid, (double)slin[tid]);
  if (stouched)
    for (int i = tid; i < 1024; i += 256)
      if (sbin[i] != 0.f) atomicAdd(binsum + i, sbin[i]);
}

// loss, and what the backward needs: pro_last [32][32], ninv[i] = 1/n_i (0 if the class is empty -> NaN loss)
__global__ void fp_final_kernel(const double* lin, const float* binsum, const float* proto, const unsigned int* cls_count,
                                int C, float* loss, float* pro_last) {
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const unsigned int nl = cls_count[C - 1] >> 5;
  float s = 0.f;
  for (int i = tid; i < 1024; i += blockDim.x) {
    const float p = nl ? binsum[i] / (float)nl : nanf("");
    pro_last[i] = p;
    const float d = p - proto[(C - 1) * 32 + (i & 31)];
    s += d * d;
  }
  s = warp_sum(s);
  if ((tid & 31) == 0) red[tid >> 5] = s;
  __syncthreads();
  if (tid < 32) {
    s = tid < (blockDim.x >> 5) ? red[tid] : 0.f;
    s = warp_sum(s);
    if (tid == 0) {
      double l = (double)s / 1024.0;
      for (int c = 0; c < C; c++) {
        const unsigned int ni = cls_count[c] >> 5;
        l += ni ? -lin[c] / (1024.0 * (double)ni) : (double)nanf("");
      }
      *loss = (float)l;
    }
  }
}

// dfeat[px][:] = g * ( -proto_i/(1024 n_i) + [i == C-1] * 2 (pro_last[bin] - proto_last)/(1024 n_last) ) for selected pixels, else 0
__global__ void __launch_bounds__(256) fp_bwd_kernel(const unsigned int* __restrict__ sidx, const unsigned char* __restrict__ lab,
                                                     const float* __restrict__ proto, const float* __restrict__ pro_last,
                                                     const unsigned int* __restrict__ cls_count, int C, long long n,
                                                     const float* __restrict__ gout, float* __restrict__ dfeat) {
  __shared__ float sproto[FP_MAXC * 32];
  __shared__ float spro[32 * 32];
  __shared__ unsigned int sstart[FP_MAXC + 1];
  const int tid = threadIdx.x;
  for (int i = tid; i < C * 32; i += 256) sproto[i] = proto[i];
  for (int i = tid; i < 1024; i += 256) spro[i] = pro_last[i];
  if (tid == 0) {
    unsigned int run = 0;
    for (int c = 0; c < C; c++) { sstart[c] = run; run += cls_count[c]; }
    sstart[C] = run;
  }
  __syncthreads();
  const float g = gout[0];
  const int sub = tid & 7;
  for (long long j = (long long)blockIdx.x * 32 + (tid >> 3); j < n; j += (long long)gridDim.x * 32) {
    const unsigned int px = sidx[j];
    const int cls = lab[px];
    const unsigned int rank = (unsigned int)(j - sstart[cls]);
    const unsigned int ni = cls_count[cls] >> 5;
    float4 o = make_float4(0, 0, 0, 0);
    if (ni > 0 && rank < 32u * ni) {
      const float k1 = -g / (1024.f * (float)ni);
      const float* pr = &sproto[cls * 32 + sub * 4];
      o = make_float4(k1 * pr[0], k1 * pr[1], k1 * pr[2], k1 * pr[3]);
      if (cls == C - 1) {
        const float k2 = 2.f * g / (1024.f * (float)ni);
        const float* pl = &spro[(rank / ni) * 32 + sub * 4];
        o.x += k2 * (pl[0] - pr[0]); o.y += k2 * (pl[1] - pr[1]); o.z += k2 * (pl[2] - pr[2]); o.w += k2 * (pl[3] - pr[3]);
      }
    }
    *reinterpret_cast<float4*>(dfeat + (size_t)px * 32 + sub * 4) = o;
  }
}

static int fp_units(long long n) { return (int)((n + FP_CHUNK - 1) / FP_CHUNK); }

// iws: unsigned int workspace of tcct_fpolar_ws_words(n) words; the first FP_MAXC words (class counts) and
// fws (double[FP_MAXC] lin | float[1024] binsum, see tcct_fpolar_fws_bytes) must be zeroed by the caller.
// After the call iws keeps the sorted pixel order for the backward.
extern "C" long long tcct_fpolar_ws_words(long long n) { return FP_MAXC + 4 * n + (long long)FP_BINS * fp_units(n) + FP_BINS; }
extern "C" long long tcct_fpolar_fws_bytes() { return FP_MAXC * 8 + 1024 * 4; }

struct FpWs {
  unsigned int *cnt, *key0, *key1, *idx0, *idx1, *hist, *rowtot;
};
static FpWs fp_ws(unsigned int* iws, long long n) {
  FpWs w;
  w.cnt = iws; w.key0 = iws + FP_MAXC; w.key1 = w.key0 + n; w.idx0 = w.key1 + n; w.idx1 = w.idx0 + n; w.hist = w.idx1 + n;
  w.rowtot = w.hist + (long long)FP_BINS * fp_units(n);
  return w;
}

extern "C" int tcct_fpolar_forward(const float* feat, const float* logits, const unsigned char* lab, const float* proto,
                                   int B, int C, int H, int W, unsigned int* iws, void* fws, float* loss, float* pro_last,
                                   void* stream) {
  TCCT_CHECK_ARG(C >= 2 && C <= FP_MAXC, "fpolar: 2 <= classes <= %d expected (got %d)", FP_MAXC, C);
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * H * W;
  TCCT_CHECK_ARG(n < (1ll << 31), "fpolar: too many pixels");
  FpWs w = fp_ws(iws, n);
  double* lin = (double*)fws;
  float* binsum = (float*)((char*)fws + FP_MAXC * 8);
  const int units = fp_units(n);
  fp_keys_kernel<<<ceil_div(units, 4), 128, 0, st>>>(logits, lab, B, C, H * W, w.key0, w.idx0, w.cnt, units, w.hist);
  TCCT_CHECK_LAUNCH("fp_keys");
  unsigned int *ki = w.key0, *ko = w.key1, *ii = w.idx0, *io = w.idx1;
  for (int pass = 0; pass < FP_PASSES; pass++) {
    if (pass > 0) {
      fp_hist_kernel<<<ceil_div(units, 4), 128, 0, st>>>(ki, ii, lab, pass, n, units, w.hist);
      TCCT_CHECK_LAUNCH("fp_hist");
    }
    fp_rowsum_kernel<<<fp_bins(pass), 256, 0, st>>>(w.hist, units, w.rowtot);
    TCCT_CHECK_LAUNCH("fp_rowsum");
    fp_rowscan_kernel<<<fp_bins(pass), 256, 0, st>>>(w.hist, units, w.rowtot);
    TCCT_CHECK_LAUNCH("fp_rowscan");
    fp_scatter_kernel<<<ceil_div(units, 4), 128, 0, st>>>(ki, ii, lab, pass, n, units, w.hist, ko, io);
    TCCT_CHECK_LAUNCH("fp_scatter");
    unsigned int* t = ki; ki = ko; ko = t;
    t = ii; ii = io; io = t;
  }
  // 4 passes: the sorted order ends back in key0/idx0
  fp_accum_kernel<<<ceil_div(n, 32 * FP_SEG), 256, 0, st>>>(ii, lab, feat, proto, w.cnt, C, n, lin, binsum);
  TCCT_CHECK_LAUNCH("fp_accum");
  fp_final_kernel<<<1, 256, 0, st>>>(lin, binsum, proto, w.cnt, C, loss, pro_last);
  TCCT_CHECK_LAUNCH("fp_final");
  return TCCT_OK;
}

extern "C" int tcct_fpolar_backward(const unsigned char* lab, const float* proto, const float* pro_last, int B, int C, int H,
                                    int W, const unsigned int* iws, const float* gout, float* dfeat, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)B * H * W;
  FpWs w = fp_ws(const_cast<unsigned int*>(iws), n);
  int blocks = ceil_div(n, 32);
  if (blocks > tcct_num_sms() * 8) blocks = tcct_num_sms() * 8;
  fp_bwd_kernel<<<blocks, 256, 0, st>>>(w.idx0, lab, proto, pro_last, w.cnt, C, n, gout, dfeat);
  TCCT_CHECK_LAUNCH("fp_bwd");
  return TCCT_OK;
}
