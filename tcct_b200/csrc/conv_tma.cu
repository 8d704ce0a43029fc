// 32->32 channel spatial convolutions (3x3, 1xk, kx1), forward and data gradient, as a TMA-fed tcgen05 pipeline:
//   TMA (cp.async.bulk.tensor, 128-byte swizzle) -> shared-memory line ring -> tcgen05.mma kind::tf32 -> TMEM ->
//   registers (+bias, BatchNorm statistics) -> swizzled shared staging tile -> TMA store.
// Reference: the nn.Conv2d calls of CrossCNNBlock (task1/nets/tcct.py:803-828), MPUpBlock.prep (887-900).
//
// Work unit: one "line tile" = 128 consecutive pixels of one line (an image row, or an image column for kx1 kernels
// so that the taps always run ALONG the line).  A ring slot holds P = 128 + KL - 1 pixels x 32 channels exactly as
// TMA lays a {32 ch, P px} box out with CU_TENSOR_MAP_SWIZZLE_128B: pixel row r at byte r*128, 16-byte chunk c at
// position c ^ (r & 7).  That is the canonical K-major SWIZZLE_128B operand layout of tcgen05 (K = 32 channels = one
// 128-byte row), so a tap shift along the line is a +128-byte shift of the A descriptor's start address and every
// tap of every kernel shape reads the SAME staged copy (each input line leaves HBM/L2 once per strip; image borders
// are TMA out-of-bounds zero fill).  3x3 kernels march down the image with the ring holding the 3 live input rows.
// M = 128 pixels, N = 32 output channels, K = 8 per MMA -> taps x 4 MMAs per line tile, fp32 accumulators double
// buffered in tensor memory.  Warp roles: 0-3 epilogue (TMEM lane quarter = warp), 4 MMA issuer, 5 TMA producer.
#include "tma.cuh"

#define CT_NS_MAX 8
#define CT_THREADS 192
#define CT_STAGE_BYTES 16384

struct LineConvArgs {
  const float* wu;       // packed fmt 2: [tap][n 32][chunk ^ (n & 7)][4] (tf32-rounded), 4096 B per tap
  const float* bias;     // [32] or null
  double* stats;         // [64] or null
  int stats_act;
  int B, H, W;
  int KL, KA;            // taps along / across the line
  int L, NL;             // line length, lines per image
  int vertical;          // 1: lines are image columns
  int strips;            // L / 128
  int tiles_total, tiles_per_cta;
  int P;                 // pixel rows per ring slot
  int NS;                // ring slots
  unsigned int slot_bytes;   // multiple of 1024
};

// Segment bookkeeping shared by all roles: the CTA owns tiles [t0, t1); a segment is a maximal run of tiles in
// the same (image, strip); within a segment output lines [l0, l1) need input lines [in0, in1].
struct LSeg { int b, strip, l0, l1, in0, in1; };
__device__ __forceinline__ bool next_lseg(const LineConvArgs& a, int& t, int t1, LSeg& s) {
  if (t >= t1) return false;
  const int per_img = a.strips * a.NL;
  s.b = t / per_img;
  const int r = t - s.b * per_img;
  s.strip = r / a.NL;
  s.l0 = r - s.strip * a.NL;
  const int n = min(a.NL - s.l0, t1 - t);
  s.l1 = s.l0 + n;
  const int pad = a.KA >> 1;
  s.in0 = max(s.l0 - pad, 0);
  s.in1 = min(s.l1 - 1 + pad, a.NL - 1);
  t += n;
  return true;
}

// TKA, TKL: compile-time tap counts (0 = read them from the arguments).  The MMA issue stream of one line tile is
// straight-line code for the shapes the network uses: one thread feeds the tensor core, so every scalar
// instruction between two tcgen05.mma is time the tensor pipe idles.
template <int TKA, int TKL>
__global__ void __launch_bounds__(CT_THREADS, 1) conv_line_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                      const __grid_constant__ CUtensorMap tmy,
                                                                      const LineConvArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KA = TKA ? TKA : a.KA, KL = TKL ? TKL : a.KL;
  const int T = KL * KA;
  const int NS = a.NS;
  // carve-up (every tile 1024-byte aligned): ring | staging x2 | weights | bias, stats, barriers, tmem ptr
  const uint32_t base_s = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_p = smem_raw + (base_s - smem_u32(smem_raw));
  const uint32_t ring_s = base_s;
  const uint32_t stage_s = ring_s + (uint32_t)NS * a.slot_bytes;
  const uint32_t w_s = stage_s + 2u * CT_STAGE_BYTES;
  unsigned char* p_stage = base_p + (size_t)NS * a.slot_bytes;
  unsigned char* p_w = p_stage + 2 * CT_STAGE_BYTES;
  float* s_bias = reinterpret_cast<float*>(p_w + (size_t)T * 4096);
  float* s_stats = s_bias + 32;                                   // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stats + 64);     // full[NS_MAX], empty[NS_MAX], tfull[2], tempty[2], wfull
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * CT_NS_MAX + 5);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * CT_NS_MAX;
  const uint32_t bar_tfull = bar_empty + 8 * CT_NS_MAX, bar_tempty = bar_tfull + 16, bar_w = bar_tempty + 16;

  const int t0 = blockIdx.x * a.tiles_per_cta;
  const int t1 = min(t0 + a.tiles_per_cta, a.tiles_total);

  // ---- one-time setup
  if (tid < 32) s_bias[tid] = a.bias ? a.bias[tid] : 0.f;
  if (tid < 64) s_stats[tid] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < NS; i++) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; i++) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 128); }
    mbar_init(bar_w, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<64>(smem_u32(s_tmem));      // two 32-column accumulators
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int padL = KL >> 1, padA = KA >> 1;

  if (warp < 4) {
    // ===================== epilogue: TMEM -> registers -> (+bias, statistics) -> staging -> TMA store =====================
    float st_sum[32], st_sq[32];
#pragma unroll
    for (int i = 0; i < 32; i++) st_sum[i] = st_sq[i] = 0.f;
    const int m = warp * 32 + lane;              // pixel of the line tile == TMEM lane
    const bool lrelu_stats = a.stats_act == ACT_LRELU;
    int t = t0, out_cnt = 0;
    LSeg s;
    while (next_lseg(a, t, t1, s)) {
      for (int l = s.l0; l < s.l1; l++, out_cnt++) {
        const int acc = out_cnt & 1;
        mbar_wait(bar_tfull + 8 * acc, (out_cnt >> 1) & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 32), v);
        tc_fence_before();
        mbar_arrive(bar_tempty + 8 * acc);
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] += s_bias[i];
        // the staging buffer written two tiles ago must have been read by its TMA store
        if (tid == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        unsigned char* row = p_stage + (size_t)acc * CT_STAGE_BYTES + (size_t)m * 128;
#pragma unroll
        for (int c = 0; c < 8; c++)
          *reinterpret_cast<float4*>(row + ((c ^ (m & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (tid == 0) {
          const uint32_t src = stage_s + (uint32_t)acc * CT_STAGE_BYTES;
          if (a.vertical) tma_store_4d(&tmy, 0, l, s.strip * 128, s.b, src);
          else tma_store_4d(&tmy, 0, s.strip * 128, l, s.b, src);
          tma_store_commit();
        }
        if (a.stats) {
          if (lrelu_stats) {
#pragma unroll
            for (int i = 0; i < 32; i++) {
              const float u = v[i] > 0.f ? v[i] : 0.01f * v[i];
              st_sum[i] += u; st_sq[i] += u * u;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; i++) {
              const float u = act_fwd(a.stats_act, v[i]);
              st_sum[i] += u; st_sq[i] += u * u;
            }
          }
        }
      }
    }
    if (tid == 0) tma_store_wait<0>();
    if (a.stats) {
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const float su = warp_sum(st_sum[i]), sq = warp_sum(st_sq[i]);
        if (lane == 0) { atomicAdd(&s_stats[i], su); atomicAdd(&s_stats[32 + i], sq); }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    // The whole warp walks the (warp-uniform) loops; one elected lane issues the tcgen05 instructions of a line tile in
    // one go.  Descriptor high words are constant; the low words (start address in 16-byte units) advance by fixed
    // steps: +8 per tap (one 128-byte pixel row), +2 per K step of 8 tf32 (32 bytes), +256 per weight tap.
    const uint32_t idesc = umma_idesc_tf32(128, 32, 0, 0);
    // K-major SWIZZLE_128B descriptors: SBO = 1024 B between 8-row groups, LBO field 1 (unused), version 1.
    // The swizzle XOR is a function of the absolute shared-memory address bits (as for TMA writes), so a tap shift
    // of whole pixel rows keeps base_offset 0 (verified on B200: (addr >> 7) & 7 there scrambles the operand).
    const uint64_t desc_hi = (uint64_t)(uint32_t)(umma_desc(0u, 16u, 1024u, 2u, 0u) >> 32) << 32;
    const uint32_t a_lo0 = (uint32_t)umma_desc(ring_s, 16u, 1024u, 2u, 0u);
    const uint32_t b_lo0 = (uint32_t)umma_desc(w_s, 16u, 1024u, 2u, 0u);
    const uint32_t slot16 = a.slot_bytes >> 4;
    int t = t0, out_cnt = 0;
    int wslot = 0, wphase = 0, waited = 0, seq_base = 0;      // next ring slot whose "full" has not been observed yet
    LSeg s;
    mbar_wait(bar_w, 0);                                        // weights have landed
    while (next_lseg(a, t, t1, s)) {
      // ring slot of input line (l0 - padA) (virtual when that line lies above the segment's first input line)
      int cur = (seq_base + (s.l0 - padA) - s.in0) % NS;
      if (cur < 0) cur += NS;
      for (int l = s.l0; l < s.l1; l++, out_cnt++) {
        const int need = seq_base + (min(l + padA, s.in1) - s.in0);       // newest input line this output line reads
        while (waited <= need) {
          mbar_wait(bar_full + 8 * wslot, wphase);
          waited++;
          if (++wslot == NS) { wslot = 0; wphase ^= 1; }
        }
        const int acc = out_cnt & 1;
        mbar_wait(bar_tempty + 8 * acc, ((out_cnt >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 32);
          uint32_t accum = 0;
          int sl = cur;
#pragma unroll
          for (int ka = 0; ka < (TKA ? TKA : 3); ka++) {
            if (ka < KA) {
              const int il = l + ka - padA;
              if (il >= 0 && il < a.NL) {                                     // zero padding across lines: skip the taps
                const uint32_t a_lo = a_lo0 + (uint32_t)sl * slot16;
                const uint32_t b_lo = b_lo0 + (uint32_t)(ka * KL) * 256u;
                if (TKL) {
#pragma unroll
                  for (int kl = 0; kl < TKL; kl++)
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                      tc_mma_tf32(d_tmem, desc_hi | (a_lo + 8u * kl + 2u * ks), desc_hi | (b_lo + 256u * kl + 2u * ks), idesc, accum);
                      accum = 1;
                    }
                } else {
                  for (int kl = 0; kl < KL; kl++)
#pragma unroll
                    for (int ks = 0; ks < 4; ks++) {
                      tc_mma_tf32(d_tmem, desc_hi | (a_lo + 8u * kl + 2u * ks), desc_hi | (b_lo + 256u * kl + 2u * ks), idesc, accum);
                      accum = 1;
                    }
                }
              }
              if (++sl == NS) sl = 0;
            }
          }
          tc_commit(bar_tfull + 8 * acc);
          // the input line (l - padA) is not needed by any later output line of the segment
          if (l - padA >= s.in0 && l + 1 < s.l1) tc_commit(bar_empty + 8 * cur);
        }
        __syncwarp();
        if (++cur == NS) cur = 0;
      }
      // end of segment: release every line still held (cur is now the slot of line l1 - padA)
      if (elect_one()) {
        int sl = cur - 1;
        if (sl < 0) sl += NS;
        for (int il = s.l1 - 1 - padA; il <= s.in1; il++) {
          if (il >= s.in0) tc_commit(bar_empty + 8 * sl);
          if (++sl == NS) sl = 0;
        }
      }
      __syncwarp();
      seq_base += s.in1 - s.in0 + 1;
    }
  } else {
    // ===================== TMA producer (one thread): weights, then input lines -> ring =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmx); tma_prefetch_desc(&tmy);
      mbar_expect_tx(bar_w, (uint32_t)T * 4096u);
      bulk_load(w_s, a.wu, (uint32_t)T * 4096u, bar_w);
      int t = t0, slot = 0, phase = 1;
      LSeg s;
      const uint32_t bytes = (uint32_t)a.P * 128u;
      while (next_lseg(a, t, t1, s)) {
        for (int il = s.in0; il <= s.in1; il++) {
          mbar_wait(bar_empty + 8 * slot, phase);
          const uint32_t dst = ring_s + (uint32_t)slot * a.slot_bytes;
          mbar_expect_tx(bar_full + 8 * slot, bytes);
          if (a.vertical) tma_load_4d(dst, &tmx, 0, il, s.strip * 128 - padL, s.b, bar_full + 8 * slot);
          else tma_load_4d(dst, &tmx, 0, s.strip * 128 - padL, il, s.b, bar_full + 8 * slot);
          if (++slot == NS) { slot = 0; phase ^= 1; }
        }
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (a.stats && tid < 64) atomicAdd(a.stats + tid, (double)s_stats[tid]);
  if (warp == 4) tmem_dealloc<64>(tmem_base);
}

template <int TKA, int TKL>
static void launch_line_conv(const CUtensorMap& tmx, const CUtensorMap& tmy, const LineConvArgs& a, int ctas, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(conv_line_tma_kernel<TKA, TKL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  conv_line_tma_kernel<TKA, TKL><<<ctas, CT_THREADS, smem, st>>>(tmx, tmy, a);
}

// 1 if this shape runs on the TMA/tcgen05 path
extern "C" int tcct_conv_tma_supported(int H, int W, int Cin, int Cout, int KH, int KW) {
  if (Cin != 32 || Cout != 32) return 0;
  if (!((KH == 3 && KW == 3) || (KH == 1 && KW >= 3 && KW <= 13 && (KW & 1)) || (KW == 1 && KH >= 3 && KH <= 13 && (KH & 1)))) return 0;
  const int L = (KW == 1) ? H : W;
  if (L % 128 != 0) return 0;
  return tcct_tensor_map_encoder() != nullptr ? 1 : 0;
}

// wu: weights packed by tcct_pack_weights with fmt = 2 ([tap][n][chunk ^ (n & 7)][4], tf32-rounded)
extern "C" int tcct_conv2d_tma(const float* x, const float* wu, const float* bias, float* y, int B, int H, int W, int KH,
                               int KW, double* stats, int stats_act, void* stream) {
  TCCT_CHECK_ARG(tcct_conv_tma_supported(H, W, 32, 32, KH, KW), "conv2d_tma: unsupported shape %dx%d kernel %dx%d", H, W, KH, KW);
  LineConvArgs a;
  a.wu = wu; a.bias = bias; a.stats = stats; a.stats_act = stats_act;
  a.B = B; a.H = H; a.W = W;
  a.vertical = (KW == 1) ? 1 : 0;
  if (a.vertical) { a.KL = KH; a.KA = 1; a.L = H; a.NL = W; }
  else { a.KL = KW; a.KA = KH; a.L = W; a.NL = H; }
  a.strips = a.L / 128;
  a.tiles_total = B * a.strips * a.NL;
  const int sms = tcct_num_sms();
  a.tiles_per_cta = ceil_div(a.tiles_total, sms);
  const int ctas = ceil_div(a.tiles_total, a.tiles_per_cta);
  a.P = 128 + a.KL - 1;
  a.slot_bytes = (unsigned int)((a.P * 128 + 1023) / 1024 * 1024);
  // as many ring slots as fit (<= 8): the lines beyond the KA live ones are the TMA prefetch distance
  {
    const size_t fixed = 1024 + 2 * CT_STAGE_BYTES + (size_t)a.KL * a.KA * 4096 + 96 * 4 + (2 * CT_NS_MAX + 5) * 8 + 16;
    a.NS = (int)((227 * 1024 - fixed) / a.slot_bytes);
    if (a.NS > CT_NS_MAX) a.NS = CT_NS_MAX;
  }
  TCCT_CHECK_ARG(a.NS >= a.KA + 2, "conv2d_tma: ring too small (%d slots)", a.NS);
  const size_t smem = 1024 + (size_t)a.NS * a.slot_bytes + 2 * CT_STAGE_BYTES + (size_t)a.KL * a.KA * 4096 + 96 * 4 +
                      (2 * CT_NS_MAX + 5) * 8 + 16;
  TCCT_CHECK_ARG(smem <= 227 * 1024, "conv2d_tma: shared memory budget exceeded (%zu B)", smem);
  CUtensorMap tmx, tmy;
  const unsigned long long dims[4] = {32ull, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
  const unsigned long long strides[3] = {128ull, (unsigned long long)W * 128ull, (unsigned long long)H * W * 128ull};
  unsigned int box_in[4] = {32u, 1u, 1u, 1u}, box_out[4] = {32u, 1u, 1u, 1u};
  box_in[a.vertical ? 2 : 1] = (unsigned int)a.P;
  box_out[a.vertical ? 2 : 1] = 128u;
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmx, x, 4, dims, strides, box_in, 1), "conv2d_tma: cuTensorMapEncodeTiled failed (input)");
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmy, y, 4, dims, strides, box_out, 1), "conv2d_tma: cuTensorMapEncodeTiled failed (output)");
  if (a.KA == 3 && a.KL == 3) launch_line_conv<3, 3>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream);
  else if (a.KA == 1 && a.KL == 13) launch_line_conv<1, 13>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream);
  else if (a.KA == 1 && a.KL == 11) launch_line_conv<1, 11>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream);
  else launch_line_conv<0, 0>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream);
  tcct_count_route(TCCT_ROUTE_CONV_TMA);
  TCCT_CHECK_LAUNCH("conv2d_tma");
  return TCCT_OK;
}
