// 32->32 channel spatial convolutions (3x3, 1xk, kx1), forward and data gradient, as a TMA-fed tcgen05 pipeline in
// INPUT-STATIONARY form with the output window resident in tensor memory:
//   TMA (cp.async.bulk.tensor, 128-byte swizzle) -> shared-memory line ring -> tcgen05.mma kind::tf32 -> TMEM ring of
//   output-line accumulators -> registers (+bias, BatchNorm statistics) -> swizzled shared staging tile -> TMA store.
// Reference: the nn.Conv2d calls of CrossCNNBlock (task1/nets/tcct.py:803-828), MPUpBlock.prep (887-900).
//
// Work unit: one "line tile" = 128 consecutive pixels of one image row.  A line buffer holds P = 128 + KL - 1 pixels x
// 32 channels exactly as TMA lays a {32 ch, P px} box out with CU_TENSOR_MAP_SWIZZLE_128B: pixel row r at byte r*128,
// 16-byte chunk c at position c ^ (r & 7) -- the canonical K-major SWIZZLE_128B operand layout of tcgen05 (K = 32
// channels = one 128-byte row), so a tap shift ALONG the line is a +128-byte shift of the A descriptor's start address
// (image borders are TMA out-of-bounds zero fill).
//
// Taps ACROSS lines (KA of them) are folded into the N dimension of the MMA: input line i contributes to the KA output
// lines i-pad .. i+pad, whose fp32 accumulators (128 pixels x 32 channels = 32 TMEM columns each) sit next to each other
// in a ring of R blocks of tensor memory, and the weights of those taps sit next to each other in shared memory in the
// matching order.  One MMA of M = 128, N = 32 * (#output lines), K = 8 therefore applies one input line tile to ALL the
// output lines it feeds: a 3x3 conv issues 13 MMAs per line tile (N = 96) instead of 36 (N = 32), a 13x1 conv ~10 instead
// of 52.  At N = 32 the tensor pipe is bound by the 4 KB A-operand read from shared memory per MMA (measured 34-44 cycles
// per MMA against a 16-cycle math floor; the round-1 kernel spent 1 570 of its 2 050 cycles per line tile blocked in MMA
// issue and 460 in mbarrier round trips, scripts/conv_ts.py); widening N amortises that read, and every input line is
// consumed by one burst of MMAs and released, so the shared-memory ring is pure prefetch depth (two lines per slot = one
// mbarrier round trip per two lines).  An output line is complete after input line j+pad; its block is drained by the
// epilogue warps while the MMAs go on in the other blocks of the ring.
// Warp roles: 0-3 epilogue (TMEM lane quarter = warp), 4 MMA issuer, 5 TMA producer.
#include "tma.cuh"
#include <stdlib.h>

#define CT_NS_MAX 12
#define CT_R_MAX 16
#define CT_THREADS 192
#define CT_STAGE_BYTES 16384

struct LineConvArgs {
  const float* wu;       // packed fmt 2: [tap = ka * KL + kl][n 32][chunk ^ (n & 7)][4] (tf32-rounded), 4096 B per tap
  const float* bias;     // [32] or null
  double* stats;         // [2 * stats_C] or null: this launch adds to sum[stats_c0 + i], sq[stats_C + stats_c0 + i], i < 32
  int stats_act, stats_C, stats_c0;
  int x_c0, y_c0;        // first of the 32 input / output channels this launch reads / writes (the tensors may carry more)
  int accumulate;        // 1: add into y (TMA reduce-add store) instead of overwriting it
  int B, H, W;
  int KL, KA;            // taps along / across the line
  int L, NL;             // line length, lines per image
  int vertical;          // 1: lines are image columns (1xk kernels: the taps then run ACROSS lines and ride in N)
  int strips;            // L / 128
  int tiles_total, tiles_per_cta;
  int P;                 // pixel rows per line buffer
  int LPS;               // input lines per ring slot (one mbarrier round trip per slot)
  int NS;                // ring slots
  int R;                 // TMEM ring: accumulator blocks (power of two, >= KA + 2)
  unsigned int line_bytes;   // multiple of 1024
  int dbg;               // experiment knobs (TCCT_CONV_DBG): 1 no staging / TMA stores, 2 first K step only, 4 no TMA loads
  long long* ts;         // experiment (TCCT_CONV_TS): per-line clock64 stamps of CTA 0 ([line][16]) or null
};
#ifdef TCCT_CONV_TIMELINE
#define TS(slot, idx) do { if (a.ts && blockIdx.x == 0 && (idx) < 64) a.ts[(size_t)(idx) * 16 + (slot)] = clock64(); } while (0)
#else
#define TS(slot, idx) do { } while (0)
#endif

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

// A run of n consecutive output-line blocks starting at ring position c (TMEM column 32 * (c mod R)) is issued as up to 3
// MMA pieces: contiguous in the TMEM ring (no wrap) and at most 8 blocks (N <= 256) each.  Plain scalars, no arrays: the
// issuing thread's instruction count per input line is what bounds the kernel once the MMAs are wide.
template <int R>
__device__ __forceinline__ void cut3(int c, int n, int& b0, int& n0, int& b1, int& n1, int& b2, int& n2) {
  b0 = c & (R - 1); n0 = min(n, min(R - b0, 8)); c += n0; n -= n0;
  b1 = c & (R - 1); n1 = min(n, min(R - b1, 8)); c += n1; n -= n1;
  b2 = c & (R - 1); n2 = n;
}
// one MMA piece: n blocks at ring block b, weights from slot `tap` on; ao / bo: descriptor low words of this K step
#define CT_MMA(b, n, tap, ao, bo, acc)                                                                                   \
  tc_mma_tf32(tmem_base + (uint32_t)(b) * 32u, desc_hi | (ao), desc_hi | ((bo) + (uint32_t)(tap) * 256u),               \
              idesc_base | ((uint32_t)(n) << 19), acc)

template <int TKL, int LOGR>
__global__ void __launch_bounds__(CT_THREADS, 1) conv_line_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                      const __grid_constant__ CUtensorMap tmy,
                                                                      const LineConvArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KA = a.KA, KL = TKL;
  const int T = KL * KA;
  const int NS = a.NS, LPS = a.LPS;
  constexpr int R = 1 << LOGR;
  // carve-up (every tile 1024-byte aligned): ring | staging x2 | weights | bias, stats, barriers, tmem ptr
  const uint32_t base_s = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_p = smem_raw + (base_s - smem_u32(smem_raw));
  const uint32_t slot_bytes = a.line_bytes * (uint32_t)LPS;
  const uint32_t ring_s = base_s;
  const uint32_t stage_s = ring_s + (uint32_t)NS * slot_bytes;
  const uint32_t w_s = stage_s + 2u * CT_STAGE_BYTES;
  unsigned char* p_stage = base_p + (size_t)NS * slot_bytes;
  unsigned char* p_w = p_stage + 2 * CT_STAGE_BYTES;
  float* s_bias = reinterpret_cast<float*>(p_w + (size_t)T * 4096);
  float* s_part = s_bias + 32;                                    // [4 warps][64] statistics partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_part + 256);     // full[NS_MAX], empty[NS_MAX], tfull[R_MAX], tempty[R_MAX], wfull
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * CT_NS_MAX + 2 * CT_R_MAX + 1);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * CT_NS_MAX;
  const uint32_t bar_tfull = bar_empty + 8 * CT_NS_MAX, bar_tempty = bar_tfull + 8 * CT_R_MAX, bar_w = bar_tempty + 8 * CT_R_MAX;

  const int t0 = blockIdx.x * a.tiles_per_cta;
  const int t1 = min(t0 + a.tiles_per_cta, a.tiles_total);

  // ---- one-time setup
  if (tid < 32) s_bias[tid] = a.bias ? a.bias[tid] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < NS; i++) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < R; i++) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 128); }
    mbar_init(bar_w, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    if (R <= 8) tmem_alloc<256>(smem_u32(s_tmem)); else tmem_alloc<512>(smem_u32(s_tmem));
  }
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
        const int blk = out_cnt & (R - 1);
        const int sb = out_cnt & 1;
        mbar_wait(bar_tfull + 8 * blk, (out_cnt >> LOGR) & 1);
        tc_fence_after();
        if (tid == 0) TS(8, out_cnt);
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(blk * 32), v);
        tc_fence_before();
        mbar_arrive(bar_tempty + 8 * blk);
        if (tid == 0) TS(9, out_cnt);
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] += s_bias[i];
        // the staging buffer written two tiles ago must have been read by its TMA store
        if (tid == 0) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        if (tid == 0) TS(10, out_cnt);
        unsigned char* row = p_stage + (size_t)sb * CT_STAGE_BYTES + (size_t)m * 128;
        if (!(a.dbg & 1)) {
#pragma unroll
        for (int c = 0; c < 8; c++)
          *reinterpret_cast<float4*>(row + ((c ^ (m & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        }
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (tid == 0) TS(11, out_cnt);
        if (tid == 0 && !(a.dbg & 1)) {
          const uint32_t src = stage_s + (uint32_t)sb * CT_STAGE_BYTES;
          const int c1 = a.vertical ? l : s.strip * 128, c2 = a.vertical ? s.strip * 128 : l;
          if (a.accumulate) tma_reduce_add_4d(&tmy, a.y_c0, c1, c2, s.b, src);
          else tma_store_4d(&tmy, a.y_c0, c1, c2, s.b, src);
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
              const float u = stat_act(a.stats_act, v[i]);
              st_sum[i] += u; st_sq[i] += u * u;
            }
          }
        }
      }
    }
    if (tid == 0) tma_store_wait<0>();
    if (a.stats) {
      // 32 per-lane partial sums x 32 lanes -> lane c holds channel c's warp total: 31 shuffles per array (butterfly with
      // halving payload) instead of 32 x 5; the four warps' totals meet in shared memory without atomics
      warp_transpose_sum(st_sum, lane);
      warp_transpose_sum(st_sq, lane);
      s_part[warp * 64 + lane] = st_sum[0];
      s_part[warp * 64 + 32 + lane] = st_sq[0];
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    // The whole warp walks the (warp-uniform) loops; one elected lane issues the tcgen05 instructions of an input line in
    // one go.  K-major SWIZZLE_128B descriptors: SBO = 1024 B between 8-row groups, LBO field 1 (unused), version 1; the
    // swizzle XOR is a function of the absolute shared-memory address bits (as for TMA writes), so a tap shift of whole
    // pixel rows keeps base_offset 0.  Descriptor low words (start address in 16-byte units) advance by +8 per tap along
    // the line (one 128-byte pixel row), +2 per K step of 8 tf32 (32 bytes), +256 per weight slot.
    const uint32_t idesc_base = umma_idesc_tf32(128, 0, 0, 0);
    const uint64_t desc_hi = (uint64_t)(uint32_t)(umma_desc(0u, 16u, 1024u, 2u, 0u) >> 32) << 32;
    const uint32_t a_lo0 = (uint32_t)umma_desc(ring_s, 16u, 1024u, 2u, 0u);
    const uint32_t b_lo0 = (uint32_t)umma_desc(w_s, 16u, 1024u, 2u, 0u);
    const uint32_t line16 = a.line_bytes >> 4;
    int t = t0, out_base = 0;                 // out_base: output lines of earlier segments (ring position of this one's first)
    int slot = 0, phase = 0;
    LSeg s;
    mbar_wait(bar_w, 0);                                        // weights have landed
    while (next_lseg(a, t, t1, s)) {
      const int nout = s.l1 - s.l0;
      for (int i0 = s.in0; i0 <= s.in1; i0 += LPS) {
        const int ni = min(LPS, s.in1 - i0 + 1);
        if (lane == 0) TS(0, i0 - s.in0);
        mbar_wait(bar_full + 8 * slot, phase);
        if (lane == 0) TS(1, i0 - s.in0);
        for (int q = 0; q < ni; q++) {
          const int i = i0 + q;
          if (lane == 0) TS(2, i - s.in0);
          // output lines fed by input line i (relative to l0), and which of them see their first contribution now
          const int j0 = max(i - padA, s.l0) - s.l0, j1 = min(i + padA, s.l1 - 1) - s.l0;
          const int jn = (i == s.in0) ? j0 : ((i + padA <= s.l1 - 1) ? j1 : j1 + 1);     // new blocks: [jn, j1]
          // the blocks the lines of this slot start replace output lines R earlier, whose accumulators must have been read.  The
          // epilogue drains in order, so one wait per ring slot, for the newest block of its last line, covers them all.
          if (q == 0 && (i == s.in0 || i + padA <= s.l1 - 1)) {
            const int c = out_base + min(i + ni - 1 + padA, s.l1 - 1) - s.l0;
            mbar_wait(bar_tempty + 8 * (c & (R - 1)), ((c >> LOGR) & 1) ^ 1);
          }
          tc_fence_after();
          if (lane == 0) TS(3, i - s.in0);
          // weight slot of relative output line j: ka' = (j + l0) - i + padA (ascending with j: weights are stored in descending
          // tap order so that ascending TMEM columns meet ascending shared-memory rows)
          const int tap0 = j0 + s.l0 - i + padA;
          const int nwin = j1 - j0 + 1;
          int b0, n0, b1, n1, b2, n2;
          cut3<R>(out_base + j0, nwin, b0, n0, b1, n1, b2, n2);
          const bool steady = (jn == j1) && (j1 > j0);          // exactly one new block: the last of the window
          const uint32_t a_lo = a_lo0 + (uint32_t)slot * (line16 * (uint32_t)LPS) + (uint32_t)q * line16;
          if (lane == 0) TS(4, i - s.in0);
          if (elect_one()) {
            TS(6, i - s.in0);
            // first K step: blocks that start their accumulation here are written (accumulate = 0), the others added to
            if (steady) {
              int c0, m0, c1, m1, c2, m2;
              cut3<R>(out_base + j0, nwin - 1, c0, m0, c1, m1, c2, m2);
              CT_MMA(c0, m0, tap0, a_lo, b_lo0, 1u);
              if (m1 > 0) CT_MMA(c1, m1, tap0 + m0, a_lo, b_lo0, 1u);
              if (m2 > 0) CT_MMA(c2, m2, tap0 + m0 + m1, a_lo, b_lo0, 1u);
              CT_MMA((out_base + j1) & (R - 1), 1, tap0 + nwin - 1, a_lo, b_lo0, 0u);
            } else {
              // segment start (every block of the window is new) or segment end (none is)
              const uint32_t acc = (jn > j1) ? 1u : 0u;
              CT_MMA(b0, n0, tap0, a_lo, b_lo0, acc);
              if (n1 > 0) CT_MMA(b1, n1, tap0 + n0, a_lo, b_lo0, acc);
              if (n2 > 0) CT_MMA(b2, n2, tap0 + n0 + n1, a_lo, b_lo0, acc);
            }
            TS(12, i - s.in0);
            if (!(a.dbg & 2)) {
              const int t1 = tap0 + n0, t2 = t1 + n1;
              if (n1 == 0) {
#pragma unroll
                for (int kl = 0; kl < TKL; kl++)
#pragma unroll
                  for (int ks = 0; ks < 4; ks++) {
                    if (kl == 0 && ks == 0) continue;
                    CT_MMA(b0, n0, tap0, a_lo + 8u * kl + 2u * ks, b_lo0 + (uint32_t)(kl * KA) * 256u + 2u * ks, 1u);
                  }
              } else if (n2 == 0) {
#pragma unroll
                for (int kl = 0; kl < TKL; kl++)
#pragma unroll
                  for (int ks = 0; ks < 4; ks++) {
                    if (kl == 0 && ks == 0) continue;
                    const uint32_t ao = a_lo + 8u * kl + 2u * ks, bo = b_lo0 + (uint32_t)(kl * KA) * 256u + 2u * ks;
                    CT_MMA(b0, n0, tap0, ao, bo, 1u);
                    CT_MMA(b1, n1, t1, ao, bo, 1u);
                  }
              } else {
#pragma unroll
                for (int kl = 0; kl < TKL; kl++)
#pragma unroll
                  for (int ks = 0; ks < 4; ks++) {
                    if (kl == 0 && ks == 0) continue;
                    const uint32_t ao = a_lo + 8u * kl + 2u * ks, bo = b_lo0 + (uint32_t)(kl * KA) * 256u + 2u * ks;
                    CT_MMA(b0, n0, tap0, ao, bo, 1u);
                    CT_MMA(b1, n1, t1, ao, bo, 1u);
                    CT_MMA(b2, n2, t2, ao, bo, 1u);
                  }
              }
            }
            TS(13, i - s.in0);
            // output lines whose last contribution this was: i - padA, or everything still open at the segment's last input line
            const int jc = i - padA - s.l0;
            if (i != s.in1) {
              if (jc >= 0) tc_commit(bar_tfull + 8 * ((out_base + jc) & (R - 1)));
            } else {
              for (int j = max(jc, 0); j < nout; j++) tc_commit(bar_tfull + 8 * ((out_base + j) & (R - 1)));
            }
            TS(14, i - s.in0);
          }
          __syncwarp();
          if (lane == 0) TS(5, i - s.in0);
        }
        if (elect_one()) tc_commit(bar_empty + 8 * slot);       // every line of the slot has been consumed
        __syncwarp();
        if (++slot == NS) { slot = 0; phase ^= 1; }
      }
      out_base += nout;
    }
  } else {
    // ===================== TMA producer (one thread): weights, then input lines -> ring =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmx); tma_prefetch_desc(&tmy);
      mbar_expect_tx(bar_w, (uint32_t)T * 4096u);
      // shared-memory weight order: [kl][KA - 1 - ka] (see the MMA issuer)
      for (int ka = 0; ka < KA; ka++)
        for (int kl = 0; kl < KL; kl++)
          bulk_load(w_s + (uint32_t)(kl * KA + (KA - 1 - ka)) * 4096u, a.wu + (size_t)(ka * KL + kl) * 1024, 4096u, bar_w);
      int t = t0, slot = 0, phase = 1;
      LSeg s;
      const uint32_t bytes = (uint32_t)a.P * 128u;
      while (next_lseg(a, t, t1, s)) {
        for (int i0 = s.in0; i0 <= s.in1; i0 += LPS) {
          const int ni = min(LPS, s.in1 - i0 + 1);
          mbar_wait(bar_empty + 8 * slot, phase);
          if (a.dbg & 4) { mbar_arrive(bar_full + 8 * slot); if (++slot == NS) { slot = 0; phase ^= 1; } continue; }
          mbar_expect_tx(bar_full + 8 * slot, bytes * (uint32_t)ni);
          for (int q = 0; q < ni; q++) {
            const uint32_t dst = ring_s + (uint32_t)slot * slot_bytes + (uint32_t)q * a.line_bytes;
            if (a.vertical) tma_load_4d(dst, &tmx, a.x_c0, i0 + q, s.strip * 128 - padL, s.b, bar_full + 8 * slot);
            else tma_load_4d(dst, &tmx, a.x_c0, s.strip * 128 - padL, i0 + q, s.b, bar_full + 8 * slot);
          }
          if (++slot == NS) { slot = 0; phase ^= 1; }
        }
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (a.stats && tid < 64) atomicAdd(a.stats + (tid < 32 ? a.stats_c0 + tid : a.stats_C + a.stats_c0 + tid - 32), (double)((s_part[tid] + s_part[64 + tid]) + (s_part[128 + tid] + s_part[192 + tid])));
  if (warp == 4) {
    if (R <= 8) tmem_dealloc<256>(tmem_base); else tmem_dealloc<512>(tmem_base);
  }
}

template <int TKL>
static void launch_line_conv(const CUtensorMap& tmx, const CUtensorMap& tmy, const LineConvArgs& a, int ctas, size_t smem, cudaStream_t st) {
  if (a.R == 8) {
    cudaFuncSetAttribute(conv_line_tma_kernel<TKL, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_line_tma_kernel<TKL, 3><<<ctas, CT_THREADS, smem, st>>>(tmx, tmy, a);
  } else {
    cudaFuncSetAttribute(conv_line_tma_kernel<TKL, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_line_tma_kernel<TKL, 4><<<ctas, CT_THREADS, smem, st>>>(tmx, tmy, a);
  }
}

// 1 if this shape runs on the TMA/tcgen05 path
extern "C" int tcct_conv_tma_supported(int H, int W, int Cin, int Cout, int KH, int KW) {
  if (Cin != 32 || Cout != 32) return 0;
  if (!((KH == 3 && KW == 3) || (KH == 1 && KW >= 3 && KW <= 13 && (KW & 1)) || (KW == 1 && KH >= 3 && KH <= 13 && (KH & 1)))) return 0;
  // 3x3 and kx1: lines are image rows (W % 128); 1xk: lines are image columns when H % 128 == 0 (taps across lines), rows otherwise
  if (KH == 1 ? (H % 128 != 0 && W % 128 != 0) : (W % 128 != 0)) return 0;
  return tcct_tensor_map_encoder() != nullptr ? 1 : 0;
}

// One 32 -> 32 channel slice of a convolution whose tensors may carry more channels: reads channels [x_c0, x_c0 + 32) of x
// ([B,H,W,x_ch]), writes (accumulate = 0) or adds into (1) channels [y_c0, y_c0 + 32) of y ([B,H,W,y_ch]).  wu: the fmt-2 pack of
// this (output tile, input tile) weight block; bias: the 32 biases of the output slice or null; stats: the FULL statistics buffer
// of y ([2 * y_ch]) or null -- only meaningful when the slice is the whole reduction (x_ch == 32).
extern "C" int tcct_conv2d_tma_slice(const float* x, int x_ch, int x_c0, const float* wu, const float* bias, float* y, int y_ch, int y_c0,
                                     int accumulate, int B, int H, int W, int KH, int KW, double* stats, int stats_act, void* stream) {
  TCCT_CHECK_ARG(tcct_conv_tma_supported(H, W, 32, 32, KH, KW), "conv2d_tma: unsupported shape %dx%d kernel %dx%d", H, W, KH, KW);
  TCCT_CHECK_ARG(x_ch % 32 == 0 && y_ch % 32 == 0 && x_c0 % 32 == 0 && y_c0 % 32 == 0 && x_c0 + 32 <= x_ch && y_c0 + 32 <= y_ch,
                 "conv2d_tma: channel slices must be 32-aligned (x %d/%d, y %d/%d)", x_c0, x_ch, y_c0, y_ch);
  LineConvArgs a;
  a.wu = wu; a.bias = bias; a.stats = stats; a.stats_act = stats_act; a.stats_C = y_ch; a.stats_c0 = y_c0;
  a.x_c0 = x_c0; a.y_c0 = y_c0; a.accumulate = accumulate;
  { const char* e = getenv("TCCT_CONV_DBG"); a.dbg = e ? atoi(e) : 0; }
  { const char* e = getenv("TCCT_CONV_TS"); a.ts = e ? (long long*)strtoull(e, nullptr, 0) : nullptr; }
  a.B = B; a.H = H; a.W = W;
  // lines are image rows: KL = taps along the row (kw), KA = taps across rows (kh); the packed tap index kh * KW + kw is
  // ka * KL + kl in this orientation
  a.vertical = (KH == 1 && H % 128 == 0) ? 1 : 0;
  { const char* e = getenv("TCCT_CONV_1XK_ROWS"); if (e && atoi(e) && W % 128 == 0) a.vertical = 0; }
  if (a.vertical) { a.KL = 1; a.KA = KW; a.L = H; a.NL = W; }     // packed tap index kw = ka
  else { a.KL = KW; a.KA = KH; a.L = W; a.NL = H; }
  a.strips = a.L / 128;
  a.tiles_total = B * a.strips * a.NL;
  const int sms = tcct_num_sms();
  a.tiles_per_cta = ceil_div(a.tiles_total, sms);
  const int ctas = ceil_div(a.tiles_total, a.tiles_per_cta);
  a.P = 128 + a.KL - 1;
  a.line_bytes = (unsigned int)((a.P * 128 + 1023) / 1024 * 1024);
  a.R = (a.KA <= 3) ? 8 : 16;
  a.LPS = 2;
  { const char* e = getenv("TCCT_CONV_LPS"); if (e && atoi(e) > 0) a.LPS = atoi(e); }
  const size_t fixed = 1024 + 2 * CT_STAGE_BYTES + (size_t)a.KL * a.KA * 4096 + 288 * 4 + (2 * CT_NS_MAX + 2 * CT_R_MAX + 1) * 8 + 16;
  a.NS = (int)((227 * 1024 - fixed) / ((size_t)a.line_bytes * a.LPS));
  if (a.NS > CT_NS_MAX) a.NS = CT_NS_MAX;
  TCCT_CHECK_ARG(a.NS >= 2, "conv2d_tma: ring too small (%d slots)", a.NS);
  const size_t smem = fixed + (size_t)a.NS * a.line_bytes * a.LPS;
  TCCT_CHECK_ARG(smem <= 227 * 1024, "conv2d_tma: shared memory budget exceeded (%zu B)", smem);
  CUtensorMap tmx, tmy;
  const unsigned long long dims_x[4] = {(unsigned long long)x_ch, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
  const unsigned long long dims_y[4] = {(unsigned long long)y_ch, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
  const unsigned long long px = 4ull * x_ch, py = 4ull * y_ch;
  const unsigned long long strides_x[3] = {px, (unsigned long long)W * px, (unsigned long long)H * W * px};
  const unsigned long long strides_y[3] = {py, (unsigned long long)W * py, (unsigned long long)H * W * py};
  unsigned int box_in[4] = {32u, 1u, 1u, 1u}, box_out[4] = {32u, 1u, 1u, 1u};
  box_in[a.vertical ? 2 : 1] = (unsigned int)a.P;
  box_out[a.vertical ? 2 : 1] = 128u;
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmx, x, 4, dims_x, strides_x, box_in, 1), "conv2d_tma: cuTensorMapEncodeTiled failed (input)");
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmy, y, 4, dims_y, strides_y, box_out, 1), "conv2d_tma: cuTensorMapEncodeTiled failed (output)");
  switch (a.KL) {
    case 1: launch_line_conv<1>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream); break;
    case 3: launch_line_conv<3>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream); break;
    case 5: launch_line_conv<5>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream); break;
    case 7: launch_line_conv<7>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream); break;
    case 9: launch_line_conv<9>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream); break;
    case 11: launch_line_conv<11>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream); break;
    default: launch_line_conv<13>(tmx, tmy, a, ctas, smem, (cudaStream_t)stream); break;
  }
  tcct_count_route(TCCT_ROUTE_CONV_TMA);
  TCCT_CHECK_LAUNCH("conv2d_tma");
  return TCCT_OK;
}

// wu: weights packed by tcct_pack_weights with fmt = 2 ([tap][n][chunk ^ (n & 7)][4], tf32-rounded)
extern "C" int tcct_conv2d_tma(const float* x, const float* wu, const float* bias, float* y, int B, int H, int W, int KH,
                               int KW, double* stats, int stats_act, void* stream) {
  return tcct_conv2d_tma_slice(x, 32, 0, wu, bias, y, 32, 0, 0, B, H, W, KH, KW, stats, stats_act, stream);
}
