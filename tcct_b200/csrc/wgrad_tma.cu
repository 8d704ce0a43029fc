// Weight gradient of the 32->32 channel spatial convolutions (3x3, 1xk, kx1) as a TMA-fed tcgen05 pipeline:
//   dW[co][ci][tap] += sum_px dy[px][co] * x[px + tap][ci]        dbias[co] += sum_px dy[px][co]
// (autograd's convolution_backward weight path for the nn.Conv2d of CrossCNNBlock, task1/nets/tcct.py:803-828).
//
// The contraction runs over PIXELS, so both operands are fed MN-major straight from the NHWC line buffers TMA writes
// (pixel row r at byte r*128, 32-byte chunks XOR-swizzled with r & 3 -- CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, the
// layout tcgen05 requires for MN-major 32-bit operands, descriptor layout type 1): 32 channels are one 128-byte MN row,
// 4 consecutive pixels one K group.  One tcgen05.mma is M = 128, N = 32, K = 8 pixels:
//   B = x line  : 32 input channels, start shifted by whole pixel rows to select the tap group
//   A = dy line : FOUR copies of the 32 output channels, copy j shifted by j pixel rows (leading-dimension byte
//                 offset = 128 B, the copies overlap in shared memory), so one MMA accumulates four neighbouring
//                 taps along the line: D[(j, co), ci] += sum_k dy[q0 + k + j][co] * x[q0 + k + 3 + sx][ci]  ->  tap
//                 offset (sx + 3 - j).  A 1x13 kernel needs 4 such tap groups (sx = -6, -2, 2, 6), a 3x3 kernel one per
//                 kernel row (the dy line pairs with the x lines above / at / below it).
// Accumulators (<= 4 groups x 32 columns) stay in tensor memory for the CTA's whole tile range; at the end every CTA
// writes its partial sums to a workspace and a follow-up launch (wgrad_line_reduce_kernel) sums the partials into dW
// (the kernel boundary is the grid-wide barrier).  dbias comes from the dy lines while they sit in shared
// memory.  Warp roles: 0-3 dbias + final read-out (TMEM lane quarter = warp), 4 MMA issuer, 5 x producer, 6 dy producer.
#include "wgrad_reduce.cuh"

#define WG_NS_MAX 8
#define WG_THREADS 224
#define WG_XTRA 8          // extra pixel rows per line buffer: the last strip of a line runs one more K step

struct WgradLineArgs {
  float* dw;             // [32][32][KH][KW]
  float* dbias;          // [32] or null
  float* ws;             // [ctas][S][128][32] partial sums
  unsigned int* counter; // unused (ABI slot of the removed software grid barrier)
  int B, H, W;
  int KL, KA;            // taps along / across the line
  int L, NL;             // line length, lines per image
  int vertical;
  int strips;
  int tiles_total, tiles_per_cta;
  int PX, PD;            // pixel rows per x / dy ring slot
  int NSX, NSD;          // ring slots
  unsigned int xslot_bytes, dslot_bytes;
  int S;                 // accumulator groups
  int dy_c0;             // first of the 32 dy channels this launch contracts (dy may carry more: Cout = 64, 96, 128)
};

struct WSeg { int b, strip, l0, l1, in0, in1; };
__device__ __forceinline__ bool next_wseg(const WgradLineArgs& a, int& t, int t1, WSeg& s) {
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

template <int TKA, int TKL>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_line_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                       const __grid_constant__ CUtensorMap tmd,
                                                                       const WgradLineArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int KA = TKA, KL = TKL;
  constexpr int S = KA == 3 ? 3 : (KL + 3) / 4;
  constexpr int padL = KL >> 1, padA = KA >> 1;
  const int NSX = a.NSX, NSD = a.NSD;
  const uint32_t base_s = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_p = smem_raw + (base_s - smem_u32(smem_raw));
  const uint32_t xring_s = base_s;
  const uint32_t dring_s = xring_s + (uint32_t)NSX * a.xslot_bytes;
  unsigned char* p_dring = base_p + (size_t)NSX * a.xslot_bytes;
  float* s_bias = reinterpret_cast<float*>(p_dring + (size_t)NSD * a.dslot_bytes);     // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 32);      // xfull[8], xempty[8], dfull[8], dempty[8], done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4 * WG_NS_MAX + 1);
  uint32_t* s_started = s_tmem + 1;
  const uint32_t bar_xfull = smem_u32(bars), bar_xempty = bar_xfull + 8 * WG_NS_MAX;
  const uint32_t bar_dfull = bar_xempty + 8 * WG_NS_MAX, bar_dempty = bar_dfull + 8 * WG_NS_MAX;
  const uint32_t bar_done = bar_dempty + 8 * WG_NS_MAX;
  const bool want_bias = a.dbias != nullptr;

  const int t0 = blockIdx.x * a.tiles_per_cta;
  const int t1 = min(t0 + a.tiles_per_cta, a.tiles_total);

  if (tid < 32) s_bias[tid] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < NSX; i++) { mbar_init(bar_xfull + 8 * i, 1); mbar_init(bar_xempty + 8 * i, 1); }
    for (int i = 0; i < NSD; i++) { mbar_init(bar_dfull + 8 * i, 1); mbar_init(bar_dempty + 8 * i, want_bias ? 5 : 1); }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<128>(smem_u32(s_tmem));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < 4) {
    // ===================== dbias from the dy lines in shared memory; final TMEM read-out =====================
    if (want_bias) {
      float bs[32];
#pragma unroll
      for (int i = 0; i < 32; i++) bs[i] = 0.f;
      const int r = 3 + tid;                      // buffer row of pixel (base + tid)
      int t = t0, slot = 0, phase = 0;
      WSeg s;
      while (next_wseg(a, t, t1, s)) {
        for (int l = s.l0; l < s.l1; l++) {
          mbar_wait(bar_dfull + 8 * slot, phase);
          const unsigned char* row = p_dring + (size_t)slot * a.dslot_bytes + (size_t)r * 128;
#pragma unroll
          for (int c = 0; c < 8; c++) {      // 16-byte chunk c lives in 32-byte chunk (c >> 1) ^ (r & 3)
            const float4 v = *reinterpret_cast<const float4*>(row + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4)));
            bs[4 * c] += v.x; bs[4 * c + 1] += v.y; bs[4 * c + 2] += v.z; bs[4 * c + 3] += v.w;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_dempty + 8 * slot);
          if (++slot == NSD) { slot = 0; phase ^= 1; }
        }
      }
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const float v = warp_sum(bs[i]);
        if (lane == 0) atomicAdd(&s_bias[i], v);
      }
    }
    // every MMA of this CTA has completed -> tensor memory holds the partial sums: lane (j, co), column (group, ci)
    mbar_wait(bar_done, 0);
    tc_fence_after();
    float* wsp = a.ws + ((size_t)blockIdx.x * S) * 4096 + (size_t)(warp * 32 + lane) * 32;
    const uint32_t started = *s_started;          // groups that never saw an MMA hold stale tensor memory: they count as zero
#pragma unroll
    for (int g = 0; g < S; g++) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(g * 32), v);
      if (!((started >> (KA == 3 ? g : 0)) & 1u)) {
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < 8; c++)
        reinterpret_cast<float4*>(wsp + (size_t)g * 4096)[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    // both operands MN-major.  N = 32 * (#accumulator groups fed by one MMA): the groups are column blocks of ONE accumulator
    // matrix, so a single MMA per K step covers them all when their B atoms are equidistant in shared memory --
    //   1xk / kx1: the S tap groups read the SAME x line at pixel-row offsets 4g: N-atoms 512 B apart (LBO), overlapping;
    //   3x3      : the three kernel rows read the x lines l-1, l, l+1: N-atoms one ring slot apart, cut where the ring wraps.
    // (At N = 32 the tensor pipe is bound by the 4 KB A read per MMA, ~40 cycles against a 16-cycle math floor: csrc/conv_tma.cu.)
    const uint32_t idesc0 = umma_idesc_tf32(128, 0, 1, 1);
    // MN-major SWIZZLE_128B_BASE32B: 32 channels = one 128-byte row; K group = 4 pixel rows = 512 B (SBO); the four
    // 32-row M groups of A are 128 B (one pixel row) apart (LBO) -- they overlap on purpose.
    const uint64_t hi_a = (uint64_t)(uint32_t)(umma_desc(0u, 128u, 512u, 1u, 0u) >> 32) << 32;
    const uint64_t hi_b = hi_a;
    // the LBO field sits in the LOW descriptor word (bits 16-29), next to the start address the K steps advance
    const uint32_t a_lo0 = (uint32_t)umma_desc(dring_s, 128u, 512u, 1u, 0u);
    const uint32_t b_lo0 = (uint32_t)umma_desc(xring_s, (KA == 3) ? a.xslot_bytes : 512u, 512u, 1u, 0u);
    const uint32_t xslot16 = a.xslot_bytes >> 4, dslot16 = a.dslot_bytes >> 4;
    int t = t0;
    int xw_slot = 0, xw_phase = 0, xwaited = 0, seq_base = 0;
    int dslot = 0, dphase = 0;
    uint32_t started = 0;                                           // bit g: accumulator group g holds data
    WSeg s;
    while (next_wseg(a, t, t1, s)) {
      int cur = (seq_base + (s.l0 - padA) - s.in0) % NSX;          // ring slot of x line (l - padA)
      if (cur < 0) cur += NSX;
      const bool extra = (s.strip == a.strips - 1);                 // the last strip of a line runs a 17th K step
      for (int l = s.l0; l < s.l1; l++) {
        const int need = seq_base + (min(l + padA, s.in1) - s.in0);
        while (xwaited <= need) {
          mbar_wait(bar_xfull + 8 * xw_slot, xw_phase);
          xwaited++;
          if (++xw_slot == NSX) { xw_slot = 0; xw_phase ^= 1; }
        }
        mbar_wait(bar_dfull + 8 * dslot, dphase);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + (uint32_t)dslot * dslot16;
        if (KA == 3) {
          // group = kernel row ka; B row offset: x pixel = q + 3 + sx with sx = -padL = -1 -> buffer row (k + 3 + sx + padL) = k + 3
          int ka = max(0, padA - l), ka_hi = min(KA - 1, a.NL - 1 - l + padA);
          while (ka <= ka_hi) {
            int sl = cur + ka;
            if (sl >= NSX) sl -= NSX;
            const uint32_t st0 = (started >> ka) & 1u;
            int n = 1;
            while (ka + n <= ka_hi && sl + n < NSX && ((started >> (ka + n)) & 1u) == st0) n++;
            if (elect_one()) {
              const uint32_t d_tmem = tmem_base + (uint32_t)(ka * 32);
              const uint32_t idesc = idesc0 | ((uint32_t)(n * 4) << 17);
              const uint32_t b_lo = b_lo0 + (uint32_t)sl * xslot16 + 8u * 3u;
              tc_mma_tf32(d_tmem, hi_a | a_lo, hi_b | b_lo, idesc, st0);
#pragma unroll
              for (int kk = 1; kk < 16; kk++) tc_mma_tf32(d_tmem, hi_a | (a_lo + 64u * kk), hi_b | (b_lo + 64u * kk), idesc, 1u);
              if (extra) tc_mma_tf32(d_tmem, hi_a | (a_lo + 64u * 16u), hi_b | (b_lo + 64u * 16u), idesc, 1u);
            }
            __syncwarp();
            started |= ((1u << n) - 1u) << ka;
            ka += n;
          }
        } else {
          if (elect_one()) {
            // groups g = 0..S-1: sx = 4g - padL -> buffer row k + 3 + sx + padL = k + 3 + 4g: one MMA, N = 32 S, LBO 512 B
            const uint32_t idesc = idesc0 | ((uint32_t)(S * 4) << 17);
            const uint32_t b_lo = b_lo0 + (uint32_t)cur * xslot16 + 8u * 3u;
            tc_mma_tf32(tmem_base, hi_a | a_lo, hi_b | b_lo, idesc, started & 1u);
#pragma unroll
            for (int kk = 1; kk < 16; kk++) tc_mma_tf32(tmem_base, hi_a | (a_lo + 64u * kk), hi_b | (b_lo + 64u * kk), idesc, 1u);
            if (extra) tc_mma_tf32(tmem_base, hi_a | (a_lo + 64u * 16u), hi_b | (b_lo + 64u * 16u), idesc, 1u);
          }
          __syncwarp();
          started |= 1u;
        }
        if (elect_one()) {
          tc_commit(bar_dempty + 8 * dslot);
          if (l - padA >= s.in0 && l + 1 < s.l1) tc_commit(bar_xempty + 8 * cur);
        }
        __syncwarp();
        if (++cur == NSX) cur = 0;
        if (++dslot == NSD) { dslot = 0; dphase ^= 1; }
      }
      if (elect_one()) {
        int sl = cur - 1;
        if (sl < 0) sl += NSX;
        for (int il = s.l1 - 1 - padA; il <= s.in1; il++) {
          if (il >= s.in0) tc_commit(bar_xempty + 8 * sl);
          if (++sl == NSX) sl = 0;
        }
      }
      __syncwarp();
      seq_base += s.in1 - s.in0 + 1;
    }
    if (elect_one()) {
      *s_started = started;
      tc_commit(bar_done);          // arrives after every MMA above has completed; the store above is ordered before it
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================== x producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmx);
      int t = t0, slot = 0, phase = 1;
      WSeg s;
      const uint32_t bytes = (uint32_t)a.PX * 128u;
      while (next_wseg(a, t, t1, s)) {
        // buffer row 0 = pixel (base - 3 - padL): row (k + 3 + sx + padL) then holds x pixel (base + k + sx)
        const int p0 = s.strip * 128 - 3 - padL;
        for (int il = s.in0; il <= s.in1; il++) {
          mbar_wait(bar_xempty + 8 * slot, phase);
          const uint32_t dst = xring_s + (uint32_t)slot * a.xslot_bytes;
          mbar_expect_tx(bar_xfull + 8 * slot, bytes);
          if (a.vertical) tma_load_4d(dst, &tmx, 0, il, p0, s.b, bar_xfull + 8 * slot);
          else tma_load_4d(dst, &tmx, 0, p0, il, s.b, bar_xfull + 8 * slot);
          if (++slot == NSX) { slot = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== dy producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmd);
      int t = t0, slot = 0, phase = 1;
      WSeg s;
      const uint32_t bytes = (uint32_t)a.PD * 128u;
      while (next_wseg(a, t, t1, s)) {
        // buffer row 0 = pixel (base - 3): M group j starts at row j, i.e. pairs dy pixel (base + k - 3 + j)
        const int p0 = s.strip * 128 - 3;
        for (int l = s.l0; l < s.l1; l++) {
          mbar_wait(bar_dempty + 8 * slot, phase);
          const uint32_t dst = dring_s + (uint32_t)slot * a.dslot_bytes;
          mbar_expect_tx(bar_dfull + 8 * slot, bytes);
          if (a.vertical) tma_load_4d(dst, &tmd, a.dy_c0, l, p0, s.b, bar_dfull + 8 * slot);
          else tma_load_4d(dst, &tmd, a.dy_c0, p0, l, s.b, bar_dfull + 8 * slot);
          if (++slot == NSD) { slot = 0; phase ^= 1; }
        }
      }
    }
  }

  // ---- the partial sums are in the workspace; wgrad_line_reduce_kernel (next launch on the stream) folds them into dW.
  // (An earlier version reduced here after a software grid barrier: unsafe when two such kernels run on parallel graph
  // branches, and as a cooperative launch it has to wait for an empty GPU.)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc<128>(tmem_base);
  if (want_bias && tid < 32) atomicAdd(a.dbias + tid, s_bias[tid]);
}

__global__ void __launch_bounds__(256) wgrad_line_reduce_kernel(const float* __restrict__ ws, int nparts, int S, int KA, int KL,
                                                                float* dw) {
  wgrad_line_reduce_body(ws, nparts, S, KA, KL, dw, blockIdx.x);
}

template <int TKA, int TKL>
static void launch_wgrad_line(const CUtensorMap& tmx, const CUtensorMap& tmd, const WgradLineArgs& a, int ctas, size_t smem, cudaStream_t st,
                              bool reduce_now) {
  cudaFuncSetAttribute(wgrad_line_tma_kernel<TKA, TKL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  wgrad_line_tma_kernel<TKA, TKL><<<ctas, WG_THREADS, smem, st>>>(tmx, tmd, a);
  if (reduce_now) {
    wgrad_line_reduce_kernel<<<wgrad_line_reduce_blocks(a.S), 256, 0, st>>>(a.ws, ctas, a.S, a.KA, a.KL, a.dw);
    tcct_count_launch();
  }
}

extern "C" int tcct_wgrad_tma_supported(int H, int W, int Cin, int Cout, int KH, int KW) {
  if (Cin != 32 || Cout % 32 != 0 || Cout < 32 || Cout > 128) return 0;
  const bool ok = (KH == 3 && KW == 3) || (KH == 1 && (KW == 13 || KW == 11)) || (KW == 1 && (KH == 13 || KH == 11));
  if (!ok) return 0;
  const int L = (KW == 1) ? H : W;
  if (L % 128 != 0) return 0;
  return tcct_tensor_map_encoder() != nullptr ? 1 : 0;
}

static int wgrad_tma_plan(int B, int H, int W, int KH, int KW, WgradLineArgs& a) {
  a.B = B; a.H = H; a.W = W;
  a.vertical = (KW == 1) ? 1 : 0;
  if (a.vertical) { a.KL = KH; a.KA = 1; a.L = H; a.NL = W; }
  else { a.KL = KW; a.KA = KH; a.L = W; a.NL = H; }
  a.strips = a.L / 128;
  a.tiles_total = B * a.strips * a.NL;
  const int sms = tcct_num_sms();
  a.tiles_per_cta = ceil_div(a.tiles_total, sms);
  a.S = a.KA == 3 ? 3 : (a.KL + 3) / 4;
  a.PD = 128 + 3 + WG_XTRA;
  a.PX = 128 + 3 + 4 * (a.KA == 3 ? 1 : a.S) + WG_XTRA;       // rows up to (k + 3 + 4(S-1)) for k < 136
  a.xslot_bytes = (unsigned int)((a.PX * 128 + 1023) / 1024 * 1024);
  a.dslot_bytes = (unsigned int)((a.PD * 128 + 1023) / 1024 * 1024);
  return ceil_div(a.tiles_total, a.tiles_per_cta);
}

// floats of workspace tcct_wgrad_tma needs (partial sums of every CTA)
extern "C" long long tcct_wgrad_tma_ws_floats(int B, int H, int W, int KH, int KW) {
  WgradLineArgs a;
  const int ctas = wgrad_tma_plan(B, H, W, KH, KW, a);
  return (long long)ctas * a.S * 4096;
}

// dw: PyTorch [Cout][32][KH][KW] (accumulated); dbias [Cout] or null (accumulated); Cout = 32, 64, 96 or 128: one launch
// pair per 32 output channels (the dy tensor map selects them); ws: tcct_wgrad_tma_ws_floats floats (reused by the
// launches); counter: unused (kept in the ABI; earlier versions ran a grid barrier on it).
static int wgrad_tma_launch(const float* x, const float* dy, float* dw, float* dbias, int B, int H, int W, int KH, int KW,
                            int Cout, float* ws, unsigned int* counter, bool reduce_now, void* stream) {
  TCCT_CHECK_ARG(tcct_wgrad_tma_supported(H, W, 32, Cout, KH, KW), "wgrad_tma: unsupported shape %dx%d kernel %dx%d Cout %d", H, W, KH, KW, Cout);
  WgradLineArgs a;
  const int ctas = wgrad_tma_plan(B, H, W, KH, KW, a);
  a.dw = dw; a.dbias = dbias; a.ws = ws; a.counter = counter; a.dy_c0 = 0;
  TCCT_CHECK_ARG(ctas <= tcct_num_sms(), "wgrad_tma: grid exceeds the SM count");
  const size_t fixed = 1024 + 32 * 4 + (4 * WG_NS_MAX + 1) * 8 + 16;
  // rings: dy needs 1 live line, x needs KA; split the rest of shared memory between them
  a.NSD = 3;
  a.NSX = (int)((227 * 1024 - fixed - (size_t)a.NSD * a.dslot_bytes) / a.xslot_bytes);
  if (a.NSX > WG_NS_MAX) a.NSX = WG_NS_MAX;
  if (a.NSX > a.KA + 3) { a.NSX = a.KA + 3; a.NSD = (int)((227 * 1024 - fixed - (size_t)a.NSX * a.xslot_bytes) / a.dslot_bytes); }
  if (a.NSD > WG_NS_MAX) a.NSD = WG_NS_MAX;
  TCCT_CHECK_ARG(a.NSX >= a.KA + 1 && a.NSD >= 2, "wgrad_tma: rings do not fit (%d, %d)", a.NSX, a.NSD);
  const size_t smem = fixed + (size_t)a.NSX * a.xslot_bytes + (size_t)a.NSD * a.dslot_bytes;
  CUtensorMap tmx, tmd;
  const unsigned long long dims[4] = {32ull, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
  const unsigned long long strides[3] = {128ull, (unsigned long long)W * 128ull, (unsigned long long)H * W * 128ull};
  unsigned int box_x[4] = {32u, 1u, 1u, 1u}, box_d[4] = {32u, 1u, 1u, 1u};
  box_x[a.vertical ? 2 : 1] = (unsigned int)a.PX;
  box_d[a.vertical ? 2 : 1] = (unsigned int)a.PD;
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmx, x, 4, dims, strides, box_x, 2), "wgrad_tma: cuTensorMapEncodeTiled failed (x)");
  const unsigned long long dims_d[4] = {(unsigned long long)Cout, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B};
  const unsigned long long strides_d[3] = {(unsigned long long)Cout * 4ull, (unsigned long long)W * Cout * 4ull,
                                           (unsigned long long)H * W * Cout * 4ull};
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmd, dy, 4, dims_d, strides_d, box_d, 2), "wgrad_tma: cuTensorMapEncodeTiled failed (dy)");
  cudaStream_t st = (cudaStream_t)stream;
  const int T = KH * KW;
  const size_t ws_slice = (size_t)ctas * a.S * 4096;
  for (int c0 = 0; c0 < Cout; c0 += 32) {
    a.dy_c0 = c0;
    a.dw = dw ? dw + (size_t)c0 * 32 * T : nullptr;
    a.dbias = dbias ? dbias + c0 : nullptr;
    if (!reduce_now) a.ws = ws + (size_t)(c0 / 32) * ws_slice;      // deferred reduction: every 32-channel slice keeps its own slabs
    if (a.KA == 3) launch_wgrad_line<3, 3>(tmx, tmd, a, ctas, smem, st, reduce_now);
    else if (a.KL == 13) launch_wgrad_line<1, 13>(tmx, tmd, a, ctas, smem, st, reduce_now);
    else launch_wgrad_line<1, 11>(tmx, tmd, a, ctas, smem, st, reduce_now);
  }
  tcct_count_route(TCCT_ROUTE_WGRAD_TMA);
  TCCT_CHECK_LAUNCH("wgrad_tma");
  return TCCT_OK;
}
extern "C" int tcct_wgrad_tma(const float* x, const float* dy, float* dw, float* dbias, int B, int H, int W, int KH, int KW,
                              int Cout, float* ws, unsigned int* counter, void* stream) {
  return wgrad_tma_launch(x, dy, dw, dbias, B, H, W, KH, KW, Cout, ws, counter, true, stream);
}
// First phase only: the partial sums stay in ws (Cout / 32 consecutive regions of tcct_wgrad_tma_ws_floats floats, one per 32 output
// channels) until tcct_wgrad_reduce_batch folds them into dW; dbias is complete after this call.  parts[4] = {partials per region, S, KA, KL}.
extern "C" int tcct_wgrad_tma_partial(const float* x, const float* dy, float* dbias, int B, int H, int W, int KH, int KW, int Cout, float* ws,
                                      int* parts, void* stream) {
  WgradLineArgs a;
  const int ctas = wgrad_tma_plan(B, H, W, KH, KW, a);
  if (parts) { parts[0] = ctas; parts[1] = a.S; parts[2] = a.KA; parts[3] = a.KL; }
  return wgrad_tma_launch(x, dy, nullptr, dbias, B, H, W, KH, KW, Cout, ws, nullptr, false, stream);
}
