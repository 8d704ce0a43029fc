// 32->32 channel spatial convolutions (3x3, 1xk, kx1) on the 5th-generation tensor cores:
// tcgen05.mma kind::tf32, accumulators in tensor memory, operands staged in shared memory in the canonical
// no-swizzle K-major layout, warp-specialised (loaders / one MMA-issuing thread / TMEM->register epilogue)
// with mbarrier pipelines.  Used for the full- and half-resolution stages of CrossResNet (94 % of the conv FLOPs,
// task1/nets/tcct.py:803-828) in the forward and data-gradient directions; smaller maps use conv_mma.cu.
//
// Work unit: one "line tile" = 128 consecutive pixels of one line (an image row, or an image column for kx1
// kernels so that the taps always run ALONG the line).  M = 128 pixels, N = 32 output channels, K = 32 input
// channels per tap, i.e. taps x 4 MMAs of K = 8 per line tile.  A line buffer holds its pixels as
// [8 channel-chunks of 16 B][P pixels][16 B] so that a tap shift is a 16-byte shift of the descriptor start
// address and every tap of every kernel shape reads the SAME staged copy of the input (read once from HBM/L2).
// 3x3 kernels march down the image with a ring of line buffers (each input row is staged once per strip).
#include "common.cuh"

#define UM_NS_MAX 8
#define UM_EPI_WARPS 4
#define UM_LOAD_WARPS 4
#define UM_THREADS ((UM_EPI_WARPS + 1 + UM_LOAD_WARPS) * 32)

struct RowConvArgs {
  const float* x;        // [B,H,W,32]
  const float* wu;       // packed [tap][chunk 8][n 32][4]  (tf32-rounded)
  const float* bias;     // [32] or null
  float* y;              // [B,H,W,32]
  double* stats;         // [64] or null
  int stats_act;
  int B, H, W;
  int KL, KA;            // taps along / across the line
  int L, NL;             // line length, lines per image
  long long lstride, pstride;   // element strides between lines / between pixels of a line
  int strips;            // L / 128
  int tiles_total, tiles_per_cta;
  int P;                 // pixel slots per line buffer (odd): 128 + KL - 1 (+1)
  int NS;                // ring slots
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned long long spins = 0;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1ull << 21)) __trap();      // a lost arrival must not hang the GPU
  } while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// shared-memory matrix descriptor, no swizzle: start address, leading/stride byte offsets (all multiples of 16 B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::tf32, fp32 accumulate, K-major A and B (or MN-major when the flags are set), M = 128
__device__ __forceinline__ uint32_t idesc_tf32(int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((128u >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// ---- the kernel ------------------------------------------------------------------------------------------
// Segment bookkeeping shared by all roles: the CTA owns tiles [t0, t1); a segment is a maximal run of tiles in
// the same (image, strip); within a segment output lines [l0, l1) need input lines [in0, in1].
struct Seg { int b, strip, l0, l1, in0, in1; };
__device__ __forceinline__ bool next_seg(const RowConvArgs& a, int& t, int t1, Seg& s) {
  if (t >= t1) return false;
  const int per_img = a.strips * a.NL;
  s.b = t / per_img;
  const int r = t - s.b * per_img;
  s.strip = r / a.NL;
  s.l0 = r - s.strip * a.NL;
  const int room = a.NL - s.l0;
  const int n = min(room, t1 - t);
  s.l1 = s.l0 + n;
  const int pad = a.KA >> 1;
  s.in0 = max(s.l0 - pad, 0);
  s.in1 = min(s.l1 - 1 + pad, a.NL - 1);
  t += n;
  return true;
}

__global__ void __launch_bounds__(UM_THREADS, 1) conv_row_umma_kernel(const RowConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.KL * a.KA;
  const int P = a.P, NS = a.NS;
  const uint32_t slot_bytes = (uint32_t)(8 * P * 16);
  // carve-up: weights | ring | bias | barriers | tmem ptr
  unsigned char* p_w = smem_raw;
  unsigned char* p_ring = p_w + (size_t)T * 4096;
  float* s_bias = reinterpret_cast<float*>(p_ring + (size_t)NS * slot_bytes);
  float* s_stats = s_bias + 32;                                   // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stats + 64);     // full[NS], empty[NS], tfull[2], tempty[2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * UM_NS_MAX + 4);
  const uint32_t w_s = smem_u32(p_w), ring_s = smem_u32(p_ring);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * UM_NS_MAX;
  const uint32_t bar_tfull = bar_empty + 8 * UM_NS_MAX, bar_tempty = bar_tfull + 16;

  const int t0 = blockIdx.x * a.tiles_per_cta;
  const int t1 = min(t0 + a.tiles_per_cta, a.tiles_total);

  // ---- one-time setup
  for (int i = tid; i < T * 1024; i += UM_THREADS) reinterpret_cast<float*>(p_w)[i] = a.wu[i];
  if (tid < 32) s_bias[tid] = a.bias ? a.bias[tid] : 0.f;
  if (tid < 64) s_stats[tid] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < NS; i++) { mbar_init(bar_full + 8 * i, UM_LOAD_WARPS * 32); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; i++) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, UM_EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == UM_EPI_WARPS) {     // the MMA warp owns the tensor-memory allocation (64 columns: 2 accumulators)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();            // the weights were written through the generic proxy, the MMA reads through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int padL = a.KL >> 1, padA = a.KA >> 1;

  if (warp < UM_EPI_WARPS) {
    // ===================== epilogue: TMEM -> registers -> (+bias, statistics) -> global =====================
    float st_sum[32], st_sq[32];
#pragma unroll
    for (int i = 0; i < 32; i++) st_sum[i] = st_sq[i] = 0.f;
    const int m = warp * 32 + lane;              // pixel of the line tile == TMEM lane
    int t = t0, out_cnt = 0;
    Seg s;
    while (next_seg(a, t, t1, s)) {
      for (int l = s.l0; l < s.l1; l++, out_cnt++) {
        const int acc = out_cnt & 1;
        mbar_wait(bar_tfull + 8 * acc, (out_cnt >> 1) & 1);
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 32), v);
        tc_fence_before();
        mbar_arrive(bar_tempty + 8 * acc);
        float* dst = a.y + (size_t)s.b * a.H * a.W * 32 + (size_t)l * a.lstride + (size_t)(s.strip * 128 + m) * a.pstride;
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] += s_bias[i];
#pragma unroll
        for (int i = 0; i < 8; i++)
          reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        if (a.stats) {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const float u = act_fwd(a.stats_act, v[i]);
            st_sum[i] += u; st_sq[i] += u * u;
          }
        }
      }
    }
    if (a.stats) {
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const float su = warp_sum(st_sum[i]), sq = warp_sum(st_sq[i]);
        if (lane == 0) { atomicAdd(&s_stats[i], su); atomicAdd(&s_stats[32 + i], sq); }
      }
    }
  } else if (warp == UM_EPI_WARPS) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = idesc_tf32(32, 0, 0);
      int t = t0, out_cnt = 0, waited = 0, seq_base = 0;
      Seg s;
      while (next_seg(a, t, t1, s)) {
        for (int l = s.l0; l < s.l1; l++, out_cnt++) {
          const int need = seq_base + (min(l + padA, s.in1) - s.in0);       // newest input line this output line reads
          while (waited <= need) { mbar_wait(bar_full + 8 * (waited % NS), (waited / NS) & 1); waited++; }
          const int acc = out_cnt & 1;
          mbar_wait(bar_tempty + 8 * acc, ((out_cnt >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 32);
          uint32_t accum = 0;
          for (int ka = 0; ka < a.KA; ka++) {
            const int il = l + ka - padA;
            if (il < 0 || il >= a.NL) continue;                              // zero padding across lines
            const uint32_t slot = (uint32_t)((seq_base + il - s.in0) % NS);
            const uint32_t abase = ring_s + slot * slot_bytes;
            for (int kl = 0; kl < a.KL; kl++) {
              const uint32_t wtap = w_s + (uint32_t)(ka * a.KL + kl) * 4096u;
#pragma unroll
              for (int ks = 0; ks < 4; ks++) {
                const uint64_t ad = smem_desc(abase + (uint32_t)(2 * ks * P + kl) * 16u, (uint32_t)P * 16u, 128u);
                const uint64_t bd = smem_desc(wtap + (uint32_t)(2 * ks) * 512u, 512u, 128u);
                tc_mma_tf32(d_tmem, ad, bd, idesc, accum);
                accum = 1;
              }
            }
          }
          tc_commit(bar_tfull + 8 * acc);
          const int dead = l - padA;                                         // input line no later output line needs
          if (dead >= s.in0 && l + 1 < s.l1) tc_commit(bar_empty + 8 * ((seq_base + dead - s.in0) % NS));
        }
        // end of segment: release every line still held
        for (int il = max(s.l1 - 1 - padA, s.in0); il <= s.in1; il++) tc_commit(bar_empty + 8 * ((seq_base + il - s.in0) % NS));
        seq_base += s.in1 - s.in0 + 1;
      }
    }
  } else {
    // ===================== loaders: global -> shared line buffers (cp.async, zero fill outside the line) =====================
    const int ltid = tid - (UM_EPI_WARPS + 1) * 32;          // 0..127
    const int nload = UM_LOAD_WARPS * 32;
    int t = t0, seq = 0, pending = -1;
    Seg s;
    while (next_seg(a, t, t1, s)) {
      for (int il = s.in0; il <= s.in1; il++, seq++) {
        const int slot = seq % NS;
        mbar_wait(bar_empty + 8 * slot, ((seq / NS) & 1) ^ 1);
        const uint32_t sbase = ring_s + (uint32_t)slot * slot_bytes;
        const float* line = a.x + (size_t)s.b * a.H * a.W * 32 + (size_t)il * a.lstride;
        for (int idx = ltid; idx < P * 8; idx += nload) {
          const int q = idx >> 3, c = idx & 7;
          const int p = s.strip * 128 + q - padL;
          const bool ok = p >= 0 && p < a.L;
          const float* src = ok ? line + (size_t)p * a.pstride + c * 4 : a.x;
          cp_async16(sbase + (uint32_t)(c * P + q) * 16u, src, ok ? 16 : 0);
        }
        cp_async_commit();
        if (pending >= 0) {
          cp_async_wait<1>();
          fence_proxy_async();
          mbar_arrive(bar_full + 8 * pending);
        }
        pending = slot;
      }
    }
    if (pending >= 0) {
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(bar_full + 8 * pending);
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (a.stats && tid < 64) atomicAdd(a.stats + tid, (double)s_stats[tid]);
  if (warp == UM_EPI_WARPS)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64));
}

// 1 if this shape runs on the tcgen05 path
extern "C" int tcct_conv_umma_supported(int H, int W, int Cin, int Cout, int KH, int KW) {
  if (Cin != 32 || Cout != 32) return 0;
  if (!((KH == 3 && KW == 3) || (KH == 1 && KW >= 3 && KW <= 13 && (KW & 1)) || (KW == 1 && KH >= 3 && KH <= 13 && (KH & 1)))) return 0;
  const int L = (KW == 1) ? H : W;
  return (L % 128 == 0) ? 1 : 0;
}

// wu: weights packed by tcct_pack_weights with fmt = 1 ([tap][chunk][n][4], tf32-rounded)
extern "C" int tcct_conv2d_umma(const float* x, const float* wu, const float* bias, float* y, int B, int H, int W, int KH,
                                int KW, double* stats, int stats_act, void* stream) {
  TCCT_CHECK_ARG(tcct_conv_umma_supported(H, W, 32, 32, KH, KW), "conv2d_umma: unsupported shape %dx%d kernel %dx%d", H, W, KH, KW);
  RowConvArgs a;
  a.x = x; a.wu = wu; a.bias = bias; a.y = y; a.stats = stats; a.stats_act = stats_act;
  a.B = B; a.H = H; a.W = W;
  if (KW == 1) {            // k x 1: lines are image columns
    a.KL = KH; a.KA = 1; a.L = H; a.NL = W; a.lstride = 32; a.pstride = (long long)W * 32;
  } else {
    a.KL = KW; a.KA = KH; a.L = W; a.NL = H; a.lstride = (long long)W * 32; a.pstride = 32;
  }
  a.strips = a.L / 128;
  a.tiles_total = B * a.strips * a.NL;
  const int sms = tcct_num_sms();
  a.tiles_per_cta = ceil_div(a.tiles_total, sms);
  const int ctas = ceil_div(a.tiles_total, a.tiles_per_cta);
  a.P = (128 + a.KL - 1) | 1;
  a.NS = a.KA == 3 ? 6 : 4;
  const size_t smem = (size_t)a.KL * a.KA * 4096 + (size_t)a.NS * 8 * a.P * 16 + 96 * 4 + (2 * UM_NS_MAX + 4) * 8 + 16;
  TCCT_CHECK_ARG(smem <= 227 * 1024, "conv2d_umma: shared memory budget exceeded (%zu B)", smem);
  cudaFuncSetAttribute(conv_row_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  conv_row_umma_kernel<<<ctas, UM_THREADS, smem, (cudaStream_t)stream>>>(a);
  TCCT_CHECK_LAUNCH("conv2d_umma");
  return TCCT_OK;
}
