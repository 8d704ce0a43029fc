// Weight gradient of the 1x1 convolutions / nn.Linear on large maps as a TMA-fed tcgen05 pipeline:
//   dW[n][k0 + k] += sum_m dy[m][n] * x[m][k]        dbias[n] += sum_m dy[m][n]
// (autograd's weight path for the GEMMs of csrc/gemm_tma.cu; reference call sites listed there).
//
// The contraction runs over PIXELS: both operands are MN-major, read straight from the row-major tiles TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (32 channels of a pixel = one 128-byte row; descriptor layout type 1, K group =
// 4 pixel rows).  A tile is PT pixels: x arrives as K/32 slabs {32 ch, PT px}, dy as N/32 slabs; the slabs of an operand
// lie PT*128 B apart, which is the leading-dimension byte offset between the 32-channel MN groups.  One tcgen05.mma is
// M = 128 (dy channels; groups past N/32 read whatever follows in shared memory and only feed accumulator rows nobody reads),
// N = K_in (<= 256), K = 8 pixels.  The accumulator (128 lanes x K_in columns) stays in tensor memory for the CTA's
// whole tile range; then: partials -> workspace -> grid-wide barrier (grid <= #SMs) -> sliced reduction into dW.
// Warp roles: 0-3 dbias + read-out, 4 MMA issuer, 5 TMA producer.
#include "wgrad_reduce.cuh"

#define WL_NS_MAX 6
#define WL_THREADS 192

struct WgradGemmArgs {
  float* dw;             // row n at dw + n * ld
  float* dbias;          // [N] or null
  float* ws;             // [ctas][128][K]
  unsigned int* counter;
  int M, K, N, ld;
  int PT;                // pixels per tile (128 or 64)
  int tiles;
  int NS;
  unsigned int stage_bytes, x_bytes, slab_bytes, pad_bytes;     // one stage = x slabs | dy slabs
};

template <int TCOLS>
__global__ void __launch_bounds__(WL_THREADS, 1) wgrad_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                       const __grid_constant__ CUtensorMap tmd,
                                                                       const WgradGemmArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = a.NS, K = a.K, N = a.N, PT = a.PT;
  const int kslabs = K >> 5, nslabs = N >> 5;
  const uint32_t base_s = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_p = smem_raw + (base_s - smem_u32(smem_raw));
  float* s_bias = reinterpret_cast<float*>(base_p + (size_t)NS * a.stage_bytes + a.pad_bytes);       // [N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + ((N + 1) & ~1));               // full[6], empty[6], done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * WL_NS_MAX + 1);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * WL_NS_MAX, bar_done = bar_empty + 8 * WL_NS_MAX;
  const bool want_bias = a.dbias != nullptr;

  for (int i = tid; i < N; i += WL_THREADS) s_bias[i] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < NS; i++) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, want_bias ? 5 : 1); }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<TCOLS>(smem_u32(s_tmem));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < 4) {
    // ===================== dbias from the dy slabs in shared memory; final TMEM read-out =====================
    if (want_bias) {
      // thread (row group rs, float4 column c) walks rows rs, rs+16, ... of every dy slab and keeps its column sums in
      // registers for the CTA's whole tile range: no shuffles or atomics per tile (a per-tile warp reduction of all 32
      // columns tripled the kernel time), one 16-way shared-memory reduction at the end
      int slot = 0, phase = 0;
      const int c = tid & 7, rs = tid >> 3;
      float4 acc[4];
#pragma unroll
      for (int sl = 0; sl < 4; sl++) acc[sl] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
        mbar_wait(bar_full + 8 * slot, phase);
#pragma unroll
        for (int sl = 0; sl < 4; sl++) {
          if (sl < nslabs) {
            const unsigned char* slab = base_p + (size_t)slot * a.stage_bytes + a.x_bytes + (size_t)sl * a.slab_bytes;
            for (int r = rs; r < PT; r += 16) {
              const float4 v = *reinterpret_cast<const float4*>(slab + (size_t)r * 128 + ((((c >> 1) ^ (r & 3)) << 5) | ((c & 1) << 4)));
              acc[sl].x += v.x; acc[sl].y += v.y; acc[sl].z += v.z; acc[sl].w += v.w;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
        if (++slot == NS) { slot = 0; phase ^= 1; }
      }
#pragma unroll
      for (int sl = 0; sl < 4; sl++) {
        if (sl < nslabs) {
          atomicAdd(&s_bias[sl * 32 + 4 * c], acc[sl].x); atomicAdd(&s_bias[sl * 32 + 4 * c + 1], acc[sl].y);
          atomicAdd(&s_bias[sl * 32 + 4 * c + 2], acc[sl].z); atomicAdd(&s_bias[sl * 32 + 4 * c + 3], acc[sl].w);
        }
      }
    }
    mbar_wait(bar_done, 0);
    tc_fence_after();
    float* wsp = a.ws + ((size_t)blockIdx.x * 128 + (size_t)(warp * 32 + lane)) * K;
    if (warp * 32 < N) {                       // accumulator rows of this lane quarter are real output channels
      for (int ch = 0; ch < kslabs; ch++) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ch * 32), v);
#pragma unroll
        for (int c = 0; c < 8; c++)
          reinterpret_cast<float4*>(wsp + ch * 32)[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_tf32(128, K, 1, 1);
    const uint32_t lbo = a.slab_bytes;
    const uint64_t desc_hi = (uint64_t)(uint32_t)(umma_desc(0u, lbo, 512u, 1u, 0u) >> 32) << 32;
    const uint32_t lo0 = (uint32_t)umma_desc(base_s, lbo, 512u, 1u, 0u);
    const int ksteps = PT >> 3;
    int slot = 0, phase = 0;
    uint32_t accum = 0;
    for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
      mbar_wait(bar_full + 8 * slot, phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_lo = lo0 + (uint32_t)slot * (a.stage_bytes >> 4);
        const uint32_t a_lo = b_lo + (a.x_bytes >> 4);
        for (int kk = 0; kk < ksteps; kk++) {
          tc_mma_tf32(tmem_base, desc_hi | (a_lo + 64u * kk), desc_hi | (b_lo + 64u * kk), idesc, accum);
          accum = 1;
        }
        tc_commit(bar_empty + 8 * slot);
      }
      __syncwarp();
      accum = 1;
      if (++slot == NS) { slot = 0; phase ^= 1; }
    }
    if (elect_one()) tc_commit(bar_done);
    __syncwarp();
  } else {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmx); tma_prefetch_desc(&tmd);
      int slot = 0, phase = 1;
      const uint32_t bytes = (uint32_t)(kslabs + nslabs) * a.slab_bytes;
      for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
        mbar_wait(bar_empty + 8 * slot, phase);
        const uint32_t st = base_s + (uint32_t)slot * a.stage_bytes;
        mbar_expect_tx(bar_full + 8 * slot, bytes);
        for (int s = 0; s < kslabs; s++) tma_load_2d(st + (uint32_t)s * a.slab_bytes, &tmx, s * 32, tile * PT, bar_full + 8 * slot);
        for (int s = 0; s < nslabs; s++) tma_load_2d(st + a.x_bytes + (uint32_t)s * a.slab_bytes, &tmd, s * 32, tile * PT, bar_full + 8 * slot);
        if (++slot == NS) { slot = 0; phase ^= 1; }
      }
    }
  }

  // ---- the partial sums are in the workspace; wgrad_gemm_reduce_kernel (next launch on the stream) folds them into dW
  // (no software grid barrier: see csrc/wgrad_tma.cu)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc<TCOLS>(tmem_base);
  if (want_bias)
    for (int i = tid; i < N; i += WL_THREADS) atomicAdd(a.dbias + i, s_bias[i]);
}

__global__ void __launch_bounds__(WL_THREADS) wgrad_gemm_reduce_kernel(const float* __restrict__ ws, int nparts, int N, int K, int ld,
                                                                       float* dw) {
  __shared__ float s_part[(WL_THREADS / 32) * 32];
  wgrad_gemm_reduce_body<WL_THREADS / 32>(ws, nparts, N, K, ld, dw, blockIdx.x, gridDim.x, s_part);
}

static int wgrad_gemm_plan(long long M, int K, int N, WgradGemmArgs& a) {
  a.M = (int)M; a.K = K; a.N = N;
  const size_t fixed = 1024 + (size_t)(N + 2) * 4 + (2 * WL_NS_MAX + 1) * 8 + 16;
  a.PT = 128;
  for (;;) {
    a.slab_bytes = (unsigned int)a.PT * 128u;
    a.x_bytes = (unsigned int)(K / 32) * a.slab_bytes;
    // a stage holds the K/32 x slabs and the N/32 dy slabs; the M groups of the A operand past N/32 read on into the
    // next stage (any mapped shared memory will do: those accumulator rows are never read), hence one pad after the last
    a.stage_bytes = a.x_bytes + (unsigned int)(N / 32) * a.slab_bytes;
    a.pad_bytes = (unsigned int)(4 - N / 32) * a.slab_bytes;
    a.NS = (int)((227 * 1024 - fixed - a.pad_bytes) / a.stage_bytes);
    if (a.NS >= 2 || a.PT == 64) break;
    a.PT = 64;
  }
  if (a.NS > WL_NS_MAX) a.NS = WL_NS_MAX;
  if (a.NS < 2 || M % a.PT != 0) return 0;
  a.tiles = (int)(M / a.PT);
  // at least four pixel tiles per CTA: every CTA leaves a [128][K] partial-sum slab behind, and on the small maps those
  // slabs (and the reduce kernel that reads them) outweigh the operands
  int ctas = tcct_num_sms();
  if (ctas > a.tiles / 4) ctas = a.tiles / 4;
  if (ctas < 1) ctas = 1;
  return ctas;
}

extern "C" int tcct_wgrad_gemm_tma_supported(long long M, int K, int N) {
  if (M <= 0 || M % 128 != 0 || M >= (1ll << 31) || M < 128ll * 64) return 0;
  if (K % 32 != 0 || N % 32 != 0 || K < 32 || N < 32 || N > 128 || K > 256) return 0;
  WgradGemmArgs a;
  if (wgrad_gemm_plan(M, K, N, a) == 0) return 0;
  return tcct_tensor_map_encoder() != nullptr ? 1 : 0;
}

extern "C" long long tcct_wgrad_gemm_tma_ws_floats(long long M, int K, int N) {
  WgradGemmArgs a;
  const int ctas = wgrad_gemm_plan(M, K, N, a);
  return (long long)ctas * 128 * K;
}

// dw: row n of the gradient at dw + n*ld (ld = row stride of the weight tensor; dw already points at column k0);
// dbias [N] or null; ws: tcct_wgrad_gemm_tma_ws_floats floats; counter: one zeroed 32-bit word.
static int wgrad_gemm_tma_launch(const float* x, const float* dy, float* dw, float* dbias, long long M, int K, int N, int ld,
                                 float* ws, unsigned int* counter, bool reduce_now, void* stream) {
  TCCT_CHECK_ARG(tcct_wgrad_gemm_tma_supported(M, K, N), "wgrad_gemm_tma: unsupported shape M=%lld K=%d N=%d", M, K, N);
  WgradGemmArgs a;
  const int ctas = wgrad_gemm_plan(M, K, N, a);
  a.dw = dw; a.dbias = dbias; a.ws = ws; a.counter = counter; a.ld = ld;
  const size_t smem = 1024 + (size_t)a.NS * a.stage_bytes + a.pad_bytes + (size_t)(N + 2) * 4 + (2 * WL_NS_MAX + 1) * 8 + 16;
  CUtensorMap tmx, tmd;
  const unsigned long long dx[2] = {(unsigned long long)K, (unsigned long long)M}, sx[1] = {(unsigned long long)K * 4ull};
  const unsigned long long dd[2] = {(unsigned long long)N, (unsigned long long)M}, sd[1] = {(unsigned long long)N * 4ull};
  const unsigned int box[2] = {32u, (unsigned int)a.PT};
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmx, x, 2, dx, sx, box, 2), "wgrad_gemm_tma: cuTensorMapEncodeTiled failed (x)");
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmd, dy, 2, dd, sd, box, 2), "wgrad_gemm_tma: cuTensorMapEncodeTiled failed (dy)");
  cudaStream_t st = (cudaStream_t)stream;
#define WL_LAUNCH(TC)                                                                                              \
  do {                                                                                                             \
    cudaFuncSetAttribute(wgrad_gemm_tma_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
    wgrad_gemm_tma_kernel<TC><<<ctas, WL_THREADS, smem, st>>>(tmx, tmd, a);                                         \
  } while (0)
  if (K <= 32) WL_LAUNCH(32);
  else if (K <= 64) WL_LAUNCH(64);
  else if (K <= 128) WL_LAUNCH(128);
  else WL_LAUNCH(256);
#undef WL_LAUNCH
  if (reduce_now) {
    wgrad_gemm_reduce_kernel<<<wgrad_gemm_reduce_blocks(N, K, tcct_num_sms()), WL_THREADS, 0, st>>>(ws, ctas, N, K, ld, dw);
    tcct_count_launch();
  }
  tcct_count_route(TCCT_ROUTE_WGRAD_GEMM_TMA);
  TCCT_CHECK_LAUNCH("wgrad_gemm_tma");
  return TCCT_OK;
}
extern "C" int tcct_wgrad_gemm_tma(const float* x, const float* dy, float* dw, float* dbias, long long M, int K, int N, int ld,
                                   float* ws, unsigned int* counter, void* stream) {
  return wgrad_gemm_tma_launch(x, dy, dw, dbias, M, K, N, ld, ws, counter, true, stream);
}
// First phase only (see tcct_wgrad_tma_partial): returns the number of partial slabs left in ws through *parts.
extern "C" int tcct_wgrad_gemm_tma_partial(const float* x, const float* dy, float* dbias, long long M, int K, int N, float* ws, int* parts,
                                           void* stream) {
  WgradGemmArgs a;
  const int ctas = wgrad_gemm_plan(M, K, N, a);
  if (parts) parts[0] = ctas;
  return wgrad_gemm_tma_launch(x, dy, nullptr, dbias, M, K, N, 0, ws, nullptr, false, stream);
}
