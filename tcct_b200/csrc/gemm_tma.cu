// 1x1 convolutions / nn.Linear over pixels on large maps as a TMA-fed tcgen05 pipeline:
//   y[M][N] = x[M][K] . W[N][K]^T (+ bias)   [; y = res + res_scale[sample] * y]   (+ per-channel sums of stats_act(y))
// Reference call sites: Conv2d_BN (task1/nets/tcct.py:55-97), DWConv2d_BN.pwconv (99-147), Mlp (29-53), the aggregate /
// tran_vit / tran_cnn / t32x / MPUpBlock.post 1x1 convs (604-616, 966-991, 887-914); the data gradient is the same GEMM
// with the transposed pack.
//
// A tile is 128 consecutive pixels x all N output channels.  x arrives as K/32 TMA boxes {32 ch, 128 px} with 128-byte
// swizzle (the canonical K-major SWIZZLE_128B operand: one 128-byte row = 32 channels of a pixel); the whole weight
// matrix sits in shared memory in the same row format, slab by slab.  Per slab 4 tcgen05.mma kind::tf32 (M128, N, K8);
// accumulators (N columns) are double buffered in tensor memory; the epilogue walks the tile in 32-column chunks:
// TMEM -> registers -> (+bias, residual) -> swizzled staging tile -> TMA store, and takes the BatchNorm sums column-wise
// from the staging tile.  The kernel is HBM-bound (AI = K*N/(2(K+N)) FLOP/B <= 40 at K = N = 160).
// Warp roles: 0-3 epilogue (TMEM lane quarter = warp), 4 MMA issuer, 5 TMA producer.
#include "tma.cuh"

#define GT_NS_MAX 8
#define GT_THREADS 192
#define GT_TILE_BYTES 16384

struct GemmTmaArgs {
  const float* wu;       // packed fmt 3: [slab][n][chunk ^ (n & 7)][4] (tf32-rounded), N*128 B per slab
  const float* bias;     // [N] or null
  const float* res;      // [M][N] or null
  const float* res_scale;   // [samples] or null
  double* stats;         // [2N] or null
  int stats_act;
  int M, K, N;
  int px_per_sample;
  int tiles;             // M / 128
  int NS;                // ring slots
  int ncol;              // TMEM columns per accumulator (N)
  const float* aux;      // MODE 2: [M][N] pre-activation saved by the forward; the result is multiplied by GELU'(aux)
};

// MODE 0: plain.  MODE 1 (Mlp.fc1, tcct.py:29-53): y receives the pre-activation and a second tensor (tmy2) GELU(y) -- the
// activation pass of the MLP rides in the epilogue.  MODE 2 (data gradient of Mlp.fc2): the result is multiplied by GELU'(aux),
// aux = the pre-activation saved by MODE 1 -- the activation's backward pass rides in the epilogue.
template <int TCOLS, int MODE>
__global__ void __launch_bounds__(GT_THREADS, 1) gemm_tma_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                 const __grid_constant__ CUtensorMap tmy,
                                                                 const __grid_constant__ CUtensorMap tmy2,
                                                                 const GemmTmaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = a.NS, N = a.N, nslab = a.K >> 5, nchunk = a.N >> 5;
  const uint32_t base_s = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* base_p = smem_raw + (base_s - smem_u32(smem_raw));
  const uint32_t ring_s = base_s;
  const uint32_t stage_s = ring_s + (uint32_t)NS * GT_TILE_BYTES;
  const uint32_t w_s = stage_s + 2u * GT_TILE_BYTES;
  unsigned char* p_stage = base_p + (size_t)NS * GT_TILE_BYTES;
  unsigned char* p_w = p_stage + 2 * GT_TILE_BYTES;
  const uint32_t wbytes = (uint32_t)a.K * (uint32_t)N * 4u;
  float* s_bias = reinterpret_cast<float*>(p_w + wbytes);          // [N]
  float* s_stats = s_bias + N;                                     // [2N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stats + 2 * N);   // full[8], empty[8], tfull[2], tempty[2], wfull
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * GT_NS_MAX + 5);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * GT_NS_MAX;
  const uint32_t bar_tfull = bar_empty + 8 * GT_NS_MAX, bar_tempty = bar_tfull + 16, bar_w = bar_tempty + 16;

  for (int i = tid; i < N; i += GT_THREADS) s_bias[i] = a.bias ? a.bias[i] : 0.f;
  for (int i = tid; i < 2 * N; i += GT_THREADS) s_stats[i] = 0.f;
  if (tid == 0) {
    for (int i = 0; i < NS; i++) { mbar_init(bar_full + 8 * i, 1); mbar_init(bar_empty + 8 * i, 1); }
    for (int i = 0; i < 2; i++) { mbar_init(bar_tfull + 8 * i, 1); mbar_init(bar_tempty + 8 * i, 128); }
    mbar_init(bar_w, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<TCOLS>(smem_u32(s_tmem));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp < 4) {
    // ===================== epilogue =====================
    const int m = warp * 32 + lane;              // row of the tile == TMEM lane
    const int scol = tid & 31, srow0 = (tid >> 5) * 32;      // column-statistics mapping: column scol, rows srow0..+31
    float st_sum[8], st_sq[8];
#pragma unroll
    for (int i = 0; i < 8; i++) st_sum[i] = st_sq[i] = 0.f;
    int it = 0, cc = 0;                           // cc: running chunk counter (staging buffer = cc & 1)
    for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      mbar_wait(bar_tfull + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      const long long row = (long long)tile * 128 + m;
      float rs = 1.f;
      if (a.res && a.res_scale) rs = a.res_scale[row / a.px_per_sample];
      for (int ch = 0; ch < nchunk; ch++, cc++) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * a.ncol + ch * 32), v);
        if (ch == nchunk - 1) { tc_fence_before(); mbar_arrive(bar_tempty + 8 * acc); }
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] += s_bias[ch * 32 + i];
        if (a.res) {
          const float4* rp = reinterpret_cast<const float4*>(a.res + row * N + ch * 32);
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const float4 r = __ldg(rp + c);
            v[4 * c] = r.x + rs * v[4 * c]; v[4 * c + 1] = r.y + rs * v[4 * c + 1];
            v[4 * c + 2] = r.z + rs * v[4 * c + 2]; v[4 * c + 3] = r.w + rs * v[4 * c + 3];
          }
        }
        if (MODE == 2) {
          const float4* ap = reinterpret_cast<const float4*>(a.aux + row * N + ch * 32);
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const float4 h = __ldg(ap + c);
            const float hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int j = 0; j < 4; j++) { float cdf, pdf; gelu_cdf_pdf(hh[j], cdf, pdf); v[4 * c + j] *= cdf + hh[j] * pdf; }
          }
        }
        const int sb = cc & 1;
        if (tid == 0) tma_store_wait_read<1>();     // the store that last read this staging buffer has drained it
        named_bar_sync(1, 128);
        unsigned char* srow = p_stage + (size_t)sb * GT_TILE_BYTES + (size_t)m * 128;
#pragma unroll
        for (int c = 0; c < 8; c++)
          *reinterpret_cast<float4*>(srow + ((c ^ (m & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        fence_proxy_async();
        named_bar_sync(1, 128);
        if (tid == 0) {
          tma_store_2d(&tmy, ch * 32, tile * 128, stage_s + (uint32_t)sb * GT_TILE_BYTES);
          tma_store_commit();
        }
        if (a.stats) {
          // column sums from the staging tile: thread (scol, srow0) adds 32 rows of one column
          const unsigned char* sbase = p_stage + (size_t)sb * GT_TILE_BYTES;
          float s = 0.f, q = 0.f;
          // the activation kind is uniform: branch once per chunk, not per element (see stat_act in common.cuh)
#define GT_COL(rr) (*reinterpret_cast<const float*>(sbase + (rr) * 128 + ((((scol >> 2) ^ ((rr) & 7)) << 4) | ((scol & 3) << 2))))
          if (a.stats_act == ACT_NONE) {
#pragma unroll 8
            for (int r = 0; r < 32; r++) { const float u = GT_COL(srow0 + r); s += u; q += u * u; }
          } else if (a.stats_act == ACT_LRELU) {
#pragma unroll 8
            for (int r = 0; r < 32; r++) { float u = GT_COL(srow0 + r); u = u > 0.f ? u : 0.01f * u; s += u; q += u * u; }
          } else {
#pragma unroll 2
            for (int r = 0; r < 32; r++) { const float u = act_fwd_rare(a.stats_act, GT_COL(srow0 + r)); s += u; q += u * u; }
          }
#undef GT_COL
          if (ch < 8) { st_sum[ch] += s; st_sq[ch] += q; }
        }
        if (MODE == 1) {       // second output: GELU of the tile just stored, through the other staging buffer
          cc++;
          const int sb2 = cc & 1;
          if (tid == 0) tma_store_wait_read<1>();
          named_bar_sync(1, 128);
          unsigned char* srow2 = p_stage + (size_t)sb2 * GT_TILE_BYTES + (size_t)m * 128;
#pragma unroll
          for (int i = 0; i < 32; i++) { float cdf, pdf; gelu_cdf_pdf(v[i], cdf, pdf); v[i] *= cdf; }
#pragma unroll
          for (int c = 0; c < 8; c++)
            *reinterpret_cast<float4*>(srow2 + ((c ^ (m & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          fence_proxy_async();
          named_bar_sync(1, 128);
          if (tid == 0) {
            tma_store_2d(&tmy2, ch * 32, tile * 128, stage_s + (uint32_t)sb2 * GT_TILE_BYTES);
            tma_store_commit();
          }
        }
      }
    }
    if (tid == 0) tma_store_wait<0>();
    if (a.stats) {
      for (int ch = 0; ch < nchunk && ch < 8; ch++) {
        atomicAdd(&s_stats[ch * 32 + scol], st_sum[ch]);
        atomicAdd(&s_stats[N + ch * 32 + scol], st_sq[ch]);
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
    const uint64_t desc_hi = (uint64_t)(uint32_t)(umma_desc(0u, 16u, 1024u, 2u, 0u) >> 32) << 32;
    const uint32_t a_lo0 = (uint32_t)umma_desc(ring_s, 16u, 1024u, 2u, 0u);
    const uint32_t b_lo0 = (uint32_t)umma_desc(w_s, 16u, 1024u, 2u, 0u);
    const uint32_t wslab16 = (uint32_t)N * 8u;                 // N rows x 128 B per slab, in 16-byte units
    int slot = 0, phase = 0, it = 0;
    mbar_wait(bar_w, 0);
    for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x, it++) {
      const int acc = it & 1;
      mbar_wait(bar_tempty + 8 * acc, ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.ncol);
      for (int s = 0; s < nslab; s++) {
        mbar_wait(bar_full + 8 * slot, phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = a_lo0 + (uint32_t)slot * (GT_TILE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + (uint32_t)s * wslab16;
#pragma unroll
          for (int ks = 0; ks < 4; ks++)
            tc_mma_tf32(d_tmem, desc_hi | (a_lo + 2u * ks), desc_hi | (b_lo + 2u * ks), idesc, (s | ks) ? 1u : 0u);
          tc_commit(bar_empty + 8 * slot);
          if (s == nslab - 1) tc_commit(bar_tfull + 8 * acc);
        }
        __syncwarp();
        if (++slot == NS) { slot = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmx); tma_prefetch_desc(&tmy);
      mbar_expect_tx(bar_w, wbytes);
      bulk_load(w_s, a.wu, wbytes, bar_w);
      int slot = 0, phase = 1;
      for (int tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
        for (int s = 0; s < nslab; s++) {
          mbar_wait(bar_empty + 8 * slot, phase);
          mbar_expect_tx(bar_full + 8 * slot, GT_TILE_BYTES);
          tma_load_2d(ring_s + (uint32_t)slot * GT_TILE_BYTES, &tmx, s * 32, tile * 128, bar_full + 8 * slot);
          if (++slot == NS) { slot = 0; phase ^= 1; }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (a.stats)
    for (int i = tid; i < 2 * N; i += GT_THREADS) atomicAdd(a.stats + i, (double)s_stats[i]);
  if (warp == 4) tmem_dealloc<TCOLS>(tmem_base);
}

static size_t gemm_tma_fixed_smem(int K, int N) {
  return 1024 + 2 * GT_TILE_BYTES + (size_t)K * N * 4 + (size_t)3 * N * 4 + (2 * GT_NS_MAX + 5) * 8 + 16;
}

// 1 if this shape runs on the TMA/tcgen05 GEMM
extern "C" int tcct_gemm_tma_supported(long long M, int K, int N) {
  if (M <= 0 || M % 128 != 0 || M >= (1ll << 31)) return 0;
  if (K % 32 != 0 || N % 32 != 0 || K < 32 || N < 32 || N > 256 || K > 512) return 0;
  if (M < 128ll * 64) return 0;                                  // small maps stay on the latency-tuned mma.sync kernel
  if (gemm_tma_fixed_smem(K, N) + 3 * GT_TILE_BYTES > 227 * 1024) return 0;
  return tcct_tensor_map_encoder() != nullptr ? 1 : 0;
}

// wu: weights packed by tcct_pack_weights with fmt = 3 ([slab][n][chunk ^ (n & 7)][4], tf32-rounded)
// y_act (or null): second output GELU(y);  mul_aux (or null): y is multiplied by GELU'(mul_aux) before it is stored
static int gemm_tma_launch(const float* x, const float* wu, const float* bias, float* y, long long M, int K, int N,
                           const float* res, const float* res_scale, int px_per_sample, double* stats, int stats_act,
                           float* y_act, const float* mul_aux, void* stream) {
  TCCT_CHECK_ARG(tcct_gemm_tma_supported(M, K, N), "gemm_tma: unsupported shape M=%lld K=%d N=%d", M, K, N);
  TCCT_CHECK_ARG(!(y_act && mul_aux), "gemm_tma: the GELU output and the GELU' factor are exclusive");
  GemmTmaArgs a;
  a.aux = mul_aux;
  a.wu = wu; a.bias = bias; a.res = res; a.res_scale = res_scale; a.stats = stats; a.stats_act = stats_act;
  a.M = (int)M; a.K = K; a.N = N; a.px_per_sample = px_per_sample > 0 ? px_per_sample : (int)M;
  a.tiles = (int)(M / 128);
  a.ncol = N;
  const size_t fixed = gemm_tma_fixed_smem(K, N);
  a.NS = (int)((227 * 1024 - fixed) / GT_TILE_BYTES);
  if (a.NS > GT_NS_MAX) a.NS = GT_NS_MAX;
  const size_t smem = fixed + (size_t)a.NS * GT_TILE_BYTES;
  CUtensorMap tmx, tmy, tmy2;
  const unsigned long long dx[2] = {(unsigned long long)K, (unsigned long long)M}, sx[1] = {(unsigned long long)K * 4ull};
  const unsigned long long dyv[2] = {(unsigned long long)N, (unsigned long long)M}, sy[1] = {(unsigned long long)N * 4ull};
  const unsigned int box[2] = {32u, 128u};
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmx, x, 2, dx, sx, box, 1), "gemm_tma: cuTensorMapEncodeTiled failed (x)");
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmy, y, 2, dyv, sy, box, 1), "gemm_tma: cuTensorMapEncodeTiled failed (y)");
  TCCT_CHECK_ARG(tcct_make_tensor_map(&tmy2, y_act ? y_act : y, 2, dyv, sy, box, 1), "gemm_tma: cuTensorMapEncodeTiled failed (y_act)");
  const int mode = y_act ? 1 : (mul_aux ? 2 : 0);
  int ctas = tcct_num_sms();
  if (ctas > a.tiles) ctas = a.tiles;
  cudaStream_t st = (cudaStream_t)stream;
  const int cols = 2 * N;
#define GT_LAUNCH_M(TC, MD)                                                                                      \
  do {                                                                                                          \
    cudaFuncSetAttribute(gemm_tma_kernel<TC, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    gemm_tma_kernel<TC, MD><<<ctas, GT_THREADS, smem, st>>>(tmx, tmy, tmy2, a);                                  \
  } while (0)
#define GT_LAUNCH(TC)                                                                                           \
  do {                                                                                                          \
    if (mode == 0) GT_LAUNCH_M(TC, 0); else if (mode == 1) GT_LAUNCH_M(TC, 1); else GT_LAUNCH_M(TC, 2);          \
  } while (0)
  if (cols <= 64) GT_LAUNCH(64);
  else if (cols <= 128) GT_LAUNCH(128);
  else if (cols <= 256) GT_LAUNCH(256);
  else GT_LAUNCH(512);
#undef GT_LAUNCH
#undef GT_LAUNCH_M
  tcct_count_route(TCCT_ROUTE_GEMM_TMA);
  TCCT_CHECK_LAUNCH("gemm_tma");
  return TCCT_OK;
}

extern "C" int tcct_gemm_tma(const float* x, const float* wu, const float* bias, float* y, long long M, int K, int N,
                             const float* res, const float* res_scale, int px_per_sample, double* stats, int stats_act,
                             void* stream) {
  return gemm_tma_launch(x, wu, bias, y, M, K, N, res, res_scale, px_per_sample, stats, stats_act, nullptr, nullptr, stream);
}
// Mlp.fc1 with its activation (tcct.py:29-53, 467-468): y = x W^T + b (kept for the backward), y_act = GELU(y) in one launch.
extern "C" int tcct_gemm_tma_gelu(const float* x, const float* wu, const float* bias, float* y, float* y_act, long long M, int K, int N,
                                  void* stream) {
  TCCT_CHECK_ARG(y_act != nullptr, "gemm_tma_gelu: y_act is required");
  return gemm_tma_launch(x, wu, bias, y, M, K, N, nullptr, nullptr, 0, nullptr, 0, y_act, nullptr, stream);
}
// Data gradient through Mlp.fc2 and the activation: dh = (dy W) * GELU'(h), h = the pre-activation tcct_gemm_tma_gelu kept.
extern "C" int tcct_gemm_tma_dgelu(const float* dy, const float* wu_t, const float* h, float* dh, long long M, int K, int N, void* stream) {
  TCCT_CHECK_ARG(h != nullptr, "gemm_tma_dgelu: the saved pre-activation is required");
  return gemm_tma_launch(dy, wu_t, nullptr, dh, M, K, N, nullptr, nullptr, 0, nullptr, 0, nullptr, h, stream);
}
