// TMA (cp.async.bulk.tensor), mbarrier and tcgen05 helpers shared by the tensor-memory kernels (sm_100a).
#pragma once
#include "common.cuh"
#include <cuda.h>      // CUtensorMap and its enums only; the driver entry point is resolved at run time (no -lcuda)

// ---- host: tensor-map encoding -------------------------------------------------------------------------------
typedef CUresult (*tcct_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                         const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                         CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tcct_encode_tiled_fn tcct_tensor_map_encoder();      // runtime.cu

// fp32 tensor of rank <= 5: dims[0] is the contiguous one; strides_bytes[i] is the stride of dims[i+1]
// swizzle: 0 none, 1 = 128-byte span / 16-byte chunks (K-major operands), 2 = 128-byte span / 32-byte chunks
// (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B: what tcgen05 wants for MN-major 32-bit operands, layout type 1)
static inline bool tcct_make_tensor_map(CUtensorMap* tm, const void* base, int rank, const unsigned long long* dims,
                                        const unsigned long long* strides_bytes, const unsigned int* box, int swizzle) {
  tcct_encode_tiled_fn enc = tcct_tensor_map_encoder();
  if (!enc) return false;
  // the encoder is a driver-API call and needs the primary context current in THIS thread; a thread whose first CUDA work is a
  // tensor-map kernel (autograd's backward thread) has only had cudaSetDevice so far, which does not bind it before CUDA 12
  static thread_local bool bound = false;
  if (!bound) { cudaFree(nullptr); bound = true; }
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; i++) gs[i] = strides_bytes[i];
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE,
             swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : (swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE),
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- device: mbarrier ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  unsigned long long spins = 0;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1ull << 22)) __trap();      // a lost arrival must not hang the GPU
  } while (!ok);
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---- device: TMA ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src) : "memory");
}
// the same as an element-wise ADD into global memory (the reduction happens at the L2: no read on the SM side)
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t src) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(src) : "memory");
}
// plain (1-D) bulk copy global -> shared, completion counted on an mbarrier; bytes % 16 == 0
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// one lane of a converged warp (the others skip); keeps the surrounding control flow warp-uniform
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}

// ---- device: tcgen05 ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "n"(COLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS));
}
// shared-memory matrix descriptor (sm_100 version 1).  layout: 0 = no swizzle, 2 = 128-byte swizzle (16-byte chunks),
// 1 = 128-byte swizzle with 32-byte chunks (MN-major 32-bit operands).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)(base_off & 7u) << 49) | ((uint64_t)(layout & 7u) << 61);
}
// kind::tf32 instruction descriptor: fp32 accumulate, M rows, N columns, operand majorness (0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
// 32 lanes x 32 columns of fp32: thread t of the warp receives row (lane quarter base + t), columns [col, col+32)
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

// ---- device: cross-CTA reduction of per-CTA partial sums (run by the follow-up reduce launch) ------------------------
// ws holds G partial vectors of `total` floats.  The calling CTA (NW warps, all threads) sums elements [e0, e1) over the
// G partials and hands every sum to `emit(e, sum)`.  32 consecutive elements per pass: lane = element (coalesced 128-byte
// loads), warp w takes partials w, w + NW, ... with several independent loads in flight; the warps' subtotals meet in
// shared memory.  s_part: NW * 32 floats.
template <int NW, class Emit>
__device__ __forceinline__ void reduce_partials(const float* __restrict__ ws, int total, unsigned int G, int e0, int e1,
                                                float* s_part, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int eb = e0; eb < e1; eb += 32) {
    const int e = eb + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (e < e1) {
      const float* p = ws + e;
      unsigned int c = (unsigned int)warp;
#pragma unroll 2
      for (; c + 3u * NW < G; c += 4u * NW) {
        s0 += __ldcg(p + (size_t)c * total);
        s1 += __ldcg(p + (size_t)(c + NW) * total);
        s2 += __ldcg(p + (size_t)(c + 2u * NW) * total);
        s3 += __ldcg(p + (size_t)(c + 3u * NW) * total);
      }
      for (; c < G; c += NW) s0 += __ldcg(p + (size_t)c * total);
    }
    s_part[warp * 32 + lane] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (warp == 0 && e < e1) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < NW; w++) sum += s_part[w * 32 + lane];
      emit(e, sum);
    }
    __syncthreads();
  }
}
