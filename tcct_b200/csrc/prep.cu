// Input / deploy formats either side of the hot path, the deterministic part of the reference's
// data path on decoded uint8 frames:
//   EyeSetResource.readPair  (task1/data/octnpy.py:117-129): rows [row0, row0+rows) of the BGR frame and of the gray label
//       PNG, label // divide (30), then alb.Resize(H, W, INTER_NEAREST) (octnpy.py:70-73,82-85); EyeSetGenerator.__getitem__
//       (task1/data/octgen.py:124-126): HWC uint8 -> CHW float / 255, clamped to [0, 1]; the label map stays a uint8 index
//       map (what the loss kernels take; the reference widens it to int64 and one-hots it, loop_seg.py:119).
//   EyeSetResource.postprocess (octnpy.py:95-112): label map * divide as uint8, alb.Resize back to the raw size
//       (INTER_NEAREST), pasted into rows [row0, row0+rows) of a zero frame.
// cv2's INTER_NEAREST maps destination index d to source index min(floor(d * src / dst), src - 1), evaluated in double.
// The random augmentations (albumentations make_tran, octgen.py:9-19) are not restated: the package is absent here.
#include "common.cuh"

__device__ __forceinline__ int nearest_src(int d, int src, int dst) {
  const int s = (int)floor((double)d * ((double)src / (double)dst));
  return s < src - 1 ? s : src - 1;
}

// img: [B][Hs][Ws][3] uint8 (cv2.IMREAD_COLOR order, kept), lab: [B][Hs][Ws] uint8 gray levels
// out_img: [B][3][H][W] float in [0,1]; out_lab: [B][H][W] uint8 class indices
__global__ void prep_pair_kernel(const unsigned char* __restrict__ img, const unsigned char* __restrict__ lab, int B, int Hs, int Ws,
                                 int row0, int rows, int H, int W, int divide, float* __restrict__ out_img,
                                 unsigned char* __restrict__ out_lab) {
  for (int r = blockIdx.x; r < B * H; r += gridDim.x) {
    const int b = r / H, y = r - b * H;
    const int sy = row0 + nearest_src(y, rows, H);
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
      const int sx = nearest_src(x, Ws, W);
      const size_t sp = ((size_t)b * Hs + sy) * Ws + sx;
      if (out_img) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float v = (float)img[sp * 3 + c] / 255.f;
          out_img[(((size_t)b * 3 + c) * H + y) * W + x] = fminf(fmaxf(v, 0.f), 1.f);
        }
      }
      if (out_lab) out_lab[((size_t)b * H + y) * W + x] = (unsigned char)(lab[sp] / divide);
    }
  }
}
extern "C" int tcct_prep_pair(const unsigned char* img, const unsigned char* lab, int B, int Hs, int Ws, int row0, int rows, int H,
                              int W, int divide, float* out_img, unsigned char* out_lab, void* stream) {
  TCCT_CHECK_ARG(B > 0 && Hs > 0 && Ws > 0 && H > 0 && W > 0, "prep_pair: empty input");
  TCCT_CHECK_ARG(row0 >= 0 && rows > 0 && row0 + rows <= Hs, "prep_pair: rows [%d, %d) outside the %d-row frame", row0, row0 + rows, Hs);
  TCCT_CHECK_ARG(divide > 0, "prep_pair: divide must be positive");
  TCCT_CHECK_ARG((img != nullptr) == (out_img != nullptr) && (lab != nullptr) == (out_lab != nullptr), "prep_pair: input/output mismatch");
  int grid = B * H;
  if (grid > tcct_num_sms() * 16) grid = tcct_num_sms() * 16;
  prep_pair_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, lab, B, Hs, Ws, row0, rows, H, W, divide, out_img, out_lab);
  TCCT_CHECK_LAUNCH("prep_pair");
  return TCCT_OK;
}

// lab: [B][H][W] uint8 class indices -> out: [B][Hfull][Wo] uint8 gray levels (index * divide), rows [row0, row0+Ho) hold the
// nearest-resized map, the rest is zero
__global__ void post_labels_kernel(const unsigned char* __restrict__ lab, int B, int H, int W, int Ho, int Wo, int row0, int Hfull,
                                   int divide, unsigned char* __restrict__ out) {
  for (int r = blockIdx.x; r < B * Hfull; r += gridDim.x) {
    const int b = r / Hfull, y = r - b * Hfull;
    const bool inside = y >= row0 && y < row0 + Ho;
    const int sy = inside ? nearest_src(y - row0, H, Ho) : 0;
    for (int x = threadIdx.x; x < Wo; x += blockDim.x) {
      unsigned char v = 0;
      if (inside) v = (unsigned char)(lab[((size_t)b * H + sy) * W + nearest_src(x, W, Wo)] * divide);
      out[((size_t)b * Hfull + y) * Wo + x] = v;
    }
  }
}
extern "C" int tcct_post_labels(const unsigned char* lab, int B, int H, int W, int Ho, int Wo, int row0, int Hfull, int divide,
                                unsigned char* out, void* stream) {
  TCCT_CHECK_ARG(B > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "post_labels: empty input");
  TCCT_CHECK_ARG(row0 >= 0 && row0 + Ho <= Hfull, "post_labels: rows [%d, %d) outside the %d-row frame", row0, row0 + Ho, Hfull);
  int grid = B * Hfull;
  if (grid > tcct_num_sms() * 16) grid = tcct_num_sms() * 16;
  post_labels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(lab, B, H, W, Ho, Wo, row0, Hfull, divide, out);
  TCCT_CHECK_LAUNCH("post_labels");
  return TCCT_OK;
}
