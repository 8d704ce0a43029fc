// Input / deploy formats either side of the hot path, the deterministic part of the reference's
// data path on decoded uint8 frames:
//   EyeSetResource.readPair  (task1/data/octnpy.py:117-129): rows [row0, row0+rows) of the BGR frame and of the gray label
//       PNG, label // divide (30), then alb.Resize(H, W, INTER_NEAREST) (octnpy.py:70-73,82-85); EyeSetGenerator.__getitem__
//       (task1/data/octgen.py:124-126): HWC uint8 -> CHW float / 255, clamped to [0, 1]; the label map stays a uint8 index
//       map (what the loss kernels take; the reference widens it to int64 and one-hots it, loop_seg.py:119).
//   EyeSetResource.postprocess (octnpy.py:95-112): label map * divide as uint8, alb.Resize back to the raw size
//       (INTER_NEAREST), pasted into rows [row0, row0+rows) of a zero frame.
// cv2's INTER_NEAREST maps destination index d to source index min(floor(d * src / dst), src - 1), evaluated in double.
//   make_tran (task1/data/octgen.py:9-19): PadIfNeeded -> CropNonEmptyMaskIfExists -> flips -> RGBShift -> HueSaturationValue ->
//       RandomContrast -> RandomBrightness on the uint8 pair, with the random draws handed in as a parameter record per sample
//       (prep_augment_kernel below: readPair + make_tran + the tensor conversion in ONE pass over the output).
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


// ----------------------------------------------------------------------------------------------------------------------------
// readPair + make_tran + tensor conversion in one launch.  Every colour step of albumentations' uint8 path is a 256-entry look-up
// table built from float arithmetic (functional.py: _shift_image_uint8, _shift_hsv_uint8, _brightness_contrast_adjust_uint) around
// cv2's 8-bit RGB<->HSV; here the table entries are evaluated per pixel with the same arithmetic in the same precision (float32
// where numpy keeps float32 tables, double where it promotes int16 + Python float), so results are identical to the tables'.
// The 8-bit HSV conversions restate OpenCV's (forward: fixed point, hsv_shift 12; backward: float32 sector arithmetic with fused
// multiply-adds, truncated) and are checked against cv2 4.13 on every input by oracle/make_golden_aug.py.
// ----------------------------------------------------------------------------------------------------------------------------
struct AugParams {          // one per sample; layout mirrored by tcct_b200/data/octgen.py (AUG_DTYPE), 64 bytes
  int y0, x0;               // crop origin in the padded frame
  int hflip, vflip;
  float rgb_shift[3];       // float32(value), as `lut += value` on a float32 table
  float contrast_alpha;     // float32(alpha)
  float brightness_add;     // float32(beta * 255)
  int flags;                // bit 0: hue table applied (hue_shift != 0), 1: saturation, 2: value, 3: contrast (alpha != 1), 4: brightness,
                            // bit 5: geometry only (readPair of the padding datasets: no colour step at all)
                            // bit 6: PadIfNeeded with cv2.BORDER_REFLECT (fedcba|abcdefgh|hgfedcb, image and mask) instead of zeros
  double hue_shift, sat_shift, val_shift;
};

__device__ __forceinline__ unsigned char u8_trunc_clip(float v) { return (unsigned char)fminf(fmaxf(v, 0.f), 255.f); }
__device__ __forceinline__ unsigned char u8_trunc_clip(double v) { return (unsigned char)fmin(fmax(v, 0.0), 255.0); }

__global__ void __launch_bounds__(256) prep_augment_kernel(const unsigned char* __restrict__ img, const unsigned char* __restrict__ lab,
                                                           const AugParams* __restrict__ params, int B, int Hs, int Ws, int row0,
                                                           int rows, int Hp, int Wp, int H, int W, int divide,
                                                           float* __restrict__ out_img, unsigned char* __restrict__ out_lab) {
  __shared__ int sdiv[256], hdiv[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {     // OpenCV's division tables: saturate_cast<int> rounds half to even
    sdiv[i] = i ? __double2int_rn((double)(255 << 12) / (double)i) : 0;
    hdiv[i] = i ? __double2int_rn((double)(180 << 12) / (6.0 * (double)i)) : 0;
  }
  __syncthreads();
  // PadIfNeeded: centred, the odd pixel goes to the bottom / right (int((min - size) / 2.0))
  const int pad_top = Hp < H ? (H - Hp) / 2 : 0, pad_left = Wp < W ? (W - Wp) / 2 : 0;
  for (int r = blockIdx.x; r < B * H; r += gridDim.x) {
    const int b = r / H, y = r - b * H;
    const AugParams p = params[b];
    int cy = (p.vflip ? H - 1 - y : y) + p.y0 - pad_top;       // row in the readPair frame
    const bool reflect = (p.flags & 64) != 0;
    if (reflect) cy = cy < 0 ? -cy - 1 : (cy >= Hp ? 2 * Hp - 1 - cy : cy);
    const bool row_in = cy >= 0 && cy < Hp;
    const int sy = row_in ? row0 + nearest_src(cy, rows, Hp) : 0;
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
      int cx = (p.hflip ? W - 1 - x : x) + p.x0 - pad_left;
      if (reflect) cx = cx < 0 ? -cx - 1 : (cx >= Wp ? 2 * Wp - 1 - cx : cx);
      const bool in = row_in && cx >= 0 && cx < Wp;
      int c[3] = {0, 0, 0};
      int l = 0;
      if (in) {
        const size_t sp = ((size_t)b * Hs + sy) * Ws + nearest_src(cx, Ws, Wp);
        c[0] = img[sp * 3]; c[1] = img[sp * 3 + 1]; c[2] = img[sp * 3 + 2];
        l = lab[sp] / divide;
      }
      out_lab[((size_t)b * H + y) * W + x] = (unsigned char)l;
      if (p.flags & 32) {
#pragma unroll
        for (int k = 0; k < 3; k++) out_img[(((size_t)b * 3 + k) * H + y) * W + x] = __fdiv_rn((float)c[k], 255.f);
        continue;
      }
      // RGBShift
#pragma unroll
      for (int k = 0; k < 3; k++) c[k] = u8_trunc_clip(__fadd_rn((float)c[k], p.rgb_shift[k]));
      // RGB -> HSV (8 bit, h in [0, 180))
      int hh, ss, vv;
      {
        const int rr = c[0], gg = c[1], bb = c[2];
        vv = max(max(rr, gg), bb);
        const int diff = vv - min(min(rr, gg), bb);
        ss = (diff * sdiv[vv] + (1 << 11)) >> 12;
        hh = vv == rr ? gg - bb : (vv == gg ? bb - rr + 2 * diff : rr - gg + 4 * diff);
        hh = (hh * hdiv[diff] + (1 << 11)) >> 12;
        if (hh < 0) hh += 180;
      }
      if (p.flags & 1) {          // np.mod(int16 + float, 180): the sign follows the divisor
        double m = fmod((double)hh + p.hue_shift, 180.0);
        if (m < 0) m += 180.0;
        hh = (int)(unsigned char)m;
      }
      if (p.flags & 2) ss = u8_trunc_clip((double)ss + p.sat_shift);
      if (p.flags & 4) vv = u8_trunc_clip((double)vv + p.val_shift);
      // HSV -> RGB
      {
        float h = __fmul_rn((float)hh, (float)(6.0 / 180.0));
        const float s = __fmul_rn((float)ss, (float)(1.0 / 255.0)), v = __fmul_rn((float)vv, (float)(1.0 / 255.0));
        const float fl = floorf(h);
        h = __fsub_rn(h, fl);
        int sector = (int)fl % 6;
        if (sector < 0) sector += 6;
        float tab[4];
        tab[0] = v;
        tab[1] = __fmul_rn(v, __fsub_rn(1.f, s));
        tab[2] = __fmul_rn(v, __fmaf_rn(-s, h, 1.f));
        tab[3] = __fmul_rn(v, __fmaf_rn(-s, __fsub_rn(1.f, h), 1.f));
        // (b, g, r) rows of OpenCV's sector table
        const int sb[6] = {1, 1, 3, 0, 0, 2}, sg[6] = {3, 0, 0, 2, 1, 1}, sr[6] = {0, 2, 1, 1, 3, 0};
        c[0] = u8_trunc_clip(__fmul_rn(tab[sr[sector]], 255.f));
        c[1] = u8_trunc_clip(__fmul_rn(tab[sg[sector]], 255.f));
        c[2] = u8_trunc_clip(__fmul_rn(tab[sb[sector]], 255.f));
      }
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (p.flags & 8) c[k] = u8_trunc_clip(__fmul_rn((float)c[k], p.contrast_alpha));
        if (p.flags & 16) c[k] = u8_trunc_clip(__fadd_rn((float)c[k], p.brightness_add));
        out_img[(((size_t)b * 3 + k) * H + y) * W + x] = __fdiv_rn((float)c[k], 255.f);
      }
    }
  }
}

extern "C" int tcct_aug_params_size(void) { return (int)sizeof(AugParams); }

// img [B,Hs,Ws,3] / lab [B,Hs,Ws]: decoded uint8 frames; rows [row0, row0+rows) are resized (cv2 INTER_NEAREST) to Hp x Wp (readPair),
// padded to at least H x W, cropped at the record's origin, flipped and colour-jittered; out_img [B,3,H,W] float, out_lab [B,H,W] uint8.
extern "C" int tcct_prep_augment(const unsigned char* img, const unsigned char* lab, const void* params_dev, int B, int Hs, int Ws, int row0,
                                 int rows, int Hp, int Wp, int H, int W, int divide, float* out_img, unsigned char* out_lab, void* stream) {
  TCCT_CHECK_ARG(B > 0 && Hs > 0 && Ws > 0 && Hp > 0 && Wp > 0 && H > 0 && W > 0, "prep_augment: empty input");
  TCCT_CHECK_ARG(row0 >= 0 && rows > 0 && row0 + rows <= Hs, "prep_augment: rows [%d, %d) outside the %d-row frame", row0, row0 + rows, Hs);
  TCCT_CHECK_ARG(divide > 0 && img && lab && params_dev && out_img && out_lab, "prep_augment: null argument");
  int grid = B * H;
  if (grid > tcct_num_sms() * 8) grid = tcct_num_sms() * 8;
  prep_augment_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, lab, (const AugParams*)params_dev, B, Hs, Ws, row0, rows, Hp, Wp, H, W,
                                                              divide, out_img, out_lab);
  TCCT_CHECK_LAUNCH("prep_augment");
  return TCCT_OK;
}
