"""CPU restatement (numpy) of the training-time augmentation of the reference -- TEST INFRASTRUCTURE, never imported by the product.

`make_tran` (task1/data/octgen.py:9-19) composes, on the uint8 HWC frame and the uint8 mask that readPair returned,
    PadIfNeeded(H, W, BORDER_CONSTANT, value 0, mask_value 0) -> CropNonEmptyMaskIfExists(H, W) -> HorizontalFlip(p=.5) -> VerticalFlip(p=.5)
    -> RGBShift(p=1) -> HueSaturationValue(p=1) -> RandomContrast(p=1) -> RandomBrightness(p=1)
and `EyeSetGenerator.__getitem__` (octgen.py:117-126) turns the result into a CHW float image / 255 and a label map.

The algorithm of every step lives in a THIRD-PARTY dependency that is absent from /root/reference and from this image: albumentations.
The reference pins no version (no requirements file); the API it uses (`RandomContrast`, `RandomBrightness`, `PadIfNeeded(mask_value=)`)
exists in albumentations 0.4 - 1.3.  Restated here from the published uint8 code path of albumentations 1.3.1
(albumentations/augmentations/functional.py: `_shift_rgb_uint8` / `_shift_image_uint8`, `_shift_hsv_uint8`,
`_brightness_contrast_adjust_uint`; geometric/functional.py: `pad_with_params`, crops/functional.py: `crop`, `hflip`, `vflip`), with the
random draws factored out into an explicit parameter record so that the kernel, this file and the goldens see the same numbers.
Those functions call two cv2 primitives, which ARE here and pin this file: cv2.LUT (a table look-up) and the 8-bit
cv2.cvtColor RGB2HSV / HSV2RGB, restated below and checked against cv2 4.13 for every possible input (16.7 M colours forward,
11.8 M (h < 180, s, v) triples backward) by oracle/make_golden_aug.py; tests/golden/aug_*.npz were written through cv2 itself.
PARITY NOTE: the composition and the parameter ranges are pinned only by this restatement (albumentations absent); the pixel
arithmetic of every step is pinned by cv2."""
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------- cv2 8-bit HSV (hrange 180)
def rgb2hsv_u8(img):
    """cv2.cvtColor(img, COLOR_RGB2HSV) for uint8 (OpenCV color_hsv: RGB2HSV_b, fixed point with hsv_shift = 12)."""
    sh = 12
    i = np.arange(256, dtype=np.float64)
    i[0] = 1
    sdiv = np.rint((255 << sh) / i).astype(np.int64)
    hdiv = np.rint((180 << sh) / (6.0 * i)).astype(np.int64)
    sdiv[0] = hdiv[0] = 0
    r, g, b = (img[..., k].astype(np.int64) for k in range(3))
    v = np.maximum(np.maximum(r, g), b)
    diff = v - np.minimum(np.minimum(r, g), b)
    s = (diff * sdiv[v] + (1 << (sh - 1))) >> sh
    h = np.where(v == r, g - b, np.where(v == g, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * hdiv[diff] + (1 << (sh - 1))) >> sh
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def _fma(a, b, c):
    """float32 fused multiply-add (the product of two float32 is exact in float64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


def hsv2rgb_u8(img):
    """cv2.cvtColor(img, COLOR_HSV2RGB) for uint8: float32 sector arithmetic with fused multiply-adds, TRUNCATED to uint8 (what
    cv2 4.13's vector path computes; bit-exact against it on every (h < 180, s, v))."""
    h = img[..., 0].astype(F32) * F32(6.0 / 180.0)
    s = img[..., 1].astype(F32) * F32(1.0 / 255.0)
    v = img[..., 2].astype(F32) * F32(1.0 / 255.0)
    sector = np.floor(h).astype(np.int32)
    h = h - sector.astype(F32)
    sector = sector % 6
    one = np.ones_like(h)
    tab = np.stack([v, v * (one - s), v * _fma(-s, h, one), v * _fma(-s, one - h, one)], -1)
    sd = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])[sector]       # (b, g, r) table rows
    pick = lambda k: np.take_along_axis(tab, sd[..., k:k + 1], -1)[..., 0]
    out = np.stack([pick(2), pick(1), pick(0)], -1) * F32(255.0)
    return np.clip(np.trunc(out), 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------- albumentations uint8 colour steps (as look-up tables)
def lut_shift(value):
    """_shift_image_uint8: lut = arange(256) as float32; lut += value; clip; astype(uint8)."""
    lut = np.arange(256).astype(F32)
    lut += value
    return np.clip(lut, 0, 255).astype(np.uint8)


def lut_hue(hue_shift):
    """_shift_hsv_uint8: lut_hue = mod(arange(256, int16) + hue_shift, 180).astype(uint8)."""
    return np.mod(np.arange(256, dtype=np.int16) + hue_shift, 180).astype(np.uint8)


def lut_clip_add(shift):
    """_shift_hsv_uint8: clip(arange(256, int16) + shift, 0, 255).astype(uint8) (saturation and value)."""
    return np.clip(np.arange(256, dtype=np.int16) + shift, 0, 255).astype(np.uint8)


def lut_brightness_contrast(alpha, beta):
    """_brightness_contrast_adjust_uint with beta_by_max=True: lut = arange(256) float32; *= alpha if alpha != 1; += beta * 255 if
    beta != 0; clip; astype(uint8).  RandomContrast is (alpha = 1 + u, beta = 0), RandomBrightness (alpha = 1, beta = u)."""
    lut = np.arange(256).astype(F32)
    if alpha != 1:
        lut *= alpha
    if beta != 0:
        lut += beta * 255
    return np.clip(lut, 0, 255).astype(np.uint8)


def colour_jitter(img, p, hsv_fwd=rgb2hsv_u8, hsv_inv=hsv2rgb_u8):
    """RGBShift -> HueSaturationValue -> RandomContrast -> RandomBrightness on a uint8 HWC image (channel 0 is treated as R, whatever
    cv2.imread's order: the reference never converts)."""
    img = np.stack([lut_shift(p["rgb_shift"][k])[img[..., k]] for k in range(3)], -1)
    hsv = hsv_fwd(img)
    h, s, v = hsv[..., 0], hsv[..., 1], hsv[..., 2]
    if p["hue_shift"] != 0:
        h = lut_hue(p["hue_shift"])[h]
    if p["sat_shift"] != 0:
        s = lut_clip_add(p["sat_shift"])[s]
    if p["val_shift"] != 0:
        v = lut_clip_add(p["val_shift"])[v]
    img = hsv_inv(np.stack([h, s, v], -1))
    img = lut_brightness_contrast(p["contrast_alpha"], 0.0)[img]
    return lut_brightness_contrast(1.0, p["brightness_beta"])[img]


# ----------------------------------------------------------------------------- geometry
def pad_if_needed(a, H, W):
    """PadIfNeeded(min_height=H, min_width=W, border_mode=BORDER_CONSTANT, value=0): centred, the odd pixel goes bottom / right."""
    rows, cols = a.shape[:2]
    top = int((H - rows) / 2.0) if rows < H else 0
    bottom = H - rows - top if rows < H else 0
    left = int((W - cols) / 2.0) if cols < W else 0
    right = W - cols - left if cols < W else 0
    pad = [(top, bottom), (left, right)] + [(0, 0)] * (a.ndim - 2)
    return np.pad(a, pad, mode="constant"), top, left


def sample_params(rng, mask, H, W):
    """The random draws of one make_tran call as a parameter record.  `mask` is the padded mask (CropNonEmptyMaskIfExists picks a
    non-zero pixel and a window around it; a window anywhere if the mask is empty).  Ranges are albumentations' defaults: shifts
    +-20 (RGB, hue, value), +-30 (saturation), contrast / brightness limit 0.2."""
    mh, mw = mask.shape
    nz = np.argwhere(mask > 0)
    if len(nz):
        y, x = nz[rng.integers(len(nz))]
        x_min = int(np.clip(x - rng.integers(0, W), 0, mw - W))
        y_min = int(np.clip(y - rng.integers(0, H), 0, mh - H))
    else:
        x_min, y_min = int(rng.integers(0, mw - W + 1)), int(rng.integers(0, mh - H + 1))
    u = lambda lim: float(rng.uniform(-lim, lim))
    return {"y0": y_min, "x0": x_min, "hflip": bool(rng.random() < 0.5), "vflip": bool(rng.random() < 0.5),
            "rgb_shift": (u(20), u(20), u(20)), "hue_shift": u(20), "sat_shift": u(30), "val_shift": u(20),
            "contrast_alpha": 1.0 + u(0.2), "brightness_beta": u(0.2)}


def make_tran_apply(img, mask, H, W, p, **hsv):
    """One make_tran call with the draws `p` (y0 / x0 are window origins in the PADDED frame) followed by the tensor conversion of
    octgen.py:124-126: (float32 [3,H,W] in [0,1], uint8 [H,W])."""
    img, _, _ = pad_if_needed(img, H, W)
    mask, _, _ = pad_if_needed(mask, H, W)
    img = img[p["y0"]:p["y0"] + H, p["x0"]:p["x0"] + W]
    mask = mask[p["y0"]:p["y0"] + H, p["x0"]:p["x0"] + W]
    if p["hflip"]:
        img, mask = img[:, ::-1], mask[:, ::-1]
    if p["vflip"]:
        img, mask = img[::-1], mask[::-1]
    img = colour_jitter(np.ascontiguousarray(img), p, **hsv)
    x = np.clip(img.transpose(2, 0, 1).astype(F32) / 255, 0, 1)
    return x, np.ascontiguousarray(mask).astype(np.uint8)
