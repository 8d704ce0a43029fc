"""Golden vectors of the training-time augmentation, made with cv2 (the primitives albumentations' uint8 code path calls):
tests/golden/aug_<db>.npz.      python oracle/make_golden_aug.py

1. pins oracle/aug_oracle.py's restatement of the 8-bit cv2.cvtColor RGB2HSV / HSV2RGB against cv2 itself for EVERY input
   (16 777 216 colours forward, 11 796 480 (h < 180, s, v) triples backward) -- aborts on the first mismatch;
2. runs make_tran (task1/data/octgen.py:9-19) with explicit draws on seeded frames, with cv2.cvtColor / cv2.LUT doing the pixel work,
   and stores the draws, a window of the output and checksums of the whole (the frames come from the same generator the tests use);
   one case per dataset family: GOALS-shaped (608x512 after readPair -> 256x256 crop), HCMS-shaped (256x512 -> 256x256) and a frame
   SMALLER than the crop (PadIfNeeded takes effect)."""
import os
import sys
import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import aug_oracle as A

ROOT = os.path.dirname(HERE)
CASES = {"goals": (608, 512, 256, 256, 5), "hcms": (256, 512, 256, 256, 9), "small": (200, 230, 256, 256, 5)}      # Hp, Wp, H, W, classes


def exhaustive_hsv():
    r = np.arange(256, dtype=np.uint8)
    R, G, B = np.meshgrid(r, r, r, indexing="ij")
    rgb = np.stack([R.ravel(), G.ravel(), B.ravel()], -1).reshape(4096, 4096, 3)
    assert np.array_equal(A.rgb2hsv_u8(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2HSV)), "RGB2HSV restatement differs from cv2"
    H, S, V = np.meshgrid(np.arange(180, dtype=np.uint8), r, r, indexing="ij")
    hsv = np.stack([H.ravel(), S.ravel(), V.ravel()], -1).reshape(180 * 16, 4096, 3)
    assert np.array_equal(A.hsv2rgb_u8(hsv), cv2.cvtColor(hsv, cv2.COLOR_HSV2RGB)), "HSV2RGB restatement differs from cv2"
    print("8-bit HSV restatement == cv2 %s on all 16 777 216 + 11 796 480 inputs" % cv2.__version__)


def frames(db):
    """What readPair returns for a seeded frame: uint8 HWC image, uint8 class-index mask (banded, with an empty margin)."""
    Hp, Wp, H, W, C = CASES[db]
    rng = np.random.default_rng(17)
    img = rng.integers(0, 256, (Hp, Wp, 3), dtype=np.uint8)
    img[:, : Wp // 3] = (img[:, : Wp // 3] // 8)             # a dark region: exercises the low end of the look-up tables
    lab = np.minimum(np.arange(Hp)[:, None] * C // Hp + rng.integers(0, 2, (Hp, Wp)), C - 1).astype(np.uint8)
    lab[:, : Wp // 5] = 0
    return img, lab


def cv_colour(img, p):
    """albumentations' uint8 path literally: cv2.LUT for every table, cv2.cvtColor for the colour space."""
    img = cv2.merge([cv2.LUT(np.ascontiguousarray(img[..., k]), A.lut_shift(p["rgb_shift"][k])) for k in range(3)])
    hsv = cv2.cvtColor(img, cv2.COLOR_RGB2HSV)
    h, s, v = cv2.split(hsv)
    if p["hue_shift"] != 0:
        h = cv2.LUT(h, A.lut_hue(p["hue_shift"]))
    if p["sat_shift"] != 0:
        s = cv2.LUT(s, A.lut_clip_add(p["sat_shift"]))
    if p["val_shift"] != 0:
        v = cv2.LUT(v, A.lut_clip_add(p["val_shift"]))
    img = cv2.cvtColor(cv2.merge((h, s, v)).astype(np.uint8), cv2.COLOR_HSV2RGB)
    img = cv2.LUT(img, A.lut_brightness_contrast(p["contrast_alpha"], 0.0))
    return cv2.LUT(img, A.lut_brightness_contrast(1.0, p["brightness_beta"]))


def cv_apply(img, mask, H, W, p):
    img = cv2.copyMakeBorder(img, *pads(img, H, W), cv2.BORDER_CONSTANT, value=0)
    mask = cv2.copyMakeBorder(mask, *pads(mask, H, W), cv2.BORDER_CONSTANT, value=0)
    img = img[p["y0"]:p["y0"] + H, p["x0"]:p["x0"] + W]
    mask = mask[p["y0"]:p["y0"] + H, p["x0"]:p["x0"] + W]
    if p["hflip"]:
        img, mask = cv2.flip(img, 1), cv2.flip(mask, 1)
    if p["vflip"]:
        img, mask = cv2.flip(img, 0), cv2.flip(mask, 0)
    img = cv_colour(np.ascontiguousarray(img), p)
    return np.clip(img.transpose(2, 0, 1).astype(np.float32) / 255, 0, 1), mask


def pads(a, H, W):
    rows, cols = a.shape[:2]
    top = int((H - rows) / 2.0) if rows < H else 0
    left = int((W - cols) / 2.0) if cols < W else 0
    return top, (H - rows - top if rows < H else 0), left, (W - cols - left if cols < W else 0)


if __name__ == "__main__":
    exhaustive_hsv()
    for db, (Hp, Wp, H, W, C) in CASES.items():
        img, lab = frames(db)
        rng = np.random.default_rng(23)
        padded_mask, _, _ = A.pad_if_needed(lab, H, W)
        draws, outs = [], {}
        for i in range(4):
            p = A.sample_params(rng, padded_mask, H, W)
            if i == 3:      # the identity-colour corner: tables that albumentations skips (shift == 0, alpha == 1, beta == 0)
                p.update(rgb_shift=(0.0, 0.0, 0.0), hue_shift=0.0, sat_shift=0.0, val_shift=0.0, contrast_alpha=1.0, brightness_beta=0.0)
            x, m = cv_apply(img, lab, H, W, p)
            ox, om = A.make_tran_apply(img, lab, H, W, p)
            assert np.array_equal(x, ox) and np.array_equal(m, om), "oracle differs from the cv2-made golden (%s, draw %d)" % (db, i)
            draws.append([p["y0"], p["x0"], int(p["hflip"]), int(p["vflip"]), *p["rgb_shift"], p["hue_shift"], p["sat_shift"], p["val_shift"],
                          p["contrast_alpha"], p["brightness_beta"]])
            outs["x_sum%d" % i] = np.float64(x.astype(np.float64).sum())
            outs["m_sum%d" % i] = np.int64(m.astype(np.int64).sum())
            outs["x_win%d" % i] = x[:, 96:160, 96:160]
            outs["m_win%d" % i] = m[96:160, 96:160]
            outs["x_rowsum%d" % i] = x.astype(np.float64).sum((0, 2))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "aug_%s.npz" % db), draws=np.array(draws, dtype=np.float64), **outs)
        print(db, (Hp, Wp), "->", (H, W), "draws", len(draws), "x_sum", [float(outs["x_sum%d" % i]) for i in range(4)])
