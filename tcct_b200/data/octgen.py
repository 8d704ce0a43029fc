"""Training-time augmentation on the GPU -- drop-in for `make_tran` / `ALB_TWIST` of task1/data/octgen.py:9-20 and the train branch of
`EyeSetGenerator.__getitem__` (octgen.py:117-126).

The reference runs albumentations on the host, per image, in four loader workers: PadIfNeeded -> CropNonEmptyMaskIfExists ->
HorizontalFlip -> VerticalFlip -> RGBShift -> HueSaturationValue -> RandomContrast -> RandomBrightness, then HWC uint8 -> CHW float / 255.
Here only the random DRAWS happen on the host (a 64-byte record per image); the pixels are produced by one kernel launch per batch
straight from the decoded frames (csrc/prep.cu: prep_augment_kernel = readPair + make_tran + tensor conversion), the label map stays
a uint8 index map.  The pixel arithmetic of every step is bit-identical to albumentations' uint8 code path on top of cv2
(oracle/aug_oracle.py, pinned by cv2-made goldens); the parameter ranges are albumentations' defaults."""
import numpy as np
import torch

from .. import _lib as L
from ..ops import _p, _stream
from .octnpy import _RESIZE_SETS, _as_u8

SIZE_IMAGEH, SIZE_IMAGEW = 256, 256            # octgen.py:8

AUG_DTYPE = np.dtype([("y0", "<i4"), ("x0", "<i4"), ("hflip", "<i4"), ("vflip", "<i4"), ("rgb_shift", "<f4", (3,)), ("contrast_alpha", "<f4"),
                      ("brightness_add", "<f4"), ("flags", "<i4"), ("hue_shift", "<f8"), ("sat_shift", "<f8"), ("val_shift", "<f8")])


def pack_params(draws, geometry_only=False, reflect=False):
    """List of draw dicts (keys of `GpuTwist.sample`) -> the kernel's record array (uint8 tensor).  geometry_only: no colour step at
    all (not even the HSV round trip albumentations performs with zero shifts) -- readPair of the padding datasets."""
    assert AUG_DTYPE.itemsize == L._lib.tcct_aug_params_size(), "AugParams ABI mismatch"
    rec = np.zeros(len(draws), dtype=AUG_DTYPE)
    for i, p in enumerate(draws):
        alpha, beta = float(p["contrast_alpha"]), float(p["brightness_beta"])
        flags = ((p["hue_shift"] != 0) * 1 + (p["sat_shift"] != 0) * 2 + (p["val_shift"] != 0) * 4 + (alpha != 1) * 8 + (beta != 0) * 16
                 + (32 if geometry_only else 0) + (64 if reflect else 0))
        rec[i] = (p["y0"], p["x0"], int(p["hflip"]), int(p["vflip"]), tuple(np.float32(v) for v in p["rgb_shift"]), np.float32(alpha),
                  np.float32(beta * 255), flags, p["hue_shift"], p["sat_shift"], p["val_shift"])
    return torch.from_numpy(rec.view(np.uint8).reshape(len(draws), -1).copy())


class GpuTwist(object):
    """What `make_tran(H, W)` returns.  `sample(mask)` draws one parameter record like the albumentations pipeline would (crop window
    around a random non-empty mask pixel, flips with p = 0.5, shifts +-20 / +-30 / +-20, contrast and brightness limits 0.2);
    `EyeSetResource.readPairAug` applies a batch of records on the device."""

    def __init__(self, height, width, seed=None):
        self.height, self.width = int(height), int(width)
        self.rng = np.random.default_rng(seed)

    def padded_shape(self, hp, wp):
        return max(hp, self.height), max(wp, self.width)

    def sample(self, mask):
        """mask: host uint8 [Hp, Wp] class-index map as readPair produces it (before padding); returns a draw dict."""
        H, W = self.height, self.width
        hp, wp = mask.shape
        mh, mw = self.padded_shape(hp, wp)
        top, left = ((H - hp) // 2 if hp < H else 0), ((W - wp) // 2 if wp < W else 0)
        rng = self.rng
        nz = np.argwhere(mask > 0)
        if len(nz):
            y, x = nz[rng.integers(len(nz))]
            x_min = int(np.clip(x + left - rng.integers(0, W), 0, mw - W))
            y_min = int(np.clip(y + top - rng.integers(0, H), 0, mh - H))
        else:
            x_min, y_min = int(rng.integers(0, mw - W + 1)), int(rng.integers(0, mh - H + 1))
        u = lambda lim: float(rng.uniform(-lim, lim))     # noqa: E731
        return {"y0": y_min, "x0": x_min, "hflip": bool(rng.random() < 0.5), "vflip": bool(rng.random() < 0.5),
                "rgb_shift": (u(20), u(20), u(20)), "hue_shift": u(20), "sat_shift": u(30), "val_shift": u(20),
                "contrast_alpha": 1.0 + u(0.2), "brightness_beta": u(0.2)}


def make_tran(SIZE_IMAGEH=SIZE_IMAGEH, SIZE_IMAGEW=SIZE_IMAGEW, seed=None):
    """octgen.py:9-19."""
    return GpuTwist(SIZE_IMAGEH, SIZE_IMAGEW, seed)


ALB_TWIST = make_tran(SIZE_IMAGEH, SIZE_IMAGEW)


def read_pair_aug(resource, img, lab, draws, twist=ALB_TWIST):
    """readPair (octnpy.py:117-129) + the train branch of __getitem__ (octgen.py:117-126) for a batch: decoded uint8 frames
    [B,Hs,Ws,3] / gray-level label maps [B,Hs,Ws] and one draw dict per image -> {'img': float32 [B,3,H,W], 'lab': uint8 [B,H,W]} on
    the device, one kernel launch."""
    img, lab = resource._decode(img, lab)
    img, lab = _as_u8(img, resource.device), _as_u8(lab, resource.device)
    if img.dim() == 3:
        img, lab = img[None], lab[None]
    B, Hs, Ws, _ = img.shape
    if len(draws) != B:
        raise RuntimeError("read_pair_aug: %d frames but %d parameter records" % (B, len(draws)))
    row0 = min(resource.height_stt, Hs)
    rows = min(resource.height_end, Hs) - row0
    Hp, Wp = resource.prep_size
    H, W = twist.height, twist.width
    params = pack_params(draws).to(resource.device, non_blocking=True)
    out_img = torch.empty((B, 3, H, W), dtype=torch.float32, device=resource.device)
    out_lab = torch.empty((B, H, W), dtype=torch.uint8, device=resource.device)
    L.prep_augment(_p(img), _p(lab), _p(params), B, Hs, Ws, row0, rows, Hp, Wp, H, W, resource.divide, _p(out_img), _p(out_lab), _stream())
    return {'img': out_img, 'lab': out_lab}
