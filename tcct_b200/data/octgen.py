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


# ----------------------------------------------------------------------------------------------------------------------------
# EyeSetGenerator -- drop-in for task1/data/octgen.py:24-129 (the dataset object KiteSeg consumes)
# ----------------------------------------------------------------------------------------------------------------------------
import glob
import os
from concurrent.futures import ThreadPoolExecutor

from .octnpy import _PAD_SETS, EyeSetResource

_OUT_CHANNELS = {"hcms": 9, "hcms1": 9, "duke": 9, "duke1": 9, "duke2": 9, "duke3": 9, "heg": 8, "goals": 5}     # octgen.py:35-60


def _listing(folder, sub):
    return (sorted(p.replace("\\", "/") for p in glob.glob(folder + "/%s/*/*.*" % sub))
            + sorted(p.replace("\\", "/") for p in glob.glob(folder + "/%s/*.*" % sub)))


class _Loader(object):
    """What `trainSet / valSet / testSet` return: a re-iterable, sized sequence of batches {'img', 'lab', 'tag'} on the device (the
    reference returns a torch DataLoader with 4 / 1 workers; here a thread pool decodes the PNGs and ONE kernel launch per batch does
    the rest)."""

    def __init__(self, ds, mode, bs, shuffle):
        self.ds, self.mode, self.bs, self.shuffle = ds, mode, int(bs), shuffle

    def __len__(self):
        n = self.ds.count(self.mode)
        return (n + self.bs - 1) // self.bs

    def __iter__(self):
        ds = self.ds
        n = ds.count(self.mode)
        order = ds.rng.permutation(n) if self.shuffle else np.arange(n)
        with ThreadPoolExecutor(max_workers=ds.workers) as pool:
            for b0 in range(0, n, self.bs):
                idx = order[b0:b0 + self.bs]
                pairs = list(pool.map(lambda i: ds.decode(self.mode, int(i)), idx))
                yield ds.collate(self.mode, pairs)


class EyeSetGenerator(EyeSetResource):
    """octgen.py:24-129.  `folder` holds train_img / train_lab (/ val_* / test_*) like the reference's dataset root; the label of an image
    is the file of the same name under *_lab.  `trainSet(bs)` yields augmented batches (make_tran on the GPU), `valSet` the reference's
    ALB_VALID (horizontal flip, vertical flip with p = 0.5), `testSet` plain readPair; `parse(batch)` = (img, lab, tag, None)."""
    exeMode = 'train'
    in_channels = 3

    def __init__(self, dbname='goals', folder=None, device=None, seed=None, workers=4, height=SIZE_IMAGEH, width=SIZE_IMAGEW, **args):
        super(EyeSetGenerator, self).__init__(dbname=dbname, device=device)
        self.folder = folder if folder is not None else r'G:\\Objects\\Cometition\\dataset\\OCTSets' + '/' + dbname      # octnpy.py:31-33
        self.src_oct = _listing(self.folder, "train_img")
        self.val_oct = _listing(self.folder, "val_img") or list(self.src_oct)
        self.inf_oct = _listing(self.folder, "test_img")
        self.src_lab = [p.replace('train_img', 'train_lab') for p in self.src_oct]
        self.val_lab = [p.replace('val_img', 'val_lab').replace('train_img', 'train_lab') for p in self.val_oct]
        self.inf_lab = [p.replace('test_img', 'test_lab') for p in self.inf_oct]
        self.lens = {'train': len(self.src_lab), 'val': len(self.val_lab), 'test': len(self.inf_oct)}
        self.out_channels = _OUT_CHANNELS.get(dbname, 8)
        self.exeNums = {'train': max(1, 735 // max(1, self.lens['train'])), 'val': 1, 'test': 1}                 # octgen.py:62
        self.twist = make_tran(height, width, seed)
        self.rng = np.random.default_rng(seed)
        self.workers = int(workers)
        print(self.__name__, self.lens)
        print('exeNums:', self.exeNums)

    # -- reference protocol
    def set_mode(self, mode='train'):
        self.exeMode = mode
        self.isTrainMode, self.isValMode, self.isTestMode = mode == 'train', mode == 'val', mode == 'test'

    def count(self, mode):
        return self.lens['train'] * self.exeNums['train'] if mode == 'train' else self.lens[mode]

    def __len__(self):
        return self.count(self.exeMode)

    def trainSet(self, bs=32, data='train'):
        self.set_mode('train')
        return _Loader(self, 'train', bs, shuffle=True)

    def valSet(self, bs=1, data='val'):
        self.set_mode('val')
        return _Loader(self, 'val', bs, shuffle=False)

    def testSet(self, bs=1, data='test'):
        self.set_mode('test')
        return _Loader(self, 'test', bs, shuffle=False)

    def parse(self, pics):
        return pics['img'], pics['lab'], pics['tag'], None

    # -- host side: decode + draws; device side: one launch per batch
    def decode(self, mode, idx):
        imgs, labs = {'train': (self.src_oct, self.src_lab), 'val': (self.val_oct, self.val_lab), 'test': (self.inf_oct, self.inf_lab)}[mode]
        i = idx % len(imgs)
        import cv2
        img, lab = cv2.imread(imgs[i], cv2.IMREAD_COLOR), cv2.imread(labs[i], cv2.IMREAD_GRAYSCALE)
        if img is None or lab is None:
            raise FileNotFoundError("cannot decode %s / %s" % (imgs[i], labs[i]))
        return img, lab, labs[i]

    def prep_geometry(self, Hs, Ws):
        """(row0, rows, Hp, Wp, pad_top, pad_left): the frame readPair makes of a raw Hs x Ws image -- resized to Hp x Wp, or its rows
        kept and centred in a (Hp, Wp) frame with those offsets."""
        row0 = min(self.height_stt, Hs)
        rows = min(self.height_end, Hs) - row0
        if self.pad_size is None:
            return row0, rows, self.prep_size[0], self.prep_size[1], 0, 0
        Hp, Wp = max(rows, self.pad_size[0]), max(Ws, self.pad_size[1])
        return row0, rows, Hp, Wp, (Hp - rows) // 2, (Wp - Ws) // 2

    def host_mask(self, lab):
        """The class-index mask readPair would return, on the host (numpy index arithmetic only: the crop window is drawn from it)."""
        Hs, Ws = lab.shape
        row0, rows, Hp, Wp, pt, pl = self.prep_geometry(Hs, Ws)
        m = (lab[row0:row0 + rows] // self.divide).astype(np.uint8)
        if self.pad_size is None:
            iy = np.minimum(np.floor(np.arange(Hp) * (rows / float(Hp))).astype(np.int64), rows - 1)
            ix = np.minimum(np.floor(np.arange(Wp) * (Ws / float(Wp))).astype(np.int64), Ws - 1)
            return m[iy][:, ix]
        return np.pad(m, ((pt, Hp - rows - pt), (pl, Wp - Ws - pl)), mode="symmetric" if self.pad_reflect else "constant")

    def collate(self, mode, pairs):
        tags = [t for _, _, t in pairs]
        shapes = {p[0].shape for p in pairs}
        if len(shapes) > 1:         # frames of different raw sizes (duke): one launch per frame, outputs have one size
            outs = [self.collate(mode, [p]) for p in pairs]
            return {'img': torch.cat([o['img'] for o in outs]), 'lab': torch.cat([o['lab'] for o in outs]), 'tag': tags}
        imgs, labs = np.stack([p[0] for p in pairs]), np.stack([p[1] for p in pairs])
        B, Hs, Ws, _ = imgs.shape
        row0, rows, Hp, Wp, pt, pl = self.prep_geometry(Hs, Ws)
        if mode == 'test':
            out = self.readPair(imgs, labs)
            return {'img': out['img'], 'lab': out['lab'], 'tag': tags}
        none = {"rgb_shift": (0.0, 0.0, 0.0), "hue_shift": 0.0, "sat_shift": 0.0, "val_shift": 0.0, "contrast_alpha": 1.0, "brightness_beta": 0.0}
        if mode == 'train':
            H, W = self.twist.height, self.twist.width
            draws = [self.twist.sample(self.host_mask(l)) for l in labs]
        else:                       # ALB_VALID (octgen.py:20-24): HorizontalFlip(p=1), VerticalFlip(p=0.5) on the whole frame
            H, W = Hp, Wp
            draws = [dict(none, y0=0, x0=0, hflip=True, vflip=bool(self.rng.random() < 0.5)) for _ in range(B)]
        # the kernel pads (rows x Ws) -> at least (H, W) itself, centred; a padding dataset's own frame is centred in (Hp, Wp): express
        # the window origin, which `sample` gives in the (Hp, Wp) frame padded further to (H, W), in the kernel's frame
        kernel_pt = (H - (rows if self.pad_size is not None else Hp)) // 2 if (rows if self.pad_size is not None else Hp) < H else 0
        kernel_pl = (W - (Ws if self.pad_size is not None else Wp)) // 2 if (Ws if self.pad_size is not None else Wp) < W else 0
        if self.pad_size is not None:
            outer_pt = (H - Hp) // 2 if Hp < H else 0
            outer_pl = (W - Wp) // 2 if Wp < W else 0
            for d in draws:
                d["y0"] += kernel_pt - pt - outer_pt
                d["x0"] += kernel_pl - pl - outer_pl
        params = pack_params(draws, geometry_only=(mode != 'train'), reflect=self.pad_reflect).to(self.device, non_blocking=True)
        img_d, lab_d = _as_u8(imgs, self.device), _as_u8(labs, self.device)
        out_img = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device)
        out_lab = torch.empty((B, H, W), dtype=torch.uint8, device=self.device)
        kHp, kWp = (rows, Ws) if self.pad_size is not None else (Hp, Wp)
        L.prep_augment(_p(img_d), _p(lab_d), _p(params), B, Hs, Ws, row0, rows, kHp, kWp, H, W, self.divide, _p(out_img), _p(out_lab), _stream())
        return {'img': out_img, 'lab': out_lab, 'tag': tags}
