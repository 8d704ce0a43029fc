"""EyeSetResource -- the deterministic part of task1/data/octnpy.py on the GPU.

`readPair` (octnpy.py:117-129) crops the rows [height_stt, height_end) of the decoded frame and of its label PNG, divides
the label gray levels by `divide` (30) and applies the dataset's `prep_tran`; `EyeSetGenerator.__getitem__`
(task1/data/octgen.py:124-126) then turns the pair into a CHW float image in [0, 1] and a label map.  `postprocess`
(octnpy.py:95-112) maps a predicted label map back to the raw frame (gray levels, `post_tran`, paste).  Here both run as one
kernel launch each on uint8 frames already in device (or pinned host) memory: csrc/prep.cu.

Built for the datasets whose `prep_tran` is `alb.Resize(..., INTER_NEAREST)` (goals, hcms, hcms1, the `else` branch) and for the
padding ones (heg, duke, duke1, duke3: `alb.PadIfNeeded(min_height, min_width, BORDER_CONSTANT, 0)`, octnpy.py:56-63, centred with
the odd pixel at the bottom / right; duke2: the same with cv2.BORDER_REFLECT, 64-66).  The random augmentations of
octgen.make_tran are tcct_b200/data/octgen.py.  File decoding stays on the host (cv2, if present)."""
import numpy as np
import torch

from .. import _lib as L
from ..ops import _p, _stream

# dbname -> (height_stt, height_end, prep (H, W), post (H, W))        octnpy.py:70-89
_RESIZE_SETS = {
    "hcms": (0, 1024, (256, 512), (128, 1024)),
    "hcms1": (0, 1024, (256, 512), (128, 1024)),
    "goals": (0, 608, (608, 512), (608, 1100)),
    "odsgh": (0, 992, (496, 512), (992, 1024)),
}


# dbname -> (height_stt, height_end, (min_height, min_width), reflect)            octnpy.py:56-66
_PAD_SETS = {
    "heg": (83, 339, (256, 672), False),
    "duke": (0, 224, (256, 576), False),
    "duke1": (0, 224, (256, 576), False),
    "duke3": (0, 224, (256, 576), False),
    "duke2": (0, 384, (384, 576), True),          # cv2.BORDER_REFLECT
}


def _as_u8(a, device):
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a))
    if a.dtype != torch.uint8:
        raise TypeError("decoded frames must be uint8, got %s" % a.dtype)
    return a.to(device, non_blocking=True).contiguous()


class EyeSetResource(object):
    divide = 30

    def __init__(self, dbname='goals', device=None, **args):
        if dbname not in _RESIZE_SETS and dbname not in _PAD_SETS:
            raise NotImplementedError("tcct_b200.data: unknown dataset %r; built: %s" % (dbname, sorted(_RESIZE_SETS) + sorted(_PAD_SETS)))
        self.__name__ = dbname
        self.pad_size, self.pad_reflect = None, False
        if dbname in _PAD_SETS:
            self.height_stt, self.height_end, self.pad_size, self.pad_reflect = _PAD_SETS[dbname]
            self.prep_size = self.post_size = None
        else:
            self.height_stt, self.height_end, self.prep_size, self.post_size = _RESIZE_SETS[dbname]
        self.device = torch.device(device if device is not None else "cuda")

    def _decode(self, img, lab):
        if isinstance(img, str) or isinstance(lab, str):
            import cv2     # host-side PNG decoding, as in the reference
            if isinstance(img, str):
                img = cv2.imread(img, cv2.IMREAD_COLOR)
            if isinstance(lab, str):
                lab = cv2.imread(lab, cv2.IMREAD_GRAYSCALE)
        return img, lab

    def readPair(self, img, lab):
        """img: path or decoded uint8 frame [Hs,Ws,3] (or a batch [B,Hs,Ws,3]); lab: path or uint8 gray-level map [Hs,Ws]
        ([B,Hs,Ws]).  Returns {'img': float32 [3,H,W] in [0,1] ([B,3,H,W]), 'lab': uint8 class indices [H,W] ([B,H,W])}
        on the device -- the reference's readPair followed by the tensor conversion of __getitem__."""
        img, lab = self._decode(img, lab)
        img, lab = _as_u8(img, self.device), _as_u8(lab, self.device)
        single = img.dim() == 3
        if single:
            img, lab = img[None], lab[None]
        if img.dim() != 4 or img.shape[-1] != 3 or lab.shape != img.shape[:3]:
            raise RuntimeError("readPair expects frames [B,Hs,Ws,3] and labels [B,Hs,Ws], got %s and %s" % (tuple(img.shape), tuple(lab.shape)))
        B, Hs, Ws, _ = img.shape
        row0 = min(self.height_stt, Hs)
        rows = min(self.height_end, Hs) - row0
        if self.pad_size is not None:
            # PadIfNeeded: the cropped rows, centred in a zero frame of at least (min_height, min_width): the geometry-only mode of the
            # augmentation kernel (no resize: Hp x Wp = rows x Ws; window origin 0)
            from .octgen import pack_params
            H, W = max(rows, self.pad_size[0]), max(Ws, self.pad_size[1])
            ident = {"y0": 0, "x0": 0, "hflip": False, "vflip": False, "rgb_shift": (0.0, 0.0, 0.0), "hue_shift": 0.0, "sat_shift": 0.0,
                     "val_shift": 0.0, "contrast_alpha": 1.0, "brightness_beta": 0.0}
            params = pack_params([ident] * B, geometry_only=True, reflect=self.pad_reflect).to(self.device, non_blocking=True)
            out_img = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device)
            out_lab = torch.empty((B, H, W), dtype=torch.uint8, device=self.device)
            L.prep_augment(_p(img), _p(lab), _p(params), B, Hs, Ws, row0, rows, rows, Ws, H, W, self.divide, _p(out_img), _p(out_lab), _stream())
            return {'img': out_img[0] if single else out_img, 'lab': out_lab[0] if single else out_lab}
        H, W = self.prep_size
        out_img = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device)
        out_lab = torch.empty((B, H, W), dtype=torch.uint8, device=self.device)
        L.prep_pair(_p(img), _p(lab), B, Hs, Ws, row0, rows, H, W, self.divide, _p(out_img), _p(out_lab), _stream())
        return {'img': out_img[0] if single else out_img, 'lab': out_lab[0] if single else out_lab}

    def readPairAug(self, img, lab, draws, twist=None):
        """readPair + make_tran + tensor conversion for a training batch (tcct_b200/data/octgen.py:read_pair_aug)."""
        from . import octgen
        if self.pad_size is not None:
            raise NotImplementedError("tcct_b200.data: readPairAug is built for the nearest-resize datasets")
        return octgen.read_pair_aug(self, img, lab, draws, twist or octgen.ALB_TWIST)

    def postprocess(self, lab, raw_height, return_lab=False, raw_width=None):
        """lab: predicted class-index map, uint8 [H,W] or [B,H,W] (what KiteSeg.predict_labels returns).  Returns the uint8
        gray-level frame [raw_height, Wpost] ([B, ...]) the reference writes to disk: index * divide, `post_tran`, pasted
        into rows [height_stt, height_end) of a zero frame (octnpy.py:95-112)."""
        lab = _as_u8(lab, self.device)
        single = lab.dim() == 2
        if single:
            lab = lab[None]
        B, H, W = lab.shape
        if self.pad_size is not None:
            # octnpy.py:101-110: CenterCrop to min(label file size, prediction size), pasted into rows [height_stt, height_end) of a
            # zero frame of the label file's size (a few strided device copies: no arithmetic beyond index * divide)
            if raw_width is None:
                raise RuntimeError("postprocess: the padding datasets need raw_width (the label file's width)")
            h, w = min(raw_height, H), min(raw_width, W)
            rows = min(self.height_end, raw_height) - self.height_stt
            if rows != h or w != raw_width:
                raise RuntimeError("postprocess: a %dx%d crop does not fill rows [%d, %d) of a %dx%d frame (the reference's numpy "
                                   "assignment fails the same way)" % (h, w, self.height_stt, self.height_stt + rows, raw_height, raw_width))
            y1, x1 = (H - h) // 2, (W - w) // 2
            out = torch.zeros((B, raw_height, raw_width), dtype=torch.uint8, device=self.device)
            out[:, self.height_stt:self.height_stt + h] = (lab[:, y1:y1 + h, x1:x1 + w].to(torch.int32) * self.divide).to(torch.uint8)
            return out[0] if single else out
        Ho, Wo = self.post_size
        row0 = self.height_stt
        if row0 + Ho > raw_height:
            raise RuntimeError("postprocess: rows [%d, %d) do not fit a %d-row frame" % (row0, row0 + Ho, raw_height))
        out = torch.empty((B, raw_height, Wo), dtype=torch.uint8, device=self.device)
        L.post_labels(_p(lab), B, H, W, Ho, Wo, row0, raw_height, self.divide, _p(out), _stream())
        return out[0] if single else out
