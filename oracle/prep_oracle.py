"""CPU restatement (numpy) of the deterministic data path of the reference -- TEST INFRASTRUCTURE, never imported by the product.

readPair / postprocess of task1/data/octnpy.py:95-129 with the dataset tables of octnpy.py:70-89 and the tensor conversion of
task1/data/octgen.py:124-126, for the datasets whose prep_tran is alb.Resize(INTER_NEAREST).  alb.Resize is cv2.resize; cv2's
INTER_NEAREST maps destination index d to min(floor(d * src / dst), src - 1).  Pinned by tests/golden/prep_*.npz, which
oracle/make_golden_prep.py generated with cv2 itself (the reference's own dependency)."""
import numpy as np

SETS = {"hcms": (0, 1024, (256, 512), (128, 1024)), "hcms1": (0, 1024, (256, 512), (128, 1024)),
        "goals": (0, 608, (608, 512), (608, 1100)), "odsgh": (0, 992, (496, 512), (992, 1024))}
DIVIDE = 30


def nearest_index(dst, src):
    d = np.arange(dst, dtype=np.float64)
    return np.minimum(np.floor(d * (float(src) / float(dst))).astype(np.int64), src - 1)


def resize_nearest(a, H, W):
    return a[nearest_index(H, a.shape[0])][:, nearest_index(W, a.shape[1])]


def read_pair(dbname, img, lab):
    """octnpy.py:117-129 + octgen.py:124-126: (float32 [3,H,W] in [0,1], uint8 [H,W])."""
    stt, end, (H, W), _ = SETS[dbname]
    img = img[stt:end]
    lab = (lab // DIVIDE)[stt:end]
    img = resize_nearest(img, H, W)
    lab = resize_nearest(lab, H, W)
    x = np.clip(img.transpose(2, 0, 1).astype(np.float32) / 255, 0, 1)
    return x, lab.astype(np.uint8)


def postprocess(dbname, lab, raw_height):
    """octnpy.py:95-112: uint8 frame [raw_height, Wpost]."""
    stt, end, _, (Ho, Wo) = SETS[dbname]
    img = (lab.astype(np.int64) * DIVIDE).astype(np.uint8)
    img = resize_nearest(img, Ho, Wo)
    out = np.zeros((raw_height, Wo), dtype=np.uint8)
    out[stt:stt + Ho] = img
    return out
