"""Golden vectors of the data path, made with cv2 (what alb.Resize calls): tests/golden/prep_<db>.npz.
    python oracle/make_golden_prep.py
Full-size frames would be megabytes; the fixtures keep (1) the index maps cv2's INTER_NEAREST produces for the real raw ->
network sizes of every dataset (from ramp images) and (2) complete input / output pairs of a frame whose width is scaled
down."""
import os
import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW = {"goals": (800, 1100), "hcms": (496, 1024)}      # raw frame sizes (H, W) of the BASELINE configs
SETS = {"hcms": (0, 1024, (256, 512), (128, 1024)), "goals": (0, 608, (608, 512), (608, 1100))}


def cv_resize(a, H, W):
    return cv2.resize(a, (W, H), interpolation=cv2.INTER_NEAREST)


for db, (Hr, Wr) in RAW.items():
    stt, end, (H, W), (Ho, Wo) = SETS[db]
    rows = min(end, Hr) - stt
    rng = np.random.default_rng(5)
    # (1) index maps for the real sizes: ramps that encode the source coordinate in 16 bits
    ry = np.repeat(np.arange(rows, dtype=np.uint16)[:, None], 8, 1)
    rx = np.repeat(np.arange(Wr, dtype=np.uint16)[None, :], 8, 0)
    sy = cv_resize(ry, H, 8)[:, 0].astype(np.int64)
    sx = cv_resize(rx, 8, W)[0].astype(np.int64)
    py = cv_resize(np.repeat(np.arange(H, dtype=np.uint16)[:, None], 8, 1), Ho, 8)[:, 0].astype(np.int64)
    px = cv_resize(np.repeat(np.arange(W, dtype=np.uint16)[None, :], 8, 0), 8, Wo)[0].astype(np.int64)
    # (2) a complete pair on the real frame size, stored compressed (speckle image, banded labels)
    img = rng.integers(0, 256, (Hr, Wr, 3), dtype=np.uint8)
    lab = (np.minimum((np.arange(Hr)[:, None] * 9 // Hr + rng.integers(0, 2, (Hr, Wr))), 8) * 30).astype(np.uint8)
    ci = cv_resize(img[stt:end], H, W)
    cl = cv_resize((lab // 30)[stt:end], H, W)
    x = np.clip(ci.transpose(2, 0, 1).astype(np.float32) / 255, 0, 1)
    pred = rng.integers(0, 9, (H, W), dtype=np.uint8)
    post = np.zeros((Hr, Wo), np.uint8)
    post[stt:stt + Ho] = cv_resize((pred.astype(np.int64) * 30).astype(np.uint8), Ho, Wo)
    # keep the fixture small: only a 64-column window of the full-size tensors, plus checksums of the whole
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "prep_%s.npz" % db), sy=sy, sx=sx, py=py, px=px,
                        img_seed=5, x_sum=np.float64(x.astype(np.float64).sum()), x_win=x[:, :, 100:164],
                        lab_sum=np.int64(cl.astype(np.int64).sum()), lab_win=cl[:, 100:164].astype(np.uint8),
                        post_sum=np.int64(post.astype(np.int64).sum()), post_win=post[:, 200:264])
    print(db, "rows", rows, "->", (H, W), "post", (Ho, Wo), "x_sum", float(x.sum()))
