"""Input / deploy formats: the numpy oracle against the cv2-made golden vectors (CPU), and the CUDA
kernels against the oracle, bit for bit (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import prep_oracle as PO  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
RAW = {"goals": (800, 1100), "hcms": (496, 1024)}


def frames(db):
    """The frames oracle/make_golden_prep.py drew (same generator, same order)."""
    Hr, Wr = RAW[db]
    H, W = PO.SETS[db][2]
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (Hr, Wr, 3), dtype=np.uint8)
    lab = (np.minimum((np.arange(Hr)[:, None] * 9 // Hr + rng.integers(0, 2, (Hr, Wr))), 8) * 30).astype(np.uint8)
    pred = rng.integers(0, 9, (H, W), dtype=np.uint8)
    return img, lab, pred


@pytest.mark.parametrize("db", ["goals", "hcms"])
def test_oracle_matches_cv2_golden(db):
    g = np.load(os.path.join(GOLDEN, "prep_%s.npz" % db))
    Hr, Wr = RAW[db]
    stt, end, (H, W), (Ho, Wo) = PO.SETS[db]
    rows = min(end, Hr) - stt
    np.testing.assert_array_equal(PO.nearest_index(H, rows), g["sy"])
    np.testing.assert_array_equal(PO.nearest_index(W, Wr), g["sx"])
    np.testing.assert_array_equal(PO.nearest_index(Ho, H), g["py"])
    np.testing.assert_array_equal(PO.nearest_index(Wo, W), g["px"])
    img, lab, pred = frames(db)
    x, l = PO.read_pair(db, img, lab)
    assert x.shape == (3, H, W) and l.shape == (H, W) and x.dtype == np.float32 and l.dtype == np.uint8
    np.testing.assert_array_equal(x[:, :, 100:164], g["x_win"])
    np.testing.assert_array_equal(l[:, 100:164], g["lab_win"])
    assert float(x.astype(np.float64).sum()) == float(g["x_sum"]) and int(l.astype(np.int64).sum()) == int(g["lab_sum"])
    post = PO.postprocess(db, pred, Hr)
    np.testing.assert_array_equal(post[:, 200:264], g["post_win"])
    assert int(post.astype(np.int64).sum()) == int(g["post_sum"])


def test_padding_datasets_are_out_of_scope():
    pytest.importorskip("torch")
    if not os.path.exists(os.path.join(ROOT, "tcct_b200", "lib", "libtcct_b200.so")):
        pytest.skip("library not built")
    from tcct_b200.data import EyeSetResource
    with pytest.raises(NotImplementedError):
        EyeSetResource("duke", device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("db", ["goals", "hcms"])
def test_kernels_match_oracle_bit_exact(db):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tcct_b200.data import EyeSetResource
    res = EyeSetResource(db, device="cuda:0")
    img, lab, pred = frames(db)
    Hr = RAW[db][0]
    x_ref, l_ref = PO.read_pair(db, img, lab)
    out = res.readPair(img, lab)
    assert torch.equal(out["lab"].cpu(), torch.from_numpy(l_ref))
    assert torch.equal(out["img"].cpu(), torch.from_numpy(x_ref))                 # uint8 / 255 in fp32: bit-exact
    # a batch of two frames, the second flipped, through the same launch
    b_img = np.stack([img, img[:, ::-1].copy()]); b_lab = np.stack([lab, lab[:, ::-1].copy()])
    outb = res.readPair(b_img, b_lab)
    x2, l2 = PO.read_pair(db, b_img[1], b_lab[1])
    assert torch.equal(outb["img"][0].cpu(), torch.from_numpy(x_ref)) and torch.equal(outb["img"][1].cpu(), torch.from_numpy(x2))
    assert torch.equal(outb["lab"][1].cpu(), torch.from_numpy(l2))
    post = res.postprocess(torch.from_numpy(pred), Hr)
    assert torch.equal(post.cpu(), torch.from_numpy(PO.postprocess(db, pred, Hr)))
    # ragged / error cases
    with pytest.raises(RuntimeError):
        res.readPair(img, lab[:-1])
    with pytest.raises(TypeError):
        res.readPair(img.astype(np.float32), lab)
