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


def test_unknown_dataset_is_refused():
    pytest.importorskip("torch")
    if not os.path.exists(os.path.join(ROOT, "tcct_b200", "lib", "libtcct_b200.so")):
        pytest.skip("library not built")
    from tcct_b200.data import EyeSetResource
    with pytest.raises(NotImplementedError):
        EyeSetResource("drive", device="cpu")


# ----------------------------------------------------------------------------- make_tran (octgen.py:9-19)
import aug_oracle as AO  # noqa: E402

AUG_CASES = {"goals": (608, 512, 256, 256, 5), "hcms": (256, 512, 256, 256, 9), "small": (200, 230, 256, 256, 5)}


def aug_frames(db):
    """The frames oracle/make_golden_aug.py drew (same generator, same order)."""
    Hp, Wp, H, W, C = AUG_CASES[db]
    rng = np.random.default_rng(17)
    img = rng.integers(0, 256, (Hp, Wp, 3), dtype=np.uint8)
    img[:, : Wp // 3] = (img[:, : Wp // 3] // 8)
    lab = np.minimum(np.arange(Hp)[:, None] * C // Hp + rng.integers(0, 2, (Hp, Wp)), C - 1).astype(np.uint8)
    lab[:, : Wp // 5] = 0
    return img, lab


def aug_draws(g):
    keys = []
    for row in g["draws"]:
        keys.append({"y0": int(row[0]), "x0": int(row[1]), "hflip": bool(row[2]), "vflip": bool(row[3]), "rgb_shift": tuple(float(v) for v in row[4:7]),
                     "hue_shift": float(row[7]), "sat_shift": float(row[8]), "val_shift": float(row[9]), "contrast_alpha": float(row[10]),
                     "brightness_beta": float(row[11])})
    return keys


@pytest.mark.parametrize("db", ["goals", "hcms", "small"])
def test_augmentation_oracle_matches_cv2_golden(db):
    """oracle/aug_oracle.py (own 8-bit HSV restatement, numpy look-up tables) against outputs made with cv2.cvtColor / cv2.LUT
    (oracle/make_golden_aug.py, which also checks the HSV restatement against cv2 on every possible input)."""
    g = np.load(os.path.join(GOLDEN, "aug_%s.npz" % db))
    Hp, Wp, H, W, C = AUG_CASES[db]
    img, lab = aug_frames(db)
    for i, p in enumerate(aug_draws(g)):
        x, m = AO.make_tran_apply(img, lab, H, W, p)
        np.testing.assert_array_equal(x[:, 96:160, 96:160], g["x_win%d" % i])
        np.testing.assert_array_equal(m[96:160, 96:160], g["m_win%d" % i])
        np.testing.assert_array_equal(x.astype(np.float64).sum((0, 2)), g["x_rowsum%d" % i])
        assert int(m.astype(np.int64).sum()) == int(g["m_sum%d" % i])


def test_hsv_restatement_on_a_colour_lattice():
    """Round trip sanity of the 8-bit HSV restatement on a coarse lattice (the exhaustive check against cv2 runs in make_golden_aug.py)."""
    r = np.arange(0, 256, 5, dtype=np.uint8)
    rgb = np.stack(np.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 1, 3)
    hsv = AO.rgb2hsv_u8(rgb)
    assert hsv[..., 0].max() < 180
    back = AO.hsv2rgb_u8(hsv).astype(np.int64)
    assert np.abs(back - rgb.astype(np.int64)).max() <= 6         # 8-bit HSV is lossy; greys come back exactly
    grey = np.repeat(np.arange(256, dtype=np.uint8)[:, None, None], 3, 2)
    np.testing.assert_array_equal(AO.hsv2rgb_u8(AO.rgb2hsv_u8(grey)), grey)


@pytest.mark.gpu
@pytest.mark.parametrize("db", ["goals", "hcms", "small"])
def test_augmentation_kernel_matches_golden_bit_exact(db):
    """csrc/prep.cu: prep_augment_kernel (pad + crop + flips + RGB shift + HSV shift + contrast + brightness + /255 in one launch) against
    the cv2-made golden windows / checksums and, pixel for pixel, the oracle."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tcct_b200.data import EyeSetResource, GpuTwist, read_pair_aug
    g = np.load(os.path.join(GOLDEN, "aug_%s.npz" % db))
    Hp, Wp, H, W, C = AUG_CASES[db]
    img, lab = aug_frames(db)
    draws = aug_draws(g)
    res = EyeSetResource("goals", device="cuda:0")
    res.height_stt, res.height_end, res.prep_size, res.divide = 0, Hp, (Hp, Wp), 1      # frames are already what readPair returns
    B = len(draws)
    out = read_pair_aug(res, np.stack([img] * B), np.stack([lab] * B), draws, GpuTwist(H, W))
    x, m = out["img"].cpu().numpy(), out["lab"].cpu().numpy()
    for i, p in enumerate(draws):
        np.testing.assert_array_equal(x[i][:, 96:160, 96:160], g["x_win%d" % i])
        np.testing.assert_array_equal(m[i][96:160, 96:160], g["m_win%d" % i])
        xo, mo = AO.make_tran_apply(img, lab, H, W, p)
        np.testing.assert_array_equal(x[i], xo)
        np.testing.assert_array_equal(m[i], mo)


@pytest.mark.gpu
def test_augmentation_from_raw_frames_and_sampler():
    """readPairAug on raw GOALS-sized frames (crop rows + label // 30 + nearest resize + make_tran in one launch) == the oracle's
    read_pair followed by make_tran_apply; draws come from GpuTwist.sample (windows stay inside the padded frame and contain mask)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tcct_b200.data import EyeSetResource, make_tran
    img, lab, _ = frames("goals")
    res = EyeSetResource("goals", device="cuda:0")
    stt, end, (Hp, Wp), _ = PO.SETS["goals"]
    prep_img = PO.resize_nearest(img[stt:end], Hp, Wp)
    prep_lab = PO.resize_nearest((lab // 30)[stt:end], Hp, Wp).astype(np.uint8)
    twist = make_tran(256, 256, seed=3)
    draws = [twist.sample(prep_lab) for _ in range(6)]
    for p in draws:
        assert 0 <= p["y0"] <= Hp - 256 and 0 <= p["x0"] <= Wp - 256
        assert prep_lab[p["y0"]:p["y0"] + 256, p["x0"]:p["x0"] + 256].any()
    out = res.readPairAug(np.stack([img] * 6), np.stack([lab] * 6), draws, twist)
    for i, p in enumerate(draws):
        xo, mo = AO.make_tran_apply(prep_img, prep_lab, 256, 256, p)
        np.testing.assert_array_equal(out["img"][i].cpu().numpy(), xo)
        np.testing.assert_array_equal(out["lab"][i].cpu().numpy(), mo)


@pytest.mark.gpu
@pytest.mark.parametrize("db,Hr,Wr", [("duke", 300, 500), ("heg", 400, 640), ("duke1", 224, 600), ("duke2", 384, 501), ("duke2", 300, 576)])
def test_padding_datasets_read_pair_and_postprocess(db, Hr, Wr):
    """readPair of the padding datasets (octnpy.py:56-66: rows [stt, end), alb.PadIfNeeded centred, zeros or cv2.BORDER_REFLECT) against
    cv2.copyMakeBorder, and postprocess (101-110: CenterCrop back to the label file's size, pasted into a zero frame) against numpy."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import cv2
    from tcct_b200.data import EyeSetResource
    from tcct_b200.data.octnpy import _PAD_SETS
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (Hr, Wr, 3), dtype=np.uint8)
    lab = (rng.integers(0, 9, (Hr, Wr)) * 30).astype(np.uint8)
    stt, end, (mh, mw), reflect = _PAD_SETS[db]

    def pad(a):
        rows, cols = a.shape[:2]
        top = int((mh - rows) / 2.0) if rows < mh else 0
        left = int((mw - cols) / 2.0) if cols < mw else 0
        return cv2.copyMakeBorder(np.ascontiguousarray(a), top, (mh - rows - top) if rows < mh else 0, left, (mw - cols - left) if cols < mw else 0,
                                  cv2.BORDER_REFLECT if reflect else cv2.BORDER_CONSTANT, value=0)
    ci, cl = pad(img[stt:end]), pad((lab // 30)[stt:end])
    res = EyeSetResource(db, device="cuda:0")
    out = res.readPair(img, lab)
    np.testing.assert_array_equal(out["img"].cpu().numpy(), np.clip(ci.transpose(2, 0, 1).astype(np.float32) / 255, 0, 1))
    np.testing.assert_array_equal(out["lab"].cpu().numpy(), cl.astype(np.uint8))
    # postprocess: the prediction has the padded size; the label file the raw one
    pred = rng.integers(0, 9, cl.shape, dtype=np.uint8)
    rows = min(end, Hr) - stt
    if rows == min(Hr, pred.shape[0]):          # the reference's `bgd[stt:end, :] = img` only works when the crop fills the row band
        h, w = min(Hr, pred.shape[0]), min(Wr, pred.shape[1])
        y1, x1 = (pred.shape[0] - h) // 2, (pred.shape[1] - w) // 2
        want = np.zeros((Hr, Wr), np.uint8)
        want[stt:stt + h] = (pred.astype(np.int64) * 30).astype(np.uint8)[y1:y1 + h, x1:x1 + w]
        got = res.postprocess(pred, Hr, raw_width=Wr)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    else:
        with pytest.raises(RuntimeError):
            res.postprocess(pred, Hr, raw_width=Wr)


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


def _write_dataset(root, db, n, Hr, Wr, C, seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "train_img"), exist_ok=True)
    os.makedirs(os.path.join(root, "train_lab"), exist_ok=True)
    frames = []
    for i in range(n):
        img = rng.integers(0, 256, (Hr, Wr, 3), dtype=np.uint8)
        lab = (np.minimum(np.arange(Hr)[:, None] * C // Hr + rng.integers(0, 2, (Hr, Wr)), C - 1) * 30).astype(np.uint8)
        lab[:, : Wr // 6] = 0
        cv2.imwrite(os.path.join(root, "train_img", "%03d.png" % i), img)
        cv2.imwrite(os.path.join(root, "train_lab", "%03d.png" % i), lab)
        frames.append((img, lab))
    return frames


@pytest.mark.gpu
@pytest.mark.parametrize("db,Hr,Wr,C", [("goals", 160, 220, 5), ("duke", 240, 300, 9), ("duke2", 300, 280, 9)])
def test_eye_set_generator_batches(tmp_path, db, Hr, Wr, C):
    """EyeSetGenerator (octgen.py:24-129) on a folder of PNGs: file listing, train batches = readPair + make_tran with the recorded
    draws (against the numpy oracle on what readPair returns), val batches = ALB_VALID flips, test batches = readPair."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tcct_b200.data import EyeSetGenerator
    frames = _write_dataset(str(tmp_path), db, 5, Hr, Wr, C)
    ds = EyeSetGenerator(db, folder=str(tmp_path), device="cuda:0", seed=5, height=128, width=128)
    assert ds.lens == {"train": 5, "val": 5, "test": 0} and ds.out_channels == C and ds.exeNums["train"] == 735 // 5
    assert len(ds.trainSet(bs=4)) == (5 * 147 + 3) // 4
    draws, orig = [], ds.twist.sample

    def sample(mask):
        d = orig(mask)
        draws.append((dict(d), mask.copy()))
        return d
    ds.twist.sample = sample
    it = iter(ds.trainSet(bs=3))
    batch = next(it)
    img, lab, tag, _ = ds.parse(batch)
    assert img.shape == (3, 3, 128, 128) and lab.shape == (3, 128, 128) and lab.dtype == torch.uint8 and len(tag) == 3
    by_name = {"%03d.png" % i: f for i, f in enumerate(frames)}
    for k in range(3):
        raw_img, raw_lab = by_name[os.path.basename(tag[k])]
        d, mask = draws[k]
        row0, rows, Hp, Wp, pt, pl = ds.prep_geometry(Hr, Wr)
        if ds.pad_size is None:
            prep_img = PO.resize_nearest(raw_img[row0:row0 + rows], Hp, Wp)
        else:
            mode = "symmetric" if ds.pad_reflect else "constant"
            prep_img = np.pad(raw_img[row0:row0 + rows], ((pt, Hp - rows - pt), (pl, Wp - Wr - pl), (0, 0)), mode=mode)
        assert mask.shape == (Hp, Wp)
        xo, mo = AO.make_tran_apply(prep_img, mask, 128, 128, d)
        np.testing.assert_array_equal(img[k].cpu().numpy(), xo)
        np.testing.assert_array_equal(lab[k].cpu().numpy(), mo)
        assert mo.any()                                           # CropNonEmptyMaskIfExists: the window holds labelled pixels
    vb = next(iter(ds.valSet(bs=2)))
    row0, rows, Hp, Wp, pt, pl = ds.prep_geometry(Hr, Wr)
    assert vb["img"].shape == (2, 3, Hp, Wp)
    for k in range(2):
        raw_img, raw_lab = by_name[os.path.basename(vb["tag"][k])]
        ref = ds.readPair(raw_img, raw_lab)
        flipped = torch.flip(ref["img"], dims=[2])
        got = vb["img"][k]
        assert torch.equal(got, flipped) or torch.equal(got, torch.flip(flipped, dims=[1]))      # HorizontalFlip(p=1), VerticalFlip(p=.5)


@pytest.mark.gpu
def test_kite_seg_trains_from_a_png_folder(tmp_path):
    """The whole input path: PNG files -> EyeSetGenerator (GPU readPair + augmentation) -> KiteSeg.train (one short epoch)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import argparse, contextlib, io
    from tcct_b200.data import EyeSetGenerator
    from tcct_b200.kite.loop_seg import KiteSeg
    from tcct_b200.nets import RegNet, stc_tt
    _write_dataset(str(tmp_path / "data"), "goals", 6, 200, 260, 5)
    with contextlib.redirect_stdout(io.StringIO()):
        ds = EyeSetGenerator("goals", folder=str(tmp_path / "data"), device="cuda:0", seed=1, height=64, width=128)
        ds.exeNums["train"] = 2
        net = RegNet(stc_tt(5), out_channels=5)
        # --udh=0: a crop window of this banded toy label holds 2-3 of the 5 classes, and feature polarisation is NaN for a class under
        # 32 pixels -- by the reference's own definition (nets/fcs.py:36; test_feature_polarisation_nan_when_class_has_under_32_pixels)
        args = argparse.Namespace(los="di", lr=1e-3, gpu="0", pl=False, bs=4, bug=False, udh=False, coff_udh=1.0, reg=True,
                                  coff_reg=0.1, epl=False, coff_epl=0.1, coff_ds=1.0, graph=True)
        seg = KiteSeg(args, model=net, dataset=ds, root=str(tmp_path / "exp"))
        loss = seg.train(0)
    assert loss == loss and 0 < loss < 1e4
