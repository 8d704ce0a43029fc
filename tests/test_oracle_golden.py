"""CPU: the oracle restatement reproduces the reference's golden vectors
(tests/golden/*.npz, written by oracle/make_golden.py from the unmodified reference)."""
import os

import numpy as np
import pytest
import torch

import tcct_oracle as O
from helpers import GOLDEN, golden_state, load, train_inputs
from tcct_b200.synth import make_bscans


@pytest.mark.parametrize("name", ["train_goals_64", "train_hcms_64x128"])
def test_train_step_matches_reference(name):
    torch.set_num_threads(8)
    g = load(name)
    n_class, seed = int(g["meta"][0]), int(g["meta"][5])
    img, lab, onehot, noise, masks = train_inputs(g["meta"])
    P = golden_state(n_class, seed)
    tr = O.OracleTrainer(P, lr=1e-4)
    tr.opt.zero_grad()
    total, parts, outs, feats = O.calc_loss(P, img, onehot, O.Ctx(True, masks), noise)
    total.backward()
    gnorm = torch.nn.utils.clip_grad_norm_([P[k] for k in tr.keys], 12)
    np.testing.assert_allclose(outs[0].detach().numpy(), g["out0"], rtol=0, atol=1e-4 * np.abs(g["out0"]).max())
    for i in (1, 2, 3):
        np.testing.assert_allclose(outs[i].detach()[:, :, ::4, ::4].numpy(), g["out%d_sub" % i], rtol=0,
                                   atol=1e-4 * np.abs(g["out0"]).max())
    np.testing.assert_allclose(feats.detach()[:, :, ::4, ::4].numpy(), g["feats_sub"], atol=1e-5)
    got = [float(parts["los"]), float(parts["udh"]), float(parts["reg"]), float(total)]
    np.testing.assert_allclose(got, g["loss"], rtol=1e-5)
    assert abs(float(gnorm) - float(g["gnorm"])) < 1e-4 * float(g["gnorm"])
    keys = [str(k) for k in g["grad_keys"]]
    assert sorted(k for k in tr.keys if P[k].grad is not None) == keys
    norms = np.array([float(P[k].grad.norm()) for k in keys])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-2, atol=1e-3 * g["grad_norms"].max())
    for k in g.files:
        if k.startswith("grad::"):
            ref = g[k]
            np.testing.assert_allclose(P[k[6:]].grad.numpy(), ref, rtol=0, atol=2e-3 * np.abs(ref).max() + 1e-6)
    tr.opt.step()
    for k in g.files:
        if k.startswith("after::") and "running" in k:
            np.testing.assert_allclose(P[k[7:]].detach().numpy(), g[k], rtol=1e-5, atol=1e-6)
        elif k.startswith("after::") and "num_batches" in k:
            assert int(P[k[7:]]) == int(g[k])
    assert int(P["lap_map.1.num_batches_tracked"]) == 2      # reg.py:128-129 runs the BN twice


@pytest.mark.parametrize("name", ["eval_goals_96x64", "eval_hcms_64"])
def test_eval_matches_reference(name):
    g = load(name)
    n_class, n_bound, batch, height, width, seed = (int(v) for v in g["meta"])
    img, _ = make_bscans(batch, height, width, n_class, n_bound, seed)
    out0, labels = O.predict_labels(golden_state(n_class, seed), img)
    np.testing.assert_allclose(out0.numpy(), g["out0"], rtol=0, atol=1e-5 * np.abs(g["out0"]).max())
    assert np.array_equal(labels.numpy().astype(np.uint8), g["labels"])
    np.testing.assert_allclose(O.soft_argmax(out0).numpy(), g["soft_argmax"], atol=1e-4)


def test_metapool_is_token_channel_plane_pool():
    """nets/tcct.py:412-415 applied to [B,N,C]: window spans neighbouring tokens AND channels."""
    t = torch.arange(2 * 4 * 3, dtype=torch.float32).view(2, 4, 3)
    got = O.meta_pool(t)
    exp = torch.zeros_like(t)
    for b in range(2):
        for n in range(4):
            for c in range(3):
                vals = [t[b, i, j] for i in range(max(0, n - 1), min(4, n + 2)) for j in range(max(0, c - 1), min(3, c + 2))]
                exp[b, n, c] = sum(vals) / len(vals) - t[b, n, c]
    assert torch.allclose(got, exp)


def test_feature_polar_nan_when_class_under_32_pixels():
    """nets/fcs.py:36: N = count//32 = 0 -> mean of empty -> NaN (Appendix A)."""
    P = {"fcp.buf_grad": torch.nn.functional.normalize(torch.rand(3, 32), dim=-1)}
    feat, logits = torch.randn(1, 32, 8, 8), torch.randn(1, 3, 8, 8)
    lab = torch.zeros(1, 8, 8, dtype=torch.long)
    lab[0, 0, :5] = 1
    lab[0, 4:, :] = 2
    onehot = torch.nn.functional.one_hot(lab, 3).permute(0, 3, 1, 2)
    assert torch.isnan(O.feature_polar(P, feat, logits, onehot))


def test_boundary_positions_of_a_synthetic_step():
    """The (unpinned) soft-argmax boundary extraction puts a sharp class transition where it is: a two-class map that switches
    from class 0 to class 1 at row 20 has its class-1 probability jump between rows 19 and 20."""
    import torch
    H, W = 48, 5
    logits = torch.full((1, 2, H, W), -8.0)
    logits[0, 0, :20] = 8.0
    logits[0, 1, 20:] = 8.0
    pos = O.boundary_positions(logits, beta=100.0)
    assert pos.shape == (1, 1, W)
    assert float((pos - 20.0).abs().max()) < 1e-3


def test_cnnu_matches_reference():
    """The CNN-only factory (nets/tcct.py:1124-1129): eval logits / labels and train-mode gradients of the oracle against vectors
    made by the unmodified reference (oracle/make_golden_cnnu.py)."""
    import torch
    import torch.nn.functional as F
    g = load("cnnu_goals_64")
    n_class, n_bound, batch, height, width, seed = (int(v) for v in g["meta"])
    img, lab = make_bscans(batch, height, width, n_class, n_bound, seed)
    out0, labels = O.predict_labels(golden_state(n_class, seed), img, flag_vit=False)
    np.testing.assert_allclose(out0.numpy(), g["out0"], rtol=0, atol=1e-5 * np.abs(g["out0"]).max())
    assert np.array_equal(labels.numpy().astype(np.uint8), g["labels"])
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in golden_state(n_class, seed).items()}
    outs, _ = O.ftc_forward(P, img, O.Ctx(True), flag_vit=False)
    onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2)
    loss = O.multi_dice(outs[0], onehot) + sum(o.mean() for o in outs[1:])
    loss.backward()
    assert abs(float(loss.detach()) - float(g["train_loss"])) <= 1e-5 * abs(float(g["train_loss"]))
    for k in [k for k in g.files if k.startswith("grad::")]:
        ref = g[k]
        np.testing.assert_allclose(P["base." + k[6:]].grad.numpy(), ref, rtol=0, atol=1e-4 * np.abs(ref).max())
    assert P["base.tran_cnn0.0.weight"].grad is None          # the fusion convs never run


def test_oracle_matches_real_weight_known_answer():
    """The shipped trained checkpoint (onnx/tcct_duke.pt) on the shipped B-scan (onnx/oct_duke.png[:224,:512]): logits, label map,
    label histogram and column profile written by the unmodified reference (oracle/make_golden_real.py)."""
    import hashlib
    g = np.load(os.path.join(GOLDEN, "real_duke.npz"))
    pt = os.path.join(GOLDEN, "tcct_duke.pt")
    assert hashlib.md5(open(pt, "rb").read()).hexdigest() == str(g["md5"])
    P = torch.load(pt, map_location="cpu")
    img = torch.from_numpy(g["image"]).float().div(255)[None, None].expand(1, 3, -1, -1).contiguous()
    logits, labels = O.predict_labels(P, img)
    ref = torch.from_numpy(g["logits_sub"])
    assert float((logits[0, :, :, ::4] - ref).abs().max()) <= 1e-5 * float(g["logit_absmax"])
    assert np.array_equal(labels[0].numpy().astype(np.uint8), g["labels"])
    assert np.bincount(g["labels"].reshape(-1), minlength=int(g["n_class"])).tolist() == g["hist"].tolist()
    assert abs(float(logits.double().sum()) - float(g["logit_sum"])) <= 1e-6 * abs(float(g["logit_sum"]))


@pytest.mark.parametrize("name", ["pnnu", "vitu"])
def test_variant_factories_match_reference(name):
    """`pnnu` (PlainCNNBlock, nets/tcct.py:830-855,1117-1122) and `vitu` (1131-1136): the oracle against vectors made by the unmodified
    reference (oracle/make_golden_variants.py): eval logits / labels, train loss and gradients."""
    import torch.nn.functional as F
    from helpers import VARIANT_KW, dp_masks, variant_state
    g = load("%s_goals_64" % name)
    n_class, n_bound, batch, height, width, seed = (int(v) for v in g["meta"])
    img, lab = make_bscans(batch, height, width, n_class, n_bound, seed)
    out0, labels = O.predict_labels(variant_state(name, n_class, seed), img, **VARIANT_KW[name])
    np.testing.assert_allclose(out0.numpy(), g["out0"], rtol=0, atol=1e-5 * np.abs(g["out0"]).max())
    assert np.array_equal(labels.numpy().astype(np.uint8), g["labels"])
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in variant_state(name, n_class, seed).items()}
    masks = dp_masks(batch, torch.Generator().manual_seed(seed + 100))
    outs, _ = O.ftc_forward(P, img, O.Ctx(True, masks), **VARIANT_KW[name])
    onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2)
    loss = O.multi_dice(outs[0], onehot) + sum(o.mean() for o in outs[1:])
    loss.backward()
    assert abs(float(loss.detach()) - float(g["train_loss"])) <= 1e-5 * abs(float(g["train_loss"]))
    for k in [k for k in g.files if k.startswith("grad::")]:
        np.testing.assert_allclose(P["base." + k[6:]].grad.numpy(), g[k], rtol=0, atol=1e-4 * np.abs(g[k]).max())


@pytest.mark.parametrize("name", ["gtc_tt", "stc_tb", "gtc_tb"])
def test_gated_and_wide_factories_match_reference(name):
    """`gtc_tt` / `gtc_tb` (GateFusion, nets/tcct.py:916-932,1050-1061) and `stc_tb` / `gtc_tb` (the 32-64-96-128-256 CrossResNet,
    861-864,1097-1102): the oracle against vectors made by the unmodified reference (oracle/make_golden_wide.py) -- eval logits / labels,
    train loss and gradients with the reference's own recorded gate fields."""
    import torch.nn.functional as F
    from helpers import dp_masks, factory_state
    g = load("%s_goals_64" % name)
    gate = name.startswith("gtc")
    n_class, n_bound, batch, height, width, seed = (int(v) for v in g["meta"])
    img, lab = make_bscans(batch, height, width, n_class, n_bound, seed)
    state = {"base." + k: v for k, v in factory_state(name, n_class, seed).items()}
    out0, labels = O.predict_labels(state, img, gate=gate)
    np.testing.assert_allclose(out0.numpy(), g["out0"], rtol=0, atol=1e-5 * np.abs(g["out0"]).max())
    assert np.array_equal(labels.numpy().astype(np.uint8), g["labels"])
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in state.items()}
    masks = dp_masks(batch, torch.Generator().manual_seed(seed + 100))
    alphas = [torch.from_numpy(g["alpha%d" % i]) for i in range(4)] if gate else None
    outs, _ = O.ftc_forward(P, img, O.Ctx(True, masks, gate_alphas=alphas), gate=gate)
    onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2)
    loss = O.multi_dice(outs[0], onehot) + sum(o.mean() for o in outs[1:])
    loss.backward()
    assert abs(float(loss.detach()) - float(g["train_loss"])) <= 1e-5 * abs(float(g["train_loss"]))
    for k in [k for k in g.files if k.startswith("grad::")]:
        np.testing.assert_allclose(P["base." + k[6:]].grad.numpy(), g[k], rtol=0, atol=1e-4 * np.abs(g[k]).max())


@pytest.mark.parametrize("tag", ["goals", "hcms"])
def test_oracle_matches_real_weight_onnx_variant(tag):
    """The deploy-time model definition (onnx/tcct_{goals,hcms}.py: older decoder tail) with the shipped trained checkpoints on the
    reference's own B-scan, against logits / labels written by the unmodified reference (oracle/make_golden_variants.py)."""
    import hashlib
    g = np.load(os.path.join(GOLDEN, "real_%s.npz" % tag))
    pt = os.path.join(GOLDEN, "tcct_%s.pt" % tag)
    assert hashlib.md5(open(pt, "rb").read()).hexdigest() == str(g["md5"])
    P = torch.load(pt, map_location="cpu")
    image = np.load(os.path.join(GOLDEN, "real_duke.npz"))["image"]
    img = torch.from_numpy(image).float().div(255)[None, None].expand(1, 3, -1, -1).contiguous()
    logits, labels = O.predict_labels(P, img, variant="onnx")
    assert float((logits[:, :, :, ::4] - torch.from_numpy(g["logits_sub"])).abs().max()) <= 1e-5 * float(g["logit_absmax"])
    assert np.array_equal(labels.numpy().astype(np.uint8), g["labels"])


@pytest.mark.parametrize("name", ["soft", "hard"])
def test_validation_scores_match_reference(name):
    """oracle.val_scores against MDiceLoss.scorem / scores / MIouLoss.scorem of the unmodified reference (oracle/make_golden_miou.py)."""
    from helpers import miou_inputs
    g = load("miou_scores")
    pr, gt = miou_inputs(g[name + "::meta"])
    dice, iou = O.val_scores(pr, gt)
    assert abs(float(dice.mean()) - float(g[name + "::dice_scorem0"])) < 1e-6 and abs(float(dice[1:].mean()) - float(g[name + "::dice_scorem1"])) < 1e-6
    assert abs(float(iou.mean()) - float(g[name + "::iou_scorem0"])) < 1e-6 and abs(float(iou[1:].mean()) - float(g[name + "::iou_scorem1"])) < 1e-6
    np.testing.assert_allclose(dice.numpy(), g[name + "::dice_scores"], atol=1e-6)
