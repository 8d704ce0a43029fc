"""GPU: the training loop (KiteSeg.train_step = calc_loss + backward + clip_grad_norm_(12) + AdamW) against the
oracle's OracleTrainer (kite/loop_seg.py:108-171, kite/loopback.py:102-128 restated), and CUDA-graph replay
against eager execution of the same steps."""
import argparse
import contextlib
import io

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import tcct_oracle as orc  # noqa: E402
from helpers import dp_masks, golden_state  # noqa: E402
from tcct_b200 import ops as O  # noqa: E402
from tcct_b200.kite.loop_seg import KiteSeg  # noqa: E402
from tcct_b200.nets import RegNet, stc_tt  # noqa: E402
from tcct_b200.nets.tcct import MHCABlock  # noqa: E402
from tcct_b200.synth import SynthOCT, make_bscans  # noqa: E402


def make_seg(C, seed, tmp, graph, udh=True, reg=True, H=64, W=64, B=2):
    with contextlib.redirect_stdout(io.StringIO()):
        net = RegNet(stc_tt(C), out_channels=C)
        net.load_state_dict(golden_state(C, seed), strict=True)
        args = argparse.Namespace(los="di", lr=1e-2, gpu="0", pl=False, bs=B, bug=False, udh=udh, coff_udh=1.0, reg=reg,
                                  coff_reg=0.1, epl=False, coff_epl=0.1, coff_ds=1.0, graph=graph)
        seg = KiteSeg(args, model=net, dataset=SynthOCT("goals" if C == 5 else "hcms", H, W, 2), root=str(tmp))
    seg.model.train()
    return seg


@pytest.mark.parametrize("C,K", [(5, 4), (9, 9)])
def test_two_train_steps_match_oracle_trainer(tmp_path, C, K):
    torch.set_num_threads(8)
    B, H, W, seed = 2, 64, 64, 17
    O.set_precision("tf32x3")
    try:
        seg = make_seg(C, seed, tmp_path, graph=False)
        lr = 1e-4
        seg.optimG.param_groups[0]["lr"] = lr
        seg.optimG.sync_lr()
        P = golden_state(C, seed)
        tr = orc.OracleTrainer(P, lr=lr)
        gen = torch.Generator().manual_seed(seed + 100)
        for step in range(2):
            img, lab = make_bscans(B, H, W, C, K, seed + step)
            onehot = F.one_hot(lab, C).permute(0, 3, 1, 2)
            noise = orc.make_noise(B, C, H, W, gen)
            masks = dp_masks(B, gen)
            total, parts, gnorm = tr.step(img, onehot, noise, [m.clone() for m in masks])
            MHCABlock.dp_tape = [m.clone() for m in masks]
            RegNet.noise_tape = noise
            got = seg.train_step(img, lab).cpu().tolist()
            MHCABlock.dp_tape = None
            want = [parts["los"], parts["udh"], parts["reg"], total]
            for name, g, w in zip(("los", "udh", "reg", "total"), got, want):
                assert abs(g - w) <= 1e-3 * max(abs(w), 1e-3), (step, name, g, w)
            # 3xTF32 products carry ~2e-6 relative error (tensor-core accumulation), which the batch-norm backward of the
            # CrossResNet branch amplifies to a few 1e-3 on its gradients (scripts/diag_parts.py, scripts/diag_x3.py)
            assert abs(seg.optimG.last_grad_norm() - gnorm) <= 1e-2 * gnorm, (seg.optimG.last_grad_norm(), gnorm)
        # weights after two AdamW steps: the first steps move every element by ~lr*sign(g), so compare the mean
        # displacement error in units of lr.  Tensors whose gradient is exactly zero in exact arithmetic (conv biases
        # feeding a BatchNorm) hold only round-off on both sides and random-walk by +-lr: they are skipped.
        sd = seg.model.state_dict()
        gmax = max(float(P[k].grad.abs().max()) for k in tr.keys if P[k].grad is not None)
        live = [k for k in tr.keys if P[k].grad is not None and float(P[k].grad.abs().max()) > 1e-4 * gmax]
        assert len(live) > 200
        worst = max((float((sd[k].cpu() - P[k].detach()).abs().mean()) / lr, k) for k in live)
        assert worst[0] <= 0.25, worst
        for k in ("base.base_cnn.cnn.1.running_var", "lap_map.1.running_mean", "base.dec4.prep.1.running_mean"):
            ref = P[k]
            assert float((sd[k].cpu() - ref).abs().max()) <= 2e-3 * float(ref.abs().max()) + 1e-6, k
        assert int(sd["lap_map.1.num_batches_tracked"]) == 4
    finally:
        O.set_precision("tf32")
        MHCABlock.dp_tape = None
        RegNet.noise_tape = None


def test_unused_parameters_are_never_touched(tmp_path):
    """Parameters that receive no gradient (crpe, cls_head, fuse, lap_epl, tau; lap_reg/lap_map when --reg=0) keep
    their values bit for bit -- torch's AdamW skips .grad=None tensors, weight decay included."""
    seg = make_seg(5, 3, tmp_path, graph=False, udh=True, reg=False)
    before = {k: v.clone() for k, v in seg.model.state_dict().items()}
    img, lab = make_bscans(2, 64, 64, 5, 4, 3)
    seg.train_step(img, lab)
    after = seg.model.state_dict()
    frozen = [k for k in before if any(u in k for u in (".crpe.", "cls_head.", "base.fuse.", "lap_epl.", "tau", "lap_reg.", "lap_map.", "fcp."))]
    assert frozen
    for k in frozen:
        assert torch.equal(before[k], after[k]), k
    moved = [k for k in before if k.endswith("block12.0.weight")]
    assert all(not torch.equal(before[k], after[k]) for k in moved)


def test_cuda_graph_replay_equals_eager(tmp_path):
    B, H, W, C, K = 2, 64, 64, 5, 4
    batches = [make_bscans(B, H, W, C, K, 100 + i) for i in range(6)]
    finals = []
    for graph in (False, True):
        seg = make_seg(C, 5, tmp_path / str(graph), graph=graph, udh=True, reg=False)
        for m in seg.model.modules():                      # DropPath / noise draws differ between capture and eager
            if isinstance(m, MHCABlock):
                m.drop_rate = 0.0
        losses = [seg.train_step(img, lab).cpu().tolist() for img, lab in batches]
        if graph:
            assert seg._graphs and all(g.graph is not None for g in seg._graphs.values())
        finals.append((losses, {k: v.clone() for k, v in seg.model.state_dict().items()}))
    (l0, s0), (l1, s1) = finals
    for a, b in zip(l0, l1):
        assert abs(a[3] - b[3]) <= 1e-3 * abs(a[3]), (a, b)
    for k in s0:
        if s0[k].dtype.is_floating_point:
            assert float((s0[k] - s1[k]).abs().max()) <= 5e-3 * float(s0[k].abs().max()) + 1e-5, k
        else:
            assert torch.equal(s0[k], s1[k]), k


def test_validation_counts_and_predict(tmp_path):
    seg = make_seg(5, 7, tmp_path, graph=False)
    logs = seg.val(epoch=0)
    assert 0.0 <= logs["val_iou"] <= 1.0 and 0.0 <= logs["val_f1s"] <= 1.0
    img, lab = make_bscans(1, 64, 64, 5, 4, 9)
    onehot_pred = seg.predict(img)
    lab8 = seg.predict_labels(img)
    assert torch.equal(onehot_pred.argmax(1).to(torch.uint8), lab8)
    # reference scores (kite/losses/miou.py:28-44,69-91) from the one-hot maps
    true = F.one_hot(lab, 5).permute(0, 3, 1, 2).float().cuda()

    def score(pr, gt, iou):
        pr, gt = pr.reshape(1, -1), gt.reshape(1, -1)
        inter = (pr * gt).sum(-1)
        return float(((inter + 1) / (pr.sum(-1) + gt.sum(-1) - inter + 1)).mean()) if iou else float(((2 * inter + 1) / (pr.sum(-1) + gt.sum(-1) + 1)).mean())
    from tcct_b200.kite.losses.miou import MDiceLoss, MIouLoss, label_counts
    counts = label_counts(lab8, O.labels_u8(lab.cuda(), 5), 5).cpu()
    f1, _ = MDiceLoss.from_counts(counts, 1)
    iou, _ = MIouLoss.from_counts(counts, 1)
    f1_ref = sum(score(onehot_pred[:, i], true[:, i], False) for i in range(1, 5)) / 4
    iou_ref = sum(score(onehot_pred[:, i], true[:, i], True) for i in range(1, 5)) / 4
    assert abs(float(f1) - f1_ref) < 1e-6 and abs(float(iou) - iou_ref) < 1e-6


@pytest.mark.parametrize("name", ["soft", "hard"])
def test_validation_score_api_matches_reference(name):
    """MDiceLoss.score / scores / scorem and MIouLoss.score / scorem (kite/losses/miou.py:28-44,69-91) on the kernel path against values
    written by the unmodified reference (oracle/make_golden_miou.py), for soft maps and hard one-hot maps, float and int64 targets."""
    import numpy as np
    from helpers import load, miou_inputs
    from tcct_b200.kite.losses.miou import MDiceLoss, MIouLoss
    g = load("miou_scores")
    pr, gt = miou_inputs(g[name + "::meta"])
    prd = pr.cuda()
    for gtd in (gt.cuda(), gt.float().cuda()):
        assert abs(float(MDiceLoss.scorem(prd, gtd)) - float(g[name + "::dice_scorem0"])) < 1e-5
        assert abs(float(MDiceLoss.scorem(prd, gtd, 1)) - float(g[name + "::dice_scorem1"])) < 1e-5
        assert abs(float(MIouLoss.scorem(prd, gtd)) - float(g[name + "::iou_scorem0"])) < 1e-5
        assert abs(float(MIouLoss.scorem(prd, gtd, 1)) - float(g[name + "::iou_scorem1"])) < 1e-5
        assert np.allclose(MDiceLoss.scores(prd, gtd), g[name + "::dice_scores"], atol=1e-5)
        assert abs(float(MDiceLoss.score(prd, gtd)) - float(g[name + "::dice_score"])) < 1e-5
        assert abs(float(MIouLoss.score(prd, gtd)) - float(g[name + "::iou_score"])) < 1e-5


@pytest.mark.parametrize("name", ["stc_tb", "gtc_tt", "gtc_tb"])
def test_gated_and_wide_models_train_under_graph(tmp_path, name):
    """`--net=stc_tb / gtc_tt / gtc_tb` through KiteSeg at 2x128x256 (tcgen05 routes on the 32-channel layers, sliced warp-level convs on
    the wide ones, GateFusion's on-device random field): eager warm-up, CUDA-graph capture and replay; the loss is finite, falls on a
    repeated batch, and every trained tensor moves."""
    import tcct_b200.nets as N
    B, H, W, C, K = 2, 128, 256, 5, 4
    with contextlib.redirect_stdout(io.StringIO()):
        net = RegNet(getattr(N, name)(C), out_channels=C)
        args = argparse.Namespace(los="di", lr=2e-3, gpu="0", pl=False, bs=B, bug=False, udh=True, coff_udh=1.0, reg=True,
                                  coff_reg=0.1, epl=False, coff_epl=0.1, coff_ds=1.0, graph=True)
        seg = KiteSeg(args, model=net, dataset=SynthOCT("goals", H, W, 2), root=str(tmp_path))
    seg.model.train()
    before = {k: v.clone() for k, v in seg.model.state_dict().items()}
    img, lab = make_bscans(B, H, W, C, K, 77)
    losses = [float(seg.train_step(img, lab)[3]) for _ in range(8)]
    assert seg._graphs and all(g.graph is not None for g in seg._graphs.values())
    assert all(l == l and abs(l) < 1e4 for l in losses), losses
    assert min(losses[4:]) < losses[0], losses
    after = seg.model.state_dict()
    for k in ("base.base_cnn.path_estan.4.block5.0.weight", "base.base_cnn.path_estan.1.block34.1.weight", "base.tran_cnn3.0.weight",
              "base.dec1.prep.0.weight", "base.base_vit.mhca_stages.2.aggregate.conv.weight"):
        assert not torch.equal(before[k], after[k]), k
