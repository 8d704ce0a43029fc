"""GPU: model-level parity AT THE BENCHMARKED CONFIGURATIONS (BASELINE.json configs 2 and 3: bs=8, 256x256 crops, C=5 / C=9)
in the shipped default precision (tf32), against the CPU oracle run live on the same seeded inputs, noise and DropPath masks.

Tolerances: north_star's for the forward (logits of all four heads and `feats` max|d| <= 1e-2 max|ref|, each loss term
1e-3 relative).  Gradients: every parameter tensor is held to a bound CALIBRATED by the oracle itself -- the oracle is run a
second time with the operands of every dense contraction truncated to TF32 (tests/helpers.py:TF32Emu, exactly what
tcgen05 kind::tf32 does with fp32 operands); a tensor passes when its relative L2 error is within GRAD_SLACK x the error of
that emulation (+ a floor).  The network at random initialisation in train mode is ill-conditioned (batch-norm backward
cancellations): the emulation itself deviates by ~1e-1 on the whole gradient, which is why a fixed tight number would be
meaningless and `cos >= 0.98` (round 1) was vacuous.  The test also asserts that the tcgen05 + TMA kernels served the step."""
import argparse
import contextlib
import io
import json
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import tcct_oracle as orc  # noqa: E402
from helpers import TF32Emu, dp_masks, golden_state, oracle_train_pass  # noqa: E402
from tcct_b200 import _lib as L  # noqa: E402
from tcct_b200 import ops as O  # noqa: E402
from tcct_b200.kite.loop_seg import KiteSeg  # noqa: E402
from tcct_b200.nets import RegNet, stc_tt  # noqa: E402
from tcct_b200.nets.tcct import MHCABlock  # noqa: E402
from tcct_b200.synth import SynthOCT, make_bscans  # noqa: E402

DEV = torch.device("cuda:0")
GRAD_SLACK, GRAD_FLOOR = 3.0, 2e-2
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.jsonl")


def report(**kw):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)


def make_seg(C, seed, tmp, B, H, W, graph=False):
    with contextlib.redirect_stdout(io.StringIO()):
        net = RegNet(stc_tt(C), out_channels=C)
        net.load_state_dict(golden_state(C, seed), strict=True)
        args = argparse.Namespace(los="di", lr=1e-2, gpu="0", pl=False, bs=B, bug=False, udh=True, coff_udh=1.0, reg=True,
                                  coff_reg=0.1, epl=False, coff_epl=0.1, coff_ds=1.0, graph=graph)
        seg = KiteSeg(args, model=net, dataset=SynthOCT("goals" if C == 5 else "hcms", H, W, 2), root=str(tmp))
    seg.model.train()
    return seg


def rl2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


@pytest.mark.parametrize("C,K", [(5, 4), (9, 9)])
def test_train_pass_at_bench_config(tmp_path, C, K):
    torch.set_num_threads(max(8, os.cpu_count() or 8))
    B, H, W, seed = 8, 256, 256, 23
    img, lab = make_bscans(B, H, W, C, K, seed)
    onehot = F.one_hot(lab, C).permute(0, 3, 1, 2)
    gen = torch.Generator().manual_seed(seed + 100)
    noise = orc.make_noise(B, C, H, W, gen)
    masks = dp_masks(B, gen)
    total, parts, outs, feats, P, keys = oracle_train_pass(C, seed, img, onehot, noise, masks)
    _, _, _, _, Pq, _ = oracle_train_pass(C, seed, img, onehot, noise, masks, quant=TF32Emu(truncate=True))

    seg = make_seg(C, seed, tmp_path, B, H, W)
    before = L.route_counts()
    MHCABlock.dp_tape = [m.clone() for m in masks]
    RegNet.noise_tape = noise
    try:
        seg.optimG.zero_grad()
        lab8 = seg._label_map(lab)
        g_total, g_parts = seg._losses(img.to(DEV), lab8)
        g_total.backward()
        torch.cuda.synchronize()
    finally:
        MHCABlock.dp_tape = None
        RegNet.noise_tape = None
    after = L.route_counts()
    routes = {k: after[k] - before[k] for k in after}
    # ---- the tcgen05 + TMA kernels served the step (full-resolution convs, 1x1 convs / Linear on the large maps)
    assert routes["conv_tma"] >= 20 and routes["wgrad_tma"] >= 10 and routes["gemm_tma"] >= 10 and routes["wgrad_gemm_tma"] >= 5, routes

    # ---- forward
    logit_err = [rel(seg.udh_out, outs[0])]
    feat_err = rel(seg.model.base.feats[0], feats)
    loss_err = {k: abs(float(g_parts[k].detach()) - float(parts[k].detach())) / max(abs(float(parts[k].detach())), 1e-6) for k in parts}

    # ---- gradients, tensor by tensor, against the TF32-emulation-calibrated bound
    named = dict(seg.model.named_parameters())
    gkeys = [k for k in keys if P[k].grad is not None]
    gnorm = {k: float(P[k].grad.double().norm()) for k in gkeys}
    gmax = max(gnorm.values())
    live = [k for k in gkeys if gnorm[k] > 1e-3 * gmax]
    bad, rows = [], []
    for k in live:
        mine = named[k].grad.detach().cpu()
        e_gpu, e_emu = rl2(mine, P[k].grad), rl2(Pq[k].grad, P[k].grad)
        rows.append((k, e_gpu, e_emu))
        if e_gpu > GRAD_SLACK * e_emu + GRAD_FLOOR:
            bad.append((k, e_gpu, e_emu))
    allm = torch.cat([named[k].grad.detach().cpu().double().flatten() for k in gkeys])
    allr = torch.cat([P[k].grad.double().flatten() for k in gkeys])
    alle = torch.cat([Pq[k].grad.double().flatten() for k in gkeys])
    whole_gpu, whole_emu = rl2(allm, allr), rl2(alle, allr)
    cos = float(F.cosine_similarity(allm, allr, 0))
    rows.sort(key=lambda r: -r[1])
    report(test="train_pass_at_bench_config", C=C, B=B, H=H, W=W, precision=O.get_precision(), routes=routes,
           logits_rel=logit_err, feats_rel=feat_err, loss_rel=loss_err, grad_whole_rl2_gpu=whole_gpu, grad_whole_rl2_tf32_emulation=whole_emu,
           grad_cos=cos, n_tensors=len(live), worst=[(k, round(a, 5), round(b, 5)) for k, a, b in rows[:8]],
           median_gpu=sorted(r[1] for r in rows)[len(rows) // 2], median_emu=sorted(r[2] for r in rows)[len(rows) // 2])
    assert logit_err[0] <= 1e-2, logit_err
    assert feat_err <= 1e-2, feat_err
    for k, e in loss_err.items():
        assert e <= 1e-3, (k, e, float(g_parts[k]), float(parts[k]))
    assert not bad, bad[:8]
    assert whole_gpu <= GRAD_SLACK * whole_emu + GRAD_FLOOR, (whole_gpu, whole_emu)
    # ---- BatchNorm running statistics (momentum 0.1, unbiased variance) of both encoders, the decoder and the loss head
    sd = seg.model.state_dict()
    for k in ("base.base_cnn.cnn.1", "base.base_cnn.path_estan.0.block5.2", "base.base_cnn.path_estan.4.block34.4",
              "base.base_vit.stem.1.bn", "base.base_vit.mhca_stages.3.aggregate.bn", "base.dec4.prep.1", "lap_map.1"):
        for s in (".running_mean", ".running_var"):
            assert rel(sd[k + s], P[k + s]) <= 2e-3, (k + s, rel(sd[k + s], P[k + s]))
    assert int(sd["lap_map.1.num_batches_tracked"]) == 2 and int(sd["base.base_cnn.cnn.1.num_batches_tracked"]) == 1


def test_all_heads_at_bench_config():
    """Logits of the four heads (y0, y1, y2, y4) in train mode at 8x256x256 against the oracle, default precision."""
    C, K, B, H, W, seed = 5, 4, 8, 256, 256, 29
    img, lab = make_bscans(B, H, W, C, K, seed)
    gen = torch.Generator().manual_seed(seed + 100)
    masks = dp_masks(B, gen)
    P = golden_state(C, seed)
    with torch.no_grad():
        outs, feats = orc.ftc_forward(P, img, orc.Ctx(True, [m.clone() for m in masks]))
    with contextlib.redirect_stdout(io.StringIO()):
        net = stc_tt(C)
    net.load_state_dict({k[5:]: v for k, v in golden_state(C, seed).items() if k.startswith("base.")}, strict=True)
    net = net.to(DEV).train()
    MHCABlock.dp_tape = [m.clone() for m in masks]
    try:
        with torch.no_grad():
            got = net(img.to(DEV))
    finally:
        MHCABlock.dp_tape = None
    errs = [rel(g, o) for g, o in zip(got, outs)]
    report(test="all_heads_at_bench_config", logits_rel=errs, feats_rel=rel(net.feats[0], feats))
    assert max(errs) <= 1e-2, errs
    assert rel(net.feats[0], feats) <= 1e-2


@pytest.mark.parametrize("C,K", [(5, 4)])
def test_two_train_steps_default_precision(tmp_path, C, K):
    """KiteSeg.train_step x2 in the shipped tf32 mode against OracleTrainer at a tcgen05-eligible size (2x128x256): loss terms
    1e-3, clipped-gradient norm 2e-2, weights after two AdamW steps (mean displacement error in units of lr)."""
    torch.set_num_threads(max(8, os.cpu_count() or 8))
    B, H, W, seed, lr = 2, 128, 256, 17, 1e-4
    seg = make_seg(C, seed, tmp_path, B, H, W)
    seg.optimG.param_groups[0]["lr"] = lr
    seg.optimG.sync_lr()
    P = golden_state(C, seed)
    tr = orc.OracleTrainer(P, lr=lr)
    gen = torch.Generator().manual_seed(seed + 100)
    before = L.route_counts()
    try:
        for step in range(2):
            img, lab = make_bscans(B, H, W, C, K, seed + step)
            onehot = F.one_hot(lab, C).permute(0, 3, 1, 2)
            noise = orc.make_noise(B, C, H, W, gen)
            masks = dp_masks(B, gen)
            total, parts, gnorm = tr.step(img, onehot, noise, [m.clone() for m in masks])
            MHCABlock.dp_tape = [m.clone() for m in masks]
            RegNet.noise_tape = noise
            got = seg.train_step(img, lab).cpu().tolist()
            want = [parts["los"], parts["udh"], parts["reg"], total]
            report(test="two_train_steps_default_precision", step=step, got=got, want=want, gnorm=[seg.optimG.last_grad_norm(), gnorm])
            for name, g, w in zip(("los", "udh", "reg", "total"), got, want):
                assert abs(g - w) <= 1e-3 * max(abs(w), 1e-3), (step, name, g, w)
            # step 0 starts from identical weights; by step 1 the two trajectories differ by +-lr on every element whose gradient sign is
            # TF32 round-off (Adam's first update is lr * sign(g)), which moves the next gradient's norm by a few percent
            assert abs(seg.optimG.last_grad_norm() - gnorm) <= (3e-2 if step == 0 else 8e-2) * gnorm, (seg.optimG.last_grad_norm(), gnorm)
    finally:
        MHCABlock.dp_tape = None
        RegNet.noise_tape = None
    after = L.route_counts()
    assert after["conv_tma"] > before["conv_tma"] and after["gemm_tma"] > before["gemm_tma"]
    sd = seg.model.state_dict()
    gmax = max(float(P[k].grad.abs().max()) for k in tr.keys if P[k].grad is not None)
    live = [k for k in tr.keys if P[k].grad is not None and float(P[k].grad.abs().max()) > 1e-3 * gmax]
    disp = sorted((float((sd[k].cpu() - P[k].detach()).abs().mean()) / lr, k) for k in live)
    report(test="two_train_steps_default_precision", worst_displacement=disp[-5:], median=disp[len(disp) // 2])
    # Adam's first steps move every element by ~lr*sign(g): the displacement error counts sign disagreements (2 lr each)
    assert disp[len(disp) // 2][0] <= 0.25 and disp[-1][0] <= 1.0, disp[-5:]


def test_graph_replay_equals_eager_with_boundary_regression(tmp_path):
    """The production step (udh + reg + CUDA graph): the early boundary-regression backward inside the captured step must
    give the same losses and weights as eager execution.  DropPath is disabled and the Gumbel noise of regular_reg is
    replaced by a fixed tape in both runs (a captured graph would otherwise replay different draws than eager)."""
    B, H, W, C, K = 2, 128, 128, 5, 4
    batches = [make_bscans(B, H, W, C, K, 200 + i) for i in range(6)]
    gen = torch.Generator().manual_seed(5)
    fixed = orc.make_noise(B, C, H, W, gen)
    finals = []
    orig = RegNet._noise

    def fixed_noise(self, B_, Cm, H_, W_, device):
        cache = self.__dict__.setdefault("_fixed_noise", None)
        if cache is None:
            eps = torch.stack([fixed[0], fixed[1]]).to(device=device, dtype=torch.float32).contiguous()
            jit = torch.stack([fixed[3].reshape(-1), fixed[2].reshape(-1)]).to(device=device, dtype=torch.float32).contiguous()
            cache = self.__dict__["_fixed_noise"] = (eps, jit)
        return cache
    RegNet._noise = fixed_noise
    try:
        for graph in (False, True):
            seg = make_seg(C, 5, tmp_path / str(graph), B, H, W, graph=graph)
            for m in seg.model.modules():
                if isinstance(m, MHCABlock):
                    m.drop_rate = 0.0
            losses = [seg.train_step(img, lab).cpu().tolist() for img, lab in batches]
            if graph:
                assert seg._graphs and all(g.graph is not None for g in seg._graphs.values())
            finals.append((losses, {k: v.clone() for k, v in seg.model.state_dict().items()}))
    finally:
        RegNet._noise = orig
    (l0, s0), (l1, s1) = finals
    for a, b in zip(l0, l1):
        for i in range(4):
            assert abs(a[i] - b[i]) <= 2e-3 * max(abs(a[i]), 1e-3), (a, b)
    for k in s0:
        if s0[k].dtype.is_floating_point:
            assert float((s0[k] - s1[k]).abs().max()) <= 5e-3 * float(s0[k].abs().max()) + 1e-5, k
        else:
            assert torch.equal(s0[k], s1[k]), k


def test_arena_survives_eval_between_graph_replays(tmp_path):
    """an eager eval forward between two train steps used to shrink what Arena.reset re-zeroes, so a later eager
    train step of a new shape found dirty 'zeroed' statistics slots.  Train (graph) -> eval -> train at a new shape must match
    the same sequence without the graph."""
    C, K = 5, 4
    res = []
    for graph in (False, True):
        seg = make_seg(C, 9, tmp_path / str(graph), 2, 64, 64, graph=graph)
        for m in seg.model.modules():
            if isinstance(m, MHCABlock):
                m.drop_rate = 0.0
        seg.args.reg = False
        for i in range(5):
            img, lab = make_bscans(2, 64, 64, C, K, 300 + i)
            seg.train_step(img, lab)
        seg.model.eval()
        seg.predict_labels(make_bscans(1, 32, 32, C, K, 1)[0])
        seg.model.train()
        img, lab = make_bscans(2, 64, 128, C, K, 400)         # new shape key: eager warm-up step
        res.append(seg.train_step(img, lab).cpu().tolist())
    for a, b in zip(*res):
        assert abs(a - b) <= 2e-3 * max(abs(a), 1e-3), res


@pytest.mark.parametrize("C,K,H,W", [(5, 4, 608, 512), (9, 9, 256, 512)])
def test_inference_at_full_frame_configs(C, K, H, W):
    """BASELINE.json configs 4 (K5): batched inference on the GOALS (608x512, C=5) and HCMS (256x512, C=9) full frames, eval mode,
    default precision, against the oracle run live on the same seeded frames (two of them: the oracle is a CPU program).  Head-0
    logits and `feats` within 1e-2 of max|ref|; the label map (`KiteSeg.predict`'s argmax) may differ only where the oracle's own
    top-1 / top-2 margin is below twice that tolerance; boundary positions within 0.05 px wherever both agree on the labels' column."""
    from tcct_b200.kite.loop_seg import argmax_labels
    torch.set_num_threads(max(8, os.cpu_count() or 8))
    B, seed = 2, 31
    img, _ = make_bscans(B, H, W, C, K, seed)
    P = golden_state(C, seed)
    with torch.no_grad():
        outs, feats = orc.ftc_forward(P, img, orc.Ctx(False))
    ref = outs[0]
    ref_lab = torch.argmax(torch.softmax(ref, 1), 1)
    with contextlib.redirect_stdout(io.StringIO()):
        net = stc_tt(C)
    net.load_state_dict({k[5:]: v for k, v in P.items() if k.startswith("base.")}, strict=True)
    net = net.to(DEV).eval()
    before = L.route_counts()
    with torch.no_grad():
        got = net(img.to(DEV))[0]
        lab = argmax_labels(got)
    torch.cuda.synchronize()
    after = L.route_counts()
    assert after["conv_tma"] > before["conv_tma"] and after["gemm_tma"] > before["gemm_tma"]
    amax = float(ref.abs().max())
    err = float((got.cpu() - ref).abs().max()) / amax
    ferr = rel(net.feats[0], feats)
    flipped = lab.cpu() != ref_lab
    top2 = ref.topk(2, 1).values
    margin = top2[:, 0] - top2[:, 1]
    worst = float(margin[flipped].max()) if bool(flipped.any()) else 0.0
    report(test="inference_at_full_frame_configs", C=C, H=H, W=W, logits_rel=err, feats_rel=ferr, flips=int(flipped.sum()),
           pixels=int(flipped.numel()), max_margin_of_flipped=worst, logit_absmax=amax)
    assert err <= 1e-2, err
    assert ferr <= 1e-2, ferr
    # random weights with untrained running statistics put many pixels on a decision boundary: the criterion is the margin one (a pixel
    # may flip only where the oracle's own top-1 / top-2 margin is below twice the logits tolerance) plus the 0.5 % cap of the other
    # inference tests
    assert int(flipped.sum()) <= 5e-3 * flipped.numel(), int(flipped.sum())
    assert worst <= 2e-2 * amax, (worst, amax)
    # boundary positions (soft-argmax extraction): compared on the oracle's own logits perturbed by nothing but the kernel's
    # arithmetic -- same input tensor on both sides
    from tcct_b200.nets import boundary_positions
    pos = boundary_positions(got, beta=100.0).cpu()
    want = orc.boundary_positions(got.cpu(), beta=100.0)
    assert float((pos - want.float()).abs().max()) <= 0.05
