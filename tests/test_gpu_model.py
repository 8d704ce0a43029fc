"""GPU: the stc_tt network on the kernel path against (a) the golden vectors written by the unmodified
reference (tests/golden/*.npz) and (b) the CPU oracle run live on the same seeded inputs.
Tolerances are north_star's: logits max|d| <= 1e-2 * max|logit|, loss 1e-3 relative; gradients are held to
2e-2 of each tensor's max (TF32 contractions, fp32 everywhere else)."""
import contextlib
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

import tcct_oracle as orc  # noqa: E402
from helpers import golden_state, load, train_inputs  # noqa: E402
from tcct_b200 import ops as O  # noqa: E402
from tcct_b200.nets import stc_tt  # noqa: E402
from tcct_b200.nets.tcct import MHCABlock  # noqa: E402
from tcct_b200.synth import make_bscans  # noqa: E402

DEV = torch.device("cuda:0")


def build(n_class, seed):
    with contextlib.redirect_stdout(io.StringIO()):
        net = stc_tt(n_class)
    state = golden_state(n_class, seed)
    net.load_state_dict({k[5:]: v for k, v in state.items() if k.startswith("base.")}, strict=True)
    return net.to(DEV), state


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max()) / (float(b.abs().max()) + 1e-30)


@pytest.mark.parametrize("name", ["eval_goals_96x64", "eval_hcms_64"])
def test_eval_logits_and_labels_match_reference(name):
    g = load(name)
    n_class, n_bound, batch, height, width, seed = (int(v) for v in g["meta"])
    img, _ = make_bscans(batch, height, width, n_class, n_bound, seed)
    net, _ = build(n_class, seed)
    net.eval()
    with torch.no_grad():
        out = net(img.to(DEV))
    ref = torch.from_numpy(g["out0"])
    assert rel(out[0], ref) <= 1e-2, rel(out[0], ref)
    # argmax label maps derived from identical logits are bit-exact
    from tcct_b200.kite.loop_seg import argmax_labels
    lab_ref = torch.argmax(torch.softmax(out[0].cpu(), 1), 1)
    assert torch.equal(argmax_labels(out[0]).cpu().long(), lab_ref)
    flips = int((lab_ref.numpy().astype(np.uint8) != g["labels"]).sum())
    assert flips <= 0.002 * lab_ref.numel(), flips


@pytest.mark.parametrize("name", ["train_goals_64", "train_hcms_64x128"])
def test_train_forward_backward_matches_oracle(name):
    torch.set_num_threads(8)
    g = load(name)
    n_class, seed = int(g["meta"][0]), int(g["meta"][5])
    img, lab, onehot, noise, masks = train_inputs(g["meta"])
    # ---- oracle (CPU, fp32): network + deep-supervised Dice only
    P = golden_state(n_class, seed)
    tr = orc.OracleTrainer(P, lr=1e-4)
    tr.opt.zero_grad()
    total, parts, outs, feats = orc.calc_loss(P, img, onehot, orc.Ctx(True, [m.clone() for m in masks]), noise,
                                              udh=False, reg=False)
    total.backward()
    # ---- kernels
    net, state = build(n_class, seed)
    net.train()
    MHCABlock.dp_tape = [m.clone() for m in masks]
    try:
        got = net(img.to(DEV))
    finally:
        MHCABlock.dp_tape = None
    ref0 = torch.from_numpy(g["out0"])
    assert rel(got[0], ref0) <= 1e-2, ("logits vs reference golden", rel(got[0], ref0))
    for i in range(4):
        assert rel(got[i], outs[i]) <= 1e-2, (i, rel(got[i], outs[i]))
    assert rel(net.feats[0], feats) <= 1e-2
    lab8 = O.labels_u8(onehot.to(DEV).contiguous(), n_class)
    loss = sum(O.DiceFn.apply(got[i], lab8, 0) for i in range(3, 0, -1)) + O.DiceFn.apply(got[0], lab8, 0)
    assert abs(float(loss) - float(total)) <= 1e-3 * abs(float(total)), (float(loss), float(total))
    loss.backward()
    worst = {}
    gmax = max(float(P[k].grad.abs().max()) for k in tr.keys if P[k].grad is not None)
    named = dict(net.named_parameters())
    for k in tr.keys:
        if P[k].grad is None or not k.startswith("base."):
            continue
        p = named[k[5:]]
        assert p.grad is not None, k
        err = float((p.grad.cpu() - P[k].grad).abs().max()) / max(float(P[k].grad.abs().max()), 1e-3 * gmax)
        worst[k] = err
    bad = {k: v for k, v in worst.items() if v > 2e-2}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:8]
    # running statistics after one training forward (momentum 0.1, unbiased variance)
    sd = net.state_dict()
    for k in ("base_cnn.cnn.1", "base_cnn.path_estan.0.block5.2", "base_vit.stem.1.bn", "dec4.prep.1"):
        for s in (".running_mean", ".running_var"):
            assert rel(sd[k + s], P["base." + k + s]) <= 2e-3, k + s
        assert int(sd[k + ".num_batches_tracked"]) == 1
