"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/task1) on seeded synthetic inputs with injected noise, and
cross-check oracle/tcct_oracle.py against it -- TEST INFRASTRUCTURE ONLY.

Run in the build container:  python oracle/make_golden.py
The vectors travel with the repo; the reference itself does not."""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()
import nets  # noqa: E402  (reference package)
from tcct_b200.synth import make_bscans, synth_state  # noqa: E402
import tcct_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
FULL_GRADS = ["base.aux0.weight", "base.base_cnn.path_estan.0.block34.0.weight", "base.base_cnn.cnn.0.weight",
              "base.base_vit.mhca_stages.1.mhca_blks.0.MHCA_layers.0.mlp.fc1.weight",
              "base.base_vit.mhca_stages.2.mhca_blks.0.cpe.proj.weight", "base.dec4.post.0.bias",
              "lap_reg.0.weight", "lap_reg.1.bias", "lap_map.0.weight", "lap_map.1.weight", "lap_map.2.bias"]
STAT_KEYS = ["base.base_cnn.cnn.1", "base.base_cnn.path_estan.0.block5.2", "base.base_vit.stem.1.bn",
             "base.base_vit.mhca_stages.3.aggregate.bn", "base.dec4.prep.1", "lap_map.1"]


def build_reference(n_class, seed):
    with contextlib.redirect_stdout(io.StringIO()):
        model = nets.RegNet(nets.stc_tt(n_class), out_channels=n_class)
    state = synth_state(model.state_dict(), seed)
    model.load_state_dict(state, strict=True)
    return model, state


class RandTape:
    """Replace torch.rand_like by a tape of pre-drawn tensors (reg.py:120,147,148)."""

    def __init__(self, tensors):
        self.tensors = list(tensors)

    def __enter__(self):
        self._orig = torch.rand_like
        torch.rand_like = lambda t, **kw: self.tensors.pop(0).to(t.dtype)
        return self

    def __exit__(self, *a):
        torch.rand_like = self._orig


def dp_masks(batch, gen):
    rates = [r for r in O.DROP_PATH if r > 0 for _ in range(2)]
    return [(torch.rand(batch, generator=gen) < 1 - r).float() for r in rates]


def train_case(name, n_class, n_bound, batch, height, width, seed):
    gen = torch.Generator().manual_seed(seed + 100)
    img, lab = make_bscans(batch, height, width, n_class, n_bound, seed)
    onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2)
    noise = O.make_noise(batch, n_class, height, width, gen)
    masks = dp_masks(batch, gen)
    # ---- reference: RegNet forward + grad_calc + regular_udh + regular_reg (loop_seg.py:146-171)
    model, state = build_reference(n_class, seed)
    model.train()
    sys.path.insert(0, os.path.join(refshim.REF_ROOT, "kite", "losses"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_loss", os.path.join(refshim.REF_ROOT, "kite/losses/loss.py"))
    ref_loss = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_loss)
    with contextlib.redirect_stdout(io.StringIO()):
        crit = ref_loss.get_loss("di")
    refshim.DropPath.tape = [m.clone() for m in masks]
    outs = model(img)
    los = sum(crit(outs[i], onehot) * 1.0 for i in range(3, 0, -1)) + crit(outs[0], onehot)
    udh = model.regular_udh(outs[0], onehot) * 1.0
    with RandTape(noise):
        reg = model.regular_reg(outs[0], onehot) * 0.1
    total = los + udh + reg
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=2e-4)
    opt.zero_grad(); total.backward()
    gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), 12)
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    feats = model.base.feats[0].detach()
    opt.step()
    after = model.state_dict()
    refshim.DropPath.tape = None
    # ---- oracle on the same inputs
    P = {k: v.clone() for k, v in state.items()}
    tr = O.OracleTrainer(P, lr=1e-4)
    ctx = O.Ctx(True, [m.clone() for m in masks])
    tr.opt.zero_grad()
    o_total, o_parts, o_outs, o_feats = O.calc_loss(P, img, onehot, ctx, noise)
    o_total.backward()
    o_gnorm = torch.nn.utils.clip_grad_norm_([P[k] for k in tr.keys], 12)
    rel = lambda a, b: float((a.detach() - b.detach()).abs().max() / (b.detach().abs().max() + 1e-30))
    print("[%s] logits rel err oracle vs reference:" % name, [rel(a, b) for a, b in zip(o_outs, outs)])
    print("   loss ref %.8f %.8f %.8f | oracle %.8f %.8f %.8f" % (
        float(los), float(udh), float(reg), float(o_parts["los"].detach()), float(o_parts["udh"].detach()), float(o_parts["reg"].detach())))
    print("   gnorm ref %.8f oracle %.8f ; feats rel %.2e" % (float(gnorm), float(o_gnorm), rel(o_feats, feats)))
    gmax = max(float(g.abs().max()) for g in grads.values())
    errs = {k: float((P[k].grad - grads[k]).abs().max()) / max(float(grads[k].abs().max()), 1e-3 * gmax)
            for k in grads if k in P and P[k].grad is not None}
    worst = max(errs.values())
    for k in sorted(errs, key=errs.get)[-3:]:
        print("      ", k, errs[k], float(grads[k].abs().max()))
    missing = [k for k in grads if k not in P or P[k].grad is None]
    extra = [k for k in tr.keys if P[k].grad is not None and k not in grads]
    print("   worst grad rel err %.2e ; missing %s extra %s ; n_grads %d" % (worst, missing, extra, len(grads)))
    assert worst < 1e-2 and not missing and not extra
    tr.opt.step()
    # Adam's first step is lr*sign(g): elements whose gradient is round-off noise flip freely,
    # so weights are compared as a mean over elements (in units of lr), not by max.
    w_after = max(float((P[k] - after[k]).abs().mean()) / 1e-4 for k in FULL_GRADS)
    s_after = max(float((P[k + s] - after[k + s]).abs().max()) for k in STAT_KEYS for s in (".running_mean", ".running_var"))
    print("   after-step mean |dW|/lr %.2e, running stats %.2e" % (w_after, s_after))
    assert w_after < 2e-2 and s_after < 1e-5
    rec = {"meta": np.array([n_class, n_bound, batch, height, width, seed], np.int64),
           "noise_seed": np.int64(seed + 100),
           "loss": np.array([float(los), float(udh), float(reg), float(total)], np.float64),
           "gnorm": np.float64(float(gnorm)),
           "out0": outs[0].detach().numpy(), "feats_sub": feats[:, :, ::4, ::4].numpy(),
           "grad_keys": np.array(sorted(grads)), "grad_norms": np.array([float(grads[k].norm()) for k in sorted(grads)]),
           "grad_sums": np.array([float(grads[k].double().sum()) for k in sorted(grads)])}
    for i in (1, 2, 3):
        rec["out%d_sub" % i] = outs[i].detach()[:, :, ::4, ::4].numpy()
    for k in FULL_GRADS:
        rec["grad::" + k] = grads[k].numpy()
        rec["after::" + k] = after[k].numpy()
    for k in STAT_KEYS:
        for s in (".running_mean", ".running_var", ".num_batches_tracked"):
            rec["after::" + k + s] = after[k + s].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)


def eval_case(name, n_class, n_bound, batch, height, width, seed):
    img, lab = make_bscans(batch, height, width, n_class, n_bound, seed)
    model, state = build_reference(n_class, seed)
    model.eval()
    with torch.no_grad():
        out0 = model(img)[0]
        labels = torch.argmax(F.softmax(out0, 1), 1)
        soft = nets.soft_argmax(out0)
    P = {k: v.clone() for k, v in state.items()}
    o_out0, o_lab = O.predict_labels(P, img)
    print("[%s] eval logits max|d| %.3e / max|ref| %.3e; label flips %d; soft_argmax d %.2e" % (
        name, float((o_out0 - out0).abs().max()), float(out0.abs().max()), int((o_lab != labels).sum()),
        float((O.soft_argmax(o_out0) - soft).abs().max())))
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        meta=np.array([n_class, n_bound, batch, height, width, seed], np.int64),
                        out0=out0.numpy(), labels=labels.numpy().astype(np.uint8), soft_argmax=soft.numpy())


def keys_fixture():
    """State-dict key/shape list of RegNet(stc_tt(C)) -- the checkpoint ABI."""
    lines = []
    for c in (5, 9):
        with contextlib.redirect_stdout(io.StringIO()):
            model = nets.RegNet(nets.stc_tt(c), out_channels=c)
        for k, v in model.state_dict().items():
            lines.append("%d %s %s %s" % (c, k, "x".join(map(str, v.shape)) or "-", str(v.dtype).replace("torch.", "")))
        req = {k for k, p in model.named_parameters() if p.requires_grad}
        lines += ["%d !trainable %s" % (c, k) for k in sorted(req)]
    with open(os.path.join(OUT, "state_keys.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    keys_fixture()
    train_case("train_goals_64", 5, 4, 2, 64, 64, 11)
    train_case("train_hcms_64x128", 9, 9, 2, 64, 128, 12)
    eval_case("eval_goals_96x64", 5, 4, 2, 96, 64, 13)
    eval_case("eval_hcms_64", 9, 9, 1, 64, 64, 14)
