"""Golden vectors from the UNMODIFIED reference for the gated / wide factories -- TEST INFRASTRUCTURE ONLY.

  * `gtc_tt` (task1/nets/tcct.py:1050-1055: GateFusion, tiny CrossResNet), `stc_tb` (1097-1102: SimpleFusion, the 32-64-96-128-256
    CrossResNet), `gtc_tb` (1056-1061: both): eval logits / labels and the loss + a few gradients of one train-mode forward/backward
    on seeded synthetic weights; the random gate fields the reference draws with torch.rand inside GateFusion.forward are RECORDED
    (torch.rand is wrapped for the duration of the call, the reference code is untouched) and stored, so that the oracle and the CUDA
    path can be fed the same fields                          -> tests/golden/{gtc_tt,stc_tb,gtc_tb}_goals_64.npz
  * the state-dict keys / shapes of the wide model           -> tests/golden/state_keys_tb.txt
and cross-checks oracle/tcct_oracle.py against each.      python oracle/make_golden_wide.py"""
import contextlib, io, os, sys
import numpy as np
import torch
import torch.nn.functional as F
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import refshim
refshim.install()
import nets
from tcct_b200.synth import make_bscans, synth_state
import tcct_oracle as O

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
torch.set_num_threads(8)


class RecordRand:
    """Records what torch.rand returns while the reference runs (GateFusion.forward is its only caller on this path: DropPath is taped
    by refshim)."""

    def __enter__(self):
        self.orig, self.fields = torch.rand, []

        def rand(*a, **k):
            t = self.orig(*a, **k)
            self.fields.append(t.clone())
            return t
        torch.rand = rand
        return self

    def __exit__(self, *exc):
        torch.rand = self.orig
        return False


def factory_case(name, gate, keys, seed):
    n_class, n_bound, batch, height, width = 5, 4, 2, 64, 64
    img, lab = make_bscans(batch, height, width, n_class, n_bound, seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model = getattr(nets, name)(n_class)
    state = synth_state(model.state_dict(), seed)
    model.load_state_dict(state, strict=True)
    model.eval()
    with torch.no_grad():
        out0 = model(img)[0]
    labels = torch.argmax(F.softmax(out0, 1), 1)
    P = {"base." + k: v.clone() for k, v in state.items()}
    o_out0, o_lab = O.predict_labels(P, img, gate=gate)
    print("[%s] eval: oracle vs reference max|d| %.3e (max|ref| %.3e), label flips %d" % (
        name, float((o_out0 - out0).abs().max()), float(out0.abs().max()), int((o_lab != labels).sum())))
    assert float((o_out0 - out0).abs().max()) <= 1e-4 * float(out0.abs().max())
    model.train()
    gen = torch.Generator().manual_seed(seed + 100)
    rates = [r for r in O.DROP_PATH if r > 0 for _ in range(2)]
    masks = [(torch.rand(batch, generator=gen) < 1 - r).float() for r in rates]      # tests/helpers.py:dp_masks
    refshim.DropPath.tape = [m.clone() for m in masks]
    torch.manual_seed(seed + 7)
    with RecordRand() as rec:
        outs = model(img)
    refshim.DropPath.tape = None
    alphas = rec.fields
    assert len(alphas) == (4 if gate else 0), len(alphas)
    onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2)
    loss = O.multi_dice(outs[0], onehot) + sum(o.mean() for o in outs[1:])
    loss.backward()
    named = dict(model.named_parameters())
    assert all(named[k].grad is not None for k in keys), [k for k in keys if named[k].grad is None]
    Pt = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in P.items()}
    o_outs, _ = O.ftc_forward(Pt, img, O.Ctx(True, [m.clone() for m in masks], gate_alphas=[a.clone() for a in alphas]), gate=gate)
    print("   train out0: oracle vs reference rel %.2e" % (float((o_outs[0] - outs[0]).abs().max()) / float(outs[0].abs().max())))
    o_loss = O.multi_dice(o_outs[0], onehot) + sum(o.mean() for o in o_outs[1:])
    o_loss.backward()
    for k in keys:
        d = float((Pt["base." + k].grad - named[k].grad).abs().max()) / float(named[k].grad.abs().max())
        print("   train grad %-45s oracle vs reference rel %.2e" % (k, d))
        assert d < 1e-3
    sd = model.state_dict()
    np.savez_compressed(os.path.join(OUT, "%s_goals_64.npz" % name), meta=np.array([n_class, n_bound, batch, height, width, seed], np.int64),
                        out0=out0.numpy(), labels=labels.numpy().astype(np.uint8), train_out0=outs[0].detach().numpy(),
                        train_loss=np.float64(float(loss)), **{"grad::" + k: named[k].grad.numpy() for k in keys},
                        **{"alpha%d" % i: a.numpy() for i, a in enumerate(alphas)},
                        cnn_running_mean=sd["base_cnn.path_estan.2.block5.2.running_mean"].numpy(),
                        vit_running_mean=sd["base_vit.stem.1.bn.running_mean"].numpy())
    return model


def keys_fixture(model, n_class):
    lines = ["%d %s %s %s" % (n_class, k, "x".join(map(str, v.shape)) or "-", str(v.dtype).replace("torch.", "")) for k, v in model.state_dict().items()]
    with open(os.path.join(OUT, "state_keys_tb.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    factory_case("gtc_tt", True, ["base_cnn.path_estan.0.block34.0.weight", "base_vit.mhca_stages.1.aggregate.conv.weight", "tran_vit2.0.weight",
                                  "tran_cnn0.1.weight", "dec2.prep.0.weight", "t323.bias", "aux0.weight"], 51)
    wide_keys = ["base_cnn.path_estan.1.block12.0.weight", "base_cnn.path_estan.2.block34.1.weight", "base_cnn.path_estan.4.block5.0.weight",
                 "base_cnn.path_estan.3.block34.0.bias", "tran_vit3.0.weight", "tran_cnn1.0.weight", "head.0.weight", "dec1.prep.0.weight",
                 "dec4.prep.0.weight", "t321.weight", "aux4.bias"]
    m = factory_case("stc_tb", False, wide_keys, 53)
    keys_fixture(m, 5)
    factory_case("gtc_tb", True, wide_keys, 55)
