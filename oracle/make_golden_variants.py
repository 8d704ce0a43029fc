"""Golden vectors from the UNMODIFIED reference for the remaining SimpleFusion factories and the deploy-time model variant
-- TEST INFRASTRUCTURE ONLY.

  * `pnnu`, `vitu` (task1/nets/tcct.py:1117-1122, 1131-1136): eval logits / labels and the loss + a few gradients of one train-mode
    forward/backward on seeded synthetic weights           -> tests/golden/{pnnu,vitu}_goals_64.npz
  * the onnx/ model definition (task1/onnx/tcct_goals.py, tcct_hcms.py: older decoder tail) with the shipped TRAINED checkpoints
    tcct_goals.pt / tcct_hcms.pt (recipe of their __main__, 1153-1164) on the reference's own B-scan onnx/oct_duke.png[:224,:512] in eval mode
                                                            -> tests/golden/real_{goals,hcms}.npz + the two checkpoints (data fixtures)
and cross-checks oracle/tcct_oracle.py against each.      python oracle/make_golden_variants.py"""
import contextlib, hashlib, importlib.util, io, os, shutil, sys
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


def factory_case(name, kw, keys, frozen, seed):
    n_class, n_bound, batch, height, width = 5, 4, 2, 64, 64
    img, lab = make_bscans(batch, height, width, n_class, n_bound, seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_full = nets.RegNet(nets.stc_tt(n_class), out_channels=n_class)      # the golden state builder of the tests
        model = getattr(nets, name)(n_class)
    full = {k[5:]: v for k, v in synth_state(ref_full.state_dict(), seed).items() if k.startswith("base.")}
    own = model.state_dict()
    # pnnu's cross convs are 1x3 / 3x1: cut the centre taps out of the stc_tt-shaped synthetic weights
    state = {}
    for k, v in own.items():
        w = full[k]
        if w.shape != v.shape:
            kh, kw_ = v.shape[2:]
            h0, w0 = (w.shape[2] - kh) // 2, (w.shape[3] - kw_) // 2
            w = w[:, :, h0:h0 + kh, w0:w0 + kw_].contiguous()
        state[k] = w
    model.load_state_dict(state, strict=True)
    model.eval()
    with torch.no_grad():
        out0 = model(img)[0]
    labels = torch.argmax(F.softmax(out0, 1), 1)
    P = {"base." + k: v.clone() for k, v in state.items()}
    o_out0, o_lab = O.predict_labels(P, img, **kw)
    print("[%s] eval: oracle vs reference max|d| %.3e (max|ref| %.3e), label flips %d" % (
        name, float((o_out0 - out0).abs().max()), float(out0.abs().max()), int((o_lab != labels).sum())))
    assert float((o_out0 - out0).abs().max()) <= 1e-4 * float(out0.abs().max())
    model.train()
    gen = torch.Generator().manual_seed(seed + 100)
    rates = [r for r in O.DROP_PATH if r > 0 for _ in range(2)]
    masks = [(torch.rand(batch, generator=gen) < 1 - r).float() for r in rates]      # tests/helpers.py:dp_masks
    refshim.DropPath.tape = [m.clone() for m in masks]
    outs = model(img)
    refshim.DropPath.tape = None
    onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2)
    loss = O.multi_dice(outs[0], onehot) + sum(o.mean() for o in outs[1:])
    loss.backward()
    named = dict(model.named_parameters())
    assert all(named[k].grad is not None for k in keys), [k for k in keys if named[k].grad is None]
    assert all(named[k].grad is None for k in frozen)
    Pt = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in P.items()}
    o_outs, _ = O.ftc_forward(Pt, img, O.Ctx(True, [m.clone() for m in masks]), **kw)
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
                        cnn_running_mean=sd["base_cnn.path_estan.2.block5.2.running_mean"].numpy(),
                        vit_running_mean=sd["base_vit.stem.1.bn.running_mean"].numpy())


def real_case(tag, n_class, n_bound, seed):
    path = os.path.join(refshim.REF_ROOT, "onnx", "tcct_%s.py" % tag)
    spec = importlib.util.spec_from_file_location("ref_onnx_" + tag, path)
    mod = importlib.util.module_from_spec(spec)
    for stub in ("pandas", "cv2", "PIL", "PIL.Image"):
        pass
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
        model = mod.RegNet(mod.stc_tt(n_class), out_channels=n_class)
    pt_path = os.path.join(refshim.REF_ROOT, "onnx", "tcct_%s.pt" % tag)
    md5 = hashlib.md5(open(pt_path, "rb").read()).hexdigest()
    state = torch.load(pt_path, map_location="cpu")
    res = model.load_state_dict(state, strict=False)
    print("[%s] load_state_dict: missing %s unexpected %s" % (tag, res.missing_keys, res.unexpected_keys))
    assert not res.missing_keys
    model.eval()
    import cv2
    im = cv2.imread(os.path.join(refshim.REF_ROOT, "onnx", "oct_duke.png"), cv2.IMREAD_COLOR)[:224, :512]      # the reference's own B-scan
    img = torch.from_numpy(im).permute(2, 0, 1)[None].float() / 255                 # = tests/golden/real_duke.npz["image"] / 255, 3 channels
    with torch.no_grad():
        logits = model(img)[0]
    labels = torch.argmax(torch.softmax(logits, 1), 1).numpy().astype(np.uint8)
    P = {k: v.clone() for k, v in state.items()}
    o_logits, o_labels = O.predict_labels(P, img, variant="onnx")
    err = float((o_logits - logits).abs().max() / logits.abs().max())
    print("[%s] oracle vs reference: logits max-rel %.2e, flips %d; max|logit| %.1f, hist %s" % (
        tag, err, int((o_labels.numpy() != labels).sum()), float(logits.abs().max()), np.bincount(labels.reshape(-1), minlength=n_class).tolist()))
    assert err < 1e-5
    top2 = torch.topk(logits, 2, dim=1).values
    shutil.copyfile(pt_path, os.path.join(OUT, "tcct_%s.pt" % tag))
    np.savez_compressed(os.path.join(OUT, "real_%s.npz" % tag), meta=np.array([n_class, n_bound, 1, 224, 512, seed], np.int64),
                        logits_sub=logits[:, :, :, ::4].numpy(), labels=labels, logit_absmax=np.float32(logits.abs().max()),
                        margin=(top2[:, 0] - top2[:, 1]).numpy().astype(np.float16), md5=np.array(md5))


if __name__ == "__main__":
    factory_case("pnnu", dict(flag_vit=False, plain=True),
                 ["base_cnn.path_estan.0.block34.0.weight", "base_cnn.path_estan.3.block34.1.weight", "dec2.prep.0.weight", "t323.bias", "aux0.weight"],
                 ["tran_cnn0.0.weight", "base_vit.stem.0.conv.weight"], 33)
    factory_case("vitu", dict(flag_vit=True, flag_cnn=False),
                 ["base_vit.stem.0.conv.weight", "base_vit.mhca_stages.1.aggregate.conv.weight", "tran_vit2.0.weight", "dec2.prep.0.weight", "t324.weight", "aux4.bias"],
                 ["tran_cnn0.0.weight", "base_cnn.cnn.0.weight", "base_cnn.path_estan.0.block12.0.weight"], 35)
    real_case("goals", 5, 4, 41)
    real_case("hcms", 9, 9, 43)
