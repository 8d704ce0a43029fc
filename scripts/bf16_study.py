"""Precision study (CPU, build container only): what happens to the logits / loss / gradients of the oracle when chosen
activation tensors are STORED in bf16 (weights of those layers rounded to bf16, products exact, fp32 accumulation — the
arithmetic of tcgen05 kind::f16 with fp32 TMEM accumulators).  Decides which chains may move to bf16 under the
north-star tolerances (logits max-rel <= 1e-2, loss <= 1e-3).

    python scripts/bf16_study.py            # needs /root/reference (duke checkpoint + B-scan) for the real-weight part
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tcct_oracle as O  # noqa: E402

MODE = {"cnn": False, "vit": False, "dec": False}


def q(x):
    return x.to(torch.bfloat16).to(torch.float32)


class QF(torch.autograd.Function):
    """bf16 storage of an activation in forward AND of its gradient in backward."""

    @staticmethod
    def forward(ctx, x):
        return q(x)

    @staticmethod
    def backward(ctx, g):
        return q(g)


_conv0, _bn0, _cross0 = O._conv, O._bn, O.cross_block


def cross_block_q(P, key, x, k, ctx):
    if not MODE["cnn"]:
        return _cross0(P, key, x, k, ctx)
    W = lambda n: QF.apply(P[n + ".weight"])
    cv = lambda n, t, pad: QF.apply(F.conv2d(t, W(n), P[n + ".bias"], 1, pad))
    x = QF.apply(x)
    a = cv(key + ".block12.1", cv(key + ".block12.0", x, 1), 1)
    a = _bn0(P, key + ".block12.3", F.leaky_relu(a, 0.01), ctx)
    b = cv(key + ".block34.0", x, (0, k // 2))
    b = cv(key + ".block34.1", b, (k // 2, 0))
    b = cv(key + ".block34.2", b, 1)
    b = _bn0(P, key + ".block34.4", F.leaky_relu(b, 0.01), ctx)
    g = QF.apply(F.gelu(a + b))
    o = cv(key + ".block5.0", g, 1)
    return QF.apply(_bn0(P, key + ".block5.2", F.leaky_relu(o, 0.01), ctx))


def run_eval(P, img):
    with torch.no_grad():
        outs, _ = O.ftc_forward(P, img, O.Ctx(False))
    return outs


def report(name, outs, ref):
    rel = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(outs, ref)]
    flips = int((outs[0].argmax(1) != ref[0].argmax(1)).sum())
    print("%-28s logits max-rel %s  argmax flips %d / %d" % (name, ["%.2e" % r for r in rel], flips, ref[0][:, 0].numel()))


def main():
    O.cross_block = cross_block_q
    torch.manual_seed(0)
    ref_root = "/root/reference/task1/onnx"
    if os.path.exists(ref_root):
        import cv2
        P = torch.load(os.path.join(ref_root, "tcct_duke.pt"), map_location="cpu")
        P = {k: v.float() if v.is_floating_point() else v for k, v in P.items()}
        im = cv2.imread(os.path.join(ref_root, "oct_duke.png"), cv2.IMREAD_COLOR)[:224, :512]
        img = torch.from_numpy(im).permute(2, 0, 1)[None].float() / 255
        MODE["cnn"] = False
        ref = run_eval(P, img)
        print("duke real weights: max|logit| %.1f" % float(ref[0].abs().max()))
        MODE["cnn"] = True
        report("duke eval, cnn chains bf16", run_eval(P, img), ref)
        MODE["cnn"] = False
    # synthetic train step, goals-shaped
    from helpers import golden_state, train_inputs
    for (C, K, B, H, W) in ((5, 4, 2, 128, 128),):
        state = golden_state(C, 0)
        img, lab, onehot, noise, masks = train_inputs((C, K, B, H, W, 0))
        res = {}
        for mode in (False, True):
            MODE["cnn"] = mode
            P = {k: v.clone() for k, v in state.items()}
            tr = O.OracleTrainer(P, lr=1e-4)
            ctx = O.Ctx(True, [m.clone() for m in masks])
            total, parts, outs, feats = O.calc_loss(P, img, onehot, ctx, noise)
            total.backward()
            res[mode] = (float(total), {k: float(v) for k, v in parts.items()}, [o.detach() for o in outs],
                         {k: P[k].grad.clone() for k in tr.keys if P[k].grad is not None})
        MODE["cnn"] = False
        t0, p0, o0, g0 = res[False]
        t1, p1, o1, g1 = res[True]
        report("synth train C=%d %dx%dx%d" % (C, B, H, W), o1, o0)
        print("   loss fp32 %.6f bf16 %.6f rel %.2e ; parts %s" % (t0, t1, abs(t1 - t0) / abs(t0),
              {k: "%.2e" % (abs(p1[k] - p0[k]) / max(abs(p0[k]), 1e-12)) for k in p0}))
        rl2 = {k: float((g1[k] - g0[k]).norm() / (g0[k].norm() + 1e-30)) for k in g0}
        cos = {k: float((g1[k] * g0[k]).sum() / (g1[k].norm() * g0[k].norm() + 1e-30)) for k in g0}
        ks = sorted(rl2, key=rl2.get)
        print("   grad rel-L2: median %.2e, 90%% %.2e, worst %.2e (%s) ; worst cos %.4f" % (
            rl2[ks[len(ks) // 2]], rl2[ks[int(len(ks) * 0.9)]], rl2[ks[-1]], ks[-1], min(cos.values())))
        allg0 = torch.cat([g0[k].flatten() for k in g0]); allg1 = torch.cat([g1[k].flatten() for k in g0])
        print("   whole-gradient rel-L2 %.2e" % float((allg1 - allg0).norm() / allg0.norm()))
        for k in ks[-5:]:
            print("      %s rl2 %.2e |g| %.2e" % (k, rl2[k], float(g0[k].norm())))


if __name__ == "__main__":
    main()
