"""Real-weight known-answer fixture (SURVEY 4 / VERDICT r1 N1) -- TEST INFRASTRUCTURE ONLY.

Runs the UNMODIFIED reference (`nets.RegNet(nets.stc_tt(9))` from /root/reference/task1, recipe of
onnx/tcct_goals.py:1153-1164) with the shipped trained checkpoint `onnx/tcct_duke.pt` on the shipped B-scan
`onnx/oct_duke.png[:224,:512]` in eval mode, and stores

    tests/golden/tcct_duke.pt        the checkpoint itself, byte for byte (a data fixture the GPU box needs; md5 in the npz)
    tests/golden/real_duke.npz       image (uint8), head-0 logits of every 4th column (fp32), the full argmax label map,
                                     sum / max of the logits, the label histogram and the column-256 run lengths

and cross-checks oracle/tcct_oracle.py against the reference on the same input.

    python oracle/make_golden_real.py"""
import contextlib
import hashlib
import io
import os
import shutil
import sys

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()
import nets  # noqa: E402  (reference package)
import tcct_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
ONNX = os.path.join(refshim.REF_ROOT, "onnx")


def main():
    pt_path = os.path.join(ONNX, "tcct_duke.pt")
    md5 = hashlib.md5(open(pt_path, "rb").read()).hexdigest()
    state = torch.load(pt_path, map_location="cpu")
    n_class = state["base.aux0.weight"].shape[0]
    with contextlib.redirect_stdout(io.StringIO()):
        model = nets.RegNet(nets.stc_tt(n_class), out_channels=n_class)
    res = model.load_state_dict(state, strict=False)
    print("load_state_dict:", res)
    assert not res.missing_keys
    model.eval()
    im = cv2.imread(os.path.join(ONNX, "oct_duke.png"), cv2.IMREAD_COLOR)[:224, :512]
    assert (im[..., 0] == im[..., 1]).all() and (im[..., 1] == im[..., 2]).all()
    img = torch.from_numpy(im).permute(2, 0, 1)[None].float() / 255          # onnx_infer.py:19-24: HWC uint8 -> /255 -> NCHW
    with torch.no_grad():
        logits = model(img)[0]
    labels = torch.argmax(torch.softmax(logits, 1), 1)[0].numpy().astype(np.uint8)
    hist = np.bincount(labels.reshape(-1), minlength=n_class)
    col = labels[:, 256]
    runs = [(int(col[0]), 1)]
    for v in col[1:]:
        runs[-1] = (runs[-1][0], runs[-1][1] + 1) if v == runs[-1][0] else runs[-1]
        if v != runs[-1][0]:
            runs.append((int(v), 1))
    print("n_class", n_class, "hist", hist.tolist(), "sum %.1f max|logit| %.1f" % (float(logits.sum()), float(logits.abs().max())))
    print("column 256 runs", runs)
    # oracle cross-check
    P = {k: v.clone() for k, v in state.items()}
    o_logits, o_labels = O.predict_labels(P, img)
    err = float((o_logits - logits).abs().max() / logits.abs().max())
    flips = int((o_labels[0].numpy() != labels).sum())
    print("oracle vs reference: logits max-rel %.2e, argmax flips %d" % (err, flips))
    assert err < 1e-5 and flips == 0
    # top-1 / top-2 margin: pixels whose decision is closer than the logits tolerance may legitimately flip
    top2 = torch.topk(logits[0], 2, dim=0).values
    margin = (top2[0] - top2[1]).numpy()
    shutil.copyfile(pt_path, os.path.join(OUT, "tcct_duke.pt"))
    np.savez_compressed(os.path.join(OUT, "real_duke.npz"), image=im[..., 0].copy(), logits_sub=logits[0, :, :, ::4].numpy(),
                        labels=labels, logit_sum=np.float64(logits.double().sum()), logit_absmax=np.float32(logits.abs().max()),
                        hist=hist, runs256=np.array(runs, dtype=np.int32), margin=margin.astype(np.float16), n_class=n_class,
                        md5=np.array(md5))
    print("wrote real_duke.npz, tcct_duke.pt (md5 %s)" % md5)


if __name__ == "__main__":
    main()
