"""Golden vectors of the validation scores (task1/kite/losses/miou.py:28-44,69-91) from the UNMODIFIED reference on seeded soft and
hard maps -- TEST INFRASTRUCTURE ONLY.      python oracle/make_golden_miou.py  ->  tests/golden/miou_scores.npz"""
import importlib.util, os, sys
import numpy as np
import torch
import torch.nn.functional as F
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
import refshim
import tcct_oracle as O
spec = importlib.util.spec_from_file_location("ref_miou", os.path.join(refshim.REF_ROOT, "kite/losses/miou.py"))
M = importlib.util.module_from_spec(spec); spec.loader.exec_module(M)
out = {}
for name, (B, C, H, W, seed, hard) in {"soft": (3, 7, 64, 48, 5, False), "hard": (2, 5, 40, 64, 6, True)}.items():
    g = torch.Generator().manual_seed(seed)
    pr = torch.softmax(torch.randn(B, C, H, W, generator=g) * 2, 1)
    gt = F.one_hot(torch.randint(0, C, (B, H, W), generator=g), C).permute(0, 3, 1, 2)
    if hard:
        pr = F.one_hot(pr.argmax(1), C).permute(0, 3, 1, 2).float()
    res = dict(dice_scorem0=float(M.MDiceLoss.scorem(pr, gt)), dice_scorem1=float(M.MDiceLoss.scorem(pr, gt, 1)),
               iou_scorem0=float(M.MIouLoss.scorem(pr, gt)), iou_scorem1=float(M.MIouLoss.scorem(pr, gt, 1)),
               dice_scores=np.array(M.MDiceLoss.scores(pr, gt)), dice_score=float(M.MDiceLoss.score(pr, gt)), iou_score=float(M.MIouLoss.score(pr, gt)))
    d, i = O.val_scores(pr, gt)
    assert abs(float(d.mean()) - res["dice_scorem0"]) < 1e-6 and abs(float(i[1:].mean()) - res["iou_scorem1"]) < 1e-6
    assert np.allclose(d.numpy(), res["dice_scores"], atol=1e-6)
    out.update({name + "::" + k: np.asarray(v) for k, v in res.items()})
    out[name + "::meta"] = np.array([B, C, H, W, seed, int(hard)])
    print(name, {k: (v if np.ndim(v) == 0 else np.round(v, 4).tolist()) for k, v in res.items()})
np.savez_compressed(os.path.join(os.path.dirname(HERE), "tests", "golden", "miou_scores.npz"), **out)
