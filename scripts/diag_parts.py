"""Diagnostic (GPU box): gradient of each loss term (Dice x4 / feature polarisation / boundary regression) on the
kernel path vs the fp64 oracle, next to the fp32 oracle's own error.  python scripts/diag_parts.py [tf32|tf32x3]"""
import argparse, contextlib, io, os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import tcct_oracle as orc
from helpers import golden_state, dp_masks
from tcct_b200 import ops as O
from tcct_b200.kite.loop_seg import KiteSeg
from tcct_b200.nets import RegNet, stc_tt
from tcct_b200.nets.tcct import MHCABlock
from tcct_b200.synth import SynthOCT, make_bscans


def run(C, K, B, H, W, seed, precision):
    O.set_precision(precision)
    torch.set_num_threads(os.cpu_count() or 8)
    img, lab = make_bscans(B, H, W, C, K, seed)
    onehot = F.one_hot(lab, C).permute(0, 3, 1, 2)
    print("== case C%d B%d %dx%d seed %d precision %s" % (C, B, H, W, seed, precision))
    for name, udh, reg in (("dice", False, False), ("dice+udh", True, False), ("dice+reg", False, True)):
        gen = torch.Generator().manual_seed(seed + 100)
        noise = orc.make_noise(B, C, H, W, gen)
        masks = dp_masks(B, gen)
        res = {}
        for dt in ((torch.float32,) if reg else (torch.float64, torch.float32)):
            P = {k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in golden_state(C, seed).items()}
            tr = orc.OracleTrainer(P, lr=1e-4)
            nz = tuple(n.to(dt) for n in noise)
            total, parts, outs, feats = orc.calc_loss(P, img.to(dt), onehot, orc.Ctx(True, [m.clone() for m in masks]), nz, udh=udh, reg=reg)
            total.backward()
            res[dt] = {k: P[k].grad.double() for k in tr.keys if P[k].grad is not None}
        with contextlib.redirect_stdout(io.StringIO()):
            net = RegNet(stc_tt(C), out_channels=C)
            net.load_state_dict(golden_state(C, seed), strict=True)
            args = argparse.Namespace(los="di", lr=1e-2, gpu="0", pl=False, bs=B, bug=False, udh=udh, coff_udh=1.0, reg=reg,
                                      coff_reg=0.1, epl=False, coff_epl=0.1, coff_ds=1.0, graph=False)
            seg = KiteSeg(args, model=net, dataset=SynthOCT("goals" if C == 5 else "hcms", H, W, 1), root="/tmp/tcct_diag")
        seg.model.train()
        MHCABlock.dp_tape = [m.clone() for m in masks]
        RegNet.noise_tape = noise
        seg.optimG.zero_grad()
        total, parts = seg._losses(seg.cuda(img).float(), seg._label_map(lab))
        total.backward()
        MHCABlock.dp_tape = None
        RegNet.noise_tape = None
        named = dict(seg.model.named_parameters())
        g64, g32 = res.get(torch.float64, res[torch.float32]), res[torch.float32]
        keys = [k for k in g64 if k in named and named[k].grad is not None]
        fm = torch.cat([named[k].grad.detach().cpu().double().flatten() for k in keys])
        fr = torch.cat([g64[k].flatten() for k in keys])
        f32 = torch.cat([g32[k].flatten() for k in keys])
        print("  %-9s |g|: gpu %.6f  o64 %.6f  o32 %.6f | relL2 gpu %.3e  o32 %.3e" % (
            name, float(fm.norm()), float(fr.norm()), float(f32.norm()), float((fm - fr).norm() / fr.norm()), float((f32 - fr).norm() / fr.norm())))
        rows = []
        for k in keys:
            mine = named[k].grad.detach().cpu().double()
            d = float((mine - g64[k]).norm())
            rows.append((d, float((g32[k] - g64[k]).norm()), float(g64[k].norm()), k))
        gmax = max(r[2] for r in rows)
        if name == "dice" or reg:
            for d, d32, n, k in rows:
                if n > 1e-3 * gmax and (d / n > 1e-3 or reg):
                    print("      %-62s rel gpu %.3e o32 %.3e  |g| %.3e" % (k, d / n, d32 / n, n))


if __name__ == "__main__":
    prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
    run(5, 4, 2, 64, 64, 17, prec)
    run(9, 9, 2, 64, 64, 17, prec)
