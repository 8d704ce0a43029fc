"""Diagnostic (GPU box): per-tensor gradient error of the kernel path vs an fp64 oracle, next to the fp32 oracle's own error."""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import tcct_oracle as orc
from helpers import golden_state, dp_masks
from tcct_b200 import ops as O
from tcct_b200.nets import stc_tt
from tcct_b200.nets.tcct import MHCABlock
from tcct_b200.synth import make_bscans
import torch.nn.functional as F

def run(n_class, n_bound, B, H, W, seed, precision='tf32'):
    O.set_precision(precision)
    torch.set_num_threads(os.cpu_count() or 8)
    img, lab = make_bscans(B, H, W, n_class, n_bound, seed)
    onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2)
    gen = torch.Generator().manual_seed(seed + 100)
    masks = dp_masks(B, gen)
    res = {}
    for dt in (torch.float64, torch.float32):
        P = {k: (v.to(dt) if v.is_floating_point() else v.clone()) for k, v in golden_state(n_class, seed).items()}
        tr = orc.OracleTrainer(P, lr=1e-4)
        total, parts, outs, feats = orc.calc_loss(P, img.to(dt), onehot, orc.Ctx(True, [m.clone() for m in masks]), None, udh=False, reg=False)
        total.backward()
        res[dt] = ({k: P[k].grad.double() for k in tr.keys if P[k].grad is not None}, [o.detach().double() for o in outs], float(total))
    with contextlib.redirect_stdout(io.StringIO()):
        net = stc_tt(n_class)
    net.load_state_dict({k[5:]: v for k, v in golden_state(n_class, seed).items() if k.startswith("base.")})
    net = net.cuda().train()
    MHCABlock.dp_tape = [m.clone() for m in masks]
    got = net(img.cuda())
    MHCABlock.dp_tape = None
    lab8 = O.labels_u8(onehot.cuda().contiguous(), n_class)
    loss = sum(O.DiceFn.apply(got[i], lab8, 0) for i in range(3, 0, -1)) + O.DiceFn.apply(got[0], lab8, 0)
    loss.backward()
    g64, o64, l64 = res[torch.float64]
    g32, o32, l32 = res[torch.float32]
    print("precision", precision); print("case %dx%d B%d C%d: loss64 %.8f loss32 %.8f gpu %.8f" % (H, W, B, n_class, l64, l32, float(loss)))
    for i in range(4):
        sc = float(o64[i].abs().max())
        print("  out%d rel err: oracle32 %.2e gpu %.2e" % (i, float((o32[i] - o64[i]).abs().max()) / sc, float((got[i].detach().cpu().double() - o64[i]).abs().max()) / sc))
    named = dict(net.named_parameters())
    rows = []
    gmax = max(float(v.abs().max()) for v in g64.values())
    for k, ref in g64.items():
        if not k.startswith("base."):
            continue
        mine = named[k[5:]].grad.detach().cpu().double()
        sc = float(ref.abs().max())
        rows.append((float((mine - ref).abs().max()) / max(sc, 1e-30), float((g32[k] - ref).abs().max()) / max(sc, 1e-30), sc / gmax,
                     float((mine - ref).norm() / (ref.norm() + 1e-30)), k))
    rows.sort(reverse=True)
    flat_m = torch.cat([named[k[5:]].grad.detach().cpu().double().flatten() for k in g64 if k.startswith("base.")])
    flat_r = torch.cat([g64[k].flatten() for k in g64 if k.startswith("base.")])
    print("  global: cos %.8f rel L2 %.3e ; oracle32 rel L2 %.3e" % (float(F.cosine_similarity(flat_m, flat_r, 0)), float((flat_m - flat_r).norm() / flat_r.norm()),
          float((torch.cat([g32[k].flatten() for k in g64 if k.startswith('base.')]) - flat_r).norm() / flat_r.norm())))
    print("  %-70s %10s %10s %10s %10s" % ("tensor", "gpu maxrel", "o32 maxrel", "|g|/gmax", "gpu relL2"))
    for r in rows[:25]:
        print("  %-70s %10.2e %10.2e %10.2e %10.2e" % (r[4][5:], r[0], r[1], r[2], r[3]))
    import statistics
    print("  median gpu maxrel %.2e ; n>2e-2: %d of %d" % (statistics.median(r[0] for r in rows), sum(r[0] > 2e-2 for r in rows), len(rows)))

if __name__ == "__main__":
    run(5, 4, 2, 64, 64, 11, 'tf32x3')
    run(9, 9, 2, 64, 128, 12, 'tf32x3')
    run(5, 4, 2, 256, 256, 21, 'tf32x3')
