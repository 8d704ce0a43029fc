import sys, os, contextlib, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT + "/oracle")
import numpy as np, torch, torch.nn.functional as F
from helpers import dp_masks, factory_state, load
from tcct_b200 import ops as O
import tcct_b200.nets as N
import tcct_oracle as orc
from tcct_b200.nets.tcct import GateFusion, MHCABlock
from tcct_b200.synth import make_bscans
DEV = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
g = load("gtc_tt_goals_64")
n_class, n_bound, batch, height, width, seed = (int(v) for v in g["meta"])
for name, gate, sd in (("stc_tt", False, seed), ("gtc_tt", True, seed), ("gtc_tt", True, seed + 1), ("stc_tt", False, seed + 1)):
    img, lab = make_bscans(batch, height, width, n_class, n_bound, sd)
    onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2).to(DEV)
    state = factory_state("gtc_tt", n_class, sd)
    al = [torch.rand(batch, 32, 3, 3, generator=torch.Generator().manual_seed(5 + i)) for i in range(4)] if gate else None
    def loss_of(outs):
        return orc.multi_dice(outs[0], onehot) + sum(o.mean() for o in outs[1:])
    P = {"base." + k: v.clone().to(DEV).requires_grad_(v.dtype.is_floating_point) for k, v in state.items()}
    masks = [m.to(DEV) for m in dp_masks(batch, torch.Generator().manual_seed(sd + 100))]
    o_outs, _ = orc.ftc_forward(P, img.to(DEV), orc.Ctx(True, masks, gate_alphas=[a.to(DEV) for a in al] if al else None), gate=gate)
    loss_of(o_outs).backward()
    for prec in ("tf32x3",):
        with contextlib.redirect_stdout(io.StringIO()):
            net = getattr(N, name)(n_class)
        net.load_state_dict(state, strict=True)
        net = net.to(DEV).train()
        O.set_precision(prec)
        MHCABlock.dp_tape = dp_masks(batch, torch.Generator().manual_seed(sd + 100))
        GateFusion.alpha_tape = al
        outs = net(img.to(DEV))
        loss_of(outs).backward()
        O.set_precision("tf32")
        errs = []
        for k, prm in net.named_parameters():
            r = P["base." + k].grad
            if prm.grad is None or r is None or float(r.abs().max()) < 1e-6:
                continue
            errs.append((float((prm.grad - r).norm() / r.norm()), k))
        errs.sort(reverse=True)
        cnn = [e for e, k in errs if k.startswith("base_cnn")]
        rest = [e for e, k in errs if not k.startswith("base_cnn")]
        print(name, "seed", sd, prec, "CNN-branch grads: max %.2e median %.2e | others: max %.2e median %.2e | worst %s" % (
            max(cnn), sorted(cnn)[len(cnn) // 2], max(rest), sorted(rest)[len(rest) // 2], errs[0][1]))
