"""Times the deep-supervision Dice paths at K2 / K3 sizes: fused (dice_multi) vs the per-head chain (3 resizes + 4 x Dice)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from tcct_b200 import ops as O
from time_kernels_util import timeit
dev = torch.device("cuda:0")
for (C, B, H, W) in ((5, 8, 256, 256), (9, 8, 256, 256), (9, 64, 256, 256)):
    zs = [[torch.randn(B, C, H // f, W // f, device=dev, requires_grad=True) for f in (1, 2, 4, 8)] for _ in range(3)]
    lab = torch.randint(0, C, (B, H, W), device=dev, dtype=torch.uint8)
    i = [0]
    def fused():
        i[0] += 1
        O.ARENA.reset(dev)
        t, _ = O.DiceMultiFn.apply(*zs[i[0] % 3], lab, 1.0)
        t.backward()
    def chain():
        i[0] += 1
        O.ARENA.reset(dev)
        z = zs[i[0] % 3]
        t = O.DiceFn.apply(z[0], lab, 0)
        for k in (1, 2, 3):
            t = t + O.DiceFn.apply(O.ResizeNCHWFn.apply(z[k], H, W), lab, 0)
        t.backward()
    def fused_fwd():
        i[0] += 1
        O.ARENA.reset(dev)
        with torch.no_grad():
            O.DiceMultiFn.apply(*zs[i[0] % 3], lab, 1.0)
    tff = timeit(fused_fwd, reps=6)
    tf, tc = timeit(fused, reps=6), timeit(chain, reps=6)
    px = B * H * W
    alg = (5.3125 * C + 1) + (2 * 5.3125 * C + 1)       # fwd + bwd bytes per pixel
    print("dice x4 C=%d B=%d %dx%d: fused fwd %.1f us, fwd+bwd %.1f us (%.0f GB/s on %.1f B/px) | per-head chain %.1f us" % (
        C, B, H, W, tff, tf, alg * px / tf / 1e3, alg, tc), flush=True)
