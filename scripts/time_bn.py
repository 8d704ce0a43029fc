"""bn_act2 backward (single operand, LeakyReLU pre-activation) at 8x256x256x32, 8x128x128x64 and 8x64x64x96, in-graph."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit
import tcct_b200._lib as L
from tcct_b200 import ops as O
from tcct_b200.ops import _p, _stream
dev = torch.device("cuda:0")
res = []
for (px, C) in ((8 * 256 * 256, 32), (8 * 128 * 128, 64), (8 * 128 * 128, 32), (8 * 64 * 64, 96), (8 * 32 * 32, 128), (8 * 16 * 16, 160)):
    xs = [torch.randn(px, C, device=dev) for _ in range(3)]
    dys = [torch.randn(px, C, device=dev) for _ in range(3)]
    gamma = torch.ones(C, device=dev)
    coef = torch.cat([torch.ones(C), torch.zeros(C), torch.zeros(C), torch.ones(C)]).to(dev)
    sums = torch.zeros(24 * C + 1, dtype=torch.float64, device=dev)
    da = torch.empty_like(xs[0]); dg = torch.zeros(C, device=dev); dbt = torch.zeros(C, device=dev)
    i = [0]
    def bb():
        i[0] += 1
        sums.zero_()
        L.bn_act2_bwd(_p(xs[i[0] % 3]), _p(coef), O.ACT_LRELU, _p(gamma), None, None, 0, None, O.ACT_NONE, _p(dys[i[0] % 3]), _p(sums),
                      _p(da), None, _p(dg), _p(dbt), None, None, px, C, _stream())
    t = timeit(bb)
    res.append("%dx%d: %.1f us (%.0f GB/s)" % (px, C, t, 12 * px * C / t / 1e3))
print("  ".join(res), flush=True)
