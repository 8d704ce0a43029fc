"""bn_act2 forward / backward for the activation configurations of the step (pre, post), at an MPViT stage-0 shape and a decoder shape."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit
import tcct_b200._lib as L
from tcct_b200 import ops as O
from tcct_b200.ops import _p, _stream
dev = torch.device("cuda:0")
N = {0: "none", 1: "lrelu", 2: "hswish", 3: "gelu"}
for (B, H, W, C) in ((8, 128, 128, 64), (8, 256, 256, 32), (8, 32, 32, 128)):
    npix = B * H * W
    xs = [torch.randn(npix, C, device=dev) for _ in range(3)]
    dys = [torch.randn(npix, C, device=dev) for _ in range(3)]
    bn = torch.nn.BatchNorm2d(C).to(dev)
    stats = torch.cat([torch.zeros(C), torch.full((C,), float(npix))]).double().to(dev)
    rec, coef = O._bn_src(bn, stats, npix, True, dev)
    out = torch.empty_like(xs[0]); da = torch.empty_like(xs[0])
    dg, dbt = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    sums = torch.zeros(24 * C + 1, dtype=torch.float64, device=dev)
    i = [0]
    for pre, post in ((1, 0), (0, 2), (0, 1), (0, 0), (0, 3)):
        def fwd():
            i[0] += 1
            L.bn_act2_fwd_bn(_p(xs[i[0] % 3]), ctypes.byref(rec), pre, None, None, 0, post, _p(out), npix, C, _stream())
        def bwd():
            i[0] += 1
            L.bn_act2_bwd(_p(xs[i[0] % 3]), _p(coef), pre, _p(bn.weight), None, None, 0, None, post, _p(dys[i[0] % 3]), _p(sums),
                          _p(da), None, _p(dg), _p(dbt), None, None, npix, C, _stream())
        print("%dx%dx%dx%d pre=%-6s post=%-6s  fwd %5.1f us   bwd pair %5.1f us" % (B, H, W, C, N[pre], N[post], timeit(fwd), timeit(bwd)), flush=True)
