"""bn_act2 forward (single operand, LeakyReLU pre-activation, train mode): the launch the step uses (BatchNorm finalisation fused into
the prologue) against the same pass with precomputed coefficients and against a plain device copy of the same tensor, in-graph."""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit
import tcct_b200._lib as L
from tcct_b200 import ops as O
from tcct_b200.ops import _p, _stream
dev = torch.device("cuda:0")
for (B, H, W, C) in ((8, 256, 256, 32), (8, 128, 128, 64), (8, 128, 128, 32), (8, 64, 64, 96), (8, 32, 32, 128), (8, 16, 16, 160)):
    xs = [torch.randn(B, H, W, C, device=dev) for _ in range(3)]
    bn = torch.nn.BatchNorm2d(C).to(dev)
    out = torch.empty_like(xs[0])
    npix = B * H * W
    stats = torch.cat([torch.zeros(C), torch.full((C,), float(npix))]).double().to(dev)
    rec, coef = O._bn_src(bn, stats, npix, True, dev)
    i = [0]
    def fused():
        i[0] += 1
        L.bn_act2_fwd_bn(_p(xs[i[0] % 3]), ctypes.byref(rec), O.ACT_LRELU, None, None, 0, 0, _p(out), npix, C, _stream())
    def plain():
        i[0] += 1
        L.bn_act2_fwd(_p(xs[i[0] % 3]), _p(coef), O.ACT_LRELU, None, None, 0, 0, _p(out), npix, C, _stream())
    def cp():
        i[0] += 1
        out.copy_(xs[i[0] % 3])
    t, t2, tc = timeit(fused), timeit(plain), timeit(cp)
    n = npix * C * 4
    print("%dx%dx%dx%d (%.1f MB): fused-finalise %.1f us (%.0f GB/s)  precomputed coef %.1f us (%.0f GB/s)  copy %.1f us (%.0f GB/s)" % (
        B, H, W, C, n / 1e6, t, 2 * n / t / 1e3, t2, 2 * n / t2 / 1e3, tc, 2 * n / tc / 1e3), flush=True)
