"""Time the loss paths (boundary regression, feature polarisation, Dice x4) and a few single ops at the K2 shape with
CUDA events over CUDA-graph replays (GPU box).   python scripts/time_losses.py [B H W C]"""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit
from tcct_b200 import ops as O
from tcct_b200.nets import RegNet, stc_tt
from tcct_b200.synth import make_bscans

_a = [v for v in sys.argv[1:] if not v.startswith('--')]
B, H, W, C = (int(v) for v in _a[:4]) if len(_a) >= 4 else (8, 256, 256, 5)
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    net = RegNet(stc_tt(C), out_channels=C)
from tcct_b200.nets.flat import FlatParams
flat = FlatParams(net, dev)
net.train()
img, lab = make_bscans(B, H, W, C, C - 1 if C == 5 else C, 7)
lab8 = lab.to(torch.uint8).to(dev)
logits = torch.randn(B, C, H, W, device=dev, requires_grad=True)
feat = torch.randn(B, H, W, 32, device=dev, requires_grad=True)
px = B * H * W


def br():
    O.ARENA.reset(dev)
    eps = torch.rand(2, B, C - 1, H, W, device=dev).clamp_(1e-6, 1 - 1e-6)
    jit = torch.rand(2, H, device=dev)
    loss = O.BoundaryRegFn.apply(logits, lab8, eps, jit, net, True)
    loss.backward()


def fp():
    O.ARENA.reset(dev)
    loss = O.FeaturePolarFn.apply(feat, logits.detach(), lab8, net.fcp.buf_grad)
    loss.backward()


if "--eager" in sys.argv:          # target for an ncu launch list
    for _ in range(3):
        br(); fp()
    torch.cuda.synchronize()
    sys.exit(0)
for name, fn, bytes_px in (("boundary regression fwd+bwd", br, 12 * (C - 1) + 1 + 32 + 8 * (C - 1)),
                           ("feature polarisation fwd+bwd", fp, 128 * 2 + 4 * C + 1 + 32)):
    t = timeit(fn, reps=4)
    print("%-32s %8.1f us   %6.1f MB algorithmic -> %7.1f GB/s" % (name, t, bytes_px * px / 1e6, bytes_px * px / t / 1e3), flush=True)
