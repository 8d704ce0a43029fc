"""Depthwise 3x3 kernels one by one (forward, data gradient, weight gradient) at the MPViT shapes, in-graph, against their bytes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit
import tcct_b200._lib as L
from tcct_b200.ops import _p, _stream
dev = torch.device("cuda:0")
SHAPES = ((8, 128, 128, 64, 1), (8, 64, 64, 96, 1), (8, 32, 32, 128, 1), (8, 16, 16, 160, 1), (32, 128, 128, 64, 1))
if not os.environ.get("TCCT_DW_ROWS"):
    SHAPES += ((8, 128, 128, 64, 2), (8, 64, 64, 96, 2), (8, 256, 256, 32, 2))
for (B, H, W, C, s) in SHAPES:
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    xs = [torch.randn(B, H, W, C, device=dev) for _ in range(3)]
    dys = [torch.randn(B, Ho, Wo, C, device=dev) for _ in range(3)]
    w = torch.randn(C, 1, 3, 3, device=dev); b = torch.randn(C, device=dev)
    y = torch.empty(B, Ho, Wo, C, device=dev); dx = torch.empty_like(xs[0])
    dw = torch.zeros_like(w); db = torch.zeros_like(b)
    i = [0]
    def fwd():
        i[0] += 1
        L.dwconv3_fwd(_p(xs[i[0] % 3]), _p(w), _p(b), _p(y), B, H, W, C, s, 0, None, _stream())
    def dgrad():
        i[0] += 1
        L.dwconv3_bwd(_p(xs[i[0] % 3]), _p(w), _p(dys[i[0] % 3]), _p(dx), None, None, B, H, W, C, s, 0, _stream())
    def wgrad():
        i[0] += 1
        L.dwconv3_bwd(_p(xs[i[0] % 3]), _p(w), _p(dys[i[0] % 3]), None, _p(dw), _p(db), B, H, W, C, s, 0, _stream())
    nx, ny = xs[0].numel() * 4 / 1e6, y.numel() * 4 / 1e6
    tf, td, tw = timeit(fwd), timeit(dgrad), timeit(wgrad)
    print("%dx%dx%dx%d stride %d: fwd %5.1f us (%4.0f GB/s)  dgrad %5.1f us (%4.0f GB/s)  wgrad %5.1f us (%4.0f GB/s)" % (
        B, H, W, C, s, tf, (nx + ny) / tf * 1e3, td, (nx + ny) / td * 1e3, tw, (nx + ny) / tw * 1e3), flush=True)
