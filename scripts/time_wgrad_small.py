"""wgrad (mma.sync) on the small maps under the TCCT_WGRAD_ZS / TCCT_WGRAD_GX knobs (set in the environment)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit
import tcct_b200._lib as L
from tcct_b200.ops import _p, _stream
dev = torch.device("cuda:0")
out = []
for (B, H, W) in ((8, 64, 64), (8, 32, 32), (8, 16, 16)):
    for (KH, KW) in ((3, 3), (1, 9)):
        T = KH * KW
        x = torch.randn(B, H, W, 32, device=dev); dy = torch.randn(B, H, W, 32, device=dev)
        dw = torch.zeros(32, 32, KH, KW, device=dev); db = torch.zeros(32, device=dev)
        def wg():
            L.wgrad(_p(x), _p(dy), _p(dw), _p(db), B, H, W, 32, 32, KH, KW, 32 * T, T, 1, 0, _stream())
        out.append("%dx%d@%d:%.1f" % (KH, KW, H, timeit(wg)))
print("ZS=%s GX=%s  " % (os.environ.get("TCCT_WGRAD_ZS", "-"), os.environ.get("TCCT_WGRAD_GX", "-")) + "  ".join(out), flush=True)
