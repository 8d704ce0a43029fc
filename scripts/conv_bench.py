"""Times the tcgen05 conv / weight-gradient kernels at the full-resolution stage shapes (CUDA-graph replay, inputs rotating through > L2)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from tcct_b200 import ops as O
import tcct_b200._lib as L
from tcct_b200.ops import _p, _stream
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv
from time_kernels_util import timeit
dev = torch.device("cuda:0")
shapes = [(8, 256, 256), (8, 128, 128)] if len(sys.argv) < 2 else [tuple(int(v) for v in sys.argv[1].split("x"))]
for (B, H, W) in shapes:
    xs = [torch.randn(B, H, W, 32, device=dev) for _ in range(3)]
    dys = [torch.randn(B, H, W, 32, device=dev) for _ in range(3)]
    px = B * H * W
    for ks in (3, (1, 13), (13, 1), (1, 11), (11, 1)):
        mod = DenseConv(32, 32, ks).to(dev)
        plan = PackPlan(mod, dev)
        O.ARENA.reset(dev); plan.run()
        KH, KW = mod.weight.shape[2:]
        T = KH * KW
        i = [0]
        res = []
        for st in (False, True):
            def fwd():
                i[0] += 1
                with torch.no_grad():
                    mod.run(xs[i[0] % 3], want_stats=st, stats_act=O.ACT_LRELU)
            O.ARENA.reset(dev)
            res.append(timeit(fwd))
        tw = float("nan")
        if L.tcct_wgrad_tma_supported(H, W, 32, 32, KH, KW):
            dw = torch.zeros_like(mod.weight); db = torch.zeros(32, device=dev)
            ws = torch.empty(int(L.tcct_wgrad_tma_ws_floats(B, H, W, KH, KW)), device=dev)
            def wg():
                i[0] += 1
                L.wgrad_tma(_p(xs[i[0] % 3]), _p(dys[i[0] % 3]), _p(dw), _p(db), B, H, W, KH, KW, 32, _p(ws), None, _stream())
            tw = timeit(wg)
        print("conv %-8s @ %dx%dx%d: fwd %.1f us (%.0f GB/s, %.0f TF/s) | fwd+stats %.1f us | wgrad %.1f us (%.0f GB/s)" % (
            ks, B, H, W, res[0], 256 * px / res[0] / 1e3, 2 * 1024 * T * px / res[0] / 1e6, res[1], tw, 256 * px / tw / 1e3), flush=True)
