"""Time the small-map kernels of CrossResNet stages 2-4 (conv fwd, wgrad, BN fwd/bwd) with CUDA events over CUDA-graph replays."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit
import tcct_b200._lib as L
from tcct_b200 import ops as O
from tcct_b200.ops import _p, _stream
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv

dev = torch.device("cuda:0")
for (B, H, W) in ((8, 128, 128), (8, 64, 64), (8, 32, 32), (8, 16, 16)):
    for ks in (3, (1, 9), (9, 1)):
        mod = DenseConv(32, 32, ks).to(dev)
        plan = PackPlan(mod, dev)
        O.ARENA.reset(dev); plan.run()
        x = torch.randn(B, H, W, 32, device=dev)
        dy = torch.randn(B, H, W, 32, device=dev)
        KH, KW = mod.weight.shape[2:]
        T = KH * KW

        def fwd():
            with torch.no_grad():
                mod.run(x, want_stats=True, stats_act=O.ACT_LRELU)
        O.ARENA.reset(dev)
        tf = timeit(fwd)
        dw = torch.zeros_like(mod.weight); db = torch.zeros(32, device=dev)

        def wg():
            L.wgrad(_p(x), _p(dy), _p(dw), _p(db), B, H, W, 32, 32, KH, KW, 32 * T, T, 1, 0, _stream())
        tw = timeit(wg)
        print("conv %-7s @ %dx%3dx%3d: fwd %6.1f us | wgrad(mma.sync) %6.1f us" % (ks, B, H, W, tf, tw), flush=True)
    # BN + act
    bn = torch.nn.BatchNorm2d(32).to(dev)
    a = torch.randn(B, H, W, 32, device=dev, requires_grad=True)
    st = torch.zeros(64, dtype=torch.float64, device=dev)
    coef = torch.cat([torch.ones(32), torch.zeros(32), torch.zeros(32), torch.ones(32)]).to(dev)
    sums = torch.zeros(8 * 96 + 1, dtype=torch.float64, device=dev)
    da = torch.empty_like(a); dg = torch.zeros(32, device=dev); dbt = torch.zeros(32, device=dev)
    out = torch.empty_like(a)
    px = B * H * W

    def bf():
        L.bn_act2_fwd(_p(a), _p(coef), O.ACT_LRELU, None, None, 0, O.ACT_NONE, _p(out), px, 32, _stream())

    def bb():
        sums.zero_()
        L.bn_act2_bwd(_p(a), _p(coef), O.ACT_LRELU, _p(bn.weight), None, None, 0, None, O.ACT_NONE, _p(dy), _p(sums),
                      _p(da), None, _p(dg), _p(dbt), None, None, px, 32, _stream())
    print("bn_act2 @ %dx%3dx%3d: fwd %6.1f us | bwd fused (+memset) %6.1f us" % (B, H, W, timeit(bf), timeit(bb)), flush=True)
