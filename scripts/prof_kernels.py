"""Run the hot kernels at the stage-0 shape of workload K2 (8x256x256x32) a few times - target for `ncu`."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tcct_b200 import ops as O
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv

dev = torch.device("cuda:0")
B, H, W = 8, 256, 256
which = sys.argv[1] if len(sys.argv) > 1 else "conv3"
if which == "bn":          # BatchNorm(train) + LeakyReLU backward, single launch (reduce + grid barrier + apply)
    import tcct_b200._lib as L
    from tcct_b200.ops import _p, _stream
    bn = torch.nn.BatchNorm2d(32).to(dev)
    xs = [torch.randn(B, H, W, 32, device=dev) for _ in range(3)]
    dys = [torch.randn(B, H, W, 32, device=dev) for _ in range(3)]
    coef = torch.cat([torch.ones(32), torch.zeros(32), torch.zeros(32), torch.ones(32)]).to(dev)
    sums = torch.zeros(8 * 96 + 1, dtype=torch.float64, device=dev)
    da = torch.empty_like(xs[0]); dg = torch.zeros(32, device=dev); dbt = torch.zeros(32, device=dev)
    for i in range(4):
        sums.zero_()
        L.bn_act2_bwd(_p(xs[i % 3]), _p(coef), O.ACT_LRELU, _p(bn.weight), None, None, 0, None, O.ACT_NONE, _p(dys[i % 3]), _p(sums),
                      _p(da), None, _p(dg), _p(dbt), None, None, B * H * W, 32, _stream())
    torch.cuda.synchronize()
    print("done bn")
    sys.exit(0)
ks = {"conv3": 3, "conv13": (1, 13), "conv13v": (13, 1), "gemm": 1}[which]
mod = DenseConv(32, 32, ks).to(dev)
plan = PackPlan(mod, dev)
O.ARENA.reset(dev); plan.run()
x = torch.randn(B, H, W, 32, device=dev, requires_grad=True)
dy = torch.randn(B, H, W, 32, device=dev)
for i in range(3):
    O.ARENA.reset(dev)
    y, st = mod.run(x, want_stats=True, stats_act=O.ACT_LRELU)
    y.backward(dy)
torch.cuda.synchronize()
print("done", which)
