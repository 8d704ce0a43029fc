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
