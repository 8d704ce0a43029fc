"""Diagnostic (GPU box): accuracy of the error-compensated 3xTF32 mode of the dense kernels vs fp64."""
import os, sys
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from tcct_b200 import ops as O
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv, DenseLinear
DEV = torch.device("cuda:0")

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())

def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
def nchw(t): return t.permute(0, 3, 1, 2).contiguous()

for prec in ("tf32", "tf32x3"):
    O.set_precision(prec)
    for (ks, B, H, W, xs, dys) in (((3, 3), 2, 16, 16, 1.0, 1.0), ((3, 3), 2, 16, 16, 1.0, 1e-4), ((1, 9), 2, 16, 16, 1.0, 1.0), ((3, 3), 2, 64, 64, 1.0, 1.0), ((3, 3), 2, 16, 16, 30.0, 1.0)):
        g = torch.Generator().manual_seed(5)
        mod = DenseConv(32, 32, ks).to(DEV)
        with torch.no_grad():
            mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.1)
            mod.bias.copy_(torch.randn(32, generator=g))
        for p in (mod.weight, mod.bias):
            p._gview = torch.zeros_like(p); p.grad = p._gview
        plan = PackPlan(mod, DEV)
        O.ARENA.reset(DEV); plan.run()
        x = torch.randn(B, 32, H, W, generator=g) * xs + (xs if xs > 1 else 0)
        dy = torch.randn(B, 32, H, W, generator=g) * dys
        xr = x.double().requires_grad_(True)
        wr, br = mod.weight.detach().cpu().double().requires_grad_(True), mod.bias.detach().cpu().double().requires_grad_(True)
        yr = F.conv2d(xr, wr, br, 1, (ks[0] // 2, ks[1] // 2)); yr.backward(dy.double())
        xg = nhwc(x).to(DEV).requires_grad_(True)
        y, st = mod.run(xg, want_stats=True, stats_act=O.ACT_LRELU)
        y.backward(nhwc(dy).to(DEV))
        print("%-7s conv %s %dx%dx%d xs=%g dys=%g: y %.2e dx %.2e dw %.2e db %.2e" % (prec, ks, B, H, W, xs, dys, rel(nchw(y), yr), rel(nchw(xg.grad), xr.grad), rel(mod.weight.grad, wr.grad), rel(mod.bias.grad, br.grad)))
    for (K, N, M) in ((64, 64, 512), (160, 160, 512)):
        g = torch.Generator().manual_seed(6)
        mod = DenseLinear(K, N).to(DEV)
        for p in (mod.weight, mod.bias):
            p._gview = torch.zeros_like(p); p.grad = p._gview
        plan = PackPlan(mod, DEV)
        O.ARENA.reset(DEV); plan.run()
        x = torch.randn(2, M // 2, K, generator=g); dy = torch.randn(2, M // 2, N, generator=g)
        xr = x.double().requires_grad_(True)
        wr, br = mod.weight.detach().cpu().double().requires_grad_(True), mod.bias.detach().cpu().double().requires_grad_(True)
        yr = F.linear(xr, wr, br); yr.backward(dy.double())
        xg = x.to(DEV).requires_grad_(True)
        y = mod.run(xg); y.backward(dy.to(DEV))
        print("%-7s linear %dx%d M=%d: y %.2e dx %.2e dw %.2e db %.2e" % (prec, K, N, M, rel(y, yr), rel(xg.grad, xr.grad), rel(mod.weight.grad, wr.grad), rel(mod.bias.grad, br.grad)))
O.set_precision("tf32")
