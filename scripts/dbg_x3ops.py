"""Accuracy of the tf32x3 mode per operator against fp64 (which operator limits the CrossResNet-branch gradients?)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from tcct_b200 import ops as O
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv
DEV = torch.device("cuda:0")
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
def nchw(t): return t.permute(0, 3, 1, 2).contiguous()
def rl2(a, b): return float((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm())
g = torch.Generator().manual_seed(0)
O.set_precision("tf32x3")
for cin, cout, ks, B, H, W in [(32, 32, (3, 3), 2, 4, 4), (32, 32, (3, 3), 2, 8, 8), (32, 32, (3, 3), 2, 64, 64), (32, 32, (1, 13), 2, 64, 64),
                               (32, 32, (13, 1), 2, 64, 64), (32, 32, (1, 5), 2, 4, 4), (32, 32, (5, 1), 2, 4, 4), (64, 64, (3, 3), 2, 32, 32),
                               (32, 32, (1, 11), 2, 32, 32), (32, 32, (1, 7), 2, 8, 8)]:
    mod = DenseConv(cin, cout, ks).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.1); mod.bias.copy_(torch.randn(cout, generator=g))
    plan = PackPlan(mod, DEV)
    O.ARENA.reset(DEV); plan.run()
    x = torch.randn(B, cin, H, W, generator=g); dy = torch.randn(B, cout, H, W, generator=g)
    xr = x.double().requires_grad_(True); wr = mod.weight.detach().cpu().double().requires_grad_(True); br = mod.bias.detach().cpu().double().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, 1, (ks[0] // 2, ks[1] // 2)); yr.backward(dy.double())
    xg = nhwc(x).to(DEV).requires_grad_(True)
    y, st = mod.run(xg, want_stats=True, stats_act=O.ACT_LRELU)
    y.backward(nhwc(dy).to(DEV))
    act = F.leaky_relu(yr, 0.01)
    ref_stats = torch.cat([act.sum((0, 2, 3)), (act * act).sum((0, 2, 3))])
    print("conv %d->%d %s @%dx%dx%d  y %.1e dx %.1e dw %.1e db %.1e stats %.1e" % (cin, cout, ks, B, H, W, rl2(nchw(y), yr), rl2(nchw(xg.grad), xr.grad),
          rl2(mod.weight.grad, wr.grad), rl2(mod.bias.grad, br.grad), rl2(st, ref_stats)))
O.set_precision("tf32")
# BatchNorm + activation, forward and backward, few samples per channel
for C, B, H, W in [(32, 2, 4, 4), (32, 2, 8, 8), (32, 2, 64, 64), (256, 2, 4, 4)]:
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * torch.randn(C, generator=g)); bn.bias.copy_(0.1 * torch.randn(C, generator=g))
    a = torch.randn(B, C, H, W, generator=g) * 3 + 1; dout = torch.randn(B, C, H, W, generator=g)
    ar = a.double().requires_grad_(True)
    ref = F.batch_norm(F.leaky_relu(ar, 0.01), None, None, bn.weight.detach().cpu().double(), bn.bias.detach().cpu().double(), True, 0.1, 1e-5)
    ref.backward(dout.double())
    ag = nhwc(a).to(DEV).requires_grad_(True)
    O.ARENA.reset(DEV)
    act = F.leaky_relu(ag.detach(), 0.01)
    stats = torch.cat([act.sum((0, 1, 2)), (act * act).sum((0, 1, 2))]).double()
    for p_ in (bn.weight, bn.bias):
        p_._gview = torch.zeros_like(p_); p_.grad = p_._gview
    out = O.bn_act2(ag, stats, bn, O.ACT_LRELU, training=True)
    out.backward(nhwc(dout).to(DEV))
    print("bn C=%d @%dx%dx%d  out %.1e da %.1e" % (C, B, H, W, rl2(nchw(out), ref), rl2(nchw(ag.grad), ar.grad)))
