"""Call the bandwidth-bound ops once (after one warm-up call) at their largest K2 shapes - a small ncu target.
    python scripts/prof_ops.py [group ...]      groups: dw meta ln stem resize gemm wgrad bn head"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tcct_b200 import ops as O
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv, DenseLinear, DwConv

dev = torch.device("cuda:0")
groups = sys.argv[1:] or ["dw", "meta", "ln", "stem", "resize", "gemm", "wgrad", "bn", "head"]
B = 8

def attach(*ps):
    for p in ps:
        if p is not None:
            p._gview = torch.zeros_like(p); p.grad = p._gview

def run2(fn):
    O.ARENA.reset(dev); fn(); torch.cuda.synchronize()
    torch.cuda.profiler.start()          # ncu --profile-from-start off: only the second call is captured
    O.ARENA.reset(dev); fn(); torch.cuda.synchronize()
    torch.cuda.profiler.stop()

if "dw" in groups:
    for C, stride in ((64, 1), (96, 2)):
        m = DwConv(C, stride, bias=True).to(dev); attach(m.weight, m.bias)
        x = torch.randn(B, 128, 128, C, device=dev, requires_grad=True)
        def f():
            y, st = m.run(x, want_stats=True); y.backward(torch.ones_like(y))
        run2(f)
if "meta" in groups:
    t = torch.randn(B, 128 * 128, 64, device=dev, requires_grad=True); cur = torch.randn_like(t, requires_grad=True)
    sc = torch.ones(B, device=dev)
    def f():
        y = O.MetaPoolFn.apply(t, cur, sc); y.backward(torch.ones_like(y))
    run2(f)
if "ln" in groups:
    x = torch.randn(B, 128 * 128, 64, device=dev, requires_grad=True)
    ln = torch.nn.LayerNorm(64).to(dev); attach(ln.weight, ln.bias)
    def f():
        y = O.LayerNormFn.apply(x, ln.weight, ln.bias, 1e-6); y.backward(torch.ones_like(y))
    run2(f)
if "stem" in groups:
    img = torch.randn(B, 3, 256, 256, device=dev)
    w = torch.randn(32, 3, 3, 3, device=dev, requires_grad=True); b = torch.randn(32, device=dev, requires_grad=True); attach(w, b)
    for stride in (1, 2):
        def f():
            y, st = O.StemConvFn.apply(img, w, b, stride, True); y.backward(torch.ones_like(y))
        run2(f)
if "resize" in groups:
    x = torch.randn(B, 128, 128, 32, device=dev, requires_grad=True); add = torch.randn(B, 256, 256, 32, device=dev)
    for align in (True, False):
        def f():
            y = O.ResizeNHWCFn.apply(x, add, 256, 256, align, 1.0); y.backward(torch.ones_like(y))
        run2(f)
    lg = torch.randn(B, 5, 64, 64, device=dev, requires_grad=True)
    def f():
        y = O.ResizeNCHWFn.apply(lg, 256, 256); y.backward(torch.ones_like(y))
    run2(f)
if "gemm" in groups:
    for K, N, hw in ((64, 64, 128), (32, 32, 256), (128, 96, 128), (96, 96, 64)):
        m = DenseLinear(K, N).to(dev); attach(m.weight, m.bias)
        plan = PackPlan(m, dev); plan.run()
        x = torch.randn(B, hw * hw, K, device=dev, requires_grad=True)
        def f():
            plan.run(); y = m.run(x); y.backward(torch.ones_like(y))
        run2(f)
if "wgrad" in groups or "conv" in groups:
    for ks, hw in ((3, 256), ((1, 13), 256), ((13, 1), 256), (3, 128), (3, 32), (3, 16)):
        m = DenseConv(32, 32, ks).to(dev); attach(m.weight, m.bias)
        plan = PackPlan(m, dev); plan.run()
        x = torch.randn(B, hw, hw, 32, device=dev, requires_grad=True)
        def f():
            plan.run(); y, st = m.run(x, want_stats=True, stats_act=O.ACT_LRELU); y.backward(torch.ones_like(y))
        run2(f)
if "bn" in groups:
    a = torch.randn(B, 256, 256, 32, device=dev, requires_grad=True); b2 = torch.randn_like(a, requires_grad=True)
    bn1, bn2 = torch.nn.BatchNorm2d(32).to(dev), torch.nn.BatchNorm2d(32).to(dev)
    attach(bn1.weight, bn1.bias, bn2.weight, bn2.bias)
    def f():
        sa = O.ARENA.take(64, dev); sb = O.ARENA.take(64, dev)
        import tcct_b200._lib as L
        L.stats_nhwc(O._p(a), a.numel() // 32, 32, O._p(sa), O._stream()); L.stats_nhwc(O._p(b2), a.numel() // 32, 32, O._p(sb), O._stream())
        y = O.bn_act2(a, sa, bn1, O.ACT_LRELU, b2, sb, bn2, O.ACT_LRELU, O.ACT_GELU, True); y.backward(torch.ones_like(y))
    run2(f)
if "head" in groups:
    x = torch.randn(B, 256, 256, 32, device=dev, requires_grad=True)
    w = torch.randn(5, 32, 1, 1, device=dev, requires_grad=True); b = torch.randn(5, device=dev, requires_grad=True); attach(w, b)
    def f():
        y = O.HeadFn.apply(x, w, b); y.backward(torch.ones_like(y))
    run2(f)
print("done")
