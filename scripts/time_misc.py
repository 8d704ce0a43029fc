"""Forward and backward of the bandwidth-bound ops at their largest K2 shapes, in-graph, against the compulsory bytes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit, side_stream
from tcct_b200 import ops as O
O.WGRAD_ASYNC = False
dev = torch.device("cuda:0")
MB = 1e6


def run(name, make, fwd, fwd_mb, bwd_mb):
    st = side_stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):          # everything autograd touches is created on the stream the capture runs on
        _run(name, make, fwd, fwd_mb, bwd_mb)
    torch.cuda.current_stream().wait_stream(st)


def _run(name, make, fwd, fwd_mb, bwd_mb):
    ins = make()
    def f():
        O.ARENA.reset(dev)
        with torch.no_grad():
            fwd(*ins)
    tf = timeit(f)
    def fb():
        O.ARENA.reset(dev)
        out = fwd(*ins)
        out = out[0] if isinstance(out, tuple) else out
        torch.autograd.grad(out, [t for t in ins if t.requires_grad], gout[0])
    O.ARENA.reset(dev)
    o = fwd(*ins); o = o[0] if isinstance(o, tuple) else o
    gout = [torch.randn_like(o)]
    tfb = timeit(fb)
    tb = tfb - tf
    print("%-34s fwd %6.1f us (%5.0f GB/s of %5.0f MB) | bwd %6.1f us (%5.0f GB/s of %5.0f MB)" % (
        name, tf, fwd_mb * MB / tf / 1e3, fwd_mb, tb, bwd_mb * MB / tb / 1e3, bwd_mb), flush=True)


def P(*shape):
    return torch.nn.Parameter(torch.randn(*shape, device=dev) * 0.1)


class Holder(torch.nn.Module):
    def __init__(s, **kw):
        super().__init__()
        for k, v in kw.items():
            setattr(s, k, v)


from tcct_b200.nets.flat import FlatParams
B = 8
for (H, W, C) in ((128, 128, 64), (64, 64, 96)):
    n = B * H * W * C * 4 / MB
    h = Holder(w=P(C, 1, 3, 3), b=P(C), g=P(C), be=P(C))
    FlatParams(h, dev)
    x = lambda: (torch.randn(B, H, W, C, device=dev, requires_grad=True),)
    run("dwconv3 s1 +bias+res %dx%dx%d" % (H, W, C), x, lambda t: O.DwConv3Fn.apply(t, h.w, h.b, 1, True, False), 2 * n, 3 * n)
    run("dwconv3 s1 stats %dx%dx%d" % (H, W, C), x, lambda t: O.DwConv3Fn.apply(t, h.w, None, 1, False, True), 2 * n, 3 * n)
    run("dwconv3 s2 %dx%dx%d" % (H, W, C), x, lambda t: O.DwConv3Fn.apply(t, h.w, None, 2, False, False), 1.25 * n, 1.5 * n)
    run("layernorm %dx%dx%d" % (H, W, C), x, lambda t: O.LayerNormFn.apply(t, h.g, h.be, 1e-6), 2 * n, 3 * n)
    x2 = lambda: (torch.randn(B, H * W, C, device=dev, requires_grad=True), torch.randn(B, H * W, C, device=dev, requires_grad=True))
    run("metapool %dx%dx%d" % (H, W, C), x2, lambda t, c: O.MetaPoolFn.apply(t, c, None), 3 * n, 3 * n)
    run("maxpool2 %dx%dx%d" % (H, W, C), x, lambda t: O.MaxPool2Fn.apply(t), 1.25 * n, 1.5 * n)
n = B * 256 * 256 * 32 * 4 / MB
x = lambda: (torch.randn(B, 256, 256, 32, device=dev, requires_grad=True),)
run("maxpool2 256x256x32", x, lambda t: O.MaxPool2Fn.apply(t), 1.25 * n, 1.5 * n)
x = lambda: (torch.randn(B, 128, 128, 32, device=dev, requires_grad=True), torch.randn(B, 256, 256, 32, device=dev, requires_grad=True))
run("resize x2 align + skip 128->256 x32", x, lambda t, s: O.ResizeNHWCFn.apply(t, s, 256, 256, True, 1.0), 2.25 * n, 1.25 * n)
hh = Holder(w=P(5, 32, 1, 1), b=P(5)); FlatParams(hh, dev)
x = lambda: (torch.randn(B, 256, 256, 32, device=dev, requires_grad=True),)
run("head 32->5 256x256", x, lambda t: O.HeadFn.apply(t, hh.w, hh.b), n * (1 + 5 / 32), n * (2 + 5 / 32))
# stems: 3 -> 32 channels, NCHW image in, NHWC out (CrossResNet.cnn at stride 1, MPViT.stem[0] at stride 2): direct C-ABI calls
import tcct_b200._lib as L
from tcct_b200.ops import _p, _stream
w_, b_ = torch.randn(32, 3, 3, 3, device=dev) * 0.1, torch.randn(32, device=dev) * 0.1
dw_, db_ = torch.zeros_like(w_), torch.zeros_like(b_)
for stride in (1, 2):
    img = torch.rand(B, 3, 256, 256, device=dev)
    Ho = 256 // stride
    ys = [torch.empty(B, Ho, Ho, 32, device=dev) for _ in range(3)]
    st = torch.zeros(64, dtype=torch.float64, device=dev)
    k = [0]
    def sf():
        k[0] += 1
        L.stem_conv_fwd(_p(img), _p(w_), _p(b_), _p(ys[k[0] % 3]), B, 256, 256, stride, _p(st), _stream())
    def sw():
        k[0] += 1
        L.stem_conv_wgrad(_p(img), _p(ys[k[0] % 3]), _p(dw_), _p(db_), B, 256, 256, stride, _stream())
    tf, tw = timeit(sf), timeit(sw)
    mb = (B * 3 * 256 * 256 * 4 + B * Ho * Ho * 32 * 4) / MB
    print("stem conv 3->32 stride %d 256x256: fwd %.1f us (%.0f GB/s of %.0f MB) | wgrad %.1f us (%.0f GB/s)" % (stride, tf, mb * MB / tf / 1e3, mb, tw, mb * MB / tw / 1e3), flush=True)
