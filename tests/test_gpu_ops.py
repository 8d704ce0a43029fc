"""GPU: every kernel behind the C ABI against the plain PyTorch fp32 restatement of the same reference op
(torch CPU ops are what oracle/tcct_oracle.py is built from).  Dense contractions run in TF32 on the
tensor cores: tolerance 5e-3 of max|ref|; everything else is fp32: 1e-4."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from tcct_b200 import ops as O  # noqa: E402
from tcct_b200.nets.flat import PackPlan  # noqa: E402
from tcct_b200.nets.tcct import DenseConv, DenseLinear, DwConv  # noqa: E402

DEV = torch.device("cuda:0")
TF32, FP32 = 5e-3, 1e-4


def close(got, ref, tol, name=""):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    err = float((got - ref).abs().max())
    scale = float(ref.abs().max()) + 1e-12
    assert err <= tol * scale, "%s: max|d| %.3e vs max|ref| %.3e (rel %.2e > %.1e)" % (name, err, scale, err / scale, tol)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def attach(*params):
    for p in params:
        if p is not None:
            p._gview = torch.zeros_like(p)
            p.grad = p._gview


def begin():
    O.ARENA.reset(DEV)


def gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("cin,cout,ks,B,H,W", [
    (32, 32, (3, 3), 2, 32, 48), (32, 32, (1, 13), 2, 32, 32), (32, 32, (13, 1), 1, 48, 32), (32, 32, (1, 5), 1, 16, 16),
    (32, 32, (7, 1), 2, 24, 40), (32, 64, (3, 3), 2, 32, 32), (32, 32, (3, 3), 3, 38, 22),
    # the wide layers of stc_tb / gtc_tb (tcct.py:861-864): reductions cut into 64 / 32-channel slices
    (64, 96, (3, 3), 2, 24, 24), (96, 128, (1, 9), 1, 16, 32), (128, 256, (7, 1), 1, 16, 16), (256, 128, (3, 3), 1, 8, 16),
    (64, 32, (3, 3), 2, 32, 32), (256, 256, (3, 3), 2, 4, 4), (64, 64, (11, 1), 1, 32, 16), (32, 64, (1, 11), 1, 16, 32)])
def test_conv2d_dense(cin, cout, ks, B, H, W):
    g = gen(1)
    mod = DenseConv(cin, cout, ks).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.1)
        mod.bias.copy_(torch.randn(cout, generator=g))
    plan = PackPlan(mod, DEV)
    begin(); plan.run()
    x = torch.randn(B, cin, H, W, generator=g)
    dy = torch.randn(B, cout, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    wr, br = mod.weight.detach().cpu().requires_grad_(True), mod.bias.detach().cpu().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, 1, (ks[0] // 2, ks[1] // 2))
    yr.backward(dy)
    xg = nhwc(x).to(DEV).requires_grad_(True)
    y, stats = mod.run(xg, want_stats=True, stats_act=O.ACT_LRELU)
    y.backward(nhwc(dy).to(DEV))
    close(nchw(y), yr, TF32, "y")
    close(nchw(xg.grad), xr.grad, TF32, "dx")
    close(mod.weight.grad, wr.grad, TF32, "dw")
    close(mod.bias.grad, br.grad, TF32, "db")
    act = F.leaky_relu(yr, 0.01)
    ref_stats = torch.cat([act.sum((0, 2, 3)), (act * act).sum((0, 2, 3))]).double()
    close(stats, ref_stats, TF32, "stats")


@pytest.mark.parametrize("ks,B,H,W", [((3, 3), 2, 24, 256), ((3, 3), 3, 130, 128), ((1, 13), 1, 20, 128), ((13, 1), 2, 40, 128),
                                       ((1, 5), 1, 7, 384), ((11, 1), 1, 24, 256), ((3, 3), 8, 256, 256),
                                       ((1, 13), 2, 128, 40), ((1, 11), 1, 256, 24), ((1, 13), 8, 256, 256), ((13, 1), 8, 256, 256)])
def test_conv2d_tma(ks, B, H, W):
    """The TMA-fed tcgen05 kernel (csrc/conv_tma.cu) against the fp32 reference and the warp-level mma.sync kernel:
    forward (+bias, LeakyReLU statistics) and data gradient."""
    import tcct_b200._lib as L
    assert L.tcct_conv_tma_supported(H, W, 32, 32, ks[0], ks[1]) == 1
    g = gen(22)
    mod = DenseConv(32, 32, ks).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.1)
        mod.bias.copy_(torch.randn(32, generator=g))
    plan = PackPlan(mod, DEV)
    x = torch.randn(B, 32, H, W, generator=g)
    dy = torch.randn(B, 32, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    wr, br = mod.weight.detach().cpu().requires_grad_(True), mod.bias.detach().cpu().requires_grad_(True)
    torch.set_num_threads(8)
    yr = F.conv2d(xr, wr, br, 1, (ks[0] // 2, ks[1] // 2))
    yr.backward(dy)
    res = {}
    for tma in (True, False):
        O.set_umma(tma)
        try:
            begin(); plan.run()
            attach(mod.weight, mod.bias)
            xg = nhwc(x).to(DEV).requires_grad_(True)
            y, stats = mod.run(xg, want_stats=True, stats_act=O.ACT_LRELU)
            y.backward(nhwc(dy).to(DEV))
            torch.cuda.synchronize()
            res[tma] = (y.detach().clone(), xg.grad.clone(), stats.clone(), mod.weight.grad.clone(), mod.bias.grad.clone())
        finally:
            O.set_umma(True)
    y, dx, stats = res[True][:3]
    close(nchw(y), yr, TF32, "y")
    close(nchw(dx), xr.grad, TF32, "dx")
    act = F.leaky_relu(yr, 0.01)
    close(stats, torch.cat([act.sum((0, 2, 3)), (act * act).sum((0, 2, 3))]).double(), TF32, "stats")
    # both are single-pass TF32: tcgen05 truncates the fp32 activations to TF32 in hardware, the mma.sync kernel rounds
    # them (cvt.rna) -> agreement at the TF32 level (2^-11 per operand), not bit for bit
    close(y, res[False][0], 2e-3, "tcgen05 vs mma.sync (y)")
    close(dx, res[False][1], 2e-3, "tcgen05 vs mma.sync (dx)")
    close(res[True][3], wr.grad, TF32, "dw")
    close(res[True][4], br.grad, TF32, "db")
    close(res[False][3], wr.grad, TF32, "dw (mma.sync)")


@pytest.mark.parametrize("cout,ks,B,H,W", [(64, (3, 3), 2, 40, 128), (64, (1, 13), 1, 20, 256)])
def test_wgrad_tma_wide_output(cout, ks, B, H, W):
    """32 -> 64 channel convs (MPViT stem): forward / dgrad on the mma.sync kernels, the weight gradient through the tcgen05
    line kernel, one launch pair per 32 output channels."""
    import tcct_b200._lib as L
    assert L.tcct_wgrad_tma_supported(H, W, 32, cout, ks[0], ks[1]) == 1
    g = gen(23)
    mod = DenseConv(32, cout, ks).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.1)
        mod.bias.copy_(torch.randn(cout, generator=g))
    plan = PackPlan(mod, DEV)
    x = torch.randn(B, 32, H, W, generator=g)
    dy = torch.randn(B, cout, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    wr, br = mod.weight.detach().cpu().requires_grad_(True), mod.bias.detach().cpu().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, 1, (ks[0] // 2, ks[1] // 2))
    yr.backward(dy)
    begin(); plan.run()
    attach(mod.weight, mod.bias)
    xg = nhwc(x).to(DEV).requires_grad_(True)
    y, _ = mod.run(xg)
    y.backward(nhwc(dy).to(DEV))
    torch.cuda.synchronize()
    close(nchw(y), yr, TF32, "y")
    close(nchw(xg.grad), xr.grad, TF32, "dx")
    close(mod.weight.grad, wr.grad, TF32, "dw")
    close(mod.bias.grad, br.grad, TF32, "db")


@pytest.mark.parametrize("K,N,M,use_res", [(32, 32, 1000, False), (64, 64, 4096, True), (96, 32, 777, False),
                                            (160, 160, 512, True), (128, 96, 300, False),
                                            # M % 128 == 0 and >= 8192: the TMA-fed tcgen05 GEMM (csrc/gemm_tma.cu), forward and dgrad
                                            (64, 64, 16384, True), (32, 32, 65536, False), (128, 96, 8192, True), (96, 128, 8192, False),
                                            (160, 32, 8192, False), (32, 160, 16384, True)])
def test_gemm_linear(K, N, M, use_res):
    g = gen(2)
    B = 4
    M = M // B * B
    mod = DenseLinear(K, N).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(torch.randn(N, K, generator=g) * 0.1)
        mod.bias.copy_(torch.randn(N, generator=g))
    plan = PackPlan(mod, DEV)
    begin(); plan.run()
    x = torch.randn(B, M // B, K, generator=g)
    res = torch.randn(B, M // B, N, generator=g) if use_res else None
    rs = torch.tensor([1.0, 0.0, 1.25, 2.0]) if use_res else None
    dy = torch.randn(B, M // B, N, generator=g)
    xr = x.clone().requires_grad_(True)
    rr = res.clone().requires_grad_(True) if use_res else None
    wr, br = mod.weight.detach().cpu().requires_grad_(True), mod.bias.detach().cpu().requires_grad_(True)
    yr = F.linear(xr, wr, br)
    if use_res:
        yr = rr + rs.view(B, 1, 1) * yr
    yr.backward(dy)
    xg = x.to(DEV).requires_grad_(True)
    rg = res.to(DEV).requires_grad_(True) if use_res else None
    import tcct_b200._lib as L
    if M >= 8192:
        assert L.tcct_gemm_tma_supported(M, K, N) == 1 and L.tcct_gemm_tma_supported(M, N, K) == 1
    torch.set_num_threads(8)
    y = mod.run(xg, res=rg, res_scale=rs.to(DEV) if use_res else None)
    y.backward(dy.to(DEV))
    close(y, yr, TF32, "y")
    close(xg.grad, xr.grad, TF32, "dx")
    close(mod.weight.grad, wr.grad, TF32, "dw")
    close(mod.bias.grad, br.grad, TF32, "db")
    if use_res:
        close(rg.grad, rr.grad, FP32, "dres")


@pytest.mark.parametrize("B,H,W", [(2, 16, 24), (2, 64, 128)])      # small: mma.sync kernel; large: TMA-fed tcgen05 GEMM
def test_gemm_concat_slices(B, H, W):
    """aggregate: 1x1 conv over cat[r, t] computed as two accumulating GEMMs (MHCA_stage.forward, tcct.py:604-616)."""
    g = gen(3)
    C, N = 64, 96
    mod = DenseConv(2 * C, N, 1, bias=False, k_slices=[(0, C), (C, C)]).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.1)
    plan = PackPlan(mod, DEV)
    begin(); plan.run()
    r, t, dy = (torch.randn(B, c, H, W, generator=g) for c in (C, C, N))
    rr, tr = r.clone().requires_grad_(True), t.clone().requires_grad_(True)
    wr = mod.weight.detach().cpu().requires_grad_(True)
    yr = F.conv2d(torch.cat([rr, tr], 1), wr)
    yr.backward(dy)
    rg, tg = nhwc(r).to(DEV).requires_grad_(True), nhwc(t).to(DEV).requires_grad_(True)
    y0, _ = mod.run(rg, part=0)
    y, stats = mod.run(tg, want_stats=True, res=y0, part=1)
    y.backward(nhwc(dy).to(DEV))
    close(nchw(y), yr, TF32, "y")
    close(nchw(rg.grad), rr.grad, TF32, "dr")
    close(nchw(tg.grad), tr.grad, TF32, "dt")
    close(mod.weight.grad, wr.grad, TF32, "dw")
    close(stats, torch.cat([yr.sum((0, 2, 3)), (yr * yr).sum((0, 2, 3))]).double(), TF32, "stats")


@pytest.mark.parametrize("C,pre,post,dual,bn_b", [(32, O.ACT_LRELU, O.ACT_GELU, True, True), (64, O.ACT_NONE, O.ACT_HSWISH, False, False),
                                                   (96, O.ACT_NONE, O.ACT_NONE, True, False), (32, O.ACT_NONE, O.ACT_LRELU, False, False),
                                                   (160, O.ACT_NONE, O.ACT_NONE, True, True)])
def test_bn_act2(C, pre, post, dual, bn_b):
    g = gen(4)
    B, H, W = 3, 10, 12
    acts = {O.ACT_NONE: lambda v: v, O.ACT_LRELU: lambda v: F.leaky_relu(v, 0.01), O.ACT_HSWISH: F.hardswish, O.ACT_GELU: F.gelu}
    bnA, bnB = torch.nn.BatchNorm2d(C), torch.nn.BatchNorm2d(C)
    for bn in (bnA, bnB):
        with torch.no_grad():
            bn.weight.copy_(1 + 0.2 * torch.randn(C, generator=g)); bn.bias.copy_(0.2 * torch.randn(C, generator=g))
    a = torch.randn(B, C, H, W, generator=g) * 2 + 0.5
    b = torch.randn(B, C, H, W, generator=g) if dual else None
    dy = torch.randn(B, C, H, W, generator=g)
    ar = a.clone().requires_grad_(True)
    brf = b.clone().requires_grad_(True) if dual else None
    z = bnA(acts[pre](ar))
    if dual:
        z = z + (bnB(acts[pre](brf)) if bn_b else brf)
    outr = acts[post](z)
    outr.backward(dy)
    import copy
    gA, gB = copy.deepcopy(bnA).to(DEV), copy.deepcopy(bnB).to(DEV)
    for bn, src in ((gA, bnA), (gB, bnB)):      # start from the pre-update running statistics
        bn.running_mean.zero_(); bn.running_var.fill_(1.0); bn.num_batches_tracked.zero_()
        attach(bn.weight, bn.bias)
    begin()
    ag = nhwc(a).to(DEV).requires_grad_(True)
    bg = nhwc(b).to(DEV).requires_grad_(True) if dual else None

    def stats_of(t, act):
        u = acts[act](t)
        return torch.cat([u.sum((0, 2, 3)), (u * u).sum((0, 2, 3))]).double().to(DEV)
    out = O.bn_act2(ag, stats_of(a, pre), gA, pre, bg, stats_of(b, pre) if (dual and bn_b) else None, gB if bn_b else None,
                    pre if bn_b else O.ACT_NONE, post, True)
    out.backward(nhwc(dy).to(DEV))
    close(nchw(out), outr, FP32, "out")
    close(nchw(ag.grad), ar.grad, 2e-4, "da")
    close(gA.weight.grad, bnA.weight.grad, 2e-4, "dgammaA")
    close(gA.bias.grad, bnA.bias.grad, 2e-4, "dbetaA")
    close(gA.running_mean, bnA.running_mean, FP32, "running_mean")
    close(gA.running_var, bnA.running_var, FP32, "running_var")
    assert int(gA.num_batches_tracked) == 1
    if dual:
        close(nchw(bg.grad), brf.grad, 2e-4, "db")
        if bn_b:
            close(gB.weight.grad, bnB.weight.grad, 2e-4, "dgammaB")


@pytest.mark.parametrize("C,B,H,W", [(32, 2, 96, 80), (64, 3, 50, 30), (320, 2, 16, 16)])
def test_bn_act2_backward_grid_barrier_path(C, B, H, W):
    """The single-launch backward with many CTAs, ragged pixel chunks and the grid barrier (train-mode BN + LeakyReLU)."""
    g = gen(41)
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.2 * torch.randn(C, generator=g)); bn.bias.copy_(0.2 * torch.randn(C, generator=g))
    a = torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3
    dy = torch.randn(B, C, H, W, generator=g)
    ar = a.clone().requires_grad_(True)
    outr = bn(F.leaky_relu(ar, 0.01))
    outr.backward(dy)
    import copy
    gbn = copy.deepcopy(bn).to(DEV)
    gbn.running_mean.zero_(); gbn.running_var.fill_(1.0); gbn.num_batches_tracked.zero_()
    attach(gbn.weight, gbn.bias)
    begin()
    ag = nhwc(a).to(DEV).requires_grad_(True)
    u = F.leaky_relu(a, 0.01)
    st = torch.cat([u.sum((0, 2, 3)), (u * u).sum((0, 2, 3))]).double().to(DEV)
    out = O.bn_act2(ag, st, gbn, O.ACT_LRELU, training=True)
    out.backward(nhwc(dy).to(DEV))
    close(nchw(out), outr, FP32, "out")
    close(nchw(ag.grad), ar.grad, 2e-4, "da")
    close(gbn.weight.grad, bn.weight.grad, 2e-4, "dgamma")
    close(gbn.bias.grad, bn.bias.grad, 2e-4, "dbeta")


def test_bn_eval_mode():
    g = gen(5)
    C = 32
    bn = torch.nn.BatchNorm2d(C).eval()
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(C, generator=g)); bn.running_var.copy_(0.5 + torch.rand(C, generator=g))
        bn.weight.copy_(torch.randn(C, generator=g))
    a = torch.randn(2, C, 8, 8, generator=g)
    ref = F.hardswish(bn(a))
    import copy
    gbn = copy.deepcopy(bn).to(DEV)
    begin()
    out = O.bn_act2(nhwc(a).to(DEV), None, gbn, post=O.ACT_HSWISH, training=False)
    close(nchw(out), ref, FP32)


def test_maxpool2():
    g = gen(6)
    x = torch.randn(2, 32, 12, 20, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 2)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xg = nhwc(x).to(DEV).requires_grad_(True)
    y = O.MaxPool2Fn.apply(xg)
    y.backward(nhwc(dy).to(DEV))
    close(nchw(y), yr, 0.0)
    close(nchw(xg.grad), xr.grad, 0.0)


@pytest.mark.parametrize("C,stride,bias,add", [(64, 1, False, False), (96, 2, False, False), (128, 1, True, True), (160, 2, False, False)])
def test_dwconv3(C, stride, bias, add):
    g = gen(7)
    mod = DwConv(C, stride, bias).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(torch.randn(mod.weight.shape, generator=g))
        if bias:
            mod.bias.copy_(torch.randn(C, generator=g))
    x = torch.randn(2, C, 14, 18, generator=g)
    xr = x.clone().requires_grad_(True)
    wr = mod.weight.detach().cpu().requires_grad_(True)
    br = mod.bias.detach().cpu().requires_grad_(True) if bias else None
    yr = F.conv2d(xr, wr, br, stride, 1, 1, C)
    if add:
        yr = yr + xr
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    begin()
    xg = nhwc(x).to(DEV).requires_grad_(True)
    y, stats = mod.run(xg, add_input=add, want_stats=True)
    y.backward(nhwc(dy).to(DEV))
    close(nchw(y), yr, FP32, "y")
    close(nchw(xg.grad), xr.grad, FP32, "dx")
    close(mod.weight.grad, wr.grad, 2e-4, "dw")
    if bias:
        close(mod.bias.grad, br.grad, 2e-4, "db")
    close(stats, torch.cat([yr.sum((0, 2, 3)), (yr * yr).sum((0, 2, 3))]).double(), FP32, "stats")


@pytest.mark.parametrize("C", [64, 96, 160])
def test_layernorm(C):
    g = gen(8)
    x = torch.randn(2, 50, C, generator=g) * 3 + 1
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (C,), gr, br, 1e-6)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xg = x.to(DEV).requires_grad_(True)
    gg, bg = gamma.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    y = O.LayerNormFn.apply(xg, gg, bg, 1e-6)
    y.backward(dy.to(DEV))
    close(y, yr, FP32, "y"); close(xg.grad, xr.grad, 2e-4, "dx"); close(gg.grad, gr.grad, 2e-4, "dg"); close(bg.grad, br.grad, 2e-4, "db")


def test_metapool_token_channel_plane():
    g = gen(9)
    B, N, C = 3, 40, 64
    t, cur = torch.randn(B, N, C, generator=g), torch.randn(B, N, C, generator=g)
    scale = torch.tensor([1.0, 0.0, 1.0 / 0.9])
    tr, cr = t.clone().requires_grad_(True), cur.clone().requires_grad_(True)
    outr = tr + scale.view(B, 1, 1) * (F.avg_pool2d(cr, 3, 1, 1, count_include_pad=False) - cr)
    dy = torch.randn(outr.shape, generator=g)
    outr.backward(dy)
    tg, cg = t.to(DEV).requires_grad_(True), cur.to(DEV).requires_grad_(True)
    out = O.MetaPoolFn.apply(tg, cg, scale.to(DEV))
    out.backward(dy.to(DEV))
    close(out, outr, FP32); close(tg.grad, tr.grad, FP32); close(cg.grad, cr.grad, FP32)


@pytest.mark.parametrize("align,h,w,f,with_add", [(True, 8, 12, 2, True), (False, 8, 12, 2, False), (False, 6, 4, 4, True), (False, 16, 16, 1, False)])
def test_resize_nhwc(align, h, w, f, with_add):
    g = gen(10)
    B, C = 2, 32
    x = torch.randn(B, C, h, w, generator=g)
    add = torch.randn(B, C, h * f, w * f, generator=g) if with_add else None
    xr = x.clone().requires_grad_(True)
    addr = add.clone().requires_grad_(True) if with_add else None
    yr = 0.5 * F.interpolate(xr, size=(h * f, w * f), mode="bilinear", align_corners=align)
    if with_add:
        yr = yr + addr
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xg = nhwc(x).to(DEV).requires_grad_(True)
    ag = nhwc(add).to(DEV).requires_grad_(True) if with_add else None
    y = O.ResizeNHWCFn.apply(xg, ag, h * f, w * f, align, 0.5)
    y.backward(nhwc(dy).to(DEV))
    close(nchw(y), yr, FP32, "y"); close(nchw(xg.grad), xr.grad, FP32, "dx")
    if with_add:
        close(nchw(ag.grad), addr.grad, FP32, "dadd")


@pytest.mark.parametrize("f", [2, 4, 8])
def test_resize_nchw_logits(f):
    g = gen(11)
    x = torch.randn(2, 5, 6, 10, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = F.interpolate(xr, size=(6 * f, 10 * f), mode="bilinear", align_corners=False)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xg = x.to(DEV).requires_grad_(True)
    y = O.ResizeNCHWFn.apply(xg, 6 * f, 10 * f)
    y.backward(dy.to(DEV))
    close(y, yr, FP32); close(xg.grad, xr.grad, FP32)


def test_l2norm32():
    g = gen(12)
    x = torch.randn(2, 32, 9, 7, generator=g)
    xr = x.clone().requires_grad_(True)
    yr = F.normalize(xr, dim=1, p=2)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xg = nhwc(x).to(DEV).requires_grad_(True)
    y = O.L2Norm32Fn.apply(xg)
    y.backward(nhwc(dy).to(DEV))
    close(nchw(y), yr, FP32); close(nchw(xg.grad), xr.grad, FP32)


def test_norm_add3_matches_reference_norm_add():
    """norm_add of the reference (tcct.py:937-942): mean of the L2-normalised maps, bilinearly up-sampled."""
    from tcct_b200.nets.tcct import norm_add
    g = gen(13)
    xs = [torch.randn(2, 32, 32, 48, generator=g), torch.randn(2, 32, 16, 24, generator=g), torch.randn(2, 32, 8, 12, generator=g)]
    xr = [x.clone().requires_grad_(True) for x in xs]
    ref = sum(F.interpolate(F.normalize(x, dim=1, p=2), size=(32, 48), mode="bilinear", align_corners=False) for x in xr) / 3
    dy = torch.randn(ref.shape, generator=g)
    ref.backward(dy)
    xg = [nhwc(x).to(DEV).requires_grad_(True) for x in xs]
    y = norm_add(xg)[0]
    y.backward(nhwc(dy).to(DEV))
    close(nchw(y), ref, FP32)
    for a, b in zip(xg, xr):
        close(nchw(a.grad), b.grad, FP32)


@pytest.mark.parametrize("stride,bias", [(1, True), (2, False)])
def test_stem_conv(stride, bias):
    g = gen(13)
    img = torch.rand(2, 3, 20, 28, generator=g)
    w = torch.randn(32, 3, 3, 3, generator=g)
    b = torch.randn(32, generator=g) if bias else None
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True) if bias else None
    yr = F.conv2d(img, wr, br, stride, 1)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    begin()
    wg = w.to(DEV).requires_grad_(True)
    bg = b.to(DEV).requires_grad_(True) if bias else None
    y, stats = O.StemConvFn.apply(img.to(DEV), wg, bg, stride, True)
    y.backward(nhwc(dy).to(DEV))
    close(nchw(y), yr, FP32, "y"); close(wg.grad, wr.grad, 2e-4, "dw")
    if bias:
        close(bg.grad, br.grad, 2e-4, "db")
    close(stats, torch.cat([yr.sum((0, 2, 3)), (yr * yr).sum((0, 2, 3))]).double(), FP32, "stats")


@pytest.mark.parametrize("C", [5, 9])
def test_head(C):
    g = gen(14)
    x = torch.randn(2, 32, 11, 13, generator=g)
    w, b = torch.randn(C, 32, 1, 1, generator=g), torch.randn(C, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, br)
    dy = torch.randn(yr.shape, generator=g)
    yr.backward(dy)
    xg = nhwc(x).to(DEV).requires_grad_(True)
    wg, bg = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    y = O.HeadFn.apply(xg, wg, bg)
    y.backward(dy.to(DEV))
    close(y, yr, FP32, "y"); close(nchw(xg.grad), xr.grad, FP32, "dx"); close(wg.grad, wr.grad, 2e-4, "dw"); close(bg.grad, br.grad, 2e-4, "db")


@pytest.mark.parametrize("C,mode", [(5, 0), (9, 0), (5, 1)])
def test_dice_multiloss(C, mode):
    import tcct_oracle as orc
    g = gen(15)
    B, H, W = 2, 24, 40
    logits = torch.randn(B, C, H, W, generator=g) * 3
    lab = torch.randint(0, C, (B, H, W), generator=g)
    onehot = F.one_hot(lab, C).permute(0, 3, 1, 2)
    lr = logits.clone().requires_grad_(True)
    if mode == 0:
        ref = orc.multi_dice(lr, onehot)
    else:
        p = torch.softmax(lr, 1)
        ref = sum(F.mse_loss(p[:, i], onehot[:, i].float()) for i in range(C))
    (ref * 0.7).backward()
    begin()
    lg = logits.to(DEV).requires_grad_(True)
    lab8 = O.labels_u8(onehot.to(DEV).contiguous(), C)
    assert torch.equal(lab8.cpu().long(), lab)
    assert torch.equal(O.labels_u8(lab.to(DEV), C).cpu().long(), lab)
    loss = O.DiceFn.apply(lg, lab8, mode)
    (loss * 0.7).backward()
    close(loss, ref, 1e-5, "loss"); close(lg.grad, lr.grad, 2e-4, "dlogits")


@pytest.mark.parametrize("C,B,H,W", [(5, 2, 64, 96), (9, 1, 40, 72), (5, 8, 256, 256)])
def test_dice_multi_matches_oracle(C, B, H, W):
    """ops.DiceMultiFn (csrc/dice_multi.cu): deep-supervision Dice over [z0, z1, z2, z3] with the auxiliary logits at native
    resolution, against F.interpolate + the oracle's multi_dice (kite/loopback.py:62-73, nets/tcct.py:1042-1044)."""
    import tcct_oracle as orc
    g = gen(31)
    lab = torch.randint(0, C, (B, H, W), generator=g)
    onehot = F.one_hot(lab, C).permute(0, 3, 1, 2)
    zs = [torch.randn(B, C, H // f, W // f, generator=g) * 2 for f in (1, 2, 4, 8)]
    w_aux = 0.7
    ref_in = [z.clone().requires_grad_(True) for z in zs]
    ups = [ref_in[0]] + [F.interpolate(z, size=(H, W), mode="bilinear", align_corners=False) for z in ref_in[1:]]
    parts_ref = [orc.multi_dice(u, onehot) for u in ups]
    total_ref = parts_ref[0] + w_aux * sum(parts_ref[1:])
    (total_ref * 1.3).backward()
    begin()
    dev_in = [z.to(DEV).requires_grad_(True) for z in zs]
    total, parts = O.DiceMultiFn.apply(*dev_in, O.labels_u8(lab.to(DEV), C), w_aux)
    (total * 1.3).backward()
    assert abs(float(total) - float(total_ref)) <= 1e-5 * abs(float(total_ref)), (float(total), float(total_ref))
    for a, b in zip(parts.tolist(), parts_ref):
        assert abs(a - float(b)) <= 1e-5 * abs(float(b))
    for k, (d, r) in enumerate(zip(dev_in, ref_in)):
        close(d.grad, r.grad, 2e-4, "dz%d" % k)


def test_torch_custom_ops_call_the_same_kernels():
    """torch.ops.tcct_b200.* (TORCH_LIBRARY shim) against the ctypes-bound path: same C ABI, same results bit for bit."""
    from tcct_b200 import torch_ops
    ns = torch_ops.load()
    g = gen(9)
    C, B, H, W = 5, 2, 64, 64
    zs = [(torch.randn(B, C, H // f, W // f, generator=g) * 2).to(DEV) for f in (1, 2, 4, 8)]
    lab = torch.randint(0, C, (B, H, W), generator=g).to(DEV).to(torch.uint8)
    begin()
    total, parts = O.DiceMultiFn.apply(*zs, lab, 0.5)
    loss, coef = ns.dice_multi_fwd(*zs, lab, 0.5)
    assert torch.equal(loss[4], total) and torch.equal(loss[:4], parts)
    d = ns.dice_multi_bwd(*zs, lab, 0.5, coef, torch.ones(1, device=DEV))
    assert d[0].shape == zs[0].shape and d[3].shape == zs[3].shape and bool(torch.isfinite(d[0]).all())
    from tcct_b200.kite.loop_seg import argmax_labels
    assert torch.equal(ns.argmax_labels(zs[0]), argmax_labels(zs[0]))
    x = torch.randn(2, 128, 128, 32, generator=g).to(DEV)
    mod = DenseConv(32, 32, 3).to(DEV)
    plan = PackPlan(mod, DEV)
    begin(); plan.run()
    before = ns.route_count(0)
    y = ns.conv2d_tma(x, mod.pk_tf, mod.bias, 3, 3, None, 0)
    assert ns.route_count(0) == before + 1
    with torch.no_grad():
        y2, _ = mod.run(x)
    assert torch.equal(y, y2)
    # the token mixer, the gate and the optimizer through the same registration
    t = torch.randn(2, 64, 96, generator=g).to(DEV)
    ln1, ln2 = torch.nn.LayerNorm(96, eps=1e-6).to(DEV), torch.nn.LayerNorm(96, eps=1e-6).to(DEV)
    t2, cur2, st = ns.ln_metapool_fwd(t, ln1.weight, ln1.bias, ln2.weight, ln2.bias, None, 1e-6)
    with torch.no_grad():
        r2, rc2 = O.LnMetaPoolFn.apply(t, ln1.weight, ln1.bias, ln2.weight, ln2.bias, None, 1e-6)
    assert torch.equal(t2, r2) and torch.equal(cur2, rc2)
    dg = [torch.zeros(96, device=DEV) for _ in range(4)]
    dt = ns.ln_metapool_bwd(t, t2, st, ln1.weight, ln2.weight, None, torch.ones_like(t), None, *dg)
    assert dt.shape == t.shape and bool(torch.isfinite(dt).all()) and float(dg[1].abs().sum()) >= 0
    a1, a2 = torch.randn(1, 8, 8, 32, generator=g).to(DEV), torch.randn(1, 8, 8, 32, generator=g).to(DEV)
    al = torch.rand(1, 32, 3, 3, generator=g).to(DEV)
    assert torch.equal(ns.gate_fuse_fwd(a1, a2, al), O.GateFuseFn.apply(a1, a2, al))
    assert torch.equal(ns.gate_fuse_fwd(a1, a2, None), 0.5 * a1 + 0.5 * a2)
    p0 = torch.randn(1000, generator=g).to(DEV)
    p, gr, m, v = p0.clone(), torch.randn(1000, generator=g).to(DEV), torch.zeros(1000, device=DEV), torch.zeros(1000, device=DEV)
    state = torch.tensor([0.0, 1e-3, 0.0, 0.0], device=DEV)
    ns.clip_adamw_step(p, gr, m, v, state, 12.0, 0.9, 0.999, 1e-8, 2e-4, 1.0)
    ref = torch.nn.Parameter(p0.clone()); ref.grad = gr.clone()
    torch.nn.utils.clip_grad_norm_([ref], 12.0)
    opt = torch.optim.AdamW([ref], lr=1e-3, weight_decay=2e-4); opt.step()
    assert float((p - ref.detach()).abs().max()) <= 1e-6 and abs(float(state[2]) - float(gr.norm())) <= 1e-3


@pytest.mark.parametrize("B,N,C", [(2, 200, 64), (1, 37, 96), (3, 16, 160), (8, 4096, 96), (2, 1, 128), (2, 50, 40), (1, 33, 64), (5, 3, 128),
                                   (1, 16384, 64)])
@pytest.mark.parametrize("scaled", [False, True])
def test_ln_metapool_fused(B, N, C, scaled):
    """ops.LnMetaPoolFn (csrc/ln_metapool.cu) against the reference ops in fp32 (MHCABlock.forward tcct.py:457-469: LayerNorm ->
    AvgPool2d(3,1,1,count_include_pad=False) on the 3-D [B,N,C] tensor -> residual with DropPath scale -> LayerNorm): both outputs and
    all gradients, with independent gradients flowing into t2 and cur2."""
    g = gen(51)
    t = torch.randn(B, N, C, generator=g)
    ws = [1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g), 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)]
    scale = (torch.rand(B, generator=g) < 0.7).float() / 0.7 if scaled else None
    d2, dc = torch.randn(B, N, C, generator=g), torch.randn(B, N, C, generator=g)
    tr = t.clone().requires_grad_(True)
    wr = [w.clone().requires_grad_(True) for w in ws]
    cur = F.layer_norm(tr, (C,), wr[0], wr[1], 1e-6)
    pooled = F.avg_pool2d(cur, 3, 1, 1, count_include_pad=False) - cur
    t2r = tr + (pooled * scale.view(-1, 1, 1) if scaled else pooled)
    c2r = F.layer_norm(t2r, (C,), wr[2], wr[3], 1e-6)
    ((t2r * d2).sum() + (c2r * dc).sum()).backward()
    begin()
    td = t.to(DEV).requires_grad_(True)
    wd = [torch.nn.Parameter(w.to(DEV)) for w in ws]
    attach(*wd)
    t2, c2 = O.LnMetaPoolFn.apply(td, wd[0], wd[1], wd[2], wd[3], scale.to(DEV) if scaled else None, 1e-6)
    ((t2 * d2.to(DEV)).sum() + (c2 * dc.to(DEV)).sum()).backward()
    close(t2, t2r, 2e-5, "t2"); close(c2, c2r, 2e-5, "cur2")
    close(td.grad, tr.grad, 1e-4, "dt")
    for k, (a, b) in enumerate(zip(wd, wr)):
        close(a.grad, b.grad, 2e-4, "dparam%d" % k)


@pytest.mark.parametrize("Cout,B,H,W", [(64, 2, 40, 128), (64, 8, 128, 128)])
def test_conv2d_tma_wide_output_slices(Cout, B, H, W):
    """32 -> 64 3x3 conv (MPViT stem, tcct.py:682-689) on the tcgen05 path as 32 -> 32 channel slices: forward with bias and
    statistics, data gradient through the TMA reduce-add store, weight gradient, against fp32 F.conv2d."""
    g = gen(27)
    mod = DenseConv(32, Cout, 3).to(DEV)
    with torch.no_grad():
        mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.1)
        mod.bias.copy_(torch.randn(Cout, generator=g))
    plan = PackPlan(mod, DEV)
    x = torch.randn(B, 32, H, W, generator=g)
    dy = torch.randn(B, Cout, H, W, generator=g)
    xr = x.clone().requires_grad_(True)
    wr, br = mod.weight.detach().cpu().requires_grad_(True), mod.bias.detach().cpu().requires_grad_(True)
    yr = F.conv2d(xr, wr, br, 1, 1)
    yr.backward(dy)
    begin(); plan.run()
    attach(mod.weight, mod.bias)
    import tcct_b200._lib as L
    before = L.route_counts()["conv_tma"]
    xg = nhwc(x).to(DEV).requires_grad_(True)
    y, stats = mod.run(xg, want_stats=True)
    y.backward(nhwc(dy).to(DEV))
    assert L.route_counts()["conv_tma"] - before == 2 * (Cout // 32)
    close(nchw(y), yr, TF32, "y"); close(nchw(xg.grad), xr.grad, TF32, "dx")
    close(mod.weight.grad, wr.grad, TF32, "dw"); close(mod.bias.grad, br.grad, TF32, "db")
    close(stats, torch.cat([yr.sum((0, 2, 3)), (yr * yr).sum((0, 2, 3))]).double(), TF32, "stats")


@pytest.mark.parametrize("B,C,H,W,train", [(2, 32, 32, 48, True), (1, 96, 16, 16, True), (2, 64, 8, 8, False), (1, 256, 4, 4, True)])
def test_gate_fusion(B, C, H, W, train):
    """GateFusion (tcct.py:916-932): the bicubic up-sampling of the random gate field evaluated in-kernel against F.interpolate."""
    g = gen(41)
    x1, x2, dy = (torch.randn(B, C, H, W, generator=g) for _ in range(3))
    alpha = torch.rand(B, C, max(3, H // 32), max(3, W // 32), generator=g) if train else None
    a1, a2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    full = F.interpolate(alpha, size=(H, W), mode="bicubic").clamp(0, 1) if train else 0.5
    ref = a1 * full + a2 * (1 - full)
    ref.backward(dy)
    g1, g2 = nhwc(x1).to(DEV).requires_grad_(True), nhwc(x2).to(DEV).requires_grad_(True)
    out = O.GateFuseFn.apply(g1, g2, alpha.to(DEV) if train else None)
    out.backward(nhwc(dy).to(DEV))
    close(nchw(out), ref, 1e-5, "out"); close(nchw(g1.grad), a1.grad, 1e-5, "d1"); close(nchw(g2.grad), a2.grad, 1e-5, "d2")


@pytest.mark.parametrize("B,N,dim,scaled", [(8, 1024, 128, True), (2, 16384, 64, False), (8, 4096, 96, True)])
def test_mlp_fused_gelu_epilogues(B, N, dim, scaled):
    """ops.MlpFn (fc1 + GELU in one tcgen05 launch, fc2 + residual; backward with GELU' in the fc2 data-gradient epilogue) against the
    unfused operators and fp32 PyTorch (Mlp tcct.py:29-53 inside MHCABlock.forward 467-468)."""
    from tcct_b200.nets.tcct import Mlp
    g = gen(61)
    mlp = Mlp(dim, dim).to(DEV)
    with torch.no_grad():
        for p in mlp.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.1 if p.dim() > 1 else 0.05))
    plan = PackPlan(mlp, DEV)
    cur, t, dy = (torch.randn(B, N, dim, generator=g) for _ in range(3))
    scale = (torch.rand(B, generator=g) < 0.7).float() / 0.7 if scaled else None
    cr, tr_ = cur.clone().requires_grad_(True), t.clone().requires_grad_(True)
    ps = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in mlp.named_parameters()}
    hid = F.gelu(F.linear(cr, ps["fc1.weight"], ps["fc1.bias"]))
    y = F.linear(hid, ps["fc2.weight"], ps["fc2.bias"])
    ref = tr_ + (y * scale.view(B, 1, 1) if scaled else y)
    ref.backward(dy)
    res = {}
    for fused in (True, False):
        begin(); plan.run()
        attach(*mlp.parameters())
        cg, tg = cur.to(DEV).requires_grad_(True), t.to(DEV).requires_grad_(True)
        sc = scale.to(DEV) if scaled else None
        if fused:
            assert O.mlp_fused_supported(B * N, dim, dim)
            out = O.MlpFn.apply(cg, tg, sc, mlp.fc1, mlp.fc2)
        else:
            h = O.bn_act2(mlp.fc1.run(cg), post=O.ACT_GELU, training=True)
            out = mlp.fc2.run(h, res=tg, res_scale=sc)
        out.backward(dy.to(DEV))
        torch.cuda.synchronize()
        res[fused] = [out.detach().clone(), cg.grad.clone(), tg.grad.clone()] + [p.grad.clone() for p in mlp.parameters()]
    names = ["out", "dcur", "dt"] + ["d" + k for k, _ in mlp.named_parameters()]
    refs = [ref, cr.grad, tr_.grad] + [ps[k].grad for k, _ in mlp.named_parameters()]
    for n, a, b, r in zip(names, res[True], res[False], refs):
        close(a, b, 1e-5, "fused vs unfused " + n)
        close(a, r, TF32, "fused vs fp32 " + n)
