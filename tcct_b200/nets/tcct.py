"""`stc_tt` -- the Tightly-Combined cross-Convolution and Transformer network -- on the B200 kernel path.

Drop-in for task1/nets/tcct.py of tyb311/TCCT: same constructor names (`stc_tt`, alias `tcct`), same
sub-module tree and therefore the same state_dict keys (tests/golden/state_keys.txt), same forward
contract (`[B,3,H,W]` fp32 image -> list of four `[B,C,H,W]` logit maps, side effect `self.feats`).
torch.nn classes are used only as parameter containers (key names, default initialisation); every
forward/backward computation runs in the kernels of csrc/ through tcct_b200.ops.

Per-stage structure follows the reference line by line in behaviour, not in code:
CrossCNNBlock tcct.py:803-828, CrossResNet 857-885, Conv2d_BN 55-97, DWConv2d_BN 99-147, ConvPosEnc 197-217,
MetaPool 405-415, MHCABlock 417-469, ResBlock 518-572, MHCA_stage 574-616, MPViT 649-753, mpvit_tiny 766-776,
MPUpBlock 887-914, norm_add 937-942, FTC 944-1047, stc_tt 1090-1096."""
import math

import torch
import torch.nn as nn

from .. import ops as O
from .flat import FlatModule

KSIZES = (13, 11, 9, 7, 5)
LN_EPS = 1e-6


# ----------------------------------------------------------------------------- parameter-holding leaves
class DenseConv(nn.Conv2d):
    """Dense conv executed by the tensor-core kernels: spatial (3x3 / 1xk / kx1) or 1x1."""

    def __init__(self, cin, cout, kernel_size=1, bias=True, k_slices=None):
        ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        super().__init__(cin, cout, ks, 1, (ks[0] // 2, ks[1] // 2), bias=bias)
        self.dense_kind = "spatial" if ks[0] * ks[1] > 1 else "gemm"
        self.k_slices = k_slices or [(0, cin)]
        self.pk_f = self.pk_b = self.pk_tf = self.pk_tb = None

    def run(self, x, want_stats=False, stats_act=O.ACT_NONE, res=None, res_scale=None, part=0):
        if self.dense_kind == "spatial":
            return O.Conv2dFn.apply(x, self.weight, self.bias, self.pk_f, self.pk_b, want_stats, stats_act, self.pk_tf, self.pk_tb)
        k0 = self.k_slices[part][0]
        bias = self.bias if part == len(self.k_slices) - 1 else None
        return O.GemmFn.apply(x, self.weight, bias, self.pk_f[part], self.pk_b[part], k0, res, res_scale,
                              want_stats, stats_act, self.pk_tf[part], self.pk_tb[part])


class DenseLinear(nn.Linear):
    def __init__(self, cin, cout):
        super().__init__(cin, cout)
        self.dense_kind = "gemm"
        self.k_slices = [(0, cin)]
        self.pk_f = self.pk_b = self.pk_tf = self.pk_tb = None

    def run(self, x, res=None, res_scale=None):
        return O.GemmFn.apply(x, self.weight, self.bias, self.pk_f[0], self.pk_b[0], 0, res, res_scale, False, 0,
                              self.pk_tf[0], self.pk_tb[0])[0]


class DwConv(nn.Conv2d):
    """Depthwise 3x3 (weights only; runs in csrc/pointwise.cu: dwconv3_*)."""

    def __init__(self, ch, stride=1, bias=False):
        super().__init__(ch, ch, 3, stride, 1, groups=ch, bias=bias)

    def run(self, x, add_input=False, want_stats=False):
        return O.DwConv3Fn.apply(x, self.weight, self.bias, self.stride[0], add_input, want_stats)


def _bn(x, stats, bn, training, pre=O.ACT_NONE, post=O.ACT_NONE):
    return O.bn_act2(x, stats, bn, pre, post=post, training=training)


# ----------------------------------------------------------------------------- MPViT-tiny branch
class Conv2d_BN(nn.Module):
    """conv (no bias) -> BN -> optional Hardswish."""

    def __init__(self, cin, cout, kernel_size=1, stride=1, act=False, stem=False, k_slices=None):
        super().__init__()
        if stem:    # 3 -> 32, stride 2: dedicated stem kernel
            self.conv = nn.Conv2d(cin, cout, kernel_size, stride, kernel_size // 2, bias=False)
        else:
            self.conv = DenseConv(cin, cout, kernel_size, bias=False, k_slices=k_slices)
        self.bn = nn.BatchNorm2d(cout)
        self.stem, self.stride, self.act = stem, stride, O.ACT_HSWISH if act else O.ACT_NONE
        fan_out = kernel_size * kernel_size * cout
        nn.init.normal_(self.conv.weight, 0.0, math.sqrt(2.0 / fan_out))

    def forward(self, x, x2=None, res=None):
        if self.stem:
            y, st = O.StemConvFn.apply(x, self.conv.weight, None, self.stride, True)
        elif x2 is not None:    # 1x1 conv over the channel concat [x, x2] without materialising it
            y, _ = self.conv.run(x, part=0)
            y, st = self.conv.run(x2, want_stats=True, res=y, part=1)
        else:
            y, st = self.conv.run(x, want_stats=True)
        if res is not None:     # res + BN(conv(x))
            return O.bn_act2(y, st, self.bn, b=res, post=self.act, training=self.training)
        return _bn(y, st, self.bn, self.training, post=self.act)


class DWConv2d_BN(nn.Module):
    """depthwise 3x3 (stride s) -> pointwise 1x1 -> BN -> Hardswish."""

    def __init__(self, ch, stride):
        super().__init__()
        self.dwconv = DwConv(ch, stride, bias=False)
        self.pwconv = DenseConv(ch, ch, 1, bias=False)
        self.bn = nn.BatchNorm2d(ch)
        for m in (self.dwconv, self.pwconv):
            n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / n))

    def forward(self, x):
        d, _ = self.dwconv.run(x)
        y, st = self.pwconv.run(d, want_stats=True)
        return _bn(y, st, self.bn, self.training, post=O.ACT_HSWISH)


class DWCPatchEmbed(nn.Module):
    def __init__(self, ch, stride):
        super().__init__()
        self.patch_conv = DWConv2d_BN(ch, stride)

    def forward(self, x):
        return self.patch_conv(x)


class Patch_Embed_stage(nn.Module):
    def __init__(self, ch, isPool):
        super().__init__()
        self.patch_embeds = nn.ModuleList([DWCPatchEmbed(ch, 2 if isPool else 1)])

    def forward(self, x):
        return self.patch_embeds[0](x)


class ConvPosEnc(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = DwConv(dim, 1, bias=True)

    def forward(self, x):
        return self.proj.run(x, add_input=True)[0]


class ConvRelPosEnc(nn.Module):
    """Parameters of the (disabled) factorised-attention relative position encoding; kept for
    checkpoint compatibility only -- the reference never executes it (tcct.py:435-449)."""

    def __init__(self, Ch, h, window):
        super().__init__()
        self.conv_list = nn.ModuleList()
        for ksize, split in window.items():
            self.conv_list.append(nn.Conv2d(split * Ch, split * Ch, ksize, padding=ksize // 2, groups=split * Ch))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = DenseLinear(dim, hidden)
        self.fc2 = DenseLinear(hidden, dim)


class MHCABlock(nn.Module):
    """cpe -> LN -> MetaPool (+res, DropPath) -> LN -> fc1 -> GELU -> fc2 (+res, DropPath)."""
    dp_tape = None      # tests: list of per-sample {0,1} keep masks consumed in call order

    def __init__(self, dim, mlp_ratio, drop_path, shared_cpe, shared_crpe):
        super().__init__()
        self.cpe, self.crpe = shared_cpe, shared_crpe
        self.mlp = Mlp(dim, dim * mlp_ratio)
        self.drop_rate = float(drop_path)
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)

    def _dp_scale(self, batch, device):
        if self.drop_rate == 0.0 or not self.training:
            return None
        keep = 1.0 - self.drop_rate
        if MHCABlock.dp_tape is not None:
            mask = MHCABlock.dp_tape.pop(0).to(device=device, dtype=torch.float32)
        else:
            mask = torch.empty(batch, dtype=torch.float32, device=device).bernoulli_(keep)
        return (mask / keep).contiguous()

    def forward(self, x):
        t = self.cpe(x)                                           # [B,h,w,C] == tokens [B,N,C]
        B = t.shape[0]
        t, cur = O.LnMetaPoolFn.apply(t, self.norm1.weight, self.norm1.bias, self.norm2.weight, self.norm2.bias,
                                      self._dp_scale(B, t.device), LN_EPS)
        fc1, fc2 = self.mlp.fc1, self.mlp.fc2
        scale = self._dp_scale(B, t.device)
        if (torch.is_grad_enabled() and fc1.pk_tf and fc1.weight.requires_grad and hasattr(fc1.weight, "_gview")
                and O.mlp_fused_supported(cur.numel() // cur.shape[-1], cur.shape[-1], fc1.weight.shape[0])):
            return O.MlpFn.apply(cur, t, scale, fc1, fc2)
        hidden = fc1.run(cur)
        hidden = O.bn_act2(hidden, post=O.ACT_GELU, training=self.training)
        return fc2.run(hidden, res=t, res_scale=scale)


class MHCAEncoder(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, drop_path):
        super().__init__()
        self.cpe = ConvPosEnc(dim)
        self.crpe = ConvRelPosEnc(dim // num_heads, num_heads, {3: 2, 5: 3, 7: 3})
        self.MHCA_layers = nn.ModuleList([MHCABlock(dim, mlp_ratio, drop_path, self.cpe, self.crpe)])

    def forward(self, x):
        for layer in self.MHCA_layers:
            x = layer(x)
        return x


class ResBlock(nn.Module):
    """x + BN(1x1( Hswish(BN(dw3x3( Hswish(BN(1x1(x))) ))) ))."""

    def __init__(self, dim):
        super().__init__()
        self.conv1 = Conv2d_BN(dim, dim, act=True)
        self.dwconv = DwConv(dim, 1, bias=False)
        self.norm = nn.BatchNorm2d(dim)
        self.conv2 = Conv2d_BN(dim, dim)
        nn.init.normal_(self.dwconv.weight, 0.0, math.sqrt(2.0 / (9 * dim)))

    def forward(self, x):
        r = self.conv1(x)
        d, st = self.dwconv.run(r, want_stats=True)
        r = _bn(d, st, self.norm, self.training, post=O.ACT_HSWISH)
        return self.conv2(r, res=x)


class MHCA_stage(nn.Module):
    def __init__(self, dim, out_dim, num_heads, mlp_ratio, drop_path):
        super().__init__()
        self.mhca_blks = nn.ModuleList([MHCAEncoder(dim, num_heads, mlp_ratio, drop_path)])
        self.InvRes = ResBlock(dim)
        self.aggregate = Conv2d_BN(dim * 2, out_dim, act=True, k_slices=[(0, dim), (dim, dim)])

    def forward(self, x):
        # the inverted-residual and the token-mixer branches are independent until `aggregate`: the MPViT encoder is the
        # longest dependent chain of the step (mostly latency-bound kernels on small maps), so they run side by side
        side = O.fork(x.device, 1)
        with O.on(side):
            if side is not None:
                x.record_stream(side)
            r = self.InvRes(x)
        t = self.mhca_blks[0](x)
        O.join(side, r)
        return self.aggregate(r, x2=t)


class Cls_head(nn.Module):
    """Unused ImageNet classifier head of MPViT (kept for checkpoint compatibility)."""

    def __init__(self, dim, num_classes):
        super().__init__()
        self.cls = nn.Linear(dim, num_classes)


class MPViT(nn.Module):
    def __init__(self, embed_dims=(64, 96, 128, 160), num_heads=(4, 4, 4, 4), mlp_ratios=(1, 1, 1, 1),
                 drop_path_rate=0.1, num_classes=1000):
        super().__init__()
        self.embed_dims = list(embed_dims)
        n = len(embed_dims)
        dpr = [drop_path_rate * i / (n - 1) for i in range(n)]      # linspace(0, rate, stages), one layer each
        self.stem = nn.Sequential(Conv2d_BN(3, embed_dims[0] // 2, 3, 2, act=True, stem=True),
                                  Conv2d_BN(embed_dims[0] // 2, embed_dims[0], 3, 1, act=True))
        self.patch_embed_stages = nn.ModuleList([Patch_Embed_stage(embed_dims[i], isPool=i > 0) for i in range(n)])
        self.mhca_stages = nn.ModuleList([
            MHCA_stage(embed_dims[i], embed_dims[i + 1] if i + 1 < n else embed_dims[i], num_heads[i], mlp_ratios[i], dpr[i])
            for i in range(n)])
        self.cls_head = Cls_head(embed_dims[-1], num_classes)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def forward_features(self, img):
        x = self.stem[1](self.stem[0](img))
        outs = []
        for pe, stage in zip(self.patch_embed_stages, self.mhca_stages):
            x = stage(pe(x))
            outs.append(x)
        return outs


def mpvit_tiny(**kw):
    return MPViT((64, 96, 128, 160), (4, 4, 4, 4), (1, 1, 1, 1), **kw)


# ----------------------------------------------------------------------------- cross-convolution branch
class CrossCNNBlock(nn.Module):
    """a = BN(lrelu(3x3(3x3 x)));  b = BN(lrelu(3x3(kx1(1xk x))));  out = BN(lrelu(3x3(GELU(a+b))))."""

    def __init__(self, in_c, out_c, ksize):
        super().__init__()
        self.block12 = nn.Sequential(DenseConv(in_c, out_c, 3), DenseConv(out_c, out_c, 3), nn.LeakyReLU(), nn.BatchNorm2d(out_c))
        self.block34 = nn.Sequential(DenseConv(in_c, out_c, (1, ksize)), DenseConv(out_c, out_c, (ksize, 1)),
                                     DenseConv(out_c, out_c, 3), nn.LeakyReLU(), nn.BatchNorm2d(out_c))
        self.block5 = nn.Sequential(DenseConv(out_c, out_c, 3), nn.LeakyReLU(), nn.BatchNorm2d(out_c))

    def forward(self, x):
        tr = self.training
        a, _ = self.block12[0].run(x)
        a, sa = self.block12[1].run(a, want_stats=True, stats_act=O.ACT_LRELU)
        b, _ = self.block34[0].run(x)
        b, _ = self.block34[1].run(b)
        b, sb = self.block34[2].run(b, want_stats=True, stats_act=O.ACT_LRELU)
        g = O.bn_act2(a, sa, self.block12[3], O.ACT_LRELU, b, sb, self.block34[4], O.ACT_LRELU, O.ACT_GELU, tr)
        o, so = self.block5[0].run(g, want_stats=True, stats_act=O.ACT_LRELU)
        return O.bn_act2(o, so, self.block5[2], O.ACT_LRELU, training=tr)


class PlainCNNBlock(CrossCNNBlock):
    """tcct.py:830-855 (`pnnu`): the same block with the cross kernels forced to 3 (1x3, 3x1)."""

    def __init__(self, in_c, out_c, ksize):
        super().__init__(in_c, out_c, 3)


class CrossResNet(nn.Module):
    __name__ = "crnet"

    def __init__(self, in_ch=3, out_ch=6, flag_tiny=False, Block=CrossCNNBlock):
        super().__init__()
        # tcct.py:861-864: the tiny branch keeps 32 channels (tcgen05 kernels on the large maps); the wide one grows to 256 and runs
        # its reductions as 64 / 32-channel slices of the warp-level kernels (ops.reduction_slices)
        layers = (32, 32, 32, 32, 32) if flag_tiny else (32, 64, 96, 128, 256)
        self.layer_dims = layers
        self.pool = nn.MaxPool2d(2)
        self.path_estan = nn.ModuleList([Block(layers[max(i - 1, 0)], layers[i], k) for i, k in enumerate(KSIZES)])
        self.cnn = nn.Sequential(nn.Conv2d(3, 32, 3, 1, 1), nn.BatchNorm2d(32))

    def forward(self, img):
        y, st = O.StemConvFn.apply(img, self.cnn[0].weight, self.cnn[0].bias, 1, True)
        x = _bn(y, st, self.cnn[1], self.training)
        outs = []
        for i, blk in enumerate(self.path_estan):
            x = blk(x)
            outs.append(x)
            if i + 1 < len(self.path_estan):
                x = O.MaxPool2Fn.apply(x)
        return outs


# ----------------------------------------------------------------------------- decoder
class MPUpBlock(nn.Module):
    """1x1( up2x_bilinear_align_corners( lrelu(BN(3x3(x))) ) + skip )."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.prep = nn.Sequential(DenseConv(in_ch, out_ch, 3), nn.BatchNorm2d(out_ch), nn.LeakyReLU(inplace=True))
        self.post = nn.Sequential(DenseConv(out_ch, out_ch, 1))

    def forward(self, x, skip):
        y, st = self.prep[0].run(x, want_stats=True)
        y = _bn(y, st, self.prep[1], self.training, post=O.ACT_LRELU)
        B, h, w, _ = y.shape
        y = O.ResizeNHWCFn.apply(y, skip, 2 * h, 2 * w, True, 1.0)
        return self.post[0].run(y)[0]


def norm_add(xs):
    """mean_k bilinear_up( L2normalize_C(x_k) ) at the resolution of xs[0] (align_corners=False)."""
    if len(xs) == 3:        # the stc_tt case: one fused pass over the full-resolution map
        n1, n2 = O.L2Norm32Fn.apply(xs[1]), O.L2Norm32Fn.apply(xs[2])
        return [O.NormAdd3Fn.apply(xs[0], n1, n2, 1.0 / 3.0)]
    ns = [O.L2Norm32Fn.apply(x) for x in xs]
    H, W = ns[0].shape[1:3]
    total = None
    for n in ns:
        total = O.ResizeNHWCFn.apply(n, total, H, W, False, 1.0 / len(ns))
    return [total]


class GateFusion(nn.Module):
    """tcct.py:916-932: train mode mixes the two branches with a random smooth field per call -- torch.rand(B, C, max(3, H/32),
    max(3, W/32)), bicubic up-sampling, clamp -- eval mode with 0.5.  The field is drawn on the device (the reference draws and
    up-samples it on the host and copies the full-size tensor over); `alpha_tape` injects fields for parity tests."""
    alpha_tape = None       # [B, C, hs, ws] tensors indexed by fusion scale (the reference's call order: x2, x3, x4, x5)

    def __init__(self):
        super().__init__()
        self.relu = nn.LeakyReLU(inplace=True)        # registered like the reference's (unused there too)

    def forward(self, x1, x2, index=0):
        alpha = None
        if self.training:
            B, H, W, C = x1.shape
            if GateFusion.alpha_tape is not None:
                alpha = GateFusion.alpha_tape[index].to(x1.device).contiguous()
            else:
                alpha = torch.rand((B, C, max(3, H // 32), max(3, W // 32)), device=x1.device)
        return O.GateFuseFn.apply(x1, x2, alpha)


class FTC(FlatModule):
    __name__ = "gtc"

    def __init__(self, base_cnn, base_vit, out_channels=5, filters=32, flag_gate=True, flag_cnn=True, flag_vit=True, variant="tcct",
                 **args):
        """variant: "tcct" = nets/tcct.py:944-1047; "onnx" = the older decoder of onnx/tcct_{goals,hcms,heg}.py:949-1035 that the
        shipped tcct_goals / tcct_hcms / tcct_heg checkpoints were trained with (no t321-t324 projections, the auxiliary heads read
        the decoder maps directly, feats = norm_add([x1, x2, x3, y0, y1, y2]))."""
        super().__init__()
        if filters != 32 or not (flag_cnn or flag_vit) or variant not in ("tcct", "onnx"):
            raise NotImplementedError("tcct_b200: FTC is built for 32 decoder filters and the tcct / onnx decoder variants")
        self.flag_cnn, self.flag_vit, self.variant = flag_cnn, flag_vit, variant
        self.base_vit, self.base_cnn = base_vit, base_cnn
        if not flag_vit:        # cnnu / pnnu (tcct.py:955-957): the MPViT branch is frozen and its features are not used
            for p in base_vit.parameters():
                p.requires_grad = False
            # the fusion convs never run either: like in torch, where their .grad stays None, the optimizer must not touch them
            self.UNUSED = FlatModule.UNUSED + ("tran_vit", "tran_cnn")
        if not flag_cnn:        # vitu (tcct.py:960-962): the CrossResNet branch is frozen; only its first block output (x1 = c1) is used
            for p in base_cnn.parameters():
                p.requires_grad = False
            self.UNUSED = FlatModule.UNUSED + ("tran_cnn",)
        ed, ld = base_vit.embed_dims, base_cnn.layer_dims
        print('DIMS-VIT:', ed)
        print('DIMS-CNN:', ld)
        print('CHES-NET:', out_channels)
        vit_in = (ed[1], ed[2], ed[3], ed[3])
        for i in range(4):
            setattr(self, "tran_vit%d" % i, nn.Sequential(DenseConv(vit_in[i], ld[i + 1], 1), nn.BatchNorm2d(ld[i + 1])))
        for i in range(4):
            setattr(self, "tran_cnn%d" % i, nn.Sequential(DenseConv(ld[i + 1], ld[i + 1], 1), nn.BatchNorm2d(ld[i + 1])))
        self.gate = GateFusion() if flag_gate else None       # SimpleFusion (x1 + x2) is folded into the BatchNorm pass of `_tran`
        self.head = nn.Sequential(DenseConv(ld[-1], ld[-1], 3), nn.BatchNorm2d(ld[-1]), nn.LeakyReLU())
        self.fuse = nn.Conv2d(ld[4], filters, kernel_size=1)         # registered, never executed (tcct.py:978)
        self.dec1, self.dec2 = MPUpBlock(ld[-1], ld[-2]), MPUpBlock(ld[-2], ld[-3])
        self.dec3, self.dec4 = MPUpBlock(ld[-3], ld[-4]), MPUpBlock(ld[-4], filters)
        if variant == "tcct":
            self.t321, self.t322 = DenseConv(ld[-2], filters, 1), DenseConv(ld[-3], filters, 1)
            self.t323, self.t324 = DenseConv(ld[-4], filters, 1), DenseConv(filters, filters, 1)
        for n in ("aux0", "aux1", "aux2", "aux4"):
            setattr(self, n, nn.Conv2d(filters, out_channels, kernel_size=1))
        self.feats = None
        self.defer_aux = False      # internal protocol with KiteSeg / KiteBack.grad_calc: return the auxiliary logits at native resolution

    def _aux_logits(self, y, aux, H, W):
        """Auxiliary head: 1x1 conv to class logits, up-sampled to the input size (tcct.py:1042-1044) -- or left at its native
        resolution when `defer_aux` is set (training loop: the deep-supervision Dice kernel up-samples in registers)."""
        z = O.HeadFn.apply(y, aux.weight, aux.bias)
        return z if self.defer_aux else O.ResizeNCHWFn.apply(z, H, W)

    def _tran(self, i, v, c):
        tv, tc = getattr(self, "tran_vit%d" % i), getattr(self, "tran_cnn%d" % i)
        yv, sv = tv[0].run(v, want_stats=True)
        yc, sc = tc[0].run(c, want_stats=True)
        if self.gate is not None:       # gtc_*: gate(BN(vit), BN(cnn)), tcct.py:1009-1012
            return self.gate(O.bn_act2(yv, sv, tv[1], training=self.training), O.bn_act2(yc, sc, tc[1], training=self.training), i)
        return O.bn_act2(yv, sv, tv[1], b=yc, stats_b=sc, bn_b=tc[1], training=self.training)

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("tcct_b200 runs on CUDA tensors only (no CPU fallback); got a %s tensor" % x.device)
        self.begin_step(x.device)
        return self.forward_impl(x)

    def forward_impl(self, x):
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] % 16 or x.shape[3] % 16 or min(x.shape[2:]) < 32:
            raise RuntimeError("stc_tt expects [B,3,H,W] with H, W multiples of 16, at least 32, got %s" % (tuple(x.shape),))
        x = x.contiguous().float()
        H, W = x.shape[2:]
        # The two encoders are independent until the fusion convs: the MPViT branch is issued on a side stream so that
        # its many small kernels overlap the CrossResNet ones (autograd replays each backward node on the stream of
        # its forward, so the backward passes overlap the same way; under CUDA-graph capture this forks the graph).
        side = O.fork(x.device, 0)
        with O.on(side):
            if side is not None:
                x.record_stream(side)
            if self.flag_vit:
                v2, v3, v4, v5 = self.base_vit.forward_features(x)
            elif self.training:
                # cnnu: the reference still runs the frozen branch (tcct.py:1003); its only effect is on the BatchNorm running
                # statistics of a train-mode forward, which stay in step with it
                with torch.no_grad():
                    self.base_vit.forward_features(x)
        if self.flag_cnn:
            c1, c2, c3, c4, c5 = self.base_cnn(x)
        else:
            with torch.no_grad():       # vitu: frozen branch, still executed (its c1 is the decoder's last skip, tcct.py:1019)
                c1, c2, c3, c4, c5 = self.base_cnn(x)
        if not self.flag_vit:
            O.join(side)
            return self._decode(x, c1, c2, c3, c4, c5, None)
        O.join(side, v2, v3, v4, v5)
        return self._decode(x, c1, c2, c3, c4, c5, (v2, v3, v4, v5))

    def _decode(self, x, c1, c2, c3, c4, c5, vit):
        H, W = x.shape[2:]
        x1 = c1
        dev = x.device
        tr = self.training

        def mark(st, *ts):            # tensors made on the current stream and read on `st`
            if st is not None:
                for t in ts:
                    t.record_stream(st)

        # The decoder is one dependent chain (head -> dec1 -> ... -> dec4) of mostly small kernels with nothing else in flight;
        # everything that hangs off it sideways runs on side streams: the two fusion blocks the chain needs last, and per scale
        # the 1x1 projection, the auxiliary head with its logit up-sampling and the normalisation for `norm_add`.
        if vit is None:         # cnnu / pnnu: the CrossResNet features feed the decoder directly (tcct.py:1017-1018)
            s_tr = None
            x2, x3, x4, x5 = c2, c3, c4, c5
        elif not self.flag_cnn:     # vitu: the projected MPViT features alone (tcct.py:1019-1020)
            s_tr = None

            def proj(i, v):
                tv = getattr(self, "tran_vit%d" % i)
                yv, sv = tv[0].run(v, want_stats=True)
                return O.bn_act2(yv, sv, tv[1], training=self.training)
            x2, x3, x4, x5 = (proj(i, v) for i, v in enumerate(vit))
        else:
            v2, v3, v4, v5 = vit
            s_tr = O.fork(dev, 5)
            with O.on(s_tr):
                mark(s_tr, v2, c2, v3, c3)
                x2, x3 = self._tran(0, v2, c2), self._tran(1, v3, c3)
            x5, x4 = self._tran(3, v5, c5), self._tran(2, v4, c4)
        y, st = self.head[0].run(x5, want_stats=True)
        y8 = _bn(y, st, self.head[1], self.training, post=O.ACT_LRELU)
        y4 = self.dec1(y8, x4)
        if self.variant == "onnx":
            # onnx/tcct_goals.py:1016-1035: no t32x projections; the auxiliary heads read the decoder maps directly,
            # feats = norm_add([x1, x2, x3, y0, y1, y2])
            O.join(s_tr, x2, x3)
            y2 = self.dec2(y4, x3)
            y1 = self.dec3(y2, x2)
            y0 = self.dec4(y1, x1)
            self.feats_nhwc = norm_add([x1, x2, x3, y0, y1, y2])[0]
            self.feats = [self.feats_nhwc.permute(0, 3, 1, 2)]
            o0 = O.HeadFn.apply(y0, self.aux0.weight, self.aux0.bias)
            return [o0, self._aux_logits(y1, self.aux1, H, W), self._aux_logits(y2, self.aux2, H, W),
                    self._aux_logits(y4, self.aux4, H, W)]
        s_aux = O.fork(dev, 4)
        with O.on(s_aux):
            mark(s_aux, x4, y4)
            y4a = self.t321.run(O.bn_act2(x4, b=y4, training=tr))[0]
            o4 = self._aux_logits(y4a, self.aux4, H, W)
        O.join(s_tr, x2, x3)
        y2 = self.dec2(y4, x3)
        s_aux = O.fork(dev, 4)
        with O.on(s_aux):
            mark(s_aux, x3, y2)
            y2a = self.t322.run(O.bn_act2(x3, b=y2, training=tr))[0]
            o2 = self._aux_logits(y2a, self.aux2, H, W)
            n2 = O.L2Norm32Fn.apply(y2a)
        y1 = self.dec3(y2, x2)
        s_aux = O.fork(dev, 4)
        with O.on(s_aux):
            mark(s_aux, x2, y1)
            y1a = self.t323.run(O.bn_act2(x2, b=y1, training=tr))[0]
            o1 = self._aux_logits(y1a, self.aux1, H, W)
            n1 = O.L2Norm32Fn.apply(y1a)
        y0 = self.dec4(y1, x1)
        y0 = self.t324.run(O.bn_act2(x1, b=y0, training=tr))[0]
        O.join(s_aux, o4, o2, o1, n1, n2)
        self.feats_nhwc = O.NormAdd3Fn.apply(y0, n1, n2, 1.0 / 3.0)   # norm_add([y0, y1, y2]); consumed by RegNet.regular_udh
        self.feats = [self.feats_nhwc.permute(0, 3, 1, 2)]             # reference layout [B,32,H,W] (a view)
        o0 = O.HeadFn.apply(y0, self.aux0.weight, self.aux0.bias)
        return [o0, o1, o2, o4]


def stc_tt(n_class=8, **args):
    net = FTC(base_vit=mpvit_tiny(), base_cnn=CrossResNet(flag_tiny=True), flag_gate=False, out_channels=n_class)
    net.__name__ = 'stctt'
    return net


tcct = stc_tt


def cnnu(n_class=8, **args):
    """tcct.py:1124-1129: the CrossResNet encoder + decoder alone (flag_vit=False); same module tree and state-dict keys as stc_tt."""
    net = FTC(base_vit=mpvit_tiny(), base_cnn=CrossResNet(flag_tiny=True), flag_gate=False, flag_vit=False, flag_cnn=True,
              out_channels=n_class)
    net.__name__ = 'cnnu'
    return net


def pnnu(n_class=8, **args):
    """tcct.py:1117-1122: `cnnu` with PlainCNNBlock (cross kernels 1x3 / 3x1)."""
    net = FTC(base_vit=mpvit_tiny(), base_cnn=CrossResNet(flag_tiny=True, Block=PlainCNNBlock), flag_gate=False, flag_vit=False,
              flag_cnn=True, out_channels=n_class)
    net.__name__ = 'pnnu'
    return net


def vitu(n_class=8, **args):
    """tcct.py:1131-1136: the MPViT encoder + decoder (flag_cnn=False: the CrossResNet branch is frozen, only x1 = c1 is used)."""
    net = FTC(base_vit=mpvit_tiny(), base_cnn=CrossResNet(flag_tiny=True), flag_gate=False, flag_vit=True, flag_cnn=False,
              out_channels=n_class)
    net.__name__ = 'vitu'
    return net


def stc_tt_onnx(n_class=8, **args):
    """`stc_tt` of onnx/tcct_{goals,hcms,heg}.py (the model definition the shipped tcct_goals / tcct_hcms / tcct_heg checkpoints load
    into; onnx/tcct_goals.py:1090-1095): same encoders, older decoder tail."""
    net = FTC(base_vit=mpvit_tiny(), base_cnn=CrossResNet(flag_tiny=True), flag_gate=False, out_channels=n_class, variant="onnx")
    net.__name__ = 'stctt'
    return net


def gtc_tt(n_class=8, **args):
    """tcct.py:1050-1055: stc_tt with GateFusion."""
    net = FTC(base_vit=mpvit_tiny(), base_cnn=CrossResNet(flag_tiny=True), flag_gate=True, out_channels=n_class)
    net.__name__ = 'gtctt'
    return net


def gtc_tb(n_class=8, **args):
    """tcct.py:1056-1061: GateFusion, wide CrossResNet (32, 64, 96, 128, 256)."""
    net = FTC(base_vit=mpvit_tiny(), base_cnn=CrossResNet(flag_tiny=False), flag_gate=True, out_channels=n_class)
    net.__name__ = 'gtctb'
    return net


def stc_tb(n_class=8, **args):
    """tcct.py:1097-1102: SimpleFusion, wide CrossResNet."""
    net = FTC(base_vit=mpvit_tiny(), base_cnn=CrossResNet(flag_tiny=False), flag_gate=False, out_channels=n_class)
    net.__name__ = 'stctb'
    return net


def _needs_mpvit_small(name):
    def factory(n_class=8, **args):
        raise NotImplementedError("tcct_b200: `%s` needs mpvit_small (tcct.py:778-788: 2-3 parallel paths per stage, up to 6 layers, "
                                  "216 / 288-channel embeddings that are not multiples of 32), which the B200 kernels do not take; built: "
                                  "stc_tt, stc_tb, gtc_tt, gtc_tb, cnnu, pnnu, vitu" % name)
    factory.__name__ = name
    return factory


stc_st, stc_sb = (_needs_mpvit_small(n) for n in ("stc_st", "stc_sb"))
