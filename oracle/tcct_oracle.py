"""CPU oracle for the TCCT `stc_tt` hot path -- TEST INFRASTRUCTURE ONLY.

A functional fp32 PyTorch restatement of the reference algorithm (tyb311/TCCT,
task1/).  It is the checker for tests/, __graft_entry__.smoke() and the
cpu_baseline / `--impl reference` legs of bench.py; nothing in tcct_b200/ may
import it.  Parity status: PINNED -- oracle/make_golden.py runs the unmodified
reference (imported from /root/reference/task1 through oracle/refshim.py) on
seeded inputs with injected noise and stores its outputs under tests/golden/;
tests/test_oracle_golden.py checks this file against those vectors.

Every function cites the reference lines it follows.  Parameters are a flat
{state_dict key: tensor} mapping using the reference's own key names
(RegNet(stc_tt(C)).state_dict(): `base.*`, `lap_reg.*`, `lap_map.*`, `fcp.*`).
"""
import math
import torch
import torch.nn.functional as F

BN_EPS, BN_MOM, LN_EPS = 1e-5, 0.1, 1e-6
KSIZES = (13, 11, 9, 7, 5)            # nets/tcct.py:866
VIT_DIMS = (64, 96, 128, 160)         # nets/tcct.py:766-776 (mpvit_tiny)
DROP_PATH = (0.0, 0.1 / 3, 0.2 / 3, 0.1)   # dpr_generator, nets/tcct.py:635-647


class Ctx:
    """Per-call options: train/eval, DropPath masks, BN running-stat updates."""

    def __init__(self, training=True, dp_masks=None, update_stats=True, gate_alphas=None):
        self.training = training
        self.dp_masks = list(dp_masks) if dp_masks is not None else None
        self.update_stats = update_stats
        self.gate_alphas = list(gate_alphas) if gate_alphas is not None else None     # GateFusion's torch.rand fields, call order


def _bn(P, key, x, ctx, eps=BN_EPS):
    """nn.BatchNorm2d (train: biased batch var for normalisation, unbiased for
    the running estimate, momentum 0.1)."""
    w, b = P[key + ".weight"], P[key + ".bias"]
    rm, rv = P[key + ".running_mean"], P[key + ".running_var"]
    if not ctx.training:
        return F.batch_norm(x, rm, rv, w, b, False, BN_MOM, eps)
    if ctx.update_stats:
        with torch.no_grad():
            n = x.numel() / x.shape[1]
            mean = x.mean((0, 2, 3))
            var = x.var((0, 2, 3), unbiased=False)
            rm.mul_(1 - BN_MOM).add_(BN_MOM * mean)
            rv.mul_(1 - BN_MOM).add_(BN_MOM * var * n / max(n - 1, 1))
            P[key + ".num_batches_tracked"] += 1
    return F.batch_norm(x, None, None, w, b, True, BN_MOM, eps)


QUANT = None    # tests only: TF32-emulation hook {inp(t): operand rounding, out(t): identity whose backward rounds dy};
                # None = exact fp32 (the oracle proper).  Used to CALIBRATE gradient tolerances of the tensor-core path.


def _dense(fn, x, w, *args):
    """A dense contraction (conv with groups=1, linear): the ops the GPU path runs on tensor cores."""
    if QUANT is None:
        return fn(x, w, *args)
    return QUANT.out(fn(QUANT.inp(x), QUANT.inp(w), *args))


def _conv(P, key, x, stride=1, pad=0, groups=1):
    if groups == 1:
        return _dense(F.conv2d, x, P[key + ".weight"], P.get(key + ".bias"), stride, pad, 1, groups)
    return F.conv2d(x, P[key + ".weight"], P.get(key + ".bias"), stride, pad, 1, groups)


def cross_block(P, key, x, k, ctx):
    """CrossCNNBlock.forward, nets/tcct.py:803-828."""
    a = _conv(P, key + ".block12.1", _conv(P, key + ".block12.0", x, pad=1), pad=1)
    a = _bn(P, key + ".block12.3", F.leaky_relu(a, 0.01), ctx)
    b = _conv(P, key + ".block34.0", x, pad=(0, k // 2))
    b = _conv(P, key + ".block34.1", b, pad=(k // 2, 0))
    b = _conv(P, key + ".block34.2", b, pad=1)
    b = _bn(P, key + ".block34.4", F.leaky_relu(b, 0.01), ctx)
    g = F.gelu(a + b)
    return _bn(P, key + ".block5.2", F.leaky_relu(_conv(P, key + ".block5.0", g, pad=1), 0.01), ctx)


def cross_resnet(P, key, x, ctx, plain=False):
    """CrossResNet.forward, nets/tcct.py:877-885 (flag_tiny: all 32 channels).  plain: Block=PlainCNNBlock (tcct.py:830-855, the
    `pnnu` factory), i.e. the cross kernels are 1x3 / 3x1 whatever the stage."""
    x = _bn(P, key + ".cnn.1", _conv(P, key + ".cnn.0", x, pad=1), ctx)
    feats = []
    for i, k in enumerate(KSIZES):
        x = cross_block(P, "%s.path_estan.%d" % (key, i), x, 3 if plain else k, ctx)
        feats.append(x)
        x = F.max_pool2d(x, 2)
    return feats


def _conv_bn(P, key, x, ctx, act=None, stride=1, pad=0):
    """Conv2d_BN, nets/tcct.py:55-97."""
    x = _bn(P, key + ".bn", _conv(P, key + ".conv", x, stride, pad), ctx)
    return F.hardswish(x) if act == "hswish" else x


def _drop_path(x, rate, ctx):
    """timm DropPath (scale_by_keep): per-sample Bernoulli(1-p) mask / (1-p)."""
    if rate == 0.0 or not ctx.training:
        return x
    if ctx.dp_masks is None:
        return x
    m = ctx.dp_masks.pop(0).to(x.dtype).view(-1, *([1] * (x.ndim - 1)))
    return x * (m / (1.0 - rate))


def meta_pool(t):
    """MetaPool.forward on a 3-D [B,N,C] tensor, nets/tcct.py:405-415: AvgPool2d
    treats B as channels and pools 3x3 over the (token, channel) plane."""
    return F.avg_pool2d(t, 3, 1, 1, count_include_pad=False) - t


def mhca_stage(P, key, x, dim, rate, ctx):
    """Patch_Embed_stage output -> MHCA_stage.forward, nets/tcct.py:604-616,
    with MHCABlock.forward 457-469, ConvPosEnc 208-217, Mlp 46-53, ResBlock 562-571."""
    B, C, H, W = x.shape
    r = _conv_bn(P, key + ".InvRes.conv1", x, ctx, "hswish")
    r = F.conv2d(r, P[key + ".InvRes.dwconv.weight"], None, 1, 1, 1, C)
    r = F.hardswish(_bn(P, key + ".InvRes.norm", r, ctx))
    r = x + _conv_bn(P, key + ".InvRes.conv2", r, ctx)
    blk = key + ".mhca_blks.0"
    img = F.conv2d(x, P[blk + ".cpe.proj.weight"], P[blk + ".cpe.proj.bias"], 1, 1, 1, C) + x
    t = img.flatten(2).transpose(1, 2)                                # [B,N,C]
    lay = blk + ".MHCA_layers.0"
    cur = F.layer_norm(t, (C,), P[lay + ".norm1.weight"], P[lay + ".norm1.bias"], LN_EPS)
    t = t + _drop_path(meta_pool(cur), rate, ctx)
    cur = F.layer_norm(t, (C,), P[lay + ".norm2.weight"], P[lay + ".norm2.bias"], LN_EPS)
    h = F.gelu(_dense(F.linear, cur, P[lay + ".mlp.fc1.weight"], P[lay + ".mlp.fc1.bias"]))
    t = t + _drop_path(_dense(F.linear, h, P[lay + ".mlp.fc2.weight"], P[lay + ".mlp.fc2.bias"]), rate, ctx)
    t = t.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _conv_bn(P, key + ".aggregate", torch.cat([r, t], 1), ctx, "hswish")


def mpvit_features(P, key, x, ctx):
    """MPViT.forward_features, nets/tcct.py:733-745 (tiny: 1 path, 1 layer)."""
    x = _conv_bn(P, key + ".stem.0", x, ctx, "hswish", 2, 1)
    x = _conv_bn(P, key + ".stem.1", x, ctx, "hswish", 1, 1)
    outs = []
    for s, dim in enumerate(VIT_DIMS):
        pe = "%s.patch_embed_stages.%d.patch_embeds.0.patch_conv" % (key, s)
        x = F.conv2d(x, P[pe + ".dwconv.weight"], None, 2 if s else 1, 1, 1, dim)
        x = _dense(F.conv2d, x, P[pe + ".pwconv.weight"])
        x = F.hardswish(_bn(P, pe + ".bn", x, ctx))
        x = mhca_stage(P, "%s.mhca_stages.%d" % (key, s), x, dim, DROP_PATH[s], ctx)
        outs.append(x)
    return outs


def _up_block(P, key, x, skip, ctx):
    """MPUpBlock.forward, nets/tcct.py:902-914."""
    x = F.leaky_relu(_bn(P, key + ".prep.1", _conv(P, key + ".prep.0", x, pad=1), ctx), 0.01)
    x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    return _conv(P, key + ".post.0", x + skip)


def norm_add(xs):
    """nets/tcct.py:937-942."""
    xs = [F.normalize(x, dim=1, p=2) for x in xs]
    xs = [F.interpolate(x, size=xs[0].shape[-2:], mode="bilinear", align_corners=False) for x in xs]
    return sum(xs) / len(xs)


def gate_fusion(x1, x2, ctx):
    """GateFusion.forward, nets/tcct.py:916-932: train -> a random field (torch.rand(B, C, max(3, H//32), max(3, W//32)), taken from
    ctx.gate_alphas) up-sampled bicubically and clamped to [0, 1]; eval -> 0.5."""
    if ctx.training:
        alpha = ctx.gate_alphas.pop(0)
        assert alpha.shape == (x1.shape[0], x1.shape[1], max(3, x1.shape[2] // 32), max(3, x1.shape[3] // 32)), alpha.shape
        alpha = F.interpolate(alpha, size=x1.shape[-2:], mode="bicubic").to(x1.device).clamp(0, 1)
    else:
        alpha = 0.5
    return x1 * alpha + x2 * (1 - alpha)


def ftc_forward(P, x, ctx, key="base", flag_vit=True, flag_cnn=True, plain=False, variant="tcct", gate=False):
    """FTC.forward, nets/tcct.py:999-1046: stc_tt (default), cnnu / pnnu (flag_vit=False: the MPViT branch still runs, frozen, and the
    CrossResNet features feed the decoder directly; plain = PlainCNNBlock), gtc_* (gate=True: GateFusion instead of the sum; the wide
    CrossResNet of *_tb needs nothing here, every width comes from the weights), vitu (flag_cnn=False: CrossResNet frozen, x1 = c1, the
    projected MPViT features alone).  variant="onnx": the older decoder of onnx/tcct_goals.py:999-1035 (no t321-t324, auxiliary heads
    on the decoder maps, feats = norm_add([x1,x2,x3,y0,y1,y2])).  Returns ([y0,y1,y2,y4], feats)."""
    if flag_cnn:
        c1, c2, c3, c4, c5 = cross_resnet(P, key + ".base_cnn", x, ctx, plain)
    else:
        with torch.no_grad():
            c1, c2, c3, c4, c5 = cross_resnet(P, key + ".base_cnn", x, ctx, plain)
    if flag_vit:
        v2, v3, v4, v5 = mpvit_features(P, key + ".base_vit", x, ctx)
    else:
        with torch.no_grad():
            mpvit_features(P, key + ".base_vit", x, ctx)

    def tran(name, t):
        return _bn(P, "%s.%s.1" % (key, name), _conv(P, "%s.%s.0" % (key, name), t), ctx)

    x1 = c1
    if flag_vit and flag_cnn:
        fuse = (lambda a, b: gate_fusion(a, b, ctx)) if gate else (lambda a, b: a + b)      # gtc_* / stc_* (tcct.py:974, 934-935)
        x2 = fuse(tran("tran_vit0", v2), tran("tran_cnn0", c2))
        x3 = fuse(tran("tran_vit1", v3), tran("tran_cnn1", c3))
        x4 = fuse(tran("tran_vit2", v4), tran("tran_cnn2", c4))
        x5 = fuse(tran("tran_vit3", v5), tran("tran_cnn3", c5))
    elif flag_cnn:
        x2, x3, x4, x5 = c2, c3, c4, c5
    else:
        x2, x3, x4, x5 = tran("tran_vit0", v2), tran("tran_vit1", v3), tran("tran_vit2", v4), tran("tran_vit3", v5)
    y8 = F.leaky_relu(_bn(P, key + ".head.1", _conv(P, key + ".head.0", x5, pad=1), ctx), 0.01)
    y4 = _up_block(P, key + ".dec1", y8, x4, ctx)
    y2 = _up_block(P, key + ".dec2", y4, x3, ctx)
    y1 = _up_block(P, key + ".dec3", y2, x2, ctx)
    y0 = _up_block(P, key + ".dec4", y1, x1, ctx)
    if variant == "onnx":
        feats = norm_add([x1, x2, x3, y0, y1, y2])
    else:
        y0 = _conv(P, key + ".t324", x1 + y0)
        y1 = _conv(P, key + ".t323", x2 + y1)
        y2 = _conv(P, key + ".t322", x3 + y2)
        y4 = _conv(P, key + ".t321", x4 + y4)
        feats = norm_add([y0, y1, y2])
    size = x.shape[-2:]
    o0 = _conv(P, key + ".aux0", y0)
    o1 = F.interpolate(_conv(P, key + ".aux1", y1), size=size, mode="bilinear", align_corners=False)
    o2 = F.interpolate(_conv(P, key + ".aux2", y2), size=size, mode="bilinear", align_corners=False)
    o4 = F.interpolate(_conv(P, key + ".aux4", y4), size=size, mode="bilinear", align_corners=False)
    return [o0, o1, o2, o4], feats


# ----------------------------------------------------------------------------- losses
def multi_dice(logits, onehot):
    """MultiLoss(DiceLoss).forward, kite/losses/loss.py:83-99 with dice 28-32:
    softmax over C, per class 1-(1+2*sum(p*g))/(1+sum(p)+sum(g)) over the whole batch."""
    p = torch.softmax(logits, 1)
    g = onehot.to(p.dtype)
    inter = (p * g).sum((0, 2, 3))
    union = p.sum((0, 2, 3)) + g.sum((0, 2, 3))
    return (1 - (1 + 2 * inter) / (1 + union)).sum()


def deep_supervision(outs, onehot, coff_ds=1.0):
    """KiteBack.grad_calc with ds=True, kite/loopback.py:62-73."""
    total = 0
    for i in range(len(outs) - 1, 0, -1):
        total = total + multi_dice(outs[i], onehot) * coff_ds
    return total + multi_dice(outs[0], onehot)


def _lap_map(P, x, ctx):
    """RegNet.lap_map, nets/reg.py:71-76: conv3x3 -> BatchNorm2d(1, eps=1) -> conv3x3 -> sigmoid."""
    x = _conv(P, "lap_map.0", x, pad=1)
    x = _bn(P, "lap_map.1", x, ctx, eps=1.0)
    return torch.sigmoid(_conv(P, "lap_map.2", x, pad=1))


def boundary_reg(P, logits, onehot, noise, ctx):
    """RegNet.regular_reg, nets/reg.py:109-156.  `noise` = (eps_pred, eps_true
    [B,C-1,H,W] in (0,1); jit_true, jit_pred [1,1,H,1] in [0,1)) in the order the
    reference draws them (torch.rand_like calls at reg.py:120,120,147,148)."""
    eps_pred, eps_true, jit_true, jit_pred = noise
    pred = logits[:, 1:]
    true = onehot[:, 1:].float()
    B, C, H, W = pred.shape
    prob_true = F.pad((true[:, :, 1:] - true[:, :, :-1]).abs(), (0, 0, 1, 0))
    prob_true = prob_true.sum(1, keepdim=True).clamp_max(1)

    def lap_reg(x):
        x = _conv(P, "lap_reg.0", x, pad=1, groups=C)
        return _conv(P, "lap_reg.1", x, pad=1, groups=C).abs()

    def sampling_softmax(x, eps):
        g = F.softmax(x - torch.log(-torch.log(eps)) / 2, dim=-2)
        return g / (1e-6 + g.sum(-2, keepdim=True))

    ps_pred = _lap_map(P, sampling_softmax(lap_reg(pred), eps_pred).sum(1, keepdim=True), ctx)
    ps_true = _lap_map(P, sampling_softmax(lap_reg(true), eps_true).sum(1, keepdim=True), ctx)
    idx = torch.arange(0, H, dtype=torch.float32, device=pred.device).view(1, 1, -1, 1)
    edge_true = (ps_true * (idx + jit_true - 0.5)).sum(-2) / H
    edge_pred = (ps_pred * (idx + jit_pred - 0.5)).sum(-2) / H
    los_edge = F.mse_loss(edge_pred, edge_true.detach()) + F.mse_loss(edge_pred.detach(), edge_true)
    los_prob = F.mse_loss(ps_true.softmax(-2), prob_true) + F.mse_loss(ps_pred.softmax(-2), prob_true)
    return los_edge + los_prob


def rank_bins(feat_rows, prob, mask, bins=32):
    """points_selection_bins, nets/fcs.py:25-50: rows of the class, sorted by
    probability (descending), cut into 32 equal rank bins, bin means."""
    sel = mask > 0.5
    f = feat_rows[sel]
    p = prob[sel]
    order = torch.sort(p, descending=True)[1]
    n = f.shape[0] // bins
    return torch.stack([f[order[i * n:(i + 1) * n]].mean(0) for i in range(bins)])


def feature_polar(P, feat, logits, onehot):
    """RegNet.regular_udh, nets/reg.py:86-105 with fcs.select1 (fcs.py:82-96),
    fcs.cosinesim/foreach_loss (63-80) and fcp.choice (fcp.py:72-75)."""
    prob = torch.softmax(logits.detach(), 1)
    L = feat.shape[1]
    rows = feat.permute(0, 2, 3, 1).reshape(-1, L)
    loss = 0
    for i in range(onehot.shape[1]):
        pro = rank_bins(rows, prob[:, i].reshape(-1), onehot[:, i].float().round().reshape(-1))
        tgt = P["fcp.buf_grad"][i].unsqueeze(0).expand(pro.shape[0], -1)
        loss = loss - torch.einsum("nc,kc->nk", pro, tgt).mean() / L
    return loss + F.mse_loss(pro, tgt)       # last class only, reg.py:102


def make_noise(batch, n_class, height, width, gen):
    """The four torch.rand_like draws of one regular_reg call (Appendix C order)."""
    shape = (batch, n_class - 1, height, width)
    return (torch.rand(shape, generator=gen), torch.rand(shape, generator=gen),
            torch.rand((1, 1, height, 1), generator=gen), torch.rand((1, 1, height, 1), generator=gen))


def calc_loss(P, img, onehot, ctx, noise=None, udh=True, reg=True,
              coff_ds=1.0, coff_udh=1.0, coff_reg=0.1):
    """KiteSeg.calc_loss, kite/loop_seg.py:146-171.  Returns (total, parts, outs, feats)."""
    outs, feats = ftc_forward(P, img, ctx)
    parts = {"los": deep_supervision(outs, onehot, coff_ds)}
    if udh:
        parts["udh"] = feature_polar(P, feats, outs[0], onehot) * coff_udh
    if reg:
        parts["reg"] = boundary_reg(P, outs[0], onehot, noise, ctx) * coff_reg
    return sum(parts.values()), parts, outs, feats


# ----------------------------------------------------------------------------- train step
def trainable(P):
    """Keys that are nn.Parameters with requires_grad in RegNet(stc_tt(C))."""
    skip = ("running_mean", "running_var", "num_batches_tracked", "cos_dist", "buf_grad", "vec_grad")
    return [k for k in P if k.rsplit(".", 1)[-1] not in skip and ".MHCA_layers.0.cpe." not in k
            and ".MHCA_layers.0.crpe." not in k]


class OracleTrainer:
    """loop_seg.py:108-142 + loopback.py:102-128: calc_loss, backward,
    clip_grad_norm_(12), AdamW(lr, wd=2e-4).  CyclicLR steps per epoch and is
    left to the caller (`lr`)."""

    def __init__(self, P, lr=1e-6, wd=2e-4, **loss_kw):
        self.P = P
        self.keys = trainable(P)
        for k in self.keys:
            P[k].requires_grad_(True)
        # shared cpe module: the alias key reads the same tensor
        for k in list(P):
            if ".MHCA_layers.0.cpe." in k:
                P[k] = P[k.replace(".MHCA_layers.0.cpe.", ".cpe.")]
        self.opt = torch.optim.AdamW([P[k] for k in self.keys], lr=lr, weight_decay=wd)
        self.loss_kw = loss_kw

    def step(self, img, onehot, noise=None, dp_masks=None):
        self.opt.zero_grad()
        ctx = Ctx(True, dp_masks)
        total, parts, outs, feats = calc_loss(self.P, img, onehot, ctx, noise, **self.loss_kw)
        total.backward()
        gnorm = torch.nn.utils.clip_grad_norm_([self.P[k] for k in self.keys], 12)
        self.opt.step()
        return float(total), {k: float(v) for k, v in parts.items()}, float(gnorm)


def predict_labels(P, img, flag_vit=True, **kw):
    """KiteSeg.predict, kite/loop_seg.py:21-33: eval forward, argmax of head 0."""
    with torch.no_grad():
        outs, _ = ftc_forward(P, img, Ctx(False), flag_vit=flag_vit, **kw)
    return outs[0], torch.argmax(torch.softmax(outs[0], 1), 1)


def boundary_positions(x, beta=100.0):
    """Soft-argmax boundary extraction of the inference contract.  PARITY UNPINNED: the reference has no
    implementation of it (only soft_argmax, nets/reg.py:27-35, and the soft-argmax edge of regular_reg, nets/reg.py:146-150);
    this restatement defines it: p = softmax_C(x); d_c[h] = |p_c[h] - p_c[h-1]|, d_c[0] = 0;
    pos[b,c-1,w] = sum_h h * softmax_H(beta * d_c)[h] for c >= 1."""
    p = F.softmax(x.double(), dim=1)[:, 1:]
    d = torch.zeros_like(p)
    d[:, :, 1:] = (p[:, :, 1:] - p[:, :, :-1]).abs()
    w = F.softmax(beta * d, dim=2)
    h = torch.arange(x.shape[2], dtype=w.dtype, device=x.device).view(1, 1, -1, 1)
    return (w * h).sum(2).float()


def soft_argmax(x, beta=100):
    """nets/reg.py:27-35."""
    sm = F.softmax(x * beta, dim=1).clamp(0, 1)
    w = torch.arange(0, x.shape[1], dtype=sm.dtype, device=x.device).view(1, -1, 1, 1)
    return (sm * w).sum(1, keepdim=True)


# ----------------------------------------------------------------------------- validation scores
def val_scores(pr, gt, smooth=1.0):
    """kite/losses/miou.py:28-44,69-91 on [B,C,H,W] maps: per-class MDiceLoss.score and MIouLoss.score (each: mean over the batch of the
    per-image ratio).  Returns (dice [C], iou [C]); scorem(start_idx) = x[start_idx:].mean(), scores = dice.tolist()."""
    pr, gt = pr.double().flatten(2), gt.double().flatten(2)
    inter, sp, sg = (pr * gt).sum(-1), pr.sum(-1), gt.sum(-1)
    dice = ((2 * inter + smooth) / (sp + sg + smooth)).mean(0)
    iou = ((inter + smooth) / (sp + sg - inter + smooth)).mean(0)
    return dice, iou
