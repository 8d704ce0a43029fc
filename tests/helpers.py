"""Shared builders for the parity tests: states, inputs and noise exactly as
oracle/make_golden.py made them."""
import os

import numpy as np
import torch
import torch.nn.functional as F

import tcct_oracle as O
from tcct_b200.synth import make_bscans, synth_state

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_state_template(n_class):
    """{key: zero tensor} in the reference's state_dict order (tests/golden/state_keys.txt)."""
    out = {}
    with open(os.path.join(GOLDEN, "state_keys.txt")) as f:
        for line in f:
            c, key, shape, *rest = line.split()
            if int(c) != n_class or key == "!trainable":
                continue
            dims = () if shape == "-" else tuple(int(s) for s in shape.split("x"))
            out[key] = torch.zeros(dims, dtype=getattr(torch, rest[0]))
    return out


def alias_template(tmpl):
    """Make the shared cpe/crpe keys alias one storage, like the reference module tree."""
    for k in list(tmpl):
        if ".MHCA_layers.0.cpe." in k or ".MHCA_layers.0.crpe." in k:
            tmpl[k] = tmpl[k.replace(".MHCA_layers.0.", ".")]
    return tmpl


def golden_state(n_class, seed):
    tmpl = alias_template(golden_state_template(n_class))
    # give every distinct tensor a distinct storage pointer so aliasing is detected by pointer
    return synth_state(tmpl, seed)


def dp_masks(batch, gen):
    rates = [r for r in O.DROP_PATH if r > 0 for _ in range(2)]
    return [(torch.rand(batch, generator=gen) < 1 - r).float() for r in rates]


def train_inputs(meta):
    n_class, n_bound, batch, height, width, seed = (int(v) for v in meta)
    gen = torch.Generator().manual_seed(seed + 100)
    img, lab = make_bscans(batch, height, width, n_class, n_bound, seed)
    onehot = F.one_hot(lab, n_class).permute(0, 3, 1, 2)
    noise = O.make_noise(batch, n_class, height, width, gen)
    masks = dp_masks(batch, gen)
    return img, lab, onehot, noise, masks


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


class TF32Emu:
    """Operand rounding of a TF32 tensor-core contraction, for tcct_oracle.QUANT: every dense conv / linear sees its
    activation, weight and (in backward) output-gradient operands with a 10-bit mantissa, products and sums stay fp32.
    `truncate` = what tcgen05 kind::tf32 does with fp32 operands; otherwise round-to-nearest (cvt.rna, the mma.sync path)."""

    def __init__(self, truncate=True):
        self.truncate = truncate

    def _q(self, t):
        i = t.contiguous().view(torch.int32)
        if not self.truncate:
            i = i + 0x1000
        return (i & ~0x1FFF).view(torch.float32)

    def inp(self, t):
        emu = self

        class Q(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x):
                return emu._q(x)

            @staticmethod
            def backward(ctx, g):
                return g
        return Q.apply(t)

    def out(self, t):
        emu = self

        class Q(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x):
                return x.view_as(x)

            @staticmethod
            def backward(ctx, g):
                return emu._q(g)
        return Q.apply(t)


def oracle_train_pass(n_class, seed, img, onehot, noise, masks, quant=None, **loss_kw):
    """One oracle calc_loss + backward (no optimizer step): (total, parts, outs, feats, P, trainable keys)."""
    old = O.QUANT
    O.QUANT = quant
    try:
        P = golden_state(n_class, seed)
        tr = O.OracleTrainer(P, lr=1e-4)
        tr.opt.zero_grad()
        total, parts, outs, feats = O.calc_loss(P, img, onehot, O.Ctx(True, [m.clone() for m in masks]), noise, **loss_kw)
        total.backward()
    finally:
        O.QUANT = old
    return total, parts, outs, feats, P, tr.keys


def variant_state(name, n_class, seed):
    """State of a SimpleFusion factory as oracle/make_golden_variants.py built it: the stc_tt-shaped synthetic state, with the centre
    taps cut out of the cross convs for `pnnu` (1x3 / 3x1 kernels)."""
    state = golden_state(n_class, seed)
    if name == "pnnu":
        for k, v in list(state.items()):
            if ".block34.0.weight" in k and v.shape[3] > 3:
                w0 = (v.shape[3] - 3) // 2
                state[k] = v[:, :, :, w0:w0 + 3].contiguous()
            elif ".block34.1.weight" in k and v.shape[2] > 3:
                h0 = (v.shape[2] - 3) // 2
                state[k] = v[:, :, h0:h0 + 3, :].contiguous()
    return state


def factory_state(name, n_class, seed):
    """FTC-level state ("base_cnn...", no RegNet extras) of a gated / wide factory as oracle/make_golden_wide.py built it:
    synth_state over the reference's own state_dict order -- stc_tt's keys for gtc_tt, tests/golden/state_keys_tb.txt (written from the
    reference's stc_tb) for the wide CrossResNet models."""
    tmpl = {}
    if name == "gtc_tt":
        for k, v in golden_state_template(n_class).items():
            if k.startswith("base."):
                tmpl[k[5:]] = v
    else:
        with open(os.path.join(GOLDEN, "state_keys_tb.txt")) as f:
            for line in f:
                c, key, shape, dt = line.split()
                dims = () if shape == "-" else tuple(int(s_) for s_ in shape.split("x"))
                tmpl[key] = torch.zeros(dims, dtype=getattr(torch, dt))
    return synth_state(alias_template(tmpl), seed)


VARIANT_KW = {"pnnu": dict(flag_vit=False, plain=True), "vitu": dict(flag_vit=True, flag_cnn=False), "cnnu": dict(flag_vit=False)}


def miou_inputs(meta):
    """The maps oracle/make_golden_miou.py scored: softmax probabilities (or their hard argmax) and an int64 one-hot target."""
    B, C, H, W, seed, hard = (int(v) for v in meta)
    g = torch.Generator().manual_seed(seed)
    pr = torch.softmax(torch.randn(B, C, H, W, generator=g) * 2, 1)
    gt = F.one_hot(torch.randint(0, C, (B, H, W), generator=g), C).permute(0, 3, 1, 2)
    if hard:
        pr = F.one_hot(pr.argmax(1), C).permute(0, 3, 1, 2).float()
    return pr, gt
