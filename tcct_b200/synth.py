"""Seeded synthetic OCT B-scans, label maps and model states.

Everything here is deterministic on the CPU generator so the oracle, the
golden-vector script and the CUDA path all see bit-identical inputs
(SynthOCT).  Shapes follow the reference's loader
contract (task1/data/octgen.py:116-128: img float [B,3,H,W] in [0,1],
lab int64 [B,H,W]).
"""
import math
import torch

# dataset name -> (classes C, boundaries K); task1/data/octgen.py:38-62
DATASETS = {"goals": (5, 4), "hcms": (9, 9), "duke": (9, 8), "heg": (8, 7)}


def _smooth_rows(gen, rows, width, knots):
    """rows x width smooth curves in [0,1]: linear interpolation of random knots."""
    ctrl = torch.rand(rows, knots, generator=gen)
    pos = torch.linspace(0, knots - 1, width)
    lo = pos.floor().long().clamp(max=knots - 2)
    frac = pos - lo
    return ctrl[:, lo] * (1 - frac) + ctrl[:, lo + 1] * frac


def make_bscans(batch, height, width, n_class, n_bound=None, seed=1234):
    """Return (img [B,3,H,W] f32 in [0,1], lab [B,H,W] int64).

    K boundary curves b_1<...<b_K per A-scan (column); label of a pixel is
    (number of boundaries above it) mod C, so every class owns a band of at
    least `gap_min` rows in every column (keeps fcs.py:36 away from N=0).
    """
    if n_bound is None:
        n_bound = n_class - 1
    gen = torch.Generator().manual_seed(int(seed))
    top, span = height / 10.0, 0.8 * height
    gap_min = min(6.0, 0.6 * span / (n_bound + 1))
    slack = (span - gap_min * (n_bound + 1)) / (n_bound + 1)
    knots = max(3, width // 32 + 2)
    lab = torch.empty(batch, height, width, dtype=torch.int64)
    rows = torch.arange(height, dtype=torch.float32).view(height, 1)
    for b in range(batch):
        gaps = gap_min + slack * _smooth_rows(gen, n_bound, width, knots)
        bounds = top + torch.cumsum(gaps, 0)                      # [K,W]
        above = (rows.unsqueeze(0) >= bounds.unsqueeze(1)).sum(0)  # [H,W]
        lab[b] = above % n_class
    refl = 0.15 + 0.75 * torch.rand(n_class, generator=gen)
    speckle = torch.rand(batch, height, width, generator=gen)
    gray = (refl[lab] * (0.35 + 0.65 * speckle)).clamp_(0, 1)
    img = gray.unsqueeze(1).expand(batch, 3, height, width).contiguous()
    return img, lab


def synth_state(state, seed=0):
    """Deterministically refill a state_dict-like {key: tensor} mapping.

    Keys that alias one storage (the shared cpe/crpe modules registered twice,
    task1/nets/tcct.py:489-503) receive identical values.  Returns a new dict.
    """
    gen = torch.Generator().manual_seed(int(seed))
    out, by_ptr = {}, {}
    for key, ref in state.items():
        ptr = (ref.data_ptr(), tuple(ref.shape)) if ref.numel() else None
        if ptr is not None and ptr in by_ptr:
            out[key] = out[by_ptr[ptr]].clone()
            continue
        shape = tuple(ref.shape)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            val = torch.zeros(shape, dtype=torch.int64)
        elif leaf == "running_var":
            val = 0.5 + torch.rand(shape, generator=gen)
        elif leaf == "running_mean":
            val = 0.1 * torch.randn(shape, generator=gen)
        elif key == "tau":
            val = torch.full(shape, 100.0)
        elif leaf == "vec_grad":
            val = torch.rand(shape, generator=gen)
        elif leaf == "buf_grad":
            val = torch.nn.functional.normalize(out[key[: -len("buf_grad")] + "vec_grad"], p=2, dim=-1)
        elif leaf == "cos_dist":
            val = ref.detach().clone().float()
        elif leaf == "bias":
            val = 0.05 * torch.randn(shape, generator=gen)
        elif len(shape) == 1:                      # BatchNorm / LayerNorm scale
            val = 1.0 + 0.1 * torch.randn(shape, generator=gen)
        else:                                      # conv / linear weight
            fan_in = max(1, math.prod(shape[1:]))
            val = torch.randn(shape, generator=gen) * math.sqrt(1.6 / fan_in)
        out[key] = val
        if ptr is not None:
            by_ptr[ptr] = key
    return out


class SynthOCT:
    """Duck-types the reference dataset object used by KiteSeg
    (task1/data/octgen.py:28-100: out_channels, trainSet, valSet, parse)."""

    def __init__(self, dbname="goals", height=256, width=256, n_batches=4, seed=1234):
        self.__name__ = dbname
        self.out_channels, self.n_bound = DATASETS[dbname]
        self.in_channels = 3
        self.height, self.width, self.n_batches, self.seed = height, width, n_batches, seed

    def _iter(self, bs, n, seed):
        for i in range(n):
            img, lab = make_bscans(bs, self.height, self.width, self.out_channels, self.n_bound, seed + i)
            yield {"img": img, "lab": lab, "tag": ["synth%d" % (seed + i)] * bs}

    def trainSet(self, bs=8, data="train"):
        return list(self._iter(bs, self.n_batches, self.seed))

    def valSet(self, bs=1, data="val"):
        return list(self._iter(bs, max(1, self.n_batches // 2), self.seed + 7919))

    testSet = valSet

    def parse(self, pics):
        return pics["img"], pics["lab"], pics["tag"], None
