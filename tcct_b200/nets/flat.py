"""Flat parameter / gradient storage and the per-step weight re-pack plan.

All parameters of a model live in ONE fp32 buffer (and their gradients in a second one) so that
clip_grad_norm_ + AdamW (kite/loopback.py:126-128, loop_seg.py:128-130 in the reference) is two kernel
launches and the data-parallel gradient exchange is one all-reduce.  `Parameter.data` / `.grad` are views,
so `state_dict()` / `load_state_dict()` / foreign optimizers keep working."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops as O
from ..ops import _p, _stream


def _align(n, a=4):
    return (n + a - 1) // a * a


class FlatParams:
    def __init__(self, module, device, unused=()):
        """`unused`: parameter-name substrings that never receive a gradient on the stc_tt path
        (crpe, cls_head, fuse, lap_epl, tau); they are laid out after the used ones so the
        optimizer touches only [0, n_used)."""
        named = list(module.named_parameters())
        used = [(k, p) for k, p in named if p.requires_grad and not any(u in k for u in unused)]
        rest = [(k, p) for k, p in named if not (p.requires_grad and not any(u in k for u in unused))]
        total = sum(_align(p.numel()) for _, p in used + rest)
        self.n_used = sum(_align(p.numel()) for _, p in used)
        self.buf = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(total, dtype=torch.float32, device=device)
        self.names, self.offsets = [], {}
        off = 0
        with torch.no_grad():
            for k, p in used + rest:
                n = p.numel()
                view = self.buf[off: off + n].view(p.shape)
                view.copy_(p.data.to(device=device, dtype=torch.float32))
                p.data = view
                p._gview = self.grad[off: off + n].view(p.shape)
                p.grad = None
                self.names.append(k)
                self.offsets[k] = (off, n)
                off += _align(n)
        self.params = [p for _, p in used + rest]
        self.used_params = [p for _, p in used]
        # non-parameter state (BN running statistics, prototypes) just moves to the device
        for mod in module.modules():
            for name, b in list(mod._buffers.items()):
                if b is not None and b.device != device:
                    mod._buffers[name] = b.to(device)

    def owns(self, p):
        return p.data.untyped_storage().data_ptr() == self.buf.untyped_storage().data_ptr()

    def attach_grads(self):
        """Make every used parameter's .grad the flat view; zero the buffer if any was detached
        (e.g. by optimizer.zero_grad(set_to_none=True))."""
        detached = any(p.grad is not p._gview for p in self.used_params)
        if detached:
            self.grad.zero_()
            for p in self.used_params:
                p.grad = p._gview

    def zero_grad(self):
        self.grad.zero_()
        for p in self.used_params:
            p.grad = p._gview


class PackPlan:
    """One kernel launch re-packs every dense conv / linear weight into MMA fragment order
    (csrc/conv_mma.cu: pack_weights_kernel), forward and transposed (dgrad) variants."""

    DTYPE = np.dtype([("w", "<u8"), ("out", "<u8"), ("N", "<i4"), ("K", "<i4"), ("T", "<i4"), ("sn", "<i4"),
                      ("sk", "<i4"), ("st", "<i4"), ("flip", "<i4"), ("first", "<i4"), ("fmt", "<i4"), ("pad", "<i4")])

    def __init__(self, module, device):
        assert self.DTYPE.itemsize == L.tcct_pack_entry_size(), "PackEntry ABI mismatch"
        specs = []      # (owner, attr, weight tensor, elem offset, N, K, T, sn, sk, st, flip)
        for mod in module.modules():
            if getattr(mod, "dense_kind", None) is None:
                continue
            w = mod.weight
            if mod.dense_kind == "spatial":
                Cout, Cin, KH, KW = w.shape
                T = KH * KW
                if Cin <= 64 and Cout <= 64:
                    specs.append((mod, "pk_f", w, 0, Cout, Cin, T, Cin * T, T, 1, 0))
                    specs.append((mod, "pk_b", w, 0, Cin, Cout, T, T, Cin * T, 1, 1))
                else:
                    # wide layers (the 64..256-channel CrossResNet / decoder of stc_tb, gtc_tb): the reduction runs in slices of
                    # 64 / 32 channels, one pack per slice and direction (ops.reduction_slices)
                    mod.pk_f, mod.pk_b = [], []
                    for (c0, sz) in O.reduction_slices(Cin):
                        specs.append((mod, "pk_f+", w, c0 * T, Cout, sz, T, Cin * T, T, 1, 0))
                    for (c0, sz) in O.reduction_slices(Cout):
                        specs.append((mod, "pk_b+", w, c0 * Cin * T, Cin, sz, T, T, Cin * T, 1, 1))
                if Cin == 32 and Cout == 32:     # tcgen05 K-major 128-byte-swizzled operand rows (csrc/conv_tma.cu)
                    specs.append((mod, "pk_tf", w, 0, Cout, Cin, T, Cin * T, T, 1, 0, 2))
                    specs.append((mod, "pk_tb", w, 0, Cin, Cout, T, T, Cin * T, 1, 1, 2))
                elif Cin == 32 and Cout % 32 == 0 and Cout <= 128:
                    # wider outputs run as 32 -> 32 channel slices: the forward pack holds one output tile after the other, the
                    # transposed (data-gradient) pack is one K = 32 slice of the output channels per entry
                    specs.append((mod, "pk_tf", w, 0, Cout, Cin, T, Cin * T, T, 1, 0, 2))
                    mod.pk_tb = []
                    for s_ in range(Cout // 32):
                        specs.append((mod, "pk_tb+", w, 32 * s_ * Cin * T, Cin, 32, T, T, Cin * T, 1, 1, 2))
            else:   # "gemm": 1x1 conv or Linear, optionally split along the input channels
                N = w.shape[0]
                ktot = w.numel() // N
                mod.pk_f, mod.pk_b, mod.pk_tf, mod.pk_tb = [], [], [], []
                for (k0, K) in mod.k_slices:
                    specs.append((mod, "pk_f+", w, k0, N, K, 1, ktot, 1, 0, 0))
                    specs.append((mod, "pk_b+", w, k0, K, N, 1, 1, ktot, 0, 0))
                    specs.append((mod, "pk_tf+", w, k0, N, K, 1, ktot, 1, 0, 0, 3))     # tcgen05 GEMM rows (csrc/gemm_tma.cu)
                    specs.append((mod, "pk_tb+", w, k0, K, N, 1, 1, ktot, 0, 0, 3))
        table = np.zeros(len(specs), dtype=self.DTYPE)
        first = 0
        sizes = []
        specs = [sp if len(sp) == 12 else sp + (0,) for sp in specs]
        for i, (mod, attr, w, k0, N, K, T, sn, sk, st, flip, fmt) in enumerate(specs):
            n = _align(N, 32) * K * T
            table[i] = (0, 0, N, K, T, sn, sk, st, flip, first, fmt, 0)
            sizes.append(n)
            first += n
        self.total = first
        self.packed = torch.zeros(2 * max(first, 1), dtype=torch.float32, device=device)   # hi plane | lo plane
        base = self.packed.data_ptr()
        for i, (mod, attr, w, k0, N, K, T, sn, sk, st, flip, fmt) in enumerate(specs):
            f = int(table[i]["first"])
            view = self.packed[f: f + sizes[i]]
            table[i]["w"] = w.data_ptr() + 4 * k0
            table[i]["out"] = base + 4 * f
            if attr.endswith("+"):
                getattr(mod, attr[:-1]).append(view)
            else:
                setattr(mod, attr, view)
        self.n = len(specs)
        self.table = torch.from_numpy(table.view(np.uint8).copy()).to(device)
        self._ptrs = [w.data_ptr() for (_, _, w, *_r) in specs]
        self._ws = [w for (_, _, w, *_r) in specs]

    def valid(self):
        return all(w.data_ptr() == p for w, p in zip(self._ws, self._ptrs))

    def run(self):
        from .. import ops as O
        L.pack_weights(_p(self.table), self.n, self.total, _stream())
        O.STATE["lo_off"] = self.total if O.STATE["x3"] else 0


class FlatModule(nn.Module):
    """Mixin for the outermost module of a model: owns the flat buffers and the pack plan."""

    UNUSED = (".crpe.", "cls_head.", "fuse.", "lap_epl.", "tau")

    def flat_state(self, device):
        st = self.__dict__.get("_flat_state")
        if st is None or st[0].buf.device != device or not all(st[0].owns(p) for p in st[0].params[:4]) or not st[1].valid():
            flat = FlatParams(self, device, self.UNUSED)
            plan = PackPlan(self, device)
            st = (flat, plan)
            self.__dict__["_flat_state"] = st
        return st

    def begin_step(self, device):
        """Called at the top of forward: re-zero scratch, attach gradients, re-pack weights."""
        from ..ops import ARENA, STEP_STREAM, reset_wgrad
        flat, plan = self.flat_state(device)
        ARENA.reset(device)
        reset_wgrad()
        STEP_STREAM[0] = torch.cuda.current_stream(device) if device.type == "cuda" else None
        if self.training and torch.is_grad_enabled():
            flat.attach_grads()
        plan.run()
        return flat
