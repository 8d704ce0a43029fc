"""clip_grad_norm_(12) + AdamW as two kernel launches over the flat parameter / gradient buffers
(task1/kite/loop_seg.py:128-130, task1/kite/loopback.py:126-128).  The step counter, learning rate and the
gradient norm live on the device, so the update is CUDA-graph capturable and never synchronises."""
import ctypes

import torch

from .. import _lib as L
from ..ops import ARENA, _p, _stream


class FlatAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction) over FlatParams.

    `n_active`: number of leading elements of the flat buffer that are trained (parameters that never
    receive a gradient are laid out behind them and, like in torch where their .grad stays None, are not
    touched -- not even by weight decay)."""

    def __init__(self, flat, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=2e-4, max_norm=12.0, n_active=None):
        self.flat = flat
        self.n_active = flat.n_used if n_active is None else n_active
        super().__init__(flat.used_params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        dev = flat.buf.device
        self.m = torch.zeros(self.n_active, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.n_active, dtype=torch.float32, device=dev)
        self.dev_state = torch.zeros(4, dtype=torch.float32, device=dev)     # [step, lr, last grad norm, -]
        self.sqnorm = torch.zeros(2, dtype=torch.float64, device=dev)
        self.max_norm = float(max_norm)
        self.grad_scale = 1.0
        self._lr_on_device = None
        self.sync_lr()

    def sync_lr(self):
        """Push param_groups[0]['lr'] (changed by the per-epoch CyclicLR) to the device scalar."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_on_device:
            self.dev_state[1].fill_(lr)
            self._lr_on_device = lr

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        n = self.n_active
        self.sqnorm.zero_()
        L.sqnorm(_p(self.flat.grad), n, _p(self.sqnorm), _stream())
        L.adamw_step(_p(self.flat.buf), _p(self.flat.grad), _p(self.m), _p(self.v), n, _p(self.sqnorm), _p(self.dev_state),
                     self.max_norm, float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                     float(self.grad_scale), _stream())

    def last_grad_norm(self):
        return float(self.dev_state[2])
