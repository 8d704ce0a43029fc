"""RegNet -- drop-in for task1/nets/reg.py:38-157: wraps the segmentation net and owns the two auxiliary
regularisers of the training step,

  * regular_reg : boundary regression   (reg.py:109-156)  -> csrc/breg.cu   via ops.BoundaryRegFn
  * regular_udh : feature polarisation  (reg.py:86-105)   -> csrc/fpolar.cu via ops.FeaturePolarFn

with the reference's parameter tree (lap_epl, lap_reg, lap_map, tau, fcs, fcp) so checkpoints load unchanged."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops as O
from .fcp import FeatConPolar
from .fcs import FeatConSuper
from .flat import FlatModule


def soft_argmax(x, beta=100):
    """reg.py:27-35 (unused by the reference's training path): sum_c c * softmax_C(beta * x) -> [B,1,H,W].  Inference helper:
    no gradient is defined for it here."""
    O._check(x)
    B, C, H, W = x.shape
    xs = O._c(x.detach().float())
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
    O.L.soft_argmax(O._p(xs), O._p(out), B, C, H * W, float(beta), O._stream())
    return out


def boundary_positions(x, beta=100.0):
    """Soft-argmax boundary extraction (this repository's definition, the reference has none):
    p = softmax_C(x); pos[b,c-1,w] = sum_h h * softmax_H(beta * |p_c[h] - p_c[h-1]|) for the classes c >= 1 -> [B,C-1,W]."""
    O._check(x)
    B, C, H, W = x.shape
    xs = O._c(x.detach().float())
    out = torch.empty((B, C - 1, W), dtype=torch.float32, device=x.device)
    O.L.boundary_positions(O._p(xs), O._p(out), B, C, H, W, float(beta), O._stream())
    return out


class RegNet(FlatModule):
    __name__ = 'reg'
    tmp = {}
    noise_tape = None     # tests: (eps_pred, eps_true, jit_true, jit_pred) consumed by the next regular_reg call

    def __init__(self, base, out_channels=5, con='cor', num_emb=32):
        super().__init__()
        self.base = base
        self.UNUSED = getattr(base, "UNUSED", FlatModule.UNUSED)
        self.__name__ = base.__name__
        self.out_channels = out_channels
        self.fcs = FeatConSuper(con=con)
        self.fcp = FeatConPolar(num_cls=out_channels, num_emb=32, init=False)
        self.lap_epl = nn.Sequential(nn.Conv2d(out_channels, 1, 3, 1, 1), nn.Conv2d(1, 1, 3, 1, 1), nn.Sigmoid())
        dim_reg = out_channels - 1
        self.lap_reg = nn.Sequential(nn.Conv2d(dim_reg, dim_reg, 3, 1, 1, groups=dim_reg),
                                     nn.Conv2d(dim_reg, dim_reg, 3, 1, 1, groups=dim_reg))
        self.lap_map = nn.Sequential(nn.Conv2d(1, 1, 3, 1, 1), nn.BatchNorm2d(1, 1), nn.Conv2d(1, 1, 3, 1, 1), nn.Sigmoid())
        self.tau = nn.Parameter(torch.ones(size=(1,), dtype=torch.float32) * 100)
        self.emb_list = self.tgt_list = None

    def forward(self, x):
        self.begin_step(x.device)
        return self.base.forward_impl(x)

    # ------------------------------------------------------------------ feature polarisation
    def regular_udh(self, pred, true, tau=5):
        lab = O.labels_u8(true, pred.shape[1])
        feat = self.base.feats_nhwc
        proto = self.fcp.buf_grad
        if proto.device != pred.device:
            proto = self.fcp.buf_grad = proto.to(pred.device)
        return O.FeaturePolarFn.apply(feat, pred.detach().contiguous(), lab, proto.contiguous())

    # ------------------------------------------------------------------ boundary regression
    loss = nn.MSELoss()

    def _noise(self, B, Cm, H, W, device):
        if RegNet.noise_tape is not None:
            eps_pred, eps_true, jit_true, jit_pred = RegNet.noise_tape
            RegNet.noise_tape = None
            eps = torch.stack([eps_pred, eps_true]).to(device=device, dtype=torch.float32).contiguous()
            jit = torch.stack([jit_pred.reshape(-1), jit_true.reshape(-1)]).to(device=device, dtype=torch.float32).contiguous()
            return eps, jit
        return torch.rand((2, B, Cm, H, W), device=device), torch.rand((2, H), device=device)

    def regular_reg(self, pred, true, tau=100):
        B, C, H, W = pred.shape
        lab = O.labels_u8(true, C)
        eps, jit = self._noise(B, C - 1, H, W, pred.device)
        return O.BoundaryRegFn.apply(pred.contiguous(), lab, eps, jit, self, self.training)
