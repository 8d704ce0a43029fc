"""Import shims that let the unmodified reference (/root/reference/task1) run in
the build container -- TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py;
the reference does not exist on the GPU box).  Stubs the packages the reference
imports but the image lacks (timm, matplotlib) and the two files it imports but
does not ship (kite/utils.py, kite/optims.py; kite/loopback.py:4,6)."""
import random
import sys
import types

import numpy as np
import torch.nn as nn

REF_ROOT = "/root/reference/task1"


class DropPath(nn.Module):
    """timm.models.layers.DropPath semantics (scale_by_keep=True); masks come from
    `DropPath.tape` (list of [B] 0/1 tensors) when set, so runs are reproducible."""
    tape = None

    def __init__(self, p=0.0):
        super().__init__()
        self.p = p

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        keep = 1 - self.p
        if DropPath.tape is not None:
            r = DropPath.tape.pop(0).to(x.dtype).view((x.shape[0],) + (1,) * (x.ndim - 1))
        else:
            r = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * (r / keep)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    _mod("timm"); _mod("timm.models")
    _mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    _mod("timm.models.layers", DropPath=DropPath, trunc_normal_=nn.init.trunc_normal_)
    _mod("matplotlib", rcParams={}); _mod("matplotlib.pyplot", rcParams={})
    _mod("kite.utils", random=random, np=np); _mod("kite.optims")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
