"""`from nets import *` surface of the reference (task1/nets/__init__.py:2-3)."""
from .tcct import *  # noqa: F401,F403
from .tcct import stc_tt, tcct, cnnu, pnnu, vitu, stc_tt_onnx, PlainCNNBlock, FTC, MPViT, mpvit_tiny, CrossResNet, CrossCNNBlock, MPUpBlock, norm_add  # noqa: F401
from .reg import RegNet, boundary_positions, soft_argmax  # noqa: F401
from .fcs import FeatConSuper, points_selection_bins  # noqa: F401
from .fcp import FeatConPolar  # noqa: F401
