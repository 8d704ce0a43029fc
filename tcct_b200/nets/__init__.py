from .tcct import *  # noqa: F401,F403
from .tcct import stc_tt, tcct, FTC, MPViT, mpvit_tiny, CrossResNet, CrossCNNBlock, MPUpBlock, norm_add  # noqa: F401
