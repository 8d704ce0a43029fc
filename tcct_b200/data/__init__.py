"""Input / deploy formats either side of the hot path (drop-in names of task1/data/octnpy.py)."""
from .octnpy import EyeSetResource  # noqa: F401
from .octgen import ALB_TWIST, EyeSetGenerator, GpuTwist, make_tran, pack_params, read_pair_aug  # noqa: F401
