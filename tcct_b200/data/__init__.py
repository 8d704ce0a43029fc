"""Input / deploy formats either side of the hot path (drop-in names of task1/data/octnpy.py)."""
from .octnpy import EyeSetResource  # noqa: F401
