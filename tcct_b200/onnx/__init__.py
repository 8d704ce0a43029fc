"""Deploy-side drop-in for task1/onnx/onnx_infer.py (the reference's onnxruntime wrapper)."""
from .onnx_infer import NetWork  # noqa: F401
