"""One launch each of the depthwise 3x3 (stride 1) and stem kernels at the K2 shapes, twice (the first round warms up), for
  ncu --set full --clock-control none --import-source on -k regex:"dwconv3_s1|stem_conv" --launch-skip 5 -c 5 -o gpurun_out/dw_stem python scripts/ncu_dw_stem.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tcct_b200._lib as L
from tcct_b200.ops import _p, _stream
dev = torch.device("cuda:0")
B, H, W, C = 8, 128, 128, 64
x = torch.randn(B, H, W, C, device=dev); dy = torch.randn(B, H, W, C, device=dev)
w = torch.randn(C, 1, 3, 3, device=dev); b = torch.randn(C, device=dev)
y = torch.empty_like(x); dx = torch.empty_like(x); dw = torch.zeros_like(w); db = torch.zeros_like(b)
img = torch.rand(8, 3, 256, 256, device=dev); sw = torch.randn(32, 3, 3, 3, device=dev); sb = torch.randn(32, device=dev)
sy = torch.empty(8, 256, 256, 32, device=dev); sdy = torch.randn(8, 256, 256, 32, device=dev)
stats = torch.zeros(64, dtype=torch.float64, device=dev); sdw = torch.zeros_like(sw); sdb = torch.zeros_like(sb)
for _ in range(2):
    L.dwconv3_fwd(_p(x), _p(w), _p(b), _p(y), B, H, W, C, 1, 0, None, _stream())
    L.dwconv3_bwd(_p(x), _p(w), _p(dy), _p(dx), None, None, B, H, W, C, 1, 0, _stream())
    L.dwconv3_bwd(_p(x), _p(w), _p(dy), None, _p(dw), _p(db), B, H, W, C, 1, 0, _stream())
    L.stem_conv_fwd(_p(img), _p(sw), _p(sb), _p(sy), 8, 256, 256, 1, _p(stats), _stream())
    L.stem_conv_wgrad(_p(img), _p(sdy), _p(sdw), _p(sdb), 8, 256, 256, 1, _stream())
    torch.cuda.synchronize()
