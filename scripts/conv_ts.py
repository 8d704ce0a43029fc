"""Experiment: per-tile clock64 stamps of CTA 0 of conv_line_tma (3x3, 8x256x256): who waits for whom."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tcct_b200 import ops as O
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv
dev = torch.device("cuda:0")
B, H, W = 8, 256, 256
for ks in (3,):
    mod = DenseConv(32, 32, ks).to(dev)
    plan = PackPlan(mod, dev)
    O.ARENA.reset(dev); plan.run()
    x = torch.randn(B, H, W, 32, device=dev)
    ts = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
    with torch.no_grad():
        for _ in range(3):
            mod.run(x, want_stats=False)
        torch.cuda.synchronize()
        os.environ["TCCT_CONV_TS"] = str(ts.data_ptr())
        # thrash L2 first
        junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev); junk.zero_()
        mod.run(x, want_stats=False)
        torch.cuda.synchronize()
        del os.environ["TCCT_CONV_TS"]
    t = ts.view(64, 16).cpu()
    t0 = int(t[0, 0])
    print("conv", ks, "line: mma[prewait, full_ok, line_start, tempty_ok, pieces, issued] epi[tfull, ld_done, bar1, staged]  (cycles since start)")
    for n in range(10):
        r = [int(v) - t0 if int(v) else -1 for v in t[n]]
        print("%2d  mma %6d %6d %6d %6d %6d [elect %6d first %6d rest %6d commit %6d] %6d | epi %6d %6d %6d %6d" % (n, r[0], r[1], r[2], r[3], r[4], r[6], r[12], r[13], r[14], r[5], r[8], r[9], r[10], r[11]))
