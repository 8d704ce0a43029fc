"""Run a few EAGER train steps (no CUDA graph) of workload K2 - a target for ncu launch lists / --set full captures.
    python scripts/prof_step.py [steps] [workload]"""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tcct_b200.kite.loop_seg import KiteSeg
from tcct_b200.nets import RegNet, stc_tt
from tcct_b200.synth import SynthOCT, make_bscans

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
wl = sys.argv[2] if len(sys.argv) > 2 else "K2"
dsname, C, K, B, H, W, desc = bench.WORKLOADS[wl]
with contextlib.redirect_stdout(io.StringIO()):
    torch.manual_seed(0)
    net = RegNet(stc_tt(C), out_channels=C)
    seg = KiteSeg(bench.make_args(bs=B, graph=False), model=net, dataset=SynthOCT(dsname, H, W, n_batches=2), root="/tmp/tcct_prof")
seg.model.train()
img, lab = make_bscans(B, H, W, C, K, 1234)
for i in range(steps):
    torch.cuda.nvtx.range_push("step%d" % i)
    parts = seg.train_step(img, lab)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("done", parts.tolist())
