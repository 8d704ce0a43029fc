"""Experiment: where does the tcgen05 conv kernel's time go?  Times conv_line_tma at 8x256x256 with the TCCT_CONV_DBG knobs
(bit 0: no staging writes / TMA stores, bit 1: first K step of every input line only, bit 2: no TMA loads), with and without the
BatchNorm statistics in the epilogue."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from tcct_b200 import ops as O
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv
from time_kernels_util import timeit
dev = torch.device("cuda:0")
B, H, W = 8, 256, 256
for ks in (3, (1, 13)):
    mod = DenseConv(32, 32, ks).to(dev)
    plan = PackPlan(mod, dev)
    O.ARENA.reset(dev); plan.run()
    xs = [torch.randn(B, H, W, 32, device=dev) for _ in range(3)]
    i = [0]
    for want_stats in (False, True):
        for dbg in (0, 1, 2, 3, 4, 5, 6, 7):
            os.environ["TCCT_CONV_DBG"] = str(dbg)
            def fwd():
                i[0] += 1
                with torch.no_grad():
                    mod.run(xs[i[0] % 3], want_stats=want_stats, stats_act=O.ACT_LRELU)
            O.ARENA.reset(dev)
            t = timeit(fwd)
            print("conv %s stats=%d dbg=%2d: %.1f us" % (ks, want_stats, dbg, t), flush=True)
os.environ["TCCT_CONV_DBG"] = "0"
