"""1x1-conv GEMM (tcgen05 gemm_tma and mma.sync gemm_px) with and without the BatchNorm-statistics epilogue."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit
from tcct_b200 import ops as O
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv
dev = torch.device("cuda:0")
for (B, H, W, K, N) in ((8, 128, 128, 64, 64), (8, 64, 64, 96, 96), (8, 32, 32, 128, 128), (8, 16, 16, 160, 160)):
    mod = DenseConv(K, N, 1).to(dev)
    plan = PackPlan(mod, dev)
    O.ARENA.reset(dev); plan.run()
    xs = [torch.randn(B, H, W, K, device=dev) for _ in range(3)]
    i = [0]
    def run(stats, act):
        def f():
            i[0] += 1
            with torch.no_grad():
                mod.run(xs[i[0] % 3], want_stats=stats, stats_act=act)
        return f
    def reset():
        O.ARENA.reset(dev)
    O.ARENA.reset(dev)
    t0 = timeit(run(False, 0))
    O.ARENA.reset(dev)
    t1 = timeit(run(True, O.ACT_NONE))
    O.ARENA.reset(dev)
    t2 = timeit(run(True, O.ACT_LRELU))
    print("%dx%dx%d %d->%d: plain %.1f us  +stats(none) %.1f us  +stats(lrelu) %.1f us" % (B, H, W, K, N, t0, t1, t2), flush=True)
