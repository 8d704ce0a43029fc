"""Time single kernels at the stage-0/1 shapes of workload K2 with CUDA events (GPU box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tcct_b200 import ops as O
from tcct_b200.nets.flat import PackPlan
from tcct_b200.nets.tcct import DenseConv

dev = torch.device("cuda:0")

def timeit(fn, reps=12):
    """Device time per call in us: `reps` calls captured into one CUDA graph (no host launch gaps), replayed 3 times."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3

for (B, H, W) in ((8, 256, 256), (8, 128, 128)):
    for ks in (3, (1, 13), (13, 1)):
        mod = DenseConv(32, 32, ks).to(dev)
        plan = PackPlan(mod, dev)
        O.ARENA.reset(dev); plan.run()
        xs = [torch.randn(B, H, W, 32, device=dev) for _ in range(3)]
        dy = torch.randn(B, H, W, 32, device=dev)
        i = [0]
        def fwd():
            i[0] += 1
            with torch.no_grad():
                mod.run(xs[i[0] % 3], want_stats=True, stats_act=O.ACT_LRELU)
        px = B * H * W
        res = {}
        for umma in (True, False):
            O.set_umma(umma)
            O.ARENA.reset(dev)
            res[umma] = timeit(fwd)
        O.set_umma(True)
        import tcct_b200._lib as L
        from tcct_b200.ops import _p, _stream
        T = mod.weight.shape[2] * mod.weight.shape[3]
        dw = torch.zeros_like(mod.weight); db = torch.zeros(32, device=dev)
        def wg():
            L.wgrad(_p(xs[0]), _p(dy), _p(dw), _p(db), B, H, W, 32, 32, mod.weight.shape[2], mod.weight.shape[3], 32 * T, T, 1, 0, _stream())
        tw = timeit(wg)
        tw2 = float("nan")
        if L.tcct_wgrad_tma_supported(H, W, 32, 32, mod.weight.shape[2], mod.weight.shape[3]):
            ws = torch.empty(int(L.tcct_wgrad_tma_ws_floats(B, H, W, mod.weight.shape[2], mod.weight.shape[3])), device=dev)
            cnt = torch.zeros(64, dtype=torch.int32, device=dev)
            def wg2():
                cnt.zero_()
                L.wgrad_tma(_p(xs[0]), _p(dy), _p(dw), _p(db), B, H, W, mod.weight.shape[2], mod.weight.shape[3], 32, _p(ws), _p(cnt), _stream())
            tw2 = timeit(wg2)
        flops = 2 * 32 * 32 * T * px
        print("conv %s @ %dx%dx%d: tcgen05+TMA %.1f us (%.0f GB/s, %.0f TF/s) | mma.sync %.1f us | wgrad tcgen05+TMA %.1f us (%.0f GB/s) | wgrad mma.sync %.1f us" % (
            ks, B, H, W, res[True], 256 * px / res[True] / 1e3, flops / res[True] / 1e6, res[False], tw2, 256 * px / tw2 / 1e3, tw), flush=True)
