"""Inference throughput: eval-mode forward + argmax label map + soft-argmax boundary extraction through
KiteSeg.predict_labels / nets.boundary_positions,
full-frame shapes, device-timed over a CUDA-graph replay (inputs resident in HBM) and end to end from pinned host memory.
    python scripts/bench_infer.py [goals|hcms] [batch]"""
import contextlib, io, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench
from tcct_b200.kite.loop_seg import KiteSeg
from tcct_b200.nets import RegNet, stc_tt
from tcct_b200.synth import SynthOCT, make_bscans

ds = sys.argv[1] if len(sys.argv) > 1 else "goals"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
C, K, H, W = (5, 4, 608, 512) if ds == "goals" else (9, 9, 256, 512)
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    torch.manual_seed(0)
    net = RegNet(stc_tt(C), out_channels=C)
    seg = KiteSeg(bench.make_args(bs=B, graph=False), model=net, dataset=SynthOCT(ds, H, W, n_batches=2), root="/tmp/tcct_infer")
seg.model.eval()
img, _ = make_bscans(B, H, W, C, K, 77)
host = img.pin_memory()
dimg = img.to(dev)
out = {}


from tcct_b200.nets import boundary_positions
from tcct_b200.kite.loop_seg import argmax_labels


def step():
    with torch.no_grad():
        logits = seg.model(dimg)[0]
        out["lab"] = argmax_labels(logits)
        out["pos"] = boundary_positions(logits, beta=100.0)


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    step()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 20
e0.record()
for _ in range(N):
    g.replay()
e1.record(); torch.cuda.synchronize()
t_dev = e0.elapsed_time(e1) / N
t0 = time.perf_counter()
for _ in range(N):
    dimg.copy_(host, non_blocking=True)
    g.replay()
    lab = out["lab"].cpu()
    pos = out["pos"].cpu()
torch.cuda.synchronize()
t_e2e = (time.perf_counter() - t0) / N * 1e3
px = B * H * W
print("inference %s full-frame %dx%d bs=%d C=%d: %.3f ms/batch device (%.0f B-scans/s, %.1f TFLOP/s fwd), %.3f ms/batch end to end (%.0f B-scans/s; H2D %d B, D2H %d B)" % (
    ds, H, W, B, C, t_dev, B / t_dev * 1e3, 223699 * px / t_dev / 1e9, t_e2e, B / t_e2e * 1e3, host.numel() * 4, lab.numel() + pos.numel() * 4), flush=True)
