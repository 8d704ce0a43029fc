"""Kernel timeline of CUDA-graph replays of the K2 train step (torch.profiler / CUPTI): name, stream, start, duration of every
kernel, written to gpurun_out/trace_step.csv.   python scripts/trace_step.py [workload]"""
import contextlib, csv, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tcct_b200.kite.loop_seg import KiteSeg
from tcct_b200.nets import RegNet, stc_tt
from tcct_b200.synth import SynthOCT, make_bscans
from torch.profiler import profile, ProfilerActivity

wl = sys.argv[1] if len(sys.argv) > 1 else "K2"
dsname, C, K, B, H, W, desc = bench.WORKLOADS[wl]
with contextlib.redirect_stdout(io.StringIO()):
    torch.manual_seed(0)
    net = RegNet(stc_tt(C), out_channels=C)
    seg = KiteSeg(bench.make_args(bs=B, graph=True), model=net, dataset=SynthOCT(dsname, H, W, n_batches=2), root="/tmp/tcct_trace")
seg.model.train()
img, lab = make_bscans(B, H, W, C, K, 1234)
img, lab = img.pin_memory(), lab.pin_memory()
for _ in range(seg.GRAPH_WARMUP + 3):
    seg.train_step(img, lab)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        seg.train_step(img, lab)
    torch.cuda.synchronize()
rows = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        rows.append((ev.time_range.start, ev.time_range.end - ev.time_range.start, getattr(ev, "device_resource_id", -1), ev.name[:100]))
rows.sort()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "trace_step.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["start_us", "dur_us", "stream", "name"])
    w.writerows(rows)
print("events", len(rows))
