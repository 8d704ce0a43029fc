"""Benchmark of the stc_tt train step (fwd + bwd + Dice x4 + feature-polarisation + boundary-regression loss +
clip_grad_norm + AdamW) on synthetic GOALS-shaped B-scans.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload K2|K3|K1]

N > 1 is launched by `python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...` (one rank per GPU,
weak scaling: bs=8 per GPU, one flat-bucket NCCL all-reduce per step).  Rank 0 prints ONE JSON line on stdout;
everything else goes to stderr.

  value        B-scans/s with the batch already resident in HBM (CUDA-graph replay, device-timed, max over ranks)
  e2e          B-scans/s through KiteSeg.train_step with pinned HOST buffers (H2D of image+labels and D2H of the
               loss inside the timed region, every step)
  roofline     the dominant kernel (3x3 conv 32->32 on the full-resolution stage) timed alone with CUDA events
  cpu_baseline the oracle (oracle/tcct_oracle.py, the CPU restatement of the reference) on the host cores
`--impl reference` times that CPU path alone (the reference itself cannot travel to the GPU box)."""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # name -> (dataset, classes, boundaries, batch per GPU, H, W, description)
    "K1": ("goals", 5, 4, 2, 256, 256, "K1: GOALS-shaped 256x256 crops, C=5, bs=2"),
    "K2": ("goals", 5, 4, 8, 256, 256, "K2: GOALS-shaped (800x1100 -> 608x512 -> 256x256 train crop per reference), C=5, bs=8 per GPU"),
    "K3": ("hcms", 9, 9, 8, 256, 256, "K3: HCMS-shaped (496x1024 -> 256x512 -> 256x256 train crop), C=9, bs=8 per GPU"),
}
TRAIN_FLOP_PER_PX = 3 * 223699          # fwd 223 699 FLOP/px (C=5), step ~ 3x


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def make_args(**kw):
    ns = argparse.Namespace(los="di", lr=1e-2, gpu="0", pl=False, bs=8, bug=False, udh=True, coff_udh=1.0, reg=True,
                            coff_reg=0.1, epl=False, coff_epl=0.1, coff_ds=1.0, graph=True)
    ns.__dict__.update(kw)
    return ns


# ----------------------------------------------------------------------------- CPU path (oracle)
def cpu_steps(wl, steps, warmup, threads):
    """The reference's train step restated on the CPU (oracle/tcct_oracle.py: calc_loss + backward + clip + AdamW)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tcct_oracle as orc
    from helpers import dp_masks, golden_state
    from tcct_b200.synth import make_bscans
    _, C, K, B, H, W, _ = WORKLOADS[wl]
    torch.set_num_threads(threads)
    P = golden_state(C, 0)
    tr = orc.OracleTrainer(P, lr=1e-6)
    gen = torch.Generator().manual_seed(4321)
    times = []
    for i in range(warmup + steps):
        img, lab = make_bscans(B, H, W, C, K, 1234 + i)
        onehot = torch.nn.functional.one_hot(lab, C).permute(0, 3, 1, 2)
        noise = orc.make_noise(B, C, H, W, gen)
        t0 = time.perf_counter()
        tr.step(img, onehot, noise, dp_masks(B, gen))
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        log("[cpu] step %d: %.3f s" % (i, dt))
    return B * len(times) / sum(times), sum(times) / len(times)


def run_reference(a):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    wl = a.workload
    threads = os.cpu_count() or 1
    steps, warmup = max(1, min(a.steps, 6)), max(1, min(a.warmup, 1))
    value, sec = cpu_steps(wl, steps, warmup, threads)
    _, C, K, B, H, W, desc = WORKLOADS[wl]
    line = {"impl": "reference", "metric": "train B-scans/s (fwd+bwd+Dice x4+FP+BR loss+clip+AdamW)", "value": value,
            "unit": "B-scans/s", "n_gpus": a.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "batch": B, "height": H, "width": W, "classes": C},
            "cpu_baseline": {"value": value, "unit": "B-scans/s", "cores": threads, "kind": "port",
                             "sample": "%d full train steps of the workload batch (bs=%d) after %d warm-up, torch CPU fp32, "
                                       "oracle/tcct_oracle.py (the reference needs /root/reference, absent on the GPU box)" % (steps, B, warmup)},
            "e2e": {"value": value, "unit": "B-scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU path
def roofline_probe(torch, B, H, W):
    """Dominant kernel: 3x3 conv 32->32 (+bias, LeakyReLU batch statistics) on the full-resolution map,
    conv_tile_kernel<32> of csrc/conv_mma.cu.  Algorithmic bytes: read 32 fp32 + write 32 fp32 per pixel = 256 B/px."""
    from tcct_b200 import ops as O
    from tcct_b200.nets.flat import PackPlan
    from tcct_b200.nets.tcct import DenseConv
    dev = torch.device("cuda", torch.cuda.current_device())
    mod = DenseConv(32, 32, 3).to(dev)
    plan = PackPlan(mod, dev)
    O.ARENA.reset(dev)
    plan.run()
    reps = 12
    xs = [torch.randn(B, H, W, 32, device=dev) for _ in range(3)]      # 3 x 67 MB inputs + outputs > L2
    with torch.no_grad():
        for x in xs:
            mod.run(x, want_stats=True, stats_act=O.ACT_LRELU)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for i in range(reps):
            O.ARENA.reset(dev)
            mod.run(xs[i % 3], want_stats=True, stats_act=O.ACT_LRELU)
        ev[1].record()
        torch.cuda.synchronize()
    sec = ev[0].elapsed_time(ev[1]) * 1e-3 / reps
    px = B * H * W
    return {"kernel": "conv_tile_kernel<32,false> 3x3 32->32 @ %dx%dx%d" % (B, H, W), "seconds": sec,
            "bytes": 256 * px, "flops": 2 * 288 * 32 * px}


def run_ours(a):
    import torch
    import torch.distributed as dist
    wl = a.workload
    dsname, C, K, B, H, W, desc = WORKLOADS[wl]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    with contextlib.redirect_stdout(sys.stderr):
        import tcct_b200._lib as L
        from tcct_b200.kite.loop_seg import KiteSeg
        from tcct_b200.nets import RegNet, stc_tt
        from tcct_b200.synth import SynthOCT, make_bscans
        if L.tcct_device_arch() != 100:
            log("warning: device arch is %d, kernels are built for sm_100a" % L.tcct_device_arch())
        torch.manual_seed(0)
        dataset = SynthOCT(dsname, H, W, n_batches=4, seed=1234 + rank)
        net = RegNet(stc_tt(C), out_channels=C)
        seg = KiteSeg(make_args(bs=B), model=net, dataset=dataset, root=os.path.join("/tmp", "tcct_bench_%d" % rank))
        n_host = 4
        host = []
        for i in range(n_host):
            img, lab = make_bscans(B, H, W, C, K, 1234 + 97 * rank + i)
            host.append((img.pin_memory(), lab.pin_memory()))          # reference loader format: f32 image, int64 labels
        dev_batches = [(i.to(dev), seg._label_map(l)) for i, l in host]
        seg.model.train()
        key = (tuple(dev_batches[0][0].shape), tuple(dev_batches[0][1].shape))

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- warm-up (eager steps, graph capture, replays)
        launches0 = L.tcct_launch_count()
        for i in range(max(a.warmup, seg.GRAPH_WARMUP + 2)):
            if i == seg.GRAPH_WARMUP:
                launches0 = L.tcct_launch_count()
            seg.train_step(*host[i % n_host])
            if i == seg.GRAPH_WARMUP:
                launches_per_step = L.tcct_launch_count() - launches0      # kernels recorded into the graph for one step
        g = seg._graphs[key]
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        # ---- device-resident throughput: inputs already in HBM, graph replay only
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        barrier()
        ev[0].record()
        for i in range(a.steps):
            img, lab8 = dev_batches[i % n_host]
            g.step(img, lab8)
        ev[1].record()
        barrier()
        t_dev = torch.tensor([ev[0].elapsed_time(ev[1]) * 1e-3], device=dev, dtype=torch.float64)
        # ---- end to end through the public API: pinned host batch -> H2D -> step -> D2H of the loss
        barrier()
        t0 = time.perf_counter()
        ev[0].record()
        last = 0.0
        for i in range(a.steps):
            parts = seg.train_step(*host[i % n_host])
            last = parts.cpu()[3].item()
        ev[1].record()
        barrier()
        t_e2e = torch.tensor([max(ev[0].elapsed_time(ev[1]) * 1e-3, 0.0)], device=dev, dtype=torch.float64)
        wall_e2e = time.perf_counter() - t0
        sampler.stop_flag = True
        if world > 1:
            dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(t_dev), float(t_e2e)
        log("rank %d: device %.3f ms/step, e2e %.3f ms/step (wall %.3f), last loss %.4f, launches/step %d" % (
            rank, t_dev / a.steps * 1e3, t_e2e / a.steps * 1e3, wall_e2e / a.steps * 1e3, last, launches_per_step))
        if rank != 0:
            return
        sampler.join(timeout=2)
        roof = roofline_probe(torch, B, H, W)
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
        achieved = roof["bytes"] / roof["seconds"] / 1e9
        cpu = None
        if world == 1 and not a.no_cpu:
            threads = os.cpu_count() or 1
            v, sec = cpu_steps(wl, 3, 1, threads)
            cpu = {"value": v, "unit": "B-scans/s", "cores": threads, "kind": "port",
                   "sample": "3 full train steps of the same batch shape (bs=%d, %dx%d, C=%d) after 1 warm-up, "
                             "oracle/tcct_oracle.py on torch CPU fp32 (%.2f s/step)" % (B, H, W, C, sec)}
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    line = {"metric": "train B-scans/s (fwd+bwd+Dice x4+FP+BR loss+clip+AdamW)", "value": world * B * a.steps / t_dev,
            "unit": "B-scans/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, seg.GRAPH_WARMUP + 2),
            "ms_per_step": t_dev / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 tensor-core contractions, fp32 storage/accumulate/statistics", "data": "synthetic",
            "config": {"workload": desc, "batch_per_gpu": B, "height": H, "width": W, "classes": C,
                       "parallelism": "dp%d" % world, "l2": "4 rotating input batches; per-step working set (~1.5 GB of activations) exceeds the 126 MB L2",
                       "step_tflops": TRAIN_FLOP_PER_PX * B * H * W * world * a.steps / t_dev / 1e12},
            "e2e": {"value": world * B * a.steps / t_e2e, "unit": "B-scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16},
            "gpu_launches": int(launches_per_step) * a.steps,
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": None, "kernel": roof["kernel"], "us_per_launch": roof["seconds"] * 1e6,
                         "tflops": roof["flops"] / roof["seconds"] / 1e12, "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)"},
            "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="K2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    if a.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(a)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
