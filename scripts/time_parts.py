"""In-graph time of the sub-networks of one train step (CUDA events around CUDA-graph replays):
CrossResNet fwd+bwd, MPViT fwd+bwd, the whole network fwd+bwd with random upstream gradients, the full step.
    python scripts/time_parts.py [workload]"""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench
from time_kernels_util import timeit
from tcct_b200 import ops as O
from tcct_b200.kite.loop_seg import KiteSeg
from tcct_b200.nets import RegNet, stc_tt
from tcct_b200.synth import SynthOCT, make_bscans

wl = sys.argv[1] if len(sys.argv) > 1 else "K2"
dsname, C, K, B, H, W, desc = bench.WORKLOADS[wl]
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    torch.manual_seed(0)
    net = RegNet(stc_tt(C), out_channels=C)
    seg = KiteSeg(bench.make_args(bs=B, graph=False), model=net, dataset=SynthOCT(dsname, H, W, n_batches=2), root="/tmp/tcct_parts")
seg.model.train()
img, lab = make_bscans(B, H, W, C, K, 1234)
img = img.to(dev); lab8 = seg._label_map(lab.to(dev)) if hasattr(seg, "_label_map") else lab.to(dev)
base = seg.model.base
gen = torch.Generator(device=dev).manual_seed(1)


def run_part(fn_outs):
    def body():
        base.begin_step(dev)
        outs = fn_outs()
        torch.autograd.backward(outs, [torch.full_like(o, 1e-3) for o in outs])      # upstream gradients made inside the capture
    return body


def cnn_outs():
    return list(base.base_cnn(img))


def vit_outs():
    return list(base.base_vit.forward_features(img))


def net_outs():
    outs = base.forward_impl(img)
    return list(outs) + [base.feats_nhwc]


def full():
    total, parts = seg._losses(img, lab8)
    total.backward()


rows = [("CrossResNet fwd+bwd", run_part(cnn_outs)), ("MPViT fwd+bwd", run_part(vit_outs)),
        ("stc_tt fwd+bwd (random upstream grads, both encoders concurrent)", run_part(net_outs)),
        ("full step without optimizer (fwd + Dice x4 + FP + BR + bwd)", full)]
for conc in (True, False):
    O.CONCURRENT = conc
    for name, fn in rows[2:] if not conc else rows:
        t = timeit(fn, reps=2, replays=5)
        print("%-75s concurrent=%d  %8.1f us" % (name, conc, t), flush=True)
