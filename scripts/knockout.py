"""Marginal cost of each kernel family in the K2 train step: replace the family's C entry points by no-ops (results become garbage,
the remaining kernels run as before), re-capture the step graph and time it.  The step is work-bound, not chain-bound, so
`baseline - knocked out` is what a family really costs inside the concurrent schedule -- the serialised launch list overstates it
by the packing factor.      python scripts/knockout.py [steps]"""
import contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import tcct_b200._lib as L
from tcct_b200.kite.loop_seg import KiteSeg
from tcct_b200.nets import RegNet, stc_tt
from tcct_b200.synth import SynthOCT, make_bscans

FAMILIES = {
    "none": [],
    "bn_fwd": ["bn_act2_fwd_bn"],
    "bn_bwd": ["bn_act2_bwd"],
    "conv_tma (fwd+dgrad)": ["conv2d_tma", "conv2d_tma_slice"],
    "wgrad_tma": ["wgrad_tma_partial", "wgrad_tma", "wgrad_reduce_batch"],
    "gemm_tma (fwd+dgrad)": ["gemm_tma", "gemm_tma_gelu", "gemm_tma_dgelu"],
    "wgrad_gemm_tma": ["wgrad_gemm_tma_partial", "wgrad_gemm_tma", "wgrad_reduce_batch"],
    "wgrad_reduce_batch": ["wgrad_reduce_batch"],
    "mma.sync conv/gemm": ["conv2d_nhwc", "gemm_px"],
    "mma.sync wgrad": ["wgrad"],
    "dwconv fwd+bwd": ["dwconv3_fwd", "dwconv3_bwd"],
    "ln_metapool": ["ln_metapool_fwd", "ln_metapool_bwd"],
    "resize/maxpool/l2norm/norm_add": ["resize_nhwc_fwd", "resize_nhwc_bwd", "maxpool2_fwd", "maxpool2_bwd", "l2norm32_fwd", "l2norm32_bwd",
                                       "norm_add3_fwd"],
    "stem + heads": ["stem_conv_fwd", "stem_conv_wgrad", "head_fwd", "head_bwd"],
    "dice": ["dice_multi_fwd", "dice_multi_bwd"],
    "boundary regression": ["breg_forward", "breg_backward"],
    "feature polarisation": ["fpolar_forward", "fpolar_backward"],
}
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dsname, C, K, B, H, W, desc = bench.WORKLOADS["K2"]
img, lab = make_bscans(B, H, W, C, K, 1234)
img, lab = img.cuda(), lab.cuda()


def noop(*a):
    return 0


def noop_parts(*a):
    """*_partial entries report the slab geometry through an int array (second to last argument)."""
    a[-2][0] = 1
    return 0


def run(names):
    saved = {}
    for n in names:
        try:
            saved[n] = getattr(L, n)
        except AttributeError:
            print("   (no entry point %s)" % n)
            continue
        setattr(L, n, noop_parts if n.endswith("_partial") else noop)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(0)
            net = RegNet(stc_tt(C), out_channels=C)
            seg = KiteSeg(bench.make_args(bs=B, graph=True), model=net, dataset=SynthOCT(dsname, H, W, n_batches=2), root="/tmp/tcct_knock")
        seg.model.train()
        for _ in range(seg.GRAPH_WARMUP + 3):
            seg.train_step(img, lab)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            seg.train_step(img, lab)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps
    finally:
        for n, f in saved.items():
            setattr(L, n, f)


base = None
for fam, names in FAMILIES.items():
    try:
        ms = run(names)
    except Exception as e:
        print("%-34s failed: %r" % (fam, e)); continue
    if base is None:
        base = ms
    print("%-34s %7.3f ms/step   marginal %6.0f us (%4.1f %%)" % (fam, ms, (base - ms) * 1e3, 100 * (base - ms) / base), flush=True)
