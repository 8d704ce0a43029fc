"""Stem convs (3 -> 32 channels, stride 1 | 2) forward and weight gradient at the K2 shape, in-graph, for one or more builds of the
library:  python scripts/time_stem.py [lib.so ...]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit, side_stream
dev = torch.device("cuda:0")
P = ctypes.c_void_p
libs = sys.argv[1:] or [os.path.join(ROOT, "tcct_b200", "lib", "libtcct_b200.so")]
B, H, W = 8, 256, 256
imgs = [torch.rand(B, 3, H, W, device=dev) for _ in range(3)]
w = torch.randn(32, 3, 3, 3, device=dev) * 0.2
b = torch.randn(32, device=dev)
for path in libs:
    lib = ctypes.CDLL(path)
    lib.tcct_stem_conv_fwd.argtypes = [P, P, P, P] + [ctypes.c_int] * 4 + [P, P]
    lib.tcct_stem_conv_wgrad.argtypes = [P, P, P, P] + [ctypes.c_int] * 4 + [P]
    for s in (1, 2):
        Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
        y = torch.empty(B, Ho, Wo, 32, device=dev)
        dys = [torch.randn(B, Ho, Wo, 32, device=dev) for _ in range(3)]
        stats = torch.zeros(64, dtype=torch.float64, device=dev)
        dw = torch.zeros_like(w); db = torch.zeros_like(b)
        i = [0]
        st = lambda: torch.cuda.current_stream().cuda_stream
        def fwd():
            i[0] += 1
            assert lib.tcct_stem_conv_fwd(imgs[i[0] % 3].data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, H, W, s, stats.data_ptr(), st()) == 0
        def wgrad():
            i[0] += 1
            assert lib.tcct_stem_conv_wgrad(imgs[i[0] % 3].data_ptr(), dys[i[0] % 3].data_ptr(), dw.data_ptr(), db.data_ptr(), B, H, W, s, st()) == 0
        tf, tw = timeit(fwd), timeit(wgrad)
        # correctness against fp32 cuDNN-free reference
        torch.backends.cudnn.allow_tf32 = False
        i[0] = 0; stats.zero_(); fwd(); torch.cuda.synchronize()
        ref = torch.nn.functional.conv2d(imgs[1], w, b, stride=s, padding=1).permute(0, 2, 3, 1)
        ef = float((y - ref).abs().max())
        es = float((stats[:32] - ref.double().sum((0, 1, 2))).abs().max() / ref.double().sum((0, 1, 2)).abs().max())
        dw.zero_(); db.zero_(); i[0] = 0; wgrad(); torch.cuda.synchronize()
        x = imgs[1].clone().requires_grad_(False)
        wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
        torch.nn.functional.conv2d(x, wr, br, stride=s, padding=1).backward(dys[1].permute(0, 3, 1, 2))
        ew = float((dw - wr.grad).abs().max() / wr.grad.abs().max()); eb = float((db - br.grad).abs().max() / br.grad.abs().max())
        mb = (imgs[0].numel() + y.numel()) * 4 / 1e6
        print("%s stride %d: fwd %5.1f us (%4.0f GB/s)  wgrad %5.1f us (%4.0f GB/s)   err fwd %.1e stats %.1e dw %.1e db %.1e" % (
            os.path.basename(path), s, tf, mb / tf * 1e3, tw, mb / tw * 1e3, ef, es, ew, eb), flush=True)
