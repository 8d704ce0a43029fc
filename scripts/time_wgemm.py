"""wgrad_gemm_tma (1x1-conv / Linear weight gradient) and gemm_tma at the shapes of the K2 step."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from time_kernels_util import timeit
import tcct_b200._lib as L
from tcct_b200.ops import _p, _stream
dev = torch.device("cuda:0")
for (M, K, N) in ((524288, 32, 32), (131072, 32, 32), (131072, 64, 64), (131072, 64, 96), (32768, 96, 96), (32768, 96, 128), (131072, 32, 64)):
    if not L.tcct_wgrad_gemm_tma_supported(M, K, N):
        print("unsupported", M, K, N); continue
    xs = [torch.randn(M, K, device=dev) for _ in range(3)]
    dys = [torch.randn(M, N, device=dev) for _ in range(3)]
    dw = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
    ws = torch.empty(int(L.tcct_wgrad_gemm_tma_ws_floats(M, K, N)), device=dev)
    cnt = torch.zeros(8, dtype=torch.int32, device=dev)
    i = [0]
    res = []
    for bias in (True, False):
        def f():
            i[0] += 1
            cnt.zero_()
            L.wgrad_gemm_tma(_p(xs[i[0] % 3]), _p(dys[i[0] % 3]), _p(dw), _p(db) if bias else None, M, K, N, K, _p(ws), _p(cnt), _stream())
        res.append(timeit(f))
    mb = M * (K + N) * 4 / 1e6
    print("wgrad_gemm_tma M=%d K=%d N=%d: %.1f us with dbias, %.1f us without (%.0f MB -> %.0f GB/s)" % (M, K, N, res[0], res[1], mb, mb / res[0] * 1e3), flush=True)
