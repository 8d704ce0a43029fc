"""Times the fused LayerNorm + MetaPool + LayerNorm token-mixer kernels at the MPViT stage shapes of workload K2 against the
unfused chain (layernorm, metapool, layernorm)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from tcct_b200 import ops as O
from time_kernels_util import timeit
dev = torch.device("cuda:0")
for (B, N, C) in ((8, 128 * 128, 64), (8, 64 * 64, 96), (8, 32 * 32, 128), (8, 16 * 16, 160)):
    ts = [torch.randn(B, N, C, device=dev, requires_grad=True) for _ in range(4)]
    ws = [torch.nn.Parameter(torch.randn(C, device=dev)) for _ in range(4)]
    for w in ws:
        w._gview = torch.zeros_like(w); w.grad = w._gview
    dy = torch.randn(B, N, C, device=dev)
    i = [0]
    def fused():
        i[0] += 1
        t2, c2 = O.LnMetaPoolFn.apply(ts[i[0] % 4], ws[0], ws[1], ws[2], ws[3], None, 1e-6)
        torch.autograd.backward([t2, c2], [dy, dy])
    def chain():
        i[0] += 1
        t = ts[i[0] % 4]
        cur = O.LayerNormFn.apply(t, ws[0], ws[1], 1e-6)
        t2 = O.MetaPoolFn.apply(t, cur, None)
        c2 = O.LayerNormFn.apply(t2, ws[2], ws[3], 1e-6)
        torch.autograd.backward([t2, c2], [dy, dy])
    def fused_fwd():
        i[0] += 1
        with torch.no_grad():
            O.LnMetaPoolFn.apply(ts[i[0] % 4], ws[0], ws[1], ws[2], ws[3], None, 1e-6)
    tf, tff, tc = timeit(fused, reps=8), timeit(fused_fwd, reps=8), timeit(chain, reps=8)
    mb = B * N * C * 4 / 1e6
    print("ln_metapool B=%d N=%d C=%d (%.1f MB/tensor): fused fwd %.1f us (%.0f GB/s), fwd+bwd %.1f us (%.0f GB/s on 8 passes) | unfused chain fwd+bwd %.1f us" % (
        B, N, C, mb, tff, 3 * mb / tff * 1e3, tf, 8 * mb / tf * 1e3, tc), flush=True)
