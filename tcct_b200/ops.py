"""Autograd bindings of the C-ABI kernels (include/tcct_b200.h).

Activations are NHWC fp32 tensors of shape [B, H, W, C] (tokens [B, N, C] share the layout); the module
boundary (images in, logits out) is NCHW like the reference.  Every op launches on torch's current CUDA
stream, allocates only through torch's caching allocator and never synchronises, so a whole train step
is capturable in one CUDA graph.  There is no CPU path: tensors must live on a CUDA device.

Weight gradients are accumulated by the kernels straight into the flat gradient buffer when the
parameter carries a `_gview` (see tcct_b200.nets.flat.FlatParams); otherwise a fresh tensor is returned
to autograd as usual."""
import ctypes
import os

import torch

from . import _lib as L

ACT_NONE, ACT_LRELU, ACT_HSWISH, ACT_GELU = 0, 1, 2, 3

# Contraction precision: "tf32" (one tensor-core product per term, what cuDNN does for the reference on a GPU
# by default) or "tf32x3" (error-compensated split products, fp32-faithful; used to calibrate parity tests).
STATE = {"x3": False, "lo_off": 0, "umma": True, "fuse_mlp": os.environ.get("TCCT_FUSE_MLP", "1") != "0"}


def set_umma(enabled):
    """Route the eligible 32->32 spatial convs through the TMA-fed tcgen05 kernel (csrc/conv_tma.cu); default on.
    Off = the warp-level mma.sync kernels for every shape (used by tests to cross-check the two)."""
    STATE["umma"] = bool(enabled)



def set_fuse_mlp(enabled):
    """MHCA-block MLP as one autograd node with the GELU / GELU' passes inside the GEMM epilogues (`MlpFn`); default on.
    Off = fc1, activation pass, fc2 as separate operators (tests cross-check the two)."""
    STATE["fuse_mlp"] = bool(enabled)


def set_precision(mode):
    if mode not in ("tf32", "tf32x3"):
        raise ValueError("precision must be 'tf32' or 'tf32x3'")
    STATE["x3"] = mode == "tf32x3"


def get_precision():
    return "tf32x3" if STATE["x3"] else "tf32"


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(*tensors):
    for t in tensors:
        if t is None:
            continue
        if isinstance(t, (list, tuple)):
            _check(*t)
            continue
        if not t.is_cuda:
            raise RuntimeError("tcct_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)
        if not t.is_contiguous():
            raise RuntimeError("tcct_b200 ops need contiguous tensors (shape %s, stride %s)" % (tuple(t.shape), t.stride()))


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------- concurrent branches
# The two encoders of the network are independent until the fusion convs, so one of them is issued on a side
# stream: most kernels of the small maps cannot fill 148 SMs on their own.  (Forking the two conv chains of a
# CrossCNNBlock or the auxiliary losses as well was measured and gave nothing: 9.98 vs 9.91 ms per step.)  autograd replays each
# backward node on the stream of its forward, so the backward overlaps the same way; under CUDA-graph capture the
# forks become parallel branches of the graph.
CONCURRENT = True
_SIDE = {}
# Priority offset per fork index relative to the dependent-chain priority (negative = more urgent; the device has four levels,
# 0 .. -3).  The MPViT encoder (0) with its inverted-residual branches (1) is the longest dependent chain of the forward and of the
# backward pass and the feature-polarisation loss (2) the longest of the loss phase: they go first; the decoder's side branches
# (4: projections / auxiliary heads, 5: the two fusion blocks needed last) yield to the decoder chain; weight gradients run at 0.
# Measured on the K2 step (one box): all chains equal 6.587 ms; side branches lower 6.546; + MPViT higher 6.518; + FP higher 6.504.
FORK_PRIORITY = {0: -1, 1: -1, 2: -1, 4: 1, 5: 1}
if os.environ.get("TCCT_FORK_PRIORITY") is not None:
    FORK_PRIORITY = {int(k): int(v) for k, v in (kv.split(":") for kv in os.environ["TCCT_FORK_PRIORITY"].split(",") if kv)}


def fork(device, idx):
    """Side stream `idx` of `device`, ordered after everything issued so far on the current stream (None when
    branch concurrency is switched off)."""
    if not CONCURRENT:
        return None
    key = (device.index, idx)
    st = _SIDE.get(key)
    if st is None:
        st = _SIDE[key] = torch.cuda.Stream(device=device, priority=max(-3, min(0, CHAIN_PRIORITY + FORK_PRIORITY.get(idx, 0))))
    st.wait_stream(torch.cuda.current_stream(device))
    return st


def join(side, *tensors):
    """The current stream waits for `side`; tensors produced there are marked as used here (allocator safety)."""
    if side is None:
        return
    cur = torch.cuda.current_stream(side.device)
    cur.wait_stream(side)
    for t in tensors:
        if t is not None:
            t.record_stream(cur)


class on:
    """`with on(side):` -- issue on the side stream, or on the current one when side is None."""

    def __init__(self, side):
        self.ctx = torch.cuda.stream(side) if side is not None else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


# --------------------------------------------------------------------------- weight gradients off the critical path
# In a backward node the data gradient feeds the next node, the weight gradient feeds only the optimizer.  The weight-
# gradient kernels are therefore issued on their own stream (ordered after the kernel that produced dy) and joined once,
# after the whole backward (`join_wgrad`, called by the training loop before clip + AdamW / the gradient all-reduce).
# On the small maps of the deep stages, where every kernel is latency-bound, this takes ~1/3 of the launches off the
# dependent chain; under CUDA-graph capture the forks become parallel branches of the graph.  The join is automatic: the
# first fork of a backward pass queues an autograd end-of-backward callback that makes every forking stream wait.
WGRAD_ASYNC = True
WGRAD_STREAMS = int(os.environ.get("TCCT_WGRAD_STREAMS", "4"))       # round-robin pool (K2 step: 1 -> 7.86, 2 -> 7.53, 4 -> 7.46, 6 -> 7.43 ms)
CHAIN_PRIORITY = -int(os.environ.get("TCCT_CHAIN_PRIORITY", "2"))     # dependent-chain streams above the weight-gradient pool (priority 0)
WGRAD_PRIORITY = -int(os.environ.get("TCCT_WGRAD_PRIORITY", "0"))     # below every dependent chain by default
_WGRAD = {}
_WGRAD_FORKERS = []
_WGRAD_NEXT = [0]


_JOIN_QUEUED = [False]


def _queue_join():
    """One end-of-backward callback per backward pass (joins the weight-gradient streams, flushes the deferred reductions)."""
    if not _JOIN_QUEUED[0]:
        _JOIN_QUEUED[0] = True
        torch.autograd.Variable._execution_engine.queue_callback(join_wgrad)


class wgrad_side:
    """`with wgrad_side(direct, x, dy):` -- issue on the weight-gradient stream when the gradients go straight into the
    flat buffer (`direct`); tensors read there are marked for the caching allocator."""

    def __init__(self, direct, *tensors):
        self.ctx = None
        if not (WGRAD_ASYNC and direct):
            return
        dev = tensors[0].device
        pool = _WGRAD.get(dev.index)
        if pool is None:
            pool = _WGRAD[dev.index] = [torch.cuda.Stream(device=dev, priority=WGRAD_PRIORITY) for _ in range(max(1, WGRAD_STREAMS))]
        st = pool[_WGRAD_NEXT[0] % len(pool)]
        _WGRAD_NEXT[0] += 1
        cur = torch.cuda.current_stream(dev)
        st.wait_stream(cur)
        _queue_join()
        if all(cur != f for f in _WGRAD_FORKERS):
            _WGRAD_FORKERS.append(cur)
        for t in tensors:
            if t is not None:
                t.record_stream(st)
        self.ctx = torch.cuda.stream(st)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def reset_wgrad():
    """Start of a step: forget forks of a backward pass that never reached its end-of-backward callback (an exception
    inside backward skips the engine's final callbacks; a stale list would suppress the join of every later pass)."""
    del _WGRAD_FORKERS[:]
    del _REDUCE_JOBS[:]
    _WGRAD_NEXT[0] = 0
    _JOIN_QUEUED[0] = False


# The tcgen05 weight-gradient kernels end in a small partial-sum reduction.  Issued per layer it is 48 latency-bound launches per
# step on the very streams that finish the step; instead the main kernels leave their partials behind (`defer_reduce`) and ONE launch
# folds all of them into the flat gradient buffer when the backward pass ends (`join_wgrad`).
STEP_STREAM = [None]    # the stream the current step was started on (nets/flat.py: begin_step)
DEFER_REDUCE = os.environ.get("TCCT_DEFER_REDUCE", "1") != "0"
_REDUCE_JOBS = []      # (ReduceJob fields, workspace tensor kept alive)


def defer_reduce(ws, dw_ptr, kind, nparts, p0, p1, p2):
    _queue_join()
    _REDUCE_JOBS.append(((ws.data_ptr() if hasattr(ws, "data_ptr") else ws, dw_ptr, kind, nparts, p0, p1, p2, 0), ws))


def flush_reduce():
    """Fold every deferred partial-sum workspace into its gradient: one launch on the current stream."""
    if not _REDUCE_JOBS:
        return
    jobs = (L.ReduceJob * len(_REDUCE_JOBS))(*[L.ReduceJob(*f) for f, _ in _REDUCE_JOBS])
    cur = torch.cuda.current_stream()
    for _, ws in _REDUCE_JOBS:
        if hasattr(ws, "record_stream"):
            ws.record_stream(cur)
    L.wgrad_reduce_batch(ctypes.cast(jobs, ctypes.c_void_p), len(_REDUCE_JOBS), _stream())
    del _REDUCE_JOBS[:]


def join_wgrad():
    """Every stream that forked weight-gradient kernels in this backward pass waits for them (end-of-backward callback)."""
    forkers = list(_WGRAD_FORKERS)
    del _WGRAD_FORKERS[:]
    _WGRAD_NEXT[0] = 0
    _JOIN_QUEUED[0] = False
    for cur in forkers:
        for st in _WGRAD.get(cur.device.index, ()):
            cur.wait_stream(st)
    if _REDUCE_JOBS:
        # the deferred reductions run once, on the stream the step continues on (clip + AdamW / the gradient all-reduce follow there);
        # it has just been made to wait for every weight-gradient stream
        home = STEP_STREAM[0] if STEP_STREAM[0] is not None else torch.cuda.current_stream()
        for st in _WGRAD.get(home.device.index, ()):
            home.wait_stream(st)
        with torch.cuda.stream(home):
            flush_reduce()
        for cur in forkers + [torch.cuda.current_stream(home.device)]:
            if cur != home:
                cur.wait_stream(home)


# --------------------------------------------------------------------------- scratch arena
class Arena:
    """Zeroed float64 scratch for per-channel statistics / reduction buffers.  Slices are handed out
    sequentially; `reset()` (start of every forward) re-zeroes everything any pass has ever dirtied: `peak` only grows,
    because a CUDA-graph replay writes up to the captured pass's high-water mark without going through `take`
    (an eager eval forward between two train steps must not shrink what the next train step zeroes)."""

    def __init__(self):
        self.buf = None
        self.off = 0
        self.peak = 0

    def reset(self, device):
        if self.buf is None or self.buf.device != device:
            self.buf = torch.zeros(1 << 18, dtype=torch.float64, device=device)
            self.peak = 0
        elif self.peak:
            self.buf[: self.peak].zero_()
        self.off = 0

    def take(self, n, device):
        n = (n + 1) & ~1
        if self.buf is None or self.buf.device != device:
            self.reset(device)
        if self.off + n > self.buf.numel():
            # start a new zeroed block; the old one stays alive through the slices already handed out
            self.buf = torch.zeros(max(self.buf.numel() * 2, n), dtype=torch.float64, device=device)
            self.off = 0
            self.peak = 0
        out = self.buf[self.off: self.off + n]
        self.off += n
        self.peak = max(self.peak, self.off)
        return out


ARENA = Arena()


def _grad_target(p):
    """(buffer the kernel accumulates into, whether it is the parameter's own flat-gradient view)."""
    gv = getattr(p, "_gview", None)
    if gv is not None:
        if p.grad is None:
            p.grad = gv
        if p.grad is gv:
            return gv, True
    return torch.zeros_like(p, memory_format=torch.contiguous_format), False


def _ret(buf, direct):
    return None if direct else buf


# --------------------------------------------------------------------------- dense convs / GEMMs
def reduction_slices(channels):
    """[(first channel, width)] with widths 64 / 32: how a conv wider than 64 channels on its reduction side is cut into launches of
    the warp-level kernel (csrc/conv_mma.cu: tcct_conv2d_nhwc_slice); nets/flat.py packs the weights slice by slice the same way."""
    out, c = [], 0
    while channels - c >= 64:
        out.append((c, 64))
        c += 64
    if channels - c:
        out.append((c, channels - c))
    return out


def _sliced_conv(x, packs, bias, y, B, H, W, Cred, Cout, KH, KW, stats, stats_act):
    """y = conv(x) as a chain of 64 / 32-channel reduction slices accumulating into y in place."""
    sl = reduction_slices(Cred)
    for i, (c0, sz) in enumerate(sl):
        L.conv2d_nhwc_slice(ctypes.c_void_p(x.data_ptr() + 4 * c0), Cred, _p(packs[i]), STATE['lo_off'], _p(bias) if i == 0 else None, _p(y),
                            B, H, W, sz, Cout, KH, KW, _p(y) if i > 0 else None, _p(stats) if i == len(sl) - 1 else None, stats_act, _stream())


class Conv2dFn(torch.autograd.Function):
    """Dense spatial conv (3x3, 1xk, kx1), stride 1, 'same' zero padding, NHWC.
    Returns (y, stats) with stats = per-channel [sum | sum sq] of stats_act(y) when want_stats."""

    @staticmethod
    def forward(ctx, x, w, b, pk_f, pk_b, want_stats, stats_act, pk_tf=None, pk_tb=None):
        _check(x, pk_f, b)
        B, H, W, Cin = x.shape
        Cout, _, KH, KW = w.shape
        y = torch.empty((B, H, W, Cout), dtype=torch.float32, device=x.device)
        stats = ARENA.take(2 * Cout, x.device) if want_stats else None
        tma = (pk_tf is not None and STATE["umma"] and not STATE["x3"] and Cin == 32 and Cout % 32 == 0
               and bool(L.tcct_conv_tma_supported(H, W, 32, 32, KH, KW)))
        if tma and Cout == 32:
            L.conv2d_tma(_p(x), _p(pk_tf), _p(b), _p(y), B, H, W, KH, KW, _p(stats), stats_act, _stream())
        elif tma:
            # 32 -> 64 (MPViT stem): one 32 -> 32 slice per output tile, each a full tcgen05 launch over the shared input
            blk = KH * KW * 1024
            for co in range(Cout // 32):
                bias = ctypes.c_void_p(b.data_ptr() + 128 * co) if b is not None else None
                L.conv2d_tma_slice(_p(x), 32, 0, ctypes.c_void_p(pk_tf.data_ptr() + 4 * blk * co), bias, _p(y), Cout, 32 * co, 0,
                                   B, H, W, KH, KW, _p(stats), stats_act, _stream())
        elif isinstance(pk_f, list):
            _sliced_conv(x, pk_f, b, y, B, H, W, Cin, Cout, KH, KW, stats, stats_act)
        else:
            L.conv2d_nhwc(_p(x), _p(pk_f), STATE['lo_off'], _p(b), _p(y), B, H, W, Cin, Cout, KH, KW, None, None, _p(stats), stats_act, _stream())
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x)
        ctx.w, ctx.b, ctx.pk_b, ctx.pk_tb, ctx.tma = w, b, pk_b, pk_tb, tma
        ctx.mark_non_differentiable(*([stats] if want_stats else []))
        return (y, stats) if want_stats else (y, None)

    @staticmethod
    def backward(ctx, dy, _ds):
        (x,) = ctx.saved_tensors
        w, b = ctx.w, ctx.b
        dy = _c(dy)
        B, H, W, Cin = x.shape
        Cout, _, KH, KW = w.shape
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if ctx.tma and Cout == 32:
                L.conv2d_tma(_p(dy), _p(ctx.pk_tb), None, _p(dx), B, H, W, KH, KW, None, 0, _stream())
            elif ctx.tma:
                # the reduction over the output channels runs as one launch per 32-channel slice; the later ones add into dx
                for k, pk in enumerate(ctx.pk_tb):
                    L.conv2d_tma_slice(_p(dy), Cout, 32 * k, _p(pk), None, _p(dx), 32, 0, int(k > 0), B, H, W, KH, KW, None, 0, _stream())
            elif isinstance(ctx.pk_b, list):
                _sliced_conv(dy, ctx.pk_b, None, dx, B, H, W, Cout, Cin, KH, KW, None, 0)
            else:
                L.conv2d_nhwc(_p(dy), _p(ctx.pk_b), STATE['lo_off'], None, _p(dx), B, H, W, Cout, Cin, KH, KW, None, None, None, 0, _stream())
        dw, dwd = _grad_target(w)
        db, dbd = _grad_target(b) if b is not None else (None, True)
        with wgrad_side(dwd and dbd, x, dy):
            if STATE["umma"] and not STATE["x3"] and bool(L.tcct_wgrad_tma_supported(H, W, Cin, Cout, KH, KW)):
                per = int(L.tcct_wgrad_tma_ws_floats(B, H, W, KH, KW))
                if DEFER_REDUCE and dwd and dbd:
                    ws = torch.empty(per * (Cout // 32), dtype=torch.float32, device=x.device)
                    parts = (ctypes.c_int * 4)()
                    L.wgrad_tma_partial(_p(x), _p(dy), _p(db), B, H, W, KH, KW, Cout, _p(ws), parts, _stream())
                    for s_ in range(Cout // 32):
                        defer_reduce(ws if s_ == 0 else ws.data_ptr() + 4 * per * s_, dw.data_ptr() + 4 * 32 * s_ * 32 * KH * KW, 0,
                                     parts[0], parts[1], parts[2], parts[3])
                else:
                    ws = torch.empty(per, dtype=torch.float32, device=x.device)
                    L.wgrad_tma(_p(x), _p(dy), _p(dw), _p(db), B, H, W, KH, KW, Cout, _p(ws), None, _stream())
            elif Cin == 32:
                L.wgrad(_p(x), _p(dy), _p(dw), _p(db), B, H, W, Cin, Cout, KH, KW, Cin * KH * KW, KH * KW, 1, int(STATE['x3']), _stream())
            else:       # wide input: one launch per 32-channel slice of x, each writing its own columns of dW
                T = KH * KW
                for s_ in range(Cin // 32):
                    L.wgrad_slice(ctypes.c_void_p(x.data_ptr() + 128 * s_), Cin, _p(dy), ctypes.c_void_p(dw.data_ptr() + 128 * s_ * T),
                                  _p(db) if s_ == 0 else None, B, H, W, Cout, KH, KW, Cin * T, T, 1, int(STATE['x3']), _stream())
        return dx, _ret(dw, dwd), _ret(db, dbd), None, None, None, None, None, None


def _wgrad_gemm(x, dy, dw_ptr, db, M, K, N, ld, direct):
    """Weight / bias gradient of a 1x1 conv or Linear on the tcgen05 kernel; the partial-sum reduction is deferred to the end of the
    backward pass when the gradient goes straight into the flat buffer."""
    ws = torch.empty(int(L.tcct_wgrad_gemm_tma_ws_floats(M, K, N)), dtype=torch.float32, device=x.device)
    if DEFER_REDUCE and direct:
        parts = (ctypes.c_int * 1)()
        L.wgrad_gemm_tma_partial(_p(x), _p(dy), _p(db), M, K, N, _p(ws), parts, _stream())
        defer_reduce(ws, dw_ptr, 1, parts[0], N, K, ld)
    else:
        L.wgrad_gemm_tma(_p(x), _p(dy), ctypes.c_void_p(dw_ptr), _p(db), M, K, N, ld, _p(ws), None, _stream())


class GemmFn(torch.autograd.Function):
    """1x1 conv / Linear over the last dim:  y = [res + res_scale[b] *] (x @ Wslice^T + bias).
    `w` is the full parameter ([N, Ktot] or [N, Ktot, 1, 1]); this op uses columns [k0, k0+K)."""

    @staticmethod
    def forward(ctx, x, w, b, pk_f, pk_b, k0, res, res_scale, want_stats, stats_act, pk_tf=None, pk_tb=None):
        _check(x, pk_f, b, res, res_scale)
        K = x.shape[-1]
        N = w.shape[0]
        M = x.numel() // K
        y = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32, device=x.device)
        stats = ARENA.take(2 * N, x.device) if want_stats else None
        pps = M // x.shape[0]
        tc = pk_tf is not None and STATE["umma"] and not STATE["x3"]
        if tc and bool(L.tcct_gemm_tma_supported(M, K, N)):
            L.gemm_tma(_p(x), _p(pk_tf), _p(b), _p(y), M, K, N, _p(res), _p(res_scale), pps, _p(stats), stats_act, _stream())
        else:
            L.gemm_px(_p(x), _p(pk_f), STATE['lo_off'], _p(b), _p(y), M, K, N, _p(res), _p(res_scale), pps, _p(stats), stats_act, _stream())
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, res_scale)
        ctx.w, ctx.b, ctx.pk_b, ctx.k0, ctx.has_res = w, b, pk_b, k0, res is not None
        ctx.pk_tb = pk_tb if tc else None
        ctx.mark_non_differentiable(*([stats] if want_stats else []))
        return (y, stats) if want_stats else (y, None)

    @staticmethod
    def backward(ctx, dy, _ds):
        x, res_scale = ctx.saved_tensors
        w, b = ctx.w, ctx.b
        dy = _c(dy)
        K = x.shape[-1]
        N = w.shape[0]
        M = x.numel() // K
        ktot = w.numel() // N
        dres = dy if ctx.has_res else None
        dacc = dy
        if ctx.has_res and res_scale is not None:
            dacc = torch.empty_like(dy)
            L.scale_per_sample(_p(dy), _p(res_scale), _p(dacc), dy.numel(), dy.numel() // dy.shape[0], _stream())
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            if ctx.pk_tb is not None and bool(L.tcct_gemm_tma_supported(M, N, K)):
                L.gemm_tma(_p(dacc), _p(ctx.pk_tb), None, _p(dx), M, N, K, None, None, 0, None, 0, _stream())
            else:
                L.gemm_px(_p(dacc), _p(ctx.pk_b), STATE['lo_off'], None, _p(dx), M, N, K, None, None, 0, None, 0, _stream())
        dw, dwd = _grad_target(w)
        db, dbd = _grad_target(b) if b is not None else (None, True)
        dw_ptr = ctypes.c_void_p(dw.data_ptr() + 4 * ctx.k0)
        with wgrad_side(dwd and dbd, x, dacc):
            if ctx.pk_tb is not None and bool(L.tcct_wgrad_gemm_tma_supported(M, K, N)):
                _wgrad_gemm(x, dacc, dw.data_ptr() + 4 * ctx.k0, db, M, K, N, ktot, dwd and dbd)
            else:
                L.wgrad(_p(x), _p(dacc), dw_ptr, _p(db), 1, 1, M, K, N, 1, 1, ktot, 1, 0, int(STATE['x3']), _stream())
        return dx, _ret(dw, dwd), _ret(db, dbd), None, None, None, dres, None, None, None, None, None


def mlp_fused_supported(M, dim, hidden):
    """Both GEMMs of the MLP, forward and data gradient, on the tcgen05 kernel (csrc/gemm_tma.cu) in the default precision."""
    return (STATE["umma"] and not STATE["x3"] and STATE["fuse_mlp"] and bool(L.tcct_gemm_tma_supported(M, dim, hidden))
            and bool(L.tcct_gemm_tma_supported(M, hidden, dim)))


class MlpFn(torch.autograd.Function):
    """MHCABlock's MLP with its residual (tcct.py:29-53, 467-468):  out = t + scale[b] * (fc2(GELU(fc1(cur)))) as ONE autograd node of
    two GEMM launches forward (fc1 writes the pre-activation and its GELU from the same epilogue; fc2 adds the residual) and two
    backward (the data gradient through fc2 multiplies by GELU' in its epilogue) + the two weight gradients.  The unfused path is
    fc1, a GELU pass, fc2 forward and fc2-dgrad, a GELU' pass, fc1-dgrad backward."""

    @staticmethod
    def forward(ctx, cur, t, scale, fc1, fc2):
        _check(cur, t, scale)
        dim, hid = cur.shape[-1], fc1.weight.shape[0]
        M = cur.numel() // dim
        h = torch.empty(cur.shape[:-1] + (hid,), dtype=torch.float32, device=cur.device)
        g = torch.empty_like(h)
        L.gemm_tma_gelu(_p(cur), _p(fc1.pk_tf[0]), _p(fc1.bias), _p(h), _p(g), M, dim, hid, _stream())
        out = torch.empty_like(t)
        L.gemm_tma(_p(g), _p(fc2.pk_tf[0]), _p(fc2.bias), _p(out), M, hid, dim, _p(t), _p(scale), M // cur.shape[0], None, 0, _stream())
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(cur, h, g, scale)
        ctx.fc1, ctx.fc2 = fc1, fc2
        return out

    @staticmethod
    def backward(ctx, dout):
        cur, h, g, scale = ctx.saved_tensors
        fc1, fc2 = ctx.fc1, ctx.fc2
        dout = _c(dout)
        dim, hid = cur.shape[-1], h.shape[-1]
        M = cur.numel() // dim
        dacc = dout
        if scale is not None:
            dacc = torch.empty_like(dout)
            L.scale_per_sample(_p(dout), _p(scale), _p(dacc), dout.numel(), dout.numel() // dout.shape[0], _stream())
        dh = torch.empty_like(h)
        L.gemm_tma_dgelu(_p(dacc), _p(fc2.pk_tb[0]), _p(h), _p(dh), M, dim, hid, _stream())
        dcur = torch.empty_like(cur)
        L.gemm_tma(_p(dh), _p(fc1.pk_tb[0]), None, _p(dcur), M, hid, dim, None, None, 0, None, 0, _stream())
        for lin, x, dy, K, N in ((fc2, g, dacc, hid, dim), (fc1, cur, dh, dim, hid)):
            dw, dwd = _grad_target(lin.weight)
            db, dbd = _grad_target(lin.bias)
            if not (dwd and dbd):
                raise RuntimeError("MlpFn: the MLP parameters must be registered in a FlatParams buffer")
            with wgrad_side(True, x, dy):
                if bool(L.tcct_wgrad_gemm_tma_supported(M, K, N)):
                    _wgrad_gemm(x, dy, dw.data_ptr(), db, M, K, N, K, True)
                else:
                    L.wgrad(_p(x), _p(dy), _p(dw), _p(db), 1, 1, M, K, N, 1, 1, K, 1, 0, 0, _stream())
        return dcur, dout, None, None, None


# --------------------------------------------------------------------------- batch norm family
def _bn_src(bn, stats, count, training, device):
    """(tcct_bn_src record, coef tensor) for one nn.BatchNorm2d operand of the fused forward."""
    C = bn.weight.numel()
    coef = torch.empty(4 * C, dtype=torch.float32, device=device)
    use_batch = training or bn.running_mean is None
    rec = L.BnSrc()
    rec.stats = stats.data_ptr() if use_batch else None
    rec.count = float(count)
    rec.gamma, rec.beta = bn.weight.data_ptr(), bn.bias.data_ptr()
    rec.eps = float(bn.eps)
    rec.momentum = float(bn.momentum if bn.momentum is not None else 0.1)
    rec.running_mean = bn.running_mean.data_ptr() if bn.running_mean is not None else None
    rec.running_var = bn.running_var.data_ptr() if bn.running_var is not None else None
    rec.num_batches = bn.num_batches_tracked.data_ptr() if bn.num_batches_tracked is not None else None
    rec.update_running = 1 if (training and bn.track_running_stats) else 0
    rec.coef = coef.data_ptr()
    return rec, coef


class BnAct2Fn(torch.autograd.Function):
    """out = post( opA(a) + opB(b) ), op(v) = BN(pre(v)) or pre(v).  bnX None -> no normalisation,
    b None -> single operand.  statsX come from the producing kernel's epilogue; the BatchNorm finalisation
    (coefficients, running statistics) runs in the prologue of the same launch."""

    @staticmethod
    def forward(ctx, a, stats_a, bn_a, pre_a, b, stats_b, bn_b, pre_b, post, training):
        _check(a, b)
        C = a.shape[-1]
        npix = a.numel() // C
        dev = a.device
        rec_a, coef_a = _bn_src(bn_a, stats_a, npix, training, dev) if bn_a is not None else (None, None)
        rec_b, coef_b = _bn_src(bn_b, stats_b, npix, training, dev) if (bn_b is not None and b is not None) else (None, None)
        out = torch.empty_like(a)
        L.bn_act2_fwd_bn(_p(a), ctypes.byref(rec_a) if rec_a is not None else None, pre_a, _p(b),
                         ctypes.byref(rec_b) if rec_b is not None else None, pre_b, post, _p(out), npix, C, _stream())
        ctx.save_for_backward(a, b, coef_a, coef_b)
        ctx.cfg = (bn_a, pre_a, bn_b, pre_b, post, training)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b, coef_a, coef_b = ctx.saved_tensors
        bn_a, pre_a, bn_b, pre_b, post, training = ctx.cfg
        dout = _c(dout)
        if coef_a is None and coef_b is None and pre_a == ACT_NONE and pre_b == ACT_NONE and post == ACT_NONE:
            # plain a (+ b): the gradient passes through unchanged, no kernel
            return dout, None, None, None, (dout if b is not None else None), None, None, None, None, None
        C = a.shape[-1]
        npix = a.numel() // C
        dev = a.device
        sums = ARENA.take(24 * C + 1, dev) if (training and (coef_a is not None or coef_b is not None)) else None   # 8 replicas of the 3 per-channel sums
        da = torch.empty_like(a)
        db = torch.empty_like(b) if b is not None else None
        ga = gb = dga = dba = dgb = dbb = None
        flags = [True] * 4
        if coef_a is not None:
            ga = bn_a.weight
            dga, flags[0] = _grad_target(bn_a.weight)
            dba, flags[1] = _grad_target(bn_a.bias)
        if coef_b is not None:
            gb = bn_b.weight
            dgb, flags[2] = _grad_target(bn_b.weight)
            dbb, flags[3] = _grad_target(bn_b.bias)
        if not all(flags):
            raise RuntimeError("BnAct2Fn: BatchNorm parameters must be registered in a FlatParams buffer")
        L.bn_act2_bwd(_p(a), _p(coef_a), pre_a, _p(ga), _p(b), _p(coef_b), pre_b, _p(gb), post, _p(dout), _p(sums),
                      _p(da), _p(db), _p(dga), _p(dba), _p(dgb), _p(dbb), npix, C, _stream())
        return da, None, None, None, db, None, None, None, None, None


def bn_act2(a, stats_a=None, bn_a=None, pre_a=ACT_NONE, b=None, stats_b=None, bn_b=None, pre_b=ACT_NONE,
            post=ACT_NONE, training=True):
    return BnAct2Fn.apply(a, stats_a, bn_a, pre_a, b, stats_b, bn_b, pre_b, post, training)


class GateFuseFn(torch.autograd.Function):
    """GateFusion (tcct.py:916-932): x1 * a + x2 * (1 - a), a = clamp(bicubic_up(alpha), 0, 1); alpha None -> 0.5 (eval)."""

    @staticmethod
    def forward(ctx, x1, x2, alpha):
        _check(x1, x2, alpha)
        B, H, W, C = x1.shape
        hs, ws = (alpha.shape[2], alpha.shape[3]) if alpha is not None else (0, 0)
        out = torch.empty_like(x1)
        L.gate_fuse_fwd(_p(x1), _p(x2), _p(alpha), _p(out), B, H, W, C, hs, ws, _stream())
        ctx.alpha, ctx.dims = alpha, (B, H, W, C, hs, ws)
        return out

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        d1, d2 = torch.empty_like(dy), torch.empty_like(dy)
        L.gate_fuse_bwd(_p(dy), _p(ctx.alpha), _p(d1), _p(d2), *ctx.dims, _stream())
        return d1, d2, None


# --------------------------------------------------------------------------- pooling / depthwise / norms
class MaxPool2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _check(x)
        B, H, W, C = x.shape
        y = torch.empty((B, H // 2, W // 2, C), dtype=torch.float32, device=x.device)
        L.maxpool2_fwd(_p(x), _p(y), B, H, W, C, _stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        B, H, W, C = x.shape
        dx = torch.empty_like(x)
        L.maxpool2_bwd(_p(x), _p(_c(dy)), _p(dx), B, H, W, C, _stream())
        return dx


class DwConv3Fn(torch.autograd.Function):
    """Depthwise 3x3, pad 1, stride 1|2; add_input: y = dw(x) + bias + x (ConvPosEnc)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, add_input, want_stats):
        _check(x, w, b)
        B, H, W, C = x.shape
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        y = torch.empty((B, Ho, Wo, C), dtype=torch.float32, device=x.device)
        stats = ARENA.take(2 * C, x.device) if want_stats else None
        L.dwconv3_fwd(_p(x), _p(w), _p(b), _p(y), B, H, W, C, stride, int(add_input), _p(stats), _stream())
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x)
        ctx.w, ctx.b, ctx.stride, ctx.add_input = w, b, stride, int(add_input)
        ctx.mark_non_differentiable(*([stats] if want_stats else []))
        return (y, stats) if want_stats else (y, None)

    @staticmethod
    def backward(ctx, dy, _ds):
        (x,) = ctx.saved_tensors
        w, b = ctx.w, ctx.b
        B, H, W, C = x.shape
        dy = _c(dy)
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw, dwd = _grad_target(w)
        db, dbd = _grad_target(b) if b is not None else (None, True)
        # data gradient on the dependent chain, weight gradient on the weight-gradient stream pool (the depthwise weight gradients of
        # the MPViT stages sat on the critical stream of the backward: scripts/trace_step.py)
        if dx is not None:
            L.dwconv3_bwd(_p(x), _p(w), _p(dy), _p(dx), None, None, B, H, W, C, ctx.stride, ctx.add_input, _stream())
        with wgrad_side(dwd and dbd, x, dy):
            L.dwconv3_bwd(_p(x), _p(w), _p(dy), None, _p(dw), _p(db), B, H, W, C, ctx.stride, ctx.add_input, _stream())
        return dx, _ret(dw, dwd), _ret(db, dbd), None, None, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        _check(x, gamma, beta)
        C = x.shape[-1]
        ntok = x.numel() // C
        y = torch.empty_like(x)
        mr = torch.empty(2 * ntok, dtype=torch.float32, device=x.device)
        L.layernorm_fwd(_p(x), _p(gamma), _p(beta), _p(y), _p(mr), ntok, C, float(eps), _stream())
        ctx.save_for_backward(x, mr)
        ctx.gamma, ctx.beta = gamma, beta
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mr = ctx.saved_tensors
        C = x.shape[-1]
        ntok = x.numel() // C
        dx = torch.empty_like(x)
        dg, dgd = _grad_target(ctx.gamma)
        db, dbd = _grad_target(ctx.beta)
        L.layernorm_bwd(_p(x), _p(ctx.gamma), _p(mr), _p(_c(dy)), _p(dx), _p(dg), _p(db), ntok, C, _stream())
        return dx, _ret(dg, dgd), _ret(db, dbd), None


class MetaPoolFn(torch.autograd.Function):
    """out = t + scale[b] * (avgpool3x3_{(token, channel) plane}(cur) - cur)."""

    @staticmethod
    def forward(ctx, t, cur, scale):
        _check(t, cur, scale)
        B, C = t.shape[0], t.shape[-1]
        N = t.numel() // (B * C)
        out = torch.empty_like(t)
        L.metapool_fwd(_p(t), _p(cur), _p(scale), _p(out), B, N, C, _stream())
        ctx.save_for_backward(scale)
        ctx.dims = (B, N, C)
        return out

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        B, N, C = ctx.dims
        dy = _c(dy)
        dcur = torch.empty_like(dy)
        L.metapool_bwd(_p(dy), _p(scale), _p(dcur), B, N, C, _stream())
        return dy, dcur, None


class LnMetaPoolFn(torch.autograd.Function):
    """MHCABlock.forward tcct.py:457-469 up to the MLP: (t2, cur2) = (t + scale[b] * (MetaPool(LN1(t))), LN2(t2)) in one kernel."""

    @staticmethod
    def forward(ctx, t, g1, b1, g2, b2, scale, eps):
        _check(t, g1, b1, g2, b2, scale)
        B, C = t.shape[0], t.shape[-1]
        N = t.numel() // (B * C)
        t2, cur2 = torch.empty_like(t), torch.empty_like(t)
        stats = torch.empty(4 * B * N, dtype=torch.float32, device=t.device)
        L.ln_metapool_fwd(_p(t), _p(g1), _p(b1), _p(g2), _p(b2), _p(scale), _p(t2), _p(cur2), _p(stats), B, N, C, float(eps), _stream())
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(t, t2, stats, scale)
        ctx.params = (g1, b1, g2, b2)
        ctx.dims = (B, N, C)
        return t2, cur2

    @staticmethod
    def backward(ctx, dt2, dcur2):
        t, t2, stats, scale = ctx.saved_tensors
        g1, b1, g2, b2 = ctx.params
        B, N, C = ctx.dims
        dt2 = _c(dt2) if dt2 is not None else None
        dcur2 = _c(dcur2) if dcur2 is not None else None
        dt = torch.empty_like(t)
        targets = [_grad_target(p) for p in (g1, b1, g2, b2)]
        L.ln_metapool_bwd(_p(t), _p(t2), _p(stats), _p(g1), _p(g2), _p(scale), _p(dt2), _p(dcur2), _p(dt),
                          *[_p(tg) for tg, _ in targets], B, N, C, _stream())
        return (dt,) + tuple(_ret(tg, d) for tg, d in targets) + (None, None)


# --------------------------------------------------------------------------- resampling / normalise
class ResizeNHWCFn(torch.autograd.Function):
    """out = alpha * bilinear(x -> [H, W]) (+ add).  align: PyTorch's align_corners."""

    @staticmethod
    def forward(ctx, x, add, H, W, align, alpha):
        _check(x, add)
        B, h, w, C = x.shape
        out = torch.empty((B, H, W, C), dtype=torch.float32, device=x.device)
        L.resize_nhwc_fwd(_p(x), _p(add), _p(out), B, h, w, H, W, C, int(align), float(alpha), 0, _stream())
        ctx.cfg = (B, h, w, H, W, C, int(align), float(alpha), add is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, h, w, H, W, C, align, alpha, has_add = ctx.cfg
        dout = _c(dout)
        dx = torch.empty((B, h, w, C), dtype=torch.float32, device=dout.device)
        L.resize_nhwc_bwd(_p(dout), _p(dx), B, h, w, H, W, C, align, alpha, _stream())
        return dx, (dout if has_add else None), None, None, None, None


class ResizeNCHWFn(torch.autograd.Function):
    """F.interpolate(x, size=(H, W), mode='bilinear', align_corners=False) on NCHW logits."""

    @staticmethod
    def forward(ctx, x, H, W):
        _check(x)
        B, C, h, w = x.shape
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
        L.resize_nchw_fwd(_p(x), _p(out), B * C, h, w, H, W, _stream())
        ctx.cfg = (B, C, h, w, H, W)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, C, h, w, H, W = ctx.cfg
        dx = torch.empty((B, C, h, w), dtype=torch.float32, device=dout.device)
        L.resize_nchw_bwd(_p(_c(dout)), _p(dx), B * C, h, w, H, W, _stream())
        return dx, None, None


class L2Norm32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _check(x)
        if x.shape[-1] != 32:
            raise RuntimeError("l2norm32: 32 channels expected")
        y = torch.empty_like(x)
        L.l2norm32_fwd(_p(x), _p(y), x.numel() // 32, 1.0, _stream())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        L.l2norm32_bwd(_p(x), _p(_c(dy)), _p(dx), x.numel() // 32, 1.0, _stream())
        return dx


class NormAdd3Fn(torch.autograd.Function):
    """norm_add (tcct.py:937-942) of three 32-channel maps: alpha * (normalize(x0) + up(n1) + up(n2)) at the resolution
    of x0 in ONE pass; n1, n2 are already-normalised lower-resolution maps (L2Norm32Fn), bilinear, align_corners=False."""

    @staticmethod
    def forward(ctx, x0, n1, n2, alpha):
        _check(x0, n1, n2)
        B, H, W, C = x0.shape
        if C != 32 or n1.shape[-1] != 32 or n2.shape[-1] != 32:
            raise RuntimeError("norm_add3: 32 channels expected")
        out = torch.empty_like(x0)
        L.norm_add3_fwd(_p(x0), _p(n1), _p(n2), _p(out), B, H, W, n1.shape[1], n1.shape[2], n2.shape[1], n2.shape[2],
                        float(alpha), _stream())
        ctx.save_for_backward(x0)
        ctx.cfg = (tuple(n1.shape), tuple(n2.shape), float(alpha))
        return out

    @staticmethod
    def backward(ctx, dout):
        (x0,) = ctx.saved_tensors
        s1, s2, alpha = ctx.cfg
        dout = _c(dout)
        B, H, W, C = x0.shape
        dx0 = torch.empty_like(x0)
        L.l2norm32_bwd(_p(x0), _p(dout), _p(dx0), x0.numel() // 32, alpha, _stream())
        dn1 = torch.empty(s1, dtype=torch.float32, device=dout.device)
        dn2 = torch.empty(s2, dtype=torch.float32, device=dout.device)
        L.resize_nhwc_bwd(_p(dout), _p(dn1), B, s1[1], s1[2], H, W, 32, 0, alpha, _stream())
        L.resize_nhwc_bwd(_p(dout), _p(dn2), B, s2[1], s2[2], H, W, 32, 0, alpha, _stream())
        return dx0, dn1, dn2, None


# --------------------------------------------------------------------------- stems and heads
class StemConvFn(torch.autograd.Function):
    """3x3 conv 3 -> 32 on the NCHW image, NHWC out (no input gradient: the image is data)."""

    @staticmethod
    def forward(ctx, img, w, b, stride, want_stats):
        _check(img, w, b)
        B, Ci, H, W = img.shape
        if Ci != 3 or tuple(w.shape) != (32, 3, 3, 3):
            raise RuntimeError("stem conv expects a 3-channel image and a [32,3,3,3] weight")
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        y = torch.empty((B, Ho, Wo, 32), dtype=torch.float32, device=img.device)
        stats = ARENA.take(64, img.device) if want_stats else None
        L.stem_conv_fwd(_p(img), _p(w), _p(b), _p(y), B, H, W, stride, _p(stats), _stream())
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(img)
        ctx.w, ctx.b, ctx.stride = w, b, stride
        ctx.mark_non_differentiable(*([stats] if want_stats else []))
        return (y, stats) if want_stats else (y, None)

    @staticmethod
    def backward(ctx, dy, _ds):
        (img,) = ctx.saved_tensors
        B, _, H, W = img.shape
        dw, dwd = _grad_target(ctx.w)
        db, dbd = _grad_target(ctx.b) if ctx.b is not None else (None, True)
        dy = _c(dy)
        with wgrad_side(dwd and dbd, img, dy):
            L.stem_conv_wgrad(_p(img), _p(dy), _p(dw), _p(db), B, H, W, ctx.stride, _stream())
        return None, _ret(dw, dwd), _ret(db, dbd), None, None


class HeadFn(torch.autograd.Function):
    """1x1 conv 32 -> n_class: NHWC features in, NCHW logits out."""

    @staticmethod
    def forward(ctx, x, w, b):
        _check(x, w, b)
        B, H, W, C = x.shape
        Cc = w.shape[0]
        out = torch.empty((B, Cc, H, W), dtype=torch.float32, device=x.device)
        L.head_fwd(_p(x), _p(w), _p(b), _p(out), B, H * W, Cc, _stream())
        ctx.save_for_backward(x)
        ctx.w, ctx.b = w, b
        return out

    @staticmethod
    def backward(ctx, dl):
        (x,) = ctx.saved_tensors
        B, H, W, C = x.shape
        w, b = ctx.w, ctx.b
        dx = torch.empty_like(x)
        dw, dwd = _grad_target(w)
        db, dbd = _grad_target(b)
        L.head_bwd(_p(x), _p(w), _p(_c(dl)), _p(dx), _p(dw), _p(db), B, H * W, w.shape[0], _stream())
        return dx, _ret(dw, dwd), _ret(db, dbd)


# --------------------------------------------------------------------------- losses
def labels_u8(gt, n_class):
    """uint8 [B,H,W] class-index map from what the reference passes around: an int64 one-hot
    [B,C,H,W] (loop_seg.py:119) or an index map [B,H,W]."""
    if gt.dtype == torch.uint8 and gt.dim() == 3:
        return _c(gt)
    _check(gt)
    if gt.dim() == 4 and gt.shape[1] == n_class:
        B, C, H, W = gt.shape
        g64 = gt if gt.dtype == torch.int64 else gt.round().to(torch.int64)
        lab = torch.empty((B, H, W), dtype=torch.uint8, device=gt.device)
        L.onehot_to_index(_p(_c(g64)), _p(lab), B, C, H * W, _stream())
        return lab
    if gt.dim() == 3:
        g64 = _c(gt.to(torch.int64))
        lab = torch.empty(gt.shape, dtype=torch.uint8, device=gt.device)
        L.index64_to_u8(_p(g64), _p(lab), g64.numel(), _stream())
        return lab
    raise RuntimeError("labels: expected one-hot [B,C,H,W] or index [B,H,W], got %s" % (tuple(gt.shape),))


class DiceFn(torch.autograd.Function):
    """MultiLoss(DiceLoss) (mode 0) / MultiLoss(MSELoss) (mode 1) on NCHW logits and a uint8 label map."""

    @staticmethod
    def forward(ctx, logits, lab, mode):
        _check(logits, lab)
        B, C, H, W = logits.shape
        sums = ARENA.take(3 * C + 1, logits.device)
        loss = torch.empty((), dtype=torch.float32, device=logits.device)
        coef = torch.empty(2 * C + 1, dtype=torch.float32, device=logits.device)
        L.dice_fwd(_p(logits), _p(lab), B, C, H * W, mode, _p(sums), _p(loss), _p(coef), _stream())
        ctx.save_for_backward(logits, lab, coef)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, lab, coef = ctx.saved_tensors
        B, C, H, W = logits.shape
        d = torch.empty_like(logits)
        L.dice_bwd(_p(logits), _p(lab), B, C, H * W, _p(coef), _p(_c(g.float())), 1.0, _p(d), 0, _stream())
        return d, None, None


class DiceMultiFn(torch.autograd.Function):
    """KiteBack.grad_calc with ds=True (loopback.py:62-73) for the Dice criterion: crit(z0) + w_aux * sum_k crit(up(z_k)), with the
    auxiliary logits z1..z3 given at their NATIVE resolution (the bilinear up-sampling of tcct.py:1042-1044 happens inside the
    kernel).  Returns (total, [four per-head losses])."""

    @staticmethod
    def forward(ctx, z0, z1, z2, z3, lab, w_aux):
        _check(z0, z1, z2, z3, lab)
        B, C, H, W = z0.shape
        dev = z0.device
        hs = (ctypes.c_int * 3)(z1.shape[2], z2.shape[2], z3.shape[2])
        ws = (ctypes.c_int * 3)(z1.shape[3], z2.shape[3], z3.shape[3])
        wt = (ctypes.c_float * 4)(1.0, float(w_aux), float(w_aux), float(w_aux))
        sums = ARENA.take(int(L.tcct_dice_multi_sums_doubles(C)), dev)
        loss = torch.empty(5, dtype=torch.float32, device=dev)
        coef = torch.empty(8 * C, dtype=torch.float32, device=dev)
        L.dice_multi_fwd(_p(z0), _p(z1), _p(z2), _p(z3), hs, ws, _p(lab), B, C, H, W, wt, _p(sums), _p(loss), _p(coef), _stream())
        ctx.save_for_backward(z0, z1, z2, z3, lab, coef)
        ctx.w_aux = float(w_aux)
        parts = loss[:4]
        ctx.mark_non_differentiable(parts)
        return loss[4], parts

    @staticmethod
    def backward(ctx, g, _gp):
        z0, z1, z2, z3, lab, coef = ctx.saved_tensors
        B, C, H, W = z0.shape
        hs = (ctypes.c_int * 3)(z1.shape[2], z2.shape[2], z3.shape[2])
        ws = (ctypes.c_int * 3)(z1.shape[3], z2.shape[3], z3.shape[3])
        wt = (ctypes.c_float * 4)(1.0, ctx.w_aux, ctx.w_aux, ctx.w_aux)
        d0 = torch.empty_like(z0)
        d1, d2, d3 = torch.zeros_like(z1), torch.zeros_like(z2), torch.zeros_like(z3)
        L.dice_multi_bwd(_p(z0), _p(z1), _p(z2), _p(z3), hs, ws, _p(lab), B, C, H, W, wt, _p(coef), _p(_c(g.float())),
                         _p(d0), _p(d1), _p(d2), _p(d3), _stream())
        return d0, d1, d2, d3, None, None


class BoundaryRegFn(torch.autograd.Function):
    """RegNet.regular_reg (reg.py:109-156).  eps: [2,B,C-1,H,W] uniform(0,1) noise (pred, true);
    jit: [2,H] uniform(0,1) row jitter (pred, true).  Gradients reach logits[:,1:] and, accumulated by the
    kernels, the lap_reg / lap_map parameters of `reg` (the RegNet module)."""

    @staticmethod
    def forward(ctx, logits, lab, eps, jit, reg, training):
        _check(logits, lab, eps, jit)
        B, C, H, W = logits.shape
        dev = logits.device
        n = B * H * W
        ws = torch.empty(int(L.tcct_breg_ws_floats(B, C, H, W)), dtype=torch.float32, device=dev)
        dws = ARENA.take(10, dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        bn = reg.lap_map[1]
        L.breg_forward(_p(logits), _p(lab), _p(eps), _p(jit), _p(reg.lap_reg[0].weight), _p(reg.lap_reg[0].bias),
                       _p(reg.lap_reg[1].weight), _p(reg.lap_reg[1].bias), _p(reg.lap_map[0].weight), _p(reg.lap_map[0].bias),
                       _p(bn.weight), _p(bn.bias), _p(reg.lap_map[2].weight), _p(reg.lap_map[2].bias), _p(bn.running_mean),
                       _p(bn.running_var), _p(bn.num_batches_tracked), int(training), B, C, H, W, _p(ws), _p(dws), _p(loss),
                       _stream())
        ctx.save_for_backward(logits, lab, eps, jit, ws, dws)
        ctx.reg, ctx.training = reg, int(training)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits, lab, eps, jit, ws, dws = ctx.saved_tensors
        reg = ctx.reg
        B, C, H, W = logits.shape
        n = B * H * W
        bws = torch.empty(int(L.tcct_breg_bwd_ws_floats(B, C, H, W)), dtype=torch.float32, device=logits.device)
        bws[6 * n:].zero_()
        dlogits = torch.zeros_like(logits)
        bn = reg.lap_map[1]
        params = [reg.lap_reg[0].weight, reg.lap_reg[0].bias, reg.lap_reg[1].weight, reg.lap_reg[1].bias,
                  reg.lap_map[0].weight, reg.lap_map[0].bias, bn.weight, bn.bias, reg.lap_map[2].weight, reg.lap_map[2].bias]
        targets = [_grad_target(p) for p in params]
        if not all(d for _, d in targets):
            raise RuntimeError("BoundaryRegFn: RegNet parameters must be registered in a FlatParams buffer")
        L.breg_backward(_p(logits), _p(lab), _p(eps), _p(jit), _p(params[0]), _p(params[1]), _p(params[2]), _p(params[3]),
                        _p(params[4]), _p(bn.weight), _p(params[8]), ctx.training, B, C, H, W, _p(ws), _p(dws), _p(bws),
                        _p(_c(g.float())), _p(dlogits), *[_p(t) for t, _ in targets], _stream())
        return dlogits, None, None, None, None, None


class FeaturePolarFn(torch.autograd.Function):
    """RegNet.regular_udh (reg.py:86-105): rank-binned class prototypes of the 32-d decoder features against
    fixed per-class targets.  feat: NHWC [B,H,W,32]; logits are detached as in the reference."""

    @staticmethod
    def forward(ctx, feat, logits, lab, proto):
        _check(feat, logits, lab, proto)
        B, C, H, W = logits.shape
        if feat.shape != (B, H, W, 32):
            raise RuntimeError("feature_polar: feat must be NHWC [B,H,W,32], got %s" % (tuple(feat.shape),))
        dev = logits.device
        n = B * H * W
        words = int(L.tcct_fpolar_ws_words(n))
        iws = torch.empty(words, dtype=torch.int32, device=dev)
        iws[:16].zero_()
        fws = torch.zeros(int(L.tcct_fpolar_fws_bytes()) // 4, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        pro_last = torch.empty(1024, dtype=torch.float32, device=dev)
        L.fpolar_forward(_p(feat), _p(logits), _p(lab), _p(proto), B, C, H, W, _p(iws), _p(fws), _p(loss), _p(pro_last), _stream())
        ctx.save_for_backward(lab, proto, pro_last, iws)
        ctx.dims = (B, C, H, W)
        return loss

    @staticmethod
    def backward(ctx, g):
        lab, proto, pro_last, iws = ctx.saved_tensors
        B, C, H, W = ctx.dims
        dfeat = torch.empty((B, H, W, 32), dtype=torch.float32, device=lab.device)
        L.fpolar_backward(_p(lab), _p(proto), _p(pro_last), B, C, H, W, _p(iws), _p(_c(g.float())), _p(dfeat), _stream())
        return dfeat, None, None, None
