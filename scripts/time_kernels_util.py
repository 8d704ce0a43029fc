import torch

_SIDE = []


def side_stream():
    if not _SIDE:
        _SIDE.append(torch.cuda.Stream())
    return _SIDE[0]


def timeit(fn, reps=12, replays=3):
    """Device time per call in us: `reps` calls captured into one CUDA graph (no host launch gaps), replayed."""
    # warm-up on a side stream (torch's CUDA-graph recipe: nothing autograd creates lazily may be bound to the legacy stream)
    if not _SIDE:
        _SIDE.append(torch.cuda.Stream())
    side = _SIDE[0]          # one stream for every warm-up AND every capture: autograd binds its accumulators to it
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (replays * reps) * 1e3
