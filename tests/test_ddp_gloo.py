"""CPU, world size 2 over gloo: the data-parallel plumbing of the training loop (tcct_b200/kite/ddp.py) -- replica
broadcast, the single flat-buffer gradient all-reduce, the 1/world averaging done by the optimizer, per-rank data
shards -- checked against the definition: averaged gradient == mean of the per-shard gradients, and replicas stay
bit-identical after a step."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _adamw_ref(p, g, m, v, step, lr, wd=2e-4, b1=0.9, b2=0.999, eps=1e-8, max_norm=12.0, grad_scale=1.0):
    """What csrc/optim.cu does on the flat buffers: scale, clip by the global norm, decoupled weight decay, AdamW."""
    g = g * grad_scale
    norm = g.norm()
    g = g * min(1.0, max_norm / (float(norm) + 1e-6))
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    mh, vh = m / (1 - b1 ** step), v / (1 - b2 ** step)
    p.mul_(1 - lr * wd).addcdiv_(mh, vh.sqrt() + eps, value=-lr)
    return float(norm)


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from tcct_b200.kite import ddp
    from tcct_b200.synth import make_bscans
    r, w, _ = ddp.init("gloo")
    assert (r, w) == (rank, world)
    n, n_active = 1000, 900
    g0 = torch.Generator().manual_seed(100 + rank)                 # replicas start DIFFERENT on purpose
    flat = torch.randn(n, generator=g0)
    running = torch.randn(8, generator=g0)
    ddp.broadcast_replica(flat, [running])
    ref = torch.randn(n, generator=torch.Generator().manual_seed(100))
    assert torch.equal(flat, ref), "rank %d: parameters differ from rank 0 after the broadcast" % rank
    # per-rank shard of the data: different seeds -> different batches
    img, lab = make_bscans(2, 32, 32, 5, 4, ddp.shard_seed(1234, rank))
    shard_grad = torch.zeros(n)
    shard_grad[:n_active] = torch.linspace(0, 1, n_active) * float(img.mean()) + float(lab.float().mean())
    local = shard_grad.clone()
    ddp.allreduce_flat(shard_grad, n_active)
    gathered = [torch.zeros(n) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert not torch.equal(gathered[0], gathered[1]), "shards must differ"
    assert torch.allclose(shard_grad[:n_active], sum(gathered)[:n_active], rtol=0, atol=1e-6)
    assert torch.equal(shard_grad[n_active:], torch.zeros(n - n_active)), "untrained tail must not be reduced"
    # optimizer step with grad_scale = 1/world == step on the mean gradient; replicas stay identical
    m, v = torch.zeros(n_active), torch.zeros(n_active)
    p = flat[:n_active].clone()
    _adamw_ref(p, shard_grad[:n_active], m, v, 1, 1e-3, grad_scale=1.0 / world)
    p2, m2, v2 = flat[:n_active].clone(), torch.zeros(n_active), torch.zeros(n_active)
    _adamw_ref(p2, torch.stack(gathered).mean(0)[:n_active], m2, v2, 1, 1e-3)
    assert torch.allclose(p, p2, rtol=0, atol=1e-7)
    every = [torch.zeros(n_active) for _ in range(world)]
    dist.all_gather(every, p)
    assert torch.equal(every[0], every[1]), "replicas diverged after the step"
    # BatchNorm running statistics drift per rank (own batches); before validation / saving they become the mean over the ranks
    bn = torch.nn.BatchNorm2d(4)
    bn.running_mean.fill_(float(rank + 1))
    bn.running_var.fill_(float(2 * rank + 1))
    ddp.average_buffers(bn)
    assert torch.equal(bn.running_mean, torch.full((4,), 1.5)) and torch.equal(bn.running_var, torch.full((4,), 2.0))
    assert int(bn.num_batches_tracked) == 0
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert sorted(out.get(timeout=5) for _ in range(2)) == [0, 1]


def test_single_process_is_a_noop():
    sys.path.insert(0, ROOT)
    from tcct_b200.kite import ddp
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    assert ddp.env_world() == (0, 1, 0)
    g = torch.arange(10.0)
    assert torch.equal(ddp.allreduce_flat(g.clone(), 5), g)
