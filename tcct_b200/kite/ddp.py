"""Data-parallel plumbing of the training loop (one process per GPU).  The reference is single-GPU
(task1/kite/loopback.py:130-139 ignores `--pl`); what a DistributedDataParallel wrap of it would do is restated over
the flat parameter / gradient buffers: replicas are made identical once, every step sums the flat gradient with ONE
all-reduce and the optimizer divides by the world size.  Backend-agnostic (nccl on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def env_world():
    """(rank, world, local_rank) from the torchrun environment; (0, 1, 0) when not launched by it."""
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    return 0, 1, 0


def init(backend="nccl", device=None):
    rank, world, local = env_world()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local


def broadcast_replica(flat_buf, buffers=()):
    """Make every rank's parameters (one flat tensor) and non-parameter state (BN running statistics, prototypes)
    equal to rank 0's."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_buf, 0)
        for b in buffers:
            dist.broadcast(b, 0)


def average_buffers(module):
    """Before validation / a checkpoint: the BatchNorm running statistics, which every rank updates from its own batches, become
    the mean over the ranks (one all-reduce of the concatenated float buffers), so that the replica rank 0 saves is the job's and
    not one shard's.  No-op for a single process."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return
    bufs = [b for n, b in module.named_buffers() if b.dtype.is_floating_point and n.rsplit(".", 1)[-1] in ("running_mean", "running_var")]
    if not bufs:
        return
    flat = torch.cat([b.reshape(-1) for b in bufs])
    dist.all_reduce(flat)
    flat /= dist.get_world_size()
    off = 0
    for b in bufs:
        b.copy_(flat[off:off + b.numel()].view_as(b))
        off += b.numel()


def allreduce_flat(grad, n_active):
    """Sum the trained prefix of the flat gradient buffer over the ranks (in place).  The mean is taken by the
    optimizer (`grad_scale = 1 / world`), so no extra pass over the buffer is needed."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(grad[:n_active])
    return grad


def shard_seed(base_seed, rank):
    """Every rank draws its own batches (weak scaling: the per-GPU batch is fixed)."""
    return base_seed + rank
