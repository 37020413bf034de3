"""Data-parallel plumbing: one process per GPU, batch sharded, ONE collective per step.

The reference is single-process (no tf.distribute / Horovod anywhere).  Samples are independent units (SURVEY.md
section 8e): every rank runs the same step on its own 1/R of the batch and the two flat gradient buffers are averaged
with an all-reduce before the (replicated) clip + momentum update.  Mean-of-means is exact because all ranks
hold the same number of meshes.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (MASTER_ADDR should be 127.0.0.1 on one node)."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def make_allreduce(world):
    """Returns f(grad_g, grad_d) that averages the flat gradient buffers over ranks in place (None if world == 1)."""
    if world <= 1:
        return None
    use_avg = dist.get_backend() == "nccl"

    def allreduce(*bufs):
        for b in bufs:
            if use_avg:
                dist.all_reduce(b, op=dist.ReduceOp.AVG)
            else:                                   # gloo has no AVG
                dist.all_reduce(b, op=dist.ReduceOp.SUM)
                b.div_(world)

    return allreduce


def shard_indices(n_total, rank, world):
    """Contiguous, equal shards of a global batch; the global batch must divide evenly (static per-GPU batch)."""
    if n_total % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (n_total, world))
    per = n_total // world
    return np.arange(rank * per, (rank + 1) * per)


def rank_seed(seed, rank):
    """eps of vae_sampling and the batch indices must differ per rank."""
    return int(seed) + 1000003 * int(rank)


def broadcast_params(flat_buffers, src=0):
    """Make replicas identical at start (rank 0's initial weights)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for b in flat_buffers:
            dist.broadcast(b, src)
