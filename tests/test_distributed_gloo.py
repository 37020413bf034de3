"""World-size-2 gloo test of the data-parallel host logic (runs on CPU)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from cape_b200 import distributed as D
    r, w, _ = D.init("gloo")
    assert (r, w) == (rank, world)
    # per-rank "gradients" of a per-rank batch mean; the all-reduce must give the global-batch mean
    rng = np.random.RandomState(D.rank_seed(7, rank))
    per_sample = torch.from_numpy(rng.normal(size=(4, 10)).astype(np.float32))
    gg = per_sample.mean(0).clone()
    gd = torch.full((3,), float(rank + 1))
    D.make_allreduce(world)(gg, gd)
    allv = [torch.zeros_like(per_sample) for _ in range(world)]
    dist.all_gather(allv, per_sample)
    want = torch.cat(allv).mean(0)
    w0 = torch.full((5,), float(rank))
    D.broadcast_params([w0])
    q.put((rank, float((gg - want).abs().max()), gd.tolist(), w0.tolist(), D.shard_indices(8, rank, world).tolist()))
    dist.destroy_process_group()


def test_gradient_allreduce_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, gd, w0, shard in res:
        assert err < 1e-6                                  # mean of per-rank means == global mean
        assert gd == [1.5, 1.5, 1.5]
        assert w0 == [0.0] * 5                             # replicas start from rank 0's weights
        assert shard == list(range(rank * 4, rank * 4 + 4))


def test_shard_requires_even_split():
    import pytest
    from cape_b200 import distributed as D
    with pytest.raises(ValueError):
        D.shard_indices(7, 0, 2)
    assert D.rank_seed(123, 0) != D.rank_seed(123, 1)
    assert D.make_allreduce(1) is None
