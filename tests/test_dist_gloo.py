"""world_size-2 gloo test of the data-parallel host logic (dist.py): shard the clips, each rank
back-propagates its shard, ONE all-reduce sums dL/dlambda -> equals the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dmel_b200 as dm
from oracle import dmel_oracle as orc


def test_shard_bounds_partition():
    for n, w in [(256, 8), (10, 4), (3, 8), (1000, 7), (8192, 8)]:
        spans = [dm.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        dm.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, x, G, lam, kw, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = dm.shard_bounds(x.shape[0], world, rank)
        p = torch.nn.Parameter(torch.tensor(lam if rank == 0 else 123.0))
        dm.broadcast_lambda([p])                      # replicas must agree on lambda (-> n_fft)
        s = orc.port_forward_batched(x[lo:hi], p, **kw)
        (torch.log(s + orc.LOG_EPS) * G[lo:hi]).sum().backward()
        dm.allreduce_lambda_grad([p])
        ret[rank] = (float(p.detach()), float(p.grad))
    finally:
        dist.destroy_process_group()


def test_two_rank_lambda_grad_equals_single_process():
    torch.manual_seed(0)
    kw = dict(hop=80, optimized=True, norm=False, n_mels=32, sample_rate=8000)
    x = torch.randn(6, 1600)
    lam = 46.667
    G = torch.randn(6, 1, 32, 21)
    _, _, want = orc.port_forward_backward(x, lam, G, use_log=True, batched=True, **kw)
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, x, G, lam, kw, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert abs(ret[0][0] - lam) < 1e-4 and abs(ret[1][0] - lam) < 1e-4
    assert ret[0][1] == ret[1][1]
    assert abs(ret[0][1] - want) <= 1e-4 * abs(want)
