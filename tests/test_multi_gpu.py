"""Needs >= 2 GPUs on one box (skipped otherwise): the peer-memory all-reduce and the fused backward+all-reduce
must equal the single-process result on the concatenated batch."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist
    import dmel_b200 as dm
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ring = None
    try:
        ring = dm.PeerRing.create(dev)
        assert ring is not None, "peer ring self-test failed"
        g = torch.Generator().manual_seed(5)
        B, N, M, hop, sr, lam = 8, 4000, 32, 80, 8000, 46.667
        x = torch.randn(B, N, generator=g)
        G = torch.randn(B, 1, M, N // hop + 1, generator=g)
        lo, hi = dm.shard_bounds(B, world, rank)
        layer = dm.MelSpectrogramLayer(torch.tensor(lam), n_mels=M, n_points=N, sample_rate=sr, hop_length=hop,
                                       device=str(dev), optimized=True, log_eps=1e-10).to(dev)
        out = {}
        for tag, need_dx in (("full", True), ("jvp", False)):
            layer.lambd.grad = None
            xin = x[lo:hi].to(dev).requires_grad_(need_dx)
            (layer(xin) * G[lo:hi].to(dev)).sum().backward()
            dm.allreduce_lambda_grad([layer.lambd], ring=ring)
            out[tag] = float(layer.lambd.grad)
        torch.cuda.synchronize()
        assert not ring.timed_out()
        ret[rank] = out
    finally:
        # orderly teardown (the order bench.py uses): drain the device, agree that nobody is still inside a ring
        # kernel, unmap the peers' buffers, then take the process group down
        torch.cuda.synchronize()
        dist.barrier()
        if ring is not None:
            ring.close()
        torch.cuda.synchronize()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_ring_gradient_equals_single_process():
    import torch.multiprocessing as mp
    import dmel_b200 as dm
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    B, N, M, hop, sr, lam = 8, 4000, 32, 80, 8000, 46.667
    x = torch.randn(B, N, generator=g)
    G = torch.randn(B, 1, M, N // hop + 1, generator=g)
    layer = dm.MelSpectrogramLayer(torch.tensor(lam), n_mels=M, n_points=N, sample_rate=sr, hop_length=hop, device="cuda:0",
                                   optimized=True, log_eps=1e-10).to(dev)
    xin = x.to(dev).requires_grad_(True)
    (layer(xin) * G.to(dev)).sum().backward()
    whole = float(layer.lambd.grad)
    for r in range(world):
        assert ret[r]["full"] == ret[0]["full"] and ret[r]["jvp"] == ret[0]["jvp"]     # bit-identical on all ranks
        assert abs(ret[r]["full"] - whole) <= 1e-5 * abs(whole)
        assert abs(ret[r]["jvp"] - whole) <= 2e-4 * abs(whole)
