"""torchrun --nproc-per-node N scripts/peer_ring_test.py : correctness + latency of the peer-memory all-reduce vs NCCL."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import dmel_b200 as dm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ring = dm.PeerRing.create(dev)
if rank == 0:
    print("ring:", "ok" if ring is not None else "UNAVAILABLE", flush=True)
if ring is not None:
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    bad = 0
    for i in range(300):
        v = torch.randn(1, device=dev, generator=g)
        want = v.clone(); dist.all_reduce(want)
        got = ring.all_reduce_(v.clone())
        if abs(float(got) - float(want)) > 1e-5 * max(1.0, abs(float(want))):
            bad += 1
    # bit-identical on all ranks
    allv = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(allv, got)
    same = all(torch.equal(allv[0], a) for a in allv)
    # latency, eager and under a CUDA graph
    v = torch.ones(1, device=dev)
    def timeit(fn, n=2000):
        for _ in range(20): fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3
    t_ring = timeit(lambda: ring.all_reduce_(v))
    t_nccl = timeit(lambda: dist.all_reduce(v))
    v.fill_(1.0)
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): ring.all_reduce_(v)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize(); dist.barrier()
    v.fill_(1.0)
    with torch.cuda.graph(gr):
        for _ in range(10): ring.all_reduce_(v)
    torch.cuda.synchronize(); dist.barrier()
    v.fill_(1.0)
    gr.replay(); torch.cuda.synchronize()
    graph_ok = abs(float(v) - float(world) ** 10) <= 1e-3 * float(world) ** 10
    t_graph = timeit(lambda: gr.replay(), n=200) / 10
    if rank == 0:
        print(f"world={world} mismatches={bad} identical_on_all_ranks={same} graph_ok={graph_ok} timed_out={ring.timed_out()} "
              f"us/allreduce: ring eager {t_ring:.1f}, ring in graph {t_graph:.1f}, nccl eager {t_nccl:.1f}", flush=True)
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
