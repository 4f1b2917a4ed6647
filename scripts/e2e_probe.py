"""Where does the end-to-end step go?  H2D bandwidth, host-side cost of one layer forward+backward, device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dmel_b200 as dm

dev = torch.device("cuda:0")
B, N, sr, hop, M, lam = 256, 8000, 8000, 80, 64, 8000 * 0.035 / 6
T = N // hop + 1
hx = torch.randn(B, N).pin_memory()
dx_ = torch.empty(B, N, device=dev)
G = torch.randn(B, 1, M, T, device=dev)
layer = dm.MelSpectrogramLayer(torch.tensor(lam), n_mels=M, n_points=N, sample_rate=sr, hop_length=hop, device="cuda:0",
                               optimized=True, log_eps=1e-10).to(dev)
def sync(): torch.cuda.synchronize()
# 1. H2D
for _ in range(5): dx_.copy_(hx, non_blocking=True)
sync(); t = time.perf_counter()
for _ in range(200): dx_.copy_(hx, non_blocking=True)
sync(); dt = (time.perf_counter() - t) / 200
print(f"H2D {hx.numel()*4/1e6:.1f} MB: {dt*1e6:.1f} us  = {hx.numel()*4/dt/1e9:.1f} GB/s")
# 2. host cost of fwd+bwd (enqueue only; queue depth kept shallow by syncing every 20)
def step():
    xin = dx_.detach().requires_grad_(True)
    layer.lambd.grad = None
    y = layer(xin)
    y.backward(G)
for _ in range(20): step()
sync(); t = time.perf_counter()
n = 400
for i in range(n): step()
t_enq = (time.perf_counter() - t) / n
sync(); t_tot = (time.perf_counter() - t) / n
print(f"fwd+bwd through nn.Module: host enqueue {t_enq*1e6:.1f} us/step, wall {t_tot*1e6:.1f} us/step")
# 3. forward only / backward only host cost
xin = dx_.detach().requires_grad_(True)
sync(); t = time.perf_counter()
for i in range(n): y = layer(xin)
t_f = (time.perf_counter() - t) / n; sync()
print(f"forward only host enqueue {t_f*1e6:.1f} us")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(200): step()
pr.disable(); sync()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

# 4. H2D of one clip batch: one copy on one stream vs two half copies on two streams (two copy engines)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
hx2 = [torch.randn(B, N).pin_memory() for _ in range(3)]
half = B // 2
def copy_one(i):
    with torch.cuda.stream(s1):
        dx_.copy_(hx2[i % 3], non_blocking=True)
def copy_split(i):
    h = hx2[i % 3]
    with torch.cuda.stream(s1):
        dx_[:half].copy_(h[:half], non_blocking=True)
    with torch.cuda.stream(s2):
        dx_[half:].copy_(h[half:], non_blocking=True)
def copy_alt(i):
    with torch.cuda.stream(s1 if i % 2 == 0 else s2):
        dx_.copy_(hx2[i % 3], non_blocking=True)
for name, fn in (("one stream", copy_one), ("two halves on two streams", copy_split), ("alternating streams", copy_alt)):
    for i in range(10): fn(i)
    sync(); t = time.perf_counter()
    for i in range(300): fn(i)
    sync(); dt = (time.perf_counter() - t) / 300
    print(f"H2D 8.2 MB, {name}: {dt*1e6:.1f} us/batch = {B*N*4/dt/1e9:.1f} GB/s")
for mb in (1, 4, 16, 64):
    hb = torch.empty(mb * 2**20 // 4).pin_memory(); db = torch.empty_like(hb, device=dev)
    for _ in range(5): db.copy_(hb, non_blocking=True)
    sync(); t = time.perf_counter()
    for _ in range(100): db.copy_(hb, non_blocking=True)
    sync(); dt = (time.perf_counter() - t) / 100
    print(f"H2D {mb} MiB back-to-back: {dt*1e6:.1f} us = {mb*2**20/dt/1e9:.1f} GB/s")
