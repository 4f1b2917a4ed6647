"""Host-side cost of one forward+backward through the module API, piece by piece (no profiler: perf_counter only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dmel_b200 as dm
from importlib import import_module
op = import_module("differentiable-mel-spectrogram_b200.op")

dev = torch.device("cuda:0")
B, N, sr, hop, M, lam = 256, 8000, 8000, 80, 64, 8000 * 0.035 / 6
T = N // hop + 1
xd = torch.randn(B, N, device=dev)
G = torch.randn(B, 1, M, T, device=dev)
layer = dm.MelSpectrogramLayer(torch.tensor(lam), n_mels=M, n_points=N, sample_rate=sr, hop_length=hop, device="cuda:0",
                               optimized=True, log_eps=1e-10).to(dev)
acc = {}
def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t; return r
    return w
real_empty = torch.empty
lib = dm.load()
class TorchShim:
    def __getattr__(self, k): return getattr(torch, k)
shim = TorchShim(); shim.empty = timed("torch.empty", real_empty)
op.torch = shim
for nm in ("dmel_forward_f32", "dmel_forward_jvp_f32", "dmel_backward_f32", "dmel_lambda_grad_f32"):
    setattr(lib, nm, timed(nm, getattr(lib, nm)))
orig_bwd = op.DmelFunction.backward
def step(need_dx):
    xin = xd.detach().requires_grad_(need_dx)
    layer.lambd.grad = None
    t0 = time.perf_counter(); y = layer(xin); t1 = time.perf_counter(); y.backward(G); t2 = time.perf_counter()
    acc["layer()"] = acc.get("layer()", 0.0) + t1 - t0
    acc["backward()"] = acc.get("backward()", 0.0) + t2 - t1
for need_dx in (True, False):
    for _ in range(30): step(need_dx)
    torch.cuda.synchronize(); acc.clear()
    n = 300
    t = time.perf_counter()
    for i in range(n):
        step(need_dx)
        if i % 8 == 7: torch.cuda.synchronize()       # keep the launch queue shallow: pure host cost
    tot = (time.perf_counter() - t) / n
    print(f"need_dx={need_dx}: {tot*1e6:.1f} us/step incl. syncs; " + ", ".join(f"{k} {v/n*1e6:.1f}" for k, v in sorted(acc.items())))
# allocator alone
for shape in ((B, N), (B, 1, M, T), ()):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(2000):
        a = real_empty(shape, dtype=torch.float32, device=dev)
    print(f"torch.empty{shape}: {(time.perf_counter()-t)/2000*1e6:.2f} us")
