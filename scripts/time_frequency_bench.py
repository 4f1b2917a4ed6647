"""BASELINE configs[0] (search space 'time_frequency'): SpectrogramLayer, optimized=False, hop 1, trainable lambda.
GPU forward+backward (dx + dlambda) through the public module vs the CPU oracle port on the same tensors."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dmel_b200 as dm
from oracle import dmel_oracle as orc

dev = torch.device("cuda:0")
for (B, N) in ((64, 256), (128, 128)):
    for lam in (0.2 * 6.38, 6.38, 5 * 6.38):
        g = torch.Generator().manual_seed(0)
        x = torch.randn(B, N, generator=g, dtype=torch.float64)
        G = torch.randn(B, 1, N + 1, N + 1, generator=g)
        layer = dm.SpectrogramLayer(torch.tensor(lam), device="cuda:0", optimized=False, size=(N + 1, N + 1), hop_length=1).to(dev)
        xd, Gd = x.to(dev).float(), G.to(dev)
        def step():
            xin = xd.detach().requires_grad_(True)
            layer.lambd.grad = None
            layer(xin).backward(Gd)
        for _ in range(10): step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(100): step()
        b.record(); torch.cuda.synchronize()
        gpu_ms = a.elapsed_time(b) / 100
        torch.set_num_threads(os.cpu_count())
        def cpu():
            xx = x.float().clone().requires_grad_(True)
            ll = torch.tensor(lam, requires_grad=True)
            s = orc.port_forward(xx, ll, hop=1, optimized=False)
            (s * G).sum().backward()
        cpu(); t = time.perf_counter(); cpu(); cpu_ms = (time.perf_counter() - t) * 1e3
        print(json.dumps({"config": "time_frequency", "B": B, "N": N, "n_fft": 2 * N, "lambda": round(lam, 3), "gpu_ms_fwd_bwd": round(gpu_ms, 4),
                          "cpu_port_ms_fwd_bwd": round(cpu_ms, 1), "cores": os.cpu_count(), "clips_per_s_gpu": round(B / gpu_ms * 1e3)}), flush=True)
