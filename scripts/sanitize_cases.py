"""Small forward+backward launches over every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dmel_b200 as dm

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
cases = [  # (N, sr, hop, M, lam, optimized)
    (700, 8000, 80, 16, 13.33, True),     # 128
    (700, 8000, 80, 16, 30.0, True),      # 256
    (900, 8000, 80, 64, 46.667, True),    # 512
    (1500, 16000, 160, 32, 93.33, True),  # 1024
    (2000, 44100, 441, 32, 257.25, True), # 2048
    (1200, 8000, 80, 16, 400.0, True),    # 4096
    (1500, 16000, 160, 16, 800.0, True),  # 8192 (big)
    (300, 8000, 40, 8, 8.0, True),        # 64 -> DFT path
    (64, 8000, 1, 0, 6.38, False),        # spectrogram, non-optimized (n_fft 128)
    (600, 8000, 50, 0, 30.0, True),       # spectrogram, optimized 256
]
only = os.environ.get("DMEL_SAN_CASES")
for ci, (N, sr, hop, M, lam, opt) in enumerate(cases):
    if only and str(ci) not in only.split(","):
        continue
    x = torch.randn(3, N, generator=g).to(dev).requires_grad_(True)
    ll = torch.tensor(lam, device=dev, requires_grad=True)
    y = dm.dmel(x, ll, hop_length=hop, optimized=opt, n_mels=M, sample_rate=sr, log_eps=1e-10)
    y.backward(torch.randn(y.shape, generator=g).to(dev))
    x2 = x.detach()
    y2 = dm.dmel(x2, ll, hop_length=hop, optimized=opt, n_mels=M, sample_rate=sr)   # lambda-only backward, linear
    y2.backward(torch.ones_like(y2))
    torch.cuda.synchronize()
    print("case", ci, tuple(y.shape), "ok", flush=True)
# round 2: 16-bit samples widened in the kernels, the fixed-lambda grid in one call (shared means), the tensor-core
# DFT with a mel bank and in spectrogram mode over several tiles, forward-mode forward at n_fft 2048 (one group per CTA)
xi = (torch.randn(3, 1600, generator=g) * 3000).to(torch.int16).to(dev)
ll = torch.tensor(46.667, device=dev, requires_grad=True)
y = dm.dmel(xi, ll, hop_length=80, optimized=True, n_mels=32, sample_rate=8000, log_eps=1e-10)
y.backward(torch.randn(y.shape, generator=g).to(dev))
ys = dm.dmel_multi(torch.randn(3, 1600, generator=g).to(dev), [8.0, 13.33, 46.667, 400.0], hop_length=80, optimized=True, n_mels=32,
                   sample_rate=8000, log_eps=1e-10)
ym = dm.dmel(torch.randn(5, 4000, generator=g).to(dev), torch.tensor(8.0, device=dev), hop_length=20, optimized=True, n_mels=24,
             sample_rate=8000, log_eps=1e-10)
ysp = dm.dmel(torch.randn(5, 4000, generator=g).to(dev), torch.tensor(4.0, device=dev), hop_length=10, optimized=True, n_mels=0)
ll2 = torch.tensor(257.25, device=dev, requires_grad=True)
yj = dm.dmel(torch.randn(2, 6000, generator=g).to(dev), ll2, hop_length=441, optimized=True, n_mels=32, sample_rate=44100, log_eps=1e-10)
yj.backward(torch.randn(yj.shape, generator=g).to(dev))
torch.cuda.synchronize()
print("round-2 cases ok", flush=True)
