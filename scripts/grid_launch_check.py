#!/usr/bin/env python
"""Systematic launch check: every (n_fft, n_mels, hop) of a grid through every entry path with a tiny batch — forward,
forward-mode forward + dot-product backward, recomputing backward with and without dL/dx, time-major, int16 input — looking
for launch failures and non-finite results (shared-memory thresholds, plan edge cases).  Found the 48 KB opt-in bug."""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dmel_b200 as dm

dev = torch.device("cuda:0")
bad = n = 0
for n_fft, M, hop in itertools.product((64, 128, 256, 512, 1024, 2048, 4096, 8192), (0, 8, 16, 24, 32, 40, 48, 64, 80, 96, 112, 128),
                                       (1, 20, 40, 80, 160, 441, 512)):
    lam = 0.75 * n_fft / 6
    N = max(4 * hop, min(3 * n_fft, 9000))
    if N // hop + 1 > 1500:
        continue
    n += 1
    tag = f"n_fft={n_fft} M={M} hop={hop} N={N}"
    try:
        g = torch.Generator().manual_seed(n)
        x = torch.randn(2, N, generator=g).to(dev)
        kw = dict(hop_length=hop, optimized=True, n_mels=M, sample_rate=16000, log_eps=1e-10)
        for need_dx in (True, False):
            xin = x.clone().requires_grad_(need_dx)
            ll = torch.tensor(lam, device=dev, requires_grad=True)
            y = dm.dmel(xin, ll, **kw)
            y.sum().backward()
            assert torch.isfinite(y).all() and torch.isfinite(ll.grad), "non-finite"
        dm.op.USE_JVP = False
        try:
            ll = torch.tensor(lam, device=dev, requires_grad=True)
            dm.dmel(x, ll, **kw).sum().backward()
        finally:
            dm.op.USE_JVP = True
        if dm.load().dmel_time_major_supported(n_fft, M):
            yt = dm.dmel(x, torch.tensor(lam, device=dev), time_major=True, **kw)
            assert torch.equal(yt, dm.dmel(x, torch.tensor(lam, device=dev), **kw).transpose(2, 3).contiguous()), "time-major differs"
        xi = (x * 1000).to(torch.int16)
        yi = dm.dmel(xi, torch.tensor(lam, device=dev), **kw)
        assert torch.isfinite(yi).all(), "int16 non-finite"
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        bad += 1
        print("FAIL", tag, "->", str(e)[:150], flush=True)
        torch.cuda.synchronize()
print(f"grid: {n - bad} of {n} geometries launch and agree", flush=True)
sys.exit(1 if bad else 0)
