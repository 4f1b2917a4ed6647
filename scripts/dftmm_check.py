#!/usr/bin/env python
"""First-light check of the tcgen05 DFT-by-matrix-multiply forward (n_fft <= 128) against the oracle, then its timing
against the CUDA-core kernels (DMEL_NO_DFTMM=1 in a child process).  Run under `timeout`: a wrong barrier hangs."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import dmel_b200 as dm
from oracle import dmel_oracle as orc

def check(tag, B, N, sr, hop, M, lam, optimized, log):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, N, generator=g)
    dev = torch.device("cuda:0")
    eps = 1e-10 if log else None
    y = dm.dmel(x.to(dev), torch.tensor(float(lam), device=dev), hop_length=hop, optimized=optimized, n_mels=M,
                sample_rate=sr, log_eps=eps)
    torch.cuda.synchronize()
    want = orc.port_forward(x, torch.tensor(float(lam)), hop=hop, optimized=optimized, norm=False, n_mels=M, sample_rate=sr)
    if log:
        want = torch.log(want + 1e-10)
    got, want = y.cpu().numpy(), want.numpy()
    atol = 1e-5 if log else 1e-5 * float(np.abs(want).max())
    err = np.abs(got - want)
    ok = bool(np.all(err <= atol + 1e-4 * np.abs(want)))
    print(f"{tag}: n_fft={orc.window_length(lam, N, optimized)[1]} shape={got.shape} max|err|={err.max():.3e} "
          f"max|want|={np.abs(want).max():.3e} {'PASS' if ok else 'FAIL'}", flush=True)
    return ok

def timing():
    dev = torch.device("cuda:0")
    N, sr, hop, M = 8000, 8000, 80, 64
    out = {}
    for B, n_fft in ((512, 128), (512, 64), (4096, 128), (4096, 64), (256, 128)):
        lam = 0.75 * n_fft / 6
        x = torch.randn(B, N, device=dev)
        ll = torch.tensor(float(lam), device=dev)
        for _ in range(5):
            dm.dmel(x, ll, hop_length=hop, optimized=True, n_mels=M, sample_rate=sr, log_eps=1e-10, lambd_value=lam)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            dm.dmel(x, ll, hop_length=hop, optimized=True, n_mels=M, sample_rate=sr, log_eps=1e-10, lambd_value=lam)
        e1.record()
        torch.cuda.synchronize()
        out[f"B{B}_L{n_fft}"] = round(e0.elapsed_time(e1) / 200, 5)
    print("TIMING", json.dumps(out), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        timing()
        sys.exit(0)
    ok = True
    ok &= check("mel128_log", 3, 8000, 8000, 80, 64, 16.0, True, True)
    ok &= check("mel128_lin", 2, 4000, 8000, 80, 64, 16.0, True, False)
    ok &= check("spec128", 2, 2000, 8000, 40, 0, 16.0, True, False)
    ok &= check("mel64", 2, 4000, 8000, 40, 32, 8.0, True, True)
    ok &= check("mel100_nonopt", 2, 4000, 8000, 40, 40, 100 / 6 + 0.05, False, True)
    ok &= check("spec36_nonopt", 2, 1000, 8000, 7, 0, 6.1, False, False)
    ok &= check("many_tiles", 40, 8000, 8000, 80, 64, 16.0, True, True)
    print("ALL PASS" if ok else "SOME FAIL", flush=True)
    for env in ({}, {"DMEL_NO_DFTMM": "1"}):
        e = dict(os.environ); e.update(env)
        print("env", env, flush=True)
        subprocess.run([sys.executable, __file__, "time"], env=e, timeout=120)
