"""Tile-shape sweep on the GPU box: times the fwd / bwd kernels (CUDA events inside the library) for every
(NG, SF) the library accepts, per workload.  Usage: python scripts/tune.py audio_mnist esc50 ..."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import dmel_b200 as dm
from importlib import import_module


def geo_big(w):
    return w["n_fft"] >= 8192


def run(wname, steps=30):
    w = bench.workload(wname)
    dev = torch.device("cuda:0")
    lib = dm.load()
    B, N, sr, hop, M, lam = w["B"], w["N"], w["sr"], w["hop"], w["M"], w["lam"]
    if wname == "large_batch":
        B = 256
    if geo_big(w):
        B = 16
    T = N // hop + 1
    geo = dm.resolve_geometry(B, N, hop, lam, True, M)
    bank = import_module("differentiable-mel-spectrogram_b200.melbank").device_bank(geo.n_freqs, 0.0, sr // 2, M, sr, dev)
    R = 3
    xs = [torch.randn(B, N, device=dev) for _ in range(R)]
    Gs = [torch.randn(B, 1, M, T, device=dev) for _ in range(R)]
    y = torch.empty(B, 1, M, T, device=dev)
    dx = torch.empty(B, N, device=dev)
    dl = torch.zeros((), device=dev)
    wsb = lib.dmel_workspace_bytes(B, geo.n_fft)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    st = int(torch.cuda.current_stream(dev).cuda_stream)
    res = []
    ngs = {128: (8, 4), 256: (8, 4), 512: (8, 4), 1024: (4, 2), 2048: (2, 1), 4096: (1,), 8192: (1,), 16384: (1,)}[geo.n_fft]
    default_only = os.environ.get("TUNE_DEFAULT") == "1"
    for ng in ((0,) if default_only else ngs):
        for sf in ((0,) if (geo_big(w) or default_only) else (2, 4, 8, 16, 32)):
            if sf:
                os.environ["DMEL_NG"], os.environ["DMEL_SF"] = str(ng), str(sf)
            def step(i, need_dx=True):
                rc = lib.dmel_forward_f32(xs[i % R].data_ptr(), B, N, N, lam, geo.win_length, geo.n_fft, hop, M, bank.ref(), 1e-10, 0,
                                          y.data_ptr(), ws.data_ptr(), wsb, st)
                if rc: return rc
                return lib.dmel_backward_f32(xs[i % R].data_ptr(), B, N, N, lam, geo.win_length, geo.n_fft, hop, M, bank.ref(), 1e-10, 0,
                                             y.data_ptr(), Gs[i % R].data_ptr(), dx.data_ptr() if need_dx else None, dl.data_ptr(),
                                             ws.data_ptr(), wsb, 1, st)
            if step(0):
                continue
            torch.cuda.synchronize()
            out = {"ng": ng, "sf": sf}
            for need_dx, tag in ((True, "bwd"), (False, "bwd_lam")):
                for i in range(3):
                    step(i, need_dx)
                lib.dmel_profile_enable(1)
                for i in range(steps):
                    step(i, need_dx)
                torch.cuda.synchronize()
                for kind, name in ((1, "fwd"), (2, tag)):
                    tot, cnt = ctypes.c_double(), ctypes.c_longlong()
                    lib.dmel_profile_read(kind, ctypes.byref(tot), ctypes.byref(cnt))
                    out[name] = round(1e3 * tot.value / max(cnt.value, 1), 2)
                lib.dmel_profile_enable(0)
            res.append(out)
            print(wname, json.dumps(out), flush=True)
    os.environ.pop("DMEL_NG", None); os.environ.pop("DMEL_SF", None)
    return res


if __name__ == "__main__":
    for wn in sys.argv[1:] or ["audio_mnist"]:
        run(wn)
