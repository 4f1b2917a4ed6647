"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (it imports /root/reference, which does not exist on the GPU
box):   PYTHONDONTWRITEBYTECODE=1 python scripts/gen_golden.py
Each fixture holds the inputs (x, lambd, upstream gradient G) and what the reference's own
`models.MelSpectrogramLayer` / `models.SpectrogramLayer` (+ `torch.log(s + 1e-10)`, models.py:72-73)
and autograd produce for them, in fp32 (the parity target) and in fp64 (the tie-breaker).
The reference ships no golden vectors of its own, so these are the pin.
"""
import os
import sys
import warnings

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = os.environ.get("DMEL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import models  # noqa: E402  (the reference)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
EPS = 1e-10

# name, kind, kwargs
CASES = [
    # --- mel, optimized (the audio search spaces, search_spaces.py:3-68) ---
    dict(name="mel_L512_am", kind="mel", B=2, N=1200, sr=8000, hop=80, M=64, lam=8000 * 0.035 / 6, opt=True, norm=False),
    dict(name="mel_L128_zero_cols", kind="mel", B=2, N=1000, sr=8000, hop=80, M=64, lam=8000 * 0.010 / 6, opt=True, norm=False),
    dict(name="mel_L256", kind="mel", B=3, N=800, sr=8000, hop=80, M=16, lam=23.0, opt=True, norm=False),
    dict(name="mel_L512_norm_neglam", kind="mel", B=2, N=900, sr=8000, hop=80, M=40, lam=-50.0, opt=True, norm=True),
    dict(name="mel_L1024_16k", kind="mel", B=1, N=3200, sr=16000, hop=160, M=128, lam=16000 * 0.035 / 6, opt=True, norm=False),
    dict(name="mel_L2048_44k", kind="mel", B=1, N=4410, sr=44100, hop=441, M=128, lam=44100 * 0.035 / 6, opt=True, norm=False),
    dict(name="mel_L4096", kind="mel", B=1, N=1600, sr=8000, hop=80, M=64, lam=400.0, opt=True, norm=False),
    dict(name="mel_L64_dft", kind="mel", B=2, N=400, sr=8000, hop=40, M=12, lam=8.0, opt=True, norm=False),
    dict(name="mel_L32_dft_norm", kind="mel", B=2, N=203, sr=8000, hop=7, M=6, lam=4.0, opt=True, norm=True),
    dict(name="mel_nonopt_N256", kind="mel", B=2, N=256, sr=8000, hop=16, M=32, lam=20.0, opt=False, norm=False),
    dict(name="mel_odd_hop_short", kind="mel", B=2, N=333, sr=8000, hop=37, M=20, lam=30.0, opt=True, norm=False),
    # --- spectrogram, non-optimized, hop 1, fp64 inputs (search_spaces.py:70-98, datasets.py:33) ---
    dict(name="spec_tf_N128_f64in", kind="spec", B=1, N=128, hop=1, lam=6.38, opt=False, norm=False, x64=True),
    dict(name="spec_tf_N64_small_lam", kind="spec", B=2, N=64, hop=1, lam=0.2 * 6.38, opt=False, norm=False),
    dict(name="spec_tf_N64_big_lam_norm", kind="spec", B=2, N=64, hop=1, lam=5 * 6.38, opt=False, norm=True),
    dict(name="spec_nonpow2_N100_dft", kind="spec", B=2, N=100, hop=5, lam=9.0, opt=False, norm=False),
    dict(name="spec_opt_L256", kind="spec", B=2, N=600, hop=50, lam=30.0, opt=True, norm=False),
    # --- big transforms: 300 ms windows (search_spaces.py:29 grid at 16 / 44.1 kHz) ---
    dict(name="mel_L8192_16k", kind="mel", B=1, N=2400, sr=16000, hop=160, M=64, lam=16000 * 0.300 / 6, opt=True, norm=False),
    dict(name="mel_L16384_44k", kind="mel", B=1, N=4410, sr=44100, hop=441, M=128, lam=44100 * 0.300 / 6, opt=True, norm=False),
    dict(name="spec_opt_L8192_norm", kind="spec", B=1, N=1600, hop=400, lam=-900.0, opt=True, norm=True),
]


def build_layer(c, dtype):
    lam = torch.tensor(c["lam"], dtype=dtype)
    if c["kind"] == "mel":
        return models.MelSpectrogramLayer(lam, n_mels=c["M"], n_points=c["N"], sample_rate=c["sr"], hop_length=c["hop"],
                                          device="cpu", optimized=c["opt"], normalize_window=c["norm"])
    # optimized SpectrogramLayer needs `size` == (F, T) of the result
    size = (512, 1024)
    if c["opt"]:
        import time_frequency as tf
        L = tf.next_power_of_2(np.float32(abs(c["lam"])) * np.float32(6))
        size = (L // 2 + 1, c["N"] // c["hop"] + 1)
    return models.SpectrogramLayer(lam, device="cpu", optimized=c["opt"], size=size, hop_length=c["hop"],
                                   normalize_window=c["norm"])


def run(c, x, G, dtype, use_log):
    layer = build_layer(c, dtype)
    xx = x.to(dtype).clone().requires_grad_(True)
    s = layer(xx)                      # (B,1,rows,T) float32 buffer (models.py:36)
    out = torch.log(s + EPS) if use_log else s
    loss = (out.to(dtype) * G.to(dtype)).sum()
    loss.backward()
    return s.detach(), out.detach(), xx.grad.detach(), layer.lambd.grad.detach()


def main():
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    for ci, c in enumerate(CASES):
        if only and c["name"] not in only:
            continue
        g = torch.Generator().manual_seed(1000 + ci)
        x = torch.randn(c["B"], c["N"], generator=g, dtype=torch.float64)
        x = x + 0.3 * torch.randn(c["B"], 1, generator=g, dtype=torch.float64)  # non-zero means
        in_dtype = torch.float64 if c.get("x64") else torch.float32
        x = x.to(in_dtype)
        layer = build_layer(c, torch.float32)
        with torch.no_grad():
            shape = layer(x.to(torch.float32)).shape
        G = torch.randn(shape, generator=g, dtype=torch.float64).to(torch.float32)
        rec = dict(x=x.numpy(), G=G.numpy(), lam=np.float64(c["lam"]))
        for tag, use_log in (("lin", False), ("log", True)):
            # fp32 = the reference exactly as its training loop runs it (fp32 parameter);
            # x stays in its dataset dtype (fp64 for the synthetic task)
            s32, o32, dx32, dl32 = run(c, x, G, torch.float32 if not c.get("x64") else torch.float64, use_log)
            if c.get("x64"):
                # reference behaviour for fp64 data: fp32 lambd parameter, fp64 signal
                layer = build_layer(c, torch.float32)
                xx = x.clone().requires_grad_(True)
                s = layer(xx)
                out = torch.log(s + EPS) if use_log else s
                (out * G).sum().backward()
                s32, o32, dx32, dl32 = s.detach(), out.detach(), xx.grad.detach(), layer.lambd.grad.detach()
            s64, o64, dx64, dl64 = run(c, x, G, torch.float64, use_log)
            rec[f"y_{tag}"] = o32.numpy().astype(np.float32)
            rec[f"dx_{tag}"] = dx32.numpy().astype(np.float32)
            rec[f"dlam_{tag}"] = np.float64(dl32.item())
            rec[f"y64_{tag}"] = o64.numpy().astype(np.float64)
            rec[f"dx64_{tag}"] = dx64.numpy().astype(np.float64)
            rec[f"dlam64_{tag}"] = np.float64(dl64.item())
        meta = {k: v for k, v in c.items() if k not in ("name",)}
        rec["meta"] = np.array(repr(meta))
        path = os.path.join(OUT, c["name"] + ".npz")
        np.savez_compressed(path, **rec)
        print(f"{c['name']:28s} y{tuple(shape)}  dlam32={rec['dlam_log']:+.6e} dlam64={rec['dlam64_log']:+.6e} "
              f"{os.path.getsize(path) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
