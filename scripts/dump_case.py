"""Write DMELCASE fixture files for tests/c/abi_run.cu (the torch-free C host of the C ABI).

Expected arrays come from the oracle (oracle/dmel_oracle.py, the CPU restatement of models.py:33-56 /
time_frequency.py:21-58 pinned to the unmodified reference by tests/golden), the sparse filterbank from the product's
own packer (melbank.pack_bank) applied to the oracle's dense matrix.  Runs on the CPU.

    python scripts/dump_case.py            -> tests/golden/cases/*.bin (small parity cases, committed)
                                              tests/golden/cases/timing/*.bin (geometry-only cases for `abi_run time`)
"""
import os
import struct
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dmel_b200  # noqa: E402
from oracle import dmel_oracle as orc  # noqa: E402

FLAG_NORM = 1


def write_case(path, *, B, N, sr, hop, M, lam, opt=True, norm=False, use_log=True, expected=True, seed=0):
    win, n_fft = orc.window_length(abs(lam), N, opt)
    F = n_fft // 2 + 1
    T = N // hop + 1
    flags = FLAG_NORM if norm else 0
    log_eps = 1e-10 if use_log else -1.0
    arrays = []
    mel_nnz = bin_nnz = 0
    if M:
        fb = orc.mel_filterbank(F, 0.0, sr // 2, M, sr).numpy()
        p = dmel_b200.pack_bank(fb)
        mel_nnz, bin_nnz = int(p.mel_off[-1]), int(p.bin_off[-1])
        arrays = [p.mel_k0.astype(np.int32), p.mel_off.astype(np.int32), p.mel_w[:mel_nnz].astype(np.float32),
                  p.bin_m0.astype(np.int32), p.bin_off.astype(np.int32), p.bin_w[:bin_nnz].astype(np.float32)]
    rows = M if M else F
    exp = []
    if expected:
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(B, N, generator=g)
        G = torch.randn(B, 1, rows, T, generator=g)
        y, dx, dl = orc.port_forward_backward(x, lam, G, use_log=use_log, batched=True, hop=hop, optimized=opt, norm=norm,
                                              n_mels=M, sample_rate=sr)
        exp = [x.numpy().astype(np.float32), G.numpy().astype(np.float32), y.astype(np.float32), dx.astype(np.float32),
               np.array([dl], np.float32)]
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"DMELCASE")
        f.write(struct.pack("<i", 1))
        f.write(struct.pack("<11i", B, N, hop, win, n_fft, M, flags, F, mel_nnz, bin_nnz, 1 if expected else 0))
        f.write(struct.pack("<2f", lam, log_eps))
        for a in arrays + exp:
            f.write(np.ascontiguousarray(a).tobytes())
    print(f"{path}: B={B} N={N} n_fft={n_fft} hop={hop} M={M} expected={expected} ({os.path.getsize(path)} bytes)")


def main():
    gold = os.path.join(ROOT, "tests", "golden", "cases")
    # small parity cases (committed): audio_mnist-shaped, large_batch-shaped (n_fft 1024), a spectrogram-mode one
    write_case(os.path.join(gold, "am_L512.bin"), B=3, N=8000, sr=8000, hop=80, M=64, lam=8000 * 0.035 / 6)
    write_case(os.path.join(gold, "lb_L1024.bin"), B=2, N=16000, sr=16000, hop=160, M=128, lam=16000 * 0.035 / 6, seed=1)
    write_case(os.path.join(gold, "spec_L256_lin.bin"), B=2, N=600, sr=0, hop=50, M=0, lam=30.0, use_log=False, seed=2)
    write_case(os.path.join(gold, "esc_L2048_norm.bin"), B=1, N=22050, sr=44100, hop=441, M=128, lam=-44100 * 0.035 / 6,
               norm=True, seed=3)
    # geometry-only cases for timing (not committed)
    tim = os.path.join(gold, "timing")
    write_case(os.path.join(tim, "audio_mnist.bin"), B=1, N=8000, sr=8000, hop=80, M=64, lam=8000 * 0.035 / 6, expected=False)
    write_case(os.path.join(tim, "large_batch.bin"), B=1, N=160000, sr=16000, hop=160, M=128, lam=16000 * 0.035 / 6,
               expected=False)
    write_case(os.path.join(tim, "esc50.bin"), B=1, N=220500, sr=44100, hop=441, M=128, lam=44100 * 0.035 / 6, expected=False)
    write_case(os.path.join(tim, "sweep256.bin"), B=1, N=8000, sr=8000, hop=80, M=64, lam=0.75 * 256 / 6, expected=False)
    write_case(os.path.join(tim, "sweep4096.bin"), B=1, N=8000, sr=8000, hop=80, M=64, lam=0.75 * 4096 / 6, expected=False)


if __name__ == "__main__":
    main()
