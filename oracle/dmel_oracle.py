"""CPU oracle for the DMEL hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this module, and only as the checker / the timed CPU baseline.  Nothing under
differentiable-mel-spectrogram_b200/ imports it, and the product path has no CPU fallback.

What it restates (all citations into /root/reference):
  models.py:33-56     MelSpectrogramLayer.forward: per-sample loop, mean removal (:38), |lambd| (:38),
                      mel filterbank per sample (:42-51), matmul (:53), pack into (B,1,M,T) fp32 (:36,:54)
  models.py:182-200   SpectrogramLayer.forward: same without the mel steps
  time_frequency.py:21-30  differentiable_gaussian_window (centre L/2, +1e-15, optional /sqrt(sum w^2))
  time_frequency.py:32-58  differentiable_spectrogram (window length, torch.stft conventions, |.|^2)
  time_frequency.py:60-65  next_power_of_2 / shift_bit_length
  models.py:72-73     torch.log(s + 1e-10) applied by the classifier heads
and, for gradients, either torch autograd over that restatement (`port_*`, what the reference itself
relies on, train.py:47) or the closed forms of SURVEY.md §3.3 in numpy fp64 (`closed_form`).

Third-party pieces the reference calls that are NOT in /root/reference: torch.stft (torch==1.13.1
pinned, requirements.txt:7; 2.11 here — center=True, pad_mode='constant', onesided, unnormalised
semantics unchanged) and torchaudio.functional.melscale_fbanks (torchaudio==0.13.1 pinned; HTK scale,
norm=None), restated in `mel_filterbank` below.

Pinning: tests/test_oracle_golden.py checks every function here against tests/golden/*.npz, which
scripts/gen_golden.py produced by running the unmodified reference (the reference has no tests or
golden vectors of its own — SURVEY.md §4 — so those outputs are the pin).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

LOG_EPS = 1e-10  # models.py:73


# ----------------------------------------------------------------------------- host integers
def next_power_of_2(x) -> int:
    """time_frequency.py:60-65."""
    xi = int(x)
    return 1 << (xi - 1).bit_length()


def window_length(lam_abs: float, n_points: int, optimized: bool) -> Tuple[int, int]:
    """(win_length, n_fft): time_frequency.py:34-41 and the n_fft argument of :48/:51."""
    if optimized:
        win = next_power_of_2(np.float32(lam_abs) * np.float32(6))
        return win, win
    return n_points, 2 * n_points


# ----------------------------------------------------------------------------- mel filterbank
def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                   dtype=torch.float32) -> torch.Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale='htk') as called at models.py:42-48;
    the reference then casts it to the spectrogram dtype (:51)."""
    hz = torch.linspace(0, sample_rate // 2, n_freqs)
    to_mel = lambda f: 2595.0 * math.log10(1.0 + f / 700.0)  # noqa: E731
    pts = torch.linspace(to_mel(f_min), to_mel(f_max), n_mels + 2)
    pts_hz = 700.0 * (10.0 ** (pts / 2595.0) - 1.0)
    gaps = pts_hz[1:] - pts_hz[:-1]
    delta = pts_hz[None, :] - hz[:, None]
    lower = -delta[:, :-2] / gaps[:-1]
    upper = delta[:, 2:] / gaps[1:]
    return torch.clamp(torch.minimum(lower, upper), min=0.0).to(dtype)


# ----------------------------------------------------------------------------- torch port (autograd)
def gaussian_window(lam_abs: torch.Tensor, win_length: int, norm: bool) -> torch.Tensor:
    """time_frequency.py:21-30 — `m` is an fp32 arange; the result takes lam's dtype."""
    m = torch.arange(0, win_length).float()
    w = torch.exp(-0.5 * torch.pow((m - win_length / 2) / (lam_abs + 1e-15), 2))
    if norm:
        return w / torch.sqrt(torch.sum(torch.pow(w, 2)))
    return w


def _spectrogram_1d(sig: torch.Tensor, lam_abs: torch.Tensor, optimized: bool, hop: int, norm: bool) -> torch.Tensor:
    """time_frequency.py:32-58 for one centred signal -> (F, T) power."""
    win, n_fft = window_length(float(lam_abs.detach()), sig.shape[-1], optimized)
    w = gaussian_window(lam_abs, win, norm)
    if w.dtype != sig.dtype:
        # torch.stft needs matching dtypes; the reference's fp64 synthetic data meets an fp32 window,
        # which torch promotes inside the window multiply (complex128 result) — same values.
        w = w.to(sig.dtype)
    s = torch.stft(sig, n_fft=n_fft, hop_length=hop, win_length=win, window=w, return_complex=True,
                   pad_mode="constant")
    return torch.pow(torch.abs(s), 2)


def port_forward(x: torch.Tensor, lambd: torch.Tensor, *, hop: int, optimized: bool, norm: bool = False,
                 n_mels: int = 0, sample_rate: int = 0, f_min: float = 0.0, f_max: Optional[float] = None,
                 remove_mean: bool = True) -> torch.Tensor:
    """The reference's forward, sample by sample like models.py:37-54 (this is the structure the CPU
    baseline is timed on: serial loop, filterbank rebuilt per sample).  Returns (B,1,rows,T) float32."""
    B, N = x.shape
    T = N // hop + 1
    fmax = f_max if f_max is not None else (sample_rate // 2 if n_mels else None)
    out = None
    for i in range(B):
        sig = x[i] - torch.mean(x[i]) if remove_mean else x[i]
        spec = _spectrogram_1d(sig, torch.abs(lambd), optimized, hop, norm)
        if n_mels:
            fb = mel_filterbank(spec.shape[0], f_min, fmax, n_mels, sample_rate, dtype=spec.dtype)
            spec = torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)
        if out is None:
            out = torch.empty((B, 1, spec.shape[0], T), dtype=torch.float32)
        out[i, :, :, :] = spec.unsqueeze(0)
    return out


def port_forward_batched(x: torch.Tensor, lambd: torch.Tensor, *, hop: int, optimized: bool, norm: bool = False,
                         n_mels: int = 0, sample_rate: int = 0, f_min: float = 0.0,
                         f_max: Optional[float] = None, remove_mean: bool = True,
                         out_dtype=torch.float32) -> torch.Tensor:
    """Same mathematics with the batch handled by one torch.stft call (fast checker for big cases)."""
    B, N = x.shape
    lam_abs = torch.abs(lambd)
    sig = x - x.mean(dim=1, keepdim=True) if remove_mean else x
    win, n_fft = window_length(float(lam_abs.detach()), N, optimized)
    w = gaussian_window(lam_abs, win, norm).to(sig.dtype)
    s = torch.stft(sig, n_fft=n_fft, hop_length=hop, win_length=win, window=w, return_complex=True,
                   pad_mode="constant")
    p = s.real ** 2 + s.imag ** 2                       # (B, F, T)
    if n_mels:
        fmax = f_max if f_max is not None else sample_rate // 2
        fb = mel_filterbank(p.shape[1], f_min, fmax, n_mels, sample_rate, dtype=p.dtype)
        p = torch.einsum("bft,fm->bmt", p, fb)
    return p.unsqueeze(1).to(out_dtype)


def port_forward_backward(x, lam, G, *, use_log: bool, dtype=torch.float32, batched: bool = False, **kw):
    """Forward + autograd backward of loss = sum(G * out); returns (out, dx, dlam) as numpy."""
    xx = torch.as_tensor(x).to(dtype).clone().requires_grad_(True)
    ll = torch.tensor(float(lam), dtype=dtype, requires_grad=True)
    fwd = port_forward_batched if batched else port_forward
    s = fwd(xx, ll, **kw)
    out = torch.log(s + LOG_EPS) if use_log else s
    (out.to(dtype) * torch.as_tensor(G).to(dtype)).sum().backward()
    return out.detach().numpy(), xx.grad.detach().numpy(), float(ll.grad)


# ----------------------------------------------------------------------------- numpy fp64 closed form
def closed_form(x: np.ndarray, lam: float, G: Optional[np.ndarray], *, hop: int, optimized: bool,
                norm: bool = False, n_mels: int = 0, sample_rate: int = 0, f_min: float = 0.0,
                f_max: Optional[float] = None, use_log: bool = False, remove_mean: bool = True):
    """Forward and the explicit adjoint in fp64 with numpy only.

    y[b,0,m,t] = sum_k fb[k,m] |rfft(frame_t * w)[k]|^2;   with G = dL/d(out):
      Gm = G/(mel+eps) (log) or G;  GP = fb @ Gm;  Z = 2 GP X;  u = Re sum_{k<=n_fft/2} Z[k] e^{+2 pi i k n/n_fft}
      dxc[t*hop+n] += u_t[n] w[n] (crop the centre padding);  dx = dxc - mean(dxc)
      dlam = sign(lam) * sum_t sum_n u_t[n] frame_t[n] dw/d|lam|[n]
    Returns (out, dx, dlam); dx/dlam are None when G is None.
    """
    x = np.asarray(x, dtype=np.float64)
    B, N = x.shape
    la = abs(float(lam))
    win, n_fft = window_length(la, N, optimized)
    F = n_fft // 2 + 1
    T = N // hop + 1
    m = np.arange(win, dtype=np.float64)
    d = m - win / 2
    den = la + 1e-15
    w0 = np.exp(-0.5 * (d / den) ** 2)
    dw0 = w0 * d * d / den ** 3
    if norm:
        nrm = np.sqrt(np.sum(w0 * w0))
        wv = w0 / nrm
        dwv = dw0 / nrm - w0 * np.sum(w0 * dw0) / nrm ** 3
    else:
        wv, dwv = w0, dw0
    left = (n_fft - win) // 2
    w = np.zeros(n_fft)
    dw = np.zeros(n_fft)
    w[left:left + win] = wv
    dw[left:left + win] = dwv
    xc = x - x.mean(axis=1, keepdims=True) if remove_mean else x
    pad = n_fft // 2
    xp = np.pad(xc, ((0, 0), (pad, pad)))
    idx = (np.arange(T) * hop)[:, None] + np.arange(n_fft)[None, :]
    frames = xp[:, idx]                                # (B, T, n_fft)
    X = np.fft.rfft(frames * w, axis=-1)               # (B, T, F)
    P = X.real ** 2 + X.imag ** 2
    if n_mels:
        fmax = f_max if f_max is not None else sample_rate // 2
        fb = mel_filterbank(F, f_min, fmax, n_mels, sample_rate).numpy().astype(np.float64)
        mel = P @ fb                                   # (B, T, M)
    else:
        fb = None
        mel = P
    lin = np.transpose(mel, (0, 2, 1))[:, None]        # (B,1,rows,T)
    out = np.log(lin + LOG_EPS) if use_log else lin
    if G is None:
        return out, None, None
    Gt = np.transpose(np.asarray(G, dtype=np.float64)[:, 0], (0, 2, 1))   # (B,T,rows)
    Gm = Gt / (mel + LOG_EPS) if use_log else Gt
    GP = Gm @ fb.T if fb is not None else Gm                             # (B,T,F)
    Z = 2.0 * GP * X
    scale = np.full(F, 0.5)
    scale[0] = 1.0
    if n_fft % 2 == 0:
        scale[-1] = 1.0
    # Re sum_{k<=n_fft/2} Z e^{+i..} == n_fft * irfft(Z * c), c = 1 at DC/Nyquist (imag dropped), 1/2 elsewhere
    u = np.fft.irfft(Z * scale, n=n_fft, axis=-1) * n_fft               # (B,T,n_fft)
    dxp = np.zeros_like(xp)
    np.add.at(dxp, (np.arange(B)[:, None, None], idx[None, :, :]), u * w)
    dxc = dxp[:, pad:pad + N]
    dx = dxc - dxc.mean(axis=1, keepdims=True) if remove_mean else dxc
    dlam = float(np.sign(lam)) * float(np.sum(u * frames * dw))
    return out, dx, dlam
