"""Host mirror of the reference's time_frequency.py public functions (time_frequency.py:21-65).

`differentiable_spectrogram` is what produce_figures.py:52-58 calls directly; it runs on the
same fused kernels as the layers (batch of one, no mean removal — the reference removes the
mean in the layer, models.py:38, not here).  `differentiable_gaussian_window` is a handful of
elementwise torch ops on L points; it is only used for plotting / `return_window=True`.
"""
from __future__ import annotations

import torch

from .op import dmel, next_power_of_2, shift_bit_length, window_length_from_lambda  # noqa: F401


def differentiable_gaussian_window(lambd, window_length, device='cpu', norm=True):
    """w[m] = exp(-0.5 ((m - L/2) / (lambd + 1e-15))^2), optionally / sqrt(sum w^2)."""
    m = torch.arange(0, window_length, dtype=torch.float32, device=device)
    w = torch.exp(-0.5 * torch.pow((m - window_length / 2) / (lambd + 1e-15), 2))
    if norm:
        return w / torch.sqrt(torch.sum(torch.pow(w, 2)))
    return w


def differentiable_spectrogram(x, lambd, optimized=False, device='cpu', hop_length=1, return_window=False,
                               norm=False, n_stds=6):
    """|STFT|^2 of one signal x (n_points,) with the Gaussian window of width lambd -> (F, T)."""
    if x.dim() != 1:
        raise ValueError("x must be one signal of shape (n_points,)")
    lam = lambd if torch.is_tensor(lambd) else torch.tensor(float(lambd))
    lam = lam.to(x.device)
    # the reference takes lambd as-is here (|.| is applied by the layers); the window only
    # depends on lambd^2, so passing it through the op's |lambd| changes nothing but the sign
    # convention of d/dlambd for negative values, which abs() restores.
    s = dmel(x.unsqueeze(0), lam, hop_length=hop_length, optimized=optimized, normalize_window=norm,
             n_mels=0, remove_mean=False, n_stds=n_stds)[0, 0]
    if not return_window:
        return s
    win_len = window_length_from_lambda(float(lam.detach()), n_stds) if optimized else x.shape[0]
    return s, differentiable_gaussian_window(lam, win_len, device=x.device, norm=norm)
