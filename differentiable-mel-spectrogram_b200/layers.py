"""Drop-in replacements for the reference's differentiable spectrogram layers.

Same constructor arguments, attributes, parameter name (`lambd`, so `spectrogram_layer.lambd`
keeps matching main.py:39 and state_dicts load, utils.py:269-270), forward signature and
output contract as
    models.py:14-56   MelSpectrogramLayer   (B, n_points) -> (B, 1, n_mels, n_points//hop + 1)
    models.py:171-200 SpectrogramLayer      (B, n_points) -> (B, 1, n_fft/2 + 1, n_points//hop + 1)
but one fused CUDA forward (and one fused backward) per batch instead of a Python loop of
~200 ATen calls and two host syncs per sample.  One keyword is appended: `log_eps` fuses the
`torch.log(s + 1e-10)` every classifier head applies (models.py:72-73); its default (None)
reproduces the reference's linear output.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .op import dmel, dmel_multi, workspace_energy


class _LambdaMirror:
    """Host copy of `lambd`, refreshed only when the parameter was written (optimizer.step(), fill_,
    load_state_dict all bump the tensor's version counter).  Removes the device->host sync that picks
    n_fft from every forward whose lambda did not change (validation loops, frozen layers, benchmarks);
    the reference pays that sync once per sample (time_frequency.py:39)."""

    def __init__(self):
        self._key = None
        self._value = 0.0

    def invalidate(self) -> None:
        self._key = None

    def get(self, lambd: torch.Tensor) -> float:
        key = (lambd._version, lambd.data_ptr(), lambd.device)
        if key != self._key:
            self._value = float(lambd.detach().to(torch.float32).item())
            self._key = key
        return self._value


def _resolve_device(module_device, lambd: torch.Tensor) -> torch.device:
    # The reference keeps a `device` attribute and also follows `net.to(device)` (main.py:25).
    # The parameter's device is the truth once the module has been moved.
    return lambd.device


class MelSpectrogramLayer(nn.Module):
    def __init__(self, init_lambd, n_mels, n_points, sample_rate, f_min=0, f_max=None, hop_length=1, device='cpu',
                 optimized=False, normalize_window=False, *, log_eps: Optional[float] = None, time_major: bool = False):
        super().__init__()
        self.hop_length = hop_length
        self.lambd = nn.Parameter(init_lambd)
        self.device = device
        self.optimized = optimized
        self.normalize_window = normalize_window

        self.f_min = f_min
        self.f_max = f_max if f_max is not None else sample_rate // 2
        self.n_mels = n_mels
        self.sample_rate = sample_rate

        self.n_freq = n_mels
        self.n_time = n_points // hop_length + 1
        self.log_eps = log_eps
        self.time_major = time_major      # True: (B, 1, n_time, n_mels), what MelPANNsNet feeds Cnn6 (models.py:162)
        self._mirror = _LambdaMirror()
        self._last_ws = None              # workspace of the last forward: holds the fused energy sum
        # optional per-mel output map (set_output_affine / fold_batchnorm); plain attributes, not buffers, so the
        # state_dict stays exactly {'lambd'} (utils.py:269-270)
        self._out_affine = None

    def set_output_affine(self, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor] = None) -> None:
        """Fold y <- scale[m]*y + shift[m] (per mel bin, after the optional log) into the kernel epilogue; `None`
        removes it.  The vectors are constants: no gradient flows to them."""
        if scale is None:
            self._out_affine = None
            return
        scale = scale.detach().to(torch.float32).reshape(-1).clone()
        shift = torch.zeros_like(scale) if shift is None else shift.detach().to(torch.float32).reshape(-1).clone()
        if scale.numel() != self.n_mels or shift.numel() != self.n_mels:
            raise ValueError(f"output affine needs n_mels={self.n_mels} values per vector")
        self._out_affine = (scale, shift)

    def fold_batchnorm(self, bn: nn.BatchNorm2d) -> None:
        """Fold an eval-mode BatchNorm over mel bins — `Cnn6.bn1`, applied to the transposed log-mel first thing in
        the PANNs head (panns.py:146, 169-172) — into the epilogue: the head then skips its `bn1` (and, with
        `time_major=True`, its transposes).  Uses the running statistics as they are now; call again after loading
        new weights.  Training-mode BatchNorm (batch statistics) cannot be folded."""
        if bn.num_features != self.n_mels:
            raise ValueError(f"BatchNorm has {bn.num_features} features, the layer has n_mels={self.n_mels}")
        if bn.running_mean is None or bn.running_var is None:
            raise ValueError("BatchNorm without running statistics cannot be folded")
        with torch.no_grad():
            inv = torch.rsqrt(bn.running_var.to(torch.float32) + bn.eps)
            scale = inv * bn.weight.to(torch.float32) if bn.weight is not None else inv
            shift = -bn.running_mean.to(torch.float32) * scale
            if bn.bias is not None:
                shift = shift + bn.bias.to(torch.float32)
        self.set_output_affine(scale, shift)

    def forward(self, x):
        # CUDA graphs: lambd goes to the kernels by value (its magnitude picks n_fft and the launch plan on the host), so a
        # captured step keeps the lambd of capture time — re-capture after every optimiser step that changes it
        dev = _resolve_device(self.device, self.lambd)
        if dev.type != 'cuda':
            raise _lib.DmelError(
                f"MelSpectrogramLayer lives on {dev}: this implementation is CUDA-only (sm_100a) and has no CPU "
                "fallback; move the module with .to('cuda') (the reference's default device='cpu' is not served)")
        if x.device != dev:
            x = x.to(dev, non_blocking=True)
        (batch_size, n_points) = x.shape
        if n_points // self.hop_length + 1 != self.n_time:
            # the reference fails on the slice assignment models.py:54
            raise RuntimeError(f"input has {n_points} points -> {n_points // self.hop_length + 1} frames, "
                               f"layer was built for n_time={self.n_time}")
        y, self._last_ws = dmel(x, self.lambd, hop_length=self.hop_length, optimized=self.optimized,
                                normalize_window=self.normalize_window, n_mels=self.n_mels,
                                sample_rate=self.sample_rate, f_min=self.f_min, f_max=self.f_max,
                                log_eps=self.log_eps, lambd_value=self._mirror.get(self.lambd), return_workspace=True,
                                time_major=self.time_major, out_affine=self._affine_on(dev))
        return y

    def forward_grid(self, x, lambdas):
        """This front-end at several FIXED window widths over one batch (the reference's fixed-lambda grid,
        search_spaces.py:22,29: one Ray trial per lambda, each re-reading the same clips): a list of outputs, entry k
        bit-identical to what this layer returns with `lambd = lambdas[k]`.  No gradients (the grid is the
        `trainable=False` case); the output affine map is not applied."""
        dev = _resolve_device(self.device, self.lambd)
        if dev.type != 'cuda':
            raise _lib.DmelError(f"MelSpectrogramLayer lives on {dev}: this implementation is CUDA-only (sm_100a)")
        if x.device != dev:
            x = x.to(dev, non_blocking=True)
        if x.shape[1] // self.hop_length + 1 != self.n_time:
            raise RuntimeError(f"input has {x.shape[1]} points, layer was built for n_time={self.n_time}")
        return dmel_multi(x, lambdas, hop_length=self.hop_length, optimized=self.optimized,
                          normalize_window=self.normalize_window, n_mels=self.n_mels, sample_rate=self.sample_rate,
                          f_min=self.f_min, f_max=self.f_max, log_eps=self.log_eps, time_major=self.time_major)

    @property
    def last_energy(self):
        """0-d fp64 device tensor: sum of the last forward's output, produced inside the kernel — what train.py:58
        copies the whole spectrogram batch to the host for.  `None` before the first forward."""
        return None if self._last_ws is None else workspace_energy(self._last_ws)

    def _affine_on(self, dev: torch.device):
        aff = self._out_affine
        if aff is not None and aff[0].device != dev:
            aff = self._out_affine = (aff[0].to(dev), aff[1].to(dev))
        return aff

    def invalidate_lambda_cache(self) -> None:
        """Force the next forward to re-read `lambd` from the device.  Only needed after writes that bypass
        autograd's version counter (`layer.lambd.data.add_(...)`-style code); optimizers, `fill_`, `copy_` and
        `load_state_dict` are picked up automatically."""
        self._mirror.invalidate()

    def extra_repr(self):
        return (f"n_mels={self.n_mels}, sample_rate={self.sample_rate}, hop_length={self.hop_length}, "
                f"optimized={self.optimized}, normalize_window={self.normalize_window}, log_eps={self.log_eps}")


class SpectrogramLayer(nn.Module):
    def __init__(self, init_lambd, device='cpu', optimized=False, size=(512, 1024), hop_length=1,
                 normalize_window=False, *, log_eps: Optional[float] = None):
        super().__init__()
        self.hop_length = hop_length
        self.lambd = nn.Parameter(init_lambd)
        self.device = device
        self.size = size
        self.optimized = optimized
        self.normalize_window = normalize_window
        self.log_eps = log_eps
        self._mirror = _LambdaMirror()
        self._last_ws = None

    @property
    def last_energy(self):
        """See MelSpectrogramLayer.last_energy."""
        return None if self._last_ws is None else workspace_energy(self._last_ws)

    def forward(self, x):
        dev = _resolve_device(self.device, self.lambd)
        if dev.type != 'cuda':
            raise _lib.DmelError(
                f"SpectrogramLayer lives on {dev}: this implementation is CUDA-only (sm_100a) and has no CPU "
                "fallback; move the module with .to('cuda')")
        if x.device != dev:
            x = x.to(dev, non_blocking=True)
        s, self._last_ws = dmel(x, self.lambd, hop_length=self.hop_length, optimized=self.optimized,
                                normalize_window=self.normalize_window, n_mels=0, log_eps=self.log_eps,
                                lambd_value=self._mirror.get(self.lambd), return_workspace=True)
        if self.optimized and tuple(s.shape[2:]) != tuple(self.size):
            # models.py:186,198: the reference allocates `size` and assigns the (F, T) spectrogram into it
            raise RuntimeError(f"spectrogram of shape {tuple(s.shape[2:])} does not fit size={tuple(self.size)}")
        return s

    def invalidate_lambda_cache(self) -> None:
        """See MelSpectrogramLayer.invalidate_lambda_cache."""
        self._mirror.invalidate()

    def extra_repr(self):
        return (f"hop_length={self.hop_length}, optimized={self.optimized}, size={self.size}, "
                f"normalize_window={self.normalize_window}, log_eps={self.log_eps}")
