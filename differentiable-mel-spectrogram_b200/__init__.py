"""B200-native (sm_100a) implementation of the DMEL layer's hot path.

Importable as `dmel_b200` (alias module at the repo root; this directory's name is not a Python
identifier).  Public surface = the reference's own: `MelSpectrogramLayer`, `SpectrogramLayer`
(models.py), `differentiable_spectrogram`, `differentiable_gaussian_window`, `next_power_of_2`
(time_frequency.py), plus the functional `dmel` op and the data-parallel helpers.
"""
from . import _lib
from ._lib import DmelError, build, load
from .dist import PeerRing, allreduce_lambda_grad, broadcast_lambda, shard_bounds
from .layers import MelSpectrogramLayer, SpectrogramLayer
from .melbank import melscale_fbanks_htk, pack_bank
from .op import (dmel, dmel_multi, next_power_of_2, query_launch, resolve_geometry, shift_bit_length, window_length_from_lambda,
                 workspace_energy)
from .time_frequency import differentiable_gaussian_window, differentiable_spectrogram

__all__ = [
    "MelSpectrogramLayer", "SpectrogramLayer", "dmel", "dmel_multi", "differentiable_spectrogram",
    "differentiable_gaussian_window", "next_power_of_2", "shift_bit_length", "window_length_from_lambda",
    "resolve_geometry", "query_launch", "melscale_fbanks_htk", "pack_bank", "allreduce_lambda_grad",
    "broadcast_lambda", "shard_bounds", "PeerRing", "DmelError", "build", "load", "workspace_energy",
]
