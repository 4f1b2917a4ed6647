"""HTK mel filterbank and its sparse device forms.

The reference rebuilds `torchaudio.functional.melscale_fbanks(n_freqs, f_min, f_max, n_mels,
sample_rate)` on the CPU for every sample of every batch (models.py:42-51) and multiplies the
dense (F, M) matrix (models.py:53).  torchaudio is a third-party dependency of the reference
(pinned torchaudio==0.13.1 in requirements.txt, not vendored), so its published algorithm is
restated here — HTK mel scale, triangular filters, norm=None — with the same fp32 operation
order, and tests/test_host_logic.py (test_melbank_bit_identical_to_torchaudio_and_pack_roundtrip) pins it bit-for-bit against
the torchaudio in this image.

The kernels never see the dense matrix: every frequency bin has at most two non-zero mel
weights, so the bank is shipped as two CSR-like forms (see include/dmel.h, dmel_melbank).
"""
from __future__ import annotations

import ctypes
import math
import threading
from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np
import torch

from ._lib import MelBankStruct


def hz_to_mel_htk(freq_hz: float) -> float:
    return 2595.0 * math.log10(1.0 + freq_hz / 700.0)


def melscale_fbanks_htk(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int) -> torch.Tensor:
    """(n_freqs, n_mels) fp32 triangular HTK-mel bank; column m peaks at mel point m+1.

    Bin centres are linspace(0, sample_rate // 2, n_freqs) (integer Nyquist, as the
    reference's dependency does); filter edges are equally spaced on the HTK mel axis
    between f_min and f_max.
    """
    bin_hz = torch.linspace(0, sample_rate // 2, n_freqs)
    mel_pts = torch.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    edge_hz = 700.0 * (10.0 ** (mel_pts / 2595.0) - 1.0)
    width = edge_hz[1:] - edge_hz[:-1]                       # (M+1,)
    dist = edge_hz.unsqueeze(0) - bin_hz.unsqueeze(1)        # (F, M+2): edge - bin
    falling = (-1.0 * dist[:, :-2]) / width[:-1]             # rises from the left edge
    rising = dist[:, 2:] / width[1:]                         # falls to the right edge
    return torch.max(torch.zeros(1), torch.min(falling, rising))


@dataclass
class PackedBank:
    """Host-side sparse forms of a dense (F, M) bank (layout of struct dmel_melbank)."""

    n_freqs: int
    n_mels: int
    mel_k0: np.ndarray   # int32 [M]
    mel_off: np.ndarray  # int32 [M+1]
    mel_w: np.ndarray    # float32 [nnz_mel]
    bin_m0: np.ndarray   # int32 [F]
    bin_off: np.ndarray  # int32 [F+1]
    bin_w: np.ndarray    # float32 [nnz_bin]


def pack_bank(fb: np.ndarray) -> PackedBank:
    """Dense (F, M) -> contiguous-run sparse forms. Works for any matrix (a dense one just
    yields full-length runs); zeros inside a run are kept so runs stay contiguous."""
    fb = np.ascontiguousarray(fb, dtype=np.float32)
    F, M = fb.shape
    nz = fb != 0.0

    def runs(mask: np.ndarray):
        # mask: (rows, cols) -> for each row first nz col, run length
        any_ = mask.any(axis=1)
        first = np.where(any_, mask.argmax(axis=1), 0)
        last = np.where(any_, mask.shape[1] - 1 - mask[:, ::-1].argmax(axis=1), -1)
        return first.astype(np.int32), (last - first + 1).clip(min=0).astype(np.int32)

    k0, kcnt = runs(nz.T)   # per mel: run of bins
    m0, mcnt = runs(nz)     # per bin: run of mels
    # the forward kernel consumes 4 taps per step: pad every mel run with zero weights to a multiple of 4
    # (bins read behind the run stay <= F+2, inside the kernel's L-entry spectrum buffer)
    kcnt_true = kcnt.copy()
    kcnt = ((kcnt + 3) // 4 * 4).astype(np.int32)
    mel_off = np.zeros(M + 1, np.int32)
    mel_off[1:] = np.cumsum(kcnt)
    bin_off = np.zeros(F + 1, np.int32)
    bin_off[1:] = np.cumsum(mcnt)
    mel_w = np.zeros(max(int(mel_off[-1]), 1), np.float32)
    for m in range(M):
        mel_w[mel_off[m]:mel_off[m] + kcnt_true[m]] = fb[k0[m]:k0[m] + kcnt_true[m], m]
    bin_w = np.zeros(max(int(bin_off[-1]), 1), np.float32)
    for k in range(F):
        bin_w[bin_off[k]:bin_off[k + 1]] = fb[k, m0[k]:m0[k] + mcnt[k]]
    return PackedBank(F, M, k0, mel_off, mel_w, m0, bin_off, bin_w)


def unpack_bank(p: PackedBank) -> Tuple[np.ndarray, np.ndarray]:
    """Inverse of pack_bank from either form (used by tests)."""
    a = np.zeros((p.n_freqs, p.n_mels), np.float32)
    b = np.zeros((p.n_freqs, p.n_mels), np.float32)
    for m in range(p.n_mels):
        c = min(p.mel_off[m + 1] - p.mel_off[m], p.n_freqs - p.mel_k0[m])   # runs are zero-padded to 4 taps
        a[p.mel_k0[m]:p.mel_k0[m] + c, m] = p.mel_w[p.mel_off[m]:p.mel_off[m] + c]
    for k in range(p.n_freqs):
        c = p.bin_off[k + 1] - p.bin_off[k]
        b[k, p.bin_m0[k]:p.bin_m0[k] + c] = p.bin_w[p.bin_off[k]:p.bin_off[k + 1]]
    return a, b


class DeviceBank:
    """A PackedBank resident on one CUDA device + the ctypes struct pointing into it."""

    def __init__(self, packed: PackedBank, device: torch.device):
        self.packed = packed
        self.device = device
        # one blob, every sub-array on a 16-byte boundary and padded to a multiple of 4 elements
        # (the kernels stage them with 16-byte cp.async copies)
        parts = [packed.mel_k0.astype(np.int32), packed.mel_off.astype(np.int32), packed.bin_m0.astype(np.int32),
                 packed.bin_off.astype(np.int32), packed.mel_w.astype(np.float32).view(np.int32),
                 packed.bin_w.astype(np.float32).view(np.int32)]
        offs, words = [], 0
        for a in parts:
            offs.append(words)
            words += (a.size + 3) // 4 * 4 + 4
        blob = np.zeros(words, np.int32)
        for o, a in zip(offs, parts):
            blob[o:o + a.size] = a
        self._blob = torch.from_numpy(blob).to(device)
        base = self._blob.data_ptr()
        assert base % 16 == 0
        s = MelBankStruct()
        s.n_freqs, s.n_mels = packed.n_freqs, packed.n_mels
        s.mel_nnz, s.bin_nnz = int(packed.mel_off[-1]), int(packed.bin_off[-1])
        s.mel_k0, s.mel_off, s.bin_m0, s.bin_off, s.mel_w, s.bin_w = (base + 4 * o for o in offs)
        self.struct = s
        # segment tables of the tiled forward (dmel.h: dmel_build_segment_tables): built once per bank by the library
        from . import _lib
        lib = _lib.load()
        n_fft = 2 * (packed.n_freqs - 1)
        nbytes = int(lib.dmel_segment_table_bytes(n_fft))
        self._seg = None
        if nbytes:
            self._seg = torch.zeros(nbytes, dtype=torch.uint8, device=device)
            idx = device.index if device.index is not None else torch.cuda.current_device()
            with torch.cuda.device(idx):
                stream = int(torch._C._cuda_getCurrentRawStream(idx))
                usable = ctypes.c_int32(0)
                _lib.check(lib.dmel_build_segment_tables(ctypes.byref(s), n_fft, self._seg.data_ptr(), nbytes, stream,
                                                         ctypes.byref(usable)), "dmel_build_segment_tables")
            if usable.value:
                s.seg_tables = self._seg.data_ptr()
            else:
                self._seg = None

    def ref(self):
        return ctypes.byref(self.struct)

    def with_affine(self, scale: torch.Tensor, shift: torch.Tensor) -> "AffineBank":
        return AffineBank(self, scale, shift)


class AffineBank:
    """A DeviceBank plus a per-mel output map y <- scale[m]*y + shift[m] (dmel.h: out_scale / out_shift): the
    eval-mode BatchNorm2d over mel bins that panns.py:169-172 applies first.  Holds the two device vectors alive."""

    def __init__(self, bank: DeviceBank, scale: torch.Tensor, shift: torch.Tensor):
        m = bank.packed.n_mels
        vec = []
        for name, v in (("scale", scale), ("shift", shift)):
            v = v.detach().to(device=bank.device, dtype=torch.float32).reshape(-1).contiguous()
            if v.numel() != m:
                raise ValueError(f"out_affine {name} must hold n_mels={m} values, got {v.numel()}")
            vec.append(v)
        self.bank, self.packed, self.device = bank, bank.packed, bank.device
        self.scale, self.shift = vec
        s = MelBankStruct()
        ctypes.memmove(ctypes.byref(s), ctypes.byref(bank.struct), ctypes.sizeof(MelBankStruct))
        s.out_scale, s.out_shift = self.scale.data_ptr(), self.shift.data_ptr()
        self.struct = s

    def ref(self):
        return ctypes.byref(self.struct)


_cache: Dict[tuple, DeviceBank] = {}
_cache_lock = threading.Lock()


def device_bank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int,
                device: torch.device) -> DeviceBank:
    """Cached per (F, f_min, f_max, M, sr, device): built once instead of once per sample."""
    device = torch.device(device)
    key = (int(n_freqs), float(f_min), float(f_max), int(n_mels), int(sample_rate), device.type, device.index)
    bank = _cache.get(key)
    if bank is None:
        with _cache_lock:
            bank = _cache.get(key)
            if bank is None:
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    fb = melscale_fbanks_htk(n_freqs, f_min, f_max, n_mels, sample_rate).numpy()
                bank = DeviceBank(pack_bank(fb), device)
                _cache[key] = bank
    return bank
