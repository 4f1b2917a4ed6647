"""The DMEL custom op: host logic + torch.autograd.Function over the C ABI (include/dmel.h).

Host logic mirrored from the reference:
  * window length  — time_frequency.py:34-43, 60-65: next_power_of_2(int(6*|lambd|)) when
    `optimized`, else len(x); n_fft = win_length (optimized) or 2*win_length
    (time_frequency.py:48 / :51).  The reference does this with one device->host sync per
    sample; here it is one `.item()` per forward for the whole batch.
  * output geometry — models.py:30 n_time = n_points // hop + 1, rows = n_mels or n_fft/2+1.
PyTorch only provides device memory and the current stream; all arithmetic on the path is
in libdmel_b200.so.  Non-CUDA tensors are rejected (no CPU fallback).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch
from torch.autograd.function import once_differentiable

from . import _lib
from .melbank import DeviceBank, device_bank


def shift_bit_length(x) -> int:
    """time_frequency.py:60-62: smallest power of two >= int(x) (0 -> 2, 1 -> 1)."""
    xi = int(x)
    return 1 << (xi - 1).bit_length()


def next_power_of_2(x) -> int:
    """time_frequency.py:64-65."""
    return shift_bit_length(x)


def window_length_from_lambda(lambd_value: float, n_stds: int = 6) -> int:
    """time_frequency.py:39 evaluated like the reference: the product is an fp32 tensor op
    (|lambd| is fp32, n_stds a Python int), `int()` truncates."""
    prod = np.float32(abs(np.float32(lambd_value))) * np.float32(n_stds)
    return next_power_of_2(prod)


@dataclass(frozen=True)
class Geometry:
    batch: int
    n_points: int
    hop: int
    win_length: int
    n_fft: int
    n_frames: int
    n_freqs: int
    rows: int          # n_mels, or n_freqs in spectrogram mode
    n_mels: int        # 0 in spectrogram mode


def resolve_geometry(batch: int, n_points: int, hop: int, lambd_value: float, optimized: bool,
                     n_mels: int = 0, n_stds: int = 6) -> Geometry:
    if optimized:
        win = window_length_from_lambda(lambd_value, n_stds)
        n_fft = win
    else:
        win = int(n_points)
        n_fft = 2 * win
    n_freqs = n_fft // 2 + 1
    return Geometry(int(batch), int(n_points), int(hop), win, n_fft, n_points // hop + 1, n_freqs,
                    n_mels if n_mels > 0 else n_freqs, n_mels)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.DmelError(
            f"{name} is on {t.device}: the B200 DMEL op has no CPU path (device='cpu' layers cannot run; "
            "use the reference itself for CPU results)")


def _stream_ptr(device: torch.device) -> int:
    # raw cudaStream_t of torch's current stream on that device (no Stream object construction)
    return int(torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device()))


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (saves two driver calls per launch)."""

    def __init__(self, dev: torch.device):
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        self._ctx = None if idx == torch.cuda.current_device() else torch.cuda.device(idx)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()

    def __exit__(self, *exc):
        if self._ctx is not None:
            self._ctx.__exit__(*exc)
        return False


# lambda-only training steps use the forward-mode (JVP) kernel; set False (or DMEL_NO_JVP=1) to always recompute in
# the backward instead (cheaper forwards when most of them are never back-propagated, e.g. long validation loops).
USE_JVP = os.environ.get("DMEL_NO_JVP", "0") != "1"

_Y_SUM_OFFSET = [-1]


_jvp_ok: dict = {}


def _jvp_supported(lib, n_fft: int) -> bool:
    ok = _jvp_ok.get(n_fft)
    if ok is None:
        ok = _jvp_ok[n_fft] = bool(lib.dmel_jvp_supported(n_fft))
    return ok


class DmelFunction(torch.autograd.Function):
    """y = DMEL(x, lambd).

    Two regimes, chosen from `needs_input_grad` at forward time:
      * x needs a gradient  -> forward saves only x (and y when the log is fused) plus the small workspace (per-clip
        means, window, twiddles); the backward kernel recomputes the spectra and produces dL/dx and dL/dlambd;
      * only lambd does (the reference's training case) -> the forward-mode kernel also writes dy/dlambd, which is
        saved instead of x; the backward is a dot product with dL/dy.
    """

    @staticmethod
    def forward(ctx, x: torch.Tensor, lambd: torch.Tensor, ws: torch.Tensor, geo: Geometry, bank: Optional[DeviceBank],
                log_eps: float, flags: int, lam_value: float, jvp: bool):
        lib = _lib.load()
        dev = x.device
        shape = (geo.batch, 1, geo.n_frames, geo.rows) if (flags & _lib.FLAG_TIME_MAJOR) else (geo.batch, 1, geo.rows, geo.n_frames)
        y = torch.empty(shape, dtype=torch.float32, device=dev)
        # The reference's training case: x is data, only lambd trains (models.py:19, main.py:27).  Then dy/dlambda is
        # produced by the forward itself (derivative-window STFT) and the backward is a dot product.
        # (decided by dmel(): grad mode is always off in here and needs_input_grad mirrors requires_grad even under
        # torch.no_grad(), so validation / inference passes would otherwise pay for the forward-mode kernel)
        dy = torch.empty_like(y) if jvp else None
        bank_ref = bank.ref() if bank is not None else None
        with _on_device(dev):
            if jvp:
                rc = lib.dmel_forward_jvp_f32(
                    x.data_ptr(), geo.batch, geo.n_points, x.stride(0), lam_value, geo.win_length, geo.n_fft, geo.hop,
                    geo.n_mels, bank_ref, log_eps, flags, y.data_ptr(), dy.data_ptr(),
                    ws.data_ptr(), ws.numel(), _stream_ptr(dev))
            else:
                rc = lib.dmel_forward_f32(
                    x.data_ptr(), geo.batch, geo.n_points, x.stride(0), lam_value, geo.win_length, geo.n_fft, geo.hop,
                    geo.n_mels, bank_ref, log_eps, flags, y.data_ptr(),
                    ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        if rc:
            _lib.check(rc, "dmel_forward_jvp_f32" if jvp else "dmel_forward_f32")
        ctx.geo, ctx.bank, ctx.log_eps, ctx.flags, ctx.lam_value = geo, bank, log_eps, flags, lam_value
        ctx.ws = ws
        ctx.jvp = jvp
        ctx.lambd_shape, ctx.lambd_dtype = lambd.shape, lambd.dtype
        if jvp:
            ctx.save_for_backward(dy)
        elif log_eps >= 0:
            ctx.save_for_backward(x, y)
        else:
            ctx.save_for_backward(x)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_y: torch.Tensor):
        lib = _lib.load()
        geo = ctx.geo
        saved = ctx.saved_tensors
        gy = grad_y
        if gy.dtype != torch.float32 or not gy.is_contiguous():
            gy = gy.to(torch.float32).contiguous()
        ws = ctx.ws
        dev = gy.device
        grad_lambda = torch.empty((), dtype=torch.float32, device=dev)
        grad_x = None
        if ctx.jvp:
            dy = saved[0]
            with _on_device(dev):
                rc = lib.dmel_lambda_grad_f32(gy.data_ptr(), dy.data_ptr(), gy.numel(), ctx.lam_value, grad_lambda.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _stream_ptr(dev))
            if rc:
                _lib.check(rc, "dmel_lambda_grad_f32")
        else:
            x = saved[0]
            y_ptr = saved[1].data_ptr() if ctx.log_eps >= 0 else None
            need_dx = ctx.needs_input_grad[0]
            if need_dx:
                grad_x = torch.empty((geo.batch, geo.n_points), dtype=torch.float32, device=dev)
            with _on_device(dev):
                rc = lib.dmel_backward_f32(
                    x.data_ptr(), geo.batch, geo.n_points, x.stride(0),
                    ctx.lam_value, geo.win_length, geo.n_fft, geo.hop,
                    geo.n_mels, ctx.bank.ref() if ctx.bank is not None else None, ctx.log_eps,
                    ctx.flags, y_ptr, gy.data_ptr(),
                    grad_x.data_ptr() if need_dx else None, grad_lambda.data_ptr(),
                    ws.data_ptr(), ws.numel(), 1, _stream_ptr(dev))
            if rc:
                _lib.check(rc, "dmel_backward_f32")
        g_lambd = None
        if ctx.needs_input_grad[1]:
            g_lambd = grad_lambda
            if ctx.lambd_shape != g_lambd.shape:
                g_lambd = g_lambd.reshape(ctx.lambd_shape)
            if ctx.lambd_dtype != torch.float32:
                g_lambd = g_lambd.to(ctx.lambd_dtype)
        return grad_x, g_lambd, None, None, None, None, None, None, None


def workspace_energy(ws: torch.Tensor) -> torch.Tensor:
    """0-d fp64 view of the fused by-product the forward leaves in its workspace header: the sum of all outputs
    (what train.py:58 copies the whole spectrogram batch to the host for).  No launch, no copy."""
    off = _Y_SUM_OFFSET[0]
    if off < 0:
        off = _Y_SUM_OFFSET[0] = _lib.load().dmel_workspace_y_sum_offset()
    return ws[off:off + 8].view(torch.float64)[0]


def dmel(x: torch.Tensor, lambd: torch.Tensor, *, hop_length: int, optimized: bool,
         normalize_window: bool = False, n_mels: int = 0, sample_rate: int = 0, f_min: float = 0.0,
         f_max: Optional[float] = None, log_eps: Optional[float] = None, remove_mean: bool = True,
         lambd_value: Optional[float] = None, return_energy: bool = False, time_major: bool = False,
         out_affine: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, return_workspace: bool = False,
         n_stds: int = 6):
    """Functional form of both layers.  `n_stds` is time_frequency.py:32's window support in standard deviations
    (optimized mode: window_length = next_power_of_2(int(n_stds * |lambd|)); the layers use the reference's 6).

    x: (B, N) CUDA tensor (fp32; fp64/fp16 inputs are cast to fp32 — the reference's fp64
    synthetic dataset ends in an fp32 output buffer anyway, models.py:189).  Returns
    (B, 1, rows, N//hop + 1) fp32, contiguous: the linear (mel) power spectrogram, or its
    log(. + log_eps) when log_eps is given (the classifiers' energy_normalize, models.py:72).
    n_mels == 0 selects spectrogram mode (models.py:171-200).  `lambd_value`, if given, must equal
    lambd's current value (it only saves the device->host read that picks n_fft).  With `return_energy`
    the result is `(y, energy)`, energy a 0-d fp64 device tensor = y.sum() computed inside the forward kernel
    (what train.py:58 copies the whole spectrogram batch to the host for).  `time_major=True` returns
    (B, 1, N//hop + 1, rows) instead — the layout MelPANNsNet transposes to before Cnn6 (models.py:162).
    `out_affine=(scale, shift)` (two vectors of n_mels values, mel mode only) folds y <- scale[m]*y + shift[m] into the
    kernel epilogue, after the log: the eval-mode BatchNorm2d over mel bins that Cnn6 applies first (panns.py:169-172).
    The pair is a constant of the call (no gradient flows to it); gradients to x and lambd pass through it.
    `return_workspace=True` returns `(y, ws)` instead, ws the call's workspace tensor (`workspace_energy(ws)` is the
    energy; the layers keep it and build that view only when `last_energy` is read).
    """
    if x.dim() != 2:
        raise ValueError(f"expected x of shape (batch, n_points), got {tuple(x.shape)}")
    _require_cuda(x, "x")
    if lambd.numel() != 1:
        raise ValueError("lambd must hold exactly one element")
    # 16-bit samples (int16 PCM — the datasets' native format, datasets.py:169,208,255 — or fp16) are widened inside
    # the kernels while the signal is staged (no intermediate fp32 tensor) when the tiled FFT path serves this n_fft and
    # no gradient with respect to x is wanted; every other dtype is cast to fp32 first, as before
    x16_flag = 0
    if x.dtype in (torch.int16, torch.float16) and not x.requires_grad:
        x16_flag = _lib.FLAG_X_I16 if x.dtype == torch.int16 else _lib.FLAG_X_F16
    elif x.dtype != torch.float32:
        x = x.to(torch.float32)
    if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    if lambd.device != x.device:
        raise _lib.DmelError(f"lambd on {lambd.device} but x on {x.device}")
    # the one host sync of the forward (the reference syncs once per sample); callers that keep a
    # host mirror of lambd (the layers do, keyed on the parameter's version counter) pass it in
    lam_value = float(lambd.detach().to(torch.float32).item()) if lambd_value is None else float(lambd_value)
    B, N = x.shape
    if log_eps is not None and log_eps < 0:
        raise ValueError("log_eps must be >= 0")
    # everything below depends only on these scalars: resolve once, reuse while they repeat (training loops)
    key = (B, N, hop_length, lam_value, optimized, normalize_window, n_mels, sample_rate, f_min, f_max, log_eps, remove_mean,
           time_major, x.device.index, x16_flag, n_stds)
    plan = _plan_cache.get(key)
    if plan is None:
        geo = resolve_geometry(B, N, hop_length, lam_value, optimized, n_mels, n_stds)
        lib = _lib.load()
        if not lib.dmel_nfft_supported(geo.n_fft):
            raise _lib.DmelError(f"n_fft={geo.n_fft} (lambd={lam_value}) has no kernel")
        bank = None
        if n_mels > 0:
            fmax = f_max if f_max is not None else sample_rate // 2
            bank = device_bank(geo.n_freqs, f_min, fmax, n_mels, sample_rate, x.device)
        flags = ((_lib.FLAG_NORMALIZE_WINDOW if normalize_window else 0) | (0 if remove_mean else _lib.FLAG_KEEP_MEAN)
                 | (_lib.FLAG_TIME_MAJOR if time_major else 0))
        if x16_flag and lib.dmel_x16_supported(geo.n_fft):
            flags |= x16_flag
        if time_major and not lib.dmel_time_major_supported(geo.n_fft, n_mels):
            raise _lib.DmelError(f"time_major output is implemented for n_fft 128..4096 (and 8192 / 16384 with a mel bank), "
                                 f"got n_fft={geo.n_fft}, n_mels={n_mels}")
        eps = -1.0 if log_eps is None else float(log_eps)
        plan = (geo, bank, eps, flags, int(lib.dmel_workspace_bytes(geo.batch, geo.n_fft)))
        if len(_plan_cache) > 64:
            _plan_cache.clear()
        _plan_cache[key] = plan
    geo, bank, eps, flags, ws_bytes = plan
    if x16_flag and not (flags & x16_flag):
        x = x.to(torch.float32)        # this n_fft is served by a kernel family without 16-bit staging
    if out_affine is not None:
        if bank is None:
            raise ValueError("out_affine needs n_mels > 0 (it is a per-mel map)")
        # the struct only holds the two pointers: reuse it while the same vectors come back (the layers keep theirs)
        direct = all(v.dtype == torch.float32 and v.device == x.device and v.is_contiguous() for v in out_affine)
        akey = (id(bank), out_affine[0].data_ptr(), out_affine[1].data_ptr())
        abank = _affine_cache.get(akey) if direct else None
        if abank is None:
            abank = bank.with_affine(*out_affine)     # converts / copies whatever is not fp32 on x's device
            if direct:
                if len(_affine_cache) > 16:
                    _affine_cache.clear()
                _affine_cache[akey] = abank
        bank = abank
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    # forward-mode kernel only when a backward can follow: grad mode on, lambd trains, x does not
    jvp = (USE_JVP and torch.is_grad_enabled() and lambd.requires_grad and not x.requires_grad
           and _jvp_supported(_lib.load(), geo.n_fft))
    y = DmelFunction.apply(x, lambd, ws, geo, bank, eps, flags, lam_value, jvp)
    if return_workspace:
        return y, ws
    return (y, workspace_energy(ws)) if return_energy else y


_plan_cache: dict = {}
_affine_cache: dict = {}


def dmel_multi(x: torch.Tensor, lambdas, *, hop_length: int, optimized: bool, normalize_window: bool = False,
               n_mels: int = 0, sample_rate: int = 0, f_min: float = 0.0, f_max: Optional[float] = None,
               log_eps: Optional[float] = None, remove_mean: bool = True, time_major: bool = False):
    """One batch through several FIXED window widths in one call: the fixed-lambda grid of the reference's search
    spaces (search_spaces.py:22,29 — `trainable=False` with lambda in {10, 35, 300 ms} * sample_rate / 6), where each
    grid point is the same front-end applied to the same clips.  Returns a list with one (B, 1, rows, T) tensor per
    lambda, each bit-identical to `dmel(x, lambda_k, ...)`; the per-clip means of models.py:38 are summed once and
    shared, and x is read (and, for 16-bit input, widened) from the one device tensor.  No autograd: the grid is the
    non-trainable case (for a trainable lambd use `dmel` / the layers)."""
    import ctypes
    if x.dim() != 2:
        raise ValueError(f"expected x of shape (batch, n_points), got {tuple(x.shape)}")
    _require_cuda(x, "x")
    lams = [float(v) for v in (lambdas.detach().flatten().tolist() if isinstance(lambdas, torch.Tensor) else lambdas)]
    if not lams:
        raise ValueError("dmel_multi needs at least one lambda")
    lib = _lib.load()
    B, N = x.shape
    geos = [resolve_geometry(B, N, hop_length, lam, optimized, n_mels) for lam in lams]
    for geo, lam in zip(geos, lams):
        if not lib.dmel_nfft_supported(geo.n_fft):
            raise _lib.DmelError(f"n_fft={geo.n_fft} (lambd={lam}) has no kernel")
        if time_major and not lib.dmel_time_major_supported(geo.n_fft, n_mels):
            raise _lib.DmelError(f"time_major output is implemented for n_fft 128..4096 (and 8192 / 16384 with a mel bank), "
                                 f"got n_fft={geo.n_fft}, n_mels={n_mels}")
    flags = ((_lib.FLAG_NORMALIZE_WINDOW if normalize_window else 0) | (0 if remove_mean else _lib.FLAG_KEEP_MEAN)
             | (_lib.FLAG_TIME_MAJOR if time_major else 0))
    # 16-bit samples are widened in the kernels only if every entry's kernel family stages them
    if x.dtype in (torch.int16, torch.float16) and all(lib.dmel_x16_supported(g.n_fft) for g in geos):
        flags |= _lib.FLAG_X_I16 if x.dtype == torch.int16 else _lib.FLAG_X_F16
    elif x.dtype != torch.float32:
        x = x.to(torch.float32)
    if x.stride(1) != 1 or x.stride(0) < x.shape[1]:
        x = x.contiguous()
    if log_eps is not None and log_eps < 0:
        raise ValueError("log_eps must be >= 0")
    eps = -1.0 if log_eps is None else float(log_eps)
    K = len(lams)
    banks = None
    bank_objs = []
    if n_mels > 0:
        fmax = f_max if f_max is not None else sample_rate // 2
        bank_objs = [device_bank(g.n_freqs, f_min, fmax, n_mels, sample_rate, x.device) for g in geos]
        banks = (ctypes.POINTER(_lib.MelBankStruct) * K)(*[ctypes.pointer(b.struct) for b in bank_objs])
    ys, wss = [], []
    for g in geos:
        shape = (B, 1, g.n_frames, g.rows) if time_major else (B, 1, g.rows, g.n_frames)
        ys.append(torch.empty(shape, dtype=torch.float32, device=x.device))
        wss.append(torch.empty(int(lib.dmel_workspace_bytes(B, g.n_fft)), dtype=torch.uint8, device=x.device))
    with _on_device(x.device):
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.dmel_forward_multi_f32(
            x.data_ptr(), B, N, x.stride(0), K, (ctypes.c_float * K)(*lams),
            (ctypes.c_int32 * K)(*[g.win_length for g in geos]), (ctypes.c_int32 * K)(*[g.n_fft for g in geos]),
            hop_length, n_mels, banks, eps, flags, (ctypes.c_void_p * K)(*[y.data_ptr() for y in ys]),
            (ctypes.c_void_p * K)(*[w.data_ptr() for w in wss]), (ctypes.c_size_t * K)(*[w.numel() for w in wss]), st),
            "dmel_forward_multi_f32")
    del bank_objs
    return ys


def query_launch(batch: int, n_points: int, n_fft: int, hop: int, rows: int, bank_nnz: int = 0,
                 which: int = 0) -> dict:
    """What the launcher would pick: frames per CTA tile, FFT groups per CTA, dynamic shared memory, threads per
    CTA and the path (0 = FFT, 1 = direct DFT).  which: 0 fwd, 1 bwd (dx + dlambda), 2 bwd (dlambda only)."""
    import ctypes
    lib = _lib.load()
    v = [ctypes.c_int32() for _ in range(5)]
    _lib.check(lib.dmel_query_launch(batch, n_points, n_fft, hop, rows, bank_nnz, which, *[ctypes.byref(a) for a in v]),
               "dmel_query_launch")
    return dict(tile_frames=v[0].value, groups=v[1].value, smem_bytes=v[2].value, threads=v[3].value, path=v[4].value)
