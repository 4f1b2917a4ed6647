"""ctypes binding of libdmel_b200.so — the C ABI declared in include/dmel.h.

There is no CPU fallback: if the shared library is missing or does not export a symbol
this module raises, and every product entry point goes through it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdmel_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")
ABI_VERSION = 4
FLAG_NORMALIZE_WINDOW = 1
FLAG_KEEP_MEAN = 2
FLAG_TIME_MAJOR = 4
FLAG_X_I16 = 8
FLAG_X_F16 = 16

# every symbol include/dmel.h declares
EXPORTED_SYMBOLS = (
    "dmel_abi_version",
    "dmel_status_string",
    "dmel_nfft_supported",
    "dmel_window_length_from_lambda",
    "dmel_workspace_bytes",
    "dmel_workspace_y_sum_offset",
    "dmel_forward_f32",
    "dmel_forward_jvp_f32",
    "dmel_forward_multi_f32",
    "dmel_jvp_supported",
    "dmel_time_major_supported",
    "dmel_x16_supported",
    "dmel_segment_table_bytes",
    "dmel_build_segment_tables",
    "dmel_lambda_grad_f32",
    "dmel_backward_f32",
    "dmel_backward_allreduce_f32",
    "dmel_query_launch",
    "dmel_peer_buffer_bytes",
    "dmel_peer_allreduce_f32",
    "dmel_peer_alloc",
    "dmel_peer_open",
    "dmel_peer_close",
    "dmel_peer_free",
    "dmel_peer_timed_out",
    "dmel_launch_count",
    "dmel_profile_enable",
    "dmel_profile_read",
)


class MelBankStruct(ctypes.Structure):
    """struct dmel_melbank (include/dmel.h)."""

    _fields_ = [
        ("n_freqs", ctypes.c_int32),
        ("n_mels", ctypes.c_int32),
        ("mel_nnz", ctypes.c_int32),
        ("bin_nnz", ctypes.c_int32),
        ("mel_k0", ctypes.c_void_p),
        ("mel_off", ctypes.c_void_p),
        ("mel_w", ctypes.c_void_p),
        ("bin_m0", ctypes.c_void_p),
        ("bin_off", ctypes.c_void_p),
        ("bin_w", ctypes.c_void_p),
        ("out_scale", ctypes.c_void_p),
        ("out_shift", ctypes.c_void_p),
        ("seg_tables", ctypes.c_void_p),
    ]


class DmelError(RuntimeError):
    pass


_lock = threading.Lock()
_lib = None


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC_DIR, f) for f in os.listdir(CSRC_DIR) if f.endswith((".cu", ".cuh")) or f == "Makefile")
    files.append(os.path.join(_HERE, "..", "include", "dmel.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a into libdmel_b200.so (in-tree) with nvcc.  Up-to-date-ness is decided by a hash of
    the sources stored next to the library (file times do not survive the copy to the GPU box)."""
    stamp = LIB_PATH + ".srchash"
    want = _source_hash()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == want:
                return LIB_PATH
    # the tiled-kernel instantiations are separate objects: build them in parallel
    proc = subprocess.run(["make", "-B", f"-j{max(1, os.cpu_count() or 1)}", "-C", CSRC_DIR], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
        print(proc.stderr)
    if proc.returncode != 0:
        raise DmelError(f"building libdmel_b200.so failed (exit {proc.returncode})")
    with open(stamp, "w") as fh:
        fh.write(want + "\n")
    return LIB_PATH


def _declare(lib: ctypes.CDLL) -> None:
    c = ctypes
    lib.dmel_abi_version.restype = c.c_int
    lib.dmel_abi_version.argtypes = []
    lib.dmel_status_string.restype = c.c_char_p
    lib.dmel_status_string.argtypes = [c.c_int]
    lib.dmel_nfft_supported.restype = c.c_int
    lib.dmel_nfft_supported.argtypes = [c.c_int32]
    lib.dmel_window_length_from_lambda.restype = c.c_int32
    lib.dmel_window_length_from_lambda.argtypes = [c.c_float]
    lib.dmel_workspace_bytes.restype = c.c_size_t
    lib.dmel_workspace_bytes.argtypes = [c.c_int64, c.c_int32]
    lib.dmel_workspace_y_sum_offset.restype = c.c_size_t
    lib.dmel_workspace_y_sum_offset.argtypes = []
    lib.dmel_forward_f32.restype = c.c_int
    lib.dmel_forward_f32.argtypes = [
        c.c_void_p, c.c_int64, c.c_int64, c.c_int64,          # x, batch, n_points, stride
        c.c_float, c.c_int32, c.c_int32, c.c_int32,           # lambda, win_length, n_fft, hop
        c.c_int32, c.POINTER(MelBankStruct), c.c_float,       # n_mels, bank, log_eps
        c.c_int32, c.c_void_p,                                # flags, y
        c.c_void_p, c.c_size_t, c.c_void_p,                   # workspace, bytes, stream
    ]
    lib.dmel_forward_jvp_f32.restype = c.c_int
    lib.dmel_forward_jvp_f32.argtypes = lib.dmel_forward_f32.argtypes[:13] + [c.c_void_p] + lib.dmel_forward_f32.argtypes[13:]
    lib.dmel_forward_multi_f32.restype = c.c_int
    lib.dmel_forward_multi_f32.argtypes = [
        c.c_void_p, c.c_int64, c.c_int64, c.c_int64,          # x, batch, n_points, stride
        c.c_int32, c.POINTER(c.c_float), c.POINTER(c.c_int32), c.POINTER(c.c_int32),   # n_lambda, lambdas, wins, n_ffts
        c.c_int32, c.c_int32, c.POINTER(c.POINTER(MelBankStruct)), c.c_float,          # hop, n_mels, banks, log_eps
        c.c_int32, c.POINTER(c.c_void_p),                     # flags, ys
        c.POINTER(c.c_void_p), c.POINTER(c.c_size_t), c.c_void_p,                      # workspaces, bytes, stream
    ]
    lib.dmel_jvp_supported.restype = c.c_int
    lib.dmel_jvp_supported.argtypes = [c.c_int32]
    lib.dmel_time_major_supported.restype = c.c_int
    lib.dmel_time_major_supported.argtypes = [c.c_int32, c.c_int32]
    lib.dmel_x16_supported.restype = c.c_int
    lib.dmel_x16_supported.argtypes = [c.c_int32]
    lib.dmel_segment_table_bytes.restype = c.c_size_t
    lib.dmel_segment_table_bytes.argtypes = [c.c_int32]
    lib.dmel_build_segment_tables.restype = c.c_int
    lib.dmel_build_segment_tables.argtypes = [c.POINTER(MelBankStruct), c.c_int32, c.c_void_p, c.c_size_t, c.c_void_p,
                                              c.POINTER(c.c_int32)]
    lib.dmel_lambda_grad_f32.restype = c.c_int
    lib.dmel_lambda_grad_f32.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_float, c.c_void_p, c.c_void_p, c.c_size_t,
                                         c.c_void_p]
    lib.dmel_peer_buffer_bytes.restype = c.c_size_t
    lib.dmel_peer_buffer_bytes.argtypes = [c.c_int32]
    lib.dmel_peer_alloc.restype = c.c_int
    lib.dmel_peer_alloc.argtypes = [c.c_int32, c.POINTER(c.c_void_p), c.c_char_p]
    lib.dmel_peer_open.restype = c.c_int
    lib.dmel_peer_open.argtypes = [c.c_char_p, c.POINTER(c.c_void_p)]
    lib.dmel_peer_close.restype = c.c_int
    lib.dmel_peer_close.argtypes = [c.c_void_p]
    lib.dmel_peer_free.restype = c.c_int
    lib.dmel_peer_free.argtypes = [c.c_void_p]
    lib.dmel_peer_timed_out.restype = c.c_int
    lib.dmel_peer_timed_out.argtypes = [c.c_void_p, c.c_int32, c.POINTER(c.c_int32)]
    lib.dmel_peer_allreduce_f32.restype = c.c_int
    lib.dmel_peer_allreduce_f32.argtypes = [c.c_void_p, c.c_void_p, c.c_int32, c.c_int32, c.c_void_p]
    lib.dmel_backward_f32.restype = c.c_int
    lib.dmel_backward_f32.argtypes = [
        c.c_void_p, c.c_int64, c.c_int64, c.c_int64,
        c.c_float, c.c_int32, c.c_int32, c.c_int32,
        c.c_int32, c.POINTER(MelBankStruct), c.c_float,
        c.c_int32, c.c_void_p, c.c_void_p,                    # flags, y, grad_y
        c.c_void_p, c.c_void_p,                               # grad_x, grad_lambda
        c.c_void_p, c.c_size_t, c.c_int32, c.c_void_p,        # workspace, bytes, prepared, stream
    ]
    lib.dmel_backward_allreduce_f32.restype = c.c_int
    lib.dmel_backward_allreduce_f32.argtypes = lib.dmel_backward_f32.argtypes + [c.c_void_p, c.c_int32, c.c_int32]
    lib.dmel_query_launch.restype = c.c_int
    lib.dmel_query_launch.argtypes = [
        c.c_int64, c.c_int64, c.c_int32, c.c_int32, c.c_int32, c.c_int32, c.c_int32,
        c.POINTER(c.c_int32), c.POINTER(c.c_int32), c.POINTER(c.c_int32), c.POINTER(c.c_int32), c.POINTER(c.c_int32),
    ]


def _declare_instrumentation(lib: ctypes.CDLL) -> None:
    c = ctypes
    lib.dmel_launch_count.restype = c.c_longlong
    lib.dmel_launch_count.argtypes = [c.c_int32]
    lib.dmel_profile_enable.restype = None
    lib.dmel_profile_enable.argtypes = [c.c_int32]
    lib.dmel_profile_read.restype = c.c_int
    lib.dmel_profile_read.argtypes = [c.c_int32, c.POINTER(c.c_double), c.POINTER(c.c_longlong)]


def load() -> ctypes.CDLL:
    """Load (once) and return the library; raises DmelError if it cannot be used."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DmelError(
                f"{LIB_PATH} not found — run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C differentiable-mel-spectrogram_b200/csrc`). There is no CPU fallback."
            )
        # DMEL_LIB_PATH: load another build of the same ABI (kernel experiments)
        lib = ctypes.CDLL(os.environ.get("DMEL_LIB_PATH") or LIB_PATH)
        missing = [s for s in EXPORTED_SYMBOLS if not hasattr(lib, s)]
        if missing:
            raise DmelError(f"{LIB_PATH} does not export {missing}")
        _declare(lib)
        _declare_instrumentation(lib)
        got = lib.dmel_abi_version()
        if got != ABI_VERSION:
            raise DmelError(f"libdmel_b200.so ABI version {got}, binding expects {ABI_VERSION}")
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().dmel_status_string(rc).decode()
        raise DmelError(f"{what}: {msg} (status {rc})")
