"""Data-parallel use of the layer: clips shard over ranks, lambda's gradient is the only exchange.

Frames never cross clips, so `x`, `y`, `dL/dy` and `dL/dx` stay rank-local; `lambd` is
replicated and its scalar gradient is summed with ONE all-reduce (NCCL over NVLink on GPUs,
gloo in the CPU tests).  Under DistributedDataParallel nothing here is needed: `lambd` is an
ordinary parameter and DDP's bucket all-reduce covers it.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n_clips: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of the batch: the first n_clips % world ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(n_clips, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_lambda_grad(params: Iterable[torch.nn.Parameter], average: bool = False, group=None, ring=None) -> None:
    """SUM (or mean) the .grad of the given lambd parameters over the process group, in place.

    Use SUM when each rank back-propagated the gradient of a global (summed) loss through its own
    shard — then the N-rank result equals the 1-rank result on the concatenated batch.
    All lambdas go out in one flat buffer so there is a single collective per step.  With `ring` (a PeerRing) the
    exchange is the library's own peer-memory kernel instead of the process group's all_reduce.
    """
    if not dist.is_available() or not dist.is_initialized():
        return
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return
    if ring is not None and all(p.grad.dtype == torch.float32 and p.grad.numel() == 1 for p in ps):
        # one-shot exchange over NVLink peer memory (PeerRing), one 32-thread launch per lambda
        for p in ps:
            g = p.grad.detach().view(1)
            ring.all_reduce_(g)
            if average:
                g /= ring.world
        return
    flat = torch.stack([p.grad.detach().reshape(()) for p in ps]) if len(ps) > 1 else ps[0].grad.detach().reshape(1)
    flat = flat.contiguous()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    for i, p in enumerate(ps):
        p.grad.copy_(flat[i].reshape(p.grad.shape))


def broadcast_lambda(params: Iterable[torch.nn.Parameter], src: int = 0, group=None) -> None:
    """Make lambd identical on all ranks (n_fft is derived from it, so it must be)."""
    if not dist.is_available() or not dist.is_initialized():
        return
    for p in params:
        dist.broadcast(p.data, src=src, group=group)


class PeerRing:
    """One-shot all-reduce of one fp32 over NVLink peer memory (include/dmel.h, dmel_peer_allreduce_f32).

    Every rank allocates a small zeroed buffer, the buffers are exchanged as CUDA-IPC handles through the
    process group, and the kernel then talks to all peers directly (stores + system-scope release/acquire flags)
    — one 32-thread launch instead of an NCCL call whose latency (~20 us at 8 ranks) is a fifth of a small
    DMEL step.  All ranks of the group must be GPUs of one box with peer access.  `PeerRing.create()` returns
    None (callers keep using NCCL) when the setup or its self-test against the group's all_reduce fails.
    Like an NCCL collective, a call waits for its peers for as long as it takes (DMEL_PEER_TIMEOUT_S, default
    1800 s, bounds it; on expiry the result is NaN on the ranks that gave up and `timed_out()` reports it).
    """

    def __init__(self, device: torch.device, group=None):
        """Collective constructor for callers that know every rank will succeed; `create()` is the safe entry point
        (same collective sequence on every rank whatever fails locally)."""
        self._init_local(device, group)
        err = self._alloc()
        gathered = self._exchange()
        if err is None:
            err = self._open(gathered)
        dist.barrier(group=group)
        if err is not None:
            self.close()
            raise err

    def _init_local(self, device: torch.device, group) -> None:
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        self._mapped: List[int] = []
        self._own = 0
        self._handle = b""
        self.bases = None

    def _alloc(self) -> Optional[Exception]:
        """Local: allocate this rank's buffer.  Returns the error instead of raising."""
        import ctypes
        try:
            lib = _lib.load()
            if lib.dmel_peer_buffer_bytes(self.world) == 0:
                raise _lib.DmelError(f"peer ring supports at most 32 ranks, got {self.world}")
            with torch.cuda.device(self.device):
                own = ctypes.c_void_p()
                handle = ctypes.create_string_buffer(64)
                _lib.check(lib.dmel_peer_alloc(self.world, ctypes.byref(own), handle), "dmel_peer_alloc")
                self._own = own.value
                self._handle = handle.raw
            return None
        except Exception as e:      # noqa: BLE001 - reported through the group, see create()
            return e

    def _exchange(self) -> List[Optional[bytes]]:
        """Collective (always runs): every rank's IPC handle, b"" for a rank whose allocation failed."""
        gathered: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(gathered, self._handle, group=self.group)
        return gathered

    def _open(self, gathered: List[Optional[bytes]]) -> Optional[Exception]:
        """Local: map the peers' buffers.  Returns the error instead of raising."""
        import ctypes
        try:
            if any(not h for h in gathered):
                raise _lib.DmelError("a peer could not allocate its buffer")
            lib = _lib.load()
            with torch.cuda.device(self.device):
                bases = []
                for r in range(self.world):
                    if r == self.rank:
                        bases.append(self._own)
                        continue
                    mapped = ctypes.c_void_p()
                    _lib.check(lib.dmel_peer_open(gathered[r], ctypes.byref(mapped)), "dmel_peer_open")
                    self._mapped.append(mapped.value)
                    bases.append(mapped.value)
                self.bases = torch.tensor(bases, dtype=torch.int64, device=self.device)
                torch.cuda.synchronize(self.device)
            return None
        except Exception as e:      # noqa: BLE001
            return e

    def close(self) -> None:
        lib = _lib.load()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for m in self._mapped:
                lib.dmel_peer_close(m)
            self._mapped = []
            if self._own:
                lib.dmel_peer_free(self._own)
                self._own = 0

    def all_reduce_(self, value: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks of a 1-element fp32 CUDA tensor, on the current stream."""
        if value.numel() != 1 or value.dtype != torch.float32 or value.device != self.device:
            raise ValueError("PeerRing reduces one fp32 element on its own device")
        lib = _lib.load()
        st = int(torch._C._cuda_getCurrentRawStream(self.device.index))
        _lib.check(lib.dmel_peer_allreduce_f32(value.data_ptr(), self.bases.data_ptr(), self.rank, self.world, st),
                   "dmel_peer_allreduce_f32")
        return value

    def timed_out(self) -> bool:
        """True if any call so far gave up waiting for a peer (its result was NaN). Synchronises the device."""
        import ctypes
        flag = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dmel_peer_timed_out(self._own, self.world, ctypes.byref(flag)), "dmel_peer_timed_out")
        return flag.value != 0

    @staticmethod
    def create(device: torch.device, group=None, trials: int = 4) -> Optional["PeerRing"]:
        """Build a ring and check it against the group's own all_reduce; None (ring closed, callers keep using the
        process group) if anything is off on ANY rank.  Every rank issues the same sequence of collectives whatever
        fails locally — all_gather_object, barrier, one MIN all-reduce of the set-up flag, `trials` x all_reduce, one
        MIN all-reduce of the verdict — so a local failure can neither hang nor mis-pair the others."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) < 2:
            return None
        device = torch.device(device)
        ring = PeerRing.__new__(PeerRing)
        ring._init_local(device, group)
        err = ring._alloc()
        gathered = ring._exchange()
        if err is None:
            err = ring._open(gathered)
        dist.barrier(group=group)
        ok = torch.ones(1, device=device) if err is None else torch.zeros(1, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)       # did every rank get its ring up?
        usable = bool(ok.item())
        good = torch.ones(1, device=device)
        g = torch.Generator(device=device).manual_seed(17 + ring.rank)
        for _ in range(trials):
            v = torch.randn(1, device=device, generator=g)
            want = v.clone()
            dist.all_reduce(want, group=group)
            if not usable:
                continue                      # the ring kernels need every rank; the group collective still ran
            try:
                got = ring.all_reduce_(v.clone())
                torch.cuda.synchronize(device)
                if ring.timed_out() or not torch.isfinite(got).all() or \
                        abs(float(got) - float(want)) > 1e-5 * max(1.0, abs(float(want))):
                    good.zero_()
            except Exception:       # noqa: BLE001
                good.zero_()
        dist.all_reduce(good, op=dist.ReduceOp.MIN, group=group)     # all ranks agree on the outcome
        if usable and bool(good.item()):
            return ring
        ring.close()
        return None
