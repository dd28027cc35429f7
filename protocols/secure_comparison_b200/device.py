"""Device-resident arrays and pinned host staging for the batched drivers.

``Initiator`` / ``KeyHolder.perform_secure_comparison`` keep a session's state (r, delta_A, z and the nonces drawn by the
device CSPRNG) in HBM between the party's phases and move only the protocol messages over PCIe.  The buffers come from
the library's own allocator (``scb_dev_alloc`` / ``scb_host_alloc``) through a small size-keyed cache per engine, so a
steady stream of sessions allocates nothing; every array a session holds is private to it (sessions may interleave at
their ``await`` points, reference ``test_secure_comparison.py:803-835``).
"""
from __future__ import annotations

import ctypes
import weakref
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native as nat


class DeviceArray:
    """A typed, shaped view of one device allocation; ``data_ptr()`` makes it usable wherever the ``*_dev`` calls take a
    device pointer."""

    def __init__(self, pool: "BufferPool", ptr: int, capacity: int, shape, dtype):
        self.pool, self.ptr, self.capacity = pool, ptr, capacity
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)

    @property
    def nbytes(self) -> int:
        return int(np.prod(self.shape)) * self.dtype.itemsize if self.shape else self.dtype.itemsize

    def data_ptr(self) -> int:
        return self.ptr

    def upload(self, host: np.ndarray) -> "DeviceArray":
        """Enqueue host -> device (``host`` must stay alive until the engine is next synchronised)."""
        host = np.ascontiguousarray(host, dtype=self.dtype)
        if host.nbytes != self.nbytes:
            raise ValueError(f"expected {self.nbytes} bytes for shape {self.shape}, got {host.nbytes} (shape {host.shape})")
        eng = self.pool.eng
        eng._check(eng.lib.scb_copy_h2d(eng.ctx, self.ptr, nat._ptr(host), host.nbytes))
        self.pool.h2d_bytes += host.nbytes
        self._keepalive = host
        return self

    def download(self, out: Optional[np.ndarray] = None, sync: bool = True) -> np.ndarray:
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype)
        elif out.nbytes != self.nbytes or not out.flags.c_contiguous:
            raise ValueError("download target must be C-contiguous and of the same size")
        eng = self.pool.eng
        eng._check(eng.lib.scb_copy_d2h(eng.ctx, nat._ptr(out), self.ptr, self.nbytes))
        self.pool.d2h_bytes += self.nbytes
        if sync:
            eng.sync()
        return out

    def release(self) -> None:
        """Back to the pool (the allocation stays cached for the next session)."""
        if self.ptr:
            self.pool._give_back_device(self.ptr, self.capacity)
            self.ptr = 0


class BufferPool:
    """Size-keyed free lists of device allocations and pinned host blocks of one engine."""

    def __init__(self, eng):
        self.eng = eng
        self._dev_free: Dict[int, List[int]] = {}
        self._dev_all: List[int] = []
        self._host_free: Dict[int, List[int]] = {}
        self._host_all: List[int] = []
        self.h2d_bytes = 0          # bytes moved by DeviceArray.upload / download since creation (bench.py reports them per step)
        self.d2h_bytes = 0

    @staticmethod
    def _bucket(nbytes: int) -> int:
        return max(256, 1 << (max(nbytes, 1) - 1).bit_length()) if nbytes < (1 << 20) else ((nbytes + (1 << 20) - 1) >> 20) << 20

    # ------------------------------------------------------------------ device
    def device(self, shape, dtype=np.uint32) -> DeviceArray:
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        cap = self._bucket(nbytes)
        free = self._dev_free.get(cap)
        if free:
            ptr = free.pop()
        else:
            out = ctypes.c_void_p()
            self.eng._check(self.eng.lib.scb_dev_alloc(self.eng.ctx, cap, ctypes.byref(out)))
            ptr = int(out.value)
            self._dev_all.append(ptr)
        return DeviceArray(self, ptr, cap, shape, dtype)

    def _give_back_device(self, ptr: int, cap: int) -> None:
        self._dev_free.setdefault(cap, []).append(ptr)

    # ------------------------------------------------------------------ pinned host
    def pinned(self, shape, dtype=np.uint32) -> Tuple[np.ndarray, Tuple[int, int]]:
        """A pinned host array and the token to hand back to ``release_pinned`` when its contents are no longer needed."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        cap = self._bucket(nbytes)
        free = self._host_free.get(cap)
        if free:
            ptr = free.pop()
        else:
            out = ctypes.c_void_p()
            if self.eng.lib.scb_host_alloc(cap, ctypes.byref(out)) != nat.SCB_OK or not out.value:
                raise MemoryError(f"scb_host_alloc({cap}) failed")
            ptr = int(out.value)
            self._host_all.append(ptr)
        buf = (ctypes.c_uint8 * cap).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr, (ptr, cap)

    def release_pinned(self, token: Tuple[int, int]) -> None:
        self._host_free.setdefault(token[1], []).append(token[0])

    def pinned_owned(self, shape, dtype=np.uint32) -> np.ndarray:
        """A pinned host array that goes back to the pool by itself when the last view of it (slices, memoryviews, arrays
        made with ``np.frombuffer`` on those) is dropped -- the lifetime of a protocol message: the sender fills it straight
        from the device, the transport or the receiver lets go of it."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        cap = self._bucket(nbytes)
        free = self._host_free.get(cap)
        if free:
            ptr = free.pop()
        else:
            out = ctypes.c_void_p()
            if self.eng.lib.scb_host_alloc(cap, ctypes.byref(out)) != nat.SCB_OK or not out.value:
                raise MemoryError(f"scb_host_alloc({cap}) failed")
            ptr = int(out.value)
            self._host_all.append(ptr)
        buf = (ctypes.c_uint8 * cap).from_address(ptr)
        weakref.finalize(buf, BufferPool._finalize_pinned, weakref.ref(self), ptr, cap)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    @staticmethod
    def _finalize_pinned(pool_ref, ptr: int, cap: int) -> None:
        pool = pool_ref()
        if pool is not None and ptr in pool._host_all:
            pool._host_free.setdefault(cap, []).append(ptr)

    def close(self) -> None:
        lib = self.eng.lib
        if getattr(self.eng, "ctx", None) is not None and self.eng.ctx.value:
            for ptr in self._dev_all:
                lib.scb_dev_free(self.eng.ctx, ptr)
        for ptr in self._host_all:
            lib.scb_host_free(ptr)
        self._dev_all.clear()
        self._host_all.clear()
        self._dev_free.clear()
        self._host_free.clear()
