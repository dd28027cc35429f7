"""Host-buffer throughput path: both parties' phases for a batch, chunked over several engine contexts.

``SchemeEngine``'s host-pointer calls are synchronous (H2D copy, kernels, D2H copy, wait).  Driving ONE context
from host buffers therefore leaves the GPU idle during every copy.  ``LocalPipeline`` owns ``n_contexts``
contexts (CUDA streams) on one GPU and runs the chunks of a batch on a small thread pool -- ctypes releases
the GIL inside every C-ABI call -- so the PCIe copies of one chunk overlap the kernels of another and the
host-buffer rate approaches the device-resident rate.  Result buffers are allocated once in pinned host memory
(``scb_host_alloc``) and reused.

This is the in-process equivalent of the reference's non-interactive flow (``README.md:117-141``: both parties'
steps called one after the other) with every ``randomize()`` of ``perform_secure_comparison`` applied.
"""
from __future__ import annotations

import ctypes
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional

import numpy as np

from . import _native as nat
from .schemes import SchemeEngine
from .sharding import shard_range


class PinnedArena:
    """Pinned host arrays owned by the library's allocator (``scb_host_alloc`` / ``scb_host_free``)."""

    def __init__(self):
        self.lib = nat.load_library()
        self._blocks: List[ctypes.c_void_p] = []

    def empty(self, shape, dtype=np.uint32) -> np.ndarray:
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = ctypes.c_void_p()
        if self.lib.scb_host_alloc(max(nbytes, 1), ctypes.byref(ptr)) != nat.SCB_OK or not ptr.value:
            raise MemoryError(f"scb_host_alloc({nbytes}) failed")
        self._blocks.append(ptr)
        buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def close(self) -> None:
        for ptr in self._blocks:
            self.lib.scb_host_free(ptr)
        self._blocks.clear()


class LocalPipeline:
    def __init__(self, scheme_paillier, scheme_dgk, l: int, device: int = 0, n_contexts: int = 2, chunks: Optional[int] = None,
                 devices: Optional[List[int]] = None):
        """``devices``: run on several GPUs of the box from ONE process -- ``n_contexts`` contexts per device, one worker
        thread per context, the batch sharded by comparison index with no device-to-device traffic.
        Default: the single GPU ``device``."""
        self.l = l
        self.engines = []
        for dev in (devices if devices is not None else [device]):
            for _ in range(n_contexts):
                eng = SchemeEngine(dev)
                eng.load_paillier(scheme_paillier)
                eng.load_dgk(scheme_dgk, l)
                self.engines.append(eng)
        self.chunks = chunks if chunks is not None else len(self.engines)      # one chunk per context measured best (tools/time_pipeline.py)
        self.pool = ThreadPoolExecutor(max_workers=len(self.engines))
        self.arena = PinnedArena()
        self._bufs: Dict[tuple, dict] = {}

    def close(self) -> None:
        self.pool.shutdown(wait=True)
        for eng in self.engines:
            eng.close()
        self.arena.close()

    def _buffers(self, slot: int, b: int) -> dict:
        key = (slot, b)
        if key not in self._bufs:
            eng, l, a = self.engines[0], self.l, self.arena
            self._bufs[key] = {
                "z_enc": a.empty((b, eng.kc)), "z": a.empty((b, eng.kN)), "db": a.empty((b, l + 1, eng.kn)), "c": a.empty((b, l + 1, eng.kn)),
                "delta_b": a.empty((b,), np.uint8), "enc3": a.empty((b, 3, eng.kc)),
            }
        return self._bufs[key]

    def _run_chunk(self, slot: int, eng: SchemeEngine, w: dict, lo: int, hi: int, result: np.ndarray) -> None:
        l = self.l
        buf = self._buffers(slot, hi - lo)
        s = slice(lo, hi)
        z_enc = eng.initiator_step1(w["x"][s], w["y"][s], w["r"][s], w["rho_z"][s], l, out=buf["z_enc"])               # A1
        z = eng.paillier_decrypt(z_enc, out=buf["z"])                                                                     # B1
        db = eng.keyholder_step4ab(z, l, w["rho_db"][s], out=buf["db"])
        c = eng.initiator_step4(db, w["r"][s], w["delta_a"][s], w["r_blind"][s], w["perm"][s], w["rho_c"][s], l, out=buf["c"])   # A2
        _, enc3 = eng.keyholder_step4j_5(c, z, w["rho_enc3"][s], l, out=(buf["delta_b"], buf["enc3"]))                  # B2
        eng.initiator_step6_7(enc3, w["r"][s], w["delta_a"][s], l, out=result[s])                                         # A3

    def compare(self, w: dict, result: Optional[np.ndarray] = None) -> np.ndarray:
        """``w``: host arrays ``x, y [B][2kN]``, ``r, rho_z [B][kN]``, ``delta_a [B]``, ``r_blind [B][l+1][rb]``, ``perm [B][l+1]``,
        ``rho_c, rho_db [B][l+1][rw]``, ``rho_enc3 [B][3][kN]`` (pinned for full copy speed).  Returns ``[[x <= y]]`` as ``[B][2kN]`` limbs."""
        B = w["x"].shape[0]
        eng0 = self.engines[0]
        if result is None:
            result = np.empty((B, eng0.kc), dtype=np.uint32)
        n = min(self.chunks, max(B, 1))
        ranges = [shard_range(B, j, n) for j in range(n)]

        def worker(slot: int) -> None:
            # a context is not re-entrant: each worker thread owns one engine and walks its own chunks in order
            for j in range(slot, n, len(self.engines)):
                lo, hi = ranges[j]
                if hi > lo:
                    self._run_chunk(j, self.engines[slot], w, lo, hi, result)

        futures = [self.pool.submit(worker, slot) for slot in range(len(self.engines))]
        for f in futures:
            f.result()
        return result

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.engines)
