"""Python int <-> little-endian uint32 limb arrays (the layout of the C ABI, include/scb200.h).

Conversion is vectorised through ``int.to_bytes`` / ``numpy.frombuffer``; nothing loops per limb in
Python.  A batch is a C-contiguous ``[B][k]`` ``uint32`` array, least significant limb first."""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np


def limbs_for_bits(bits: int) -> int:
    return (bits + 31) // 32


def int_to_limbs(value: int, k: int) -> np.ndarray:
    if value < 0:
        raise ValueError("limb arrays hold non-negative integers")
    return np.frombuffer(value.to_bytes(4 * k, "little"), dtype="<u4").astype(np.uint32)


def ints_to_limbs(values: Sequence[int], k: int) -> np.ndarray:
    """``[B]`` Python ints -> ``[B][k]`` uint32 (raises OverflowError when a value needs more than k limbs)."""
    n = 4 * k
    buf = b"".join(int(v).to_bytes(n, "little") for v in values)
    return np.frombuffer(buf, dtype="<u4").astype(np.uint32).reshape(len(values), k)


def limbs_to_int(row: np.ndarray) -> int:
    return int.from_bytes(np.ascontiguousarray(row, dtype="<u4").tobytes(), "little")


def limbs_to_ints(arr: np.ndarray) -> List[int]:
    """``[B][k]`` uint32 -> list of Python ints."""
    arr = np.ascontiguousarray(arr, dtype="<u4")
    if arr.ndim == 1:
        return [limbs_to_int(arr)]
    flat = arr.reshape(-1, arr.shape[-1])
    raw, n = flat.tobytes(), 4 * flat.shape[1]
    return [int.from_bytes(raw[i * n:(i + 1) * n], "little") for i in range(flat.shape[0])]


def random_limbs(rng: np.random.Generator, shape: Iterable[int]) -> np.ndarray:
    return rng.integers(0, 1 << 32, size=tuple(shape), dtype=np.uint64).astype(np.uint32)
