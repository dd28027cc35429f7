"""Bit (de)composition and small host-side limb helpers.

``to_bits`` / ``from_bits`` keep the reference's scalar contract (``utils.py:6-38``: LSB first, and
``to_bits`` asserts ``integer < 2**bit_length``); the ``*_batch`` variants are the vectorised forms the
batched steps use.  The limb helpers below only move / compare / shift plaintext bookkeeping values
(r, z); every modular multiplication, inversion and exponentiation happens on the GPU."""
from __future__ import annotations

import secrets
from typing import List, Sequence

import numpy as np

from .limbs import int_to_limbs, ints_to_limbs


def to_bits(integer: int, bit_length: int) -> List[int]:
    """``bit_length`` bits of ``integer``, least significant first (contract of reference ``utils.py:6-21``, including its
    assertion that the integer fits)."""
    assert integer < (1 << bit_length)
    return [(integer >> position) & 1 for position in range(bit_length)]


def from_bits(bits: Sequence[int]) -> int:
    """Inverse of ``to_bits`` (reference ``utils.py:24-38``): bits are least significant first; only entries equal to 1 count."""
    return sum(1 << position for position, bit in enumerate(bits) if bit == 1)


def to_bits_batch(values: np.ndarray, bit_length: int) -> np.ndarray:
    """``[B]`` uint64 -> ``[B][bit_length]`` uint8, LSB first (same assertion as ``to_bits``)."""
    values = np.asarray(values, dtype=np.uint64)
    if bit_length < 64:
        assert bool(np.all(values < (np.uint64(1) << np.uint64(bit_length))))
    shifts = np.arange(bit_length, dtype=np.uint64)
    return ((values[:, None] >> shifts[None, :]) & np.uint64(1)).astype(np.uint8)


def from_bits_batch(bits: np.ndarray) -> np.ndarray:
    bits = np.asarray(bits, dtype=np.uint64)
    shifts = np.arange(bits.shape[1], dtype=np.uint64)
    return (bits << shifts[None, :]).sum(axis=1, dtype=np.uint64)


# ---------------------------------------------------------------------------------------------
# limb-array bookkeeping (plaintext values only)
# ---------------------------------------------------------------------------------------------
def low_bits(a: np.ndarray, l: int) -> np.ndarray:
    """``a mod 2**l`` for ``[B][k]`` limbs, l <= 64, as uint64."""
    lo = a[:, 0].astype(np.uint64)
    if a.shape[1] > 1:
        lo |= a[:, 1].astype(np.uint64) << np.uint64(32)
    if l < 64:
        lo &= (np.uint64(1) << np.uint64(l)) - np.uint64(1)
    return lo


def less_than_const(a: np.ndarray, const: int) -> np.ndarray:
    """Row-wise ``a < const`` for ``[B][k]`` limbs."""
    k = a.shape[1]
    if const >= 1 << (32 * k):
        return np.ones(a.shape[0], dtype=bool)
    if const <= 0:
        return np.zeros(a.shape[0], dtype=bool)
    c = int_to_limbs(const, k)
    diff = a != c[None, :]
    any_diff = diff.any(axis=1)
    top = k - 1 - np.argmax(diff[:, ::-1], axis=1)          # most significant differing limb
    rows = np.arange(a.shape[0])
    return any_diff & (a[rows, top] < c[top])


def shift_right(a: np.ndarray, l: int) -> np.ndarray:
    """Row-wise ``a >> l`` for ``[B][k]`` limbs (same limb count)."""
    B, k = a.shape
    ws, bs = l // 32, l % 32
    padded = np.zeros((B, k + ws + 1), dtype=np.uint64)
    padded[:, :k] = a
    lo = padded[:, ws:ws + k]
    hi = padded[:, ws + 1:ws + k + 1]
    if bs == 0:
        return lo.astype(np.uint32)
    return (((lo >> np.uint64(bs)) | (hi << np.uint64(32 - bs))) & np.uint64(0xFFFFFFFF)).astype(np.uint32)


def random_below(bound: int, count: int, k: int, minimum: int = 0) -> np.ndarray:
    """``count`` values uniform in ``[minimum, bound)`` from ``secrets`` (as the reference draws r, r_i)."""
    span = bound - minimum
    return ints_to_limbs([secrets.randbelow(span) + minimum for _ in range(count)], k)


def random_bits(bits: int, count: int, k: int) -> np.ndarray:
    return ints_to_limbs([secrets.randbits(bits) for _ in range(count)], k)


def random_permutations(count: int, n: int) -> np.ndarray:
    """Uniform permutations drawn with ``secrets`` (reference ``Initiator.shuffle``, initiator.py:212-226)."""
    rng = secrets.SystemRandom()
    out = np.empty((count, n), dtype=np.uint8)
    for i in range(count):
        p = list(range(n))
        rng.shuffle(p)
        out[i] = p
    return out
