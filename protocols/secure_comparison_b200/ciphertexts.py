"""Batched ciphertext containers with the reference's operator algebra.

The reference reaches all of its arithmetic through Python operators on ``PaillierCiphertext`` / ``DGKCiphertext``
objects of ``tno.mpc.encryption_schemes`` (``+ - * += *=`` at ``initiator.py:254-256, 320, 371, 406, 476-484, 512, 531,
561-564``).  ``BatchedPaillierCiphertext`` / ``BatchedDGKCiphertext`` hold a whole batch of raw ciphertext values as one
``[..batch.., k]`` uint32 limb array plus the scheme and the ``fresh`` flag, and map the same operators onto the batched
GPU primitives exactly as the scheme templates do:

    a + b        -> mulmod          (scheme.add)             a + int  -> a + unsafe_encrypt(int)
    -a           -> modinv          (scheme.neg)             a - b    -> a + (-b);   int - a -> (-a) + int
    a * k        -> modexp          (scheme.mul; k < 0: (-a) * |k|;  k may be one int or one int per element)
    randomize()  -> a * rho^N mod N^2  /  a * h^rho mod n    (device CSPRNG unless ``rho`` is injected); sets ``fresh``

so code written against the reference's ciphertext types (e.g. the body of ``Initiator.step_4h``) runs on a batch.
Every ``step_*`` of the batched ``Initiator`` / ``KeyHolder`` returns these containers and accepts them (or bare limb
arrays, or sequences of scheme ciphertext objects).  No arithmetic happens on the host."""
from __future__ import annotations

import warnings
from typing import Optional, Sequence, Union

import numpy as np

from . import _native as nat
from .limbs import int_to_limbs, ints_to_limbs, limbs_for_bits, limbs_to_ints

Plain = Union[int, np.integer, Sequence[int], np.ndarray]


def _is_plain_scalar(v) -> bool:
    return isinstance(v, (int, np.integer)) and not isinstance(v, bool) or isinstance(v, bool)


class BatchedCiphertext:
    """Common behaviour; subclasses bind the modulus slot and the scheme-specific ``unsafe_encrypt`` / randomizer."""

    _slot: int = -1

    def __init__(self, raw, scheme, fresh: bool = False):
        raw = np.ascontiguousarray(raw, dtype=np.uint32)
        if raw.ndim < 1:
            raise ValueError("raw must be a [..., k] limb array")
        self._raw = raw
        self.scheme = scheme
        self._fresh = bool(fresh)
        k = self._engine().k32(self._slot)
        if raw.shape[-1] != k:
            raise ValueError(f"ciphertexts of this scheme have {k} limbs, got {raw.shape[-1]}")

    # ------------------------------------------------------------------ container protocol
    def _engine(self):
        raise NotImplementedError

    def _like(self, raw, fresh: bool = False):
        out = object.__new__(type(self))
        out._raw, out.scheme, out._fresh = np.ascontiguousarray(raw, dtype=np.uint32), self.scheme, fresh
        return out

    @property
    def fresh(self) -> bool:
        return self._fresh

    @property
    def shape(self):
        """Batch shape (the limb axis is not part of it)."""
        return self._raw.shape[:-1]

    def __len__(self) -> int:
        if self._raw.ndim == 1:
            raise TypeError("a single ciphertext has no length")
        return self._raw.shape[0]

    def __getitem__(self, idx):
        sub = self._raw[idx]
        if sub.ndim == 0 or sub.shape[-1:] != self._raw.shape[-1:]:
            raise IndexError("index the batch axes only, not the limbs")
        return self._like(sub, self._fresh)

    def __setitem__(self, idx, value) -> None:
        self._raw[idx] = self._coerce(value)._raw
        self._fresh = False

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __array__(self, dtype=None, copy=None):
        return self._raw if dtype is None else self._raw.astype(dtype, copy=False)

    def peek_value(self) -> np.ndarray:
        """The raw values (limb array), without touching the ``fresh`` flag."""
        return self._raw

    def get_value(self) -> np.ndarray:
        """The raw values; like the reference's ciphertexts, reading them for use marks the randomness as spent."""
        self._fresh = False
        return self._raw

    def to_ints(self):
        """Raw values as (nested) Python ints -- for interoperability with scalar scheme objects."""
        flat = limbs_to_ints(self._raw.reshape(-1, self._raw.shape[-1]))
        return flat[0] if self._raw.ndim == 1 else np.array(flat, dtype=object).reshape(self.shape).tolist()

    def copy(self):
        return self._like(self._raw.copy(), self._fresh)

    def __repr__(self) -> str:
        return f"{type(self).__name__}(batch={self.shape}, fresh={self._fresh})"

    # ------------------------------------------------------------------ helpers
    def _flat(self, raw: Optional[np.ndarray] = None) -> np.ndarray:
        raw = self._raw if raw is None else raw
        return np.ascontiguousarray(raw.reshape(-1, raw.shape[-1]))

    def _use(self) -> np.ndarray:
        if self._fresh:
            warnings.warn("arithmetic on a fresh ciphertext: its randomness is spent (randomize again before sending)", UserWarning, stacklevel=3)
        self._fresh = False
        return self._raw

    def _coerce(self, other):
        """Other operand of ``add``: a ciphertext of the same scheme, or plaintext(s) to ``unsafe_encrypt`` first."""
        if isinstance(other, BatchedCiphertext):
            if type(other) is not type(self):
                raise TypeError("cannot combine ciphertexts of different schemes")
            return other
        if hasattr(other, "peek_value"):                      # a scalar scheme ciphertext object
            return self._like(int_to_limbs(int(other.peek_value()), self._raw.shape[-1]))
        return self._encrypt_plain(other)

    def _encrypt_plain(self, plain: Plain):
        raise NotImplementedError

    def _broadcast(self, a: np.ndarray, b: np.ndarray):
        shape = np.broadcast_shapes(a.shape[:-1], b.shape[:-1]) + (a.shape[-1],)
        return np.broadcast_to(a, shape), np.broadcast_to(b, shape), shape

    # ------------------------------------------------------------------ scheme.add / neg / mul
    def __add__(self, other):
        o = self._coerce(other)
        a, b, shape = self._broadcast(self._use(), o._use() if o is not self else self._raw)
        eng = self._engine()
        out = eng.modmul(self._slot, self._flat(a), self._flat(b))
        return self._like(out.reshape(shape))

    __radd__ = __add__

    def __iadd__(self, other):
        res = self + other
        self._raw, self._fresh = res._raw, False
        return self

    def __neg__(self):
        eng = self._engine()
        return self._like(eng.modinv(self._slot, self._flat(self._use())).reshape(self._raw.shape))

    def __sub__(self, other):
        o = self._coerce(other)
        return self + (-o)

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, k):
        eng = self._engine()
        raw = self._use()
        kk = raw.shape[-1]
        if _is_plain_scalar(k):
            k = int(k)
            if k == 0:
                return self._like(np.broadcast_to(int_to_limbs(1, kk), raw.shape).copy())
            base = self._flat(raw)
            if k < 0:
                base = eng.modinv(self._slot, base)
                k = -k
            if k == 1:
                return self._like(base.reshape(raw.shape).copy())
            ew = limbs_for_bits(k.bit_length())
            out = eng.modexp(self._slot, base, int_to_limbs(k, ew), exp_bits=k.bit_length())
            return self._like(out.reshape(raw.shape))
        ks = np.asarray(k, dtype=object)
        shape = np.broadcast_shapes(raw.shape[:-1], ks.shape)
        base = self._flat(np.broadcast_to(raw, shape + (kk,)))
        flat_k = [int(v) for v in np.broadcast_to(ks, shape).reshape(-1)]
        neg = np.array([v < 0 for v in flat_k], dtype=bool)
        if neg.any():
            inv = eng.modinv(self._slot, base)
            base = np.where(neg[:, None], inv, base)
        mags = [abs(v) for v in flat_k]
        bits = max(1, max(m.bit_length() for m in mags)) if mags else 1
        ew = limbs_for_bits(bits)
        out = eng.modexp(self._slot, np.ascontiguousarray(base), ints_to_limbs(mags, ew), exp_bits=bits)
        return self._like(out.reshape(shape + (kk,)))

    __rmul__ = __mul__

    def __imul__(self, k):
        res = self * k
        self._raw, self._fresh = res._raw, False
        return self

    def __eq__(self, other):
        if isinstance(other, BatchedCiphertext):
            return type(other) is type(self) and self._raw.shape == other._raw.shape and bool(np.array_equal(self._raw, other._raw))
        return NotImplemented

    __hash__ = None


class BatchedPaillierCiphertext(BatchedCiphertext):
    """Batch of Paillier ciphertexts: raw values mod N^2, ``[..., 2 kN]`` limbs."""

    _slot = nat.MOD_PAILLIER_N2

    def _engine(self):
        from .schemes import get_engine
        return get_engine(self.scheme)

    def _encrypt_plain(self, plain: Plain):
        """``scheme.unsafe_encrypt(plain)`` with the integer encoding (negatives wrap mod N): raw value 1 + m N."""
        eng = self._engine()
        N = eng.paillier_n
        vals = np.asarray(plain, dtype=object)
        flat = [int(v) % N for v in vals.reshape(-1)] if vals.ndim else [int(vals) % N]
        raw = eng.paillier_encrypt(ints_to_limbs(flat, eng.kN))
        return self._like(raw.reshape(vals.shape + (eng.kc,)))

    def randomize(self, rho: Optional[np.ndarray] = None):
        """``c <- c * rho^N mod N^2`` (reference ``randomize()``, U2); ``rho``: ``[..., kN]`` limbs in [1, N), drawn by the
        device CSPRNG when omitted.  Returns ``self`` (in place, like the reference)."""
        eng = self._engine()
        M = int(np.prod(self.shape)) if self.shape else 1
        if rho is None:
            rho = eng.rng_below(eng.paillier_n, M, eng.kN, min_one=True)
        rho = np.ascontiguousarray(rho, dtype=np.uint32).reshape(M, eng.kN)
        exp = int_to_limbs(eng.paillier_n, eng.kN)
        wide = np.zeros((M, eng.kc), dtype=np.uint32)
        wide[:, :eng.kN] = rho
        rn = eng.modexp(self._slot, wide, exp, exp_bits=eng.paillier_n.bit_length())
        self._raw = eng.modmul(self._slot, self._flat(), rn).reshape(self._raw.shape)
        self._fresh = True
        return self


class BatchedDGKCiphertext(BatchedCiphertext):
    """Batch of DGK ciphertexts: raw values mod n, ``[..., kn]`` limbs."""

    _slot = nat.MOD_DGK_N

    def _engine(self):
        from .schemes import get_engine
        return get_engine(None, self.scheme)

    def _encrypt_plain(self, plain: Plain):
        """``scheme.unsafe_encrypt(plain)``: g^m mod n; negative m through the inverse of g (``pow_mod``)."""
        eng = self._engine()
        kn = eng.kn
        vals = np.asarray(plain, dtype=object)
        flat = [int(v) for v in vals.reshape(-1)] if vals.ndim else [int(vals)]
        table = {0: int_to_limbs(1, kn), 1: int_to_limbs(eng.dgk_g, kn), -1: int_to_limbs(eng.dgk_g_neg, kn)}
        if all(v in table for v in flat):
            raw = np.stack([table[v] for v in flat])
        else:
            g = self._like(np.broadcast_to(table[1], (len(flat), kn)).copy())
            raw = (g * np.array(flat, dtype=object))._raw
        return self._like(raw.reshape(vals.shape + (kn,)))

    def randomize(self, rho: Optional[np.ndarray] = None):
        """``c <- c * h^rho mod n`` (reference ``randomize()``, U6) through the fixed-base table of h; ``rho``: ``[..., words]``."""
        eng = self._engine()
        M = int(np.prod(self.shape)) if self.shape else 1
        bits = eng._fixed_base[0][1]
        if rho is None:
            rho = eng.rng_bits(bits, M, eng.rand_words)
        rho = np.ascontiguousarray(rho, dtype=np.uint32).reshape(M, -1)
        self._raw = eng.fixed_base_exp(0, rho, mul_in=self._flat()).reshape(self._raw.shape)
        self._fresh = True
        return self


def raw_limbs(values, k: int) -> Optional[np.ndarray]:
    """Limb array of a batched container (None when ``values`` is not one)."""
    if isinstance(values, BatchedCiphertext):
        raw = values.peek_value()
        if raw.shape[-1] != k:
            raise ValueError(f"limb arrays for this modulus need {k} limbs per value, got {raw.shape[-1]}")
        return raw
    return None
