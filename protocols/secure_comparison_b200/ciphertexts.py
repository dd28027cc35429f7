"""Batched ciphertext containers with the reference's operator algebra.

The reference reaches all of its arithmetic through Python operators on ``PaillierCiphertext`` / ``DGKCiphertext``
objects of ``tno.mpc.encryption_schemes`` (``+ - * += *=`` at ``initiator.py:254-256, 320, 371, 406, 476-484, 512, 531,
561-564``).  ``BatchedPaillierCiphertext`` / ``BatchedDGKCiphertext`` are ``[..batch.., k]`` uint32 limb arrays (numpy
subclasses: slicing, stacking, ``np.where`` and the C-ABI calls treat them as the limb arrays they are) that also carry
the scheme and the ``fresh`` flag, and map the same operators onto the batched GPU primitives exactly as the scheme
templates do:

    a + b        -> mulmod          (scheme.add)             a + int  -> a + unsafe_encrypt(int)
    -a           -> modinv          (scheme.neg)             a - b    -> a + (-b);   int - a -> (-a) + int
    a * k        -> modexp          (scheme.mul; k < 0: (-a) * |k|;  k is one int, or one int per element)
    randomize()  -> a * rho^N mod N^2  /  a * h^rho mod n    (device CSPRNG unless ``rho`` is injected); sets ``fresh``

so code written against the reference's ciphertext types (e.g. the body of ``Initiator.step_4h``) runs on a whole batch.
Every ``step_*`` of the batched ``Initiator`` / ``KeyHolder`` returns these containers and accepts them (or bare limb
arrays, or sequences of scalar scheme ciphertext objects).  No arithmetic happens on the host."""
from __future__ import annotations

import warnings
from typing import Optional, Sequence, Union

import numpy as np

from . import _native as nat
from .limbs import int_to_limbs, ints_to_limbs, limbs_for_bits, limbs_to_ints

Plain = Union[int, np.integer, Sequence[int], np.ndarray]


class BatchedCiphertext(np.ndarray):
    """Common behaviour; subclasses bind the modulus slot and the scheme-specific ``unsafe_encrypt`` / randomizer."""

    _slot: int = -1
    scheme = None
    _fresh = False
    __array_priority__ = 100.0

    def __new__(cls, raw, scheme, fresh: bool = False):
        arr = np.ascontiguousarray(raw, dtype=np.uint32)
        if arr.ndim < 1:
            raise ValueError("raw must be a [..., k] limb array")
        obj = arr.view(cls)
        obj.scheme = scheme
        obj._fresh = bool(fresh)
        k = obj._engine().k32(cls._slot)
        if arr.shape[-1] != k:
            raise ValueError(f"ciphertexts of this scheme have {k} limbs, got {arr.shape[-1]}")
        return obj

    def __array_finalize__(self, obj) -> None:
        if obj is None:
            return
        self.scheme = getattr(obj, "scheme", None)
        self._fresh = getattr(obj, "_fresh", False)

    # ------------------------------------------------------------------ container protocol
    def _engine(self):
        raise NotImplementedError

    def _like(self, raw, fresh: bool = False):
        out = np.ascontiguousarray(raw, dtype=np.uint32).view(type(self))
        out.scheme, out._fresh = self.scheme, fresh
        return out

    @property
    def fresh(self) -> bool:
        return self._fresh

    @property
    def batch_shape(self):
        return self.shape[:-1]

    def peek_value(self) -> np.ndarray:
        """The raw values (plain limb array view), without touching the ``fresh`` flag."""
        return self.view(np.ndarray)

    def get_value(self) -> np.ndarray:
        """The raw values; like the reference's ciphertexts, reading them for use marks the randomness as spent."""
        self._fresh = False
        return self.view(np.ndarray)

    def to_ints(self):
        """Raw values as (nested) Python ints -- for interoperability with scalar scheme objects."""
        flat = limbs_to_ints(self.view(np.ndarray).reshape(-1, self.shape[-1]))
        return flat[0] if self.ndim == 1 else np.array(flat, dtype=object).reshape(self.batch_shape).tolist()

    def __repr__(self) -> str:
        return f"{type(self).__name__}(batch={self.batch_shape}, fresh={self._fresh})"

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _flat(raw: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(raw).reshape(-1, raw.shape[-1]))

    def _use(self) -> np.ndarray:
        if self._fresh:
            warnings.warn("arithmetic on a fresh ciphertext: its randomness is spent (randomize again before sending)", UserWarning, stacklevel=3)
        self._fresh = False
        return self.view(np.ndarray)

    def _coerce(self, other):
        """Other operand of ``add``: a ciphertext of the same scheme, or plaintext(s) to ``unsafe_encrypt`` first."""
        if isinstance(other, BatchedCiphertext):
            if type(other) is not type(self):
                raise TypeError("cannot combine ciphertexts of different schemes")
            return other
        if hasattr(other, "peek_value"):                      # a scalar scheme ciphertext object
            return self._like(int_to_limbs(int(other.peek_value()), self.shape[-1]))
        return self._encrypt_plain(other)

    def _encrypt_plain(self, plain: Plain):
        raise NotImplementedError

    # ------------------------------------------------------------------ scheme.add / neg / mul
    def __add__(self, other):
        o = self._coerce(other)
        a, b = self._use(), o._use()
        shape = np.broadcast_shapes(a.shape[:-1], b.shape[:-1]) + (a.shape[-1],)
        eng = self._engine()
        out = eng.modmul(self._slot, self._flat(np.broadcast_to(a, shape)), self._flat(np.broadcast_to(b, shape)))
        return self._like(out.reshape(shape))

    __radd__ = __add__

    def __iadd__(self, other):
        res = self + other
        if res.shape != self.shape:
            raise ValueError("in-place addition cannot change the batch shape")
        self.view(np.ndarray)[...] = res.view(np.ndarray)
        self._fresh = False
        return self

    def __neg__(self):
        raw = self._use()
        return self._like(self._engine().modinv(self._slot, self._flat(raw)).reshape(raw.shape))

    def __sub__(self, other):
        return self + (-self._coerce(other))

    def __rsub__(self, other):
        return (-self) + other

    def __mul__(self, k):
        if isinstance(k, BatchedCiphertext):
            raise TypeError("a ciphertext can only be multiplied by plaintext integers")
        eng = self._engine()
        raw = self._use()
        kk = raw.shape[-1]
        if isinstance(k, (int, np.integer)):
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
            base = np.where(neg[:, None], eng.modinv(self._slot, base), base)
        mags = [abs(v) for v in flat_k]
        bits = max(1, max((m.bit_length() for m in mags), default=1))
        out = eng.modexp(self._slot, np.ascontiguousarray(base), ints_to_limbs(mags, limbs_for_bits(bits)), exp_bits=bits)
        return self._like(out.reshape(shape + (kk,)))

    __rmul__ = __mul__

    def __imul__(self, k):
        res = self * k
        if res.shape != self.shape:
            raise ValueError("in-place multiplication cannot change the batch shape")
        self.view(np.ndarray)[...] = res.view(np.ndarray)
        self._fresh = False
        return self


class BatchedPaillierCiphertext(BatchedCiphertext):
    """Batch of Paillier ciphertexts: raw values mod N^2, ``[..., 2 kN]`` limbs."""

    _slot = nat.MOD_PAILLIER_N2

    def _engine(self):
        from .schemes import get_engine
        return get_engine(self.scheme)

    def _encrypt_plain(self, plain: Plain):
        """``scheme.unsafe_encrypt(plain)`` with the integer encoding (negatives wrap mod N): raw value 1 + m N."""
        eng = self._engine()
        vals = np.asarray(plain, dtype=object)
        raw = eng.paillier_encrypt(eng.encode_paillier_plain(vals))
        return self._like(raw.reshape(vals.shape + (eng.kc,)))

    def randomize(self, rho: Optional[np.ndarray] = None):
        """``c <- c * rho^N mod N^2`` in place (reference ``randomize()``, U2); ``rho``: ``[..., kN]`` limbs in [1, N), drawn by
        the device CSPRNG when omitted.  Returns ``self``."""
        eng = self._engine()
        raw = self.view(np.ndarray)
        M = int(np.prod(self.batch_shape)) if self.batch_shape else 1
        if rho is None:
            eng.ensure_rng()
            rho = eng.rng_below(eng.paillier_n, M, eng.kN, min_one=True)
        wide = np.zeros((M, eng.kc), dtype=np.uint32)
        wide[:, :eng.kN] = np.ascontiguousarray(rho, dtype=np.uint32).reshape(M, eng.kN)
        rn = eng.modexp(self._slot, wide, int_to_limbs(eng.paillier_n, eng.kN), exp_bits=eng.paillier_n.bit_length())
        raw[...] = eng.modmul(self._slot, self._flat(raw), rn).reshape(raw.shape)
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
        flat = [int(v) for v in vals.reshape(-1)]
        table = {0: int_to_limbs(1, kn), 1: int_to_limbs(eng.dgk_g, kn), -1: int_to_limbs(eng.dgk_g_neg, kn)}
        if all(v in table for v in flat):
            raw = np.stack([table[v] for v in flat])
        else:
            g = self._like(np.broadcast_to(table[1], (len(flat), kn)).copy())
            raw = (g * np.array(flat, dtype=object)).view(np.ndarray)
        return self._like(raw.reshape(vals.shape + (kn,)))

    def randomize(self, rho: Optional[np.ndarray] = None):
        """``c <- c * h^rho mod n`` in place (reference ``randomize()``, U6) through the fixed-base table of h."""
        eng = self._engine()
        raw = self.view(np.ndarray)
        M = int(np.prod(self.batch_shape)) if self.batch_shape else 1
        if rho is None:
            eng.ensure_rng()
            rho = eng.rng_bits(eng._fixed_base[0][1], M, eng.rand_words)
        rho = np.ascontiguousarray(rho, dtype=np.uint32).reshape(M, -1)
        raw[...] = eng.fixed_base_exp(0, rho, mul_in=self._flat(raw)).reshape(raw.shape)
        self._fresh = True
        return self


def paillier_batch(raw, scheme, fresh: bool = False) -> BatchedPaillierCiphertext:
    """Wrap a limb array produced by the engine (no copy, no validation beyond the limb count)."""
    out = np.asarray(raw, dtype=np.uint32).view(BatchedPaillierCiphertext)
    out.scheme, out._fresh = scheme, fresh
    return out


def dgk_batch(raw, scheme, fresh: bool = False) -> BatchedDGKCiphertext:
    out = np.asarray(raw, dtype=np.uint32).view(BatchedDGKCiphertext)
    out.scheme, out._fresh = scheme, fresh
    return out
