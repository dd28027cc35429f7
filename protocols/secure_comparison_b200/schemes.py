"""Scheme-level engine: key loading from (duck-typed) Paillier / DGK scheme objects and the batched
scheme operations + fused protocol phases of the C ABI.

The scheme objects are consumed through exactly the attribute names the reference uses
(``paillier.public_key.{n, n_squared, g}``, ``paillier.secret_key.{lambda_, mu}``,
``dgk.public_key.{n, g, h, u, t}``, ``dgk.secret_key.{v_p, p}``), so real
``tno.mpc.encryption_schemes`` objects drop in when that package is installed.  Nothing here computes
a modular exponentiation / inverse / product on the host: the host only derives one-off key constants
(e.g. the raw ciphertext of plaintext -1) and moves limb arrays.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as nat
from .engine import Engine
from .limbs import int_to_limbs, ints_to_limbs, limbs_for_bits


def _raw(ct) -> int:
    """Raw integer of a ciphertext-like object (scheme ciphertext, or a plain int)."""
    if isinstance(ct, (int, np.integer)):
        return int(ct)
    if hasattr(ct, "peek_value"):
        return int(ct.peek_value())
    if hasattr(ct, "get_value"):
        return int(ct.get_value())
    raise TypeError(f"cannot read a raw ciphertext value from {type(ct).__name__}")


def _paillier_primes(sk, n: int, lam: int):
    """(p, q) of a Paillier secret key, or None when they cannot be recovered (lambda = lcm variants)."""
    import math
    p, q = getattr(sk, "p", None), getattr(sk, "q", None)
    if p is not None and q is not None and int(p) * int(q) == n:
        return int(p), int(q)
    s = n + 1 - lam                      # p + q when lambda = (p - 1)(q - 1)
    disc = s * s - 4 * n
    if disc < 0:
        return None
    root = math.isqrt(disc)
    if root * root != disc or (s + root) % 2:
        return None
    p, q = (s + root) // 2, (s - root) // 2
    return (p, q) if p * q == n and q > 1 else None


def as_limbs(values, k: int) -> np.ndarray:
    """Batch container -> ``[..][k]`` uint32 limbs.  Accepts batched ciphertext containers (``ciphertexts.py``), limb
    arrays (returned as is), a single scheme ciphertext / raw int (batch of one), or (nested) sequences of those."""
    if isinstance(values, np.ndarray) and values.dtype == np.uint32:
        if values.shape[-1] != k:
            raise ValueError(f"limb arrays for this modulus need {k} limbs per value, got {values.shape[-1]}")
        return np.ascontiguousarray(values)          # (containers are uint32 ndarray subclasses: this drops the wrapper)
    if isinstance(values, (int, np.integer)) or hasattr(values, "peek_value"):
        return ints_to_limbs([_raw(values)], k)
    if isinstance(values, np.ndarray):
        raise TypeError("limb arrays must have dtype uint32")
    values = list(values)
    if values and isinstance(values[0], (list, tuple, np.ndarray)):
        return np.stack([as_limbs(v, k) for v in values])
    return ints_to_limbs([_raw(v) for v in values], k)


def is_ciphertext_batch(values) -> bool:
    """True for what the drivers treat as ENCRYPTED input (batched containers, uint32 limb arrays, scheme ciphertext objects or
    sequences of them); everything else numeric is a plaintext to encrypt first (reference ``initiator.py:93-102``)."""
    from .ciphertexts import BatchedCiphertext
    if isinstance(values, BatchedCiphertext) or hasattr(values, "peek_value"):
        return True
    if isinstance(values, np.ndarray):
        return values.dtype == np.uint32 and values.ndim >= 2
    if isinstance(values, (list, tuple)) and values:
        return is_ciphertext_batch(values[0])
    return False


def _check_shape(name: str, arr: np.ndarray, shape) -> np.ndarray:
    if tuple(arr.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(arr.shape)}")
    return arr


def _dgk_raw_of_minus_one(dgk) -> Optional[int]:
    """The raw ciphertext the scheme itself produces for plaintext -1 without encoding
    (``scheme_dgk.unsafe_encrypt(s, apply_encoding=False)`` with s = -1, reference initiator.py:460).
    Asking the scheme keeps us bit-compatible whatever its negative-exponent convention is."""
    if hasattr(dgk, "unsafe_encrypt"):
        return _raw(dgk.unsafe_encrypt(-1, apply_encoding=False))
    return None      # plain key holders: the engine derives g^-1 mod n on the device (pow_mod's convention, U4)


def _out(out, shape, dtype=np.uint32) -> np.ndarray:
    """Caller-provided result buffer (e.g. pinned host memory reused across calls) or a fresh array."""
    if out is None:
        return np.empty(shape, dtype=dtype)
    if out.dtype != dtype or tuple(out.shape) != tuple(shape) or not out.flags.c_contiguous:
        raise ValueError(f"out must be a C-contiguous {np.dtype(dtype).name} array of shape {tuple(shape)}")
    return out


class SchemeEngine(Engine):
    """One GPU context with a Paillier and a DGK key loaded."""

    def __init__(self, device: int = 0):
        super().__init__(device)
        self.paillier_n: Optional[int] = None
        self.dgk_n: Optional[int] = None
        self.kN = self.kc = self.kn = 0
        self.l = 0
        self.u = 0
        self.rand_words = 0
        self.rb_words = 0
        self.has_paillier_secret = False
        self.has_paillier_crt = False
        self.has_dgk_secret = False
        from .device import BufferPool
        self.pool = BufferPool(self)
        self._rng_seeded = False

    def close(self) -> None:
        if getattr(self, "pool", None) is not None:
            self.pool.close()
        super().close()

    def rng_seed(self, key=None, nonce=None) -> None:
        super().rng_seed(key, nonce)
        self._rng_seeded = True

    def ensure_rng(self) -> None:
        """Key the device CSPRNG from ``secrets`` once (the drivers draw every nonce on the device by default)."""
        if not self._rng_seeded:
            self.rng_seed()

    def encode_paillier_plain(self, values) -> np.ndarray:
        """Plaintext integers -> ``[B][kN]`` limbs of the Paillier plaintext space: the integer encoding of the reference's
        ``scheme.unsafe_encrypt(x)`` (negative values wrap modulo N, ``test_secure_comparison.py:27-37`` uses -400, -1)."""
        N = self.paillier_n
        flat = np.asarray(values, dtype=object).reshape(-1)
        out = []
        for v in flat:
            if isinstance(v, (float, np.floating)):
                if not float(v).is_integer():
                    raise ValueError("only integral plaintexts can be compared")
                v = int(v)
            out.append(int(v) % N)
        return ints_to_limbs(out, self.kN)

    # ------------------------------------------------------------------ keys
    def load_paillier(self, scheme) -> None:
        pk = scheme.public_key
        if getattr(pk, "g", pk.n + 1) != pk.n + 1:
            raise ValueError("the engine relies on g = n + 1 (reference U1)")
        n = int(pk.n)
        if self.paillier_n != n:
            kN = limbs_for_bits(n.bit_length())
            n_w = int_to_limbs(n, kN)                      # keep the arrays alive across the C call
            self._check(self.lib.scb_paillier_set_public(self.ctx, nat._ptr(n_w), kN))
            self.paillier_n, self.kN, self.kc = n, kN, 2 * kN
            self._moduli[nat.MOD_PAILLIER_N2] = (n * n, 2 * kN)
            self._moduli[nat.MOD_PAILLIER_N] = (n, kN)
            self.has_paillier_secret = False
        sk = getattr(scheme, "secret_key", None)
        if sk is not None and not self.has_paillier_secret:
            lam, mu = int(sk.lambda_), int(sk.mu)
            lw = limbs_for_bits(lam.bit_length())
            lam_w, mu_w = int_to_limbs(lam, lw), int_to_limbs(mu, self.kN)
            self._check(self.lib.scb_paillier_set_private(self.ctx, nat._ptr(lam_w), lw, nat._ptr(mu_w)))
            self.has_paillier_secret = True
            # CRT form for the KeyHolder's Paillier work (decrypt, randomize): needs the primes.  The upstream
            # secret key stores (lambda_, mu, n); with lambda = (p-1)(q-1) the primes are the roots of
            # X^2 - (N + 1 - lambda) X + N.  A one-off key-setup computation, like key generation itself.
            pq = _paillier_primes(sk, n, lam)
            self.has_paillier_crt = False
            if pq is not None and self.kN % 2 == 0 and all(limbs_for_bits(v.bit_length()) <= self.kN // 2 for v in pq):
                kP = self.kN // 2
                p_w, q_w = int_to_limbs(pq[0], kP), int_to_limbs(pq[1], kP)
                self._check(self.lib.scb_paillier_set_private_crt(self.ctx, nat._ptr(p_w), nat._ptr(q_w), kP))
                self.has_paillier_crt = True

    def load_dgk(self, scheme, l: int, rand_bits: Optional[int] = None) -> None:
        pk = scheme.public_key
        n, u = int(pk.n), int(pk.u)
        if rand_bits is None:
            rand_bits = int(getattr(scheme, "randomness_bits", (5 * int(pk.t)) // 2))
        if self.dgk_n != n or self.l != l or self.rand_words != limbs_for_bits(rand_bits):
            kn = limbs_for_bits(n.bit_length())
            g_neg = _dgk_raw_of_minus_one(scheme)
            n_w, g_w, h_w = (int_to_limbs(v, kn) for v in (n, int(pk.g), int(pk.h)))
            gn_w = None if g_neg is None else int_to_limbs(g_neg, kn)
            self._check(self.lib.scb_dgk_set_public(self.ctx, nat._ptr(n_w), nat._ptr(g_w), nat._ptr(h_w), nat._ptr(gn_w), kn,
                                                    u.bit_length(), l, rand_bits))
            self.dgk_n, self.kn, self.l, self.u = n, kn, l, u
            self._moduli[nat.MOD_DGK_N] = (n, kn)
            if g_neg is None:
                from .limbs import limbs_to_int
                g_neg = limbs_to_int(self.modinv(nat.MOD_DGK_N, g_w.reshape(1, kn))[0])
            self.dgk_g, self.dgk_g_neg = int(pk.g), g_neg
            self.rand_words = limbs_for_bits(rand_bits)
            self.rb_words = limbs_for_bits(u.bit_length())
            self._moduli[nat.MOD_DGK_N] = (n, kn)
            self._fixed_base[0] = (nat.MOD_DGK_N, rand_bits)
            self.has_dgk_secret = False
        sk = getattr(scheme, "secret_key", None)
        if sk is not None and not self.has_dgk_secret:
            p, v_p = int(sk.p), int(sk.v_p)
            kp = limbs_for_bits(p.bit_length())
            vw = limbs_for_bits(v_p.bit_length())
            p_w, vp_w = int_to_limbs(p, kp), int_to_limbs(v_p, vw)
            self._check(self.lib.scb_dgk_set_private(self.ctx, nat._ptr(p_w), kp, nat._ptr(vp_w), vw, v_p.bit_length()))
            self._moduli[nat.MOD_DGK_P] = (p, kp)
            self.has_dgk_secret = True

    def set_paillier_crt(self, enabled: bool) -> None:
        """Switch the KeyHolder's Paillier work between the CRT form (default when the primes are known) and the
        public formulas c^lambda mod N^2 / rho^N mod N^2.  Results are bit-identical either way."""
        self._check(self.lib.scb_set_option(self.ctx, nat.OPT_PAILLIER_CRT, int(bool(enabled))))

    # ------------------------------------------------------------------ scheme operations (host buffers)
    def paillier_encrypt(self, plain: np.ndarray, rho: Optional[np.ndarray] = None) -> np.ndarray:
        """``(1 + m N) * rho^N mod N^2`` -- ``unsafe_encrypt(m, apply_encoding=False)`` [+ ``randomize()``]."""
        plain = nat._host_u32(plain)
        if plain.ndim < 1 or plain.shape[-1] != self.kN:
            raise ValueError(f"plaintexts must be [..][{self.kN}] limbs")
        shape = plain.shape[:-1]
        M = int(np.prod(shape)) if shape else 1
        rho_a = None if rho is None else nat._host_u32(rho, plain.shape)
        out = np.empty(shape + (self.kc,), dtype=np.uint32)
        self._check(self.lib.scb_paillier_encrypt(self.ctx, nat._ptr(plain), nat._ptr(rho_a), nat._ptr(out), M))
        return out

    def paillier_decrypt(self, c: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Raw plaintext in ``[0, N)`` -- ``decrypt(c, apply_encoding=False)``."""
        c = nat._host_u32(c)
        if c.ndim < 1 or c.shape[-1] != self.kc:
            raise ValueError(f"Paillier ciphertexts must be [..][{self.kc}] limbs")
        shape = c.shape[:-1]
        M = int(np.prod(shape)) if shape else 1
        out = _out(out, shape + (self.kN,))
        self._check(self.lib.scb_paillier_decrypt(self.ctx, nat._ptr(c), nat._ptr(out), M))
        return out

    def dgk_is_zero(self, c: np.ndarray) -> np.ndarray:
        c = nat._host_u32(c)
        if c.ndim < 1 or c.shape[-1] != self.kn:
            raise ValueError(f"DGK ciphertexts must be [..][{self.kn}] limbs")
        shape = c.shape[:-1]
        M = int(np.prod(shape)) if shape else 1
        flags = np.zeros(shape, dtype=np.uint8)
        self._check(self.lib.scb_dgk_is_zero(self.ctx, nat._ptr(c), nat._ptr(flags), M))
        return flags

    # ------------------------------------------------------------------ fused phases (host buffers)
    def initiator_step1(self, x: np.ndarray, y: np.ndarray, r: np.ndarray, rho: Optional[np.ndarray], l: int,
                        out: Optional[np.ndarray] = None) -> np.ndarray:
        x = nat._host_u32(x)
        B = x.shape[0]
        _check_shape("x", x, (B, self.kc))
        y = nat._host_u32(y, (B, self.kc))
        r = nat._host_u32(r, (B, self.kN))
        rho_a = None if rho is None else nat._host_u32(rho, (B, self.kN))
        out = _out(out, (B, self.kc))
        self._check(self.lib.scb_initiator_step1(self.ctx, nat._ptr(x), nat._ptr(y), nat._ptr(r), nat._ptr(rho_a), l, nat._ptr(out), B))
        return out

    def keyholder_step4ab(self, z: np.ndarray, l: int, rho: Optional[np.ndarray], out: Optional[np.ndarray] = None) -> np.ndarray:
        z = nat._host_u32(z)
        B = z.shape[0]
        _check_shape("z", z, (B, self.kN))
        rho_a, rw = None, 0
        if rho is not None:
            rho_a = nat._host_u32(rho)
            if rho_a.ndim != 3 or rho_a.shape[:2] != (B, l + 1):
                raise ValueError("rho must be [B][l+1][words]")
            rw = rho_a.shape[2]
        out = _out(out, (B, l + 1, self.kn))
        self._check(self.lib.scb_keyholder_step4ab(self.ctx, nat._ptr(z), l, nat._ptr(rho_a), rw, nat._ptr(out), B))
        return out

    def initiator_step4(self, db_enc: np.ndarray, r: np.ndarray, delta_a: np.ndarray, r_blind: Optional[np.ndarray],
                        perm: Optional[np.ndarray], rho: Optional[np.ndarray], l: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        db_enc = nat._host_u32(db_enc)
        B = db_enc.shape[0]
        if db_enc.shape != (B, l + 1, self.kn):
            raise ValueError(f"db_enc must be [B][{l + 1}][{self.kn}]")
        r = nat._host_u32(r, (B, self.kN))
        delta_a = np.ascontiguousarray(delta_a, dtype=np.uint8).reshape(B)
        rb, rbw = None, 0
        if r_blind is not None:
            rb = nat._host_u32(r_blind)
            if rb.ndim != 3 or rb.shape[:2] != (B, l + 1):
                raise ValueError("r_blind must be [B][l+1][words]")
            rbw = rb.shape[2]
        pm = None if perm is None else np.ascontiguousarray(perm, dtype=np.uint8).reshape(B, l + 1)
        rho_a, rw = None, 0
        if rho is not None:
            rho_a = nat._host_u32(rho)
            if rho_a.ndim != 3 or rho_a.shape[:2] != (B, l + 1):
                raise ValueError("rho must be [B][l+1][words]")
            rw = rho_a.shape[2]
        out = _out(out, (B, l + 1, self.kn))
        self._check(self.lib.scb_initiator_step4(self.ctx, nat._ptr(db_enc), nat._ptr(r), nat._ptr(delta_a), nat._ptr(rb), rbw, nat._ptr(pm),
                                                 nat._ptr(rho_a), rw, l, nat._ptr(out), B))
        return out

    def keyholder_step4j_5(self, c: np.ndarray, z: np.ndarray, rho: Optional[np.ndarray], l: int,
                           out: Optional[Tuple[np.ndarray, np.ndarray]] = None) -> Tuple[np.ndarray, np.ndarray]:
        c = nat._host_u32(c)
        B = c.shape[0]
        _check_shape("c_is_enc", c, (B, l + 1, self.kn))
        z = nat._host_u32(z, (B, self.kN))
        rho_a = None if rho is None else nat._host_u32(rho, (B, 3, self.kN))
        delta_b = _out(None if out is None else out[0], (B,), np.uint8)
        enc3 = _out(None if out is None else out[1], (B, 3, self.kc))
        self._check(self.lib.scb_keyholder_step4j_5(self.ctx, nat._ptr(c), nat._ptr(z), nat._ptr(rho_a), l, nat._ptr(delta_b), nat._ptr(enc3), B))
        return delta_b, enc3

    def keyholder_step5(self, z: np.ndarray, delta_b: np.ndarray, rho: Optional[np.ndarray], l: int) -> np.ndarray:
        z = nat._host_u32(z)
        B = z.shape[0]
        _check_shape("z", z, (B, self.kN))
        delta_b = np.ascontiguousarray(delta_b, dtype=np.uint8).reshape(B)
        rho_a = None if rho is None else nat._host_u32(rho, (B, 3, self.kN))
        enc3 = np.empty((B, 3, self.kc), dtype=np.uint32)
        self._check(self.lib.scb_keyholder_step5(self.ctx, nat._ptr(z), nat._ptr(delta_b), nat._ptr(rho_a), l, nat._ptr(enc3), B))
        return enc3

    def initiator_step6_7(self, enc3: np.ndarray, r: np.ndarray, delta_a: np.ndarray, l: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        enc3 = nat._host_u32(enc3)
        B = enc3.shape[0]
        _check_shape("enc3", enc3, (B, 3, self.kc))
        r = nat._host_u32(r, (B, self.kN))
        delta_a = np.ascontiguousarray(delta_a, dtype=np.uint8).reshape(B)
        out = _out(out, (B, self.kc))
        self._check(self.lib.scb_initiator_step6_7(self.ctx, nat._ptr(enc3), nat._ptr(r), nat._ptr(delta_a), l, nat._ptr(out), B))
        return out


    # ------------------------------------------------------------------ fused phases (device buffers: anything with data_ptr())
    def initiator_step1_dev(self, x, y, r, rho, l: int, out, B: int) -> None:
        self._check(self.lib.scb_initiator_step1_dev(self.ctx, nat._ptr(x), nat._ptr(y), nat._ptr(r), nat._ptr(rho), l, nat._ptr(out), B))

    def paillier_encrypt_dev(self, plain, rho, out, M: int) -> None:
        self._check(self.lib.scb_paillier_encrypt_dev(self.ctx, nat._ptr(plain), nat._ptr(rho), nat._ptr(out), M))

    def paillier_decrypt_dev(self, c, out, M: int) -> None:
        self._check(self.lib.scb_paillier_decrypt_dev(self.ctx, nat._ptr(c), nat._ptr(out), M))

    def keyholder_step4ab_dev(self, z, l: int, rho, rho_words: int, out, B: int) -> None:
        self._check(self.lib.scb_keyholder_step4ab_dev(self.ctx, nat._ptr(z), l, nat._ptr(rho), rho_words, nat._ptr(out), B))

    def initiator_step4_dev(self, db_enc, r, delta_a, r_blind, rb_words: int, perm, rho, rho_words: int, l: int, out, B: int) -> None:
        self._check(self.lib.scb_initiator_step4_dev(self.ctx, nat._ptr(db_enc), nat._ptr(r), nat._ptr(delta_a), nat._ptr(r_blind), rb_words,
                                                     nat._ptr(perm), nat._ptr(rho), rho_words, l, nat._ptr(out), B))

    def keyholder_step4j_5_dev(self, c, z, rho, l: int, delta_b, enc3, B: int) -> None:
        self._check(self.lib.scb_keyholder_step4j_5_dev(self.ctx, nat._ptr(c), nat._ptr(z), nat._ptr(rho), l, nat._ptr(delta_b), nat._ptr(enc3), B))

    def initiator_step6_7_dev(self, enc3, r, delta_a, l: int, out, B: int) -> None:
        self._check(self.lib.scb_initiator_step6_7_dev(self.ctx, nat._ptr(enc3), nat._ptr(r), nat._ptr(delta_a), l, nat._ptr(out), B))


# -------------------------------------------------------------------------------------------------
# engine registry: the reference's step functions are static methods that only receive scheme objects,
# so the batched mirror looks the engine up from the schemes it is given.
# -------------------------------------------------------------------------------------------------
_ENGINES: dict = {}
_DEFAULT_DEVICE = 0


def set_default_device(device: int) -> None:
    """CUDA device ordinal the scheme-driven API (``Initiator`` / ``KeyHolder`` steps and drivers, ciphertext containers) runs
    on in this process -- one process per GPU sets it to its local rank."""
    global _DEFAULT_DEVICE
    _DEFAULT_DEVICE = int(device)


def get_engine(scheme_paillier=None, scheme_dgk=None, l: Optional[int] = None, device: Optional[int] = None) -> SchemeEngine:
    device = _DEFAULT_DEVICE if device is None else device
    eng = _ENGINES.get(device)
    if eng is None:
        eng = _ENGINES[device] = SchemeEngine(device)
    if scheme_paillier is not None:
        eng.load_paillier(scheme_paillier)
    if scheme_dgk is not None:
        if l is None:
            l = eng.l or max(1, int(scheme_dgk.public_key.u).bit_length() - 3)
        eng.load_dgk(scheme_dgk, l)
    return eng


def reset_engines() -> None:
    for eng in _ENGINES.values():
        eng.close()
    _ENGINES.clear()
