"""Scheme stand-ins of the oracle: Paillier and DGK with the attribute surface the reference uses.

TEST INFRASTRUCTURE.  The reference never touches a big integer itself: every operation goes through
``tno.mpc.encryption_schemes.paillier.Paillier`` / ``.dgk.DGK`` objects and operator overloading on
their ciphertexts (call sites: reference ``initiator.py:96,101,109,153-154,249-256,287-290,320,371,
406,460-484,503-513,531,559-564``; ``keyholder.py:106-108,126-128,156-166,195,212-216,230-233,249,
274-287``).  Those packages are not vendored and not installable here, so this module restates the
semantics the protocol relies on:

U1  Paillier ``unsafe_encrypt(m, apply_encoding=False)`` = ``(1 + m*N) mod N^2``  (g = N+1).
U2  Paillier ``randomize()``: ``c <- c * rho^N mod N^2``, rho in [1, N).
U3  Paillier ``decrypt(c, apply_encoding=False)`` = ``((c^lambda mod N^2 - 1)//N * mu) mod N``.
U4  DGK ``unsafe_encrypt(m, apply_encoding=False)`` = ``pow_mod(g, m, n)``; m < 0 goes through the
    modular inverse (switchable: ``negative_exponent_mode``).
U5  ``neg`` = modular inverse, ``mul(c, k)`` = ``c^k`` (k < 0: inverse first), ``add`` = modular
    product; ``ct + int`` encrypts the int first; ``k - ct`` = ``add(neg(ct), k)``.
U6  DGK ``randomize()``: ``c <- c * h^rho mod n``, rho of ``randomness_bits`` (~2.5 t) bits.
U7  DGK ``is_zero(c)`` = ``pow(c, v_p, p) == 1``.
U8  ``fresh`` flag: set by ``randomize()``, cleared by ``get_value()``; arithmetic on a fresh
    ciphertext warns.

Randomness is injectable (``scheme.randomness_source``) so that the CUDA engine and this oracle can
be driven with the same nonces; the default source is ``secrets`` like the reference.
"""
from __future__ import annotations

import random
import secrets
import warnings
from dataclasses import dataclass, field
from typing import Any, Callable, Iterable, Iterator, Optional, Union

from .bigint import PythonBackend, get_backend

Number = Union[int, float]


# ---------------------------------------------------------------------------------------------
# primes (host-side, one-off; keygen is out of the hot path -- SURVEY.md section 3.3)
# ---------------------------------------------------------------------------------------------
_SMALL_PRIMES = [p for p in range(3, 2000, 2) if all(p % d for d in range(3, int(p ** 0.5) + 1, 2))]


_GMP_PRIMALITY = None


def is_probable_prime(n: int, rounds: int = 24) -> bool:
    """Miller-Rabin with fixed + seeded bases (deterministic for a given n); large candidates go to
    GMP's ``mpz_probab_prime_p`` when libgmp is loadable (same verdict, much faster)."""
    global _GMP_PRIMALITY
    if n.bit_length() > 256:
        if _GMP_PRIMALITY is None:
            try:
                from .bigint import GmpBackend
                _GMP_PRIMALITY = GmpBackend()
            except OSError:
                _GMP_PRIMALITY = False
        if _GMP_PRIMALITY:
            return _GMP_PRIMALITY.is_probable_prime(n, rounds)
    if n < 2:
        return False
    if n in (2, 3):
        return True
    if n % 2 == 0:
        return False
    for p in _SMALL_PRIMES:
        if n == p:
            return True
        if n % p == 0:
            return False
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    rng = random.Random(n & 0xFFFFFFFFFFFF)
    for i in range(rounds):
        a = 2 if i == 0 else rng.randrange(3, n - 1)
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def next_prime(n: int) -> int:
    """Smallest prime strictly greater than ``n`` (``tno.mpc.encryption_schemes.utils.next_prime``;
    used by the reference for ``u = next_prime(1 << (l + 2))``, ``keyholder.py:164``)."""
    c = n + 1
    if c <= 2:
        return 2
    c |= 1
    while not is_probable_prime(c):
        c += 2
    return c


def randprime(bits: int, rng: random.Random) -> int:
    """Random prime of exactly ``bits`` bits with the two top bits set (so products have full length)."""
    while True:
        c = rng.getrandbits(bits) | (3 << (bits - 2)) | 1
        if is_probable_prime(c):
            return c


# ---------------------------------------------------------------------------------------------
# randomness sources
# ---------------------------------------------------------------------------------------------
class _SecretsRng:
    def randbelow(self, n: int) -> int:
        return secrets.randbelow(n)

    def randbits(self, k: int) -> int:
        return secrets.randbits(k)


class InjectedRandomness:
    """Hands out pre-drawn raw nonces (rho) in order; raises when exhausted, which mirrors the
    reference's "randomness pool exhausted" warning-turned-error in its strict test fixtures
    (``test/conftest.py:26-36``)."""

    def __init__(self, values: Iterable[int]):
        self._it: Iterator[int] = iter(list(values))
        self.drawn = 0

    def __call__(self) -> int:
        try:
            value = next(self._it)
        except StopIteration as exc:
            raise RuntimeError("injected randomness exhausted") from exc
        self.drawn += 1
        return value


# ---------------------------------------------------------------------------------------------
# ciphertext algebra (templates semantics: dunder -> scheme.add/neg/mul)
# ---------------------------------------------------------------------------------------------
class _Ciphertext:
    __slots__ = ("_raw_value", "scheme", "_fresh")

    def __init__(self, raw_value: int, scheme: Any, *, fresh: bool = False):
        self._raw_value = int(raw_value)
        self.scheme = scheme
        self._fresh = fresh

    # -- value access / freshness (U8) --
    @property
    def fresh(self) -> bool:
        return self._fresh

    def peek_value(self) -> int:
        return self._raw_value

    def get_value(self) -> int:
        self._fresh = False
        return self._raw_value

    @property
    def value(self) -> int:  # legacy alias kept by the upstream templates
        return self.get_value()

    def randomize(self) -> "_Ciphertext":
        self._raw_value = self.scheme._apply_randomness(self._raw_value)
        self._fresh = True
        return self

    def _use(self) -> int:
        if self._fresh:
            warnings.warn(
                "Arithmetic on a fresh ciphertext: its randomness is reused and the ciphertext is no longer fresh.",
                UserWarning,
                stacklevel=3,
            )
            self._fresh = False
        return self._raw_value

    # -- operators --
    def _coerce(self, other: Any) -> "_Ciphertext":
        if isinstance(other, _Ciphertext):
            if other.scheme != self.scheme:
                raise AttributeError("The schemes of the ciphertexts differ.")
            return other
        return self.scheme.unsafe_encrypt(other)

    def __neg__(self):
        return self.scheme.neg(self)

    def __add__(self, other):
        return self.scheme.add(self, self._coerce(other))

    __radd__ = __add__

    def __sub__(self, other):
        if isinstance(other, _Ciphertext):
            return self.scheme.add(self, self.scheme.neg(self._coerce(other)))
        return self.scheme.add(self, self.scheme.unsafe_encrypt(-other))

    def __rsub__(self, other):
        return self.scheme.add(self.scheme.neg(self), self._coerce(other))

    def __mul__(self, scalar):
        return self.scheme.mul(self, scalar)

    __rmul__ = __mul__

    def __eq__(self, other):
        return (
            isinstance(other, _Ciphertext)
            and self._raw_value == other._raw_value
            and self.scheme == other.scheme
        )

    def __hash__(self):
        return hash((self._raw_value, id(self.scheme)))

    def __repr__(self):
        return f"{type(self).__name__}(0x{self._raw_value:x})"


class PaillierCiphertext(_Ciphertext):
    """``PaillierCiphertext(raw_value, scheme)`` -- constructed directly in the reference's tests
    (``test/unit/test_secure_comparison.py:176``)."""


class DGKCiphertext(_Ciphertext):
    pass


class _Scheme:
    ciphertext_cls = _Ciphertext
    modulus: int

    def __init__(self, backend=None, randomness_source: Optional[Callable[[], int]] = None):
        self.backend = backend if backend is not None else get_backend("auto")
        self.randomness_source = randomness_source
        self._rng = _SecretsRng()

    # randomness pool surface of the templates package -- no threads/processes here
    def boot_randomness_generation(self, amount: int = 0) -> None:  # noqa: D401
        return None

    def boot_generation(self, amount: int = 0) -> None:
        return None

    def shut_down(self) -> None:
        return None

    def get_randomness(self) -> int:
        raise NotImplementedError

    def _apply_randomness(self, raw: int) -> int:
        return self.backend.mulmod(raw, self.get_randomness(), self.modulus)

    def encrypt(self, plaintext: Number, apply_encoding: bool = True):
        return self.unsafe_encrypt(plaintext, apply_encoding).randomize()

    # homomorphic algebra (U5)
    def neg(self, ct):
        return self.ciphertext_cls(self.backend.invert(ct._use(), self.modulus), self)

    def add(self, a, b):
        if not isinstance(b, _Ciphertext):
            b = self.unsafe_encrypt(b)
        return self.ciphertext_cls(self.backend.mulmod(a._use(), b._use(), self.modulus), self)

    def sub(self, a, b):
        return self.add(a, self.neg(b))

    def mul(self, ct, scalar: int):
        if not isinstance(scalar, int):
            raise TypeError("scalar must be an int")
        base = ct._use()
        if scalar < 0:
            base, scalar = self.backend.invert(base, self.modulus), -scalar
        return self.ciphertext_cls(self.backend.powmod(base, scalar, self.modulus), self)


# ---------------------------------------------------------------------------------------------
# Paillier
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class PaillierPublicKey:
    n: int
    g: int
    n_squared: int = field(init=False)

    def __post_init__(self):
        object.__setattr__(self, "n_squared", self.n * self.n)


@dataclass(frozen=True)
class PaillierSecretKey:
    lambda_: int
    mu: int
    n: int
    p: Optional[int] = None
    q: Optional[int] = None


class Paillier(_Scheme):
    ciphertext_cls = PaillierCiphertext

    def __init__(self, public_key: PaillierPublicKey, secret_key: Optional[PaillierSecretKey] = None, **kw):
        super().__init__(**kw)
        self.public_key = public_key
        self.secret_key = secret_key
        self.modulus = public_key.n_squared
        self.precision = 0
        self.max_value = public_key.n // 2
        self.min_value = -(public_key.n // 2)

    @classmethod
    def from_security_parameter(cls, key_length: int = 2048, rng: Optional[random.Random] = None, **kw) -> "Paillier":
        """Key generation as the reference requests it (``keyholder.py:156-158``): ``n = p*q`` of exactly
        ``key_length`` bits, ``g = n + 1``, ``lambda = (p-1)(q-1)``, ``mu = lambda^-1 mod n``."""
        rng = rng or random.Random(secrets.randbits(128))
        while True:
            p = randprime(key_length // 2, rng)
            q = randprime(key_length - key_length // 2, rng)
            if p != q and (p * q).bit_length() == key_length:
                break
        n = p * q
        lam = (p - 1) * (q - 1)
        return cls(PaillierPublicKey(n, n + 1), PaillierSecretKey(lam, pow(lam, -1, n), n, p, q), **kw)

    def public_only(self) -> "Paillier":
        return Paillier(self.public_key, None, backend=self.backend)

    def __eq__(self, other):
        return isinstance(other, Paillier) and self.public_key == other.public_key

    def __hash__(self):
        return hash(self.public_key.n)

    # encoding (fixed point with precision 0: ints pass through, negatives wrap mod n)
    def encode(self, value: Number) -> int:
        iv = int(value)
        if iv != value:
            raise ValueError("precision 0 scheme: only integral plaintexts")
        if not self.min_value <= iv <= self.max_value:
            warnings.warn("Encoding of plaintext outside the supported plaintext range.", UserWarning)
        return iv % self.public_key.n

    def decode(self, raw: int) -> int:
        return raw if raw <= self.max_value else raw - self.public_key.n

    def unsafe_encrypt(self, plaintext: Number, apply_encoding: bool = True) -> PaillierCiphertext:
        m = self.encode(plaintext) if apply_encoding else int(plaintext)
        assert self.public_key.g == self.public_key.n + 1
        return PaillierCiphertext((1 + m * self.public_key.n) % self.modulus, self)

    def decrypt(self, ct: PaillierCiphertext, apply_encoding: bool = True) -> int:
        if self.secret_key is None:
            raise ValueError("This scheme only has a public key.")
        n = self.public_key.n
        x = self.backend.powmod(ct.peek_value(), self.secret_key.lambda_, self.modulus)
        raw = ((x - 1) // n) * self.secret_key.mu % n
        return self.decode(raw) if apply_encoding else raw

    def random_rho(self) -> int:
        if self.randomness_source is not None:
            return self.randomness_source()
        return self._rng.randbelow(self.public_key.n - 1) + 1

    def get_randomness(self) -> int:
        return self.backend.powmod(self.random_rho(), self.public_key.n, self.modulus)


# ---------------------------------------------------------------------------------------------
# DGK
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class DGKPublicKey:
    g: int
    h: int
    u: int
    n: int
    t: int


@dataclass(frozen=True)
class DGKSecretKey:
    v_p: int
    v_q: int
    p: int
    q: int


class DGK(_Scheme):
    ciphertext_cls = DGKCiphertext
    #: U4 switch: how a negative raw plaintext is exponentiated. "inverse" = pow_mod semantics.
    negative_exponent_mode = "inverse"

    def __init__(self, public_key: DGKPublicKey, secret_key: Optional[DGKSecretKey] = None,
                 full_decryption: bool = False, randomness_bits: Optional[int] = None, **kw):
        super().__init__(**kw)
        self.public_key = public_key
        self.secret_key = secret_key
        self.modulus = public_key.n
        self.full_decryption = full_decryption
        self.randomness_bits = randomness_bits if randomness_bits is not None else (5 * public_key.t) // 2
        self.max_value = public_key.u // 2
        self.min_value = -(public_key.u // 2)
        self._table: Optional[dict[int, int]] = None

    @classmethod
    def from_security_parameter(cls, v_bits: int = 160, n_bits: int = 2048, u: int = 2 ** 18 + 3,
                                full_decryption: bool = False, rng: Optional[random.Random] = None, **kw) -> "DGK":
        """Standard DGK key generation as requested at reference ``keyholder.py:161-166``: t-bit primes
        v_p, v_q; primes p, q with ``u*v_p | p-1`` and ``u*v_q | q-1``; ``n = p*q`` of exactly ``n_bits``;
        g of order ``u*v_p*v_q``, h of order ``v_p*v_q`` in Z_n^*."""
        rng = rng or random.Random(secrets.randbits(128))
        if not is_probable_prime(u):
            raise ValueError("u must be prime")
        half = n_bits // 2

        def prime_with_factor(bits: int, factor: int) -> int:
            # p = 2 * factor * r + 1 with r random of the right size
            rbits = bits - (2 * factor).bit_length()
            if rbits < 2:
                raise ValueError("n_bits too small for u and v_bits")
            while True:
                r = rng.getrandbits(rbits + 1) | (1 << rbits)
                cand = 2 * factor * r + 1
                if cand.bit_length() == bits and (cand >> (bits - 2)) == 3 and is_probable_prime(cand):
                    return cand

        while True:
            v_p = randprime(v_bits, rng)
            v_q = randprime(v_bits, rng)
            if v_p == v_q or v_p == u or v_q == u:
                continue
            p = prime_with_factor(half, u * v_p)
            q = prime_with_factor(n_bits - half, u * v_q)
            if p != q and (p * q).bit_length() == n_bits:
                break
        n = p * q

        def element_of_order(prime: int, order_factors: list[int]) -> int:
            order = 1
            for f in order_factors:
                order *= f
            while True:
                x = rng.randrange(2, prime - 1)
                e = pow(x, (prime - 1) // order, prime)
                if all(pow(e, order // f, prime) != 1 for f in order_factors):
                    return e

        def crt(a_p: int, a_q: int) -> int:
            return (a_p + p * ((a_q - a_p) * pow(p, -1, q) % q)) % n

        g = crt(element_of_order(p, [u, v_p]), element_of_order(q, [u, v_q]))
        h = crt(element_of_order(p, [v_p]), element_of_order(q, [v_q]))
        return cls(DGKPublicKey(g, h, u, n, v_bits), DGKSecretKey(v_p, v_q, p, q), full_decryption, **kw)

    def public_only(self) -> "DGK":
        return DGK(self.public_key, None, False, self.randomness_bits, backend=self.backend)

    def __eq__(self, other):
        return isinstance(other, DGK) and self.public_key == other.public_key

    def __hash__(self):
        return hash(self.public_key.n)

    def encode(self, value: Number) -> int:
        iv = int(value)
        if iv != value:
            raise ValueError("DGK plaintexts are integers")
        return iv % self.public_key.u

    def decode(self, raw: int) -> int:
        return raw if raw <= self.max_value else raw - self.public_key.u

    def unsafe_encrypt(self, plaintext: Number, apply_encoding: bool = True) -> DGKCiphertext:
        m = self.encode(plaintext) if apply_encoding else int(plaintext)
        if m < 0 and self.negative_exponent_mode == "mod_u":
            m %= self.public_key.u
        return DGKCiphertext(self.backend.powmod(self.public_key.g, m, self.modulus), self)

    def is_zero(self, ct: DGKCiphertext) -> bool:
        if self.secret_key is None:
            raise ValueError("This scheme only has a public key.")
        sk = self.secret_key
        return self.backend.powmod(ct.peek_value() % sk.p, sk.v_p, sk.p) == 1

    def decrypt(self, ct: DGKCiphertext, apply_encoding: bool = True) -> int:
        """Full decryption through a lookup table of ``(g^{v_p})^m mod p`` (only when the scheme was
        made with ``full_decryption=True``, as the reference's ``dgk_full`` test scheme is)."""
        if self.secret_key is None:
            raise ValueError("This scheme only has a public key.")
        if not self.full_decryption:
            raise ValueError("This DGK scheme only supports is_zero (full_decryption=False).")
        sk = self.secret_key
        if self._table is None:
            base = pow(self.public_key.g, sk.v_p, sk.p)
            table, acc = {}, 1
            for m in range(self.public_key.u):
                table[acc] = m
                acc = acc * base % sk.p
            self._table = table
        raw = self._table[pow(ct.peek_value() % sk.p, sk.v_p, sk.p)]
        return self.decode(raw) if apply_encoding else raw

    def random_rho(self) -> int:
        if self.randomness_source is not None:
            return self.randomness_source()
        return self._rng.randbits(self.randomness_bits)

    def get_randomness(self) -> int:
        return self.backend.powmod(self.public_key.h, self.random_rho(), self.modulus)
