"""Closed-form CPU restatement of the DGK/Veugen secure comparison steps (TEST INFRASTRUCTURE).

Every function works on RAW ciphertext values (Python ints in ``[0, modulus)``) and takes every random
choice as an explicit argument, so the CUDA engine and this oracle can be driven with identical
nonces.  Each function cites the reference lines it restates (``/root/reference/src/tno/mpc/
protocols/secure_comparison/``); the closed forms are those of SURVEY.md section 8a.  Because every
result is a unique group element, the evaluation order chosen here (which differs from the
reference's operator-by-operator order) cannot change a single bit -- ``tests/test_oracle_golden.py``
checks that claim against vectors produced by the unmodified reference code.

Keys are passed as the duck-typed scheme objects of ``oracle/schemes.py`` (or real
``tno.mpc.encryption_schemes`` objects -- only ``public_key`` / ``secret_key`` attributes are read).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

from .bigint import get_backend


def to_bits(integer: int, bit_length: int) -> List[int]:
    """LSB-first bits (reference ``utils.py:6-21``, including its assertion)."""
    assert integer < (1 << bit_length)
    return [(integer >> i) & 1 for i in range(bit_length)]


def from_bits(bits: Sequence[int]) -> int:
    """Reference ``utils.py:24-38``."""
    return sum(1 << i for i, bit in enumerate(bits) if bit == 1)


@dataclass
class Nonces:
    """All randomness of ONE comparison, in the reference's draw order."""

    r: int                      # initiator.py:250   randbelow(N)
    rho_z: int                  # initiator.py:109   Paillier randomizer base for [[z]]
    delta_a: int                # initiator.py:420   randbelow(2)
    r_blind: List[int]          # initiator.py:512   randbelow(u-1)+1, list order 0..l
    perm: List[int]             # initiator.py:212-226 out[k] = masked[perm[k]]
    rho_c: List[int]            # initiator.py:153-154 DGK randomizer exponents, shuffled order
    rho_d: int                  # keyholder.py:106
    rho_beta: List[int]         # keyholder.py:107-108
    rho_zeta1: int              # keyholder.py:126
    rho_zeta2: int              # keyholder.py:127
    rho_delta_b: int            # keyholder.py:128


@dataclass
class Transcript:
    """Every intermediate of one comparison (raw values) -- what the parity tests compare."""

    z_enc_step1: int = 0
    z_enc: int = 0
    z: int = 0
    beta: int = 0
    alpha: List[int] = field(default_factory=list)
    d_enc_4a: int = 0
    beta_enc_4b: List[int] = field(default_factory=list)
    d_enc_sent: int = 0
    beta_enc_sent: List[int] = field(default_factory=list)
    d_enc_4c: int = 0
    xor_enc: List[int] = field(default_factory=list)
    w_enc_4e: List[int] = field(default_factory=list)
    alpha_tilde: List[int] = field(default_factory=list)
    w_enc_4f: List[int] = field(default_factory=list)
    c_enc_4h: List[int] = field(default_factory=list)
    c_enc_4i: List[int] = field(default_factory=list)
    c_enc_sent: List[int] = field(default_factory=list)
    delta_b: int = 0
    zeta1_enc: int = 0
    zeta2_enc: int = 0
    delta_b_enc: int = 0
    zeta1_sent: int = 0
    zeta2_sent: int = 0
    delta_b_sent: int = 0
    beta_lt_alpha_enc: int = 0
    result_enc: int = 0
    result_bit: Optional[int] = None


class Oracle:
    """Protocol steps on raw integers.  ``backend`` = ``"cpython"`` | ``"gmp"`` | ``"auto"``."""

    def __init__(self, paillier, dgk, backend="auto", dgk_negative_exponent_mode: str = "inverse"):
        self.be = get_backend(backend) if isinstance(backend, str) else backend
        self.paillier = paillier
        self.dgk = dgk
        self.N = paillier.public_key.n
        self.N2 = self.N * self.N
        self.n = dgk.public_key.n
        self.g = dgk.public_key.g
        self.h = dgk.public_key.h
        self.u = dgk.public_key.u
        self.neg_mode = dgk_negative_exponent_mode
        self._g_inv: Optional[int] = None

    # ------------------------------------------------------------------ scheme primitives (U1-U7)
    def paillier_encrypt_raw(self, m: int) -> int:
        """U1: ``unsafe_encrypt(m, apply_encoding=False)`` = (1 + m N) mod N^2."""
        return (1 + m * self.N) % self.N2

    def paillier_randomize(self, c: int, rho: int) -> int:
        """U2: c * rho^N mod N^2."""
        return self.be.mulmod(c, self.be.powmod(rho, self.N, self.N2), self.N2)

    def paillier_decrypt_raw(self, c: int) -> int:
        """U3: L(c^lambda mod N^2) * mu mod N."""
        sk = self.paillier.secret_key
        x = self.be.powmod(c, sk.lambda_, self.N2)
        return ((x - 1) // self.N) * sk.mu % self.N

    def paillier_decrypt_signed(self, c: int) -> int:
        raw = self.paillier_decrypt_raw(c)
        return raw if raw <= self.N // 2 else raw - self.N

    def dgk_encrypt_raw(self, m: int) -> int:
        """U4: g^m mod n, negative m through the inverse (or reduced mod u when so configured)."""
        if m < 0:
            if self.neg_mode == "mod_u":
                return self.be.powmod(self.g, m % self.u, self.n)
            if self._g_inv is None:
                self._g_inv = self.be.invert(self.g, self.n)
            return self.be.powmod(self._g_inv, -m, self.n)
        return self.be.powmod(self.g, m, self.n)

    def dgk_randomize(self, c: int, rho: int) -> int:
        """U6: c * h^rho mod n."""
        return self.be.mulmod(c, self.be.powmod(self.h, rho, self.n), self.n)

    def dgk_is_zero(self, c: int) -> bool:
        """U7: c^{v_p} mod p == 1."""
        sk = self.dgk.secret_key
        return self.be.powmod(c % sk.p, sk.v_p, sk.p) == 1

    # ------------------------------------------------------------------ Initiator (Alice)
    def step_1(self, x_enc: int, y_enc: int, l: int, r: int) -> int:
        """``Initiator.step_1`` (initiator.py:228-258): [[z]] = y * x^-1 * (1 + (2^l + r) N) mod N^2."""
        assert (1 << (l + 2)) < self.N // 2
        t = self.be.mulmod(y_enc, self.be.invert(x_enc, self.N2), self.N2)
        return self.be.mulmod(t, self.paillier_encrypt_raw((1 << l) + r), self.N2)

    def step_3(self, r: int, l: int) -> List[int]:
        """``Initiator.step_3`` (initiator.py:260-270)."""
        return to_bits(r % (1 << l), l)

    def step_4c(self, d_enc: int, r: int) -> int:
        """``Initiator.step_4c`` (initiator.py:272-291): [d] <- [0] = 1 when r < (N-1)//2."""
        assert 0 <= r < self.N
        return 1 if r < (self.N - 1) // 2 else d_enc

    def step_4d(self, alpha: Sequence[int], beta_enc: Sequence[int]) -> List[int]:
        """``Initiator.step_4d`` (initiator.py:293-328): alpha_i = 1 -> g * [beta_i]^-1 mod n."""
        return [
            b if a == 0 else self.be.mulmod(self.be.invert(b, self.n), self.g, self.n)
            for a, b in zip(alpha, beta_enc)
        ]

    def step_4e(self, r: int, alpha: Sequence[int], xor_enc: Sequence[int], d_enc: int) -> Tuple[List[int], List[int]]:
        """``Initiator.step_4e`` (initiator.py:330-384)."""
        l = len(xor_enc)
        alpha_tilde = to_bits((r - self.N) % (1 << l), l)
        d_inv = None
        out = []
        for a, at, x in zip(alpha, alpha_tilde, xor_enc):
            if a == at:
                out.append(x)
            else:
                if d_inv is None:
                    d_inv = self.be.invert(d_enc, self.n)
                out.append(self.be.mulmod(x, d_inv, self.n))
        return out, alpha_tilde

    def step_4f(self, w_enc: Sequence[int]) -> List[int]:
        """``Initiator.step_4f`` (initiator.py:386-410): [w_i]^(2^i) mod n."""
        return [self.be.powmod(w, 1 << i, self.n) for i, w in enumerate(w_enc)]

    def step_4h(self, s: int, alpha: Sequence[int], alpha_tilde: Sequence[int], d_enc: int,
                beta_enc: Sequence[int], w_enc: Sequence[int], delta_a: int) -> List[int]:
        """``Initiator.step_4h`` (initiator.py:423-485).  out[i+1] = g^s g^{alpha_i} [d]^{at_i-a_i}
        [beta_i]^-1 (prod_{j>i} [w_j])^3 ; out[0] = g^{delta_a} prod_j [w_j]  (all mod n)."""
        l = len(beta_enc)
        mm, n = self.be.mulmod, self.n
        g_s = self.dgk_encrypt_raw(s)
        d_pow = {0: 1, 1: d_enc % n, -1: None}
        out = [0] * l
        w_sum = 1  # the reference's ``w_is_enc_sum`` (int 0 -> encrypts to 1)
        for i in range(l - 1, -1, -1):
            e = alpha_tilde[i] - alpha[i]
            if e == -1 and d_pow[-1] is None:
                d_pow[-1] = self.be.invert(d_enc, n)
            v = mm(g_s, self.g if alpha[i] else 1, n)
            v = mm(v, d_pow[e], n)
            v = mm(v, self.be.invert(beta_enc[i], n), n)
            v = mm(v, self.be.powmod(w_sum, 3, n), n)
            out[i] = v
            w_sum = mm(w_sum, w_enc[i], n)
        c_minus1 = mm(w_sum, self.g if delta_a else 1, n)
        return [c_minus1] + out

    def step_4i(self, c_enc: Sequence[int], r_blind: Sequence[int], perm: Optional[Sequence[int]]) -> List[int]:
        """``Initiator.step_4i`` (initiator.py:487-516): c_i^{r_i}, then out[k] = masked[perm[k]]."""
        masked = [self.be.powmod(c, r, self.n) for c, r in zip(c_enc, r_blind)]
        return masked if perm is None else [masked[p] for p in perm]

    def step_6(self, delta_a: int, delta_b_enc: int) -> int:
        """``Initiator.step_6`` (initiator.py:518-531)."""
        if delta_a == 1:
            return delta_b_enc
        return self.be.mulmod(self.be.invert(delta_b_enc, self.N2), self.paillier_encrypt_raw(1), self.N2)

    def step_7(self, zeta1_enc: int, zeta2_enc: int, r: int, l: int, beta_lt_alpha_enc: int) -> int:
        """``Initiator.step_7`` (initiator.py:533-564)."""
        zeta = zeta1_enc if r < (self.N - 1) // 2 else zeta2_enc
        t = self.be.mulmod(self.paillier_encrypt_raw(r >> l), beta_lt_alpha_enc, self.N2)
        return self.be.mulmod(zeta, self.be.invert(t, self.N2), self.N2)

    # ------------------------------------------------------------------ KeyHolder (Bob)
    def step_2(self, z_enc: int, l: int) -> Tuple[int, int]:
        """``KeyHolder.step_2`` (keyholder.py:181-196)."""
        z = self.paillier_decrypt_raw(z_enc)
        return z, z % (1 << l)

    def step_4a(self, z: int, l: int) -> int:
        """``KeyHolder.step_4a`` (keyholder.py:198-216): [d] = g^d, d = (z < (N-1)//2)."""
        assert self.u > (1 << (l + 2))
        return self.dgk_encrypt_raw(int(z < (self.N - 1) // 2))

    def step_4b(self, beta: int, l: int) -> List[int]:
        """``KeyHolder.step_4b`` (keyholder.py:218-233)."""
        return [self.dgk_encrypt_raw(bit) for bit in to_bits(beta, l)]

    def step_4j(self, c_enc: Sequence[int]) -> int:
        """``KeyHolder.step_4j`` (keyholder.py:235-253)."""
        return int(any(self.dgk_is_zero(c) for c in c_enc))

    def step_5(self, z: int, l: int, delta_b: int) -> Tuple[int, int, int]:
        """``KeyHolder.step_5`` (keyholder.py:255-287)."""
        zeta1 = self.paillier_encrypt_raw(z >> l)
        zeta2 = self.paillier_encrypt_raw((z + self.N) >> l) if z < (self.N - 1) // 2 else zeta1
        return zeta1, zeta2, self.paillier_encrypt_raw(delta_b)

    # ------------------------------------------------------------------ whole comparison
    def compare(self, x_enc: int, y_enc: int, l: int, nc: Nonces, interactive: bool = True,
                decrypt_result: bool = True) -> Transcript:
        """One comparison in the order of ``perform_secure_comparison`` (initiator.py:69-175,
        keyholder.py:70-133).  ``interactive=False`` skips every ``randomize()`` like the
        step-by-step flow of ``README.md:117-141`` / ``test_secure_comparison.py:71-124``."""
        t = Transcript()
        t.z_enc_step1 = self.step_1(x_enc, y_enc, l, nc.r)
        t.z_enc = self.paillier_randomize(t.z_enc_step1, nc.rho_z) if interactive else t.z_enc_step1
        t.alpha = self.step_3(nc.r, l)
        t.z, t.beta = self.step_2(t.z_enc, l)
        t.d_enc_4a = self.step_4a(t.z, l)
        t.beta_enc_4b = self.step_4b(t.beta, l)
        if interactive:
            t.d_enc_sent = self.dgk_randomize(t.d_enc_4a, nc.rho_d)
            t.beta_enc_sent = [self.dgk_randomize(b, rho) for b, rho in zip(t.beta_enc_4b, nc.rho_beta)]
        else:
            t.d_enc_sent, t.beta_enc_sent = t.d_enc_4a, list(t.beta_enc_4b)
        t.d_enc_4c = self.step_4c(t.d_enc_sent, nc.r)
        t.xor_enc = self.step_4d(t.alpha, t.beta_enc_sent)
        t.w_enc_4e, t.alpha_tilde = self.step_4e(nc.r, t.alpha, t.xor_enc, t.d_enc_4c)
        t.w_enc_4f = self.step_4f(t.w_enc_4e)
        s = 1 - 2 * nc.delta_a
        t.c_enc_4h = self.step_4h(s, t.alpha, t.alpha_tilde, t.d_enc_4c, t.beta_enc_sent, t.w_enc_4f, nc.delta_a)
        t.c_enc_4i = self.step_4i(t.c_enc_4h, nc.r_blind, nc.perm)
        if interactive:
            t.c_enc_sent = [self.dgk_randomize(c, rho) for c, rho in zip(t.c_enc_4i, nc.rho_c)]
        else:
            t.c_enc_sent = list(t.c_enc_4i)
        t.delta_b = self.step_4j(t.c_enc_sent)
        t.zeta1_enc, t.zeta2_enc, t.delta_b_enc = self.step_5(t.z, l, t.delta_b)
        if interactive:
            t.zeta1_sent = self.paillier_randomize(t.zeta1_enc, nc.rho_zeta1)
            t.zeta2_sent = self.paillier_randomize(t.zeta2_enc, nc.rho_zeta2)
            t.delta_b_sent = self.paillier_randomize(t.delta_b_enc, nc.rho_delta_b)
        else:
            t.zeta1_sent, t.zeta2_sent, t.delta_b_sent = t.zeta1_enc, t.zeta2_enc, t.delta_b_enc
        t.beta_lt_alpha_enc = self.step_6(nc.delta_a, t.delta_b_sent)
        t.result_enc = self.step_7(t.zeta1_sent, t.zeta2_sent, nc.r, l, t.beta_lt_alpha_enc)
        if decrypt_result:
            t.result_bit = self.paillier_decrypt_signed(t.result_enc)
        return t
