"""Batched ``Initiator`` (Alice) -- same step names and argument order as the reference
(``tno/mpc/protocols/secure_comparison/initiator.py:15-564``), every argument being a batch.

Batch containers: ``[B][k]`` uint32 limb arrays (preferred), or sequences of scheme ciphertext objects /
ints (converted once at the surface); every ciphertext result is a ``BatchedPaillierCiphertext`` /
``BatchedDGKCiphertext`` (``ciphertexts.py``: a limb array that also carries the scheme, the ``fresh`` flag and the
reference's ``+ - *`` operator algebra).  Bit vectors are ``[B][l]`` uint8 arrays, LSB first.  Every random
choice of the reference (``secrets.randbelow`` / ``choice`` at initiator.py:223,250,420,512) can be injected
through keyword arguments so that the CUDA path can be compared bit for bit with the CPU oracle; when
left out it is drawn by the device CSPRNG (ChaCha20 keyed from ``secrets``, ``csrc/scb200_rng.cu``).

The individual ``step_*`` methods mirror the reference one to one and are built from the engine's
batched primitives; ``steps_4c_to_4i`` and ``steps_6_7`` are the fused single-launch-sequence phases
the throughput path uses (same results, see tests/test_gpu_protocol.py).
"""
from __future__ import annotations

import secrets
from typing import Any, Optional, Tuple

import numpy as np

from . import _native as nat
from .ciphertexts import BatchedDGKCiphertext, BatchedPaillierCiphertext, dgk_batch, paillier_batch
from .communicator import Communicator
from .limbs import int_to_limbs, ints_to_limbs
from .schemes import SchemeEngine, as_limbs, get_engine
from .sessions import InitiatorSession
from .utils import less_than_const, low_bits, random_permutations, shift_right, to_bits_batch


def _bits(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a.reshape(1, -1) if a.ndim == 1 else a


def _scheme_of(container, fallback=None):
    """The reference's ciphertext objects carry their scheme; limb arrays do not (then the scheme given
    explicitly, or the one already loaded in the engine, is used)."""
    if fallback is not None:
        return fallback
    if getattr(container, "scheme", None) is not None:
        return container.scheme
    item = container
    while isinstance(item, (list, tuple)) and item:
        item = item[0]
    return getattr(item, "scheme", None)


class Initiator:
    """Player Alice in the secure comparison protocol, initiates."""

    def __init__(self, l_maximum_bit_length: int, communicator: Optional[Communicator] = None, other_party: str = "",
                 scheme_paillier: Any = None, scheme_dgk: Any = None, session_id: int = 0, *, wire_format: bool = False) -> None:
        # wire_format (new, keyword-only): exchange each protocol message as one packed blob (wire.py) instead of a numpy array
        self.wire_format = wire_format
        self.l_maximum_bit_length = l_maximum_bit_length
        self.communicator = communicator
        self.other_party = other_party
        self._scheme_paillier = scheme_paillier
        self._scheme_dgk = scheme_dgk
        self.session_id = session_id

    @property
    def scheme_paillier(self):
        if self._scheme_paillier is None:
            raise ValueError("No Paillier scheme has been initialized or received.")
        return self._scheme_paillier

    @property
    def scheme_dgk(self):
        if self._scheme_dgk is None:
            raise ValueError("No DGK scheme has been initialized or received.")
        return self._scheme_dgk

    # ------------------------------------------------------------------ driver (initiator.py:69-203)
    async def perform_secure_comparison(self, x, y, nonces: Optional[dict] = None) -> BatchedPaillierCiphertext:
        """All of Alice's steps for a BATCH of comparisons, with the reference's message ids (``initiator.py:69-175``).

        ``x`` / ``y``: batches of Paillier ciphertexts (containers, ``[B][2kN]`` limb arrays, sequences of scheme ciphertext
        objects) or -- like the reference, ``initiator.py:93-102`` -- plaintext numbers (an int, a sequence or an integer array;
        negatives allowed), which are encrypted first.  Returns ``[[x <= y]]`` as a ``BatchedPaillierCiphertext`` (``[B][2kN]``).

        The session state (r, delta_A) stays on the device between the phases and every nonce not injected through ``nonces``
        (keys ``r, rho_z, delta_a, r_blind, perm, rho_c``; parity runs) is drawn by the device CSPRNG; only the four protocol
        messages cross PCIe.  Sessions of one ``Initiator`` may overlap (``asyncio.gather``): nothing of a session is kept on
        ``self`` except the monotonically increasing ``session_id``."""
        if self.communicator is None:
            raise ValueError("Communicator not properly initialized.")
        self.session_id += 1
        session_id = self.session_id
        await self.receive_encryption_schemes(session_id)
        l = self.l_maximum_bit_length
        eng = get_engine(self.scheme_paillier, self.scheme_dgk, l)
        sess = InitiatorSession(eng, l, nonces, self.wire_format)
        try:
            z_msg = sess.phase_a1(x, y)                                                                   # step 1 + randomize
            await self.communicator.send(self.other_party, z_msg, msg_id=f"step_1_session_{session_id}")
            del z_msg
            db_msg = await self.communicator.recv(self.other_party, msg_id=f"step_4b_session_{session_id}")
            c_msg = sess.phase_a2(db_msg)                                                                 # steps 3, 4c-4i + randomize
            del db_msg
            await self.communicator.send(self.other_party, c_msg, msg_id=f"step_4i_session_{session_id}")
            del c_msg
            enc3_msg = await self.communicator.recv(self.other_party, msg_id=f"step_5_session_{session_id}")
            return paillier_batch(sess.phase_a3(enc3_msg), self.scheme_paillier)                          # steps 6 + 7
        finally:
            sess.close()

    async def receive_encryption_schemes(self, session_id: int = 1) -> None:
        """Receive the schemes (public parts) from Bob and check them against pre-set ones
        (initiator.py:177-203, including its ValueErrors)."""
        if self.communicator is None:
            raise ValueError("Communicator not properly initialized.")
        message = await self.communicator.recv(self.other_party, msg_id=f"schemes_session_{session_id}")
        from . import wire
        if isinstance(message, (bytes, bytearray, memoryview)):
            received_paillier, received_dgk = wire.unpack_schemes(bytes(message))
        else:
            received_paillier, received_dgk = message
        differs = lambda mine, got: not wire.same_public_keys(mine, got)  # noqa: E731  (public parts travel, public parts compare)
        if self._scheme_paillier is None:
            self._scheme_paillier = received_paillier
        elif differs(self.scheme_paillier, received_paillier):
            raise ValueError("Readily available Paillier scheme and received Paillier scheme are different.")
        if self._scheme_dgk is None:
            self._scheme_dgk = received_dgk
        elif differs(self.scheme_dgk, received_dgk):
            raise ValueError("Readily available DGK scheme and received DGK scheme are different.")

    @staticmethod
    def shuffle(values, perm: Optional[np.ndarray] = None):
        """Shuffle in random order (initiator.py:212-226).  A plain list is shuffled like the reference does (uniform
        permutation from ``secrets``); a batch ``[B][n][...]`` array is shuffled independently per comparison along axis 1
        (``perm``: ``[B][n]``, output position k takes element ``perm[b][k]``)."""
        if isinstance(values, np.ndarray) and values.ndim >= 2:
            if perm is None:
                perm = random_permutations(values.shape[0], values.shape[1])
            idx = np.asarray(perm, dtype=np.int64).reshape(values.shape[0], values.shape[1], *([1] * (values.ndim - 2)))
            return np.take_along_axis(values, idx, axis=1)
        values = list(values)
        order = random_permutations(1, len(values))[0] if perm is None else np.asarray(perm).reshape(-1)
        return [values[int(i)] for i in order]

    # ------------------------------------------------------------------ steps (static, as in the reference)
    @staticmethod
    def step_1(x_enc, y_enc, l: int, scheme_paillier, r: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
        """``[[z]] = [[y]] * [[x]]^-1 * [[2^l + r]] mod N^2`` (initiator.py:228-258).  Returns ``(z_enc, r)``."""
        N = int(scheme_paillier.public_key.n)
        assert (1 << (l + 2)) < N // 2
        eng = get_engine(scheme_paillier)
        x, y = as_limbs(x_enc, eng.kc), as_limbs(y_enc, eng.kc)
        if r is None:
            eng.ensure_rng()
            r = eng.rng_below(N, x.shape[0], eng.kN)
        r = as_limbs(r, eng.kN)
        return paillier_batch(eng.initiator_step1(x, y, r, None, l), scheme_paillier), r

    @staticmethod
    def step_3(r, l: int) -> np.ndarray:
        """``alpha = r mod 2^l`` as bits (initiator.py:260-270)."""
        r = np.ascontiguousarray(r, dtype=np.uint32)
        return to_bits_batch(low_bits(r, l), l)

    @staticmethod
    def step_4c(d_enc, r, scheme_dgk, scheme_paillier) -> np.ndarray:
        """``[d] <- [0]`` (raw value 1) wherever ``r < (N-1)//2`` (initiator.py:272-291)."""
        N = int(scheme_paillier.public_key.n)
        r = np.ascontiguousarray(r, dtype=np.uint32)
        assert bool(np.all(less_than_const(r, N)))          # 0 <= r < N
        kn = (int(scheme_dgk.public_key.n).bit_length() + 31) // 32
        d = as_limbs(d_enc, kn).copy()
        small = less_than_const(r, (N - 1) // 2)
        d[small] = int_to_limbs(1, kn)
        return dgk_batch(d, scheme_dgk)

    @staticmethod
    def step_4d(alpha, beta_is_enc, scheme_dgk=None) -> np.ndarray:
        """``[alpha_i xor beta_i]``: ``[beta_i]`` or ``g * [beta_i]^-1`` (initiator.py:293-328).
        (The reference finds the scheme on the ciphertext objects; limb arrays need it passed.)"""
        eng_scheme = _scheme_of(beta_is_enc, scheme_dgk)
        eng = get_engine(None, eng_scheme)
        alpha = _bits(alpha)
        beta = as_limbs(beta_is_enc, eng.kn)
        B, l, kn = beta.shape
        flat = beta.reshape(B * l, kn)
        inv = eng.modinv(nat.MOD_DGK_N, flat)
        g = np.broadcast_to(int_to_limbs(eng.dgk_g, kn), flat.shape)
        flipped = eng.modmul(nat.MOD_DGK_N, inv, np.ascontiguousarray(g))
        return dgk_batch(np.where(alpha.reshape(B * l, 1) != 0, flipped, flat).reshape(B, l, kn), eng_scheme)

    @staticmethod
    def step_4e(r, alpha, alpha_is_xor_beta_is_enc, d_enc, scheme_paillier, scheme_dgk=None) -> Tuple[np.ndarray, np.ndarray]:
        """``[w_i]`` and ``alpha~ = (r - N) mod 2^l`` (initiator.py:330-384)."""
        eng_scheme = _scheme_of(alpha_is_xor_beta_is_enc, scheme_dgk)
        eng = get_engine(None, eng_scheme)
        N = int(scheme_paillier.public_key.n)
        r = np.ascontiguousarray(r, dtype=np.uint32)
        x = as_limbs(alpha_is_xor_beta_is_enc, eng.kn)
        B, l, kn = x.shape
        mask = (1 << l) - 1
        alpha_tilde = to_bits_batch((low_bits(r, 64) - np.uint64(N & 0xFFFFFFFFFFFFFFFF)) & np.uint64(mask), l)
        alpha = _bits(alpha)
        d = as_limbs(d_enc, kn)
        d_inv = eng.modinv(nat.MOD_DGK_N, d)
        with_d = eng.modmul(nat.MOD_DGK_N, x.reshape(B * l, kn), np.ascontiguousarray(np.repeat(d_inv, l, axis=0)))
        w = np.where((alpha != alpha_tilde).reshape(B * l, 1), with_d, x.reshape(B * l, kn)).reshape(B, l, kn)
        return dgk_batch(w, eng_scheme), alpha_tilde

    @staticmethod
    def step_4f(w_is_enc, scheme_dgk=None) -> np.ndarray:
        """``[w_i] <- [w_i]^(2^i) mod n`` (initiator.py:386-410)."""
        eng_scheme = _scheme_of(w_is_enc, scheme_dgk)
        eng = get_engine(None, eng_scheme)
        w = as_limbs(w_is_enc, eng.kn)
        B, l, kn = w.shape
        ew = (l + 31) // 32 + 1
        exps = np.tile(ints_to_limbs([1 << i for i in range(l)], ew), (B, 1))
        return dgk_batch(eng.modexp(nat.MOD_DGK_N, w.reshape(B * l, kn), exps, exp_bits=l).reshape(B, l, kn), eng_scheme)

    @staticmethod
    def step_4g(batch: int = 1, delta_a: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
        """Uniform bits ``delta_A`` and ``s = 1 - 2 delta_A`` (initiator.py:412-421)."""
        if delta_a is None:
            delta_a = np.frombuffer(secrets.token_bytes(batch), dtype=np.uint8) & 1
        delta_a = np.ascontiguousarray(delta_a, dtype=np.uint8)
        return 1 - 2 * delta_a.astype(np.int8), delta_a

    @staticmethod
    def step_4h(s, alpha, alpha_tilde, d_enc, beta_is_enc, w_is_enc, delta_a, scheme_dgk) -> np.ndarray:
        """``[c_i]`` with the prepended ``[c_{-1}]`` (initiator.py:423-485); output ``[B][l+1][kn]``.
        Built from the generic primitives (one batched call per operator of the reference expression)."""
        eng = get_engine(None, scheme_dgk)
        M = nat.MOD_DGK_N
        alpha, alpha_tilde = _bits(alpha), _bits(alpha_tilde)
        beta = as_limbs(beta_is_enc, eng.kn)
        w = as_limbs(w_is_enc, eng.kn)
        B, l, kn = beta.shape
        d = as_limbs(d_enc, kn)
        s = np.asarray(s).reshape(B)
        delta_a = np.asarray(delta_a).reshape(B)
        pk = scheme_dgk.public_key
        g = int(pk.g)
        g_neg = eng.dgk_g_neg
        one = int_to_limbs(1, kn)
        mm = lambda a, b: eng.modmul(M, np.ascontiguousarray(a), np.ascontiguousarray(b))  # noqa: E731
        g_s = np.where((s == 1)[:, None], int_to_limbs(g, kn)[None, :], int_to_limbs(g_neg, kn)[None, :])
        d_inv = eng.modinv(M, d)
        beta_inv = eng.modinv(M, beta.reshape(B * l, kn)).reshape(B, l, kn)
        out = np.empty((B, l + 1, kn), dtype=np.uint32)
        w_sum = np.tile(one, (B, 1))
        for i in range(l - 1, -1, -1):
            v = mm(g_s, np.where((alpha[:, i] != 0)[:, None], int_to_limbs(g, kn)[None, :], one[None, :]))
            e = alpha_tilde[:, i].astype(np.int8) - alpha[:, i].astype(np.int8)
            d_pow = np.where((e == 0)[:, None], one[None, :], np.where((e > 0)[:, None], d, d_inv))
            v = mm(v, d_pow)
            v = mm(v, beta_inv[:, i])
            cube = mm(mm(w_sum, w_sum), w_sum)
            out[:, i + 1] = mm(v, cube)
            w_sum = mm(w_sum, w[:, i])
        out[:, 0] = mm(w_sum, np.where((delta_a != 0)[:, None], int_to_limbs(g, kn)[None, :], one[None, :]))
        return dgk_batch(out, scheme_dgk)

    @staticmethod
    def step_4i(c_is_enc, scheme_dgk, do_shuffle: bool = True, r_blind: Optional[np.ndarray] = None,
                perm: Optional[np.ndarray] = None) -> np.ndarray:
        """Blind with ``r_i in [1, u-1]`` and shuffle (initiator.py:487-516)."""
        eng = get_engine(None, scheme_dgk)
        c = as_limbs(c_is_enc, eng.kn)
        B, n_c, kn = c.shape
        u = int(scheme_dgk.public_key.u)
        rbw = (u.bit_length() + 31) // 32
        if r_blind is None:
            eng.ensure_rng()
            r_blind = eng.rng_below(u, B * n_c, rbw, min_one=True).reshape(B, n_c, rbw)
        r_blind = np.ascontiguousarray(r_blind, dtype=np.uint32).reshape(B * n_c, -1)
        masked = eng.modexp(nat.MOD_DGK_N, c.reshape(B * n_c, kn), r_blind, exp_bits=u.bit_length()).reshape(B, n_c, kn)
        if not do_shuffle:
            return dgk_batch(masked, scheme_dgk)
        if perm is None:
            eng.ensure_rng()
            perm = eng.rng_permutations(n_c, B)
        perm = np.asarray(perm, dtype=np.int64)
        return dgk_batch(np.take_along_axis(masked, perm[:, :, None], axis=1), scheme_dgk)

    @staticmethod
    def step_6(delta_a, delta_b_enc, scheme_paillier=None) -> np.ndarray:
        """``[[beta < alpha]]``: ``[[delta_B]]`` or ``[[1]] * [[delta_B]]^-1`` (initiator.py:518-531)."""
        eng = get_engine(scheme_paillier)
        db = as_limbs(delta_b_enc, eng.kc)
        delta_a = np.asarray(delta_a).reshape(db.shape[0])
        inv = eng.modinv(nat.MOD_PAILLIER_N2, db)
        one_enc = np.tile(int_to_limbs(1 + eng.paillier_n, eng.kc), (db.shape[0], 1))
        flipped = eng.modmul(nat.MOD_PAILLIER_N2, inv, one_enc)
        return paillier_batch(np.where((delta_a == 1)[:, None], db, flipped), scheme_paillier if scheme_paillier is not None else getattr(delta_b_enc, 'scheme', None))

    @staticmethod
    def step_7(zeta_1_enc, zeta_2_enc, r, l: int, beta_lt_alpha_enc, scheme_paillier) -> np.ndarray:
        """``[[x <= y]] = [[zeta]] * ([[r div 2^l]] * [[beta < alpha]])^-1`` (initiator.py:533-564)."""
        eng = get_engine(scheme_paillier)
        N = eng.paillier_n
        r = np.ascontiguousarray(r, dtype=np.uint32)
        z1, z2 = as_limbs(zeta_1_enc, eng.kc), as_limbs(zeta_2_enc, eng.kc)
        zeta = np.where(less_than_const(r, (N - 1) // 2)[:, None], z1, z2)
        r_div = eng.paillier_encrypt(shift_right(r, l))
        t = eng.modmul(nat.MOD_PAILLIER_N2, r_div, as_limbs(beta_lt_alpha_enc, eng.kc))
        return paillier_batch(eng.modmul(nat.MOD_PAILLIER_N2, np.ascontiguousarray(zeta), eng.modinv(nat.MOD_PAILLIER_N2, t)), scheme_paillier)

    # ------------------------------------------------------------------ fused phases
    @staticmethod
    def steps_4c_to_4i(db_enc, r, l: int, scheme_dgk, scheme_paillier, delta_a, r_blind=None, perm=None, rho=None,
                       randomize: bool = True, do_shuffle: bool = True) -> BatchedDGKCiphertext:
        """Steps 3, 4c-4f, 4h, 4i (+ ``randomize()`` of every ``[c_i]``, initiator.py:153-154) in one device pass.
        ``db_enc``: ``[B][l+1][kn]`` (``[d]`` then ``[beta_i]``) as produced by ``KeyHolder.steps_4a_4b``; ``delta_a``: the ``[B]`` coins
        of ``step_4g`` (drawn by the caller, as in the reference flow, and needed again by ``steps_6_7``).  The other nonces
        default to the device CSPRNG; pass arrays to inject them (parity runs)."""
        eng = get_engine(scheme_paillier, scheme_dgk, l)
        db = as_limbs(db_enc, eng.kn)
        B = db.shape[0]
        r = np.ascontiguousarray(r, dtype=np.uint32)
        assert bool(np.all(less_than_const(r, eng.paillier_n)))
        delta_a = np.ascontiguousarray(delta_a, dtype=np.uint8).reshape(B)
        eng.ensure_rng()
        if r_blind is None:
            r_blind = eng.rng_below(eng.u, B * (l + 1), eng.rb_words, min_one=True).reshape(B, l + 1, eng.rb_words)
        if perm is None and do_shuffle:
            perm = eng.rng_permutations(l + 1, B)
        if rho is None and randomize:
            rho = eng.rng_bits(eng._fixed_base[0][1], B * (l + 1), eng.rand_words).reshape(B, l + 1, eng.rand_words)
        c = eng.initiator_step4(db, r, delta_a, r_blind, perm if do_shuffle else None, rho if randomize else None, l)
        return dgk_batch(c, scheme_dgk, fresh=randomize)

    @staticmethod
    def steps_6_7(enc3, r, delta_a, l: int, scheme_paillier) -> BatchedPaillierCiphertext:
        """Steps 6 and 7 fused: ``enc3`` = ``[B][3][2kN]`` (``[[zeta_1]], [[zeta_2]], [[delta_B]]``)."""
        eng = get_engine(scheme_paillier)
        return paillier_batch(eng.initiator_step6_7(np.ascontiguousarray(enc3, dtype=np.uint32), np.ascontiguousarray(r, dtype=np.uint32),
                                                    np.ascontiguousarray(delta_a, dtype=np.uint8), l), scheme_paillier)
