"""Batched ``KeyHolder`` (Bob) -- same step names and argument order as the reference
(``tno/mpc/protocols/secure_comparison/keyholder.py:15-287``), every argument being a batch.

``steps_4a_4b`` and ``steps_4j_5`` are the fused phases (with the ``randomize()`` calls of
``perform_secure_comparison``, keyholder.py:106-108 and 126-128, when nonces are requested).  Ciphertext results are
``BatchedDGKCiphertext`` / ``BatchedPaillierCiphertext`` containers (``ciphertexts.py``); randomizer exponents that are not
injected come from the device CSPRNG."""
from __future__ import annotations

from typing import Any, Optional, Tuple

import numpy as np

from . import _native as nat
from .ciphertexts import BatchedDGKCiphertext, BatchedPaillierCiphertext, dgk_batch, paillier_batch
from .communicator import Communicator
from .schemes import as_limbs, get_engine
from .sessions import KeyHolderSession
from .utils import low_bits, to_bits_batch


class KeyHolder:
    """Player Bob in the secure comparison protocol, holds the keys."""

    def __init__(self, l_maximum_bit_length: int, communicator: Optional[Communicator] = None, other_party: str = "",
                 scheme_paillier: Any = None, scheme_dgk: Any = None, session_id: int = 0, *, wire_format: bool = False) -> None:
        self.wire_format = wire_format      # new, keyword-only: one packed blob per protocol message (wire.py)
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

    # ------------------------------------------------------------------ driver (keyholder.py:70-172)
    async def perform_secure_comparison(self, nonces: Optional[dict] = None) -> None:
        """All of Bob's steps for a batch (``keyholder.py:70-133``).  z stays on the device between his two phases; the randomizer
        exponents not injected through ``nonces`` (keys ``rho_db``, ``rho_enc3``) come from the device CSPRNG.  Sessions may overlap."""
        if self.communicator is None:
            raise ValueError("Communicator not properly initialized.")
        self.session_id += 1
        session_id = self.session_id
        await self.make_and_send_encryption_schemes(session_id)
        l = self.l_maximum_bit_length
        eng = get_engine(self.scheme_paillier, self.scheme_dgk, l)
        sess = KeyHolderSession(eng, l, nonces, self.wire_format)
        try:
            z_msg = await self.communicator.recv(self.other_party, msg_id=f"step_1_session_{session_id}")
            db_msg = sess.phase_b1(z_msg)                                                      # step 2, steps 4a, 4b + randomize
            del z_msg
            await self.communicator.send(self.other_party, db_msg, msg_id=f"step_4b_session_{session_id}")
            del db_msg
            c_msg = await self.communicator.recv(self.other_party, msg_id=f"step_4i_session_{session_id}")
            enc3_msg = sess.phase_b2(c_msg)                                                    # steps 4j, 5 + randomize
            del c_msg
            await self.communicator.send(self.other_party, enc3_msg, msg_id=f"step_5_session_{session_id}")
        finally:
            sess.close()

    async def make_and_send_encryption_schemes(self, session_id: int = 1, key_length_paillier: int = 2048, v_bits_dgk: int = 160,
                                               n_bits_dgk: int = 2048) -> None:
        """Send the PUBLIC parts of both schemes to Alice (keyholder.py:135-172).  What travels is always a freshly built
        public-only holder (n; n, g, h, u, t) or its packed form -- never the scheme objects themselves, which carry the secret
        keys.  Key GENERATION is a one-off host job outside the engine: the schemes must have been supplied to the constructor."""
        if self.communicator is None:
            raise ValueError("Communicator not properly initialized.")
        if self._scheme_paillier is None or self._scheme_dgk is None:
            raise ValueError("No scheme has been initialized: the batched KeyHolder does not generate keys (host-side, one-off).")
        from . import wire
        if self.wire_format:
            message = wire.pack_schemes(self.scheme_paillier, self.scheme_dgk)      # public parts only, one blob
        else:
            message = wire.public_only(self.scheme_paillier, self.scheme_dgk)
            if any(getattr(m, "secret_key", None) is not None for m in message):
                raise RuntimeError("refusing to send a scheme that still carries a secret key")
        await self.communicator.send(self.other_party, message, msg_id=f"schemes_session_{session_id}")

    # ------------------------------------------------------------------ steps
    @staticmethod
    def step_2(z_enc, l: int, scheme_paillier) -> Tuple[np.ndarray, np.ndarray]:
        """Decrypt ``[[z]]``; ``beta = z mod 2^l`` (keyholder.py:181-196).  Returns ``(z [B][kN], beta [B] uint64)``."""
        eng = get_engine(scheme_paillier)
        z = eng.paillier_decrypt(as_limbs(z_enc, eng.kc))
        return z, low_bits(z, l)

    @staticmethod
    def step_4a(z, scheme_dgk, scheme_paillier, l: int) -> np.ndarray:
        """``[d]``, ``d = (z < (N-1)//2)`` (keyholder.py:198-216)."""
        assert int(scheme_dgk.public_key.u) > (1 << (l + 2))
        eng = get_engine(scheme_paillier, scheme_dgk, l)
        return dgk_batch(eng.keyholder_step4ab(as_limbs(z, eng.kN), l, None)[:, 0], scheme_dgk)

    @staticmethod
    def step_4b(beta, l: int, scheme_dgk) -> np.ndarray:
        """``[beta_i]`` LSB first (keyholder.py:218-233).  ``beta``: ``[B]`` integers < 2^l."""
        eng = get_engine(None, scheme_dgk, l)
        beta = np.asarray(beta, dtype=np.uint64).reshape(-1)
        to_bits_batch(beta, l)                              # the reference's to_bits assertion
        # step_4ab wants z; any z with the right low bits gives the same [beta_i] (d is discarded here)
        z = np.zeros((beta.shape[0], eng.kN), dtype=np.uint32)
        z[:, 0] = (beta & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        if eng.kN > 1:
            z[:, 1] = (beta >> np.uint64(32)).astype(np.uint32)
        return dgk_batch(eng.keyholder_step4ab(z, l, None)[:, 1:], scheme_dgk)

    @staticmethod
    def step_4j(c_is_enc, scheme_dgk) -> np.ndarray:
        """``delta_B = any(is_zero(c_i))`` (keyholder.py:235-253).  Returns ``[B]`` uint8."""
        eng = get_engine(None, scheme_dgk)
        c = as_limbs(c_is_enc, eng.kn)
        if c.ndim == 2:
            c = c[None]
        return eng.dgk_is_zero(c).any(axis=1).astype(np.uint8)

    @staticmethod
    def step_5(z, l: int, delta_b, scheme_paillier) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """``[[zeta_1]], [[zeta_2]], [[delta_B]]`` (keyholder.py:255-287)."""
        eng = get_engine(scheme_paillier)
        enc3 = eng.keyholder_step5(as_limbs(z, eng.kN), np.asarray(delta_b, dtype=np.uint8).reshape(-1), None, l)
        return tuple(paillier_batch(enc3[:, i], scheme_paillier) for i in range(3))

    # ------------------------------------------------------------------ fused phases
    @staticmethod
    def steps_4a_4b(z, l: int, scheme_dgk, scheme_paillier, rho=None, randomize: bool = True) -> BatchedDGKCiphertext:
        """Steps 4a + 4b (+ randomize): ``[B][l+1][kn]`` = ``[d], [beta_0] .. [beta_{l-1}]``."""
        assert int(scheme_dgk.public_key.u) > (1 << (l + 2))
        eng = get_engine(scheme_paillier, scheme_dgk, l)
        z = as_limbs(z, eng.kN)
        if rho is None and randomize:
            eng.ensure_rng()
            rho = eng.rng_bits(eng._fixed_base[0][1], z.shape[0] * (l + 1), eng.rand_words).reshape(z.shape[0], l + 1, eng.rand_words)
        return dgk_batch(eng.keyholder_step4ab(z, l, rho if randomize else None), scheme_dgk, fresh=randomize)

    @staticmethod
    def steps_4j_5(c_is_enc, z, l: int, scheme_dgk, scheme_paillier, rho=None, randomize: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        """Steps 4j + 5 (+ randomize): returns ``(delta_b [B], enc3 [B][3][2kN])``."""
        eng = get_engine(scheme_paillier, scheme_dgk, l)
        z = as_limbs(z, eng.kN)
        if rho is None and randomize:
            eng.ensure_rng()
            rho = eng.rng_below(eng.paillier_n, z.shape[0] * 3, eng.kN, min_one=True).reshape(z.shape[0], 3, eng.kN)
        delta_b, enc3 = eng.keyholder_step4j_5(as_limbs(c_is_enc, eng.kn), z, rho if randomize else None, l)
        return delta_b, paillier_batch(enc3, scheme_paillier, fresh=randomize)
