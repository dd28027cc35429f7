"""Batched ``KeyHolder`` (Bob) -- same step names and argument order as the reference
(``tno/mpc/protocols/secure_comparison/keyholder.py:15-287``), every argument being a batch.

``steps_4a_4b`` and ``steps_4j_5`` are the fused phases (with the ``randomize()`` calls of
``perform_secure_comparison``, keyholder.py:106-108 and 126-128, when nonces are requested)."""
from __future__ import annotations

from typing import Any, Optional, Tuple

import numpy as np

from . import _native as nat
from .communicator import Communicator
from .limbs import ints_to_limbs
from .schemes import as_limbs, get_engine
from .utils import less_than_const, low_bits, random_below, random_bits, to_bits_batch


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
        if self.communicator is None:
            raise ValueError("Communicator not properly initialized.")
        self.session_id += 1
        session_id = self.session_id
        await self.make_and_send_encryption_schemes(session_id)
        l = self.l_maximum_bit_length
        nonces = nonces or {}
        from . import wire
        out = (lambda a, kind: wire.pack(a, kind)) if self.wire_format else (lambda a, kind: a)
        inp = (lambda m, kind: wire.unpack(m, kind) if isinstance(m, (bytes, bytearray, memoryview)) else m)
        z_enc = inp(await self.communicator.recv(self.other_party, msg_id=f"step_1_session_{session_id}"), wire.KIND_PAILLIER_CT)
        z, _beta = KeyHolder.step_2(z_enc, l, self.scheme_paillier)
        db_enc = KeyHolder.steps_4a_4b(z, l, self.scheme_dgk, self.scheme_paillier, rho=nonces.get("rho_db"), randomize=True)
        await self.communicator.send(self.other_party, out(db_enc, wire.KIND_DGK_VECTOR), msg_id=f"step_4b_session_{session_id}")
        c_is_enc = inp(await self.communicator.recv(self.other_party, msg_id=f"step_4i_session_{session_id}"), wire.KIND_DGK_VECTOR)
        _delta_b, enc3 = KeyHolder.steps_4j_5(c_is_enc, z, l, self.scheme_dgk, self.scheme_paillier, rho=nonces.get("rho_enc3"), randomize=True)
        await self.communicator.send(self.other_party, out(enc3, wire.KIND_PAILLIER_TRIPLE), msg_id=f"step_5_session_{session_id}")

    async def make_and_send_encryption_schemes(self, session_id: int = 1, key_length_paillier: int = 2048, v_bits_dgk: int = 160,
                                               n_bits_dgk: int = 2048) -> None:
        """Send the public parts of both schemes to Alice (keyholder.py:135-172).  Key GENERATION is a
        one-off host job outside the engine: the schemes must have been supplied to the constructor."""
        if self.communicator is None:
            raise ValueError("Communicator not properly initialized.")
        if self._scheme_paillier is None or self._scheme_dgk is None:
            raise ValueError("No scheme has been initialized: the batched KeyHolder does not generate keys (host-side, one-off).")
        if self.wire_format:
            from . import wire
            message = wire.pack_schemes(self.scheme_paillier, self.scheme_dgk)      # public parts only, one blob
        else:
            pub = lambda s: s.public_only() if hasattr(s, "public_only") else s  # noqa: E731
            message = (pub(self.scheme_paillier), pub(self.scheme_dgk))
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
        return eng.keyholder_step4ab(as_limbs(z, eng.kN), l, None)[:, 0]

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
        return eng.keyholder_step4ab(z, l, None)[:, 1:]

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
        return enc3[:, 0], enc3[:, 1], enc3[:, 2]

    # ------------------------------------------------------------------ fused phases
    @staticmethod
    def steps_4a_4b(z, l: int, scheme_dgk, scheme_paillier, rho=None, randomize: bool = True) -> np.ndarray:
        """Steps 4a + 4b (+ randomize): ``[B][l+1][kn]`` = ``[d], [beta_0] .. [beta_{l-1}]``."""
        assert int(scheme_dgk.public_key.u) > (1 << (l + 2))
        eng = get_engine(scheme_paillier, scheme_dgk, l)
        z = as_limbs(z, eng.kN)
        if rho is None and randomize:
            rho = random_bits(eng._fixed_base[0][1], z.shape[0] * (l + 1), eng.rand_words).reshape(z.shape[0], l + 1, eng.rand_words)
        return eng.keyholder_step4ab(z, l, rho if randomize else None)

    @staticmethod
    def steps_4j_5(c_is_enc, z, l: int, scheme_dgk, scheme_paillier, rho=None, randomize: bool = True) -> Tuple[np.ndarray, np.ndarray]:
        """Steps 4j + 5 (+ randomize): returns ``(delta_b [B], enc3 [B][3][2kN])``."""
        eng = get_engine(scheme_paillier, scheme_dgk, l)
        z = as_limbs(z, eng.kN)
        if rho is None and randomize:
            rho = random_below(eng.paillier_n, z.shape[0] * 3, eng.kN, minimum=1).reshape(z.shape[0], 3, eng.kN)
        return eng.keyholder_step4j_5(as_limbs(c_is_enc, eng.kn), z, rho if randomize else None, l)
