"""Device-resident party sessions: what ``Initiator`` / ``KeyHolder.perform_secure_comparison`` run between messages.

One object per protocol session and party.  It keeps the party's state in HBM across its phases (Alice: r and delta_A;
Bob: z), draws every nonce the caller did not inject with the device CSPRNG (``scb_rng_*``: the batch-scale replacement
of ``secrets.randbelow`` / ``choice`` and of the schemes' randomness pools, reference ``initiator.py:205-226, 250, 420,
512``; ``keyholder.py:174-179``), calls the fused ``*_dev`` phases of the C ABI and moves only the protocol messages over
PCIe -- each outgoing message is copied from the device straight into (pinned) message memory.  Everything a session
holds is private to it, so sessions of one party may interleave at their ``await`` points (reference
``test_secure_comparison.py:803-835``); a context executes one call at a time, which the single-threaded event loop of the
drivers guarantees.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import wire
from .device import DeviceArray
from .schemes import SchemeEngine, as_limbs, is_ciphertext_batch
from .utils import less_than_const


class _Session:
    def __init__(self, eng: SchemeEngine, l: int, nonces: Optional[dict], wire_format: bool):
        self.eng, self.l, self.nonces, self.wire_format = eng, l, dict(nonces or {}), wire_format
        self._held: List[DeviceArray] = []
        eng.ensure_rng()

    # ------------------------------------------------------------------ buffers
    def dev(self, shape, dtype=np.uint32) -> DeviceArray:
        a = self.eng.pool.device(shape, dtype)
        self._held.append(a)
        return a

    def drop(self, *arrays: DeviceArray) -> None:
        for a in arrays:
            if a in self._held:
                self._held.remove(a)
                a.release()

    def close(self) -> None:
        """Give every buffer back (after the stream has drained: a buffer may be reused by the next session at once)."""
        if self._held:
            self.eng.sync()
        for a in self._held:
            a.release()
        self._held.clear()

    def injected(self, name: str, shape, dtype=np.uint32) -> Optional[DeviceArray]:
        """Upload a host-supplied nonce array (parity runs), or None when that nonce is to be drawn on the device."""
        v = self.nonces.get(name)
        if v is None:
            return None
        v = np.ascontiguousarray(v, dtype=dtype)
        if int(np.prod(v.shape)) != int(np.prod(shape)):
            raise ValueError(f"nonces[{name!r}] must have shape {tuple(shape)}, got {tuple(v.shape)}")
        return self.dev(shape, dtype).upload(v.reshape(shape))

    # ------------------------------------------------------------------ messages
    def emit(self, d: DeviceArray, kind: int):
        """Device array -> outgoing message: a limb array, or one packed blob (``wire_format``); the device-to-host copy lands
        directly in the message's own pinned memory."""
        pool = self.eng.pool
        if self.wire_format:
            blob, payload = wire.blob_with_payload_view(lambda n: pool.pinned_owned((n,), np.uint8), kind, d.shape, d.dtype)
            d.download(payload)
            return blob
        return d.download(pool.pinned_owned(d.shape, d.dtype))

    def ingest(self, message, kind: int, shape, what: str) -> DeviceArray:
        """Incoming message -> device array.  The shape is checked against what THIS session expects before a single byte
        is copied: the other party controls the message (a short one must not make the library read past its end)."""
        if isinstance(message, (bytes, bytearray, memoryview)):
            arr = wire.unpack(message, kind)
        else:
            arr = np.asarray(message)
            if arr.dtype != np.uint32:
                raise ValueError(f"{what}: expected a uint32 limb array, got dtype {arr.dtype}")
        if tuple(arr.shape) != tuple(shape):
            raise ValueError(f"{what}: expected shape {tuple(shape)} for this session, got {tuple(arr.shape)}")
        return self.dev(shape).upload(arr)


class InitiatorSession(_Session):
    """Alice's phases A1 (step 1 + randomize), A2 (steps 3, 4c-4i + randomize), A3 (steps 6, 7)."""

    def phase_a1(self, x, y):
        eng, l = self.eng, self.l
        d_x, d_y = self._input_ciphertexts(x), self._input_ciphertexts(y)
        if d_x.shape != d_y.shape:
            raise ValueError(f"x and y must hold the same number of values ({d_x.shape[0]} vs {d_y.shape[0]})")
        self.B = B = d_x.shape[0]
        N = eng.paillier_n
        assert (1 << (l + 2)) < N // 2                                   # reference initiator.py:249
        r_host = self.nonces.get("r")
        if r_host is not None:
            r_host = np.ascontiguousarray(r_host, dtype=np.uint32).reshape(B, eng.kN)
            assert bool(np.all(less_than_const(r_host, N)))              # 0 <= r < N (reference initiator.py:286-288)
        self.d_r = self.injected("r", (B, eng.kN))
        if self.d_r is None:
            self.d_r = self.dev((B, eng.kN))
            eng.rng_below(N, B, eng.kN, False, out=self.d_r)
        d_rho = self.injected("rho_z", (B, eng.kN))
        if d_rho is None:
            d_rho = self.dev((B, eng.kN))
            eng.rng_below(N, B, eng.kN, True, out=d_rho)
        d_z = self.dev((B, eng.kc))
        eng.initiator_step1_dev(d_x, d_y, self.d_r, d_rho, l, d_z, B)
        msg = self.emit(d_z, wire.KIND_PAILLIER_CT)
        self.drop(d_x, d_y, d_rho, d_z)
        return msg

    def _input_ciphertexts(self, v) -> DeviceArray:
        """``x`` / ``y`` of ``perform_secure_comparison``: ciphertexts are uploaded as they are, anything else is a plaintext
        and is encrypted first -- ``scheme.unsafe_encrypt(x)`` (reference ``initiator.py:93-102``)."""
        eng = self.eng
        if is_ciphertext_batch(v):
            raw = as_limbs(v, eng.kc)
            raw = raw.reshape(-1, eng.kc)
            return self.dev(raw.shape).upload(raw)
        plain = eng.encode_paillier_plain(v)
        d_plain = self.dev(plain.shape).upload(plain)
        d_ct = self.dev((plain.shape[0], eng.kc))
        eng.paillier_encrypt_dev(d_plain, None, d_ct, plain.shape[0])
        self.drop(d_plain)
        return d_ct

    def phase_a2(self, db_message, do_shuffle: bool = True):
        eng, l, B = self.eng, self.l, self.B
        d_db = self.ingest(db_message, wire.KIND_DGK_VECTOR, (B, l + 1, eng.kn), "step_4b message")
        self.d_delta_a = self.injected("delta_a", (B,), np.uint8)
        if self.d_delta_a is None:
            self.d_delta_a = self.dev((B,), np.uint8)
            eng.rng_coins(B, out=self.d_delta_a)
        d_rb = self.injected("r_blind", (B, l + 1, eng.rb_words))
        if d_rb is None:
            d_rb = self.dev((B, l + 1, eng.rb_words))
            eng.rng_below(eng.u, B * (l + 1), eng.rb_words, True, out=d_rb)
        d_perm = None
        if do_shuffle:
            d_perm = self.injected("perm", (B, l + 1), np.uint8)
            if d_perm is None:
                d_perm = self.dev((B, l + 1), np.uint8)
                eng.rng_permutations(l + 1, B, out=d_perm)
        d_rho = self.injected("rho_c", (B, l + 1, eng.rand_words))
        if d_rho is None:
            d_rho = self.dev((B, l + 1, eng.rand_words))
            eng.rng_bits(eng._fixed_base[0][1], B * (l + 1), eng.rand_words, out=d_rho)
        d_c = self.dev((B, l + 1, eng.kn))
        eng.initiator_step4_dev(d_db, self.d_r, self.d_delta_a, d_rb, eng.rb_words, d_perm, d_rho, eng.rand_words, l, d_c, B)
        msg = self.emit(d_c, wire.KIND_DGK_VECTOR)
        self.drop(d_db, d_rb, d_rho, d_c, *([d_perm] if d_perm is not None else []))
        return msg

    def phase_a3(self, enc3_message) -> np.ndarray:
        eng, l, B = self.eng, self.l, self.B
        d_enc3 = self.ingest(enc3_message, wire.KIND_PAILLIER_TRIPLE, (B, 3, eng.kc), "step_5 message")
        d_res = self.dev((B, eng.kc))
        eng.initiator_step6_7_dev(d_enc3, self.d_r, self.d_delta_a, l, d_res, B)
        return d_res.download()                                           # an ordinary array: it outlives the engine's pools


class KeyHolderSession(_Session):
    """Bob's phases B1 (step 2, steps 4a, 4b + randomize) and B2 (steps 4j, 5 + randomize)."""

    def phase_b1(self, z_message):
        eng, l = self.eng, self.l
        assert eng.u > (1 << (l + 2))                                     # reference keyholder.py:212
        if isinstance(z_message, (bytes, bytearray, memoryview)):
            shape = wire.unpack(z_message, wire.KIND_PAILLIER_CT).shape
        else:
            shape = np.shape(z_message)
        if len(shape) != 2 or shape[1] != eng.kc:
            raise ValueError(f"step_1 message: expected [B][{eng.kc}] Paillier ciphertext limbs, got shape {tuple(shape)}")
        self.B = B = int(shape[0])
        d_zenc = self.ingest(z_message, wire.KIND_PAILLIER_CT, (B, eng.kc), "step_1 message")
        self.d_z = self.dev((B, eng.kN))
        eng.paillier_decrypt_dev(d_zenc, self.d_z, B)
        d_rho = self.injected("rho_db", (B, l + 1, eng.rand_words))
        if d_rho is None:
            d_rho = self.dev((B, l + 1, eng.rand_words))
            eng.rng_bits(eng._fixed_base[0][1], B * (l + 1), eng.rand_words, out=d_rho)
        d_db = self.dev((B, l + 1, eng.kn))
        eng.keyholder_step4ab_dev(self.d_z, l, d_rho, eng.rand_words, d_db, B)
        msg = self.emit(d_db, wire.KIND_DGK_VECTOR)
        self.drop(d_zenc, d_rho, d_db)
        return msg

    def phase_b2(self, c_message):
        eng, l, B = self.eng, self.l, self.B
        d_c = self.ingest(c_message, wire.KIND_DGK_VECTOR, (B, l + 1, eng.kn), "step_4i message")
        d_rho = self.injected("rho_enc3", (B, 3, eng.kN))
        if d_rho is None:
            d_rho = self.dev((B, 3, eng.kN))
            eng.rng_below(eng.paillier_n, 3 * B, eng.kN, True, out=d_rho)
        d_delta_b = self.dev((B,), np.uint8)
        d_enc3 = self.dev((B, 3, eng.kc))
        eng.keyholder_step4j_5_dev(d_c, self.d_z, d_rho, l, d_delta_b, d_enc3, B)
        msg = self.emit(d_enc3, wire.KIND_PAILLIER_TRIPLE)
        self.drop(d_c, d_rho, d_delta_b, d_enc3)
        return msg
