"""CPU test: what the KeyHolder puts on the wire for ``schemes_session_k`` never contains secret key material
(reference keyholder.py:168-172 sends scheme objects whose serialisation drops the secret key; the batched KeyHolder
builds explicit public-only holders)."""
import asyncio
import os

from protocols.secure_comparison_b200 import InMemoryCommunicator, KeyHolder
from protocols.secure_comparison_b200.keyfile import schemes_from_json

KEYS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keys_p1024_d128_v20_l16.json")


def _secret_values(paillier, dgk):
    sk, dk = paillier.secret_key, dgk.secret_key
    return {int(v) for v in (sk.lambda_, sk.mu, sk.p, sk.q, dk.p, dk.v_p, dk.q, dk.v_q) if v is not None}


def test_schemes_message_is_public_only_in_both_encodings():
    paillier, dgk = schemes_from_json(KEYS)
    assert paillier.secret_key is not None and dgk.secret_key is not None
    secrets_ = _secret_values(paillier, dgk)
    for wire_format in (False, True):
        store = {}
        bob = KeyHolder(16, InMemoryCommunicator(store), "alice", scheme_paillier=paillier, scheme_dgk=dgk, wire_format=wire_format)
        asyncio.run(bob.make_and_send_encryption_schemes(1))
        message = store["schemes_session_1"]
        if wire_format:
            assert isinstance(message, bytes)
            for v in secrets_:
                assert v.to_bytes((v.bit_length() + 7) // 8, "little") not in message
        else:
            got_p, got_d = message
            assert got_p is not paillier and got_d is not dgk
            assert got_p.secret_key is None and got_d.secret_key is None
            assert got_p.public_key.n == paillier.public_key.n
            assert (got_d.public_key.n, got_d.public_key.g, got_d.public_key.h, got_d.public_key.u, got_d.public_key.t) == \
                   (dgk.public_key.n, dgk.public_key.g, dgk.public_key.h, dgk.public_key.u, dgk.public_key.t)

            def walk(obj, depth=0):
                for v in vars(obj).values():
                    if isinstance(v, int):
                        assert v not in secrets_
                    elif hasattr(v, "__dict__") and depth < 3:
                        walk(v, depth + 1)
            walk(got_p)
            walk(got_d)
