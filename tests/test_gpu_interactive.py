"""GPU test of the batched interactive drivers (Initiator / KeyHolder.perform_secure_comparison) over an in-memory
communicator -- the reference's own way of testing two parties in one event loop (test/conftest.py:162-198,
test_secure_comparison.py:127-147): message ids, scheme exchange and equality check, both message encodings
(numpy arrays and the packed wire format), two concurrent sessions, and bit-exactness against the CPU oracle."""
import asyncio
import random

import numpy as np
import pytest

from helpers import pack_case_arrays
from oracle.protocol import Oracle
from oracle.workload import draw_inputs, draw_nonces, encrypt_input, make_schemes
from protocols.secure_comparison_b200 import Initiator, KeyHolder, get_engine, limbs_to_ints, reset_engines

pytestmark = pytest.mark.gpu


def DictionaryCommunicator(store, log):
    """The package's loop-back transport on a shared mailbox (the counterpart of the reference's test double)."""
    from protocols.secure_comparison_b200 import InMemoryCommunicator
    return InMemoryCommunicator(store, log)


@pytest.mark.parametrize("wire_format", [False, True], ids=["arrays", "packed"])
def test_interactive_batch_matches_oracle(wire_format):
    l = 16
    paillier, dgk = make_schemes(1024, 128, 20, l)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    oracle = Oracle(paillier, dgk, "auto")
    rng = random.Random(11)
    pairs = draw_inputs(rng, 64, l)
    cases = [{"x": x, "y": y, "nonces": draw_nonces(rng, paillier, dgk, l),
              "x_enc": encrypt_input(paillier, x, rng.randrange(1, paillier.public_key.n)),
              "y_enc": encrypt_input(paillier, y, rng.randrange(1, paillier.public_key.n))} for x, y in pairs]
    arr = pack_case_arrays(cases, l, eng.kN, eng.kn, eng.rand_words, eng.rb_words)
    store, log = {}, []
    alice = Initiator(l, DictionaryCommunicator(store, log), "bob", wire_format=wire_format)      # receives the schemes from Bob
    bob = KeyHolder(l, DictionaryCommunicator(store, log), "alice", scheme_paillier=paillier, scheme_dgk=dgk, wire_format=wire_format)
    a_nonces = {k: arr[k] for k in ("r", "rho_z", "delta_a", "r_blind", "perm", "rho_c")}
    b_nonces = {"rho_db": arr["rho_db"], "rho_enc3": arr["rho_enc3"]}

    async def both():
        return await asyncio.gather(alice.perform_secure_comparison(arr["x"], arr["y"], a_nonces), bob.perform_secure_comparison(b_nonces))

    result, _ = asyncio.run(both())
    ts = [oracle.compare(c["x_enc"], c["y_enc"], l, c["nonces"], interactive=True) for c in cases]
    assert limbs_to_ints(result) == [t.result_enc for t in ts]
    assert limbs_to_ints(eng.paillier_decrypt(result)) == [int(x <= y) for x, y in pairs]
    assert [m for m, _, _ in log] == ["schemes_session_1", "step_1_session_1", "step_4b_session_1", "step_4i_session_1", "step_5_session_1"]
    assert all((kind == "bytes") == wire_format for _, kind, _ in log)
    assert alice.scheme_paillier.public_key.n == paillier.public_key.n and getattr(alice.scheme_paillier, "secret_key", None) is None


def test_two_sessions_and_scheme_mismatch():
    """Second session reuses the schemes Alice already holds (equality check passes); a different key raises ValueError
    (initiator.py:192-203, test_secure_comparison.py:838-865)."""
    l = 16
    paillier, dgk = make_schemes(1024, 128, 20, l)
    other_paillier, other_dgk = make_schemes(512, 192, 24, 8)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    x = eng.paillier_encrypt(np.array([[5] + [0] * (eng.kN - 1), [9] + [0] * (eng.kN - 1)], dtype=np.uint32))
    y = eng.paillier_encrypt(np.array([[7] + [0] * (eng.kN - 1), [2] + [0] * (eng.kN - 1)], dtype=np.uint32))
    store, log = {}, []
    alice = Initiator(l, DictionaryCommunicator(store, log), "bob", wire_format=True)
    bob = KeyHolder(l, DictionaryCommunicator(store, log), "alice", scheme_paillier=paillier, scheme_dgk=dgk, wire_format=True)

    async def run_once():
        return await asyncio.gather(alice.perform_secure_comparison(x, y), bob.perform_secure_comparison())

    for session in (1, 2):                               # fresh randomness from `secrets` each time
        result, _ = asyncio.run(run_once())
        assert limbs_to_ints(eng.paillier_decrypt(result)) == [1, 0]
        assert alice.session_id == bob.session_id == session
    mallory = KeyHolder(l, DictionaryCommunicator(store, log), "alice", scheme_paillier=other_paillier, scheme_dgk=other_dgk,
                        session_id=2, wire_format=True)

    async def mismatch():
        await mallory.make_and_send_encryption_schemes(3)
        await alice.receive_encryption_schemes(3)

    with pytest.raises(ValueError):
        asyncio.run(mismatch())
