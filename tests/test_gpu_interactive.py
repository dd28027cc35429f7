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
    assert all((kind in ("bytes", "memoryview")) == wire_format for _, kind, _ in log)
    assert all(size > 0 for _, _, size in log[1:])
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


def _big_keys():
    from bench import key_file
    from protocols.secure_comparison_b200.keyfile import schemes_from_json
    return schemes_from_json(key_file(2048, 2048, 16))


def test_unencrypted_inputs_are_encrypted_first():
    """``perform_secure_comparison(x, y)`` with plaintext numbers, negatives included (reference initiator.py:93-102,
    test_secure_comparison.py:27-37, 776-800), at Paillier-2048 / DGK-2048; also one side encrypted, one side plain."""
    l = 16
    paillier, dgk = _big_keys()
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    pairs = [(-400, -383), (-1, 0), (0, 2), (1, 10), (230, 269), (1508, 2408), (3122, 6048), (4250, 7804), (8668, 9015), (7, 7), (9015, 8668), (0, -1)]
    xs, ys = [p[0] for p in pairs], [p[1] for p in pairs]
    store = {}
    alice = Initiator(l, DictionaryCommunicator(store, None), "bob", wire_format=True)
    bob = KeyHolder(l, DictionaryCommunicator(store, None), "alice", scheme_paillier=paillier, scheme_dgk=dgk, wire_format=True)

    async def run(x, y):
        return (await asyncio.gather(alice.perform_secure_comparison(x, y), bob.perform_secure_comparison()))[0]

    want = [int(x <= y) for x, y in pairs]
    assert limbs_to_ints(eng.paillier_decrypt(asyncio.run(run(xs, ys)))) == want
    y_enc = eng.paillier_encrypt(eng.encode_paillier_plain(ys))
    assert limbs_to_ints(eng.paillier_decrypt(asyncio.run(run(np.array(xs), y_enc)))) == want
    assert limbs_to_ints(eng.paillier_decrypt(asyncio.run(run(23, 42)))) == [1]          # the README example, scalars


def test_parallel_runs_on_one_player_pair():
    """Two sessions in flight on ONE Initiator / KeyHolder pair (reference test_secure_comparison.py:803-835), different batch
    sizes, device-drawn nonces: nothing of a session may live on the player objects."""
    l = 16
    paillier, dgk = _big_keys()
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    store = {}
    alice = Initiator(l, DictionaryCommunicator(store, None), "bob")
    bob = KeyHolder(l, DictionaryCommunicator(store, None), "alice", scheme_paillier=paillier, scheme_dgk=dgk)
    rng = random.Random(5)
    x1, y1 = [rng.randrange(1 << l) for _ in range(40)], [rng.randrange(1 << l) for _ in range(40)]
    x2, y2 = [11, 41, 12], [12, 41, 11]

    async def both():
        tasks = [asyncio.create_task(alice.perform_secure_comparison(x1, y1)), asyncio.create_task(alice.perform_secure_comparison(x2, y2)),
                 asyncio.create_task(bob.perform_secure_comparison()), asyncio.create_task(bob.perform_secure_comparison())]
        return await asyncio.gather(*tasks)

    r1, r2, _, _ = asyncio.run(both())
    assert limbs_to_ints(eng.paillier_decrypt(r1)) == [int(a <= b) for a, b in zip(x1, y1)]
    assert limbs_to_ints(eng.paillier_decrypt(r2)) == [1, 1, 0]
    assert alice.session_id == bob.session_id == 2


def test_malformed_messages_are_rejected_before_any_copy():
    """A message whose shape does not fit the session (the other party controls it) raises ValueError instead of making the
    library read past the end of the buffer."""
    from protocols.secure_comparison_b200 import wire
    from protocols.secure_comparison_b200.sessions import InitiatorSession, KeyHolderSession
    l = 16
    paillier, dgk = make_schemes(1024, 128, 20, l)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    a = InitiatorSession(eng, l, None, False)
    a.phase_a1([1, 2, 3], [3, 2, 1])
    short = np.zeros((2, l + 1, eng.kn), dtype=np.uint32)
    with pytest.raises(ValueError):
        a.phase_a2(short)
    with pytest.raises(ValueError):
        a.phase_a2(wire.pack(short, wire.KIND_DGK_VECTOR))
    with pytest.raises(ValueError):
        a.phase_a2(np.zeros((3, l + 1, eng.kn), dtype=np.uint64))
    a.close()
    b = KeyHolderSession(eng, l, None, False)
    with pytest.raises(ValueError):
        b.phase_b1(np.zeros((3, eng.kc - 1), dtype=np.uint32))
    b.phase_b1(np.ones((3, eng.kc), dtype=np.uint32))
    with pytest.raises(ValueError):
        b.phase_b2(np.zeros((3, l, eng.kn), dtype=np.uint32))
    b.close()
    with pytest.raises(ValueError):
        eng.initiator_step6_7(np.zeros((3, 2, eng.kc), dtype=np.uint32), np.zeros((3, eng.kN), dtype=np.uint32), np.zeros(3, dtype=np.uint8), l)
    with pytest.raises(ValueError):
        eng.keyholder_step4j_5(np.zeros((3, l, eng.kn), dtype=np.uint32), np.zeros((3, eng.kN), dtype=np.uint32), None, l)


@pytest.mark.parametrize("wire_format", [False, True], ids=["arrays", "packed"])
def test_empty_and_single_comparison_through_the_drivers(wire_format):
    """Edge batches of the interactive flow: B = 0 (every message is an empty array) and B = 1 (what the reference does)."""
    l = 16
    paillier, dgk = make_schemes(1024, 128, 20, l)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    store = {}
    alice = Initiator(l, DictionaryCommunicator(store, None), "bob", wire_format=wire_format)
    bob = KeyHolder(l, DictionaryCommunicator(store, None), "alice", scheme_paillier=paillier, scheme_dgk=dgk, wire_format=wire_format)

    async def run(x, y):
        return (await asyncio.gather(alice.perform_secure_comparison(x, y), bob.perform_secure_comparison()))[0]

    empty = asyncio.run(run([], []))
    assert np.asarray(empty).shape == (0, eng.kc)
    one = asyncio.run(run([-5], [3]))
    assert np.asarray(one).shape == (1, eng.kc) and limbs_to_ints(eng.paillier_decrypt(one)) == [1]
    with pytest.raises(ValueError):
        asyncio.run(run([1, 2], [1]))                       # x and y of different lengths
    store.clear()
