"""CPU tests of the packed wire format (one blob per protocol message)."""
import numpy as np
import pytest

from protocols.secure_comparison_b200 import wire
from protocols.secure_comparison_b200.keyfile import schemes_from_json
import os

KEYS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keys_p1024_d128_v20_l16.json")


@pytest.mark.parametrize("shape,dtype", [((0, 64), np.uint32), ((5, 17, 64), np.uint32), ((7,), np.uint8), ((3, 2), np.uint64), ((1, 3, 128), np.uint32)])
def test_roundtrip(shape, dtype):
    rng = np.random.default_rng(1)
    a = rng.integers(0, 255, size=shape, dtype=np.uint64).astype(dtype)
    blob = wire.pack(a, wire.KIND_DGK_VECTOR)
    assert len(blob) == 8 + 4 * len(shape) + a.nbytes
    b = wire.unpack(blob, wire.KIND_DGK_VECTOR)
    assert b.dtype == a.dtype and b.shape == a.shape and np.array_equal(a, b)


def test_rejects_wrong_kind_and_corruption():
    blob = wire.pack(np.zeros((2, 4), np.uint32), wire.KIND_PAILLIER_CT)
    with pytest.raises(ValueError):
        wire.unpack(blob, wire.KIND_DGK_VECTOR)
    with pytest.raises(ValueError):
        wire.unpack(blob[:-1])
    with pytest.raises(ValueError):
        wire.unpack(b"XXXX" + blob[4:])
    with pytest.raises(TypeError):
        wire.pack(np.zeros(3, np.float32))


def test_schemes_roundtrip_public_parts_only():
    paillier, dgk = schemes_from_json(KEYS)
    blob = wire.pack_schemes(paillier, dgk)
    p2, d2 = wire.unpack_schemes(blob)
    assert p2.public_key.n == paillier.public_key.n and p2.secret_key is None
    assert wire.same_public_keys(d2, dgk) and d2.secret_key is None
    assert hex(dgk.secret_key.p)[2:].encode() not in blob and dgk.secret_key.p.to_bytes(32, "little")[:8] not in blob
    with pytest.raises(ValueError):
        wire.unpack_schemes(blob + b"\0")


def test_initiator_shuffle_mirrors_the_reference():
    """Initiator.shuffle (initiator.py:212-226): a uniform permutation of a list; per-comparison along axis 1 for batches."""
    from protocols.secure_comparison_b200.initiator import Initiator
    values = list(range(17))
    out = Initiator.shuffle(values)
    assert sorted(out) == values and values == list(range(17))          # input untouched
    batch = np.arange(2 * 5 * 3).reshape(2, 5, 3)
    out = Initiator.shuffle(batch, perm=np.array([[4, 3, 2, 1, 0], [0, 1, 2, 3, 4]]))
    assert np.array_equal(out[0], batch[0, ::-1]) and np.array_equal(out[1], batch[1])
    out = Initiator.shuffle(batch)
    assert sorted(out[1, :, 0].tolist()) == batch[1, :, 0].tolist()


def test_in_memory_communicator_round_trip():
    import asyncio
    from protocols.secure_comparison_b200 import Communicator, InMemoryCommunicator
    log = []
    alice = InMemoryCommunicator(log=log)
    bob = alice.peer()
    assert isinstance(alice, Communicator) and isinstance(bob, Communicator)

    async def talk():
        async def late_send():
            await asyncio.sleep(0.01)
            await bob.send("alice", wire.pack(np.arange(6, dtype=np.uint32).reshape(2, 3)), msg_id="m1")
        task = asyncio.ensure_future(late_send())
        got = await alice.recv("bob", msg_id="m1")                 # waits for the message
        await task
        return got

    got = asyncio.run(talk())
    assert np.array_equal(wire.unpack(got), np.arange(6, dtype=np.uint32).reshape(2, 3))
    assert log == [("m1", "bytes", 8 + 8 + 24)]
    quick = InMemoryCommunicator(timeout_s=0.01)
    with pytest.raises(KeyError):
        asyncio.run(quick.recv("nobody", msg_id="missing"))
