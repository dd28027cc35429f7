"""CPU tests of round-2 host logic: in-place wire blobs, the public-only scheme view, the build id of the library, the
plaintext / ciphertext classification of driver inputs, and the oracle-side parity gate bench.py runs (self-consistency on
a tiny batch: transcripts rebuilt from limb arrays must reproduce themselves, and a flipped bit must be reported)."""
import os
import random

import numpy as np
import pytest

from protocols.secure_comparison_b200 import wire
from protocols.secure_comparison_b200.limbs import ints_to_limbs


def test_blob_built_in_place_round_trips():
    shape = (5, 3, 7)
    held = []

    def alloc(n):
        held.append(np.zeros(n + 13, dtype=np.uint8))          # an allocator may hand out more than asked for
        return held[-1]

    blob, payload = wire.blob_with_payload_view(alloc, wire.KIND_DGK_VECTOR, shape, np.uint32)
    payload[...] = np.arange(int(np.prod(shape)), dtype=np.uint32).reshape(shape)
    assert isinstance(blob, memoryview) and blob.nbytes == len(wire.header(wire.KIND_DGK_VECTOR, np.uint32, shape)) + 4 * int(np.prod(shape))
    got = wire.unpack(blob, wire.KIND_DGK_VECTOR)
    assert got.shape == shape and np.array_equal(got, payload)
    assert bytes(blob) == wire.pack(payload, wire.KIND_DGK_VECTOR)      # identical to the copying encoder
    with pytest.raises(ValueError):
        wire.unpack(blob, wire.KIND_PAILLIER_CT)
    with pytest.raises(ValueError):
        wire.unpack(bytes(blob)[:-4], wire.KIND_DGK_VECTOR)


def test_public_only_view_drops_the_secret_keys():
    from protocols.secure_comparison_b200.keyfile import schemes_from_json
    keys = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "keys_p1024_d128_v20_l16.json")
    paillier, dgk = schemes_from_json(keys)
    p2, d2 = wire.public_only(paillier, dgk)
    assert p2.secret_key is None and d2.secret_key is None
    assert wire.same_public_keys(paillier, p2) and wire.same_public_keys(dgk, d2)
    assert d2.randomness_bits == dgk.randomness_bits


def test_library_was_built_from_this_tree():
    from protocols.secure_comparison_b200 import _native
    from protocols.secure_comparison_b200.build import build, built_id, source_id
    build()                                                        # no-op when the ids already agree
    assert built_id() == source_id() != ""
    lib = _native.load_library()
    assert lib.scb_build_id().decode() == source_id() and lib.scb_version() >= 200


def test_driver_input_classification():
    from protocols.secure_comparison_b200.schemes import is_ciphertext_batch
    limbs = np.zeros((4, 64), dtype=np.uint32)

    class Ct:
        def peek_value(self):
            return 5

    assert is_ciphertext_batch(limbs) and is_ciphertext_batch(Ct()) and is_ciphertext_batch([Ct(), Ct()])
    for plain in (5, -3, [1, 2, 3], np.array([1, 2, 3]), np.array([1, 2], dtype=np.uint64), np.zeros(4, dtype=np.uint32), (7, 8)):
        assert not is_ciphertext_batch(plain)


def test_oracle_parity_gate_detects_a_flipped_bit():
    from oracle.parallel import cases_from_arrays, check_messages_against_oracle, oracle_transcripts_parallel
    from oracle.workload import draw_nonces, encrypt_input, make_schemes
    from helpers import pack_case_arrays
    key_spec, l = (512, 192, 24, 8), 8
    paillier, dgk = make_schemes(*key_spec)
    rng = random.Random(3)
    kN = (paillier.public_key.n.bit_length() + 31) // 32
    kn = (dgk.public_key.n.bit_length() + 31) // 32
    rand_words = (dgk.randomness_bits + 31) // 32
    rb_words = (dgk.public_key.u.bit_length() + 31) // 32
    cases = []
    for x, y in [(3, 200), (200, 3), (77, 77), (0, 255), (255, 0), (9, 10)]:
        cases.append({"x": x, "y": y, "nonces": draw_nonces(rng, paillier, dgk, l),
                      "x_enc": encrypt_input(paillier, x, rng.randrange(1, paillier.public_key.n)),
                      "y_enc": encrypt_input(paillier, y, rng.randrange(1, paillier.public_key.n))})
    w = pack_case_arrays(cases, l, kN, kn, rand_words, rb_words)
    idx = [0, 2, 3, 5]
    rebuilt = cases_from_arrays(w, idx)
    assert [c["x_enc"] for c in rebuilt] == [cases[i]["x_enc"] for i in idx]
    assert [c["nonces"] for c in rebuilt] == [cases[i]["nonces"] for i in idx]
    ts = oracle_transcripts_parallel(key_spec, l, rebuilt, processes=2)
    got = {"z_enc": ints_to_limbs([t[0] for t in ts], 2 * kN), "z": ints_to_limbs([t[1] for t in ts], kN),
           "db": np.stack([ints_to_limbs(t[2], kn) for t in ts]), "c": np.stack([ints_to_limbs(t[3], kn) for t in ts]),
           "delta_b": np.array([t[4] for t in ts], dtype=np.uint8), "enc3": np.stack([ints_to_limbs(list(t[5]), 2 * kN) for t in ts]),
           "result": ints_to_limbs([t[6] for t in ts], 2 * kN)}
    res = check_messages_against_oracle(key_spec, l, w, got, idx, processes=2)
    assert res["bit_identical"] and res["comparisons_checked"] == 4 and res["ciphertexts_checked"] == 4 * (2 + 2 * (l + 1) + 1 + 3 + 1)
    got["c"][2, 4, 0] ^= 1
    res = check_messages_against_oracle(key_spec, l, w, got, idx, processes=2)
    assert not res["bit_identical"] and res["mismatches"] == [{"index": 3, "message": "c"}]
