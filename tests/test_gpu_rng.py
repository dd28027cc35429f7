"""GPU tests of the device CSPRNG (csrc/scb200_rng.cu): RFC 8439 known-answer block, every sampler bit for bit against
the Python model (tests/chacha_model.py), and a whole protocol pass driven by device-generated nonces."""
import struct

import numpy as np
import pytest

from chacha_model import Model
from oracle.workload import make_schemes
from protocols.secure_comparison_b200 import get_engine, limbs_to_ints, reset_engines
from protocols.secure_comparison_b200 import _native as nat

pytestmark = pytest.mark.gpu

KEY, NONCE = bytes(range(32)), bytes([7, 0, 0, 1, 9, 9, 9, 2])


@pytest.fixture(scope="module")
def setup():
    paillier, dgk = make_schemes(1024, 128, 20, 16)
    reset_engines()
    return paillier, dgk, get_engine(paillier, dgk, 16)


def test_rfc8439_block_on_device(setup):
    _, _, eng = setup
    eng.rng_seed(KEY, NONCE)
    out = eng.rng_block([1, 0x09000000, 0x4A000000, 0])
    assert out[:4].tolist() == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3]
    assert out[12:].tolist() == [0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]


def test_samplers_match_model(setup):
    paillier, dgk, eng = setup
    eng.rng_seed(KEY, NONCE)
    model = Model(KEY, NONCE)
    N, u = paillier.public_key.n, dgk.public_key.u
    big = (1 << 4095) + (1 << 777) + 5
    for bound, B, min_one in ((N, 300, False), (N, 257, True), (u, 1000, True), (3, 600, False), (2, 100, True), (big, 20, False)):
        k = (bound.bit_length() + 31) // 32
        got = limbs_to_ints(eng.rng_below(bound, B, k, min_one))
        assert got == model.below(bound, B, min_one), f"below({bound.bit_length()} bits, min_one={min_one})"
        assert all((1 if min_one else 0) <= v < bound for v in got)
    for bits, words, B in ((50, 2, 500), (400, 13, 200), (32, 1, 64), (513, 17, 40)):
        assert limbs_to_ints(eng.rng_bits(bits, B, words)) == model.bits(bits, B, words)
    assert eng.rng_coins(1300).tolist() == model.coins(1300)
    for n, B in ((17, 400), (65, 100), (1, 3), (2, 50)):
        assert eng.rng_permutations(n, B).tolist() == model.permutations(n, B)
    # statistical sanity: coins are balanced, permutations move every position
    coins = eng.rng_coins(1 << 16)
    assert abs(int(coins.sum()) - (1 << 15)) < 1200
    perms = eng.rng_permutations(17, 4096)
    assert all(abs(float((perms[:, 0] == v).mean()) - 1 / 17) < 0.02 for v in range(17))


def test_protocol_pass_with_device_nonces(setup):
    """All nonces generated on the GPU (fresh key from `secrets`), device-pointer phases end to end."""
    from helpers import run_device_pass
    paillier, dgk, eng = setup
    l, B = 16, 512
    rng = np.random.default_rng(3)
    x = rng.integers(0, 1 << l, size=B, dtype=np.uint64)
    y = rng.integers(0, 1 << l, size=B, dtype=np.uint64)
    x[:3], y[:3] = [23, 42, 7], [42, 23, 7]
    bits, r_vals = run_device_pass(eng, paillier, dgk, l, x, y)
    assert np.array_equal(bits[:, 0].astype(np.uint8), (x <= y).astype(np.uint8)) and not bits[:, 1:].any()
    assert len(set(r_vals)) == B and all(v < paillier.public_key.n for v in r_vals)


@pytest.mark.parametrize("pbits,dbits,vbits,l,B", [(2048, 2048, 160, 16, 8192), (2048, 2048, 160, 32, 4096), (4096, 3072, 160, 64, 1024)])
def test_larger_batches_decrypt_to_the_truth(pbits, dbits, vbits, l, B):
    """Property check at sizes the oracle cannot follow (BASELINE configs[1..3] key sizes and bit lengths): every decrypted
    comparison bit equals the plaintext truth, with equal pairs, extremes and adjacent values forced in."""
    from helpers import run_device_pass
    paillier, dgk = make_schemes(pbits, dbits, vbits, l)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    rng = np.random.default_rng(l)
    hi = (1 << l) - 1
    x = rng.integers(0, hi, size=B, dtype=np.uint64, endpoint=True)
    y = rng.integers(0, hi, size=B, dtype=np.uint64, endpoint=True)
    y[: B // 8] = x[: B // 8]                                        # equal pairs
    y[B // 8: B // 4] = np.minimum(x[B // 8: B // 4], np.uint64(hi - 1)) + np.uint64(1)     # y = x + 1
    x[-4:], y[-4:] = [0, hi, 0, hi], [0, hi, hi, 0]
    bits, _ = run_device_pass(eng, paillier, dgk, l, x, y)
    assert not bits[:, 1:].any()
    assert np.array_equal(bits[:, 0].astype(np.uint8), (x <= y).astype(np.uint8))
