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
    import torch
    paillier, dgk, eng = setup
    l, B = 16, 512
    eng.rng_seed()
    rng = np.random.default_rng(3)
    x = rng.integers(0, 1 << l, size=B, dtype=np.uint64)
    y = rng.integers(0, 1 << l, size=B, dtype=np.uint64)
    x[:3], y[:3] = [23, 42, 7], [42, 23, 7]
    plain = np.zeros((2, B, eng.kN), dtype=np.uint32)
    plain[0, :, 0], plain[1, :, 0] = x.astype(np.uint32), y.astype(np.uint32)
    N = paillier.public_key.n
    i32 = dict(dtype=torch.int32, device="cuda")
    dev = lambda *shape: torch.empty(shape, **i32)  # noqa: E731
    d_rho_in, d_r, d_rho_z, d_rho3 = dev(2 * B, eng.kN), dev(B, eng.kN), dev(B, eng.kN), dev(3 * B, eng.kN)
    d_rblind, d_rho_c, d_rho_db = dev(B * (l + 1), eng.rb_words), dev(B * (l + 1), eng.rand_words), dev(B * (l + 1), eng.rand_words)
    d_da = torch.empty((B,), dtype=torch.uint8, device="cuda")
    d_perm = torch.empty((B, l + 1), dtype=torch.uint8, device="cuda")
    eng.rng_below(N, 2 * B, eng.kN, True, out=d_rho_in)
    eng.rng_below(N, B, eng.kN, False, out=d_r)                                   # r = randbelow(N)
    eng.rng_below(N, B, eng.kN, True, out=d_rho_z)
    eng.rng_below(N, 3 * B, eng.kN, True, out=d_rho3)
    eng.rng_below(dgk.public_key.u, B * (l + 1), eng.rb_words, True, out=d_rblind)  # r_i in [1, u-1]
    eng.rng_bits(dgk.randomness_bits, B * (l + 1), eng.rand_words, out=d_rho_c)
    eng.rng_bits(dgk.randomness_bits, B * (l + 1), eng.rand_words, out=d_rho_db)
    eng.rng_coins(B, out=d_da)
    eng.rng_permutations(l + 1, B, out=d_perm)
    d_plain = torch.from_numpy(plain.reshape(2 * B, eng.kN).view(np.int32)).cuda()
    d_in, d_z_enc, d_z = dev(2 * B, eng.kc), dev(B, eng.kc), dev(B, eng.kN)
    d_db, d_c, d_enc3, d_res, d_bit = dev(B, l + 1, eng.kn), dev(B, l + 1, eng.kn), dev(B, 3, eng.kc), dev(B, eng.kc), dev(B, eng.kN)
    d_delta_b = torch.empty((B,), dtype=torch.uint8, device="cuda")
    lib, ctx, p = eng.lib, eng.ctx, nat._ptr
    torch.cuda.synchronize()
    eng._check(lib.scb_paillier_encrypt_dev(ctx, p(d_plain), p(d_rho_in), p(d_in), 2 * B))
    eng._check(lib.scb_initiator_step1_dev(ctx, p(d_in[:B]), p(d_in[B:]), p(d_r), p(d_rho_z), l, p(d_z_enc), B))
    eng._check(lib.scb_paillier_decrypt_dev(ctx, p(d_z_enc), p(d_z), B))
    eng._check(lib.scb_keyholder_step4ab_dev(ctx, p(d_z), l, p(d_rho_db), eng.rand_words, p(d_db), B))
    eng._check(lib.scb_initiator_step4_dev(ctx, p(d_db), p(d_r), p(d_da), p(d_rblind), eng.rb_words, p(d_perm), p(d_rho_c), eng.rand_words, l, p(d_c), B))
    eng._check(lib.scb_keyholder_step4j_5_dev(ctx, p(d_c), p(d_z), p(d_rho3), l, p(d_delta_b), p(d_enc3), B))
    eng._check(lib.scb_initiator_step6_7_dev(ctx, p(d_enc3), p(d_r), p(d_da), l, p(d_res), B))
    eng._check(lib.scb_paillier_decrypt_dev(ctx, p(d_res), p(d_bit), B))
    eng.sync()
    bits = d_bit.cpu().numpy().view(np.uint32)
    assert np.array_equal(bits[:, 0].astype(np.uint8), (x <= y).astype(np.uint8)) and not bits[:, 1:].any()
    r_vals = limbs_to_ints(d_r.cpu().numpy().view(np.uint32))
    assert len(set(r_vals)) == B and all(v < N for v in r_vals)
