"""GPU tests of the batched ciphertext containers: the reference's operator algebra on whole
batches against Python integers, the ``fresh`` flag, and the reference's way of writing ``Initiator.step_4h`` -- operators on
ciphertext objects (initiator.py:459-484) -- run on a batch and compared bit for bit with the fused kernel."""
import random
import warnings

import numpy as np
import pytest

from oracle.workload import make_schemes
from protocols.secure_comparison_b200 import (BatchedDGKCiphertext, BatchedPaillierCiphertext, Initiator, KeyHolder, get_engine, ints_to_limbs,
                                              limbs_to_ints, reset_engines)

pytestmark = pytest.mark.gpu
L_BITS = 16


@pytest.fixture(scope="module")
def setup():
    paillier, dgk = make_schemes(1024, 128, 20, L_BITS)
    reset_engines()
    eng = get_engine(paillier, dgk, L_BITS)
    return paillier, dgk, eng


def test_paillier_operator_algebra(setup):
    paillier, _, eng = setup
    N = paillier.public_key.n
    N2 = N * N
    rng = random.Random(1)
    B = 24
    av, bv = [rng.randrange(1, N2) for _ in range(B)], [rng.randrange(1, N2) for _ in range(B)]
    a = BatchedPaillierCiphertext(ints_to_limbs(av, eng.kc), paillier)
    b = BatchedPaillierCiphertext(ints_to_limbs(bv, eng.kc), paillier)
    assert a.scheme is paillier and not a.fresh and a.batch_shape == (B,)
    assert limbs_to_ints(a + b) == [x * y % N2 for x, y in zip(av, bv)]
    assert limbs_to_ints(-a) == [pow(x, -1, N2) for x in av]
    assert limbs_to_ints(a - b) == [x * pow(y, -1, N2) % N2 for x, y in zip(av, bv)]
    assert limbs_to_ints(a * 5) == [pow(x, 5, N2) for x in av]
    assert limbs_to_ints(3 * a) == [pow(x, 3, N2) for x in av]
    assert limbs_to_ints(a * -2) == [pow(pow(x, -1, N2), 2, N2) for x in av]
    assert limbs_to_ints(a * 0) == [1] * B and limbs_to_ints(a * 1) == av
    ks = [rng.randrange(-50, 1 << 40) for _ in range(B)]
    assert limbs_to_ints(a * np.array(ks, dtype=object)) == [pow(x, k, N2) for x, k in zip(av, ks)]
    # a + int: add(a, unsafe_encrypt(int)) with the integer encoding (negatives wrap mod N)
    assert limbs_to_ints(a + 7) == [x * (1 + 7 * N) % N2 for x in av]
    assert limbs_to_ints(-3 + a) == [x * (1 + (N - 3) * N) % N2 for x in av]
    assert limbs_to_ints(1 - a) == [pow(x, -1, N2) * (1 + N) % N2 for x in av]
    ms = [rng.randrange(-1000, 1000) for _ in range(B)]
    assert limbs_to_ints(a + ms) == [x * (1 + (m % N) * N) % N2 for x, m in zip(av, ms)]
    # broadcasting a single ciphertext over the batch, indexing, in-place forms
    assert limbs_to_ints(a + b[0]) == [x * bv[0] % N2 for x in av]
    c = a.copy()
    c += b
    c *= 3
    assert isinstance(c, BatchedPaillierCiphertext) and limbs_to_ints(c) == [pow(x * y % N2, 3, N2) for x, y in zip(av, bv)]
    assert limbs_to_ints(a) == av                       # the operands are untouched
    view = a[3:7]
    assert isinstance(view, BatchedPaillierCiphertext) and view.scheme is paillier and limbs_to_ints(view) == av[3:7]


def test_randomize_and_fresh_flag(setup):
    paillier, dgk, eng = setup
    N, n = paillier.public_key.n, dgk.public_key.n
    rng = random.Random(2)
    B = 9
    ms = [rng.randrange(1 << 16) for _ in range(B)]
    ct = BatchedPaillierCiphertext(eng.paillier_encrypt(eng.encode_paillier_plain(ms)), paillier)
    before = limbs_to_ints(ct)
    rho = [rng.randrange(1, N) for _ in range(B)]
    assert ct.randomize(ints_to_limbs(rho, eng.kN)) is ct and ct.fresh
    assert limbs_to_ints(ct) == [c * pow(r, N, N * N) % (N * N) for c, r in zip(before, rho)]
    assert [int(v) for v in limbs_to_ints(eng.paillier_decrypt(ct))] == ms
    assert ct.peek_value() is not None and ct.fresh
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        _ = ct + 1                                       # arithmetic on a fresh ciphertext spends its randomness
    assert any("ciphertext" in str(w.message) for w in caught) and not ct.fresh
    ct.randomize()                                       # device CSPRNG
    assert ct.fresh and [int(v) for v in limbs_to_ints(eng.paillier_decrypt(ct))] == ms
    ct.get_value()
    assert not ct.fresh
    g, h = dgk.public_key.g, dgk.public_key.h
    d = BatchedDGKCiphertext(ints_to_limbs([pow(g, m % 2, n) for m in ms], eng.kn), dgk)
    rho_d = [rng.getrandbits(dgk.randomness_bits) for _ in range(B)]
    base = limbs_to_ints(d)
    d.randomize(ints_to_limbs(rho_d, eng.rand_words))
    assert d.fresh and limbs_to_ints(d) == [c * pow(h, r, n) % n for c, r in zip(base, rho_d)]
    assert limbs_to_ints(d.copy() * 0 + [1, 0, -1, 2][0]) == [g % n] * B    # DGK a + int -> a * g^int


def reference_style_step_4h(s, alpha, alpha_tilde, d_enc, beta_is_enc, w_is_enc, delta_a, scheme_dgk, enc):
    """``Initiator.step_4h`` written the way the reference writes it -- nothing but operators on ciphertext objects
    (initiator.py:459-484) -- with every ciphertext being a whole batch.  ``enc(m)`` = ``scheme_dgk.unsafe_encrypt(m, apply_encoding=False)``."""
    l = len(beta_is_enc)
    c_is_enc = [enc(s) for _ in range(l)]
    w_sum = 0
    d_table = {-1: d_enc * -1, 0: d_enc * 0, 1: d_enc * 1}
    for i, alpha_i, alpha_tilde_i in zip(range(l - 1, -1, -1), reversed(alpha), reversed(alpha_tilde)):
        c_is_enc[i] += int(alpha_i) + d_table[alpha_tilde_i - alpha_i] - beta_is_enc[i] + 3 * w_sum
        w_sum += w_is_enc[i]
    c_is_enc.insert(0, delta_a + w_sum)
    return c_is_enc


@pytest.mark.parametrize("delta_a", [0, 1])
def test_reference_style_step_4h_on_a_batch(setup, delta_a):
    paillier, dgk, eng = setup
    l, N = L_BITS, paillier.public_key.n
    rng = random.Random(3 + delta_a)
    B = 40
    r_int = rng.randrange(N)                             # one r for the whole batch: alpha, alpha~ are then plain ints as in the reference
    r = np.tile(ints_to_limbs([r_int], eng.kN), (B, 1))
    z = ints_to_limbs([rng.randrange(N) for _ in range(B)], eng.kN)
    db = KeyHolder.steps_4a_4b(z, l, dgk, paillier, randomize=False)
    d_enc, beta_enc = db[:, 0], db[:, 1:]
    assert isinstance(d_enc, BatchedDGKCiphertext) and d_enc.scheme is dgk
    alpha = Initiator.step_3(r, l)
    d_enc = Initiator.step_4c(d_enc, r, dgk, paillier)
    xor_enc = Initiator.step_4d(alpha, beta_enc, dgk)
    w_enc, alpha_tilde = Initiator.step_4e(r, alpha, xor_enc, d_enc, paillier, dgk)
    w_enc = Initiator.step_4f(w_enc, dgk)
    s = 1 - 2 * delta_a
    kn = eng.kn
    enc = lambda m: BatchedDGKCiphertext(np.tile(ints_to_limbs([pow(dgk.public_key.g, m, dgk.public_key.n)], kn), (B, 1)), dgk)  # noqa: E731
    c_list = reference_style_step_4h(s, [int(v) for v in alpha[0]], [int(v) for v in alpha_tilde[0]], d_enc,
                                     [beta_enc[:, i] for i in range(l)], [w_enc[:, i] for i in range(l)], delta_a, dgk, enc)
    assert len(c_list) == l + 1 and all(isinstance(c, BatchedDGKCiphertext) for c in c_list)
    by_operators = np.stack([np.asarray(c) for c in c_list], axis=1)
    deltas = np.full(B, delta_a, dtype=np.uint8)
    by_primitives = Initiator.step_4h(np.full(B, s), alpha, alpha_tilde, d_enc, beta_enc, w_enc, deltas, dgk)
    fused = eng.initiator_step4(np.asarray(db), r, deltas, None, None, None, l)          # stops after step_4h
    assert np.array_equal(by_operators, np.asarray(by_primitives))
    assert np.array_equal(by_operators, fused)
    # step_4i the reference's way: c_i_enc *= r_i
    r_blind = [rng.randrange(1, dgk.public_key.u) for _ in range(l + 1)]
    masked = []
    for c, ri in zip(c_list, r_blind):
        c *= ri
        masked.append(np.asarray(c))
    rb = np.tile(ints_to_limbs(r_blind, eng.rb_words)[None], (B, 1, 1))
    assert np.array_equal(np.stack(masked, axis=1), eng.initiator_step4(np.asarray(db), r, deltas, rb, None, None, l))
