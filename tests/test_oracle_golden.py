"""CPU tests: the oracle restatement against the golden vectors minted from the unmodified reference,
the two big-integer backends against each other, and the reference's own per-step assertions
(test/unit/test_secure_comparison.py) re-run on the oracle's scheme stand-ins."""
import random

import pytest

from helpers import GOLDEN_FILES, load_golden
from oracle.bigint import GmpBackend, PythonBackend, gmp_available
from oracle.protocol import Oracle, from_bits, to_bits
from oracle.reference_runner import reference_available, run_reference
from oracle.schemes import PaillierCiphertext, next_prime
from oracle.workload import draw_inputs, draw_nonces, encrypt_input, make_schemes

L = 16


@pytest.mark.parametrize("name", GOLDEN_FILES)
@pytest.mark.parametrize("backend", ["cpython", "gmp"])
def test_oracle_reproduces_golden_vectors(name, backend):
    if backend == "gmp" and not gmp_available():
        pytest.skip("libgmp not loadable")
    l, paillier, dgk, cases = load_golden(name)
    oracle = Oracle(paillier, dgk, backend)
    if "2048" in name and backend == "cpython":
        cases = cases[:2]  # CPython pow on 4096-bit moduli is slow; GMP covers the rest
    for case in cases:
        for mode in ("interactive", "non_interactive"):
            got = oracle.compare(case["x_enc"], case["y_enc"], l, case["nonces"], interactive=(mode == "interactive"))
            assert got == case[mode], f"{name} {mode} x={case['x']} y={case['y']}"
            assert got.result_bit == int(case["x"] <= case["y"])


@pytest.mark.skipif(not reference_available(), reason="/root/reference only exists in the build container")
def test_oracle_matches_unmodified_reference_on_fresh_inputs():
    paillier, dgk = make_schemes(1024, 128, 20, L)
    oracle = Oracle(paillier, dgk, "cpython")
    rng = random.Random(4242)
    for x, y in draw_inputs(rng, 12, L, include_reference_pairs=False):
        nc = draw_nonces(rng, paillier, dgk, L)
        x_enc = encrypt_input(paillier, x, rng.randrange(1, paillier.public_key.n))
        y_enc = encrypt_input(paillier, y, rng.randrange(1, paillier.public_key.n))
        for interactive in (True, False):
            assert oracle.compare(x_enc, y_enc, L, nc, interactive) == run_reference(paillier, dgk, x_enc, y_enc, L, nc, interactive)


@pytest.mark.skipif(not gmp_available(), reason="libgmp not loadable")
def test_backends_agree():
    gmp, py = GmpBackend(), PythonBackend()
    rng = random.Random(1)
    for bits in (64, 1024, 2048, 4096):
        m = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        for _ in range(6):
            b, e = rng.getrandbits(bits) % m, rng.getrandbits(bits // 2)
            assert gmp.powmod(b, e, m) == py.powmod(b, e, m)
            assert gmp.mulmod(b, e, m) == py.mulmod(b, e, m)
            try:
                inv = py.invert(b, m)
            except ZeroDivisionError:
                with pytest.raises(ZeroDivisionError):
                    gmp.invert(b, m)
            else:
                assert gmp.invert(b, m) == inv
    for edge_b, edge_e in ((0, 5), (1, 0), (0, 0), (6, 1)):
        assert gmp.powmod(edge_b, edge_e, 7) == pow(edge_b, edge_e, 7)
    with pytest.raises(ZeroDivisionError):
        gmp.invert(0, 9)
    with pytest.raises(ZeroDivisionError):
        gmp.invert(3, 9)


# --------------------------------------------------------------------------------------------
# the reference's per-step assertions (test_secure_comparison.py:150-634) on the stand-in schemes
# --------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def schemes():
    paillier, dgk_full = make_schemes(1024, 128, 20, L, full_decryption=True)
    return paillier, dgk_full, Oracle(paillier, dgk_full, "cpython")


PAIRS = [(-400, -383), (-1, 0), (0, 2), (1, 10), (230, 269), (1508, 2408), (3122, 6048), (4250, 7804), (8668, 9015)]


@pytest.mark.parametrize("number", [v for pair in PAIRS for v in pair if v >= 0])
def test_bit_conversion(number):  # test_secure_comparison.py:150-163
    assert from_bits(to_bits(number, L)) == number


def test_to_bits_asserts_range():  # utils.py:16
    with pytest.raises(AssertionError):
        to_bits(1 << L, L)


def test_modulo_n_squared(schemes):  # test_secure_comparison.py:166-182
    paillier, _, oracle = schemes
    n2 = paillier.public_key.n_squared
    x_enc = oracle.paillier_encrypt_raw(0)
    y_enc = n2 - 2
    y = paillier.decrypt(PaillierCiphertext(y_enc, paillier), apply_encoding=False)
    r = n2
    z_enc = oracle.step_1(x_enc, y_enc, L, r)
    assert (y - 0 + (1 << L) + r) % n2 % paillier.public_key.n == oracle.paillier_decrypt_raw(z_enc)


@pytest.mark.parametrize("x,y", PAIRS)
def test_steps_algebra(schemes, x, y):
    """test_step_1/2/4b/4d/4e/4f/4h/4j/7 (test_secure_comparison.py:185-634) in one walk."""
    paillier, dgk, oracle = schemes
    N, u = paillier.public_key.n, dgk.public_key.u
    rng = random.Random(x * 1000003 + y)
    nc = draw_nonces(rng, paillier, dgk, L)
    dec = lambda raw: dgk.decrypt(type(dgk).ciphertext_cls(raw, dgk), apply_encoding=False)  # noqa: E731
    x_enc, y_enc = encrypt_input(paillier, x, 3), encrypt_input(paillier, y, 5)
    t = oracle.compare(x_enc, y_enc, L, nc, interactive=False)
    r = nc.r
    assert oracle.paillier_decrypt_raw(t.z_enc) == (y - x + (1 << L) + r) % N              # step 1
    assert t.beta == t.z % 2 ** L                                                           # step 2
    assert dec(t.d_enc_4a) == int(t.z < (N - 1) // 2)                                       # step 4a
    assert [dec(b) for b in t.beta_enc_4b] == to_bits(t.beta, L)                            # step 4b
    d = dec(t.d_enc_4c)
    assert d == (0 if r < (N - 1) // 2 else dec(t.d_enc_4a))                                # step 4c
    beta_bits = to_bits(t.beta, L)
    assert [dec(v) for v in t.xor_enc] == [a ^ b for a, b in zip(t.alpha, beta_bits)]       # step 4d
    assert t.alpha_tilde == to_bits((r - N) % (1 << L), L)                                  # step 4e
    for i in range(L):
        expect = (t.alpha[i] ^ beta_bits[i]) - (0 if t.alpha[i] == t.alpha_tilde[i] else d)
        assert dec(t.w_enc_4e[i]) == expect % u
        assert dec(t.w_enc_4f[i]) == dec(t.w_enc_4e[i]) * 2 ** i % u                        # step 4f
    s = 1 - 2 * nc.delta_a
    w_plain = [dec(w) for w in t.w_enc_4f]
    for i in range(L):                                                                      # step 4h
        temp = s + t.alpha[i] + (t.alpha_tilde[i] - t.alpha[i]) * d - beta_bits[i] + 3 * sum(w_plain[i + 1:])
        assert temp % u == dec(t.c_enc_4h[i + 1]) % u
    assert dec(t.c_enc_4h[0]) == (nc.delta_a + sum(w_plain)) % u
    assert t.delta_b == int(any(dec(c) == 0 for c in t.c_enc_4i))                           # step 4i/4j
    assert sorted(t.c_enc_4i) == sorted(pow(c, e, dgk.public_key.n) for c, e in zip(t.c_enc_4h, nc.r_blind))
    assert oracle.paillier_decrypt_raw(t.zeta1_enc) == t.z // 2 ** L                        # step 5
    assert oracle.paillier_decrypt_raw(t.delta_b_enc) == t.delta_b
    beta_lt_alpha = int(t.beta < from_bits(t.alpha))
    assert oracle.paillier_decrypt_raw(t.beta_lt_alpha_enc) == beta_lt_alpha                # step 6
    assert t.result_bit == 1                                                                # x < y
    assert oracle.compare(y_enc, x_enc, L, nc, interactive=False).result_bit == 0           # test_greater_than
    assert oracle.compare(x_enc, x_enc, L, nc, interactive=False).result_bit == 1           # test_equal_to


def test_step_1_rejects_large_l(schemes):  # test_secure_comparison.py:240-251
    paillier, _, oracle = schemes
    with pytest.raises(AssertionError):
        oracle.step_1(oracle.paillier_encrypt_raw(1), oracle.paillier_encrypt_raw(2), paillier.public_key.n.bit_length() - 2, 5)


@pytest.mark.parametrize("z_kind", ["zero", "n", "half"])
def test_step_4a(schemes, z_kind):  # test_secure_comparison.py:279-291, TEST_4A
    paillier, dgk, oracle = schemes
    N = paillier.public_key.n
    z = {"zero": 0, "n": N, "half": (N - 1) // 2}[z_kind]
    assert dgk.decrypt(type(dgk).ciphertext_cls(oracle.step_4a(z, L), dgk)) == int(z < (N - 1) // 2)


def test_step_4c(schemes):  # test_secure_comparison.py:324-386
    paillier, dgk, oracle = schemes
    N = paillier.public_key.n
    one = oracle.dgk_encrypt_raw(1)
    assert oracle.step_4c(one, 0) == 1 and oracle.step_4c(one, 100) == 1
    assert oracle.step_4c(one, N - 100) == one
    for bad in (-10, N + 10):
        with pytest.raises(AssertionError):
            oracle.step_4c(one, bad)


@pytest.mark.parametrize("z_is_n,delta_b", [(True, 0), (False, 0), (True, 1), (False, 1)])
def test_step_5(schemes, z_is_n, delta_b):  # test_secure_comparison.py:547-565, TEST_5
    paillier, _, oracle = schemes
    N = paillier.public_key.n
    z = N if z_is_n else 100
    z1, z2, db = oracle.step_5(z, L, delta_b)
    assert oracle.paillier_decrypt_raw(z1) == (z // 2 ** L) % N
    assert oracle.paillier_decrypt_raw(z2) == (int(z < (N - 1) // 2) * ((z + N) // 2 ** L) + int(z >= (N - 1) // 2) * (z // 2 ** L)) % N
    assert oracle.paillier_decrypt_raw(db) == delta_b


@pytest.mark.parametrize("delta_a,delta_b", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_step_6(schemes, delta_a, delta_b):  # test_secure_comparison.py:568-585, TEST_6
    _, _, oracle = schemes
    out = oracle.step_6(delta_a, oracle.paillier_encrypt_raw(delta_b))
    assert oracle.paillier_decrypt_raw(out) == (delta_b if delta_a else 1 - delta_b)


def test_next_prime_matches_reference_choice_of_u():
    assert next_prime(1 << 18) == 262147 and next_prime(1 << 10) == 1031
