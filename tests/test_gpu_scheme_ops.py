"""GPU parity tests of the scheme-level operations (Paillier encrypt / randomize / decrypt, DGK bit
encryption / is_zero, the phase building blocks) against Python integers, one operation at a time."""
import random

import numpy as np
import pytest

from oracle.protocol import Oracle
from oracle.workload import make_schemes
from protocols.secure_comparison_b200 import get_engine, ints_to_limbs, limbs_to_ints, reset_engines
from protocols.secure_comparison_b200 import _native as nat

pytestmark = pytest.mark.gpu

KEYS = [(1024, 128, 20, 16), (512, 192, 24, 8), (2048, 2048, 160, 16)]


@pytest.fixture(scope="module", params=KEYS, ids=lambda k: f"p{k[0]}d{k[1]}l{k[3]}")
def setup(request):
    pb, db, vb, l = request.param
    paillier, dgk = make_schemes(pb, db, vb, l)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    return paillier, dgk, l, eng, Oracle(paillier, dgk, "auto"), random.Random(pb + l)


def test_paillier_encrypt_and_randomize(setup):
    paillier, dgk, l, eng, oracle, rng = setup
    N, N2 = paillier.public_key.n, paillier.public_key.n_squared
    B = 37
    m = [rng.randrange(N) for _ in range(B)]
    m[:4] = [0, 1, N - 1, 2 ** l]
    rho = [rng.randrange(1, N) for _ in range(B)]
    out = eng.paillier_encrypt(ints_to_limbs(m, eng.kN))
    assert limbs_to_ints(out) == [(1 + v * N) % N2 for v in m]
    out = eng.paillier_encrypt(ints_to_limbs(m, eng.kN), ints_to_limbs(rho, eng.kN))
    assert limbs_to_ints(out) == [(1 + v * N) * pow(r, N, N2) % N2 for v, r in zip(m, rho)]


def test_paillier_decrypt(setup):
    paillier, dgk, l, eng, oracle, rng = setup
    N, N2 = paillier.public_key.n, paillier.public_key.n_squared
    m = [rng.randrange(N) for _ in range(20)] + [0, 1, N - 1]
    c = [(1 + v * N) * pow(rng.randrange(1, N), N, N2) % N2 for v in m]
    c += [1, N2 - 2, rng.randrange(1, N2)]          # arbitrary group elements decrypt like the oracle says
    want = [oracle.paillier_decrypt_raw(v) for v in c]
    assert want[:len(m)] == m
    assert limbs_to_ints(eng.paillier_decrypt(ints_to_limbs(c, eng.kc))) == want


def test_paillier_crt_and_public_formulas_agree(setup):
    """KeyHolder-side CRT (p, p^2, q, q^2) vs the public formulas c^lambda mod N^2 / rho^N mod N^2: identical bits."""
    paillier, dgk, l, eng, oracle, rng = setup
    N, N2 = paillier.public_key.n, paillier.public_key.n_squared
    assert eng.has_paillier_crt
    B = 33
    m = [rng.randrange(N) for _ in range(B)]
    m[:3] = [0, 1, N - 1]
    rho = [rng.randrange(1, N) for _ in range(B)]
    rho[:3] = [1, N - 1, paillier.secret_key.p]             # a multiple of p: the lift (rho mod p)^... = 0 mod p^2
    want = [(1 + v * N) * pow(r, N, N2) % N2 for v, r in zip(m, rho)]
    c = [rng.randrange(1, N2) for _ in range(B)]
    want_dec = [oracle.paillier_decrypt_raw(v) for v in c]
    try:
        for crt in (True, False):
            for fp64, sliding in ((True, True), (True, False), (False, True)):
                eng.set_paillier_crt(crt)
                eng.set_modexp_fp64(fp64)
                eng.set_sliding_windows(sliding)
                tag = f"crt={crt}, fp64={fp64}, sliding={sliding}"
                got = eng.paillier_encrypt(ints_to_limbs(m, eng.kN), ints_to_limbs(rho, eng.kN))
                assert limbs_to_ints(got) == want, f"encrypt+randomize, {tag}"
                assert limbs_to_ints(eng.paillier_decrypt(ints_to_limbs(c, eng.kc))) == want_dec, f"decrypt, {tag}"
    finally:
        eng.set_paillier_crt(True)
        eng.set_modexp_fp64(True)
        eng.set_sliding_windows(True)


def test_crt_halves_on_two_streams_are_bit_identical(setup):
    """SCB_OPT_OVERLAP_CRT: the q half of the CRT forms runs on the context's second stream.  Back-to-back calls (buffers
    reused across the fork / join) give the same bits as the single-stream order and as the closed forms."""
    paillier, dgk, l, eng, oracle, rng = setup
    N, N2 = paillier.public_key.n, paillier.public_key.n_squared
    B = 1500                                          # several warps per launch on both streams
    m = [rng.randrange(N) for _ in range(B)]
    rho = [rng.randrange(1, N) for _ in range(B)]
    M, RHO = ints_to_limbs(m, eng.kN), ints_to_limbs(rho, eng.kN)
    outs = {}
    try:
        for overlap in (True, False, True):
            eng.set_overlap_crt(overlap)
            c1 = eng.paillier_encrypt(M, RHO)
            d1 = eng.paillier_decrypt(c1)             # consumes c1 right after the join
            c2 = eng.paillier_encrypt(d1, RHO)        # same plaintexts again
            outs.setdefault(overlap, []).append((c1.copy(), d1.copy(), c2.copy()))
    finally:
        eng.set_overlap_crt(True)
    ref = outs[False][0]
    for trip in outs[True]:
        for got, want in zip(trip, ref):
            assert np.array_equal(got, want)
    assert np.array_equal(ref[0], ref[2])
    assert limbs_to_ints(ref[1]) == m
    idx = [0, 1, B // 2, B - 1]
    assert [limbs_to_ints(ref[0])[i] for i in idx] == [(1 + m[i] * N) * pow(rho[i], N, N2) % N2 for i in idx]


def test_step1_pieces(setup):
    paillier, dgk, l, eng, oracle, rng = setup
    N, N2 = paillier.public_key.n, paillier.public_key.n_squared
    B = 45                                           # > one inversion chunk, ragged second chunk
    x = [rng.randrange(1, N2) for _ in range(B)]
    y = [rng.randrange(1, N2) for _ in range(B)]
    r = [rng.randrange(N) for _ in range(B)]
    r[:3] = [0, N - 1, (N - 1) // 2]
    rho = [rng.randrange(1, N) for _ in range(B)]
    X, Y, R, RHO = (ints_to_limbs(v, k) for v, k in ((x, eng.kc), (y, eng.kc), (r, eng.kN), (rho, eng.kN)))
    got = limbs_to_ints(eng.initiator_step1(X, Y, R, None, l))
    want = [oracle.step_1(a, b, l, c) for a, b, c in zip(x, y, r)]
    bad = [i for i in range(B) if got[i] != want[i]]
    assert not bad, f"step_1 without randomize differs at {bad[:10]}"
    want = [oracle.paillier_randomize(oracle.step_1(a, b, l, c), d) for a, b, c, d in zip(x, y, r, rho)]
    try:
        for overlap in (True, False, True):          # rho^N on the second stream next to the inversion of x, or one stream
            eng.set_overlap_crt(overlap)
            got = limbs_to_ints(eng.initiator_step1(X, Y, R, RHO, l))
            bad = [i for i in range(B) if got[i] != want[i]]
            assert not bad, f"step_1 + randomize (two streams: {overlap}) differs at {bad[:10]}"
    finally:
        eng.set_overlap_crt(True)


def test_dgk_bit_encryption(setup):
    paillier, dgk, l, eng, oracle, rng = setup
    N = paillier.public_key.n
    B = 23
    z = [rng.randrange(N) for _ in range(B)]
    z[:3] = [0, (N - 1) // 2, (N - 1) // 2 - 1]
    bits = dgk.randomness_bits
    rho = [[rng.getrandbits(bits) for _ in range(l + 1)] for _ in range(B)]
    out = eng.keyholder_step4ab(ints_to_limbs(z, eng.kN), l, None)
    want = [[oracle.step_4a(v, l)] + oracle.step_4b(v % (1 << l), l) for v in z]
    assert limbs_to_ints(out.reshape(B * (l + 1), -1)) == [c for row in want for c in row]
    RHO = ints_to_limbs([e for row in rho for e in row], eng.rand_words).reshape(B, l + 1, eng.rand_words)
    out = eng.keyholder_step4ab(ints_to_limbs(z, eng.kN), l, RHO)
    want_r = [oracle.dgk_randomize(c, e) for row, erow in zip(want, rho) for c, e in zip(row, erow)]
    assert limbs_to_ints(out.reshape(B * (l + 1), -1)) == want_r


def test_dgk_is_zero(setup):
    paillier, dgk, l, eng, oracle, rng = setup
    n, u = dgk.public_key.n, dgk.public_key.u
    ms = [0, 1, u - 1, 0, 5, 0] + [rng.randrange(u) for _ in range(30)]
    c = [pow(dgk.public_key.g, m, n) * pow(dgk.public_key.h, rng.getrandbits(dgk.randomness_bits), n) % n for m in ms]
    c += [1, n - 1, rng.randrange(1, n)]
    flags = eng.dgk_is_zero(ints_to_limbs(c, eng.kn))
    assert flags.tolist() == [int(oracle.dgk_is_zero(v)) for v in c]
    assert flags[:6].tolist() == [1, 0, 0, 1, 0, 1]


def test_step5_and_step67(setup):
    paillier, dgk, l, eng, oracle, rng = setup
    N, N2 = paillier.public_key.n, paillier.public_key.n_squared
    B = 40
    z = [rng.randrange(N) for _ in range(B)]
    z[:4] = [0, 100, (N - 1) // 2, N - 1]
    db = [rng.randrange(2) for _ in range(B)]
    rho = [[rng.randrange(1, N) for _ in range(3)] for _ in range(B)]
    r = [rng.randrange(N) for _ in range(B)]
    r[:4] = [0, N - 1, (1 << l) - 1, (N - 1) // 2]
    da = [rng.randrange(2) for _ in range(B)]
    enc3 = eng.keyholder_step5(ints_to_limbs(z, eng.kN), np.array(db, dtype=np.uint8), None, l)
    want = [oracle.step_5(a, l, b) for a, b in zip(z, db)]
    assert limbs_to_ints(enc3.reshape(3 * B, -1)) == [v for t in want for v in t]
    RHO = ints_to_limbs([e for row in rho for e in row], eng.kN).reshape(B, 3, eng.kN)
    enc3r = eng.keyholder_step5(ints_to_limbs(z, eng.kN), np.array(db, dtype=np.uint8), RHO, l)
    want_r = [[oracle.paillier_randomize(c, e) for c, e in zip(t, er)] for t, er in zip(want, rho)]
    assert limbs_to_ints(enc3r.reshape(3 * B, -1)) == [v for t in want_r for v in t]
    res = eng.initiator_step6_7(enc3r, ints_to_limbs(r, eng.kN), np.array(da, dtype=np.uint8), l)
    want_res = [oracle.step_7(t[0], t[1], rr, l, oracle.step_6(a, t[2])) for t, rr, a in zip(want_r, r, da)]
    bad = [i for i in range(B) if limbs_to_ints(res)[i] != want_res[i]]
    assert not bad, f"steps 6+7 differ at {bad[:10]} (delta_a {[da[i] for i in bad[:10]]})"


def test_step4_without_blinding_and_with(setup):
    paillier, dgk, l, eng, oracle, rng = setup
    N, n, u = paillier.public_key.n, dgk.public_key.n, dgk.public_key.u
    B = 19
    cases = []
    for i in range(B):
        z = rng.randrange(N)
        r = rng.randrange(N) if i % 3 else rng.randrange((N - 1) // 2)
        d = oracle.dgk_randomize(oracle.step_4a(z, l), rng.getrandbits(64))
        beta = [oracle.dgk_randomize(b, rng.getrandbits(64)) for b in oracle.step_4b(z % (1 << l), l)]
        cases.append((r, d, beta, rng.randrange(2)))
    DB = ints_to_limbs([v for r, d, beta, da in cases for v in [d] + beta], eng.kn).reshape(B, l + 1, eng.kn)
    R = ints_to_limbs([c[0] for c in cases], eng.kN)
    DA = np.array([c[3] for c in cases], dtype=np.uint8)

    def oracle_4h(r, d, beta, da):
        alpha = oracle.step_3(r, l)
        d4c = oracle.step_4c(d, r)
        xor = oracle.step_4d(alpha, beta)
        w, at = oracle.step_4e(r, alpha, xor, d4c)
        w = oracle.step_4f(w)
        return oracle.step_4h(1 - 2 * da, alpha, at, d4c, beta, w, da)

    want = [oracle_4h(*c) for c in cases]
    got = eng.initiator_step4(DB, R, DA, None, None, None, l)
    assert limbs_to_ints(got.reshape(B * (l + 1), -1)) == [v for row in want for v in row], "steps 4c-4h"
    rb = [[rng.randrange(1, u) for _ in range(l + 1)] for _ in range(B)]
    perm = [rng.sample(range(l + 1), l + 1) for _ in range(B)]
    RB = ints_to_limbs([v for row in rb for v in row], eng.rb_words).reshape(B, l + 1, eng.rb_words)
    got = eng.initiator_step4(DB, R, DA, RB, np.array(perm, dtype=np.uint8), None, l)
    want_i = [oracle.step_4i(row, e, p) for row, e, p in zip(want, rb, perm)]
    assert limbs_to_ints(got.reshape(B * (l + 1), -1)) == [v for row in want_i for v in row], "step 4i"


def test_empty_and_tiny_batches(setup):
    """B = 0 is a no-op for every phase (shapes preserved); B = 1, 2, 3 agree with the oracle (ragged warps)."""
    paillier, dgk, l, eng, oracle, rng = setup
    kN, kc, kn = eng.kN, eng.kc, eng.kn
    z0 = np.zeros((0, kN), np.uint32)
    assert eng.paillier_encrypt(z0).shape == (0, kc)
    assert eng.paillier_decrypt(np.zeros((0, kc), np.uint32)).shape == (0, kN)
    assert eng.initiator_step1(np.zeros((0, kc), np.uint32), np.zeros((0, kc), np.uint32), z0, None, l).shape == (0, kc)
    assert eng.keyholder_step4ab(z0, l, None).shape == (0, l + 1, kn)
    assert eng.initiator_step4(np.zeros((0, l + 1, kn), np.uint32), z0, np.zeros(0, np.uint8), None, None, None, l).shape == (0, l + 1, kn)
    db, enc3 = eng.keyholder_step4j_5(np.zeros((0, l + 1, kn), np.uint32), z0, None, l)
    assert db.shape == (0,) and enc3.shape == (0, 3, kc)
    assert eng.initiator_step6_7(np.zeros((0, 3, kc), np.uint32), z0, np.zeros(0, np.uint8), l).shape == (0, kc)
    assert eng.dgk_is_zero(np.zeros((0, kn), np.uint32)).shape == (0,)
    N, N2 = paillier.public_key.n, paillier.public_key.n_squared
    for B in (1, 2, 3):
        m = [rng.randrange(N) for _ in range(B)]
        rho = [rng.randrange(1, N) for _ in range(B)]
        got = eng.paillier_encrypt(ints_to_limbs(m, kN), ints_to_limbs(rho, kN))
        assert limbs_to_ints(got) == [(1 + v * N) * pow(r, N, N2) % N2 for v, r in zip(m, rho)]
        assert limbs_to_ints(eng.paillier_decrypt(got)) == m


def test_plaintexts_beyond_n_reduce_like_the_reference(setup):
    """test_modulo_n_squared (test_secure_comparison.py:166-182): unsafe_encrypt of a plaintext >= N is (1 + m N) mod N^2,
    and step_1 reduces mod N^2 whatever the magnitude of its inputs (y_enc = N^2 - 2, x_enc = enc(0))."""
    paillier, dgk, l, eng, oracle, rng = setup
    N, N2 = paillier.public_key.n, paillier.public_key.n_squared
    top = (1 << (32 * eng.kN)) - 1                                   # largest value the [.., kN] plaintext layout can carry
    m = [N, N + 5, top, (1 << l) + N - 1]
    assert limbs_to_ints(eng.paillier_encrypt(ints_to_limbs(m, eng.kN))) == [(1 + v * N) % N2 for v in m]
    x_enc, y_enc, r = 1, N2 - 2, N - 1                               # enc(0) = 1
    got = limbs_to_ints(eng.initiator_step1(ints_to_limbs([x_enc], eng.kc), ints_to_limbs([y_enc], eng.kc), ints_to_limbs([r], eng.kN), None, l))
    assert got == [oracle.step_1(x_enc, y_enc, l, r)]
    y = oracle.paillier_decrypt_raw(y_enc)
    assert limbs_to_ints(eng.paillier_decrypt(ints_to_limbs(got, eng.kc))) == [(y - 0 + (1 << l) + r) % N]
