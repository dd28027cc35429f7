"""GPU parity tests of the generic batched primitives (through the C ABI) against Python integers
(CPython pow == GMP mpz_powm/mpz_invert results; see tests/test_oracle_golden.py::test_backends_agree)."""
import random

import numpy as np
import pytest

from protocols.secure_comparison_b200 import Engine, NotInvertibleError, ints_to_limbs, limbs_to_ints
from protocols.secure_comparison_b200 import _native as nat

pytestmark = pytest.mark.gpu

SLOT = nat.MOD_USER0


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def rand_modulus(rng, bits):
    return rng.getrandbits(bits) | 1 | (1 << (bits - 1))


# (modulus bits, exponent bits, batch) -- every (L, TPI) kernel configuration is covered:
# K=36 (<=1042 bits), 54, 72, 108, 144, 216, 288
SIZES = [(64, 64, 40), (128, 160, 70), (1024, 160, 300), (1024, 1024, 200), (1536, 160, 150), (2048, 400, 300), (2048, 2048, 130),
         (3072, 400, 100), (4096, 2048, 70), (6144, 64, 40), (8192, 96, 40), (1000, 70, 64), (2040, 333, 64), (4090, 100, 33)]


@pytest.mark.parametrize("engine", ["fp64", "fp64-wide-lanes", "imad"])
@pytest.mark.parametrize("bits,ebits,B", SIZES)
def test_modexp_matches_pow(eng, bits, ebits, B, engine):
    """Both multiplier engines (FP64-pipe 52-bit digits, IMAD 29-bit digits) against Python's pow.  These batches are small, so by
    default the FP64 engine spreads every number over twice the lanes (5 digits per lane); "fp64" switches that off and
    runs the throughput geometry (10 digits per lane) that full-size batches use."""
    eng.set_modexp_fp64(engine != "imad")
    eng.set_small_batch(engine == "fp64-wide-lanes")
    try:
        _modexp_matches_pow(eng, bits, ebits, B)
    finally:
        eng.set_modexp_fp64(True)
        eng.set_small_batch(True)


def _modexp_matches_pow(eng, bits, ebits, B):
    rng = random.Random(bits * 7 + ebits)
    n = rand_modulus(rng, bits)
    k = eng.load_modulus(SLOT, n)
    bases = [rng.randrange(n) for _ in range(B)]
    exps = [rng.getrandbits(ebits) for _ in range(B)]
    # edge cases: base 0, 1, n-1; exponent 0, 1, all ones
    bases[:3] = [0, 1, n - 1]
    exps[3:6] = [0, 1, (1 << ebits) - 1]
    ew = (ebits + 31) // 32
    out = eng.modexp(SLOT, ints_to_limbs(bases, k), ints_to_limbs(exps, ew), ebits)
    assert limbs_to_ints(out) == [pow(b, e, n) for b, e in zip(bases, exps)]
    # shared exponent + fused multiplication (the "randomize" shape: c * rho^N)
    e_shared = rng.getrandbits(ebits) | (1 << (ebits - 1))
    muls = [rng.randrange(n) for _ in range(B)]
    out = eng.modexp(SLOT, ints_to_limbs(bases, k), ints_to_limbs([e_shared], ew)[0], ebits, ints_to_limbs(muls, k))
    assert limbs_to_ints(out) == [m * pow(b, e_shared, n) % n for b, m in zip(bases, muls)]


# (root bits, exponent bits, batch): every pair-form geometry -- root K = 20 (10x2), 30 (15x2), 40 (10x4), 60 (15x4), 80 (10x8);
# 1036 and 2076 bits are the sizes with the least Montgomery slack (R = 16 n)
SQUARE_SIZES = [(512, 512, 40), (1024, 1024, 150), (1036, 200, 64), (1536, 160, 64), (2048, 2048, 100), (2076, 96, 40), (3072, 128, 40),
                (4096, 300, 40), (100, 64, 33)]


@pytest.mark.parametrize("wide_lanes", [False, True], ids=["10-digits-per-lane", "5-digits-per-lane"])
@pytest.mark.parametrize("bits,ebits,B", SQUARE_SIZES)
def test_modexp_square_modulus_pair_form(eng, bits, ebits, B, wide_lanes):
    """Exponentiation modulo a linked square n^2 (base-n pair arithmetic, tests/pairmul_model.py) against Python's pow and
    against the plain engine on the same slot; bases wider than n (reduced on load), per-element and shared exponents."""
    eng.set_small_batch(wide_lanes)
    rng = random.Random(bits * 13 + ebits)
    n = rand_modulus(rng, bits)
    root, sq = nat.MOD_USER0 + 1, nat.MOD_USER0 + 2
    eng.load_modulus(root, n)
    k2 = eng.load_modulus(sq, n * n, 2 * ((bits + 31) // 32))
    eng.link_square(sq, root)
    bases = [rng.randrange(n * n) for _ in range(B)]
    bases[:6] = [0, 1, n * n - 1, n, n - 1, (1 << (32 * k2)) - 1 if (1 << (32 * k2)) - 1 < n * n else n + 1]
    exps = [rng.getrandbits(ebits) for _ in range(B)]
    exps[6:9] = [0, 1, (1 << ebits) - 1]
    ew = (ebits + 31) // 32
    want = [pow(b, e, n * n) for b, e in zip(bases, exps)]
    out = eng.modexp(sq, ints_to_limbs(bases, k2), ints_to_limbs(exps, ew), ebits)
    assert limbs_to_ints(out) == want
    e_shared = rng.getrandbits(ebits) | (1 << (ebits - 1)) | 1
    want_shared = [pow(b, e_shared, n * n) for b in bases]
    out = eng.modexp(sq, ints_to_limbs(bases, k2), ints_to_limbs([e_shared], ew)[0], ebits)
    assert limbs_to_ints(out) == want_shared
    eng.set_sliding_windows(False)
    try:
        out = eng.modexp(sq, ints_to_limbs(bases, k2), ints_to_limbs([e_shared], ew)[0], ebits)
        assert limbs_to_ints(out) == want_shared
    finally:
        eng.set_sliding_windows(True)
    eng.set_square_modulus(False)                    # the same slot as an ordinary 2K-digit modulus
    try:
        out = eng.modexp(sq, ints_to_limbs(bases, k2), ints_to_limbs([e_shared], ew)[0], ebits)
        assert limbs_to_ints(out) == want_shared
    finally:
        eng.set_square_modulus(True)
    with pytest.raises(ValueError):
        eng.link_square(root, sq)                    # not a square of that slot
    eng.load_modulus(root, n)                        # reloading the root drops the link (results stay correct)
    out = eng.modexp(sq, ints_to_limbs(bases[:8], k2), ints_to_limbs([e_shared], ew)[0], ebits)
    assert limbs_to_ints(out) == want_shared[:8]
    eng.set_small_batch(True)


@pytest.mark.parametrize("bits", [64, 1024, 2048, 3072, 4096, 8192])
def test_modmul_and_edge_moduli(eng, bits):
    rng = random.Random(bits)
    for n in (rand_modulus(rng, bits), (1 << bits) - 1, (1 << (bits - 1)) + 1):
        k = eng.load_modulus(SLOT, n)
        a = [rng.randrange(n) for _ in range(50)] + [0, n - 1, n - 1, 1]
        b = [rng.randrange(n) for _ in range(50)] + [5, n - 1, 1, 1]
        out = eng.modmul(SLOT, ints_to_limbs(a, k), ints_to_limbs(b, k))
        assert limbs_to_ints(out) == [x * y % n for x, y in zip(a, b)]


def _inv_or_none(v, n):
    try:
        return pow(v, -1, n)
    except ValueError:
        return None


@pytest.mark.parametrize("bits", [128, 2048, 4096])
def test_modinv(eng, bits):
    rng = random.Random(bits + 1)
    p, q = rand_modulus(rng, bits // 2), rand_modulus(rng, bits - bits // 2)
    n = p * q                                    # p, q are random odd numbers: n has small factors too
    k = eng.load_modulus(SLOT, n)
    vals = [rng.randrange(1, n) for _ in range(60)] + [1, n - 1, 2, 0, p, (q * 3) % n]
    expect = [_inv_or_none(v, n) for v in vals]
    assert any(e is None for e in expect) and sum(e is not None for e in expect) > 10
    out, ok = eng.modinv(SLOT, ints_to_limbs(vals, k), raise_on_failure=False)
    assert ok.tolist() == [int(e is not None) for e in expect]
    assert limbs_to_ints(out) == [e or 0 for e in expect]
    with pytest.raises(NotInvertibleError):
        eng.modinv(SLOT, ints_to_limbs(vals, k))
    good = [v for v, e in zip(vals, expect) if e is not None]
    assert limbs_to_ints(eng.modinv(SLOT, ints_to_limbs(good, k))) == [pow(v, -1, n) for v in good]


@pytest.mark.parametrize("wide", [True, False], ids=["w16", "w8"])
@pytest.mark.parametrize("bits,ebits", [(128, 50), (2048, 400), (3072, 400), (1024, 31)])
def test_fixed_base(eng, bits, ebits, wide):
    """8-bit-window tables and the 16-bit-window tables built from them give the same (exact) powers."""
    rng = random.Random(bits + ebits)
    n = rand_modulus(rng, bits)
    k = eng.load_modulus(SLOT, n)
    h = rng.randrange(2, n)
    eng.set_fixed_base_wide(wide)
    try:
        eng.fixed_base_load(0, SLOT, h, ebits)
    finally:
        eng.set_fixed_base_wide(True)
    B = 200
    exps = [rng.getrandbits(ebits) for _ in range(B)]
    exps[:3] = [0, 1, (1 << ebits) - 1]
    muls = [rng.randrange(n) for _ in range(B)]
    ew = (ebits + 31) // 32
    out = eng.fixed_base_exp(0, ints_to_limbs(exps, ew))
    assert limbs_to_ints(out) == [pow(h, e, n) for e in exps]
    out = eng.fixed_base_exp(0, ints_to_limbs(exps, ew), ints_to_limbs(muls, k))
    assert limbs_to_ints(out) == [m * pow(h, e, n) % n for e, m in zip(exps, muls)]


def test_argument_errors(eng):
    with pytest.raises(ValueError):
        eng.load_modulus(SLOT, 1 << 64)           # even modulus
    k = eng.load_modulus(SLOT, (1 << 127) - 1)
    with pytest.raises(ValueError):
        eng.modmul(SLOT, np.zeros((3, k + 1), np.uint32), np.zeros((3, k + 1), np.uint32))
    assert eng.modmul(SLOT, np.zeros((0, k), np.uint32), np.zeros((0, k), np.uint32)).shape == (0, k)


@pytest.mark.parametrize("L,TPI", [(10, 4), (10, 2), (10, 8), (10, 16), (5, 4), (5, 8), (5, 16), (15, 2), (15, 4), (15, 8)],
                         ids=lambda v: str(v))
def test_fp64_engine_carry_resolution_rare_paths(eng, L, TPI):
    """resolve_carries of the FP64 engine on crafted columns, in every geometry the engine instantiates: all-ones digits
    with a carry arriving from below (the ripple loop behind the common path: about 2^-46 per lane and product on random
    data, i.e. it DOES run in production, a few times per day and GPU), big lane carries, carries that stop inside a lane,
    at a lane boundary, or run through every lane, mixed random cases."""
    K = L * TPI
    M52 = (1 << 52) - 1
    rng = random.Random(52 + 100 * L + TPI)
    cases = []
    ones = [M52] * K
    zero_hi = [0] * TPI
    cases.append(([1] + [0] * (K - 1), zero_hi))                                     # trivial
    cases.append(([M52 + 1] + ones[1:K - 1] + [0], zero_hi))                          # +1 ripples through every lane to the top digit
    if TPI >= 3:
        cases.append(([M52 + 5] + ones[1:2 * L] + [7] + [0] * (K - 2 * L - 1), zero_hi))   # ripple across two lane boundaries, stops
    cases.append(([(1 << 58) + 3] * (K - 2) + [0, 0], zero_hi))                       # large column values everywhere
    cases.append(([M52 + 1] + ones[1:K - 1] + [0], zero_hi))
    cases.append(([M52] * (K - 1) + [0], [1] + [0] * (TPI - 1)))                      # carry enters through the lane's high accumulator
    cases.append(([M52] * (K - 1) + [0], [(1 << 57) + 1] + [1] * (TPI - 2) + [0] if TPI >= 2 else [0]))
    for stop in range(1, min(K - 1, 2 * L + 2)):                                       # a carry out of digit 0 that dies at digit `stop`
        c = [M52 + 1] + [M52] * (stop - 1) + [rng.getrandbits(51)] + [rng.getrandbits(52) for _ in range(K - stop - 2)] + [0]
        cases.append((c[:K], zero_hi))
    for lane in range(1, TPI):                                                         # digits 0 / 1 of a lane take the neighbour's big carry-out
        c = [rng.getrandbits(52) for _ in range(K)]
        c[lane * L - 1] = (1 << 62) + M52
        c[lane * L] = M52
        c[lane * L + 1] = M52
        c[K - 1] = 0
        if lane * L + 1 == K - 1:
            c[lane * L] = 0
        cases.append((c, zero_hi))
    for _ in range(60):
        cols = [rng.choice([M52, M52, M52 + 1, rng.getrandbits(60), 0, 1]) for _ in range(K - 1)] + [0]
        hi = [rng.choice([0, 1, rng.getrandbits(56)]) for _ in range(TPI - 1)] + [0]
        cases.append((cols, hi))
    kept = []
    for c, h in cases:
        value = sum(int(v) << (52 * i) for i, v in enumerate(c)) + sum(int(v) << (52 * L * (t + 1)) for t, v in enumerate(h))
        if value < 1 << (52 * K):                                                      # must fit K digits
            kept.append((c, h, value))
    assert len(kept) >= 40
    G = len(kept)
    cols = np.array([c for c, _, _ in kept], dtype=np.uint64)
    hi = np.array([h for _, h, _ in kept], dtype=np.uint64)
    out = np.zeros((G, K), dtype=np.uint64)
    eng._check(eng.lib.scb_debug_resolve52_geo(eng.ctx, L, TPI, nat._ptr(cols), nat._ptr(hi), nat._ptr(out), G))
    for (c, h, value), digits in zip(kept, out.tolist()):
        assert all(d <= M52 for d in digits)
        assert sum(d << (52 * i) for i, d in enumerate(digits)) == value
    if (L, TPI) == (10, 4):                                                            # the fixed-geometry entry point is the same code
        out2 = np.zeros((G, K), dtype=np.uint64)
        eng._check(eng.lib.scb_debug_resolve52(eng.ctx, nat._ptr(cols), nat._ptr(hi), nat._ptr(out2), G))
        assert np.array_equal(out, out2)


@pytest.mark.parametrize("fp64", [True, False], ids=["fp64", "imad"])
@pytest.mark.parametrize("sliding", [True, False], ids=["sliding", "fixed"])
def test_shared_exponent_programs_on_crafted_exponents(eng, fp64, sliding):
    """The host-pointer modexp compiles a SHARED exponent into a sliding-window program: exponents with zero runs longer
    than one program entry can hold (255 squarings), a single set bit, all ones, trailing zeros, and short exponents that
    stay on the fixed-window loop -- all against Python's pow, on both engines."""
    rng = random.Random(77)
    bits = 2048
    n = rand_modulus(rng, bits)
    k = eng.load_modulus(SLOT, n)
    B = 24
    bases = [rng.randrange(n) for _ in range(B)]
    bases[:3] = [0, 1, n - 1]
    exps = [(1 << 1500) + 1, 1 << 2047, (1 << 2048) - 1, ((rng.getrandbits(700) | 1) << 1200) | rng.getrandbits(500) | (1 << 499),
            rng.getrandbits(2048) << 0 | (1 << 2047), (rng.getrandbits(1000) | (1 << 999)) << 600, (1 << 63) | 5, 12345, 1 << 64, 3]
    eng.set_modexp_fp64(fp64)
    eng.set_sliding_windows(sliding)
    try:
        for e in exps:
            ew = max(1, (e.bit_length() + 31) // 32)
            out = eng.modexp(SLOT, ints_to_limbs(bases, k), ints_to_limbs([e], ew)[0], e.bit_length())
            assert limbs_to_ints(out) == [pow(b, e, n) for b in bases], f"exponent of {e.bit_length()} bits"
    finally:
        eng.set_modexp_fp64(True)
        eng.set_sliding_windows(True)
