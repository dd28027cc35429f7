"""GPU parity tests of the protocol path (through the C ABI) against
 * the committed golden vectors minted from the UNMODIFIED reference (tests/golden/*.json), and
 * the CPU oracle (oracle/protocol.py) on fresh seeded batches,
bit for bit on every ciphertext that crosses the party boundary, plus the decrypted comparison bits.
Both flows are covered: interactive (every randomize() of perform_secure_comparison, injected nonces)
and non-interactive (the step-by-step flow of README.md:117-141), fused phases and individual steps."""
import random

import numpy as np
import pytest

from helpers import GOLDEN_FILES, load_golden, pack_case_arrays
from oracle.protocol import Oracle
from oracle.workload import draw_inputs, draw_nonces, encrypt_input, make_schemes
from protocols.secure_comparison_b200 import Initiator, KeyHolder, get_engine, ints_to_limbs, limbs_to_ints, reset_engines

pytestmark = pytest.mark.gpu


def run_fused(eng, paillier, dgk, l, arr, interactive):
    """The five party-phases on the GPU; returns every message that crosses the party boundary."""
    rho = (lambda k: arr[k]) if interactive else (lambda k: None)
    out = {}
    out["z_enc"] = eng.initiator_step1(arr["x"], arr["y"], arr["r"], rho("rho_z"), l)                       # A1
    out["z"] = eng.paillier_decrypt(out["z_enc"])                                                             # B1: step 2
    out["db_enc"] = eng.keyholder_step4ab(out["z"], l, rho("rho_db"))                                         #     steps 4a, 4b
    out["c_enc"] = eng.initiator_step4(out["db_enc"], arr["r"], arr["delta_a"], arr["r_blind"], arr["perm"], rho("rho_c"), l)   # A2
    out["delta_b"], out["enc3"] = eng.keyholder_step4j_5(out["c_enc"], out["z"], rho("rho_enc3"), l)          # B2
    out["result"] = eng.initiator_step6_7(out["enc3"], arr["r"], arr["delta_a"], l)                           # A3
    out["bit_raw"] = eng.paillier_decrypt(out["result"])
    return out


def check_against_transcripts(out, transcripts, l):
    B = len(transcripts)
    assert limbs_to_ints(out["z_enc"]) == [t.z_enc for t in transcripts]
    assert limbs_to_ints(out["z"]) == [t.z for t in transcripts]
    got_db = limbs_to_ints(out["db_enc"].reshape(B * (l + 1), -1))
    want_db = [v for t in transcripts for v in [t.d_enc_sent] + list(t.beta_enc_sent)]
    assert got_db == want_db
    assert limbs_to_ints(out["c_enc"].reshape(B * (l + 1), -1)) == [v for t in transcripts for v in t.c_enc_sent]
    assert out["delta_b"].tolist() == [t.delta_b for t in transcripts]
    assert limbs_to_ints(out["enc3"].reshape(B * 3, -1)) == [v for t in transcripts for v in (t.zeta1_sent, t.zeta2_sent, t.delta_b_sent)]
    assert limbs_to_ints(out["result"]) == [t.result_enc for t in transcripts]
    assert limbs_to_ints(out["bit_raw"]) == [t.result_bit for t in transcripts]


@pytest.mark.parametrize("name", GOLDEN_FILES)
@pytest.mark.parametrize("mode", ["interactive", "non_interactive"])
def test_fused_phases_reproduce_golden_vectors(name, mode):
    l, paillier, dgk, cases = load_golden(name)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    arr = pack_case_arrays(cases, l, eng.kN, eng.kn, eng.rand_words, eng.rb_words)
    out = run_fused(eng, paillier, dgk, l, arr, interactive=(mode == "interactive"))
    check_against_transcripts(out, [c[mode] for c in cases], l)
    assert limbs_to_ints(out["bit_raw"]) == [int(c["x"] <= c["y"]) for c in cases]


@pytest.mark.parametrize("name", ["golden_p1024_d128_l16.json", "golden_p512_d192_l8.json"])
def test_individual_steps_reproduce_golden_vectors(name):
    """Initiator.step_* / KeyHolder.step_* one by one (the reference's step-by-step flow)."""
    l, paillier, dgk, cases = load_golden(name)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    ts = [c["non_interactive"] for c in cases]
    B = len(cases)
    arr = pack_case_arrays(cases, l, eng.kN, eng.kn, eng.rand_words, eng.rb_words)
    flat = lambda a: limbs_to_ints(a.reshape(-1, a.shape[-1]))  # noqa: E731
    z_enc, r = Initiator.step_1(arr["x"], arr["y"], l, paillier, r=arr["r"])
    assert flat(z_enc) == [t.z_enc_step1 for t in ts]
    z, beta = KeyHolder.step_2(z_enc, l, paillier)
    assert flat(z) == [t.z for t in ts] and beta.tolist() == [t.beta for t in ts]
    alpha = Initiator.step_3(r, l)
    assert alpha.tolist() == [t.alpha for t in ts]
    d_enc = KeyHolder.step_4a(z, dgk, paillier, l)
    assert flat(d_enc) == [t.d_enc_4a for t in ts]
    beta_enc = KeyHolder.step_4b(beta, l, dgk)
    assert flat(beta_enc) == [v for t in ts for v in t.beta_enc_4b]
    d_enc = Initiator.step_4c(d_enc, r, dgk, paillier)
    assert flat(d_enc) == [t.d_enc_4c for t in ts]
    xor_enc = Initiator.step_4d(alpha, beta_enc, dgk)
    assert flat(xor_enc) == [v for t in ts for v in t.xor_enc]
    w_enc, alpha_tilde = Initiator.step_4e(r, alpha, xor_enc, d_enc, paillier, dgk)
    assert alpha_tilde.tolist() == [t.alpha_tilde for t in ts]
    assert flat(w_enc) == [v for t in ts for v in t.w_enc_4e]
    w_enc = Initiator.step_4f(w_enc, dgk)
    assert flat(w_enc) == [v for t in ts for v in t.w_enc_4f]
    s, delta_a = Initiator.step_4g(B, delta_a=arr["delta_a"])
    c_enc = Initiator.step_4h(s, alpha, alpha_tilde, d_enc, beta_enc, w_enc, delta_a, dgk)
    assert flat(c_enc) == [v for t in ts for v in t.c_enc_4h]
    c_enc = Initiator.step_4i(c_enc, dgk, do_shuffle=True, r_blind=arr["r_blind"], perm=arr["perm"])
    assert flat(c_enc) == [v for t in ts for v in t.c_enc_4i]
    delta_b = KeyHolder.step_4j(c_enc, dgk)
    assert delta_b.tolist() == [t.delta_b for t in ts]
    zeta1, zeta2, delta_b_enc = KeyHolder.step_5(z, l, delta_b, paillier)
    assert flat(zeta1) == [t.zeta1_enc for t in ts] and flat(zeta2) == [t.zeta2_enc for t in ts]
    assert flat(delta_b_enc) == [t.delta_b_enc for t in ts]
    blt = Initiator.step_6(delta_a, delta_b_enc, paillier)
    assert flat(blt) == [t.beta_lt_alpha_enc for t in ts]
    result = Initiator.step_7(zeta1, zeta2, r, l, blt, paillier)
    assert flat(result) == [t.result_enc for t in ts]
    assert flat(eng.paillier_decrypt(result)) == [int(c["x"] <= c["y"]) for c in cases]
    # the fused A2 / A3 phases give the same bits as the step-by-step path
    fused_c = eng.initiator_step4(np.concatenate([KeyHolder.step_4a(z, dgk, paillier, l)[:, None], beta_enc], axis=1), r, delta_a,
                                  arr["r_blind"], arr["perm"], None, l)
    assert np.array_equal(fused_c, c_enc)
    assert np.array_equal(Initiator.steps_6_7(np.stack([zeta1, zeta2, delta_b_enc], axis=1), r, delta_a, l, paillier), result)


@pytest.mark.parametrize("pbits,dbits,vbits,l,B", [(1024, 128, 20, 16, 700), (512, 192, 24, 8, 300), (2048, 2048, 160, 16, 24),
                                                 (2048, 2048, 160, 32, 10), (4096, 3072, 160, 64, 4)])
def test_fused_phases_match_oracle_on_fresh_batches(pbits, dbits, vbits, l, B):
    paillier, dgk = make_schemes(pbits, dbits, vbits, l)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    oracle = Oracle(paillier, dgk, "auto")
    rng = random.Random(pbits + B)
    pairs = draw_inputs(rng, B, l)
    cases = []
    for x, y in pairs:
        nc = draw_nonces(rng, paillier, dgk, l)
        cases.append({"x": x, "y": y, "nonces": nc,
                      "x_enc": encrypt_input(paillier, x, rng.randrange(1, paillier.public_key.n)),
                      "y_enc": encrypt_input(paillier, y, rng.randrange(1, paillier.public_key.n))})
    arr = pack_case_arrays(cases, l, eng.kN, eng.kn, eng.rand_words, eng.rb_words)
    for interactive in (True, False):
        out = run_fused(eng, paillier, dgk, l, arr, interactive)
        ts = [oracle.compare(c["x_enc"], c["y_enc"], l, c["nonces"], interactive) for c in cases]
        check_against_transcripts(out, ts, l)
        assert limbs_to_ints(out["bit_raw"]) == [int(x <= y) for x, y in pairs]


def test_error_contract():
    """AssertionError / ValueError behaviour of the reference (initiator.py:249,286-288; keyholder.py:212)."""
    l = 16
    paillier, dgk = make_schemes(1024, 128, 20, l)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    x = ints_to_limbs([encrypt_input(paillier, 1, 7)], eng.kc)
    with pytest.raises(AssertionError):                       # 2^(l+2) < N//2 violated
        Initiator.step_1(x, x, paillier.public_key.n.bit_length() - 2, paillier)
    one = ints_to_limbs([1], eng.kn)
    with pytest.raises(AssertionError):                       # r out of [0, N)
        Initiator.step_4c(one, ints_to_limbs([paillier.public_key.n + 10], eng.kN), dgk, paillier)
    with pytest.raises(AssertionError):                       # u must exceed 2^(l+2)
        KeyHolder.step_4a(ints_to_limbs([5], eng.kN), dgk, paillier, 17)
    with pytest.raises(ZeroDivisionError):                    # non-invertible ciphertext (x = 0)
        eng.initiator_step1(np.zeros_like(x), x, ints_to_limbs([5], eng.kN), None, l)
    assert Initiator(l).l_maximum_bit_length == l
    with pytest.raises(ValueError):
        Initiator(l).scheme_paillier
    with pytest.raises(ValueError):
        KeyHolder(l).scheme_dgk


@pytest.mark.parametrize("key_spec,B", [((2048, 2048, 160, 16), 4096), ((2048, 2048, 160, 32), 4096), ((4096, 3072, 160, 64), 1024)],
                         ids=["config1-l16", "config2-l32", "config3-p4096-d3072-l64"])
def test_ciphertexts_match_oracle_at_scale(key_spec, B):
    """BASELINE configs[1..3] key sizes and bit lengths: EVERY message of the interactive flow bit-identical to the CPU
    oracle on batches of 4,096 / 4,096 / 1,024 comparisons (the oracle runs on all host cores); bench.py applies the same
    gate to a 4,096-index sample of its timed full-size batch."""
    from helpers import oracle_transcripts_parallel
    l = key_spec[3]
    paillier, dgk = make_schemes(*key_spec)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    eng.set_small_batch(False)       # the throughput geometry of full-size batches (these batches alone would get 5 digits per lane)
    rng = random.Random(65536)
    pairs = draw_inputs(rng, B, l)
    N = paillier.public_key.n
    # input ciphertexts are made on the GPU (any valid ciphertext of x / y will do; CPython would need minutes for 3,072 of them)
    plain = ints_to_limbs([v % N for pair in pairs for v in pair], eng.kN)
    rho_in = ints_to_limbs([rng.randrange(1, N) for _ in range(2 * B)], eng.kN)
    enc = limbs_to_ints(eng.paillier_encrypt(plain, rho_in))
    cases = [{"x": x, "y": y, "nonces": draw_nonces(rng, paillier, dgk, l), "x_enc": enc[2 * i], "y_enc": enc[2 * i + 1]}
             for i, (x, y) in enumerate(pairs)]
    want = oracle_transcripts_parallel(key_spec, l, cases)
    arr = pack_case_arrays(cases, l, eng.kN, eng.kn, eng.rand_words, eng.rb_words)
    out = run_fused(eng, paillier, dgk, l, arr, interactive=True)
    assert limbs_to_ints(out["z_enc"]) == [t[0] for t in want]
    assert limbs_to_ints(out["z"]) == [t[1] for t in want]
    assert limbs_to_ints(out["db_enc"].reshape(B * (l + 1), -1)) == [v for t in want for v in t[2]]
    assert limbs_to_ints(out["c_enc"].reshape(B * (l + 1), -1)) == [v for t in want for v in t[3]]
    assert out["delta_b"].tolist() == [t[4] for t in want]
    assert limbs_to_ints(out["enc3"].reshape(B * 3, -1)) == [v for t in want for v in t[5]]
    assert limbs_to_ints(out["result"]) == [t[6] for t in want]
    assert limbs_to_ints(out["bit_raw"]) == [int(x <= y) for x, y in pairs]


@pytest.mark.parametrize("x,y", [(23, 42), (42, 23), (7, 7), (-400, -383)])
def test_readme_example_call_sequence(x, y):
    """BASELINE configs[0]: the README's step-by-step flow (README.md:117-141), call for call and with the README's argument
    lists (no scheme argument where the reference has none), scalar scheme ciphertext OBJECTS as inputs, l = 16,
    Paillier-2048 / DGK-2048 (v = 160): x = 23, y = 42 -> 1."""
    from oracle.schemes import PaillierCiphertext
    l = 16
    scheme_paillier, scheme_dgk = make_schemes(2048, 2048, 160, l)
    reset_engines()
    x_enc = scheme_paillier.unsafe_encrypt(x)
    y_enc = scheme_paillier.unsafe_encrypt(y)
    alice = Initiator(l)
    bob = KeyHolder(l, scheme_paillier=scheme_paillier, scheme_dgk=scheme_dgk)
    z_enc, r = alice.step_1(x_enc, y_enc, l, scheme_paillier)
    z, beta = bob.step_2(z_enc, l, scheme_paillier)
    alpha = alice.step_3(r, l)
    d_enc = bob.step_4a(z, scheme_dgk, scheme_paillier, l)
    beta_is_enc = bob.step_4b(beta, l, scheme_dgk)
    d_enc = alice.step_4c(d_enc, r, scheme_dgk, scheme_paillier)
    alpha_is_xor_beta_is_enc = alice.step_4d(alpha, beta_is_enc)
    w_is_enc, alpha_tilde = alice.step_4e(r, alpha, alpha_is_xor_beta_is_enc, d_enc, scheme_paillier)
    w_is_enc = alice.step_4f(w_is_enc)
    s, delta_a = alice.step_4g()
    c_is_enc = alice.step_4h(s, alpha, alpha_tilde, d_enc, beta_is_enc, w_is_enc, delta_a, scheme_dgk)
    c_is_enc = alice.step_4i(c_is_enc, scheme_dgk)
    delta_b = bob.step_4j(c_is_enc, scheme_dgk)
    zeta_1_enc, zeta_2_enc, delta_b_enc = bob.step_5(z, l, delta_b, scheme_paillier)
    beta_lt_alpha_enc = alice.step_6(delta_a, delta_b_enc)
    x_leq_y_enc = alice.step_7(zeta_1_enc, zeta_2_enc, r, l, beta_lt_alpha_enc, scheme_paillier)
    raw = limbs_to_ints(x_leq_y_enc)[0]
    x_leq_y = scheme_paillier.decrypt(PaillierCiphertext(raw, scheme_paillier))       # the scheme's own decrypt, as in the README
    assert x_leq_y == int(x <= y)
    assert limbs_to_ints(get_engine(scheme_paillier).paillier_decrypt(x_leq_y_enc)) == [int(x <= y)]
