"""GPU tests of the host-buffer pipeline: chunked over several contexts (and several GPUs when the box has them), the
result must equal a single context's, run after run (a context is not re-entrant: each worker thread owns one)."""
import random

import numpy as np
import pytest

from helpers import pack_case_arrays
from oracle.workload import draw_inputs, draw_nonces, encrypt_input, make_schemes
from protocols.secure_comparison_b200 import get_engine, limbs_to_ints, reset_engines
from protocols.secure_comparison_b200.pipeline import LocalPipeline

pytestmark = pytest.mark.gpu


def _workload(B, l, paillier, dgk, eng):
    rng = random.Random(B)
    pairs = draw_inputs(rng, B, l)
    cases = [{"x": x, "y": y, "nonces": draw_nonces(rng, paillier, dgk, l),
              "x_enc": encrypt_input(paillier, x, rng.randrange(1, paillier.public_key.n)),
              "y_enc": encrypt_input(paillier, y, rng.randrange(1, paillier.public_key.n))} for x, y in pairs]
    return pairs, pack_case_arrays(cases, l, eng.kN, eng.kn, eng.rand_words, eng.rb_words)


def _single(eng, w, l):
    z_enc = eng.initiator_step1(w["x"], w["y"], w["r"], w["rho_z"], l)
    z = eng.paillier_decrypt(z_enc)
    db = eng.keyholder_step4ab(z, l, w["rho_db"])
    c = eng.initiator_step4(db, w["r"], w["delta_a"], w["r_blind"], w["perm"], w["rho_c"], l)
    _, enc3 = eng.keyholder_step4j_5(c, z, w["rho_enc3"], l)
    return eng.initiator_step6_7(enc3, w["r"], w["delta_a"], l)


@pytest.mark.parametrize("n_contexts,chunks", [(1, 3), (2, 2), (2, 5), (3, 7)])
def test_chunked_pipeline_equals_single_context(n_contexts, chunks):
    l, B = 16, 333                                       # ragged chunks
    paillier, dgk = make_schemes(1024, 128, 20, l)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    pairs, w = _workload(B, l, paillier, dgk, eng)
    want = _single(eng, w, l)
    pipe = LocalPipeline(paillier, dgk, l, n_contexts=n_contexts, chunks=chunks)
    try:
        for _ in range(3):
            assert np.array_equal(pipe.compare(w), want)
    finally:
        pipe.close()
    assert limbs_to_ints(eng.paillier_decrypt(want)) == [int(x <= y) for x, y in pairs]


def test_pipeline_over_two_gpus_in_one_process():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    l, B = 16, 500
    paillier, dgk = make_schemes(1024, 128, 20, l)
    reset_engines()
    eng = get_engine(paillier, dgk, l)
    pairs, w = _workload(B, l, paillier, dgk, eng)
    want = _single(eng, w, l)
    pipe = LocalPipeline(paillier, dgk, l, n_contexts=2, devices=[0, 1])
    try:
        assert np.array_equal(pipe.compare(w), want)
    finally:
        pipe.close()
