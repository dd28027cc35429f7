"""Shared helpers for the test-suite: golden-vector loading."""
import json
import os

from oracle.protocol import Nonces, Transcript
from oracle.workload import schemes_from_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_FILES = ["golden_p1024_d128_l16.json", "golden_p512_d192_l8.json", "golden_p2048_d2048_l16.json",
                "golden_p2048_d2048_l32.json", "golden_p4096_d3072_l64.json"]   # the last two: BASELINE configs[2] / configs[3] sizes


def unhex(obj):
    if isinstance(obj, str) and (obj.startswith("0x") or obj.startswith("-0x")):
        return int(obj, 16)
    if isinstance(obj, list):
        return [unhex(v) for v in obj]
    if isinstance(obj, dict):
        return {k: unhex(v) for k, v in obj.items()}
    return obj


def load_golden(name, backend=None, full_decryption=False):
    with open(os.path.join(GOLDEN, name)) as fh:
        raw = json.load(fh)
    paillier, dgk = schemes_from_dict(raw["keys"], backend, full_decryption)
    cases = []
    for c in raw["cases"]:
        cases.append({
            "x": c["x"], "y": c["y"], "x_enc": int(c["x_enc"], 16), "y_enc": int(c["y_enc"], 16),
            "nonces": Nonces(**unhex(c["nonces"])),
            "interactive": Transcript(**unhex(c["interactive"])),
            "non_interactive": Transcript(**unhex(c["non_interactive"])),
        })
    return raw["l"], paillier, dgk, cases


# ---------------------------------------------------------------------------------------------
# packing oracle-side Python ints into the limb arrays of the C ABI
# ---------------------------------------------------------------------------------------------
def pack_case_arrays(cases, l, kN, kn, rand_words, rb_words):
    """cases: list of dicts with x_enc, y_enc (ints) and nonces (oracle.protocol.Nonces)."""
    import numpy as np

    from protocols.secure_comparison_b200.limbs import ints_to_limbs

    B = len(cases)
    nc = [c["nonces"] for c in cases]
    arr = {
        "x": ints_to_limbs([c["x_enc"] for c in cases], 2 * kN),
        "y": ints_to_limbs([c["y_enc"] for c in cases], 2 * kN),
        "r": ints_to_limbs([n.r for n in nc], kN),
        "rho_z": ints_to_limbs([n.rho_z for n in nc], kN),
        "delta_a": np.array([n.delta_a for n in nc], dtype=np.uint8),
        "r_blind": ints_to_limbs([v for n in nc for v in n.r_blind], rb_words).reshape(B, l + 1, rb_words),
        "perm": np.array([n.perm for n in nc], dtype=np.uint8),
        "rho_c": ints_to_limbs([v for n in nc for v in n.rho_c], rand_words).reshape(B, l + 1, rand_words),
        "rho_db": ints_to_limbs([v for n in nc for v in [n.rho_d] + list(n.rho_beta)], rand_words).reshape(B, l + 1, rand_words),
        "rho_enc3": ints_to_limbs([v for n in nc for v in (n.rho_zeta1, n.rho_zeta2, n.rho_delta_b)], kN).reshape(B, 3, kN),
    }
    return arr


# ---------------------------------------------------------------------------------------------
# a whole protocol pass on device buffers with device-generated nonces (property tests at larger batches)
# ---------------------------------------------------------------------------------------------
def run_device_pass(eng, paillier, dgk, l, x, y):
    """x, y: [B] uint64 plaintexts in [0, 2^l).  Encrypts them, runs all five party-phases through the device-pointer
    C-ABI entry points with nonces from the device CSPRNG, decrypts the result.  Returns (bits [B][kN] uint32, r [B] ints)."""
    import numpy as np
    import torch

    from protocols.secure_comparison_b200 import _native as nat
    from protocols.secure_comparison_b200.limbs import limbs_to_ints

    B = len(x)
    eng.rng_seed()
    plain = np.zeros((2, B, eng.kN), dtype=np.uint32)
    for j, v in enumerate((x, y)):
        plain[j, :, 0] = (v & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        plain[j, :, 1] = (v >> np.uint64(32)).astype(np.uint32)
    N = paillier.public_key.n
    i32 = dict(dtype=torch.int32, device="cuda")
    dev = lambda *shape: torch.empty(shape, **i32)  # noqa: E731
    d_rho_in, d_r, d_rho_z, d_rho3 = dev(2 * B, eng.kN), dev(B, eng.kN), dev(B, eng.kN), dev(3 * B, eng.kN)
    d_rblind, d_rho_c, d_rho_db = dev(B * (l + 1), eng.rb_words), dev(B * (l + 1), eng.rand_words), dev(B * (l + 1), eng.rand_words)
    d_da = torch.empty((B,), dtype=torch.uint8, device="cuda")
    d_perm = torch.empty((B, l + 1), dtype=torch.uint8, device="cuda")
    rbits = getattr(dgk, "randomness_bits", (5 * int(dgk.public_key.t)) // 2)
    eng.rng_below(N, 2 * B, eng.kN, True, out=d_rho_in)
    eng.rng_below(N, B, eng.kN, False, out=d_r)
    eng.rng_below(N, B, eng.kN, True, out=d_rho_z)
    eng.rng_below(N, 3 * B, eng.kN, True, out=d_rho3)
    eng.rng_below(dgk.public_key.u, B * (l + 1), eng.rb_words, True, out=d_rblind)
    eng.rng_bits(rbits, B * (l + 1), eng.rand_words, out=d_rho_c)
    eng.rng_bits(rbits, B * (l + 1), eng.rand_words, out=d_rho_db)
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
    return d_bit.cpu().numpy().view(np.uint32), limbs_to_ints(d_r.cpu().numpy().view(np.uint32))


# ---------------------------------------------------------------------------------------------
# oracle transcripts on all host cores: oracle/parallel.py (shared with bench.py's parity gate)
# ---------------------------------------------------------------------------------------------
from oracle.parallel import oracle_transcripts_parallel  # noqa: E402,F401
