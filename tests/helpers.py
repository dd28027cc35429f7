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
