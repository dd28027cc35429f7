"""Shared helpers for the test-suite: golden-vector loading."""
import json
import os

from oracle.protocol import Nonces, Transcript
from oracle.workload import schemes_from_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_FILES = ["golden_p1024_d128_l16.json", "golden_p512_d192_l8.json", "golden_p2048_d2048_l16.json"]


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
