"""Mint the golden vectors of tests/golden/*.json  (run in the build container only).

    python tests/golden/make_golden.py

Imports the UNMODIFIED reference (``/root/reference/src``) on top of the scheme stand-ins in
``oracle/tno_shim`` (the reference's real dependencies are not installable here), drives it with seeded
keys, inputs and injected nonces, and records every raw intermediate of
``perform_secure_comparison`` (interactive, with ``randomize()``) and of the step-by-step flow of
``README.md:117-141`` (non-interactive).  The committed JSON is what travels to the GPU box; the CPU
oracle (``oracle/protocol.py``) and the CUDA engine are both checked against it.
"""
from __future__ import annotations

import dataclasses
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.reference_runner import run_reference  # noqa: E402
from oracle.workload import (draw_inputs, draw_nonces, encrypt_input, make_schemes, schemes_to_dict)  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hexify(obj):
    if isinstance(obj, bool) or obj is None:
        return obj
    if isinstance(obj, int):
        return hex(obj)
    if isinstance(obj, (list, tuple)):
        return [hexify(v) for v in obj]
    if isinstance(obj, dict):
        return {k: hexify(v) for k, v in obj.items()}
    return obj


def make(name: str, paillier_bits: int, dgk_bits: int, v_bits: int, l: int, pairs, seed: int) -> None:
    paillier, dgk = make_schemes(paillier_bits, dgk_bits, v_bits, l)
    rng = random.Random(seed)
    cases = []
    for x, y in pairs:
        nc = draw_nonces(rng, paillier, dgk, l)
        x_enc = encrypt_input(paillier, x, rng.randrange(1, paillier.public_key.n))
        y_enc = encrypt_input(paillier, y, rng.randrange(1, paillier.public_key.n))
        case = {"x": x, "y": y, "x_enc": hex(x_enc), "y_enc": hex(y_enc), "nonces": hexify(dataclasses.asdict(nc))}
        for mode, interactive in (("interactive", True), ("non_interactive", False)):
            t = run_reference(paillier, dgk, x_enc, y_enc, l, nc, interactive=interactive)
            assert t.result_bit == int(x <= y), (x, y, t.result_bit)
            case[mode] = hexify(dataclasses.asdict(t))
        cases.append(case)
    out = {"generator": "tests/golden/make_golden.py", "reference": "tno.mpc.protocols.secure_comparison 4.4.0",
           "l": l, "seed": seed, "keys": schemes_to_dict(paillier, dgk), "cases": cases}
    path = os.path.join(HERE, name)
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0)
    print(f"wrote {path}: {len(cases)} cases, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    rng = random.Random(99)
    only_large = "--only-large" in sys.argv
    if only_large:
        sys.argv.append("--large")
    # the reference's own test sizes: Paillier-1024, DGK n=128/v=20, l=16 (test/conftest.py:22-23,65-66)
    make("golden_p1024_d128_l16.json", 1024, 128, 20, 16, draw_inputs(rng, 40, 16), seed=1)
    # README / BASELINE config 1: x=23, y=42, l=16, Paillier-2048, DGK-2048 v=160 (+ boundary pairs)
    make("golden_p2048_d2048_l16.json", 2048, 2048, 160, 16, [(23, 42), (42, 23), (7, 7), (-400, -383), (0, 65535)],
         seed=2)
    # a second bit length on small keys (l = 8; u = next_prime(2^10))
    make("golden_p512_d192_l8.json", 512, 192, 24, 8, draw_inputs(rng, 12, 8, include_reference_pairs=False), seed=3)
    if "--large" in sys.argv:
        # BASELINE configs[2] and configs[3] key sizes and bit lengths (a few comparisons each: the files stay small)
        make("golden_p2048_d2048_l32.json", 2048, 2048, 160, 32, [(23, 42), ((1 << 32) - 1, 5), (123456789, 123456789)], seed=4)
        make("golden_p4096_d3072_l64.json", 4096, 3072, 160, 64, [(23, 42), ((1 << 64) - 1, (1 << 63) + 12345), (1 << 40, 1 << 40)], seed=5)
