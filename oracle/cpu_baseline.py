"""CPU baseline timing (TEST INFRASTRUCTURE; used only by bench.py's ``cpu_baseline`` / ``--impl reference`` legs).

Times the oracle's restatement of ``perform_secure_comparison`` -- the identical per-comparison operation
list of the reference (every randomizer modexp included, so the cost
of the reference's randomness pool is counted) -- with the GMP backend (``mpz_powm`` / ``mpz_invert`` /
``mpz_mul``+``mpz_mod`` of libgmp.so.10, the routines gmpy2 calls), on all host cores through
``multiprocessing``.  kind = "port": the real ``tno.*`` + ``gmpy2`` stack is not installable offline.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import random
import time
from typing import Dict, Tuple

from .bigint import gmp_available
from .protocol import Oracle
from .workload import draw_inputs, draw_nonces, encrypt_input, make_schemes

_STATE: Dict[str, object] = {}


def _init(key_spec: Tuple[int, int, int, int], backend: str) -> None:
    paillier, dgk = make_schemes(*key_spec)
    _STATE["oracle"] = Oracle(paillier, dgk, backend)
    _STATE["schemes"] = (paillier, dgk)
    _STATE["l"] = key_spec[3]


def _work(args: Tuple[int, int]) -> Tuple[int, int]:
    seed, count = args
    oracle: Oracle = _STATE["oracle"]  # type: ignore[assignment]
    paillier, dgk = _STATE["schemes"]  # type: ignore[misc]
    l: int = _STATE["l"]  # type: ignore[assignment]
    rng = random.Random(seed)
    pairs = draw_inputs(rng, count, l, include_reference_pairs=False)
    prepared = []
    for x, y in pairs:
        nc = draw_nonces(rng, paillier, dgk, l)
        # inputs are deterministic encryptions (1 + mN): their randomisation is not part of the protocol
        prepared.append((x, y, encrypt_input(paillier, x, None) * 1, encrypt_input(paillier, y, None), nc))
    ok = 0
    t0 = time.perf_counter()
    for x, y, xe, ye, nc in prepared:
        t = oracle.compare(xe, ye, l, nc, interactive=True, decrypt_result=False)
        ok += 1 if t.result_enc else 0
    dt = time.perf_counter() - t0
    return ok, int(dt * 1e9)


def time_comparisons(paillier_bits: int = 2048, dgk_bits: int = 2048, v_bits: int = 160, l: int = 16, per_core: int = 8,
                     cores: int | None = None) -> dict:
    """comparisons/s of the CPU path on ``cores`` processes (default: all).  Bounded: per_core comparisons each."""
    cores = cores or os.cpu_count() or 1
    backend = "gmp" if gmp_available() else "cpython"
    spec = (paillier_bits, dgk_bits, v_bits, l)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_init, initargs=(spec, backend)) as pool:
        pool.map(_work, [(i, 1) for i in range(cores)])          # warm-up: key load, GMP init
        t0 = time.perf_counter()
        results = pool.map(_work, [(1000 + i, per_core) for i in range(cores)])
        wall = time.perf_counter() - t0
    done = sum(r[0] for r in results)
    busy = sum(r[1] for r in results) * 1e-9
    return {
        "value": done / wall, "unit": "comparisons/s", "cores": cores, "kind": "port", "backend": backend,
        "sample": f"{done} interactive comparisons (l={l}, Paillier-{paillier_bits}/DGK-{dgk_bits}), {per_core} per process on {cores} processes",
        "wall_s": wall, "per_core_comparisons_per_s": done / busy if busy else 0.0,
    }


def time_modexp(bits: int = 2048, count: int = 64) -> dict:
    """Single-core full-length modexp/s (mpz_powm)."""
    from .bigint import get_backend
    be = get_backend("auto")
    rng = random.Random(bits)
    n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
    ops = [(rng.randrange(n), rng.getrandbits(bits) | (1 << (bits - 1))) for _ in range(count)]
    t0 = time.perf_counter()
    for b, e in ops:
        be.powmod(b, e, n)
    dt = time.perf_counter() - t0
    return {"bits": bits, "modexp_per_s_per_core": count / dt, "backend": be.name}


if __name__ == "__main__":
    import json
    print(json.dumps(time_modexp()))
    print(json.dumps(time_comparisons(per_core=2)))
