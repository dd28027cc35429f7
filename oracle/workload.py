"""Seeded synthetic keys, inputs and nonces (TEST INFRASTRUCTURE; SURVEY.md section 8d "Synthetic inputs").

Keys are generated once from a fixed seed and cached as JSON under ``tests/golden/`` (key generation is
a one-off host-side job -- reference ``keyholder.py:135-172`` -- and is not part of the hot path).
"""
from __future__ import annotations

import json
import os
import random
from typing import Dict, List, Optional, Tuple

from .protocol import Nonces
from .schemes import (DGK, DGKPublicKey, DGKSecretKey, Paillier, PaillierPublicKey, PaillierSecretKey, next_prime)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

#: (x, y) pairs with x < y used throughout the reference's tests (test_secure_comparison.py:27-37)
REFERENCE_PAIRS = [(-400, -383), (-1, 0), (0, 2), (1, 10), (230, 269), (1508, 2408), (3122, 6048), (4250, 7804),
                   (8668, 9015)]


DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(GOLDEN_DIR)), "data")


def key_file(paillier_bits: int, dgk_bits: int, v_bits: int, l: int) -> str:
    """Seeded key pairs: the three BASELINE key sizes live under ``data/`` (bench.py's product path loads them through the
    package's own key-file reader); the small test-size pairs under ``tests/golden/``."""
    name = f"keys_p{paillier_bits}_d{dgk_bits}_v{v_bits}_l{l}.json"
    in_data = os.path.join(DATA_DIR, name)
    return in_data if os.path.exists(in_data) else os.path.join(GOLDEN_DIR, name)


def schemes_to_dict(paillier: Paillier, dgk: DGK) -> Dict[str, str]:
    pk, sk, dpk, dsk = paillier.public_key, paillier.secret_key, dgk.public_key, dgk.secret_key
    d = {"paillier_n": pk.n, "paillier_p": sk.p, "paillier_q": sk.q, "dgk_n": dpk.n, "dgk_g": dpk.g, "dgk_h": dpk.h,
         "dgk_u": dpk.u, "dgk_t": dpk.t, "dgk_p": dsk.p, "dgk_q": dsk.q, "dgk_v_p": dsk.v_p, "dgk_v_q": dsk.v_q}
    return {k: hex(v) for k, v in d.items()}


def schemes_from_dict(d: Dict[str, str], backend=None, full_decryption: bool = False) -> Tuple[Paillier, DGK]:
    v = {k: int(x, 16) for k, x in d.items()}
    n, p, q = v["paillier_n"], v["paillier_p"], v["paillier_q"]
    lam = (p - 1) * (q - 1)
    paillier = Paillier(PaillierPublicKey(n, n + 1), PaillierSecretKey(lam, pow(lam, -1, n), n, p, q), backend=backend)
    dgk = DGK(DGKPublicKey(v["dgk_g"], v["dgk_h"], v["dgk_u"], v["dgk_n"], v["dgk_t"]),
              DGKSecretKey(v["dgk_v_p"], v["dgk_v_q"], v["dgk_p"], v["dgk_q"]), full_decryption, backend=backend)
    return paillier, dgk


def make_schemes(paillier_bits: int, dgk_bits: int, v_bits: int, l: int, seed: int = 20261018, backend=None,
                 full_decryption: bool = False, cache: bool = True) -> Tuple[Paillier, DGK]:
    """Seeded key pair (cached under tests/golden when ``cache``).  u = next_prime(2^(l+2)) as the
    reference's KeyHolder chooses it (keyholder.py:164)."""
    path = key_file(paillier_bits, dgk_bits, v_bits, l)
    if cache and os.path.exists(path):
        with open(path) as fh:
            return schemes_from_dict(json.load(fh), backend, full_decryption)
    rng = random.Random(f"{seed}/{paillier_bits}/{dgk_bits}/{v_bits}/{l}")
    paillier = Paillier.from_security_parameter(paillier_bits, rng=rng, backend=backend)
    dgk = DGK.from_security_parameter(v_bits, dgk_bits, next_prime(1 << (l + 2)), full_decryption, rng=rng, backend=backend)
    if cache:
        os.makedirs(GOLDEN_DIR, exist_ok=True)
        with open(path, "w") as fh:
            json.dump(schemes_to_dict(paillier, dgk), fh, indent=1)
    return paillier, dgk


def draw_nonces(rng: random.Random, paillier: Paillier, dgk: DGK, l: int, shuffle: bool = True) -> Nonces:
    """One comparison's randomness, drawn in the reference's order (last row)."""
    N, u, rb = paillier.public_key.n, dgk.public_key.u, dgk.randomness_bits
    r = rng.randrange(N)
    rho_z = rng.randrange(1, N)
    delta_a = rng.randrange(2)
    r_blind = [rng.randrange(u - 1) + 1 for _ in range(l + 1)]
    perm = list(range(l + 1))
    if shuffle:
        rng.shuffle(perm)
    rho_c = [rng.getrandbits(rb) for _ in range(l + 1)]
    rho_d = rng.getrandbits(rb)
    rho_beta = [rng.getrandbits(rb) for _ in range(l)]
    return Nonces(r, rho_z, delta_a, r_blind, perm, rho_c, rho_d, rho_beta,
                  rng.randrange(1, N), rng.randrange(1, N), rng.randrange(1, N))


def draw_inputs(rng: random.Random, batch: int, l: int, include_reference_pairs: bool = True) -> List[Tuple[int, int]]:
    """Plaintext pairs: the reference's fixed pairs (both orders, incl. negatives for l >= 16), forced
    equalities, then uniform values in [0, 2^l)."""
    pairs: List[Tuple[int, int]] = []
    if include_reference_pairs and l >= 16:
        for x, y in REFERENCE_PAIRS:
            pairs += [(x, y), (y, x), (y, y)]
    pairs += [(0, 0), ((1 << l) - 1, (1 << l) - 1), (0, (1 << l) - 1), ((1 << l) - 1, 0), (23, 42)]
    pairs = pairs[:batch]
    while len(pairs) < batch:
        pairs.append((rng.randrange(1 << l), rng.randrange(1 << l)))
    return pairs


def encrypt_input(paillier: Paillier, value: int, rho: Optional[int]) -> int:
    """Input ciphertext [[value]] (with encoding, so negatives wrap mod N) optionally randomised with rho."""
    N, N2 = paillier.public_key.n, paillier.public_key.n_squared
    c = (1 + (value % N) * N) % N2
    if rho is not None:
        c = c * pow(rho, N, N2) % N2
    return c
