"""Key files: a Paillier + DGK key pair stored as hex integers in JSON, loaded into plain scheme objects.

The engine consumes scheme objects duck-typed through the attribute names of
``tno.mpc.encryption_schemes`` (``paillier.public_key.{n, n_squared, g}``,
``paillier.secret_key.{lambda_, mu}``, ``dgk.public_key.{n, g, h, u, t}``, ``dgk.secret_key.{v_p, p}``);
this module provides the minimal holders for deployments (and ``bench.py``) that keep keys in a file
instead of generating them per run (key generation is a one-off host job, reference
``keyholder.py:135-172``, "may take up to a minute" ``README.md:84``).  No arithmetic on ciphertexts
happens here.

File format (all values hex strings): ``paillier_n, paillier_p, paillier_q, dgk_n, dgk_g, dgk_h, dgk_u,
dgk_t, dgk_p, dgk_q, dgk_v_p, dgk_v_q`` -- secret fields may be absent (Initiator side).
"""
from __future__ import annotations

import json
from types import SimpleNamespace
from typing import Tuple


def schemes_from_json(path: str) -> Tuple[SimpleNamespace, SimpleNamespace]:
    with open(path) as fh:
        v = {k: int(x, 16) for k, x in json.load(fh).items()}
    n = v["paillier_n"]
    paillier = SimpleNamespace(public_key=SimpleNamespace(n=n, n_squared=n * n, g=n + 1), secret_key=None)
    if "paillier_p" in v and "paillier_q" in v:
        p, q = v["paillier_p"], v["paillier_q"]
        lam = (p - 1) * (q - 1)
        paillier.secret_key = SimpleNamespace(lambda_=lam, mu=pow(lam, -1, n), n=n, p=p, q=q)
    dgk = SimpleNamespace(public_key=SimpleNamespace(n=v["dgk_n"], g=v["dgk_g"], h=v["dgk_h"], u=v["dgk_u"], t=v["dgk_t"]),
                          secret_key=None, randomness_bits=(5 * v["dgk_t"]) // 2)
    if "dgk_p" in v and "dgk_v_p" in v:
        dgk.secret_key = SimpleNamespace(p=v["dgk_p"], v_p=v["dgk_v_p"], q=v.get("dgk_q"), v_q=v.get("dgk_v_q"))
    return paillier, dgk
