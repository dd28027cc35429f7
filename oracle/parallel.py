"""Oracle transcripts on all host cores (TEST INFRASTRUCTURE: the checker of tests/ and of bench.py's parity gate; never on
the product path).  The oracle needs ~56 ms per Paillier-2048 comparison on one core, so ciphertext-level parity of a few
thousand comparisons is computed in worker processes (spawned: no CUDA state is inherited)."""
from __future__ import annotations

import dataclasses
import multiprocessing as mp
import os
from typing import Dict, List, Optional, Sequence, Tuple


def _oracle_worker(args):
    key_spec, l, chunk = args
    from oracle.protocol import Nonces, Oracle
    from oracle.workload import make_schemes
    paillier, dgk = make_schemes(*key_spec)
    oracle = Oracle(paillier, dgk, "auto")
    out = []
    for x_enc, y_enc, nonce_fields in chunk:
        t = oracle.compare(x_enc, y_enc, l, Nonces(**nonce_fields), interactive=True, decrypt_result=False)
        out.append((t.z_enc, t.z, [t.d_enc_sent] + list(t.beta_enc_sent), list(t.c_enc_sent), t.delta_b,
                    (t.zeta1_sent, t.zeta2_sent, t.delta_b_sent), t.result_enc))
    return out


def oracle_transcripts_parallel(key_spec, l, cases, processes=None):
    """[(z_enc, z, db_sent[l+1], c_sent[l+1], delta_b, enc3, result_enc)] for every case (dicts with ``x_enc``, ``y_enc`` ints and
    ``nonces``: an ``oracle.protocol.Nonces``), computed in ``processes`` workers."""
    processes = max(1, min(processes or (os.cpu_count() or 1), len(cases)))
    items = [(c["x_enc"], c["y_enc"], dataclasses.asdict(c["nonces"])) for c in cases]
    chunks = [items[i::processes] for i in range(processes)]
    with mp.get_context("spawn").Pool(processes) as pool:
        parts = pool.map(_oracle_worker, [(key_spec, l, ch) for ch in chunks])
    out = [None] * len(items)
    for p, part in enumerate(parts):
        for j, t in enumerate(part):
            out[p + j * processes] = t
    return out


def _ints(row) -> int:
    return int.from_bytes(row.astype("<u4", copy=False).tobytes(), "little")


def cases_from_arrays(w: Dict[str, "object"], indices: Sequence[int]) -> List[dict]:
    """Rebuild oracle inputs (Python ints + ``Nonces``) from the limb arrays of a workload dict (``x, y, r, rho_z, delta_a,
    r_blind, perm, rho_c, rho_db, rho_enc3`` as laid out for the C ABI) at the given batch indices."""
    from oracle.protocol import Nonces
    cases = []
    for i in indices:
        rho_db = [_ints(v) for v in w["rho_db"][i]]
        rho3 = [_ints(v) for v in w["rho_enc3"][i]]
        nc = Nonces(_ints(w["r"][i]), _ints(w["rho_z"][i]), int(w["delta_a"][i]), [_ints(v) for v in w["r_blind"][i]],
                    [int(v) for v in w["perm"][i]], [_ints(v) for v in w["rho_c"][i]], rho_db[0], rho_db[1:], rho3[0], rho3[1], rho3[2])
        cases.append({"x_enc": _ints(w["x"][i]), "y_enc": _ints(w["y"][i]), "nonces": nc})
    return cases


def check_messages_against_oracle(key_spec: Tuple[int, int, int, int], l: int, w: dict, got: dict, indices: Sequence[int],
                                  processes: Optional[int] = None) -> dict:
    """Ciphertext-level parity gate: every boundary message of the protocol at ``indices`` (``got``: rows of ``z_enc, z, db, c,
    delta_b, enc3, result`` ALREADY sliced to those indices, limb arrays) against the oracle run on the same inputs and nonces.
    Returns ``{"ciphertexts_checked": n, "bit_identical": bool, "mismatches": [...]}``."""
    want = oracle_transcripts_parallel(key_spec, l, cases_from_arrays(w, indices), processes)
    bad = []
    checked = 0
    for j, (idx, t) in enumerate(zip(indices, want)):
        z_enc, z, db, c, delta_b, enc3, result = t
        rows = {
            "z_enc": ([_ints(got["z_enc"][j])], [z_enc]), "z": ([_ints(got["z"][j])], [z]),
            "db": ([_ints(v) for v in got["db"][j]], db), "c": ([_ints(v) for v in got["c"][j]], c),
            "delta_b": ([int(got["delta_b"][j])], [delta_b]), "enc3": ([_ints(v) for v in got["enc3"][j]], list(enc3)),
            "result": ([_ints(got["result"][j])], [result]),
        }
        for name, (have, expect) in rows.items():
            checked += len(expect)
            if have != list(expect) and len(bad) < 8:
                bad.append({"index": int(idx), "message": name})
    return {"comparisons_checked": len(indices), "ciphertexts_checked": checked, "bit_identical": not bad, "mismatches": bad,
            "messages": ["z_enc", "z", "db_enc", "c_enc", "delta_b", "enc3", "result"]}
