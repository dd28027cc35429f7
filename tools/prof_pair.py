#!/usr/bin/env python3
"""Launch the square-modulus exponentiations of the protocol in isolation (for ncu and for kernel tuning).

    python tools/prof_pair.py [--batch 32768] [--reps 2] [--plain]        (run on the GPU box)

rho^N mod N^2 (root N: 2048 bits, geometry 10x4) and b^p mod p^2 (root p: 1024 bits, geometry 10x2) with the key pair of
the headline configuration; --plain switches the pair form off (n^2 as an ordinary modulus) for comparison.  Prints
one JSON line per launch kind with the CUDA-event time and checks a sample against Python's pow."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from protocols.secure_comparison_b200 import SchemeEngine, limbs_to_ints
from protocols.secure_comparison_b200 import _native as nat
from protocols.secure_comparison_b200.keyfile import schemes_from_json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32768)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--plain", action="store_true")
    args = ap.parse_args()
    paillier, dgk = schemes_from_json(bench.key_file(2048, 2048, 16))
    eng = SchemeEngine(0)
    eng.load_paillier(paillier)
    eng.load_dgk(dgk, 16)
    if args.plain:
        eng.set_square_modulus(False)
    B, kN = args.batch, eng.kN
    stream = torch.cuda.ExternalStream(eng.stream_ptr)
    rng = np.random.default_rng(3)
    rho = rng.integers(0, 1 << 32, size=(B, kN), dtype=np.uint64).astype(np.uint32)
    rho[:, -1] &= 0x7FFFFFFF
    plain = np.zeros((B, kN), dtype=np.uint32)
    d_rho = torch.from_numpy(rho.view(np.int32)).cuda()
    d_plain = torch.from_numpy(plain.view(np.int32)).cuda()
    d_out = torch.empty((B, 2 * kN), dtype=torch.int32, device="cuda")
    N, p = paillier.public_key.n, paillier.secret_key.p
    for name, crt in (("alice_rho_N_mod_N2", False), ("bob_crt_randomizer", True)):
        eng.set_paillier_crt(crt)
        eng._check(eng.lib.scb_paillier_encrypt_dev(eng.ctx, nat._ptr(d_plain), nat._ptr(d_rho), nat._ptr(d_out), B))   # warm-up
        eng.sync()
        eng.profile_begin()
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.reps):
                eng._check(eng.lib.scb_paillier_encrypt_dev(eng.ctx, nat._ptr(d_plain), nat._ptr(d_rho), nat._ptr(d_out), B))
            e1.record(stream)
        kernels = eng.profile_end()
        ms = e0.elapsed_time(e1) / args.reps
        got = limbs_to_ints(d_out[:4].cpu().numpy().view(np.uint32))
        want = [pow(int(r), N, N * N) for r in limbs_to_ints(rho[:4])]
        print(json.dumps({"launch": name, "batch": B, "ms": round(ms, 3), "per_s": round(B / ms * 1e3, 1), "ok": got == want,
                          "pair_form": not args.plain, "kernels_ms": {k: round(v["ms"] / args.reps, 3) for k, v in kernels.items()}}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
