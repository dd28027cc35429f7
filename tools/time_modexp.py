#!/usr/bin/env python3
"""Kernel sweep (BASELINE configs[4]): device-side timing of the batched modexp on both multiplier engines and of the
DGK is-zero exponentiation, against the integer-multiply roofline.

    python tools/time_modexp.py [--bits 2048,3072,4096] [--batches 1024,4096,...] [--engines fp64,imad]   (run on the GPU box)

One JSON line per (size, engine, batch).  Inputs are seeded random (modulus odd with the top bit set, bases < n,
full-length exponents with the top bit set); a sample of every result is checked against Python's pow."""
import argparse
import json
import random
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from protocols.secure_comparison_b200 import Engine, limbs_to_ints
from protocols.secure_comparison_b200 import _native as nat

PEAK_MAC = 148 * 32 * 1.965e9   # IMAD.WIDE.U32 roofline (32 lanes/clk/SM)


def algorithmic_mac(bits, ebits):
    """fixed-window (w = 5) Montgomery modexp on k 32-bit limbs."""
    k = bits // 32
    return (ebits + -(-ebits // 5) + 32) * (2 * k * k + k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", default="2048,3072,4096")
    ap.add_argument("--batches", default="1024,4096,16384,65536")
    ap.add_argument("--engines", default="fp64,imad")
    ap.add_argument("--ebits", default="", help="exponent length (default: the modulus length); e.g. 160 with --bits 1024 = DGK is-zero")
    ap.add_argument("--no-wide-lanes", action="store_true", help="keep 10 digits per lane for small batches too (SCB_OPT_SMALL_BATCH = 0)")
    ap.add_argument("--max-seconds", type=float, default=0.0, help="skip batch sizes whose predicted time (from the previous size) exceeds this")
    args = ap.parse_args()
    eng = Engine(0)
    if args.no_wide_lanes:
        eng.set_small_batch(False)
    stream = torch.cuda.ExternalStream(eng.stream_ptr)
    rng = random.Random(1)
    for bits in [int(v) for v in args.bits.split(",")]:
        ebits = int(args.ebits) if args.ebits else bits
        n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        k = eng.load_modulus(nat.MOD_USER0, n)
        ew = (ebits + 31) // 32
        for engine in args.engines.split(","):
            eng.set_modexp_fp64(engine == "fp64")
            last = None
            for B in [int(v) for v in args.batches.split(",")]:
                if args.max_seconds and last and last[1] * B / last[0] > args.max_seconds * 1e3:
                    continue
                base = np.random.default_rng(0).integers(0, 1 << 32, size=(B, k), dtype=np.uint64).astype(np.uint32)
                base[:, -1] &= 0x7FFFFFFF
                exp = np.random.default_rng(1).integers(0, 1 << 32, size=(B, ew), dtype=np.uint64).astype(np.uint32)
                if ebits % 32:
                    exp[:, -1] &= (1 << (ebits % 32)) - 1
                exp[:, -1] |= 1 << ((ebits - 1) % 32)
                d_base = torch.from_numpy(base.view(np.int32)).cuda()
                d_exp = torch.from_numpy(exp.view(np.int32)).cuda()
                d_out = torch.empty_like(d_base)
                torch.cuda.synchronize()
                with torch.cuda.stream(stream):
                    eng.modexp_dev(nat.MOD_USER0, d_base, d_exp, ew, False, ebits, None, d_out, B)   # warm-up
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    eng.modexp_dev(nat.MOD_USER0, d_base, d_exp, ew, False, ebits, None, d_out, B)
                    e1.record(stream)
                eng.sync()
                ms = e0.elapsed_time(e1)
                got = limbs_to_ints(d_out[:4].cpu().numpy().view(np.uint32))
                want = [pow(int.from_bytes(base[i].tobytes(), "little"), int.from_bytes(exp[i].tobytes(), "little"), n) for i in range(4)]
                rate = B / (ms * 1e-3)
                last = (B, ms)
                print(json.dumps({"bits": bits, "exponent_bits": ebits, "engine": engine, "B": B, "ms": round(ms, 3), "modexp_per_s": round(rate, 1),
                                  "parity_sample": got == want, "wide_lanes_for_small_batches": not args.no_wide_lanes,
                                  "frac_of_int_roofline": round(rate * algorithmic_mac(bits, ebits) / PEAK_MAC, 4)}), flush=True)


if __name__ == "__main__":
    main()
