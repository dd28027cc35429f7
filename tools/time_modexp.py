#!/usr/bin/env python3
"""Device-side timing of the batched modexp kernel (C5 kernel sweep): modexp/s vs the integer roofline.

    python tools/time_modexp.py [bits ...]      (run on the GPU box)
"""
import json
import random
import sys

import numpy as np
import torch

from protocols.secure_comparison_b200 import Engine, ints_to_limbs, limbs_to_ints
from protocols.secure_comparison_b200 import _native as nat

PEAK_MAC = 18.4e12   # measured IMAD.WIDE.U32 rate, profiles/r01_imad_microbench.jsonl


def algorithmic_mac(bits, ebits):
    k = bits // 32
    return (ebits + -(-ebits // 5) + 32) * (2 * k * k + k)


def main():
    args = sys.argv[1:]
    batches = None
    if "--batches" in args:
        j = args.index("--batches")
        batches = [int(v) for v in args[j + 1].split(",")]
        args = args[:j] + args[j + 2:]
    sizes = [int(a) for a in args] or [2048, 4096]
    eng = Engine(0)
    stream = torch.cuda.ExternalStream(eng.stream_ptr)
    rng = random.Random(1)
    for bits in sizes:
        n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        k = eng.load_modulus(nat.MOD_USER0, n)
        for B in (batches or (1 << 12, 1 << 14, 1 << 16)):
            if batches is None and bits >= 4096 and B > (1 << 14):
                continue
            base = np.random.default_rng(0).integers(0, 1 << 32, size=(B, k), dtype=np.uint64).astype(np.uint32)
            base[:, -1] &= 0x7FFFFFFF
            exp = np.random.default_rng(1).integers(0, 1 << 32, size=(B, k), dtype=np.uint64).astype(np.uint32)
            exp[:, -1] |= 0x80000000
            d_base = torch.from_numpy(base.view(np.int32)).cuda()
            d_exp = torch.from_numpy(exp.view(np.int32)).cuda()
            d_out = torch.empty_like(d_base)
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                eng.modexp_dev(nat.MOD_USER0, d_base, d_exp, k, False, bits, None, d_out, B)   # warm-up
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.modexp_dev(nat.MOD_USER0, d_base, d_exp, k, False, bits, None, d_out, B)
                e1.record(stream)
            eng.sync()
            ms = e0.elapsed_time(e1)
            got = limbs_to_ints(d_out[:4].cpu().numpy().view(np.uint32))
            want = [pow(int.from_bytes(base[i].tobytes(), "little"), int.from_bytes(exp[i].tobytes(), "little"), n) for i in range(4)]
            rate = B / (ms * 1e-3)
            print(json.dumps({"bits": bits, "B": B, "ms": round(ms, 3), "modexp_per_s": round(rate, 1), "parity_sample": got == want,
                              "frac_of_int_roofline": round(rate * algorithmic_mac(bits, bits) / PEAK_MAC, 4)}), flush=True)


if __name__ == "__main__":
    main()
