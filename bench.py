#!/usr/bin/env python3
"""bench.py -- secure comparisons/s of the batched DGK/Veugen protocol on B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 2 --warmup 3                    # our arm (CUDA engine)
    python bench.py --impl reference --steps 1 --warmup 1            # CPU reference arm (oracle port over GMP)
    python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...

One "step" = one pass of the whole protocol (all five party-phases: A1 step_1+randomize, B1 decrypt +
steps 4a/4b+randomize, A2 steps 3,4c-4i+randomize, B2 steps 4j,5+randomize, A3 steps 6,7) over one
batch of B comparisons with synthetic seeded inputs and host-supplied nonces.  Workload at N=1:
BASELINE.json configs[1]: B = 65,536, l = 16, Paillier-2048 / DGK-2048 (v = 160).  Multi-GPU: the batch
shards by comparison index, one process + one context per GPU, no collective on the data path (weak
scaling: B per GPU is fixed; NCCL only carries the barrier and the max-over-ranks of the timings).

`value`   : comparisons/s with every input resident in HBM (device-pointer C-ABI entry points), CUDA
            events on the engine's stream, max over ranks.
`e2e`     : the same through the host-buffer C-ABI calls (pinned host inputs; H2D of every input and D2H
            of every message/result inside the timed region).
`roofline`: integer-multiply pipe (IMAD.WIDE.U32 on the fmaheavy pipe), not HBM / tensor: DESIGN.md "Roofline".

The product path never touches ``oracle/``: keys come from a key file through the package's own loader.
``oracle`` is imported only by the ``cpu_baseline`` leg and by ``--impl reference``.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "secure comparisons/sec (Paillier-2048/DGK-2048, l=16)"   # the headline configuration; other --l / key sizes are labelled in `config`
UNIT = "comparisons/s"
KEY_FILE = os.path.join(ROOT, "tests", "golden", "keys_p2048_d2048_v160_l16.json")   # seeded synthetic key pair (data)


def key_file(paillier_bits: int, dgk_bits: int, l: int) -> str:
    """Seeded synthetic key pairs (data, committed under tests/golden): the headline configuration and the key sizes /
    bit lengths of BASELINE configs[2] (l = 32) and configs[3] (Paillier-4096 / DGK-3072, l = 64)."""
    return os.path.join(ROOT, "tests", "golden", f"keys_p{paillier_bits}_d{dgk_bits}_v160_l{l}.json")
# Integer-multiply peak: IMAD.WIDE.U32 (32x32+64 -> 64) issues on the fmaheavy pipe only, 8 lanes/clk/SMSP =
# 32 MAC/clk/SM (ncu: sm__pipe_fmaheavy_cycles_active; microbenchmark profiles/r01_imad_microbench2.jsonl measured
# 8.6 TMAC/s sustained = 29.6/clk/SM at the nominal clock).  Peak used for `frac` = 148 SMs x 32 x 1.965 GHz.
INT_PEAK_MAC_PER_S = 148 * 32 * 1.965e9
INT_PEAK_MEASURED_MAC_PER_S = 8.596e12
# SURVEY.md section 8d algorithmic work, 32x32->64 MAC units (fixed-window modexp, k 32-bit limbs: (e + e/5 + 32)(2k^2 + k))
MAC_PER_COMPARISON_L16 = 567e6      # reference-faithful operation list of one interactive comparison
MAC_PER_MODEXP_4096_2048 = 81.91e6  # one rho^N mod N^2 (k = 128 limbs, 2048-bit exponent)
# fixed (x, y) pairs of the reference's tests (test_secure_comparison.py:27-37), non-negative ones
FIXED_PAIRS = [(0, 2), (1, 10), (230, 269), (1508, 2408), (3122, 6048), (4250, 7804), (8668, 9015), (23, 42), (42, 23), (7, 7), (0, 0)]
# ncu --set full of k_modexp52<10,8> at B = 32768 (profiles/r01y_ncu_k_modexp52_10x8_final.txt): dram read + write bytes per launch
NCU_TRAFFIC_BYTES_PER_COMPARISON_STEP1 = (2.916283e9 + 1.426988e9) / 32768


def _measured_hbm_gbs() -> float:
    """HBM copy bandwidth measured by the driver on this pool (MEASURED_PEAKS.json), else the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"])
    except Exception:
        return 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="comparisons per GPU per step")
    ap.add_argument("--l", type=int, default=16)
    ap.add_argument("--paillier-bits", type=int, default=2048)
    ap.add_argument("--dgk-bits", type=int, default=2048)
    ap.add_argument("--cpu-per-core", type=int, default=128, help="comparisons per host core in the CPU sample (~7 s per core at 2048 bits)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        pw = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------- workload
def make_workload(B: int, l: int, seed: int, eng):
    """Seeded synthetic batch: plaintext pairs, nonces (host numpy), input ciphertexts made on the GPU (untimed)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    x = rng.integers(0, 1 << l, size=B, dtype=np.uint64)
    y = rng.integers(0, 1 << l, size=B, dtype=np.uint64)
    fixed = FIXED_PAIRS + [((1 << l) - 1, (1 << l) - 1)]
    for i, (a, b) in enumerate(fixed[:B]):
        x[i], y[i] = a, b
    kN = eng.kN

    def rand_limbs(shape, top_mask=0x7FFFFFFF):
        a = rng.integers(0, 1 << 32, size=shape, dtype=np.uint64).astype(np.uint32)
        a[..., -1] &= top_mask          # < 2^(32k-1) <= N: N has its top bit set
        a[..., 0] |= 1                  # never 0
        return a

    def rand_bits(shape_words, bits):
        a = rng.integers(0, 1 << 32, size=shape_words, dtype=np.uint64).astype(np.uint32)
        extra = 32 * shape_words[-1] - bits
        if extra:
            a[..., -1] &= (1 << (32 - extra)) - 1
        return a

    rand_bits_n = eng._fixed_base[0][1]
    u = eng.u
    # blinding exponents in [1, u-1]: odd values below 2^(bitlen(u)-1) (a subset of the range; the arithmetic is the same)
    r_blind = rand_bits((B, l + 1, eng.rb_words), u.bit_length() - 1)
    r_blind[..., 0] |= 1
    perm = np.argsort(rng.random((B, l + 1)), axis=1).astype(np.uint8)
    plain = np.zeros((2, B, kN), dtype=np.uint32)
    plain[0, :, 0] = (x & 0xFFFFFFFF).astype(np.uint32)
    plain[1, :, 0] = (y & 0xFFFFFFFF).astype(np.uint32)
    if kN > 1:
        plain[0, :, 1] = (x >> 32).astype(np.uint32)
        plain[1, :, 1] = (y >> 32).astype(np.uint32)
    rho_in = rand_limbs((2, B, kN))
    enc = eng.paillier_encrypt(plain.reshape(2 * B, kN), rho_in.reshape(2 * B, kN)).reshape(2, B, 2 * kN)
    w = {
        "x": enc[0], "y": enc[1], "r": rand_limbs((B, kN)), "rho_z": rand_limbs((B, kN)),
        "delta_a": rng.integers(0, 2, size=B, dtype=np.uint8), "r_blind": r_blind, "perm": perm,
        "rho_c": rand_bits((B, l + 1, eng.rand_words), rand_bits_n), "rho_db": rand_bits((B, l + 1, eng.rand_words), rand_bits_n),
        "rho_enc3": rand_limbs((B, 3, kN)),
    }
    truth = (x <= y).astype(np.uint8)
    return w, truth


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch

    from protocols.secure_comparison_b200 import SchemeEngine
    from protocols.secure_comparison_b200 import _native as nat
    from protocols.secure_comparison_b200.keyfile import schemes_from_json
    from protocols.secure_comparison_b200.sharding import shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, l = args.batch, args.l
    # weak scaling: the global batch is world * B comparisons; this rank owns the contiguous block [lo, hi)
    lo, hi = shard_range(world * B, rank, world)
    assert hi - lo == B
    paillier, dgk = schemes_from_json(key_file(args.paillier_bits, args.dgk_bits, l))
    eng = SchemeEngine(local)
    eng.load_paillier(paillier)
    eng.load_dgk(dgk, l)
    w, truth = make_workload(B, l, seed=1234 + rank, eng=eng)
    kN, kc, kn = eng.kN, eng.kc, eng.kn
    stream = torch.cuda.ExternalStream(eng.stream_ptr)
    lib, ctx = eng.lib, eng.ctx

    def dev(a):
        return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).cuda()

    d = {k: dev(v) for k, v in w.items()}
    i32 = dict(dtype=torch.int32, device="cuda")
    d_z_enc = torch.empty((B, kc), **i32)
    d_z = torch.empty((B, kN), **i32)
    d_db = torch.empty((B, l + 1, kn), **i32)
    d_c = torch.empty((B, l + 1, kn), **i32)
    d_delta_b = torch.empty((B,), dtype=torch.uint8, device="cuda")
    d_enc3 = torch.empty((B, 3, kc), **i32)
    d_res = torch.empty((B, kc), **i32)
    d_bit = torch.empty((B, kN), **i32)
    p = nat._ptr

    def step_device():
        eng._check(lib.scb_initiator_step1_dev(ctx, p(d["x"]), p(d["y"]), p(d["r"]), p(d["rho_z"]), l, p(d_z_enc), B))
        eng._check(lib.scb_paillier_decrypt_dev(ctx, p(d_z_enc), p(d_z), B))
        eng._check(lib.scb_keyholder_step4ab_dev(ctx, p(d_z), l, p(d["rho_db"]), eng.rand_words, p(d_db), B))
        eng._check(lib.scb_initiator_step4_dev(ctx, p(d_db), p(d["r"]), p(d["delta_a"]), p(d["r_blind"]), eng.rb_words, p(d["perm"]),
                                               p(d["rho_c"]), eng.rand_words, l, p(d_c), B))
        eng._check(lib.scb_keyholder_step4j_5_dev(ctx, p(d_c), p(d_z), p(d["rho_enc3"]), l, p(d_delta_b), p(d_enc3), B))
        eng._check(lib.scb_initiator_step6_7_dev(ctx, p(d_enc3), p(d["r"]), p(d["delta_a"]), l, p(d_res), B))

    def barrier():
        torch.cuda.synchronize()
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also grows every scratch buffer to its final size)
    for _ in range(max(args.warmup, 1)):
        step_device()
    eng.sync()
    launches_before = eng.launch_count

    # ---- timed region, device resident
    barrier()
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.steps):
                step_device()
            e1.record(stream)
        eng.sync()
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launch_count - launches_before
    if world > 1:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- parity gate on the last step's output: decrypted bits == plaintext truth for 100 % of the batch
    eng._check(lib.scb_paillier_decrypt_dev(ctx, p(d_res), p(d_bit), B))
    eng.sync()
    bits = d_bit.cpu().numpy().view(np.uint32)
    bits_ok = bool(np.array_equal(bits[:, 0].astype(np.uint8), truth) and not bits[:, 1:].any())

    # ---- per-kernel timing of one more step (one CUDA event after every launch, on the engine's stream) for the
    #      roofline of the dominant kernel: k_step1 = B modexps rho^N mod N^2 (4096-bit modulus, 2048-bit exponent)
    eng.profile_begin()
    step_device()
    kernels = eng.profile_end()
    step_ms_prof = sum(k["ms"] for k in kernels.values())
    top_name = max(kernels, key=lambda k: kernels[k]["ms"])
    # the B modexps rho^N mod N^2 of Initiator.step_1 + randomize: one launch of k_modexp52<10,8> (FP64 engine, default) or,
    # with the IMAD engine, the modexp fused inside k_step1<18,8>
    ebits_N = eng.paillier_n.bit_length()
    f64_names = [k for k in kernels if k.startswith("k_modexp52") and k.endswith(f"[e={ebits_N}]")]
    fp64_engine = bool(f64_names)
    k1_name = f64_names[0] if fp64_engine else "k_step1"
    k1 = kernels.get(k1_name, {"ms": 0.0, "launches": 1})
    k1_ms = k1["ms"] / max(k1["launches"], 1)
    # algorithmic MACs of B randomizers rho^N mod N^2 (SURVEY 8d formula; 81.91 M each at Paillier-2048)
    k1_mac = B * (ebits_N + -(-ebits_N // 5) + 32) * (2 * kc * kc + kc)
    achieved = k1_mac / (k1_ms * 1e-3) / 1e12 if k1_ms else 0.0
    bytes_per_launch = B * (kN * 4 + kc * 4) if fp64_engine else B * (2 * kc * 4 + 2 * kN * 4 + kc * 4)
    K52 = {4096: 80, 8192: 160}.get(64 * kN, (64 * kN + 4 + 51) // 52)
    nprod = ebits_N + -(-ebits_N // 6) + 64                 # squarings + 6-bit-window products + table (2,455 at 2048 bits)
    fp64_ops = B * nprod * (2 * K52 * K52 * 3 + K52 * 4)    # DFMA/DADD thread-instructions of one launch
    roofline = {
        "bound": "int-pipe",
        "bound_detail": "multiplier pipes, not hbm and not tensor (north star: integer-multiply roofline): " + (
            "FP64 pipe used as an exact 52x52-bit integer multiplier (DFMA anchor trick; throughput-limited on the DFMA + IADD3 instruction mix)"
            if fp64_engine else "int32-multiply pipe (IMAD.WIDE.U32 on fmaheavy)"),
        "kernel": k1_name, "kernel_share_of_step": round(k1["ms"] / step_ms_prof, 4) if step_ms_prof else None,
        "kernel_ms_per_launch": round(k1_ms, 3),
        "achieved": round(achieved, 4), "peak": round(INT_PEAK_MAC_PER_S / 1e12, 3), "unit": "TMAC/s (32x32->64, algorithmic)",
        "frac": round(achieved * 1e12 / INT_PEAK_MAC_PER_S, 4),
        "peak_source": "integer-multiply roofline of the north star: IMAD.WIDE.U32 = 32 lanes/clk/SM x 148 SMs x 1.965 GHz (what ncu's fmaheavy "
                       "pct_of_peak refers to; the IMAD microbenchmark on this pool sustains 8.6 TMAC/s, profiles/r01_imad_microbench2.jsonl)",
        "frac_of_measured_microbench_peak": round(achieved * 1e12 / INT_PEAK_MEASURED_MAC_PER_S, 4),
        "fp64_pipe_frac": round(fp64_ops / (k1_ms * 1e-3) / (148 * 64 * 1.965e9), 4) if (fp64_engine and k1_ms) else None,
        "fp64_pipe_note": "DFMA+DADD thread-instructions of the launch / (64 lanes/clk/SM x 148 x 1.965 GHz); ncu: sm__pipe_fp64_cycles_active 62 %, "
                          "ALU pipe 50 %, issue slots 60 % (profiles/r01y_ncu_k_modexp52_10x8_final.txt); a pure DFMA stream sustains 54.5/clk/SM, mixed 1:1 with "
                          "integer adds 46.7 (profiles/r01_fp64_issue_microbench.jsonl)",
        "imad_engine_reference": "same launch on the IMAD engine: 833 ms at B=65536 = 6.45 TMAC/s algorithmic, fmaheavy pipe 84-87 % busy "
                                 "(profiles/r01e_ncu_k_step1.txt, profiles/r01s_ncu_k_modexp_imad_18x4.txt)",
        "traffic": int(NCU_TRAFFIC_BYTES_PER_COMPARISON_STEP1 * B),
        "traffic_note": "dram bytes per launch scaled from an ncu capture (window tables of the ~9,500 resident groups exceed L2): "
                        "tens of GB/s, non-binding next to hbm_gbs = 6,550",
        "hbm_peak_gbs_measured": _measured_hbm_gbs(),
        "hbm_frac_of_measured": round(NCU_TRAFFIC_BYTES_PER_COMPARISON_STEP1 * B / (k1_ms * 1e-3) / 1e9 / _measured_hbm_gbs(), 5) if k1_ms else None,
        "algorithmic_bytes": int(bytes_per_launch),
        "hbm_gbs_algorithmic": round(bytes_per_launch / (k1_ms * 1e-3) / 1e9, 2) if k1_ms else None,
        "whole_step_frac": round(value / world * MAC_PER_COMPARISON_L16 / INT_PEAK_MAC_PER_S, 4),
        "whole_step_note": "whole_step_frac counts the reference-faithful 567 M MAC per comparison against the INTEGER roofline; it exceeds 1 because the "
                           "KeyHolder's CRT and the fixed-base tables do fewer multiply-adds for identical outputs and because the exponentiations run on the FP64 pipe",
        "top_kernel_by_time": top_name,
        "kernels_ms": {k: round(v["ms"], 3) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])},
    }

    # ---- the same pass with the exponentiations on the IMAD engine (the integer-pipe design of the north star; identical
    #      outputs): one warm-up pass, one timed pass, results compared bit for bit with the FP64-engine pass
    ref_res = d_res.clone()
    eng.set_modexp_fp64(False)
    step_device()
    eng.sync()
    with torch.cuda.stream(stream):
        i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        i0.record(stream)
        step_device()
        i1.record(stream)
    eng.sync()
    imad_ms = i0.elapsed_time(i1)
    imad_same = bool(torch.equal(ref_res, d_res))
    eng.set_modexp_fp64(True)
    if world > 1:
        t = torch.tensor([imad_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        imad_ms = float(t.item())
    imad_engine = {"value": round(world * B / (imad_ms * 1e-3), 2), "unit": UNIT, "ms_per_step": round(imad_ms, 3),
                   "bit_identical_to_fp64_engine": imad_same,
                   "note": "SCB_OPT_MODEXP_FP64 = 0: every exponentiation on the IMAD.WIDE engine (ncu: fmaheavy pipe 87 % busy, "
                           "profiles/r01s_ncu_k_modexp_imad_18x4.txt)"}

    # ---- production mode: every nonce drawn on the device (ChaCha20 CSPRNG keyed from `secrets`) INSIDE the timed pass,
    #      instead of the host-supplied nonce arrays the parity path uses; checked by decrypting all result bits
    eng.rng_seed()
    N_int, u_int = eng.paillier_n, eng.u

    def draw_nonces_device():
        eng.rng_below(N_int, B, kN, False, out=d["r"])
        eng.rng_below(N_int, B, kN, True, out=d["rho_z"])
        eng.rng_below(N_int, 3 * B, kN, True, out=d["rho_enc3"])
        eng.rng_below(u_int, B * (l + 1), eng.rb_words, True, out=d["r_blind"])
        eng.rng_bits(rand_bits_dgk, B * (l + 1), eng.rand_words, out=d["rho_c"])
        eng.rng_bits(rand_bits_dgk, B * (l + 1), eng.rand_words, out=d["rho_db"])
        eng.rng_coins(B, out=d["delta_a"])
        eng.rng_permutations(l + 1, B, out=d["perm"])

    rand_bits_dgk = eng._fixed_base[0][1]
    saved = {k: d[k].clone() for k in ("r", "rho_z", "rho_enc3", "r_blind", "rho_c", "rho_db", "delta_a", "perm")}
    draw_nonces_device()
    step_device()
    eng.sync()
    with torch.cuda.stream(stream):
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        draw_nonces_device()
        step_device()
        r1.record(stream)
    eng.sync()
    rng_ms = r0.elapsed_time(r1)
    eng._check(lib.scb_paillier_decrypt_dev(ctx, p(d_res), p(d_bit), B))
    eng.sync()
    rbits = d_bit.cpu().numpy().view(np.uint32)
    rng_ok = bool(np.array_equal(rbits[:, 0].astype(np.uint8), truth) and not rbits[:, 1:].any())
    for k, v in saved.items():
        d[k].copy_(v)                                   # back to the host-supplied nonces for the parity / e2e comparisons below
    torch.cuda.synchronize()
    step_device()
    eng.sync()
    if world > 1:
        t = torch.tensor([rng_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rng_ms = float(t.item())
    device_nonces = {"value": round(world * B / (rng_ms * 1e-3), 2), "unit": UNIT, "ms_per_step": round(rng_ms, 3),
                     "decrypted_bits_match_truth": rng_ok,
                     "note": "nonces (r, rho's, r_i, delta_A, permutations) generated on the GPU inside the timed pass: scb_rng_* (ChaCha20)"}

    # ---- second half of the BASELINE metric: 2048-bit modexp/s (full-length per-element exponents, modulus = DGK n),
    #      device resident, k_modexp<18,4>, SURVEY 8d unit = 20.56 M MAC per modexp
    MB = B
    g = torch.Generator(device="cuda").manual_seed(99 + rank)
    d_mb = torch.randint(-2 ** 31, 2 ** 31 - 1, (MB, kn), dtype=torch.int32, device="cuda", generator=g)
    d_mb[:, kn - 1] &= 0x3FFFFFFF                                   # < 2^(32 kn - 2) < n
    d_me = torch.randint(-2 ** 31, 2 ** 31 - 1, (MB, kn), dtype=torch.int32, device="cuda", generator=g)
    d_mo = torch.empty((MB, kn), **i32)
    torch.cuda.synchronize()

    def modexp_pass():
        eng.modexp_dev(nat.MOD_DGK_N, d_mb, d_me, kn, False, 32 * kn, None, d_mo, MB)

    modexp_pass()
    eng.sync()
    with torch.cuda.stream(stream):
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        modexp_pass()
        modexp_pass()
        m1.record(stream)
    eng.sync()
    mx_ms = m0.elapsed_time(m1) / 2
    if world > 1:
        t = torch.tensor([mx_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mx_ms = float(t.item())
    modexp_rate = world * MB / (mx_ms * 1e-3)
    modexp_2048 = {"value": round(modexp_rate, 1), "unit": "modexp/s", "bits": 32 * kn, "exponent_bits": 32 * kn, "batch_per_gpu": MB,
                   "ms": round(mx_ms, 3), "frac_of_int_peak": round(modexp_rate / world * 20.56e6 / INT_PEAK_MAC_PER_S, 4)}

    # ---- e2e: host buffers through the synchronous C-ABI calls (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        h = {k: torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).pin_memory().numpy() for k, v in w.items()}
        h = {k: (v.view(np.uint32) if v.dtype == np.int32 else v) for k, v in h.items()}

        # the public host-buffer API: LocalPipeline = the five party-phases through the synchronous host-pointer C-ABI calls,
        # batch split in two halves over two contexts (streams) so that one half's PCIe copies overlap the other's kernels
        from protocols.secure_comparison_b200.pipeline import LocalPipeline
        pipe = LocalPipeline(paillier, dgk, l, device=local, n_contexts=2, chunks=2)    # best of tools/time_pipeline.py
        res = pipe.arena.empty((B, kc))

        pipe.compare(h, res)                      # warm-up: scratch growth, pinned result buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pipe.compare(h, res)
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        E = (l + 1) * kn * 4
        h2d = B * (2 * kc * 4 + 2 * kN * 4) + B * kc * 4 + B * (kN * 4 + (l + 1) * eng.rand_words * 4) + \
            B * (E + kN * 4 + 1 + (l + 1) * (eng.rb_words * 4 + 1 + eng.rand_words * 4)) + B * (E + kN * 4 + 3 * kN * 4) + B * (3 * kc * 4 + kN * 4 + 1)
        d2h = B * kc * 4 + B * kN * 4 + B * E + B * E + B * (1 + 3 * kc * 4) + B * kc * 4
        same = bool(np.array_equal(res, d_res.cpu().numpy().view(np.uint32)))
        e2e = {"value": round(world * B * args.steps / dt, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": round(dt / args.steps * 1e3, 2), "matches_device_path": same,
               "api": "LocalPipeline.compare (host-pointer C-ABI calls, 2 contexts x 2 chunks, pinned buffers)"}
        pipe.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle as the timed CPU port, never on the product path
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_baseline import time_comparisons, time_modexp
        cpu = time_comparisons(args.paillier_bits, args.dgk_bits, 160, l, per_core=args.cpu_per_core)
        cpu["modexp_2048_per_s_per_core"] = time_modexp(2048, 96)["modexp_per_s_per_core"]     # full-length exponent, mpz_powm, one core
        cpu["speedup_e2e_vs_all_cores"] = (e2e["value"] / cpu["value"]) if e2e else None
        cpu["speedup_e2e_vs_one_core"] = (e2e["value"] / cpu["per_core_comparisons_per_s"]) if e2e else None
        cpu = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in cpu.items()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "exact integers: 52-bit digits in f64 (exponentiations) + 29-bit digits in u32/u64 (everything else)",
            "data": "synthetic (seeded keys, inputs and host-supplied nonces)",
            "config": {"workload": f"batch of {B} comparisons per GPU, l={l}, Paillier-{args.paillier_bits}/DGK-{args.dgk_bits} (v=160), interactive flow (all randomize() calls)",
                       "inputs_vs_l2": "per-step working set ~1.4 GB >> 126 MB L2", "batch_per_gpu": B, "l": l,
                       "sharding": "contiguous block of the global batch per rank, no data-path collective"},
            "parity": {"decrypted_bits_match_truth": bits_ok, "checked": int(B)},
            "gpu_launches": int(launches), "clocks": clocks.summary(), "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu,
            "modexp_2048": modexp_2048, "imad_engine": imad_engine, "device_nonces": device_nonces,
            "modexp_4096_per_s": round(B / (k1_ms * 1e-3), 1) if k1_ms else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if not bits_ok:
        raise SystemExit("bench.py: parity gate failed (decrypted bits differ from the plaintext truth)")


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU path (oracle port over GMP: the reference + gmpy2 cannot be installed offline),
    all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle.cpu_baseline import time_comparisons
    vals = []
    for _ in range(min(args.warmup, 1)):
        time_comparisons(2048, 2048, 160, args.l, per_core=1)
    res = None
    for _ in range(args.steps):
        res = time_comparisons(2048, 2048, 160, args.l, per_core=args.cpu_per_core)
        vals.append(res["value"])
    value = sum(vals) / len(vals)
    done_per_step = args.cpu_per_core * res["cores"]
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(done_per_step / value * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "arbitrary-precision int (GMP mpz)",
        "data": "synthetic (seeded keys, inputs and nonces)",
        "config": {"workload": f"bounded sample of the batch-of-{args.batch} workload: {done_per_step} comparisons per step, l={args.l}, "
                               "Paillier-2048/DGK-2048 (v=160), interactive flow", "l": args.l},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": res["sample"]},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
