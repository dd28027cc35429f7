"""GPU debugging aid: candidate matching for steps 1 and 6+7 (which wrong value did the kernel compute?)."""
import random, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.protocol import Oracle
from oracle.workload import make_schemes
from protocols.secure_comparison_b200 import get_engine, ints_to_limbs, limbs_to_ints, reset_engines

for (pb, db_, vb, l) in [(512, 192, 24, 8), (1024, 128, 20, 16)]:
    pa, dgk = make_schemes(pb, db_, vb, l)
    reset_engines()
    eng = get_engine(pa, dgk, l)
    o = Oracle(pa, dgk, "auto")
    N, N2 = o.N, o.N2
    rng = random.Random(5)
    for B in (1, 2, 5, 40):
        z = [rng.randrange(N) for _ in range(B)]
        dbit = [rng.randrange(2) for _ in range(B)]
        r = [rng.randrange(N) for _ in range(B)]
        da = [i % 2 for i in range(B)]
        t3 = [[o.paillier_randomize(c, rng.randrange(1, N)) for c in o.step_5(a, l, b)] for a, b in zip(z, dbit)]
        enc3 = ints_to_limbs([v for t in t3 for v in t], eng.kc).reshape(B, 3, eng.kc)
        res = limbs_to_ints(eng.initiator_step6_7(enc3, ints_to_limbs(r, eng.kN), np.array(da, dtype=np.uint8), l))
        for i in range(min(B, 6)):
            t = t3[i]
            want = o.step_7(t[0], t[1], r[i], l, o.step_6(da[i], t[2]))
            if res[i] == want:
                print(pb, "B", B, i, "OK"); continue
            a = (r[i] >> l) + (0 if da[i] else 1)
            cands = {}
            for zn, zeta in (("z1", t[0]), ("z2", t[1]), ("db", t[2])):
                for en, ev in (("e", (N - a) % N), ("a", a), ("0", 0), ("e'", (N - (r[i] >> l) - (1 if da[i] else 0)) % N)):
                    for vn, V in (("inv", pow(t[2], -1, N2)), ("db", t[2]), ("1", 1), ("z1", t[0]), ("z2", t[1])):
                        cands[(zn, en, vn)] = zeta * (1 + ev * N) * V % N2
            hit = [k for k, v in cands.items() if v == res[i]]
            print(pb, "B", B, i, "da", da[i], "rsmall", r[i] < (N - 1) // 2, "MISMATCH; matches:", hit)
    # step 1
    for B in (24, 45):
        x = [rng.randrange(1, N2) for _ in range(B)]
        y = [rng.randrange(1, N2) for _ in range(B)]
        rr = [rng.randrange(N) for _ in range(B)]
        got = limbs_to_ints(eng.initiator_step1(ints_to_limbs(x, eng.kc), ints_to_limbs(y, eng.kc), ints_to_limbs(rr, eng.kN), None, l))
        want = [o.step_1(a, b, l, c) for a, b, c in zip(x, y, rr)]
        bad = [i for i in range(B) if got[i] != want[i]]
        print(pb, "step1 B", B, "bad", bad)
        for i in bad[:3]:
            xi = pow(x[i], -1, N2)
            c = {"y*xinv": y[i] * xi % N2, "want+N2?": want[i], "diff": (got[i] - want[i]) % N2}
            print("   got*x/y == enc?", (got[i] * x[i] * pow(y[i], -1, N2)) % N2 == (1 + ((1 << l) + rr[i]) * N) % N2,
                  " got-want =", hex((got[i] - want[i]) % N2)[:40], "got>=N2", got[i] >= N2)
