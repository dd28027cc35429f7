"""Lane-level Python model of csrc/bn29.cuh::mont_mul (radix 2^29, TPI lanes x L digits).

Used by the CPU test-suite to pin the kernel's algorithm (column shifts, cross-lane hand-overs, lazy
digit bounds, 64-bit accumulator head-room) against Python integers; the CUDA code is a transcription
of this data flow."""
W = 29
MASK = (1 << W) - 1
U64 = (1 << 64) - 1


def to_digits(x, K):
    return [(x >> (W * i)) & MASK for i in range(K)]


def from_digits(d):
    return sum(v << (W * i) for i, v in enumerate(d))


def mont_consts(n, K):
    R = 1 << (W * K)
    assert R > 4 * n and n % 2 == 1
    return {"np": (-pow(n, -1, 1 << W)) % (1 << W), "one": R % n, "r2": R * R % n, "r3": R * R * R % n, "R": R}


def mont_mul_lanes(a, b, n, np_, L, TPI):
    """a, b, n: digit lists of length K = L*TPI (lazy digits <= 2^29 allowed). Returns lazy digit list."""
    K = L * TPI
    acc = [[0] * L for _ in range(TPI)]
    peak = 0
    for j in range(K):
        bj = b[j]
        for t in range(TPI):
            for i in range(L):
                acc[t][i] += a[t * L + i] * bj
        q = ((acc[0][0] & 0xFFFFFFFF) * np_) & MASK
        t0 = [acc[t][0] + n[t * L] * q for t in range(TPI)]
        for t in range(TPI):
            for i in range(1, L):
                acc[t][i - 1] = acc[t][i] + n[t * L + i] * q
                peak = max(peak, acc[t][i - 1])
        assert t0[0] & MASK == 0
        for t in range(TPI):
            peak = max(peak, t0[t])
            acc[t][0] += t0[t] >> W
            acc[t][L - 1] = 0 if t == TPI - 1 else (t0[t + 1] & MASK)
    assert peak <= U64, "64-bit column accumulator overflow"
    r = [0] * K
    carry_out = []
    for t in range(TPI):
        c = 0
        for i in range(L):
            c += acc[t][i]
            r[t * L + i] = c & MASK
            c >>= W
        carry_out.append(c)
    assert carry_out[-1] == 0
    for t in range(1, TPI):
        c2 = carry_out[t - 1] + r[t * L]
        r[t * L] = c2 & MASK
        c2 = (c2 >> W) + r[t * L + 1]
        r[t * L + 1] = c2 & MASK
        r[t * L + 2] += c2 >> W
        assert r[t * L + 2] <= (1 << W)
    return r
