"""Lane-level Python model of the FP64-engine row loop (csrc/bn52.cuh ``mmd_rows_core``), raw bit patterns included.

The CUDA loop never sees "integers": it adds the BIT PATTERNS of doubles of the form 2^104 + hi * 2^52 and 2^52 + lo into
64-bit column accumulators and cancels the exponent fields with constants pre-loaded into every column.  The schedule
is software-pipelined (the product a_0 * b_{j+1} and the complement digit of the pair arithmetic enter one row EARLY, so
that the serial quotient chain of row j+1 can start as soon as column 0 of row j retires), which makes the constant
bookkeeping the error-prone part.  This model replays exactly that data flow with Python integers modulo 2^64 -- same
initial constants, same renaming, same hand-overs between lanes -- and the CPU test-suite checks it against big-integer
arithmetic for every geometry; the kernel is a transcription of it."""
W = 52
MASK = (1 << W) - 1
U64 = (1 << 64) - 1
BITS_2P52 = 0x4330000000000000
BITS_2P104 = 0x4670000000000000


def raw_hi(a, b):
    """bit pattern of fma_rz(a, b, 2^104) for integers a, b < 2^52"""
    return BITS_2P104 + ((a * b) >> W)


def raw_lo(a, b):
    """bit pattern of fma_rz(a, b, (2^104 + 2^52) - hi)"""
    return BITS_2P52 + ((a * b) & MASK)


def digits(x, K):
    assert 0 <= x < 1 << (W * K)
    return [(x >> (W * i)) & MASK for i in range(K)]


def undigits(d):
    return sum(v << (W * i) for i, v in enumerate(d))


def mmd_rows_core(a, b, n, np_, L, TPI, comp=None, want_q=False):
    """(a * b + [R - m] + q n) / R with R = 2^(52 K), K = L * TPI.  a, b, n: digit lists (digits < 2^52).
    comp: the K complement digits (MASK - m_j) of a previous pass, or None.  Returns (unresolved columns per lane
    [TPI][L], hi_last per lane [TPI], quotient digits)."""
    K = L * TPI
    bpad = list(b) + [0, 0]                               # two zero pad rows: the loop touches b[j+1] and loads b[j+2]
    cpad = (list(comp) + [0]) if comp is not None else None
    AL = [[(0 - 2 * (i + 1) * BITS_2P52) & U64 for i in range(L)] for _ in range(TPI)]
    AH = [[(0 - 2 * (i + 1) * BITS_2P104) & U64 for i in range(L)] for _ in range(TPI)]
    # prologue: P(0) of row 0, and the first complement digit (+1) into column 0 of lane 0
    for t in range(TPI):
        AL[t][0] = (AL[t][0] + raw_lo(a[t * L], bpad[0])) & U64
        AH[t][0] = (AH[t][0] + raw_hi(a[t * L], bpad[0])) & U64
    if cpad is not None:
        AL[0][0] = (AL[0][0] + cpad[0] + 1) & U64
    qs = []
    j = 0
    for blk in range(TPI):
        for rr in range(L):
            c = lambda i: (i + rr) % L                    # noqa: E731  logical column i of this row
            bj, bn = bpad[j], bpad[j + 1]
            # chain A: quotient digit from the (complete) column 0 of lane 0
            v0 = AL[0][c(0)] & MASK
            q = (v0 * np_) & MASK
            qs.append(q)
            for t in range(TPI):
                # Y: a_i b_j for i >= 1, and a_0 b_{j+1} one row early (into the column that is column 0 of the next row)
                for i in range(1, L):
                    AL[t][c(i)] = (AL[t][c(i)] + raw_lo(a[t * L + i], bj)) & U64
                    AH[t][c(i)] = (AH[t][c(i)] + raw_hi(a[t * L + i], bj)) & U64
                AL[t][c(1)] = (AL[t][c(1)] + raw_lo(a[t * L], bn)) & U64
                AH[t][c(1)] = (AH[t][c(1)] + raw_hi(a[t * L], bn)) & U64
                # B + Z: n_i q
                for i in range(L):
                    AL[t][c(i)] = (AL[t][c(i)] + raw_lo(n[t * L + i], q)) & U64
                    AH[t][c(i)] = (AH[t][c(i)] + raw_hi(n[t * L + i], q)) & U64
            if cpad is not None:
                AL[0][c(1)] = (AL[0][c(1)] + cpad[j + 1]) & U64
            # retire column 0: carry up inside the lane, low part down to the neighbour lane
            lows = []
            for t in range(TPI):
                t0 = AL[t][c(0)]
                AL[t][c(1)] = (AL[t][c(1)] + AH[t][c(0)] + (t0 >> W)) & U64
                lows.append(t0 & MASK)
            assert lows[0] == 0                           # Montgomery: the lowest column of the number is cleared
            for t in range(TPI):
                incoming = 0 if t == TPI - 1 else lows[t + 1]
                AL[t][c(0)] = (incoming - 2 * L * BITS_2P52) & U64
                AH[t][c(0)] = (0 - 2 * L * BITS_2P104) & U64
            j += 1
    cols, hi_last = [], []
    for t in range(TPI):
        # the column at final position p lived only L-1-p rows in this lane: give the constants of the rows it did not see
        # back; position 0 additionally received the (zero) product a_0 * b_K one row early
        al = [(AL[t][p] + 2 * (p + 1) * BITS_2P52) & U64 for p in range(L)]
        ah = [(AH[t][p] + 2 * (p + 1) * BITS_2P104) & U64 for p in range(L)]
        al[0] = (al[0] - BITS_2P52) & U64
        ah[0] = (ah[0] - BITS_2P104) & U64
        col = [al[0]] + [(al[p] + ah[p - 1]) & U64 for p in range(1, L)]
        cols.append(col)
        hi_last.append(ah[L - 1])
    return cols, hi_last, qs


def resolve(cols, hi_last, L, TPI):
    """Exact carry resolution across the lanes -> K normalised digits (value semantics of resolve_carries)."""
    total = 0
    for t in range(TPI):
        for p in range(L):
            assert cols[t][p] < 1 << 63
            total += cols[t][p] << (W * (t * L + p))
        assert hi_last[t] < 1 << 63
        total += hi_last[t] << (W * (t * L + L))
    return total


def mont_pass(a, b, n, L, TPI, comp=None):
    K = L * TPI
    np_ = (-pow(n, -1, 1 << W)) % (1 << W)
    cols, hi_last, qs = mmd_rows_core(digits(a, K), digits(b, K), digits(n, K), np_, L, TPI, comp)
    return resolve(cols, hi_last, L, TPI), qs
