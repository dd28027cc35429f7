"""Python model of the square-modulus ("pair") arithmetic of csrc/bn52_pair.cuh / kernels_f64_pair.cuh.

A residue x mod n^2 is held as a pair (a, b) with x = a + b*n (mod n^2), both halves K-digit numbers in
radix 2^52 kept lazily in [0, 4n).  The Montgomery product for the radix R = 2^(52K) of the ROOT modulus n

    (a, b) (x) (c, d) = (a + b n)(c + d n) R^-1  mod n^2

costs three K-digit Montgomery passes (two for a squaring) instead of one 2K-digit product (= four K-digit
ones): with  a c + m n = t R  (pass 1, m = the K quotient digits of the reduction)

    x y = a c + (a d + b c) n            (mod n^2)
        = t R + (a d + b c - m) n        (mod n^2)
    x y R^-1 = t + [ (a d + b c - m) R^-1 mod n ] n
             = t + [ REDC(a d + (R - m)) + REDC(b c) - 1 ] n            (pass 2, pass 0)

(R - m is added digit by digit as the complement of m plus one, so no pass ever sees a negative number; the
-1 becomes + (n - 1)).  The kernels are a transcription of this data flow; the CPU test-suite pins it against
Python's pow()."""
W = 52
MASK = (1 << W) - 1


def K_for_bits(bits):
    for K in (20, 30, 40, 60, 80, 120, 160):
        if W * K >= bits + 4:
            return K
    raise ValueError("modulus too large")


def digits(x, K):
    assert 0 <= x < 1 << (W * K)
    return [(x >> (W * i)) & MASK for i in range(K)]


def undigits(d):
    return sum(v << (W * i) for i, v in enumerate(d))


class PairCtx:
    def __init__(self, n, K=None):
        assert n % 2 == 1 and n >= 3
        self.n = n
        self.bits = n.bit_length()
        self.K = K or K_for_bits(self.bits)
        self.R = 1 << (W * self.K)
        assert self.R >= 16 * n
        self.np = (-pow(n, -1, 1 << W)) % (1 << W)
        self.n2 = n * n
        self.w = self.bits - 1                          # piece width of the input conversion: 2^w <= n
        split = lambda v: (v % n, v // n)               # noqa: E731  host-side, one-off (key load)
        self.one = split(self.R % self.n2)              # Montgomery form of 1
        self.r2 = split(self.R * self.R % self.n2)      # to-Montgomery multiplier
        self.sR = split((1 << self.w) * self.R % self.n2)   # Montgomery form of the Horner radix 2^w

    # ---- one K-row CIOS pass: (a * stream + inject) R^-1 mod n, lazily; also returns the quotient digits
    def redc_pass(self, a, stream, comp=None):
        """a, stream: integers < R with a * stream < 16 n^2.  comp: K complement digits (MASK - m_j) of a
        previous pass, injected at column j (+1 at column 0), or None."""
        K, n = self.K, self.n
        sd = digits(stream, K)
        T = 0
        q = []
        for j in range(K):
            T += a * sd[j]
            if comp is not None:
                T += comp[j] + (1 if j == 0 else 0)
            qj = ((T & MASK) * self.np) & MASK
            T += qj * n
            assert T & MASK == 0
            T >>= W
            q.append(qj)
        return T, q

    def norm(self, v):
        assert 0 <= v < self.R                          # fits K normalised digits
        return v

    def sqr(self, x):
        a, b = x
        n = self.n
        assert a < 4 * n and b < 4 * n
        t, m = self.redc_pass(a, a)
        comp = [MASK - mj for mj in m]
        s, _ = self.redc_pass(a, self.norm(2 * b), comp)
        return self.norm(t), self.norm(s + n - 1)

    def mul(self, x, y):
        (a, b), (c, d) = x, y
        n = self.n
        assert max(a, b, c, d) < 4 * n
        t, m = self.redc_pass(a, c)
        comp = [MASK - mj for mj in m]
        s1, _ = self.redc_pass(a, d, comp)
        s2, _ = self.redc_pass(b, c)
        return self.norm(t), self.norm(s1 + s2 + n - 1)

    def redc_pass_dual(self, a, stream_a, b, stream_b, comp):
        """(a * stream_a + b * stream_b + R - m) R^-1 mod n in ONE K-row pass (one reduction for both products): the form
        the exponentiation kernel uses for window products (mmd_rows_dual)."""
        K, n = self.K, self.n
        sa, sb = digits(stream_a, K), digits(stream_b, K)
        T = 0
        for j in range(K):
            T += a * sa[j] + b * sb[j] + comp[j] + (1 if j == 0 else 0)
            qj = ((T & MASK) * self.np) & MASK
            T += qj * n
            assert T & MASK == 0
            T >>= W
        return T

    def mul_dual(self, x, y):
        """Same value as mul() with two passes: t = REDC(a c), s = REDC(a d + b c + (R - m)) + (n - 1)."""
        (a, b), (c, d) = x, y
        n = self.n
        assert max(a, b, c, d) < 4 * n
        t, m = self.redc_pass(a, c)
        comp = [MASK - mj for mj in m]
        s = self.redc_pass_dual(a, d, b, c, comp)
        return self.norm(t), self.norm(s + n - 1)

    def value(self, x):
        return (x[0] + x[1] * self.n) % self.n2

    # ---- conversions
    def pieces(self, x, in_bits):
        """x < 2^in_bits as w-bit pieces, least significant first (each < 2^w <= n)."""
        cnt = max(1, -(-in_bits // self.w))
        return [(x >> (self.w * i)) & ((1 << self.w) - 1) for i in range(cnt)]

    def to_mont(self, x, in_bits):
        """Plain integer (any value < 2^in_bits) -> Montgomery pair of x mod n^2, by Horner over the pieces."""
        ps = self.pieces(x, in_bits)
        acc = (ps[-1], 0)
        for p in reversed(ps[:-1]):
            acc = self.mul(acc, self.sR)                # acc * 2^w (plain domain)
            acc = (self.norm(acc[0] + p), acc[1])
        return self.mul(acc, self.r2)

    def from_mont(self, x):
        """Montgomery pair -> canonical integer in [0, n^2)."""
        a, b = self.mul(x, (1, 0))
        n = self.n
        carry = 0
        while a >= n:
            a -= n
            carry += 1
        assert carry <= 3
        b += carry
        cnt = 0
        while b >= n:
            b -= n
            cnt += 1
        assert cnt <= 5
        return a + b * n

    # ---- exponentiation loops, as the kernels run them
    def modexp_fixed(self, x, e, ebits, wbits, in_bits):
        xm = self.to_mont(x, in_bits)
        nwin = -(-ebits // wbits)
        if nwin == 0:
            return self.from_mont(self.one)
        table = [self.one, xm]
        for _ in range((1 << wbits) - 2):
            table.append(self.mul(table[-1], xm))
        r = table[(e >> ((nwin - 1) * wbits)) & ((1 << wbits) - 1)]
        for wi in range(nwin - 2, -1, -1):
            for _ in range(wbits):
                r = self.sqr(r)
            r = self.mul(r, table[(e >> (wi * wbits)) & ((1 << wbits) - 1)])
        return self.from_mont(r)

    def modexp_prog(self, x, prog, pw, in_bits):
        xm = self.to_mont(x, in_bits)
        x2 = self.sqr(xm)
        table = [xm]
        for _ in range((1 << (pw - 1)) - 1):
            table.append(self.mul(table[-1], x2))
        r = table[prog[0]]
        for entry in prog[1:]:
            for _ in range(entry >> 8):
                r = self.sqr(r)
            if entry & 0xFF != 0xFF:
                r = self.mul(r, table[entry & 0xFF])
        return self.from_mont(r)
