"""Limb-level Python model of csrc/kernels_generic.cuh::k_modinv (Bernstein-Yang "safegcd" divsteps, 30 divsteps
per transition matrix, signed 30-bit limbs).  The CUDA kernel is a transcription of this data flow; the CPU
test-suite pins the model against Python's pow(x, -1, n) (iteration cap, limb counts, sign handling)."""
M30 = (1 << 30) - 1
U32 = (1 << 32) - 1


def s32(x):
    x &= U32
    return x - (1 << 32) if x >> 31 else x


def to_limbs30(x, ln):
    return [(x >> (30 * i)) & M30 for i in range(ln)]


def from_limbs30(v):
    return sum(int(w) << (30 * i) for i, w in enumerate(v))      # top limb may be negative


def inv_len(k32):
    return (32 * k32 + 2 + 29) // 30


def inv_cap(k32):
    bits = 32 * k32
    return (49 * bits + 57) // 17 // 30 + 2


def divsteps_30(zeta, f0, g0):
    u, v, q, r = 1, 0, 0, 1
    f, g = f0 & U32, g0 & U32
    for _ in range(30):
        c1 = U32 if zeta < 0 else 0
        c2 = (-(g & 1)) & U32
        x = ((f ^ c1) - c1) & U32
        y = ((u ^ c1) - c1) & U32
        z = ((v ^ c1) - c1) & U32
        g = (g + (x & c2)) & U32
        q = (q + (y & c2)) & U32
        r = (r + (z & c2)) & U32
        c1 &= c2
        zeta = (zeta ^ s32(c1)) - 1
        f = (f + (g & c1)) & U32
        u = (u + (q & c1)) & U32
        v = (v + (r & c1)) & U32
        g >>= 1
        u = (u << 1) & U32
        v = (v << 1) & U32
    return zeta, (s32(u), s32(v), s32(q), s32(r))


def update_de(d, e, t, mod, ninv30):
    u, v, q, r = t
    ln = len(d)
    sd, se = (-1 if d[-1] < 0 else 0), (-1 if e[-1] < 0 else 0)
    md = (u & sd) + (v & se)
    me = (q & sd) + (r & se)
    cd = u * d[0] + v * e[0]
    ce = q * d[0] + r * e[0]
    md -= (ninv30 * (cd & U32) + md) & M30
    me -= (ninv30 * (ce & U32) + me) & M30
    cd += mod[0] * md
    ce += mod[0] * me
    assert cd & M30 == 0 and ce & M30 == 0
    cd >>= 30
    ce >>= 30
    for i in range(1, ln):
        cd += u * d[i] + v * e[i] + mod[i] * md
        ce += q * d[i] + r * e[i] + mod[i] * me
        assert -(1 << 63) <= cd < (1 << 63) and -(1 << 63) <= ce < (1 << 63)
        d[i - 1] = cd & M30
        e[i - 1] = ce & M30
        cd >>= 30
        ce >>= 30
    assert -(1 << 31) <= cd < (1 << 31) and -(1 << 31) <= ce < (1 << 31)
    d[ln - 1], e[ln - 1] = cd, ce


def update_fg(f, g, t):
    u, v, q, r = t
    ln = len(f)
    cf = u * f[0] + v * g[0]
    cg = q * f[0] + r * g[0]
    assert cf & M30 == 0 and cg & M30 == 0
    cf >>= 30
    cg >>= 30
    for i in range(1, ln):
        cf += u * f[i] + v * g[i]
        cg += q * f[i] + r * g[i]
        assert -(1 << 63) <= cf < (1 << 63) and -(1 << 63) <= cg < (1 << 63)
        f[i - 1] = cf & M30
        g[i - 1] = cg & M30
        cf >>= 30
        cg >>= 30
    assert -(1 << 31) <= cf < (1 << 31) and -(1 << 31) <= cg < (1 << 31)
    f[ln - 1], g[ln - 1] = cf, cg


def normalize(r, f_negative, mod):
    ln = len(r)
    for _pass in range(2):
        add = r[-1] < 0
        neg = f_negative and _pass == 0
        for i in range(ln):
            if add:
                r[i] += mod[i]
            if neg:
                r[i] = -r[i]
        for i in range(ln - 1):
            r[i + 1] += r[i] >> 30
            r[i] &= M30
    return r


def modinv(x, n, k32):
    """Returns (inverse or 0, ok, outer iterations)."""
    ln = inv_len(k32)
    mod = to_limbs30(n, ln)
    ninv30 = pow(n, -1, 1 << 30)
    f, g = list(mod), to_limbs30(x, ln)
    d, e = [0] * ln, [0] * ln
    e[0] = 1
    zeta = -1
    its = 0
    for its in range(1, inv_cap(k32) + 1):
        zeta, t = divsteps_30(zeta, f[0] | ((f[1] & 3) << 30), g[0] | ((g[1] & 3) << 30))
        update_de(d, e, t, mod, ninv30)
        update_fg(f, g, t)
        if not any(g):
            break
    else:
        return 0, False, its
    fv = from_limbs30(f)
    if fv not in (1, -1):
        return 0, False, its
    r = normalize(d, fv < 0, mod)
    return from_limbs30(r), True, its
