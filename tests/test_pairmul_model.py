"""CPU pin of the square-modulus pair arithmetic (tests/pairmul_model.py = the data flow of csrc/kernels_f64_pair.cuh)
against Python's pow, including the sizes with the least Montgomery slack and inputs wider than the modulus."""
import random

import pytest

from pairmul_model import PairCtx
from test_exponent_program import compile_program as compile_program_py


def rand_odd(rng, bits):
    return rng.getrandbits(bits) | (1 << (bits - 1)) | 1


@pytest.mark.parametrize("bits", [100, 512, 1024, 1036, 1536, 2048, 2076, 4096])
def test_pair_modexp_fixed_windows(bits):
    rng = random.Random(bits)
    n = rand_odd(rng, bits)
    ctx = PairCtx(n)
    assert ctx.R >= 16 * n
    for trial in range(8):
        in_bits = rng.choice([32 * ((bits + 31) // 32), 64 * ((bits + 31) // 32), 128 * ((bits + 31) // 32)])
        x = [0, (n * n - 1) % (1 << in_bits), n, (1 << in_bits) - 1][trial] if trial < 4 else rng.getrandbits(in_bits)
        ebits = rng.choice([1, 5, 17, 64, 160])
        e = rng.getrandbits(ebits)
        assert ctx.modexp_fixed(x, e, ebits, rng.choice([1, 2, 3, 4, 5]), in_bits) == pow(x, e, n * n)


@pytest.mark.parametrize("bits,ebits", [(512, 512), (1024, 1024), (2048, 300)])
def test_pair_modexp_sliding_program(bits, ebits):
    rng = random.Random(bits + ebits)
    n = rand_odd(rng, bits)
    ctx = PairCtx(n)
    e = rng.getrandbits(ebits) | (1 << (ebits - 1))
    prog, w = compile_program_py(e)
    for x in (0, 1, n - 1, n * n - 1, rng.randrange(n * n), rng.randrange(n)):
        assert ctx.modexp_prog(x, prog, w, 64 * ((bits + 31) // 32)) == pow(x, e, n * n)


def test_two_pass_window_product_of_the_model():
    """mul_dual (a d + b c + (R - m) reduced ONCE) equals mul and keeps the (2n, 4n) bounds, also at the two sizes with the
    least Montgomery slack.  The kernel keeps the three-pass form (kernels_f64_pair.cuh explains why); the model pins the
    algebra of the alternative."""
    rng = random.Random(11)
    for bits in (1024, 1036, 2048, 2076):
        n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        ctx = PairCtx(n)
        for it in range(40):
            x = (rng.randrange(4 * n), rng.randrange(4 * n)) if it else (4 * n - 1, 4 * n - 1)
            y = (rng.randrange(4 * n), rng.randrange(4 * n)) if it else (4 * n - 1, 4 * n - 1)
            r = ctx.mul_dual(x, y)
            assert ctx.value(r) == ctx.value(ctx.mul(x, y))
            assert r[0] < 2 * n + 1 and r[1] < 4 * n
