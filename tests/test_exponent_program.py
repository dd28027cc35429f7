"""CPU test (no GPU): the sliding-window programs the library compiles for shared exponents, interpreted with Python
integers, compute base^exponent for random, structured and adversarial exponents."""
import ctypes
import random

import numpy as np
import pytest

from protocols.secure_comparison_b200 import _native as nat
from protocols.secure_comparison_b200.limbs import int_to_limbs


def compile_program(e: int):
    lib = nat.load_library()
    words = int_to_limbs(e, max(1, (e.bit_length() + 31) // 32))
    prog = np.zeros(4096, dtype=np.uint16)
    w = ctypes.c_int(0)
    n = lib.scb_debug_compile_exponent(nat._ptr(words), len(words), nat._ptr(prog), len(prog), ctypes.byref(w))
    assert n >= 0
    return prog[:n].tolist(), w.value


def run_program(prog, w, x, n):
    table = [pow(x, 2 * j + 1, n) for j in range(1 << (w - 1))]
    r = table[prog[0]]
    products = (1 << (w - 1))                      # x^2 and the odd powers
    for entry in prog[1:]:
        nsq, idx = entry >> 8, entry & 0xFF
        for _ in range(nsq):
            r = r * r % n
        products += nsq
        if idx != 0xFF:
            r = r * table[idx] % n
            products += 1
    return r, products


@pytest.mark.parametrize("bits", [64, 65, 160, 400, 1022, 1024, 2048, 4096])
def test_programs_compute_the_power(bits):
    rng = random.Random(bits)
    n = rng.getrandbits(521) | 1
    for trial in range(6):
        e = rng.getrandbits(bits) | (1 << (bits - 1))
        prog, w = compile_program(e)
        assert 2 <= w <= 7 and prog
        x = rng.randrange(2, n)
        got, products = run_program(prog, w, x, n)
        assert got == pow(x, e, n)
        if bits >= 1022:                           # cheaper than the fixed-window loop it replaces
            wf = 5 if bits <= 1100 else 6
            assert products < bits + -(-bits // wf) + (1 << wf) - 2


def test_structured_exponents():
    n = (1 << 255) - 19
    x = 0x1234567
    for e in [(1 << 1500) + 1, 1 << 2047, (1 << 2048) - 1, (0xABCDEF << 900) | 1, ((1 << 300) - 1) << 700, (1 << 63), (1 << 64) + (1 << 63)]:
        prog, w = compile_program(e)
        assert prog and all((p >> 8) <= 255 for p in prog[1:])
        assert run_program(prog, w, x, n)[0] == pow(x, e, n)
    assert any((p >> 8) == 255 and (p & 0xFF) == 0xFF for p in compile_program((1 << 1500) + 1)[0])   # a split zero run


def test_short_exponents_have_no_program():
    for e in (0, 1, 3, 12345, (1 << 63) - 1):
        prog, w = compile_program(e)
        assert prog == [] and w == 0
