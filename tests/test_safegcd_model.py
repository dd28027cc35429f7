"""CPU pin of the modular-inverse algorithm of k_modinv (tests/safegcd_model.py) against Python's pow(x, -1, n)."""
import math
import random

import pytest

from safegcd_model import inv_cap, inv_len, modinv


@pytest.mark.parametrize("bits", [33, 64, 192, 1024, 2048, 4096])
def test_safegcd_model_matches_python(bits):
    rng = random.Random(bits)
    k32 = (bits + 31) // 32
    worst = 0
    for _ in range(12 if bits <= 1024 else 3):
        n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
        for x in [1, 2, n - 1, n - 2, rng.randrange(1, n), (1 << (32 * k32)) - 1, 3 * rng.randrange(1, n // 3)]:
            inv, ok, its = modinv(x, n, k32)
            worst = max(worst, its)
            if math.gcd(x, n) == 1:
                assert ok and inv == pow(x, -1, n)
            else:
                assert not ok and inv == 0
    assert worst < inv_cap(k32)          # the cap is never what ends the loop
    assert inv_len(k32) * 30 >= 32 * k32 + 2


def test_safegcd_model_rejects_non_units():
    n = 3 * 5 * 7 * 1000003
    for x in (0, 3, 15, 21 * 1000003, n):
        assert modinv(x, n, 2)[:2] == (0, False)
    assert modinv(2, n, 2)[:2] == (pow(2, -1, n), True)
