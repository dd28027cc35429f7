"""CPU tests: the C-ABI library loads and exports every symbol include/scb200.h declares; limb helpers."""
import os
import random
import re

import numpy as np
import pytest

from protocols.secure_comparison_b200 import _native as nat
from protocols.secure_comparison_b200.limbs import int_to_limbs, ints_to_limbs, limbs_to_int, limbs_to_ints

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "scb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scb_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_exports_every_declared_symbol():
    from protocols.secure_comparison_b200.build import build
    build()
    lib = nat.load_library()
    declared = header_symbols()
    assert declared, "no declarations parsed from include/scb200.h"
    for name in declared:
        assert hasattr(lib, name), f"libscb200.so does not export {name}"
    assert set(declared) == set(nat.declared_symbols()), "ctypes binding and header disagree"
    assert lib.scb_version() >= 100


def test_limb_round_trip():
    rng = random.Random(0)
    vals = [0, 1, (1 << 2048) - 1] + [rng.getrandbits(2048) for _ in range(20)]
    arr = ints_to_limbs(vals, 64)
    assert arr.shape == (23, 64) and arr.dtype == np.uint32
    assert limbs_to_ints(arr) == vals
    assert limbs_to_int(int_to_limbs(12345678901234567890, 3)) == 12345678901234567890
    assert arr[2].tolist() == [0xFFFFFFFF] * 64
    with pytest.raises(OverflowError):
        ints_to_limbs([1 << 64], 2)
