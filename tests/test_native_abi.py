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


def test_option_and_slot_constants_match_the_header():
    """Every SCB_OPT_* / SCB_PEAK_* enumerator of include/scb200.h has the same value in the ctypes binding."""
    text = open(os.path.join(ROOT, "include", "scb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    enums = {name: int(val) for name, val in re.findall(r"\b(SCB_(?:OPT|PEAK)_[A-Z0-9_]+)\s*=\s*(\d+)", text)}
    assert len([k for k in enums if k.startswith("SCB_OPT_")]) >= 7
    for name, val in enums.items():
        py = name[len("SCB_"):]
        assert hasattr(nat, py), f"_native.py lacks {py}"
        assert getattr(nat, py) == val, name


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


def _compile_c_smoke(tmp_path):
    import subprocess
    exe = str(tmp_path / "c_abi_smoke")
    libdir = os.path.dirname(nat.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c_abi_smoke.c"),
           "-L", libdir, "-lscb200", f"-Wl,-rpath,{libdir}", "-o", exe]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr
    return exe


def test_header_compiles_as_c99_and_links(tmp_path):
    """include/scb200.h is plain C (no C++-isms) and a C program links against the shared library."""
    from protocols.secure_comparison_b200.build import build
    build()
    _compile_c_smoke(tmp_path)


@pytest.mark.gpu
def test_c_program_runs_the_abi(tmp_path):
    import subprocess
    exe = _compile_c_smoke(tmp_path)
    proc = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    assert "c_abi_smoke ok" in proc.stdout
