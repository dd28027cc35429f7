"""Big-integer backends of the oracle (TEST INFRASTRUCTURE).

The reference reaches its arithmetic through ``tno.mpc.encryption_schemes.utils.pow_mod`` /
``mod_inv`` which dispatch to ``gmpy2.powmod`` / ``gmpy2.invert`` when the ``[gmpy]`` extra is
installed (reference ``pyproject.toml:41-44``, ``README.md:49-53``) and to CPython ``pow`` otherwise.
Both dispatch targets are restated here:

* ``PythonBackend`` -- CPython ``int`` / three-argument ``pow``.
* ``GmpBackend``    -- GMP (``libgmp.so.10``, the library gmpy2 wraps) through ctypes with
  hand-declared prototypes (the image ships the runtime library without headers):
  ``mpz_powm``, ``mpz_invert``, ``mpz_mul``/``mpz_mod``.

Every result is the canonical residue in ``[0, modulus)``, which is what both gmpy2 and CPython
return; a non-invertible input raises ``ZeroDivisionError`` (gmpy2) / ``ValueError`` (CPython) in the
reference -- both are mapped to ``ZeroDivisionError`` here.
"""
from __future__ import annotations

import ctypes
import ctypes.util
from typing import Optional


class PythonBackend:
    name = "cpython"

    @staticmethod
    def powmod(base: int, exp: int, mod: int) -> int:
        try:
            return pow(base, exp, mod)
        except ValueError as exc:  # negative exponent with a non-invertible base
            raise ZeroDivisionError(str(exc)) from exc

    @staticmethod
    def invert(a: int, mod: int) -> int:
        try:
            return pow(a, -1, mod)
        except ValueError as exc:
            raise ZeroDivisionError(str(exc)) from exc

    @staticmethod
    def mulmod(a: int, b: int, mod: int) -> int:
        return a * b % mod


class _Mpz(ctypes.Structure):
    # struct __mpz_struct { int _mp_alloc; int _mp_size; mp_limb_t *_mp_d; }  (GMP 6.x ABI)
    _fields_ = [("alloc", ctypes.c_int), ("size", ctypes.c_int), ("d", ctypes.c_void_p)]


class _GmpFns:
    """Plain-named handles on the ``__gmpz_*`` exports (avoids Python's class-private name mangling)."""

    def __init__(self, lib: ctypes.CDLL) -> None:
        P = ctypes.POINTER(_Mpz)
        sigs = {
            "init": ([P], None),
            "clear": ([P], None),
            "import_": ([P, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.c_size_t,
                         ctypes.c_char_p], None),
            "export": ([ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t), ctypes.c_int, ctypes.c_size_t,
                        ctypes.c_int, ctypes.c_size_t, P], ctypes.c_void_p),
            "powm": ([P, P, P, P], None),
            "invert": ([P, P, P], ctypes.c_int),
            "mul": ([P, P, P], None),
            "mod": ([P, P, P], None),
            "sizeinbase": ([P, ctypes.c_int], ctypes.c_size_t),
            "probab_prime_p": ([P, ctypes.c_int], ctypes.c_int),
        }
        for name, (argtypes, restype) in sigs.items():
            fn = getattr(lib, "__gmpz_" + name.rstrip("_"))
            fn.argtypes = argtypes
            fn.restype = restype
            setattr(self, name, fn)


def _load_gmp() -> Optional[ctypes.CDLL]:
    for cand in ("libgmp.so.10", ctypes.util.find_library("gmp")):
        if not cand:
            continue
        try:
            return ctypes.CDLL(cand)
        except OSError:
            continue
    return None


class GmpBackend:
    """GMP through ctypes.  One instance owns a few scratch ``mpz_t``; not thread-safe."""

    name = "gmp"
    _lib = None

    def __init__(self) -> None:
        if GmpBackend._lib is None:
            lib = _load_gmp()
            if lib is None:
                raise OSError("libgmp.so.10 not found")
            GmpBackend._lib = _GmpFns(lib)
        self._z = [_Mpz() for _ in range(5)]
        for z in self._z:
            GmpBackend._lib.init(ctypes.byref(z))
        self._mod_cache: dict[int, _Mpz] = {}

    def __del__(self) -> None:  # pragma: no cover - interpreter teardown order
        lib = GmpBackend._lib
        if lib is None:
            return
        try:
            for z in self._z:
                lib.clear(ctypes.byref(z))
            for z in self._mod_cache.values():
                lib.clear(ctypes.byref(z))
        except Exception:
            pass

    # -- conversions ---------------------------------------------------------------------------
    def _set(self, z: _Mpz, value: int) -> None:
        if value < 0:
            raise ValueError("GmpBackend operands must be non-negative")
        raw = value.to_bytes((value.bit_length() + 7) // 8 or 1, "little")
        self._lib.import_(ctypes.byref(z), len(raw), -1, 1, 0, 0, raw)

    def _get(self, z: _Mpz) -> int:
        nbytes = (self._lib.sizeinbase(ctypes.byref(z), 2) + 7) // 8
        buf = ctypes.create_string_buffer(nbytes or 1)
        count = ctypes.c_size_t(0)
        self._lib.export(buf, ctypes.byref(count), -1, 1, 0, 0, ctypes.byref(z))
        return int.from_bytes(buf.raw[: count.value], "little")

    def _modulus(self, mod: int) -> _Mpz:
        z = self._mod_cache.get(mod)
        if z is None:
            if len(self._mod_cache) > 64:
                for old in self._mod_cache.values():
                    self._lib.clear(ctypes.byref(old))
                self._mod_cache.clear()
            z = _Mpz()
            self._lib.init(ctypes.byref(z))
            self._set(z, mod)
            self._mod_cache[mod] = z
        return z

    # -- the three primitives gmpy2 offers on this path -----------------------------------------
    def powmod(self, base: int, exp: int, mod: int) -> int:
        if exp < 0:
            base, exp = self.invert(base % mod, mod), -exp
        r, b, e = self._z[0], self._z[1], self._z[2]
        self._set(b, base % mod if base < 0 else base)
        self._set(e, exp)
        self._lib.powm(ctypes.byref(r), ctypes.byref(b), ctypes.byref(e), ctypes.byref(self._modulus(mod)))
        return self._get(r)

    def invert(self, a: int, mod: int) -> int:
        r, x = self._z[0], self._z[1]
        self._set(x, a % mod)
        ok = self._lib.invert(ctypes.byref(r), ctypes.byref(x), ctypes.byref(self._modulus(mod)))
        if not ok:
            raise ZeroDivisionError("invert() no inverse exists")
        return self._get(r)

    def mulmod(self, a: int, b: int, mod: int) -> int:
        r, x, y = self._z[0], self._z[1], self._z[2]
        self._set(x, a)
        self._set(y, b)
        self._lib.mul(ctypes.byref(r), ctypes.byref(x), ctypes.byref(y))
        self._lib.mod(ctypes.byref(r), ctypes.byref(r), ctypes.byref(self._modulus(mod)))
        return self._get(r)

    def is_probable_prime(self, value: int, reps: int = 32) -> bool:
        z = self._z[3]
        self._set(z, value)
        return self._lib.probab_prime_p(ctypes.byref(z), reps) > 0


def gmp_available() -> bool:
    return _load_gmp() is not None


def get_backend(name: str = "auto"):
    """``"cpython"``, ``"gmp"`` or ``"auto"`` (GMP when loadable, as the reference does with gmpy2)."""
    if name == "cpython":
        return PythonBackend()
    if name == "gmp":
        return GmpBackend()
    return GmpBackend() if gmp_available() else PythonBackend()
