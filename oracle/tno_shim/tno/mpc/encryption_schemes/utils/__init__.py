"""Import-name stand-in for ``tno.mpc.encryption_schemes.utils`` (see ../../../../README.md)."""
from oracle.schemes import is_probable_prime, next_prime, randprime  # noqa: F401


def pow_mod(base: int, exponent: int, modulus: int) -> int:
    return pow(base, exponent, modulus)


def mod_inv(value: int, modulus: int) -> int:
    return pow(value, -1, modulus)
