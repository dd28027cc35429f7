"""Import-name stand-in for ``tno.mpc.encryption_schemes.dgk`` (see ../../../../README.md)."""
from oracle.schemes import DGK, DGKCiphertext, DGKPublicKey, DGKSecretKey  # noqa: F401
