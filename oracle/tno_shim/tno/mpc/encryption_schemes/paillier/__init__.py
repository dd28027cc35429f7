"""Import-name stand-in for ``tno.mpc.encryption_schemes.paillier`` (see ../../../../README.md)."""
from oracle.schemes import Paillier, PaillierCiphertext, PaillierPublicKey, PaillierSecretKey  # noqa: F401
