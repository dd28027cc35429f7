"""Import-name stand-in for ``tno.mpc.encryption_schemes.templates`` (nothing is imported from it by
the reference package itself; present so ``import tno.mpc.encryption_schemes.templates`` resolves)."""
