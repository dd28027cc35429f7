"""Run the UNMODIFIED reference step functions on the oracle's stand-in schemes (TEST INFRASTRUCTURE).

Only usable where ``/root/reference`` exists (the build container); never on the GPU box.  The
reference's randomness (``secrets.randbelow`` / ``secrets.choice`` imported by name at
``initiator.py:5``, and each scheme's randomness pool) is replaced by injected nonces in the draw
order documented in SURVEY.md section 8c.
"""
from __future__ import annotations

import os
import sys
from typing import List

from .protocol import Nonces, Transcript
from .schemes import DGK, InjectedRandomness, Paillier, PaillierCiphertext

REFERENCE_SRC = "/root/reference/src"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tno_shim")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "tno", "mpc", "protocols", "secure_comparison"))


def import_reference():
    """Import ``tno.mpc.protocols.secure_comparison`` from /root/reference/src over the shim."""
    if not reference_available():
        raise ImportError("/root/reference is not present on this machine")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in (root, _SHIM, REFERENCE_SRC):
        if path not in sys.path:
            sys.path.insert(0, path)
    import tno.mpc.protocols.secure_comparison as ref  # type: ignore
    import tno.mpc.protocols.secure_comparison.initiator as ref_initiator  # type: ignore
    return ref, ref_initiator


class _Scripted:
    """Replacement for ``randbelow`` / ``choice`` inside the reference's initiator module."""

    def __init__(self, nc: Nonces, n: int, u: int):
        self.nc, self.n, self.u = nc, n, u
        self.blind = iter(nc.r_blind)
        self.perm = iter(nc.perm)
        self.remaining: List[int] = []

    def randbelow(self, bound: int) -> int:
        if bound == self.n:
            return self.nc.r
        if bound == 2:
            return self.nc.delta_a
        if bound == self.u - 1:
            return next(self.blind) - 1
        raise AssertionError(f"unexpected randbelow({bound})")

    def choice(self, values):
        # reference shuffle (initiator.py:212-226): repeatedly choice() + remove() from a shrinking copy.
        # perm[k] is the ORIGINAL index of the element emitted k-th.
        if not self.remaining or len(values) == len(self.nc.perm):
            self.remaining = list(range(len(values)))
        orig = next(self.perm)
        pos = self.remaining.index(orig)
        self.remaining.pop(pos)
        return values[pos]


def run_reference(paillier: Paillier, dgk: DGK, x_enc: int, y_enc: int, l: int, nc: Nonces,
                  interactive: bool = True) -> Transcript:
    """One comparison through the reference's own ``Initiator.step_*`` / ``KeyHolder.step_*`` in the
    order of ``perform_secure_comparison`` (with its ``randomize()`` calls when ``interactive``)."""
    ref, ref_initiator = import_reference()
    Initiator, KeyHolder = ref.Initiator, ref.KeyHolder
    script = _Scripted(nc, paillier.public_key.n, dgk.public_key.u)
    old = ref_initiator.randbelow, ref_initiator.choice
    ref_initiator.randbelow, ref_initiator.choice = script.randbelow, script.choice
    old_src = paillier.randomness_source, dgk.randomness_source
    t = Transcript()
    try:
        def send(ct):  # serialisation reads the value and clears the fresh flag (U8)
            return ct.get_value()

        x_ct, y_ct = PaillierCiphertext(x_enc, paillier), PaillierCiphertext(y_enc, paillier)
        z_ct, r = Initiator.step_1(x_ct, y_ct, l, paillier)
        assert r == nc.r
        t.z_enc_step1 = z_ct.peek_value()
        if interactive:
            paillier.randomness_source = InjectedRandomness([nc.rho_z])
            z_ct.randomize()
        t.z_enc = send(z_ct)
        t.alpha = Initiator.step_3(r, l)
        t.z, t.beta = KeyHolder.step_2(z_ct, l, paillier)
        d_ct = KeyHolder.step_4a(t.z, dgk, paillier, l)
        beta_cts = KeyHolder.step_4b(t.beta, l, dgk)
        t.d_enc_4a = d_ct.peek_value()
        t.beta_enc_4b = [b.peek_value() for b in beta_cts]
        if interactive:
            dgk.randomness_source = InjectedRandomness([nc.rho_d] + list(nc.rho_beta))
            d_ct.randomize()
            for b in beta_cts:
                b.randomize()
        t.d_enc_sent = send(d_ct)
        t.beta_enc_sent = [send(b) for b in beta_cts]
        d_ct = Initiator.step_4c(d_ct, r, dgk, paillier)
        t.d_enc_4c = d_ct.peek_value()
        xor_cts = Initiator.step_4d(t.alpha, beta_cts)
        t.xor_enc = [c.peek_value() for c in xor_cts]
        w_cts, t.alpha_tilde = Initiator.step_4e(r, t.alpha, xor_cts, d_ct, paillier)
        t.w_enc_4e = [c.peek_value() for c in w_cts]
        w_cts = Initiator.step_4f(w_cts)
        t.w_enc_4f = [c.peek_value() for c in w_cts]
        s, delta_a = Initiator.step_4g()
        assert delta_a == nc.delta_a
        c_cts = Initiator.step_4h(s, t.alpha, t.alpha_tilde, d_ct, beta_cts, w_cts, delta_a, dgk)
        t.c_enc_4h = [c.peek_value() for c in c_cts]
        c_cts = Initiator.step_4i(c_cts, dgk, do_shuffle=True)
        t.c_enc_4i = [c.peek_value() for c in c_cts]
        if interactive:
            dgk.randomness_source = InjectedRandomness(nc.rho_c)
            for c in c_cts:
                c.randomize()
        t.c_enc_sent = [send(c) for c in c_cts]
        t.delta_b = KeyHolder.step_4j(c_cts, dgk)
        z1, z2, db = KeyHolder.step_5(t.z, l, t.delta_b, paillier)
        t.zeta1_enc, t.zeta2_enc, t.delta_b_enc = z1.peek_value(), z2.peek_value(), db.peek_value()
        if interactive:
            paillier.randomness_source = InjectedRandomness([nc.rho_zeta1, nc.rho_zeta2, nc.rho_delta_b])
            z1.randomize()
            z2.randomize()
            db.randomize()
        t.zeta1_sent, t.zeta2_sent, t.delta_b_sent = send(z1), send(z2), send(db)
        blt = Initiator.step_6(delta_a, db)
        t.beta_lt_alpha_enc = blt.peek_value()
        res = Initiator.step_7(z1, z2, r, l, blt, paillier)
        t.result_enc = res.peek_value()
        t.result_bit = int(paillier.decrypt(res))
    finally:
        ref_initiator.randbelow, ref_initiator.choice = old
        paillier.randomness_source, dgk.randomness_source = old_src
    return t
