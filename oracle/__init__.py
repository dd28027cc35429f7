"""CPU oracle for the secure-comparison hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the big-integer algorithm of
``tno.mpc.protocols.secure_comparison`` v4.4.0 (reference files ``initiator.py:228-564``,
``keyholder.py:181-287``, ``utils.py:6-38``) together with the semantics of the un-vendored scheme
packages it delegates to (``tno.mpc.encryption_schemes.{paillier ~=3.0, dgk ~=3.0, templates ~=4.1,
utils ~=0.10}``, pins at ``pyproject.toml:32-38``; assumptions U1-U8 of SURVEY.md section 8c).

Nothing in the product package ``protocols.secure_comparison_b200`` may import, link or execute
anything under ``oracle/``.  The only legitimate callers are ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only as the
checker or as the timed CPU baseline, never as the thing shipped.

Parity status: PINNED at the protocol level -- ``tests/golden/make_golden.py`` runs the UNMODIFIED
reference step functions (imported from /root/reference/src) on top of the scheme stand-ins in
``oracle/tno_shim`` with injected nonces and records every raw ciphertext; the closed-form
restatement in ``oracle/protocol.py`` must reproduce those vectors bit for bit.  The scheme layer
itself (the absent ``tno.mpc.encryption_schemes`` packages) has no golden vectors anywhere, so its
semantics rest on assumptions U1-U8; the highest-risk one (U4, DGK encryption of s = -1) is
switchable (``DGK.negative_exponent_mode``).
"""
