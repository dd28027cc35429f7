"""Engine: one CUDA context (one GPU) of the batched big-integer engine, generic primitives.

Thin, typed Python over the C ABI (include/scb200.h).  Host entry points take/return numpy ``uint32``
limb arrays (``[B][k]``, little-endian); ``*_dev`` entry points take device tensors (anything with
``data_ptr()``, e.g. ``torch.uint8/int32`` tensors viewed as 32-bit words) and only enqueue work.

These primitives stand where the reference calls ``tno.mpc.encryption_schemes.utils.pow_mod`` /
``mod_inv`` (-> ``gmpy2.powmod`` / ``gmpy2.invert``) and the modular product inside ``scheme.add``.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _native as nat
from .limbs import int_to_limbs, limbs_for_bits


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[scb {code}] {message}")
        self.code = code


class NotInvertibleError(ZeroDivisionError):
    """An element had no modular inverse (gmpy2 raises ZeroDivisionError, CPython ValueError)."""


class Engine:
    def __init__(self, device: int = 0):
        self.lib = nat.load_library()
        handle = ctypes.c_void_p()
        rc = self.lib.scb_ctx_create(device, ctypes.byref(handle))
        if rc != nat.SCB_OK or not handle.value:
            raise EngineError(rc, f"cannot create a CUDA context on device {device} (no GPU visible? there is no CPU fallback)")
        self.ctx = handle
        self.device = device
        self._moduli: dict[int, tuple[int, int]] = {}   # slot -> (modulus, k32)
        self._fixed_base: dict[int, tuple[int, int]] = {}  # tslot -> (slot, exp_bits)

    # ------------------------------------------------------------------ plumbing
    def close(self) -> None:
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.scb_ctx_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc == nat.SCB_OK:
            return
        msg = self.lib.scb_last_error(self.ctx).decode(errors="replace")
        if rc == nat.SCB_ENOTINV:
            raise NotInvertibleError(msg)
        if rc == nat.SCB_EINVAL:
            raise ValueError(msg)
        raise EngineError(rc, msg)

    def sync(self) -> None:
        self._check(self.lib.scb_sync(self.ctx))

    @property
    def stream_ptr(self) -> int:
        return int(self.lib.scb_stream(self.ctx) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.lib.scb_launch_count(self.ctx))

    def measure_peak(self, which: str) -> float:
        """Issue-rate ceiling of a multiplier pipe measured on this GPU, thread-level ops/s: ``"imad_wide"`` or ``"dfma"``."""
        out = ctypes.c_double(0.0)
        self._check(self.lib.scb_measure_peak(self.ctx, {"imad_wide": nat.PEAK_IMAD_WIDE, "dfma": nat.PEAK_DFMA}[which], ctypes.byref(out)))
        return float(out.value)

    def profile_begin(self) -> None:
        self._check(self.lib.scb_profile_begin(self.ctx))

    def profile_end(self) -> dict:
        import json
        buf = ctypes.create_string_buffer(1 << 16)
        self._check(self.lib.scb_profile_end(self.ctx, buf, len(buf)))
        return json.loads(buf.value.decode())

    # ------------------------------------------------------------------ device CSPRNG (nonces)
    def rng_seed(self, key: Optional[bytes] = None, nonce: Optional[bytes] = None) -> None:
        """Key the device ChaCha20 generator (32-byte key, 8-byte nonce; default: fresh from ``secrets``)."""
        import secrets
        key = secrets.token_bytes(32) if key is None else bytes(key)
        nonce = secrets.token_bytes(8) if nonce is None else bytes(nonce)
        if len(key) != 32 or len(nonce) != 8:
            raise ValueError("key must be 32 bytes and nonce 8 bytes")
        k = np.frombuffer(key, dtype="<u4").astype(np.uint32)
        n = np.frombuffer(nonce, dtype="<u4").astype(np.uint32)
        self._check(self.lib.scb_rng_seed(self.ctx, nat._ptr(k), nat._ptr(n)))

    def rng_below(self, bound: int, B: int, k: Optional[int] = None, min_one: bool = False, out=None) -> Optional[np.ndarray]:
        """``[B][k]`` uniform integers in ``[min_one, bound)``.  ``out``: a device buffer (anything with ``data_ptr()``)
        to fill in place; without it the values are copied back to a numpy array."""
        k = k or limbs_for_bits(bound.bit_length())
        bw = int_to_limbs(bound, k)
        if out is not None:
            self._check(self.lib.scb_rng_below_dev(self.ctx, nat._ptr(bw), k, int(min_one), nat._ptr(out), B))
            return None
        host = np.empty((B, k), dtype=np.uint32)
        self._check(self.lib.scb_rng_fetch(self.ctx, nat.RNG_BELOW, nat._ptr(bw), k, int(min_one), nat._ptr(host), B))
        return host

    def rng_bits(self, bits: int, B: int, words: Optional[int] = None, out=None) -> Optional[np.ndarray]:
        words = words or limbs_for_bits(bits)
        if out is not None:
            self._check(self.lib.scb_rng_bits_dev(self.ctx, bits, words, nat._ptr(out), B))
            return None
        host = np.empty((B, words), dtype=np.uint32)
        self._check(self.lib.scb_rng_fetch(self.ctx, nat.RNG_BITS, None, words, bits, nat._ptr(host), B))
        return host

    def rng_coins(self, B: int, out=None) -> Optional[np.ndarray]:
        if out is not None:
            self._check(self.lib.scb_rng_coins_dev(self.ctx, nat._ptr(out), B))
            return None
        host = np.empty((B,), dtype=np.uint8)
        self._check(self.lib.scb_rng_fetch(self.ctx, nat.RNG_COINS, None, 0, 0, nat._ptr(host), B))
        return host

    def rng_permutations(self, n: int, B: int, out=None) -> Optional[np.ndarray]:
        if out is not None:
            self._check(self.lib.scb_rng_permutations_dev(self.ctx, n, nat._ptr(out), B))
            return None
        host = np.empty((B, n), dtype=np.uint8)
        self._check(self.lib.scb_rng_fetch(self.ctx, nat.RNG_PERMUTATIONS, None, 0, n, nat._ptr(host), B))
        return host

    def rng_block(self, state12_15) -> np.ndarray:
        """One raw ChaCha20 block for explicit state words 12..15 (known-answer tests)."""
        cn = np.asarray(state12_15, dtype=np.uint32)
        host = np.empty((16,), dtype=np.uint32)
        self._check(self.lib.scb_rng_fetch(self.ctx, nat.RNG_BLOCK, nat._ptr(cn), 0, 0, nat._ptr(host), 1))
        return host

    # ------------------------------------------------------------------ moduli
    def load_modulus(self, slot: int, modulus: int, k32: Optional[int] = None) -> int:
        """Load an odd modulus; returns the limb count ``k32`` every operand of this slot must have."""
        if modulus % 2 == 0 or modulus < 3:
            raise ValueError("modulus must be odd and >= 3")
        k = k32 if k32 is not None else limbs_for_bits(modulus.bit_length())
        words = int_to_limbs(modulus, k)
        self._check(self.lib.scb_modulus_load(self.ctx, slot, nat._ptr(words), k))
        self._moduli[slot] = (modulus, k)
        return k

    def link_square(self, square_slot: int, root_slot: int) -> None:
        """Declare ``square_slot`` (loaded with n^2) the square of ``root_slot`` (loaded with n): its exponentiations then
        run in base-n pair form on the FP64 engine -- half the digit products, bit-identical results."""
        self._check(self.lib.scb_modulus_link_square(self.ctx, square_slot, root_slot))

    def set_square_modulus(self, enabled: bool) -> None:
        """Pair-form exponentiation for linked square moduli (default) or n^2 as an ordinary modulus; bit-identical results."""
        self._check(self.lib.scb_set_option(self.ctx, nat.OPT_SQUARE_MODULUS, int(bool(enabled))))

    def set_small_batch(self, enabled: bool) -> None:
        """Small batches on twice the lanes per number (default) or always the 10-digits-per-lane geometry; bit-identical results."""
        self._check(self.lib.scb_set_option(self.ctx, nat.OPT_SMALL_BATCH, int(bool(enabled))))

    def set_overlap_crt(self, enabled: bool) -> None:
        """The p and q halves of the KeyHolder's CRT forms on two streams (default) or one after the other; bit-identical results."""
        self._check(self.lib.scb_set_option(self.ctx, nat.OPT_OVERLAP_CRT, int(bool(enabled))))

    def modulus(self, slot: int) -> int:
        return self._moduli[slot][0]

    def k32(self, slot: int) -> int:
        return self._moduli[slot][1]

    # ------------------------------------------------------------------ host-buffer primitives
    def modexp(self, slot: int, base: np.ndarray, exp: np.ndarray, exp_bits: Optional[int] = None,
               mul_in: Optional[np.ndarray] = None) -> np.ndarray:
        """``out[i] = (mul_in[i] *) base[i] ** exp[i] mod n``.  ``exp``: ``[B][ew]`` or shared ``[ew]``."""
        k = self.k32(slot)
        base = nat._host_u32(base)
        if base.ndim != 2 or base.shape[1] != k:
            raise ValueError(f"base must be [B][{k}]")
        B = base.shape[0]
        exp = nat._host_u32(exp)
        shared = exp.ndim == 1
        ew = exp.shape[-1]
        if not shared and exp.shape[0] != B:
            raise ValueError("exp must be [B][ew] or [ew]")
        if exp_bits is None:
            exp_bits = 32 * ew
        if mul_in is not None:
            mul_in = nat._host_u32(mul_in, (B, k))
        out = np.empty((B, k), dtype=np.uint32)
        self._check(self.lib.scb_modexp(self.ctx, slot, nat._ptr(base), nat._ptr(exp), ew, int(shared), int(exp_bits),
                                        nat._ptr(mul_in), nat._ptr(out), B))
        return out

    def modmul(self, slot: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        k = self.k32(slot)
        a = nat._host_u32(a)
        b = nat._host_u32(b, a.shape)
        if a.ndim != 2 or a.shape[1] != k:
            raise ValueError(f"operands must be [B][{k}]")
        out = np.empty_like(a)
        self._check(self.lib.scb_modmul(self.ctx, slot, nat._ptr(a), nat._ptr(b), nat._ptr(out), a.shape[0]))
        return out

    def modinv(self, slot: int, a: np.ndarray, raise_on_failure: bool = True):
        """Returns ``out`` (and the per-element ``ok`` flags when ``raise_on_failure`` is False)."""
        k = self.k32(slot)
        a = nat._host_u32(a)
        if a.ndim != 2 or a.shape[1] != k:
            raise ValueError(f"operand must be [B][{k}]")
        out = np.empty_like(a)
        ok = np.zeros(a.shape[0], dtype=np.uint8)
        rc = self.lib.scb_modinv(self.ctx, slot, nat._ptr(a), nat._ptr(out), nat._ptr(ok), a.shape[0])
        if rc == nat.SCB_ENOTINV and not raise_on_failure:
            return out, ok
        self._check(rc)
        return out if raise_on_failure else (out, ok)

    def set_modexp_fp64(self, enabled: bool) -> None:
        """Large exponentiations on the FP64-pipe multiplier (default) or on the IMAD one; bit-identical results."""
        self._check(self.lib.scb_set_option(self.ctx, nat.OPT_MODEXP_FP64, int(bool(enabled))))

    def set_sliding_windows(self, enabled: bool) -> None:
        """Sliding-window programs for the library's shared exponents (default) or fixed windows; bit-identical results."""
        self._check(self.lib.scb_set_option(self.ctx, nat.OPT_SLIDING_WINDOWS, int(bool(enabled))))

    def set_fixed_base_wide(self, enabled: bool) -> None:
        """16-bit-window fixed-base tables (default) vs 8-bit windows only; applies to tables loaded afterwards."""
        self._check(self.lib.scb_set_option(self.ctx, nat.OPT_FIXED_BASE_WIDE, int(bool(enabled))))

    def fixed_base_load(self, tslot: int, slot: int, base: int, exp_bits: int) -> None:
        k = self.k32(slot)
        words = int_to_limbs(base % self.modulus(slot), k)
        self._check(self.lib.scb_fixed_base_load(self.ctx, tslot, slot, nat._ptr(words), int(exp_bits)))
        self._fixed_base[tslot] = (slot, exp_bits)

    def fixed_base_exp(self, tslot: int, exp: np.ndarray, mul_in: Optional[np.ndarray] = None) -> np.ndarray:
        slot, _ = self._fixed_base[tslot]
        k = self.k32(slot)
        exp = nat._host_u32(exp)
        if exp.ndim != 2:
            raise ValueError("exp must be [B][ew]")
        B, ew = exp.shape
        if mul_in is not None:
            mul_in = nat._host_u32(mul_in, (B, k))
        out = np.empty((B, k), dtype=np.uint32)
        self._check(self.lib.scb_fixed_base_exp(self.ctx, tslot, nat._ptr(exp), ew, nat._ptr(mul_in), nat._ptr(out), B))
        return out

    # ------------------------------------------------------------------ device-buffer primitives
    def modexp_dev(self, slot: int, base, exp, exp_words: int, exp_shared: bool, exp_bits: int, mul_in, out, B: int) -> None:
        self._check(self.lib.scb_modexp_dev(self.ctx, slot, nat._ptr(base), nat._ptr(exp), exp_words, int(exp_shared), exp_bits,
                                            nat._ptr(mul_in), nat._ptr(out), B))

    def modmul_dev(self, slot: int, a, b, out, B: int) -> None:
        self._check(self.lib.scb_modmul_dev(self.ctx, slot, nat._ptr(a), nat._ptr(b), nat._ptr(out), B))

    def modinv_dev(self, slot: int, a, out, ok, B: int) -> None:
        self._check(self.lib.scb_modinv_dev(self.ctx, slot, nat._ptr(a), nat._ptr(out), nat._ptr(ok), B))

    def fixed_base_exp_dev(self, tslot: int, exp, exp_words: int, mul_in, out, B: int) -> None:
        self._check(self.lib.scb_fixed_base_exp_dev(self.ctx, tslot, nat._ptr(exp), exp_words, nat._ptr(mul_in), nat._ptr(out), B))
