"""ctypes binding of libscb200.so (C ABI: include/scb200.h).

The product path has NO CPU fallback: if the CUDA library is missing or no device is visible the
functions here raise, loudly.  (The CPU oracle under ``oracle/`` is test infrastructure and is never
imported from this package.)"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libscb200.so")

OPT_PAILLIER_CRT = 1
OPT_FIXED_BASE_WIDE = 2
OPT_MODEXP_FP64 = 3
OPT_SLIDING_WINDOWS = 4
OPT_SQUARE_MODULUS = 5
OPT_SMALL_BATCH = 6
OPT_OVERLAP_CRT = 7
PEAK_IMAD_WIDE = 0
PEAK_DFMA = 1
RNG_BELOW, RNG_BITS, RNG_COINS, RNG_PERMUTATIONS, RNG_BLOCK = 0, 1, 2, 3, 4
SCB_OK, SCB_EINVAL, SCB_ECUDA, SCB_ENOMEM, SCB_ENOTINV, SCB_ESTATE = 0, -1, -2, -3, -4, -5

MOD_PAILLIER_N2, MOD_PAILLIER_N, MOD_DGK_N, MOD_DGK_P = 0, 1, 2, 3
MOD_PAILLIER_P2, MOD_PAILLIER_Q2, MOD_PAILLIER_P, MOD_PAILLIER_Q, MOD_USER0 = 4, 5, 6, 7, 8

_u32p = ctypes.POINTER(ctypes.c_uint32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_vp = ctypes.c_void_p

_SIGNATURES = {
    "scb_ctx_create": ([ctypes.c_int, ctypes.POINTER(_vp)], ctypes.c_int),
    "scb_ctx_destroy": ([_vp], None),
    "scb_last_error": ([_vp], ctypes.c_char_p),
    "scb_sync": ([_vp], ctypes.c_int),
    "scb_stream": ([_vp], _vp),
    "scb_version": ([], ctypes.c_int),
    "scb_build_id": ([], ctypes.c_char_p),
    "scb_launch_count": ([_vp], ctypes.c_ulonglong),
    "scb_profile_begin": ([_vp], ctypes.c_int),
    "scb_profile_end": ([_vp, ctypes.c_char_p, ctypes.c_size_t], ctypes.c_int),
    "scb_measure_peak": ([_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double)], ctypes.c_int),
    "scb_host_alloc": ([ctypes.c_size_t, ctypes.POINTER(_vp)], ctypes.c_int),
    "scb_host_free": ([_vp], ctypes.c_int),
    "scb_dev_alloc": ([_vp, ctypes.c_size_t, ctypes.POINTER(_vp)], ctypes.c_int),
    "scb_dev_free": ([_vp, _vp], ctypes.c_int),
    "scb_copy_h2d": ([_vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_copy_d2h": ([_vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_modulus_load": ([_vp, ctypes.c_int, _vp, ctypes.c_int], ctypes.c_int),
    "scb_modulus_link_square": ([_vp, ctypes.c_int, ctypes.c_int], ctypes.c_int),
    "scb_modexp": ([_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_modexp_dev": ([_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_modmul": ([_vp, ctypes.c_int, _vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_modmul_dev": ([_vp, ctypes.c_int, _vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_modinv": ([_vp, ctypes.c_int, _vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_modinv_dev": ([_vp, ctypes.c_int, _vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_fixed_base_load": ([_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int], ctypes.c_int),
    "scb_fixed_base_exp": ([_vp, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_fixed_base_exp_dev": ([_vp, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_paillier_set_public": ([_vp, _vp, ctypes.c_int], ctypes.c_int),
    "scb_paillier_set_private": ([_vp, _vp, ctypes.c_int, _vp], ctypes.c_int),
    "scb_paillier_set_private_crt": ([_vp, _vp, _vp, ctypes.c_int], ctypes.c_int),
    "scb_set_option": ([_vp, ctypes.c_int, ctypes.c_int], ctypes.c_int),
    "scb_dgk_set_public": ([_vp, _vp, _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int], ctypes.c_int),
    "scb_dgk_set_private": ([_vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int], ctypes.c_int),
    "scb_debug_compile_exponent": ([_vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int)], ctypes.c_int),
    "scb_debug_resolve52": ([_vp, _vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_debug_resolve52_geo": ([_vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_rng_seed": ([_vp, _vp, _vp], ctypes.c_int),
    "scb_rng_below_dev": ([_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_rng_bits_dev": ([_vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_rng_coins_dev": ([_vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_rng_permutations_dev": ([_vp, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_rng_chacha_block_dev": ([_vp, _vp, _vp], ctypes.c_int),
    "scb_rng_fetch": ([_vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_paillier_encrypt": ([_vp, _vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_paillier_encrypt_dev": ([_vp, _vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_paillier_decrypt": ([_vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_paillier_decrypt_dev": ([_vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_dgk_is_zero": ([_vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_dgk_is_zero_dev": ([_vp, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_initiator_step1": ([_vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_initiator_step1_dev": ([_vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_keyholder_step4ab": ([_vp, _vp, ctypes.c_int, _vp, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_keyholder_step4ab_dev": ([_vp, _vp, ctypes.c_int, _vp, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_initiator_step4": ([_vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_initiator_step4_dev": ([_vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_keyholder_step4j_5": ([_vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_keyholder_step4j_5_dev": ([_vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_keyholder_step5": ([_vp, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_keyholder_step5_dev": ([_vp, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_initiator_step6_7": ([_vp, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
    "scb_initiator_step6_7_dev": ([_vp, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_size_t], ctypes.c_int),
}

_lib: Optional[ctypes.CDLL] = None


class NativeLibraryError(RuntimeError):
    pass


def declared_symbols():
    return sorted(_SIGNATURES)


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """Load libscb200.so and bind every declared entry point (no GPU needed for this).  ``SCB200_LIB`` selects another
    build of the same library (kernel-tuning experiments)."""
    global _lib
    if _lib is not None:
        return _lib
    if path is None:
        path = os.environ.get("SCB200_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise NativeLibraryError(
            f"{path} is missing: build it with `python -m protocols.secure_comparison_b200.build` "
            "(there is no CPU fallback for the product path)")
    lib = ctypes.CDLL(path)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def _ptr(arr) -> Optional[int]:
    """Pointer of a numpy array (host) or of anything with ``data_ptr()`` (torch tensor, device)."""
    if arr is None:
        return None
    if isinstance(arr, int):
        return arr
    if hasattr(arr, "data_ptr"):
        return arr.data_ptr()
    return arr.ctypes.data


def _host_u32(arr, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(arr, dtype=np.uint32)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a
