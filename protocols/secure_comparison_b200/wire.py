"""Packed wire format for the batched protocol messages.

The reference serialises one scheme-ciphertext object per value (``Serialization.pack`` of
``tno.mpc.communication``; message tuples at ``keyholder.py:109-113, 129-133``, ``initiator.py:110-112, 155-157``).
A batch of 65,536 comparisons would be >10^6 Python objects per message, so the batched drivers can send ONE
self-describing blob per protocol message instead: a fixed header and the raw little-endian limb array.

    blob = b"SCB1" | kind u8 | dtype u8 | ndim u8 | reserved u8 | shape (ndim x u32 LE) | payload (C order, LE)

``kind`` tags what the array is (Paillier ciphertexts, DGK ciphertext vectors, ...), so a receiver can reject a
message of the wrong type; ``pack_schemes`` / ``unpack_schemes`` carry the PUBLIC parts of both schemes for the
``schemes_session_k`` message (``keyholder.py:168-172``).  Any ``Communicator`` that moves ``bytes`` works.
"""
from __future__ import annotations

import struct
from types import SimpleNamespace
from typing import Tuple

import numpy as np

MAGIC = b"SCB1"
KIND_PAILLIER_CT, KIND_DGK_VECTOR, KIND_PAILLIER_TRIPLE, KIND_GENERIC = 1, 2, 3, 0
_DTYPES = {0: np.dtype("<u4"), 1: np.dtype("u1"), 2: np.dtype("<u8")}
_CODES = {v: k for k, v in _DTYPES.items()}


def pack(array: np.ndarray, kind: int = KIND_GENERIC) -> bytes:
    a = np.ascontiguousarray(array)
    dt = a.dtype.newbyteorder("<") if a.dtype.itemsize > 1 else a.dtype
    if dt not in _CODES:
        raise TypeError(f"unsupported dtype {a.dtype}")
    if a.ndim > 8:
        raise ValueError("too many dimensions")
    header = MAGIC + struct.pack("<BBBB", kind, _CODES[dt], a.ndim, 0) + struct.pack(f"<{a.ndim}I", *a.shape)
    return header + a.astype(dt, copy=False).tobytes()


def header(kind: int, dtype, shape) -> bytes:
    """The SCB1 header of an array message; ``len(header) + payload bytes`` is the blob size."""
    dt = np.dtype(dtype)
    dt = dt.newbyteorder("<") if dt.itemsize > 1 else dt
    if dt not in _CODES:
        raise TypeError(f"unsupported dtype {dtype}")
    if len(shape) > 8:
        raise ValueError("too many dimensions")
    return MAGIC + struct.pack("<BBBB", kind, _CODES[dt], len(shape), 0) + struct.pack(f"<{len(shape)}I", *shape)


def blob_with_payload_view(alloc, kind: int, shape, dtype=np.uint32):
    """Build a message IN PLACE: ``alloc(nbytes)`` returns a writable uint8 array (e.g. pinned host memory); the header is
    written at its start and the returned ``payload`` is a ``shape`` / ``dtype`` view of the rest, to be filled by the caller
    (a device-to-host copy lands directly in the message, no intermediate array).  Returns ``(blob_memoryview, payload)``."""
    head = header(kind, dtype, shape)
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dt.itemsize
    blob = alloc(len(head) + nbytes)
    blob[:len(head)] = np.frombuffer(head, dtype=np.uint8)
    payload = blob[len(head):len(head) + nbytes].view(dt).reshape(shape)
    return memoryview(blob)[:len(head) + nbytes], payload


def unpack(blob: bytes, expect_kind: int | None = None) -> np.ndarray:
    if len(blob) < 8 or blob[:4] != MAGIC:
        raise ValueError("not an SCB1 message")
    kind, code, ndim, _ = struct.unpack_from("<BBBB", blob, 4)
    if expect_kind is not None and kind != expect_kind:
        raise ValueError(f"message kind {kind}, expected {expect_kind}")
    if code not in _DTYPES or ndim > 8:
        raise ValueError("corrupt SCB1 header")
    shape = struct.unpack_from(f"<{ndim}I", blob, 8)
    dt = _DTYPES[code]
    off = 8 + 4 * ndim
    count = int(np.prod(shape)) if ndim else 1
    if len(blob) != off + count * dt.itemsize:
        raise ValueError("SCB1 payload length does not match the header")
    out = np.frombuffer(blob, dtype=dt, count=count, offset=off).reshape(shape)
    return out.astype(dt.newbyteorder("="), copy=False) if dt.itemsize > 1 else out


def _int_to_bytes(v: int) -> bytes:
    raw = int(v).to_bytes((int(v).bit_length() + 7) // 8 or 1, "little")
    return struct.pack("<I", len(raw)) + raw


def pack_schemes(scheme_paillier, scheme_dgk) -> bytes:
    """Public parts only: Paillier n; DGK n, g, h, u, t."""
    pk, dk = scheme_paillier.public_key, scheme_dgk.public_key
    return b"SCBK" + b"".join(_int_to_bytes(v) for v in (pk.n, dk.n, dk.g, dk.h, dk.u, dk.t))


def unpack_schemes(blob: bytes) -> Tuple[SimpleNamespace, SimpleNamespace]:
    if blob[:4] != b"SCBK":
        raise ValueError("not an SCBK key message")
    vals, off = [], 4
    for _ in range(6):
        (n,) = struct.unpack_from("<I", blob, off)
        off += 4
        vals.append(int.from_bytes(blob[off:off + n], "little"))
        off += n
    if off != len(blob):
        raise ValueError("trailing bytes in SCBK message")
    pn, dn, g, h, u, t = vals
    paillier = SimpleNamespace(public_key=SimpleNamespace(n=pn, n_squared=pn * pn, g=pn + 1), secret_key=None)
    dgk = SimpleNamespace(public_key=SimpleNamespace(n=dn, g=g, h=h, u=u, t=t), secret_key=None, randomness_bits=(5 * t) // 2)
    return paillier, dgk


def public_only(scheme_paillier, scheme_dgk) -> Tuple[SimpleNamespace, SimpleNamespace]:
    """Fresh holders carrying ONLY the public parts of both schemes (what ``unpack_schemes`` would rebuild on the other
    side).  The KeyHolder sends these -- never the scheme objects themselves, which carry the secret keys."""
    return unpack_schemes(pack_schemes(scheme_paillier, scheme_dgk))


def same_public_keys(a, b) -> bool:
    """Scheme equality as the Initiator needs it (initiator.py:192-203): the public parts agree."""
    fields = ("n", "g", "h", "u", "t")
    return all(getattr(a.public_key, f, None) == getattr(b.public_key, f, None) for f in fields)
