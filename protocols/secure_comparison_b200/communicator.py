"""Party-to-party transport as the batched drivers see it.

The engine does not move bytes between machines: ``Initiator`` / ``KeyHolder.perform_secure_comparison`` hand every
protocol message (one limb array, or one packed blob with ``wire_format=True``) to an object that offers the two
coroutines below -- the same duck-typed contract the reference's drivers use (reference ``communicator.py:8-11``; its
users pass a ``tno.mpc.communication.Pool``).  ``InMemoryCommunicator`` is a loop-back transport for single-process
use (tests, notebooks, benchmarks): both parties share one mailbox keyed by message id, like the reference's own test
double (``test/conftest.py:162-198``).
"""
from __future__ import annotations

import asyncio
from typing import Any, Dict, List, Optional, Protocol, Tuple, runtime_checkable


@runtime_checkable
class Communicator(Protocol):
    """Anything with awaitable ``send(party_id, message, msg_id)`` and ``recv(party_id, msg_id)``."""

    async def send(self, party_id: str, message: Any, msg_id: Optional[str] = None) -> None:
        """Deliver ``message`` to ``party_id`` under the identifier ``msg_id``."""

    async def recv(self, party_id: str, msg_id: Optional[str] = None) -> Any:
        """Wait for and return the message ``party_id`` sent under ``msg_id``."""


class InMemoryCommunicator:
    """Loop-back transport: every instance created from the same ``mailbox`` dict talks to the others.

    ``log`` (optional list) records ``(msg_id, python type name, payload bytes)`` of everything sent -- handy to
    check message ids and sizes.  ``recv`` polls the mailbox cooperatively (``poll_s`` between looks) and gives up
    with ``KeyError`` after ``timeout_s``.
    """

    def __init__(self, mailbox: Optional[Dict[str, Any]] = None, log: Optional[List[Tuple[str, str, int]]] = None,
                 poll_s: float = 0.0005, timeout_s: float = 600.0):
        self.mailbox = {} if mailbox is None else mailbox
        self.log = log
        self.poll_s, self.timeout_s = poll_s, timeout_s

    def peer(self) -> "InMemoryCommunicator":
        """A second endpoint on the same mailbox (give one to each party)."""
        return InMemoryCommunicator(self.mailbox, self.log, self.poll_s, self.timeout_s)

    async def send(self, party_id: str, message: Any, msg_id: Optional[str] = None) -> None:
        if self.log is not None:
            size = int(getattr(message, "nbytes", 0)) or (len(message) if isinstance(message, (bytes, bytearray)) else 0)
            self.log.append((str(msg_id), type(message).__name__, size))
        self.mailbox[str(msg_id)] = message

    async def recv(self, party_id: str, msg_id: Optional[str] = None) -> Any:
        key, waited = str(msg_id), 0.0
        while key not in self.mailbox:
            if waited >= self.timeout_s:
                raise KeyError(f"no message {key!r} from {party_id!r} after {self.timeout_s} s")
            await asyncio.sleep(self.poll_s)
            waited += self.poll_s
        return self.mailbox.pop(key)
