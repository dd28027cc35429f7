"""Transport protocol of the two parties -- same shape as the reference's ``communicator.py:8-11``.
Party-to-party transport is out of scope for the engine; any object with these two coroutines works
(the reference uses ``tno.mpc.communication.Pool``)."""
from __future__ import annotations

from typing import Any, Protocol


class Communicator(Protocol):
    async def send(self, party_id: str, message: Any, msg_id: str | None = None) -> None: ...

    async def recv(self, party_id: str, msg_id: str | None = None) -> Any: ...
