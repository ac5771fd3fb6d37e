from __future__ import annotations

import asyncio
from dataclasses import dataclass, field
from typing import Any, Generic

from calfkit._types import OutputT
from calfkit.client.deserialize import _UNSET, deserialize_to_node_result
from calfkit.client.node_result import NodeResult
from calfkit.models.envelope import Envelope


@dataclass
class InvocationHandle(Generic[OutputT]):
    """reference calfkit/client/invocation_handle.py:13-41"""
    correlation_id: str
    topic: str
    reply_topic: str
    _future: asyncio.Future[Envelope] = field(repr=False, compare=False)
    _output_type: type[Any] = field(default=_UNSET, repr=False, compare=False)

    async def result(self, timeout: float | None = None) -> NodeResult[OutputT]:
        if self._future is None:
            raise RuntimeError("This handle has no associated future — was the client's reply dispatcher configured?")
        envelope = await (asyncio.wait_for(self._future, timeout=timeout) if timeout is not None else self._future)
        return deserialize_to_node_result(envelope, self._output_type)
