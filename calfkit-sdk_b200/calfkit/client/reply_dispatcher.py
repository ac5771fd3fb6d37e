"""correlation_id -> Future table (reference calfkit/client/reply_dispatcher.py:15-53)."""
from __future__ import annotations

import asyncio
import logging

from calfkit.models.envelope import Envelope

logger = logging.getLogger(__name__)


class _ReplyDispatcher:
    def __init__(self) -> None:
        self._pending: dict[str, asyncio.Future[Envelope]] = {}
        self._topic: str | None = None

    def register(self, broker, reply_topic: str, group_id: str) -> None:
        self._topic = reply_topic
        self._sub = broker.subscriber(reply_topic, group_id=group_id, auto_offset_reset="latest")
        self._sub(self._handle_reply)

    async def _handle_reply(self, envelope: Envelope, correlation_id: str) -> None:
        future = self._pending.pop(correlation_id, None)
        if future is None:
            logger.warning("[%s] reply received but no pending future", correlation_id[:8])
            return
        if future.cancelled():
            return
        future.set_result(envelope)

    async def drain(self, broker) -> int:
        """deliver every queued reply record (the batch worker's stand-in for FastStream's consume task)"""
        recs = broker.poll_batch((self._topic,), 1 << 16) if self._topic else []
        for r in recs:
            envelope = Envelope.model_validate_json(r.value)     # reply -> Python objects at the user-API edge
            corr = r.correlation_id or envelope.context.deps.correlation_id
            await self._handle_reply(envelope, corr)
        return len(recs)

    def expect(self, correlation_id: str) -> asyncio.Future[Envelope]:
        if correlation_id in self._pending:
            raise RuntimeError(f"Duplicate correlation_id: {correlation_id}")
        future: asyncio.Future[Envelope] = asyncio.get_running_loop().create_future()
        self._pending[correlation_id] = future
        future.add_done_callback(lambda _: self._pending.pop(correlation_id, None))
        return future

    def close(self) -> None:
        for future in self._pending.values():
            if not future.done():
                future.cancel()
        self._pending.clear()
