"""BaseClient (mirrors reference calfkit/client/base.py:27-172): connect(), broker, reply_topic,
_invoke (first Envelope of a correlation chain + reply future), close(), async context manager."""
from __future__ import annotations

import logging
import os
from collections.abc import Iterable, Sequence
from typing import Any

from typing_extensions import Self

from calfkit._ids import uuid7_hex
from calfkit.broker import KafkaBroker
from calfkit.client.deserialize import _UNSET
from calfkit.client.invocation_handle import InvocationHandle
from calfkit.client.middleware import ContextInjectionMiddleware
from calfkit.client.reply_dispatcher import _ReplyDispatcher
from calfkit.models import State
from calfkit.models.envelope import Envelope
from calfkit.models.session_context import CallFrame, CallFrameStack, Deps, SessionRunContext, WorkflowState
from calfkit.models.state import OverridesState

logger = logging.getLogger(__name__)


class BaseClient:
    def __init__(self, connection: KafkaBroker, reply_topic: str, dispatcher: _ReplyDispatcher) -> None:
        self._connection = connection
        self._reply_topic = reply_topic
        self._dispatcher = dispatcher

    @classmethod
    def connect(cls, server_urls: str | Iterable[str] | None = None, reply_topic: str | None = None,
                **broker_kwargs: Any) -> Self:
        if server_urls is None:
            server_urls = os.getenv("CALF_HOST_URL") or "localhost"
        client_id = uuid7_hex()
        if reply_topic is None:
            reply_topic = f"calf-client-reply-{client_id}"
        group_id = f"calf-client-reply-{client_id}"
        broker_connection = KafkaBroker(server_urls, middlewares=[ContextInjectionMiddleware], **broker_kwargs)
        dispatcher = _ReplyDispatcher()
        dispatcher.register(broker_connection, reply_topic, group_id)
        return cls(broker_connection, reply_topic, dispatcher)

    @property
    def broker(self) -> KafkaBroker:
        return self._connection

    @property
    def reply_topic(self) -> str:
        return self._reply_topic

    async def _invoke(self, topic: str, reply_topic: str, correlation_id: str, state: State,
                      overrides: OverridesState | None = None, run_args: Sequence[Any] | None = None,
                      deps: dict[str, Any] | None = None, output_type: type[Any] = _UNSET) -> InvocationHandle:
        future = self._dispatcher.expect(correlation_id)
        logger.debug("[%s] invoke topic=%s reply=%s", correlation_id[:8], topic, reply_topic)
        if not self._connection._connection:
            await self._connection.start()
        call_stack = CallFrameStack()
        call_stack.push(CallFrame(target_topic=topic, callback_topic=reply_topic, input_args=run_args, overrides=overrides))
        envelope = Envelope(internal_workflow_state=WorkflowState(call_stack=call_stack),
                            context=SessionRunContext(state=state, deps=Deps(correlation_id=correlation_id,
                                                                             provided_deps=deps or dict())))
        # the first hop is unkeyed, exactly like the reference (client/base.py:147)
        await self._connection.publish(envelope, topic=topic, correlation_id=correlation_id)
        return InvocationHandle(correlation_id=correlation_id, topic=topic, reply_topic=reply_topic, _future=future,
                                _output_type=output_type)

    async def close(self) -> None:
        self._dispatcher.close()
        await self._connection.stop()

    async def __aenter__(self) -> Self:
        return self

    async def __aexit__(self, *exc: object) -> None:
        await self.close()
