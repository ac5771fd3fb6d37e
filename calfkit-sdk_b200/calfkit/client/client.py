"""TRANSCRIPTION: signatures and bodies restate the reference's calfkit/client/client.py (the user-facing API of the drop-in
boundary) — not original work.

Client (mirrors reference calfkit/client/client.py:20-218): invoke_node / execute_node."""
from __future__ import annotations

from collections.abc import Sequence
from typing import Any

from calfkit._ids import uuid7_hex
from calfkit.client.base import BaseClient
from calfkit.client.deserialize import _UNSET
from calfkit.client.invocation_handle import InvocationHandle
from calfkit.client.node_result import NodeResult
from calfkit.models import State
from calfkit.models.messages import ModelMessage, ModelRequest
from calfkit.models.node_schema import BaseToolNodeSchema
from calfkit.models.state import OverridesState


class Client(BaseClient):
    async def invoke_node(self, user_prompt: str, topic: str, *, tool_overrides: list[BaseToolNodeSchema] | None = None,
                          output_type: type[Any] = _UNSET, reply_topic: str | None = None, correlation_id: str | None = None,
                          temp_instructions: str | None = None, message_history: list[ModelMessage] | None = None,
                          run_args: Sequence[Any] | None = None, deps: dict[str, Any] | None = None) -> InvocationHandle[Any]:
        if correlation_id is None:
            correlation_id = uuid7_hex()
        if reply_topic is None:
            reply_topic = self._reply_topic
        state = State(message_history=message_history or list(), temp_instructions=temp_instructions)
        state.stage_message(ModelRequest.user_text_prompt(user_prompt))
        return await self._invoke(topic=topic, reply_topic=reply_topic, correlation_id=correlation_id, run_args=run_args,
                                  state=state, deps=deps, output_type=output_type,
                                  overrides=OverridesState(override_agent_tools=tool_overrides) if tool_overrides is not None else None)

    async def execute_node(self, user_prompt: str, topic: str, *, tool_overrides: list[BaseToolNodeSchema] | None = None,
                           output_type: type[Any] = _UNSET, reply_topic: str | None = None, correlation_id: str | None = None,
                           temp_instructions: str | None = None, message_history: list[ModelMessage] | None = None,
                           run_args: Sequence[Any] | None = None, deps: dict[str, Any] | None = None,
                           timeout: float | None = None) -> NodeResult[Any]:
        handle = await self.invoke_node(user_prompt, topic, tool_overrides=tool_overrides, output_type=output_type,
                                        reply_topic=reply_topic, correlation_id=correlation_id,
                                        temp_instructions=temp_instructions, message_history=message_history,
                                        run_args=run_args, deps=deps)
        return await handle.result(timeout=timeout)
