"""Agent state carried inside every Envelope (reference calfkit/models/state.py:18-141).

Canonical key order of `state` on the wire is the MRO-driven field order of `State`:
tool_calls, tool_results, uncommitted_message, message_history, final_output_parts,
temp_instructions, metadata, overrides  (SURVEY.md Appendix A)."""
import logging
from dataclasses import dataclass, field
from typing import Any

from pydantic import BaseModel, ConfigDict, Field

from calfkit.models.messages import ModelMessage, ModelRequest, ToolCallPart, ToolCallResult
from calfkit.models.node_schema import BaseToolNodeSchema
from calfkit.models.payload import ContentPart


class BaseAgentActivityState(BaseModel):
    model_config = ConfigDict(extra="ignore")


class OverridesState(BaseAgentActivityState):
    model_config = ConfigDict(extra="ignore")
    override_agent_tools: list[BaseToolNodeSchema] | None


class CoreMessageState(BaseAgentActivityState):
    model_config = ConfigDict(extra="ignore")
    uncommitted_message: ModelMessage | None = None
    message_history: list[ModelMessage] = Field(default_factory=list)
    final_output_parts: list[ContentPart] = Field(default_factory=list)
    temp_instructions: str | None = None

    def latest_tool_calls(self) -> list[ToolCallPart]:
        """Tool calls of the trailing run of responses (state.py:39-46)."""
        pending: list[ToolCallPart] = []
        for msg in reversed(self.message_history):
            if isinstance(msg, ModelRequest):
                break
            pending.extend(msg.tool_calls)
        return pending

    def stage_message(self, message: ModelMessage) -> None:
        self.uncommitted_message = message

    def commit_message_to_history(self) -> None:
        if self.uncommitted_message is None:
            msg = "The staged message(uncommitted_message) is None, can't be committed to history."
            logging.error(msg)
            raise RuntimeError(msg)
        self.message_history.append(self.uncommitted_message)
        self.uncommitted_message = None


class InFlightToolsState(BaseAgentActivityState):
    model_config = ConfigDict(extra="ignore")
    tool_calls: dict[str, ToolCallPart] = Field(default_factory=dict)
    tool_results: dict[str, ToolCallResult | Any] = Field(default_factory=dict)

    def add_tool_call(self, tool_call: ToolCallPart) -> None:
        self.tool_calls[tool_call.tool_call_id] = tool_call

    def add_tool_result(self, tool_call_id: str, tool_result: Any) -> None:
        self.tool_results[tool_call_id] = tool_result

    def get_tool_call(self, tool_call_id: str) -> ToolCallPart | None:
        return self.tool_calls.get(tool_call_id)

    def get_tool_result(self, tool_call_id: str) -> Any | None:
        return self.tool_results.get(tool_call_id)

    def all_call_ids_complete(self, *call_ids: str) -> bool:
        for call_id in call_ids:
            _ = self.tool_calls[call_id]
            if call_id not in self.tool_results:
                return False
        return True


class State(CoreMessageState, InFlightToolsState):
    model_config = ConfigDict(extra="ignore")
    metadata: Any = Field(default=None)
    overrides: OverridesState | None = None


@dataclass
class PendingToolBatch:
    """One in-flight parallel tool-call batch per correlation chain (state.py:127-141)."""
    expected_tool_call_ids: frozenset[str]
    base_state: State
    collected_results: dict[str, Any] = field(default_factory=dict)

    @property
    def is_complete(self) -> bool:
        return self.expected_tool_call_ids == frozenset(self.collected_results.keys())
