"""TRANSCRIPTION of the reference's model declarations (field names, order, defaults, config = the byte contract), with the
helper bodies the reference gives them (commit_message_to_history, Stack.push/pop/peek, invoke_frame / unwind_frame) restated
as they are — not original work; the product's own work is the CUDA path that implements this schema.

Wire-format models of the hot path in ONE module, in dependency order: content parts -> node routing schema ->
agent state -> call stack / session context -> Envelope.  Field names, order, defaults and config are the byte
contract (`Envelope.model_dump_json()`); the reference spreads the same declarations over
calfkit/models/{payload,node_schema,state,session_context,envelope}.py — those module paths still exist here and
re-export from this file.  The CUDA walker (csrc/ck_walk.cuh) and canonicaliser (csrc/ck_canon.cuh) implement
exactly the schema declared below; tests/test_api_surface.py diffs its JSON schema against the reference's."""
import logging
from collections.abc import Sequence
from dataclasses import KW_ONLY, dataclass, field
from typing import Annotated, Any, Generic, Literal, Union

from pydantic import BaseModel, ConfigDict, Discriminator, Field

from calfkit._ids import uuid7_hex
from calfkit._types import DepsT, StackItemT, StateT
from calfkit.models.actions import _Call
from calfkit.models.messages import ModelMessage, ModelRequest, ToolCallResult, ToolDefinition
from calfkit.models.messages import ToolCallPart as _ModelToolCallPart      # the LLM-side tool call (state.tool_calls)


# ----------------------------------------------------------------------------------------------------
# final-output content parts  (reference models/payload.py:6-35; `kind` is the first key of every part)
# ----------------------------------------------------------------------------------------------------
class TextPart(BaseModel):
    kind: Literal["text"] = "text"
    text: str
    metadata: dict[str, Any] | None = None


class FilePart(BaseModel):
    kind: Literal["file"] = "file"
    media_type: str
    uri: str | None = None
    data: str | None = None
    metadata: dict[str, Any] | None = None


class DataPart(BaseModel):
    kind: Literal["data"] = "data"
    data: dict[str, Any] | list[Any] | Any
    # model_dump_json() emits "schema_" (no by_alias) — SURVEY.md Appendix C item 2.
    schema_: dict[str, Any] | None = Field(default=None, alias="schema")
    metadata: dict[str, Any] | None = None


class ToolCallPart(BaseModel):
    kind: Literal["tool"] = "tool"
    tool_call_id: str
    kwargs: dict[str, Any]
    tool_name: str
    metadata: dict[str, Any] | None = None


ContentPart = Annotated[Union[TextPart, FilePart, DataPart, ToolCallPart], Discriminator("kind")]


# ----------------------------------------------------------------------------------------------------
# node routing data  (reference models/node_schema.py:6-21)
# ----------------------------------------------------------------------------------------------------
@dataclass
class BaseNodeSchema:
    _: KW_ONLY
    node_id: str
    subscribe_topics: list[str]
    publish_topic: str | None

    def __post_init__(self) -> None:
        if not isinstance(self.subscribe_topics, (list, tuple)):
            self.subscribe_topics = [self.subscribe_topics]


@dataclass
class BaseToolNodeSchema(BaseNodeSchema):
    _: KW_ONLY
    tool_schema: ToolDefinition


# ----------------------------------------------------------------------------------------------------
# agent state  (reference models/state.py:18-141; wire order: tool_calls, tool_results, uncommitted_message,
# message_history, final_output_parts, temp_instructions, metadata, overrides)
# ----------------------------------------------------------------------------------------------------
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

    def latest_tool_calls(self) -> list[_ModelToolCallPart]:
        """Tool calls of the trailing run of responses (state.py:39-46)."""
        pending: list[_ModelToolCallPart] = []
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
    tool_calls: dict[str, _ModelToolCallPart] = Field(default_factory=dict)
    tool_results: dict[str, ToolCallResult | Any] = Field(default_factory=dict)

    def add_tool_call(self, tool_call: _ModelToolCallPart) -> None:
        self.tool_calls[tool_call.tool_call_id] = tool_call

    def add_tool_result(self, tool_call_id: str, tool_result: Any) -> None:
        self.tool_results[tool_call_id] = tool_result

    def get_tool_call(self, tool_call_id: str) -> _ModelToolCallPart | None:
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


# ----------------------------------------------------------------------------------------------------
# call stack + session context  (reference models/session_context.py:13-91)
# ----------------------------------------------------------------------------------------------------
@dataclass
class Stack(Generic[StackItemT]):
    _internal_list: list[StackItemT] = field(default_factory=list)

    def push(self, item: StackItemT) -> None:
        self._internal_list.append(item)

    def pop(self) -> StackItemT:
        try:
            return self._internal_list.pop()
        except Exception as e:
            raise Exception("An exception occurred when popping from execution stack") from e

    def peek(self) -> StackItemT:
        try:
            return self._internal_list[-1]
        except Exception as e:
            raise Exception("An exception occurred when peeking from execution stack") from e


@dataclass(frozen=True)
class CallFrame:
    target_topic: str
    callback_topic: str
    input_args: Sequence[Any] | None = field(default=None)
    frame_id: str = field(default_factory=uuid7_hex)
    overrides: OverridesState | None = field(default=None)


CallFrameStack = Stack[CallFrame]


class WorkflowState(BaseModel):
    model_config = ConfigDict(extra="ignore")
    call_stack: CallFrameStack
    metadata: Any = Field(default=None)

    @property
    def current_frame(self) -> CallFrame:
        return self.call_stack.peek()

    def unwind_frame(self) -> CallFrame:
        return self.call_stack.pop()

    def invoke_frame(self, call: _Call, callback_topic: str) -> None:
        if call.target_topic is None:
            raise Exception("")
        self.call_stack.push(CallFrame(target_topic=call.target_topic, callback_topic=callback_topic,
                                       input_args=call.input_args))


class Deps(BaseModel):
    model_config = ConfigDict(extra="ignore", frozen=True)
    correlation_id: str
    provided_deps: dict[str, Any] = Field(description="user-provided agent dependencies")


class BaseSessionRunContext(BaseModel, Generic[StateT, DepsT]):
    state: StateT
    deps: DepsT


SessionRunContext = BaseSessionRunContext[State, Deps]


# ----------------------------------------------------------------------------------------------------
# the envelope  (reference models/envelope.py:9-17)
# ----------------------------------------------------------------------------------------------------
class Envelope(BaseModel):
    context: SessionRunContext
    internal_workflow_state: WorkflowState = Field(description="framework-level workflow state")

