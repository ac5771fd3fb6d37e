"""Call stack + session context (reference calfkit/models/session_context.py:13-91)."""
from collections.abc import Sequence
from dataclasses import dataclass, field
from typing import Any, Generic

from pydantic import BaseModel, ConfigDict, Field

from calfkit._ids import uuid7_hex
from calfkit._types import DepsT, StackItemT, StateT
from calfkit.models.actions import _Call
from calfkit.models.state import OverridesState, State


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
