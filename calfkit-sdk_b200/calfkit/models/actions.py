"""What a node's run() may return.  TRANSCRIPTION, not original work: these dataclasses restate the declarations of the
reference's calfkit/models/actions.py (:26-49 _Call / Call / TailCall / ReturnCall, :68-70 Silent, :73-80 NodeResult) because
their names and fields are the user-facing API of the drop-in boundary.  The reference's other action
types (Reply, Delegate, Sequential, Emit, Parallel, :14-66) are not produced by any node on this path and are not mirrored.
The batch engine maps each kept type onto a call-stack operation: Call = push, ReturnCall = pop, TailCall = pop+push,
list[Call] = fan-out of pushes, Silent = nothing (reference nodes/base.py:70-147)."""
from collections.abc import Sequence
from dataclasses import dataclass
from typing import Any, Generic

from typing_extensions import TypeAliasType, TypeVar

from calfkit._types import StateT


@dataclass(init=False)
class _Call(Generic[StateT]):
    target_topic: str
    state: StateT
    input_args: Sequence[Any] | None

    def __init__(self, target_topic: str, state: StateT, *input_args: Any):
        self.target_topic = target_topic
        self.state = state
        self.input_args = input_args or None  # () -> None, else a tuple (actions.py:66)


class Call(Generic[StateT], _Call[StateT]):
    """Call another node; the target calls back with the state when done."""


class TailCall(Generic[StateT], _Call[StateT]):
    """Call another node; the callee inherits this frame's callback."""


@dataclass
class ReturnCall(Generic[StateT]):
    state: StateT


@dataclass
class Silent:
    """No publish; end of this event stream."""


_T = TypeVar("_T")
NodeResult = TypeAliasType(
    "NodeResult", Silent | Call[_T] | list[Call[_T]] | ReturnCall[_T] | TailCall[_T], type_params=(_T,))
