from typing import Any, TypeVar

StateT = TypeVar("StateT")
DepsT = TypeVar("DepsT")
StackItemT = TypeVar("StackItemT")
OutputT = TypeVar("OutputT")
AgentOutputT = TypeVar("AgentOutputT")
AnyT = Any
