"""Context handed to @agent_tool functions whose first parameter is a ToolContext
(reference calfkit/models/tool_context.py:7-20; vendored RunContext
calfkit/_vendor/pydantic_ai/_run_context.py:30-).  Only the attributes the reference's tool
node fills (nodes/tool.py:53-60) are kept; `messages` is decoded lazily from the record's
message_history span because most tools never read it."""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Generic

from calfkit._types import DepsT


@dataclasses.dataclass(kw_only=True)
class RunContext(Generic[DepsT]):
    deps: DepsT
    tool_call_id: str | None = None
    tool_name: str | None = None
    run_id: str | None = None
    retry: int = 0
    max_retries: int = 0
    _messages: Any = dataclasses.field(default=None, repr=False)
    _messages_loader: Callable[[], list] | None = dataclasses.field(default=None, repr=False)

    @property
    def messages(self) -> list:
        if self._messages is None and self._messages_loader is not None:
            self._messages = self._messages_loader()
        return self._messages if self._messages is not None else []


@dataclasses.dataclass(kw_only=True)
class ToolContext(RunContext[Any]):
    agent_name: str | None = None
