from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Generic

from calfkit._types import OutputT
from calfkit.models import ContentPart
from calfkit.models.messages import ModelMessage


@dataclass(frozen=True)
class NodeResult(Generic[OutputT]):
    """Client-facing projection of a reply envelope (reference calfkit/client/node_result.py:11-32)."""
    output: OutputT
    output_parts: list[ContentPart]
    message_history: list[ModelMessage]
    metadata: Any
    correlation_id: str
