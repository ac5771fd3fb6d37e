"""Reply envelope -> NodeResult (reference calfkit/client/deserialize.py:15-89): first DataPart.data,
else first TextPart.text; `output_type=str` forces text, any other type validates DataPart.data."""
from __future__ import annotations

from typing import Any

from pydantic import TypeAdapter

from calfkit.client.node_result import NodeResult
from calfkit.exceptions import DeserializationError
from calfkit.models import DataPart, TextPart
from calfkit.models.envelope import Envelope

_UNSET: Any = object()


def deserialize_to_node_result(envelope: Envelope, output_type: type[Any] = _UNSET) -> NodeResult[Any]:
    state = envelope.context.state
    return NodeResult(output=_extract_output(state.final_output_parts, output_type), output_parts=state.final_output_parts,
                      message_history=state.message_history, metadata=state.metadata,
                      correlation_id=envelope.context.deps.correlation_id)


def _extract_output(parts: list[Any], output_type: type[Any]) -> Any:
    if output_type is _UNSET:
        for part in parts:
            if isinstance(part, DataPart):
                return part.data
        for part in parts:
            if isinstance(part, TextPart):
                return part.text
        raise DeserializationError("No DataPart or TextPart found in final_output_parts; cannot auto-detect output.")
    if output_type is str:
        for part in parts:
            if isinstance(part, TextPart):
                return part.text
        raise DeserializationError("No TextPart found in final_output_parts; expected output_type=str.")
    for part in parts:
        if isinstance(part, DataPart):
            return TypeAdapter(output_type).validate_python(part.data)
    raise DeserializationError("No DataPart found in final_output_parts; expected output_type="
                               f"{getattr(output_type, '__name__', str(output_type))}.")
