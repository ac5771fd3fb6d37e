"""Final-output content parts (reference calfkit/models/payload.py:6-35).  `kind` is the first
key of every part, which is what lets the device matcher dispatch on it without look-ahead."""
from typing import Annotated, Any, Literal, Union

from pydantic import BaseModel, Discriminator, Field


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
