"""Wire types embedded in the Envelope JSON: LLM messages, parts, tool definitions, tool results.

These restate — field for field, in declaration order, with the same defaults and tags — the
subset of the vendored pydantic_ai types that the reference puts on the wire
(reference calfkit/_vendor/pydantic_ai/messages.py:112-153 SystemPromptPart, :676-699 ToolReturn,
:739-760 UserPromptPart, :811-913 tool-return parts, :918-960 RetryPromptPart, :1014-1041
ModelRequest, :1059-1147 TextPart/ThinkingPart, :1187-1283 tool-call parts, :1292-1345
ModelResponse; usage.py RequestUsage; tools.py:189-210,474-540 DeferredToolCallResult and
ToolDefinition; exceptions.py:35-70 ModelRetry).  Declaration order is the canonical key order of
the byte contract, so it must not change.

Out of scope: multi-modal user content (ImageUrl / AudioUrl /
DocumentUrl / VideoUrl / BinaryContent / CachePoint) and the response FilePart — the LLM provider
layer that produces them is not on the hot path.  Such payloads fail validation here and get
status CK_UNSUPPORTED on the device.
"""
from __future__ import annotations

from collections.abc import Sequence
from dataclasses import KW_ONLY, field
from datetime import datetime, timezone
from typing import Annotated, Any, Literal, Union

import pydantic
import pydantic_core
from pydantic import Discriminator, Tag
from pydantic.dataclasses import dataclass
from pydantic_core import core_schema

from calfkit._ids import generate_tool_call_id


def now_utc() -> datetime:
    return datetime.now(tz=timezone.utc)


FinishReason = Literal["stop", "length", "content_filter", "tool_call", "error"]
ToolKind = Literal["function", "output", "external", "unapproved"]


# ----------------------------------------------------------------------------- request parts
@dataclass(repr=False)
class SystemPromptPart:
    content: str
    _: KW_ONLY
    timestamp: datetime = field(default_factory=now_utc)
    dynamic_ref: str | None = None
    name: str | None = None
    part_kind: Literal["system-prompt"] = "system-prompt"


@dataclass(repr=False)
class UserPromptPart:
    content: str | Sequence[str]
    _: KW_ONLY
    timestamp: datetime = field(default_factory=now_utc)
    name: str | None = None
    part_kind: Literal["user-prompt"] = "user-prompt"


@dataclass(repr=False)
class ToolReturnPart:
    tool_name: str
    content: Any
    tool_call_id: str = field(default_factory=generate_tool_call_id)
    _: KW_ONLY
    metadata: Any = None
    timestamp: datetime = field(default_factory=now_utc)
    part_kind: Literal["tool-return"] = "tool-return"

    def model_response_str(self) -> str:
        if isinstance(self.content, str):
            return self.content
        return pydantic_core.to_json(self.content).decode()


@dataclass(repr=False)
class RetryPromptPart:
    content: list[pydantic_core.ErrorDetails] | str
    _: KW_ONLY
    tool_name: str | None = None
    tool_call_id: str = field(default_factory=generate_tool_call_id)
    timestamp: datetime = field(default_factory=now_utc)
    part_kind: Literal["retry-prompt"] = "retry-prompt"

    def model_response(self) -> str:
        if isinstance(self.content, str):
            if self.tool_name is None:
                return f"Validation feedback:\n{self.content}\n\nFix the errors and try again."
            return f"{self.content}\n\nFix the errors and try again."
        return pydantic_core.to_json(self.content).decode() + "\n\nFix the errors and try again."


ModelRequestPart = Annotated[
    Union[SystemPromptPart, UserPromptPart, ToolReturnPart, RetryPromptPart],
    Discriminator("part_kind"),
]


@dataclass(repr=False)
class ModelRequest:
    parts: Sequence[ModelRequestPart]
    _: KW_ONLY
    timestamp: datetime | None = None
    instructions: str | None = None
    kind: Literal["request"] = "request"
    run_id: str | None = None
    metadata: dict[str, Any] | None = None

    @classmethod
    def user_text_prompt(cls, user_prompt: str, *, instructions: str | None = None,
                         name: str | None = None) -> "ModelRequest":
        return cls(parts=[UserPromptPart(user_prompt, name=name)], instructions=instructions)


# ---------------------------------------------------------------------------- response parts
@dataclass(repr=False)
class TextPart:
    content: str
    _: KW_ONLY
    id: str | None = None
    provider_name: str | None = None
    provider_details: dict[str, Any] | None = None
    part_kind: Literal["text"] = "text"


@dataclass(repr=False)
class ThinkingPart:
    content: str
    _: KW_ONLY
    id: str | None = None
    signature: str | None = None
    provider_name: str | None = None
    provider_details: dict[str, Any] | None = None
    part_kind: Literal["thinking"] = "thinking"


@dataclass(repr=False)
class _BaseToolCallPart:
    tool_name: str
    args: str | dict[str, Any] | None = None
    tool_call_id: str = field(default_factory=generate_tool_call_id)
    _: KW_ONLY
    id: str | None = None
    provider_name: str | None = None
    provider_details: dict[str, Any] | None = None

    def args_as_dict(self) -> dict[str, Any]:
        """dict passthrough, JSON string parsed, falsy -> {} (reference messages.py:1229-1240)."""
        if not self.args:
            return {}
        if isinstance(self.args, dict):
            return self.args
        parsed = pydantic_core.from_json(self.args)
        assert isinstance(parsed, dict), "args should be a dict"
        return parsed

    def args_as_json_str(self) -> str:
        if not self.args:
            return "{}"
        if isinstance(self.args, str):
            return self.args
        return pydantic_core.to_json(self.args).decode()


@dataclass(repr=False)
class ToolCallPart(_BaseToolCallPart):
    _: KW_ONLY
    part_kind: Literal["tool-call"] = "tool-call"


@dataclass(repr=False)
class BuiltinToolCallPart(_BaseToolCallPart):
    _: KW_ONLY
    part_kind: Literal["builtin-tool-call"] = "builtin-tool-call"


@dataclass(repr=False)
class BuiltinToolReturnPart:
    tool_name: str
    content: Any
    tool_call_id: str = field(default_factory=generate_tool_call_id)
    _: KW_ONLY
    metadata: Any = None
    timestamp: datetime = field(default_factory=now_utc)
    provider_name: str | None = None
    provider_details: dict[str, Any] | None = None
    part_kind: Literal["builtin-tool-return"] = "builtin-tool-return"


ModelResponsePart = Annotated[
    Union[TextPart, ToolCallPart, BuiltinToolCallPart, BuiltinToolReturnPart, ThinkingPart],
    Discriminator("part_kind"),
]


@dataclass(repr=False, kw_only=True)
class RequestUsage:
    input_tokens: int = 0
    cache_write_tokens: int = 0
    cache_read_tokens: int = 0
    output_tokens: int = 0
    input_audio_tokens: int = 0
    cache_audio_read_tokens: int = 0
    output_audio_tokens: int = 0
    details: dict[str, int] = field(default_factory=dict)


@dataclass(repr=False)
class ModelResponse:
    parts: Sequence[ModelResponsePart]
    _: KW_ONLY
    usage: RequestUsage = field(default_factory=RequestUsage)
    model_name: str | None = None
    name: str | None = None
    timestamp: datetime = field(default_factory=now_utc)
    kind: Literal["response"] = "response"
    provider_name: str | None = None
    provider_url: str | None = None
    provider_details: Annotated[
        dict[str, Any] | None,
        pydantic.Field(validation_alias=pydantic.AliasChoices("provider_details", "vendor_details")),
    ] = None
    provider_response_id: Annotated[
        str | None,
        pydantic.Field(validation_alias=pydantic.AliasChoices("provider_response_id", "vendor_id")),
    ] = None
    finish_reason: FinishReason | None = None
    run_id: str | None = None
    metadata: dict[str, Any] | None = None

    @property
    def tool_calls(self) -> list[ToolCallPart]:
        return [p for p in self.parts if isinstance(p, ToolCallPart)]

    @property
    def text(self) -> str | None:
        texts = [p.content for p in self.parts if isinstance(p, TextPart)]
        return "\n\n".join(texts) if texts else None


ModelMessage = Annotated[Union[ModelRequest, ModelResponse], Discriminator("kind")]
ModelMessagesTypeAdapter = pydantic.TypeAdapter(list[ModelMessage])


# ------------------------------------------------------------------------------ tool results
@dataclass(repr=False)
class ToolReturn:
    return_value: Any
    _: KW_ONLY
    content: str | Sequence[str] | None = None
    metadata: Any = None
    kind: Literal["tool-return"] = "tool-return"


class ModelRetry(Exception):
    """Raised by a tool to ask the model to retry; serialises as {"message","kind":"model-retry"}."""

    message: str

    def __init__(self, message: str):
        self.message = message
        super().__init__(message)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, self.__class__) and other.message == self.message

    def __hash__(self) -> int:
        return hash((self.__class__, self.message))

    @classmethod
    def __get_pydantic_core_schema__(cls, _source: Any, _handler: Any) -> core_schema.CoreSchema:
        shape = core_schema.typed_dict_schema({
            "message": core_schema.typed_dict_field(core_schema.str_schema()),
            "kind": core_schema.typed_dict_field(core_schema.literal_schema(["model-retry"])),
        })
        return core_schema.no_info_after_validator_function(
            lambda d: ModelRetry(d["message"]),
            shape,
            serialization=core_schema.plain_serializer_function_ser_schema(
                lambda x: {"message": x.message, "kind": "model-retry"}, return_schema=shape),
        )


def _tool_result_tag(x: Any) -> str | None:
    if isinstance(x, dict):
        if "kind" in x:
            return x["kind"]
        if "part_kind" in x:
            return x["part_kind"]
        return None
    if hasattr(x, "kind"):
        return x.kind
    if hasattr(x, "part_kind"):
        return x.part_kind
    return None


ToolCallResult = Annotated[
    Union[
        Annotated[ToolReturn, Tag("tool-return")],
        Annotated[ModelRetry, Tag("model-retry")],
        Annotated[RetryPromptPart, Tag("retry-prompt")],
    ],
    Discriminator(_tool_result_tag),
]


@dataclass(repr=False, kw_only=True)
class ToolDefinition:
    name: str
    parameters_json_schema: dict[str, Any] = field(
        default_factory=lambda: {"type": "object", "properties": {}})
    description: str | None = None
    outer_typed_dict_key: str | None = None
    strict: bool | None = None
    sequential: bool = False
    kind: ToolKind = field(default="function")
    metadata: dict[str, Any] | None = None
    timeout: float | None = None
