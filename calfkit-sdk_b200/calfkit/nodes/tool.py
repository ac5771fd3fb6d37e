"""@agent_tool nodes (mirrors reference calfkit/nodes/tool.py:19-95).

Two kinds of tool are distinguished at decoration time:
  * device tools  — `@agent_tool(device_template="It's sunny in {location}")`: the return value is a
    pure string template of string arguments; ToolNodeDef.run happens entirely in
    ck_plan_tool_kernel, no host round trip.
  * host tools    — any Python callable: the engine gathers each record's `args` JSON on the device,
    the callable runs on the host (the one thing a GPU cannot do: user Python), and its JSON result is
    spliced in on the device.  Argument decoding / result encoding of the *user's values* happens here
    with pydantic_core, exactly as reference nodes/tool.py:64 + messages.py:1229-1240 do.
"""
from __future__ import annotations

import inspect
import logging
from collections.abc import Awaitable, Callable
from dataclasses import dataclass, field
from typing import Any

import numpy as np
import pydantic
import pydantic_core
from typing_extensions import Self

from calfkit.broker import Record
from calfkit.engine._lib import CK_ACT_HOST_TOOL, CK_ACT_RAISES, CK_OK, COL, STATUS_NAMES
from calfkit.engine.batch import ToolTemplate
from calfkit.models import SessionRunContext, State, ToolContext
from calfkit.models.actions import NodeResult
from calfkit.models.messages import ModelMessagesTypeAdapter, ToolDefinition
from calfkit.models.node_schema import BaseToolNodeSchema
from calfkit.nodes.base import BaseNodeDef, pack_records

logger = logging.getLogger(__name__)


@dataclass
class Tool:
    """Minimal stand-in for pydantic_ai.Tool: the callable + its ToolDefinition (name, docstring,
    JSON schema of the keyword parameters).  reference: calfkit/_vendor/pydantic_ai/tools.py Tool."""
    function: Callable[..., Any]
    takes_ctx: bool = False
    tool_def: ToolDefinition = field(init=False)

    def __post_init__(self) -> None:
        sig = inspect.signature(self.function)
        params = list(sig.parameters.values())
        if params and (params[0].annotation is ToolContext or params[0].name == "ctx"):
            self.takes_ctx = True
            params = params[1:]
        fields = {p.name: ((p.annotation if p.annotation is not inspect.Parameter.empty else Any),
                           (... if p.default is inspect.Parameter.empty else p.default)) for p in params}
        model = pydantic.create_model(f"{self.function.__name__}_args", **fields)      # type: ignore[call-overload]
        schema = model.model_json_schema()
        schema.pop("title", None)
        for prop in schema.get("properties", {}).values():
            prop.pop("title", None)
        schema.setdefault("additionalProperties", False)
        self.tool_def = ToolDefinition(name=self.function.__name__, parameters_json_schema=schema,
                                       description=inspect.getdoc(self.function))


@dataclass
class BaseToolNodeDef(BaseToolNodeSchema, BaseNodeDef):
    _tool: Tool
    _template: ToolTemplate | None = None


class ToolNodeDef(BaseToolNodeDef):
    @classmethod
    def create_tool_node(cls, func: Callable[..., Any], subscribe_topics: str | list[str], publish_topic: str,
                         device_template: str | None = None) -> Self:
        if not isinstance(subscribe_topics, (list, tuple)):
            subscribe_topics = [subscribe_topics]
        tool = Tool(func)
        return cls(node_id=f"tool_{func.__name__}", tool_schema=tool.tool_def, subscribe_topics=subscribe_topics,
                   publish_topic=publish_topic, _tool=tool,
                   _template=ToolTemplate.from_format(device_template) if device_template else None)

    async def run(self, ctx: SessionRunContext, tool_call_id: str, source_node_name: str) -> NodeResult[State]:
        raise RuntimeError("ToolNodeDef.run is executed by the CUDA engine (ck_plan_tool_kernel); use process_batch")

    # ---- batch path ---------------------------------------------------------------------------------
    def configure_engine(self, engine) -> None:
        engine.set_tool_node(self.publish_topic, self._template)

    def _call_host(self, args_json: bytes, rec_bytes: memoryview, cols: np.ndarray, i: int) -> bytes:
        v = pydantic_core.from_json(args_json)
        if isinstance(v, str):
            v = pydantic_core.from_json(v)               # args_as_dict: JSON string -> dict
        kwargs = v or {}
        if self._tool.takes_ctx:
            def span(name: str) -> bytes:
                o, n = int(cols[COL[name + "_OFF"], i]), int(cols[COL[name + "_LEN"], i])
                return bytes(rec_bytes[o:o + n])
            deps = {"correlation_id": pydantic_core.from_json(b'"' + span("CORR") + b'"'),
                    "provided_deps": pydantic_core.from_json(span("PD"))}
            from calfkit.models import Deps
            hist = span("HIST")
            ctx = ToolContext(deps=Deps(**deps), agent_name=_jstr(span("ARG1")) if cols[COL["ARGKINDS"], i] & 2 else None,
                              tool_call_id=_jstr(span("ARG0")), tool_name=_jstr(span("TNAME")), run_id=deps["correlation_id"],
                              _messages_loader=lambda: ModelMessagesTypeAdapter.validate_json(hist))
            result = self._tool.function(ctx, **kwargs)
        else:
            result = self._tool.function(**kwargs)
        if inspect.isawaitable(result):
            raise TypeError("async tools must be awaited by the caller: wrap them with asyncio.run or use a sync tool")
        return pydantic_core.to_json(result)

    def process_batch(self, engine, records: list[Record]) -> list[Record]:
        data, offsets = pack_records(records)
        engine.submit(data, offsets)
        if self._template is not None:
            engine.tool_plan()
        else:
            blob, off, ln = engine.tool_args()
            cols = engine.columns()
            mv = memoryview(data)
            results: list[bytes] = []
            for i in range(len(records)):
                if cols[COL["ACTION"], i] == CK_ACT_HOST_TOOL:
                    rec = mv[offsets[i]:offsets[i + 1]]
                    results.append(self._call_host(blob[off[i]:off[i] + ln[i]].tobytes(), rec, cols, i))
                else:
                    results.append(b"")
            aux_off = np.zeros(len(records) + 1, dtype=np.int64)
            np.cumsum([len(r) for r in results], out=aux_off[1:])
            engine.tool_plan(np.frombuffer(b"".join(results) or b"\0", dtype=np.uint8), aux_off)
        out = engine.fetch()
        for i in np.nonzero(out.cols[COL["STATUS"]] != CK_OK)[0]:
            logger.error("record %d rejected: %s at byte %d", i, STATUS_NAMES[int(out.cols[COL["STATUS"], i])],
                         int(out.cols[COL["ERR"], i]))
        for i in np.nonzero(out.cols[COL["ACTION"]] == CK_ACT_RAISES)[0]:
            logger.error("record %d: the reference handler would raise here (bad input_args / empty call stack)", i)
        produced = []
        for p in out.publishes():
            corr = records[p.record].correlation_id
            if corr is None and p.key is not None:
                corr = p.key.decode()
            produced.append(Record(p.topic, p.payload, p.key, corr))
        return produced


def _jstr(raw: bytes) -> str:
    return pydantic_core.from_json(b'"' + raw + b'"')


def agent_tool(func: Callable[..., Any] | Callable[..., Awaitable[Any]] | None = None, *,
               device_template: str | None = None):
    """Decorator turning a function into a deployable tool node (reference nodes/tool.py:89-95):
    subscribes `tool.<name>.input`, publishes `tool.<name>.output`, node_id `tool_<name>`."""
    def make(f):
        return ToolNodeDef.create_tool_node(func=f, subscribe_topics=f"tool.{f.__name__}.input",
                                            publish_topic=f"tool.{f.__name__}.output", device_template=device_template)
    return make(func) if func is not None else make
