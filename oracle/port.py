"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's hot path.

May be imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, and there only as the checker or the timed CPU baseline.  The product
(calfkit-sdk_b200/) never imports this module and has no CPU fallback.

What it restates, with the reference lines each function follows:

  decode / encode     Envelope.model_validate_json / model_dump_json           models/envelope.py:9-17
                      (arithmetic = pydantic-core 2.46.4 via pydantic 2.13.4, third-party, present
                      in the image and unpinned by the reference: pyproject.toml:24 `pydantic>=2.12.5`)
  prepare_context     deep copy + frame overrides                              nodes/base.py:64-68
  handler             input_args dispatch rule                                 nodes/base.py:149-164
  tool_run            ToolNodeDef.run                                          nodes/tool.py:37-86
  publish_action      action -> [(topic, key, correlation_id, Envelope)]       nodes/base.py:70-147
  worker_publish      handler return value -> node.publish_topic               worker/worker.py:52-53
  agent_fanout        pending tool calls -> list[Call]                         nodes/agent.py:177-211
  aggregate           _parallel_state_aggregation                              nodes/agent.py:57-68

Pinning: tests/golden/*.json were produced by the UNMODIFIED reference (oracle/ref_harness.py +
tests/golden/make_golden.py, run in the build container where /root/reference exists);
tests/test_oracle.py checks this port against every one of them.  The wire models it uses are the
host-side mirror in calfkit-sdk_b200/calfkit/models (their JSON schema is diffed against the
reference's in make_golden.py).
"""
from __future__ import annotations

import inspect
import os
import sys
from dataclasses import dataclass
from typing import Any, Callable

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "calfkit-sdk_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from calfkit.models import (Call, Envelope, ReturnCall, SessionRunContext, Silent, State, TailCall)  # noqa: E402
from calfkit.models.messages import ToolReturn  # noqa: E402
from calfkit.models.state import PendingToolBatch  # noqa: E402

Published = tuple[str, bytes | None, str, bytes]   # (topic, key, correlation_id, payload)


def decode(payload: bytes) -> Envelope:
    return Envelope.model_validate_json(payload)


def encode(envelope: Envelope) -> bytes:
    return envelope.model_dump_json().encode()


@dataclass
class ToolNode:
    """Routing data + callable of one @agent_tool node (nodes/tool.py:24-35,89-95)."""
    func: Callable[..., Any]
    node_id: str
    subscribe_topics: list[str]
    publish_topic: str | None
    takes_ctx: bool = False

    @classmethod
    def of(cls, func: Callable[..., Any]) -> "ToolNode":
        params = list(inspect.signature(func).parameters.values())
        takes_ctx = bool(params) and params[0].name == "ctx"
        return cls(func, f"tool_{func.__name__}", [f"tool.{func.__name__}.input"], f"tool.{func.__name__}.output", takes_ctx)


def prepare_context(envelope: Envelope) -> SessionRunContext:
    ctx = envelope.context.model_copy(deep=True)
    if envelope.internal_workflow_state.current_frame.overrides:
        ctx.state.overrides = envelope.internal_workflow_state.current_frame.overrides
    return ctx


def tool_run(node: ToolNode, ctx: SessionRunContext, tool_call_id: str, source_node_name: str):
    part = ctx.state.get_tool_call(tool_call_id)
    if part is None:
        return Silent()
    kwargs = part.args_as_dict()
    if node.takes_ctx:
        from calfkit.models import ToolContext
        tctx = ToolContext(deps=ctx.deps, agent_name=source_node_name, tool_call_id=part.tool_call_id,
                           tool_name=part.tool_name, run_id=ctx.deps.correlation_id,
                           _messages=ctx.state.message_history)
        result = node.func(tctx, **kwargs)
    else:
        result = node.func(**kwargs)
    ctx.state.add_tool_result(part.tool_call_id,
                              ToolReturn(return_value=result, metadata={"tool_call_id": part.tool_call_id}))
    return ReturnCall(state=ctx.state)


def publish_action(subscribe_topic0: str, output: Any, envelope: Envelope, correlation_id: str
                   ) -> tuple[list[tuple[str, bytes, str, Envelope]], Envelope]:
    """Returns ([(topic, key, correlation_id, envelope-to-publish)], handler return value)."""
    key = correlation_id.encode()
    pubs: list[tuple[str, bytes, str, Envelope]] = []
    if isinstance(output, list) and all(isinstance(c, Call) for c in output):
        for call in output:
            wf = envelope.internal_workflow_state.model_copy(deep=True)
            wf.invoke_frame(call, subscribe_topic0)
            env = Envelope(context=SessionRunContext(state=call.state, deps=envelope.context.deps),
                           internal_workflow_state=wf)
            pubs.append((wf.current_frame.target_topic, key, correlation_id, env))
        return pubs, envelope
    if isinstance(output, Call):
        envelope.internal_workflow_state.invoke_frame(output, subscribe_topic0)
        env = Envelope(context=SessionRunContext(state=output.state, deps=envelope.context.deps),
                       internal_workflow_state=envelope.internal_workflow_state)
        pubs.append((envelope.internal_workflow_state.current_frame.target_topic, key, correlation_id, env))
        return pubs, env
    if isinstance(output, ReturnCall):
        fr = envelope.internal_workflow_state.unwind_frame()
        env = Envelope(context=SessionRunContext(state=output.state, deps=envelope.context.deps),
                       internal_workflow_state=envelope.internal_workflow_state)
        pubs.append((fr.callback_topic, key, correlation_id, env))
        return pubs, env
    if isinstance(output, TailCall):
        fr = envelope.internal_workflow_state.unwind_frame()
        envelope.internal_workflow_state.invoke_frame(output, fr.callback_topic)
        env = Envelope(context=SessionRunContext(state=output.state, deps=envelope.context.deps),
                       internal_workflow_state=envelope.internal_workflow_state)
        pubs.append((envelope.internal_workflow_state.current_frame.target_topic, key, correlation_id, env))
        return pubs, env
    return pubs, envelope   # Silent / unknown: nothing published, input envelope returned


def tool_node_event(node: ToolNode, payload: bytes, correlation_id: str | None = None) -> list[Published]:
    """One inbound record through a tool node, end to end: decode -> handler -> publishes
    (callback publish, then the handler-return publish to node.publish_topic)."""
    envelope = decode(payload)
    if correlation_id is None:
        correlation_id = envelope.context.deps.correlation_id   # header value == deps value on this path
    ctx = prepare_context(envelope)
    args = envelope.internal_workflow_state.current_frame.input_args
    if args is not None:
        output = tool_run(node, ctx, *args)
    else:
        raise TypeError("ToolNodeDef.run() missing tool_call_id/source_node_name")
    pubs, returned = publish_action(node.subscribe_topics[0], output, envelope, correlation_id)
    out: list[Published] = [(t, k, c, encode(e)) for (t, k, c, e) in pubs]
    if node.publish_topic:
        out.append((node.publish_topic, None, correlation_id, encode(returned)))
    return out


def agent_fanout(agent_name: str, subscribe_topic0: str, publish_topic: str | None,
                 registry: dict[str, str], payload: bytes, sequential: bool = False, inbound: bytes | None = None) -> list[Published]:
    """The data-parallel half of Agent.run after the (out-of-scope) LLM step (agent.py:132-211): `payload` is the
    envelope with the post-LLM state (tool calls added, invalid ones already answered with a RetryPromptPart);
    pending ones become Call(registry[name], state copy, id, agent_name).  `inbound`: the record the handler was
    invoked with — for list[Call] the handler's return value, and so the publish_topic payload, is THAT envelope,
    untouched (run() works on prepare_context's deep copy, nodes/base.py:64-68,88); without it the post-LLM
    envelope stands in (what the device plan does: the host layer substitutes, calfkit/nodes/agent.py)."""
    envelope = decode(payload)
    correlation_id = envelope.context.deps.correlation_id
    ctx = prepare_context(envelope)
    latest = ctx.state.latest_tool_calls()                      # state.py:39-46
    pending = [tc for tc in latest if tc.tool_call_id not in ctx.state.tool_results]
    if not pending:
        output: Any = TailCall(subscribe_topic0, ctx.state)
    elif sequential or len(pending) == 1:
        tc = pending[0]
        output = Call(registry[tc.tool_name], ctx.state, tc.tool_call_id, agent_name)
    else:
        output = [Call(registry[tc.tool_name], ctx.state.model_copy(deep=True), tc.tool_call_id, agent_name)
                  for tc in pending]
    pubs, returned = publish_action(subscribe_topic0, output, envelope, correlation_id)
    if isinstance(output, list) and inbound is not None:
        returned = decode(inbound)
    out: list[Published] = [(t, k, c, encode(e)) for (t, k, c, e) in pubs]
    if publish_topic:
        out.append((publish_topic, None, correlation_id, encode(returned)))
    return out


def aggregate(batches: dict[str, PendingToolBatch], state: State, correlation_id: str) -> State | None:
    """_parallel_state_aggregation: returns the merged base state when the batch completes,
    None while it is still incomplete (caller returns Silent), `state` if no batch is pending."""
    batch = batches.get(correlation_id)
    if batch is None:
        return state
    for cid in batch.expected_tool_call_ids:
        if cid not in batch.collected_results and cid in state.tool_results:
            batch.collected_results[cid] = state.tool_results[cid]
    if batch.is_complete:
        for cid, res in batch.collected_results.items():
            batch.base_state.add_tool_result(cid, res)
        del batches[correlation_id]
        return batch.base_state
    return None


# ---------------------------------------------------------------------------------------------------
# client reply projection (reference calfkit/client/deserialize.py:15-89)
# ---------------------------------------------------------------------------------------------------
REPLY_UNSET: Any = object()


def reply_output(payload: bytes, output_type: Any = REPLY_UNSET) -> tuple[str, bytes]:
    """-> (correlation_id, JSON of NodeResult.output); raises what the reference raises (DeserializationError when
    the wanted part is missing, pydantic ValidationError when a typed output does not validate)."""
    from pydantic import TypeAdapter
    import pydantic_core
    from calfkit.exceptions import DeserializationError
    from calfkit.models import DataPart, TextPart
    envelope = decode(payload)
    parts = envelope.context.state.final_output_parts               # deserialize.py:38-43
    corr = envelope.context.deps.correlation_id
    if output_type is REPLY_UNSET:                                     # _extract_auto, deserialize.py:64-72
        for p in parts:
            if isinstance(p, DataPart):
                return corr, pydantic_core.to_json(p.data)
        for p in parts:
            if isinstance(p, TextPart):
                return corr, pydantic_core.to_json(p.text)
        raise DeserializationError("No DataPart or TextPart found in final_output_parts; cannot auto-detect output.")
    if output_type is str:                                           # _extract_text, deserialize.py:75-80
        for p in parts:
            if isinstance(p, TextPart):
                return corr, pydantic_core.to_json(p.text)
        raise DeserializationError("No TextPart found in final_output_parts; expected output_type=str.")
    for p in parts:                                                  # _extract_data, deserialize.py:83-89
        if isinstance(p, DataPart):
            return corr, pydantic_core.to_json(TypeAdapter(output_type).validate_python(p.data))
    raise DeserializationError("No DataPart found in final_output_parts; expected output_type="
                               f"{getattr(output_type, '__name__', str(output_type))}.")
