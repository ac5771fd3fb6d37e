"""Generates tests/golden/tool_returns.json: how the UNMODIFIED reference's tool node (calfkit/nodes/tool.py:37-86 through
BaseNodeDef.handler + _publish_action) turns a tool's Python return value into envelope bytes — dicts, lists, None, pydantic
models, dataclasses, datetimes, Decimal / UUID / Enum / bytes ...  Build container only:

    python tests/golden/make_golden_returns.py"""
import asyncio
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

ref = rh.load_reference()
import tools_more  # noqa: E402
import importlib  # noqa: E402
tools_more.ToolContext = importlib.import_module("calfkit.models.tool_context").ToolContext

Envelope = ref.Envelope
BASE = ('{"context":{"state":{"tool_calls":{"call_1":{"tool_name":"%s","args":%s,"tool_call_id":"call_1","id":null,"provider_name":null,'
        '"provider_details":null,"part_kind":"tool-call"}},"tool_results":{},"uncommitted_message":null,"message_history":[{"parts":[{"content":"hello","timestamp":"2026-01-01T00:00:00Z","part_kind":"user-prompt"}],'
        '"timestamp":null,"instructions":null,"kind":"request","run_id":null,"metadata":null}],"final_output_parts":[],'
        '"temp_instructions":null,"metadata":null,"overrides":null},"deps":{"correlation_id":"%s","provided_deps":{"tenant":"acme"}}},"internal_workflow_state":'
        '{"call_stack":{"_internal_list":[{"target_topic":"planner.input","callback_topic":"reply","input_args":null,"frame_id":"%s","overrides":null},'
        '{"target_topic":"tool.%s.input","callback_topic":"planner.input","input_args":["call_1","planner"],"frame_id":"%s","overrides":null}]},"metadata":null}}')
cases = []
for k, (name, (fn, args)) in enumerate(tools_more.RETURNS.items()):
    node = ref.agent_tool(fn)
    payload = (BASE % (name, json.dumps(args, separators=(",", ":"), ensure_ascii=False), f"{k:032x}", "a" * 32, name, "b" * 32)).encode()
    env = Envelope.model_validate_json(payload)
    payload = env.model_dump_json().encode()
    br = rh.CaptureBroker()
    err = None
    try:
        asyncio.run(node.handler(env, env.context.deps.correlation_id, br))
    except Exception as e:  # noqa: BLE001
        err = f"{type(e).__name__}: {e}"[:200]
    pubs = [{"topic": t, "key": kk.decode() if kk else None, "payload": p.decode()} for (t, kk, c, p) in br.published]
    cases.append({"name": name, "input": payload.decode(), "args": args, "raises": err, "publishes": pubs})
    print(name, err, pubs[0]["payload"][pubs[0]["payload"].index('"tool_results"'):][:160] if pubs else None)
json.dump({"generated_by": "tests/golden/make_golden_returns.py", "cases": cases}, open(os.path.join(HERE, "tool_returns.json"), "w"),
          ensure_ascii=False, indent=0)
