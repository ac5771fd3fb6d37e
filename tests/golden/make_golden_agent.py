"""Generates tests/golden/agent_run.json by running the UNMODIFIED reference's Agent node
(/root/reference/calfkit/nodes/agent.py:70-220 run() + nodes/base.py:70-164 handler / _publish_action) with a scripted
stand-in for the LLM loop (`_agent_loop.run`, the network-facing part).  Build container only:

    python tests/golden/make_golden_agent.py

Each case: inbound envelope bytes, what the scripted model answered, -> the action type, every publish (topic, key,
payload) in order incl. the handler-return publish to publish_topic, and the post-LLM state run() left in its context.
Frame ids come from a counter (f'{n:032x}'), tool-call ids are scripted."""
import asyncio
import importlib
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

ref = rh.load_reference()
import tools_def  # noqa: E402

pai = sys.modules["calfkit._vendor.pydantic_ai"]
tools_mod = sys.modules["calfkit._vendor.pydantic_ai.tools"]
messages = sys.modules["calfkit._vendor.pydantic_ai.messages"]


class _Loop:                                      # stands in for the vendored agent graph: never constructed for real
    def __init__(self, *a, **k):
        pass

    def __class_getitem__(cls, item):
        return cls


pai.Agent = _Loop
pai.DeferredToolRequests = tools_mod.DeferredToolRequests
for name, attrs in (("calfkit._vendor.pydantic_ai.output", {"OutputSpec": type("OutputSpec", (), {"__class_getitem__": classmethod(lambda c, i: c)})}),
                    ("calfkit._vendor.pydantic_ai.toolsets", {}),
                    ("calfkit._vendor.pydantic_ai.toolsets.external", {"ExternalToolset": lambda defs: ("external", defs)}),
                    ("calfkit.providers", {}), ("calfkit.providers.pydantic_ai", {}),
                    ("calfkit.providers.pydantic_ai.model_client", {"PydanticModelClient": object})):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
sys.path.insert(0, rh.REF_ROOT)
agent_mod = importlib.import_module("calfkit.nodes.agent")
sys.path.remove(rh.REF_ROOT)
assert agent_mod.__file__.startswith("/root/reference/")
models = importlib.import_module("calfkit.models")
session = importlib.import_module("calfkit.models.session_context")
Envelope, State = ref.Envelope, models.State

TOOLS = {name: ref.agent_tool(fn) for name, fn in tools_def.TOOLS.items()}
FIXED_TS = "2026-01-01T00:00:00Z"

# freeze the clock: every `timestamp` default_factory of the vendored message types is _utils.now_utc, which reads
# `datetime.now(tz=...)` through its module global — a frozen datetime class there makes this file reproducible byte for byte
import datetime as _dt  # noqa: E402


class _FrozenDatetime(_dt.datetime):
    @classmethod
    def now(cls, tz=None):
        return cls(2026, 1, 1, 0, 0, 1, tzinfo=tz)


sys.modules["calfkit._vendor.pydantic_ai._utils"].datetime = _FrozenDatetime


def inbound(corr: str, history_json: str = "[]", stage_prompt: str | None = "What's the weather in Paris and Tokyo?") -> bytes:
    unc = "null"
    if stage_prompt is not None:
        unc = ('{"parts":[{"content":%s,"timestamp":"%s","part_kind":"user-prompt"}],"timestamp":"%s","instructions":null,"kind":"request","run_id":null,"metadata":null}'
               % (json.dumps(stage_prompt), FIXED_TS, FIXED_TS))
    raw = ('{"context":{"state":{"tool_calls":{},"tool_results":{},"uncommitted_message":%s,"message_history":%s,"final_output_parts":[],'
           '"temp_instructions":null,"metadata":null,"overrides":null},"deps":{"correlation_id":"%s","provided_deps":{"tenant":"t1"}}},'
           '"internal_workflow_state":{"call_stack":{"_internal_list":[{"target_topic":"planner.input","callback_topic":"calf-client-reply-1",'
           '"input_args":null,"frame_id":"%s","overrides":null}]},"metadata":null}}' % (unc, history_json, corr, "a" * 32))
    return Envelope.model_validate_json(raw).model_dump_json().encode()


def response(calls=None, text=None):
    parts = []
    for (tool, args, cid) in calls or []:
        parts.append(messages.ToolCallPart(tool_name=tool, args=args, tool_call_id=cid))
    if text is not None:
        parts.append(messages.TextPart(content=text))
    import datetime
    return messages.ModelResponse(parts=parts, timestamp=datetime.datetime(2026, 1, 1, tzinfo=datetime.timezone.utc))


def run_case(name, payload: bytes, model_answer, *, sequential=False, publish_topic="planner.output"):
    counter = [0]

    def det():
        counter[0] += 1
        return f"{counter[0]:032x}"
    rh.set_uuid_source(det)
    node = agent_mod.BaseAgentNodeDef("planner", subscribe_topics="planner.input", publish_topic=publish_topic,
                                      tools=[TOOLS["get_weather"], TOOLS["get_temperature"], TOOLS["count_chars"]], model_client=None,
                                      sequential_only_mode=sequential)
    seen = {}

    class FakeLoop:
        async def run(self, **kw):
            seen["n_history"] = len(kw["message_history"])
            seen["deferred"] = sorted(kw["deferred_tool_results"].calls) if kw.get("deferred_tool_results") is not None else None
            resp = model_answer
            if isinstance(resp, str):
                r = response(text=resp)
                return types.SimpleNamespace(output=resp, new_messages=lambda: [r])
            r = response(calls=resp)
            return types.SimpleNamespace(output=tools_mod.DeferredToolRequests(calls=list(r.parts)), new_messages=lambda: [r])
    node._agent_loop = FakeLoop()
    post = {}
    orig = node._publish_action

    async def spy(output, envelope, correlation_id, broker):      # observe only: what run() returned and left in its context
        post["action"] = "list[Call]" if isinstance(output, list) else type(output).__name__
        st = output[0].state if isinstance(output, list) else getattr(output, "state", None)
        post["state"] = st.model_dump_json() if st is not None else None
        return await orig(output, envelope, correlation_id, broker)
    node._publish_action = spy
    env = Envelope.model_validate_json(payload)
    br = rh.CaptureBroker()
    err = None
    try:
        ret = asyncio.run(node.handler(env, env.context.deps.correlation_id, br))
    except Exception as e:  # noqa: BLE001
        err = type(e).__name__
        ret = None
    pubs = [{"topic": t, "key": k.decode() if k else None, "payload": p.decode()} for (t, k, c, p) in br.published]
    if ret is not None and publish_topic:
        pubs.append({"topic": publish_topic, "key": None, "payload": ret.model_dump_json()})
    return {"name": name, "input": payload.decode(), "sequential": sequential, "model_answer": model_answer, "llm_saw": seen,
            "action": post.get("action"), "post_llm_state": post.get("state"), "raises": err, "publishes": pubs,
            "pending_batch_ids": sorted(next(iter(node._pending_batches.values())).expected_tool_call_ids) if node._pending_batches else None}


cases = []
cases.append(run_case("parallel_3", inbound("1" * 32), [("get_weather", {"location": "Paris"}, "call_a"), ("get_temperature", {"location": "Tokyo"}, "call_b"), ("count_chars", {"text": "héllo"}, "call_c")]))
cases.append(run_case("single_call", inbound("2" * 32), [("get_weather", {"location": "Kraków"}, "call_a")]))
cases.append(run_case("sequential_3", inbound("3" * 32), [("get_weather", {"location": "Paris"}, "call_a"), ("get_temperature", {"location": "Tokyo"}, "call_b"), ("count_chars", {"text": "x"}, "call_c")], sequential=True))
cases.append(run_case("all_invalid_tailcall", inbound("4" * 32), [("nope", {}, "call_a"), ("nada", {"x": 1}, "call_b")]))
cases.append(run_case("one_invalid_two_valid", inbound("5" * 32), [("nope", {}, "call_a"), ("get_weather", {"location": "Paris"}, "call_b"), ("get_temperature", {"location": "Oslo"}, "call_c")]))
cases.append(run_case("final_text", inbound("6" * 32), "It's sunny in Paris"))
cases.append(run_case("no_publish_topic_parallel", inbound("7" * 32), [("get_weather", {"location": "Paris"}, "call_a"), ("get_temperature", {"location": "Tokyo"}, "call_b")], publish_topic=None))
# sequential mode, second pass: one of three results is in, no LLM call, next pending call goes out
seq = [c for c in cases if c["name"] == "sequential_3"][0]
st = json.loads(seq["post_llm_state"])
st["tool_results"]["call_a"] = {"return_value": "It's sunny in Paris", "content": None, "metadata": {"tool_call_id": "call_a"}, "kind": "tool-return"}
env2 = json.loads(seq["input"])
env2["context"]["state"] = st
payload2 = Envelope.model_validate_json(json.dumps(env2)).model_dump_json().encode()
cases.append(run_case("sequential_second_pass", payload2, "unused", sequential=True))
json.dump({"generated_by": "tests/golden/make_golden_agent.py", "frame_id_source": "counter: f'{n:032x}' starting at 1 per case", "cases": cases},
          open(os.path.join(HERE, "agent_run.json"), "w"), ensure_ascii=False, indent=0)
for c in cases:
    print(c["name"], c["action"], c["raises"], [(p["topic"], p["key"] is not None, len(p["payload"])) for p in c["publishes"]], c["llm_saw"])
