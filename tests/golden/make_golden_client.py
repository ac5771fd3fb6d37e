"""Generates tests/golden/client_invoke.json: the first envelope of a correlation chain as the UNMODIFIED reference's
Client builds and publishes it (calfkit/client/client.py invoke_node + calfkit/client/base.py:109-153 _invoke), for a
range of arguments (deps, temp_instructions, message_history, run_args, tool_overrides).  Build container only:

    python tests/golden/make_golden_client.py

FastStream (absent third-party dependency) is replaced by inert names; the connection handed to the client records what
is published.  The clock behind the message timestamps is frozen (the reference reads it through
calfkit/_vendor/pydantic_ai/_utils.now_utc)."""
import asyncio
import datetime as _dt
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
import tools_more  # noqa: E402

fs = sys.modules["faststream"]
_Sub = type("_Sub", (), {"__class_getitem__": classmethod(lambda c, i: c)})       # subscriptable placeholder for annotations
for name, attrs in (("faststream.message", {"StreamMessage": _Sub}), ("faststream.types", {"AsyncFuncAny": _Sub})):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
pkg = types.ModuleType("calfkit.client")
pkg.__path__ = [os.path.join(rh.REF_ROOT, "calfkit", "client")]
sys.modules["calfkit.client"] = pkg
sys.path.insert(0, rh.REF_ROOT)
client_mod = importlib.import_module("calfkit.client.client")
utils = importlib.import_module("calfkit._vendor.pydantic_ai._utils")
sys.path.remove(rh.REF_ROOT)
assert client_mod.__file__.startswith("/root/reference/")
tools_more.ToolContext = importlib.import_module("calfkit.models.tool_context").ToolContext

FIXED = _dt.datetime(2026, 1, 2, 3, 4, 5, 678901, tzinfo=_dt.timezone.utc)


class _FrozenDateTime(_dt.datetime):
    @classmethod
    def now(cls, tz=None):
        return FIXED


utils.datetime = _FrozenDateTime
messages = sys.modules["calfkit._vendor.pydantic_ai.messages"]


class Conn:                                   # what BaseClient needs from its broker connection
    _connection = True

    def __init__(self):
        self.published = []

    async def publish(self, envelope, topic, correlation_id, **kw):
        self.published.append({"topic": topic, "correlation_id": correlation_id, "key": kw.get("key"), "payload": envelope.model_dump_json()})


class Disp:
    def expect(self, correlation_id):
        return None


counter = [0]


def det():
    counter[0] += 1
    return f"{counter[0]:032x}"


rh.set_uuid_source(det)
history = [messages.ModelRequest(parts=[messages.UserPromptPart(content="earlier question")]),
           messages.ModelResponse(parts=[messages.TextPart(content="earlier answer")], timestamp=FIXED)]
overrides = [ref.agent_tool(tools_def.TOOLS["get_weather"]), ref.agent_tool(tools_more.with_defaults), ref.agent_tool(tools_more.google_multiline)]
ARGS = {
    "plain": dict(user_prompt="What's the weather in Tokyo?", topic="weather_agent.input"),
    "deps_instructions": dict(user_prompt="hi — ünïcode \"quoted\"", topic="planner.input", deps={"tenant": "t1", "n": [1, 2.5, None]},
                              temp_instructions="be brief", correlation_id="c" * 32, reply_topic="my.replies"),
    "history": dict(user_prompt="follow-up", topic="planner.input", message_history="HISTORY"),
    "run_args": dict(user_prompt="x", topic="node.input", run_args=["a", 2, {"k": None}]),
    "tool_overrides": dict(user_prompt="use the tools", topic="planner.input", tool_overrides="OVERRIDES"),
}
cases = []
for name, a in ARGS.items():
    counter[0] = 0
    kw = dict(a)
    if kw.get("message_history") == "HISTORY":
        kw["message_history"] = list(history)
    if kw.get("tool_overrides") == "OVERRIDES":
        kw["tool_overrides"] = list(overrides)
    conn = Conn()
    client = client_mod.Client(conn, "calf-client-reply-test", Disp())
    handle = asyncio.run(client.invoke_node(**kw))
    assert len(conn.published) == 1
    cases.append({"name": name, "args": a, "publish": conn.published[0],
                  "handle": {"correlation_id": handle.correlation_id, "topic": handle.topic, "reply_topic": handle.reply_topic}})
json.dump({"generated_by": "tests/golden/make_golden_client.py", "frozen_clock": FIXED.isoformat(), "id_source": "counter f'{n:032x}' from 1 per case",
           "cases": cases}, open(os.path.join(HERE, "client_invoke.json"), "w"), ensure_ascii=False, indent=1)
for c in cases:
    print(c["name"], c["publish"]["topic"], c["publish"]["key"], len(c["publish"]["payload"]))
