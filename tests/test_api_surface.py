"""The drop-in surface of the hot path (SURVEY.md §8b "upward" face): names, signatures and the
error behaviour the reference pins.  CPU only: nothing here launches a kernel."""
import inspect

import pytest


def test_public_names():
    import calfkit
    for name in ["Client", "InvocationHandle", "NodeResult", "ToolContext", "Agent", "BaseNodeDef", "NodeDef", "ToolNodeDef",
                 "agent_tool", "Worker"]:          # reference calfkit/__init__.py:10-30 (providers are out of scope)
        assert hasattr(calfkit, name), name


def test_worker_signature_and_guard():
    from calfkit import Client, Worker
    sig = inspect.signature(Worker.__init__)
    assert list(sig.parameters)[:7] == ["self", "client", "nodes", "max_workers", "group_id", "extra_publish_kwargs",
                                        "extra_subscribe_kwargs"]       # reference worker/worker.py:13-21
    assert sig.parameters["max_workers"].default == 1 and sig.parameters["group_id"].default is None
    client = Client.connect("localhost")
    w = Worker(client, nodes=[])
    w.register_handlers()
    with pytest.raises(RuntimeError, match="already called"):                 # worker.py:34-35
        w.register_handlers()
    assert inspect.iscoroutinefunction(Worker.run)


def test_agent_tool_naming_and_schema():
    from calfkit import agent_tool

    @agent_tool
    def get_weather(location: str) -> str:
        """Get the current weather at a location"""
        return f"It's sunny in {location}"

    assert get_weather.node_id == "tool_get_weather"                            # nodes/tool.py:30
    assert get_weather.subscribe_topics == ["tool.get_weather.input"]           # nodes/tool.py:91
    assert get_weather.publish_topic == "tool.get_weather.output"               # nodes/tool.py:92
    td = get_weather.tool_schema
    assert td.name == "get_weather" and td.description == "Get the current weather at a location"
    assert td.parameters_json_schema["properties"]["location"]["type"] == "string"
    assert td.parameters_json_schema["required"] == ["location"]
    assert get_weather.name == get_weather.id == "tool_get_weather"
    assert get_weather._return_topic == "tool_get_weather.private.return"       # nodes/base.py:174-176


def test_client_connect_and_signatures():
    from calfkit import Client
    c = Client.connect()
    assert c.reply_topic.startswith("calf-client-reply-") and len(c.reply_topic) == len("calf-client-reply-") + 32
    assert c.broker is c._connection
    c2 = Client.connect("k:9092", reply_topic="my-replies")
    assert c2.reply_topic == "my-replies"
    p = inspect.signature(Client.execute_node).parameters                         # client/client.py:154-218
    assert list(p)[:3] == ["self", "user_prompt", "topic"]
    for kw in ["tool_overrides", "output_type", "reply_topic", "correlation_id", "temp_instructions", "message_history",
               "run_args", "deps", "timeout"]:
        assert p[kw].kind is inspect.Parameter.KEYWORD_ONLY
    assert "timeout" not in inspect.signature(Client.invoke_node).parameters


def test_call_input_args_rule():
    from calfkit.models import Call, State
    assert Call("t", State()).input_args is None                                  # models/actions.py:66
    assert Call("t", State(), "a", "b").input_args == ("a", "b")


def test_first_envelope_is_unkeyed_and_canonical():
    """client/base.py:140-147: one frame (target, callback=reply topic), unkeyed publish; the bytes are
    a fixed point of the codec, which is what the device walker accepts."""
    import asyncio
    from calfkit import Client
    from hostsim import walk

    async def go():
        c = Client.connect()
        h = await c.invoke_node("What's the weather in Tokyo?", "weather_agent.input", deps={"k": 1})
        rec = c.broker.queues["weather_agent.input"][0]
        assert rec.key is None and rec.correlation_id == h.correlation_id
        ok, cols = walk(rec.value)
        assert ok, rec.value[int(cols[2]) - 30:int(cols[2]) + 30]
        with pytest.raises(RuntimeError, match="Duplicate correlation_id"):
            await c.invoke_node("x", "weather_agent.input", correlation_id=h.correlation_id)
        await c.close()
    asyncio.run(go())


def test_tool_template_and_uuid7():
    from calfkit.engine.batch import ToolTemplate, device_uuid7_hex
    t = ToolTemplate.from_format('It\'s "sunny"\n in {location}!')
    assert t.kinds == [0, 1, 0] and t.pieces == [b'"It\'s \\"sunny\\"\\n in ', b"location", b'!"']
    u = device_uuid7_hex(1767225600000, 7, 3)
    assert len(u) == 32 and u[12] == "7" and u[16] in "89ab" and int(u[:12], 16) == 1767225600000


def test_murmur2_vectorised_matches_kafka_reference():
    import numpy as np
    sys_path_hack = __import__("sys").path
    sys_path_hack.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
    from bench import np_murmur2_32
    from test_gpu_parity import _murmur2
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 256, size=(200, 32), dtype=np.uint8)
    got = np_murmur2_32(keys)
    assert [int(x) for x in got] == [_murmur2(bytes(k)) for k in keys]
    # known answers of Kafka's Utils.murmur2 (kafka-python / aiokafka partitioner test vectors)
    assert _murmur2(b"21") == 0xFFFFFFFF & -973932308
    assert _murmur2(b"foobar") == 0xFFFFFFFF & -790332482
    assert _murmur2(b"a-little-bit-long-string") == 0xFFFFFFFF & -985981536
    assert _murmur2(b"") == 275646681


def test_tool_definitions_match_reference_byte_for_byte():
    """What @agent_tool derives from a function — node id, topics and the ToolDefinition (name, description incl. the
    <summary>/<returns> wrapping, JSON schema with per-parameter descriptions, key order) — against the unmodified
    reference (tests/golden/tool_schemas.json, made by tests/golden/make_golden_schemas.py).  The definition travels in
    OverridesState on the wire, so its byte order matters.  A context parameter is recognised by its ANNOTATION, as in
    the reference (_function_schema._takes_ctx): an un-annotated `ctx` is an ordinary argument."""
    import json
    import os
    import sys
    from pydantic import TypeAdapter
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import tools_def
    import tools_more
    from calfkit import agent_tool
    from calfkit.models import ToolContext
    tools_more.ToolContext = ToolContext
    gold = {c["name"]: c for c in json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tool_schemas.json")))["cases"]}
    fns = {**tools_def.TOOLS, **tools_more.MORE}
    assert set(fns) == set(gold)
    for name, fn in fns.items():
        node = agent_tool(fn)
        g = gold[name]
        assert (node.node_id, list(node.subscribe_topics), node.publish_topic) == (g["node_id"], g["subscribe_topics"], g["publish_topic"])
        mine = json.dumps(json.loads(TypeAdapter(type(node.tool_schema)).dump_json(node.tool_schema)))
        assert mine == json.dumps(g["tool_schema"]), name
    assert agent_tool(tools_more.with_ctx)._tool.takes_ctx and not agent_tool(tools_more.contextual)._tool.takes_ctx


def test_host_tool_call_with_context_from_columns():
    """the host half of a contextual tool call (ToolNodeDef._call_host): the ToolContext is rebuilt from the column spans
    of the record — here produced by the CPU build of the walker (tests/hostsim) — exactly as the GPU path hands them over"""
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import tools_more
    from hostsim import walk
    from calfkit import agent_tool, synth
    from calfkit.models import ToolContext
    tools_more.ToolContext = ToolContext
    node = agent_tool(tools_more.with_ctx)
    rec = synth.tool_events(1, seed=5)[0]
    ok, cols = walk(rec)
    assert ok
    i = rec.index(b'"provided_deps":{') + len(b'"provided_deps":{')
    rec2 = rec[:i] + b'"tenant":"acme",' + rec[i:]
    ok, cols = walk(rec2)
    assert ok
    out = node._call_host(b'{"q":"ab","n":3}', memoryview(rec2), cols.reshape(-1, 1), 0)
    assert out == b'"acme:ababab"'


def test_client_first_envelope_matches_reference_bytes():
    """Client.invoke_node -> the first envelope of a correlation chain, byte for byte what the unmodified reference's client
    publishes for the same arguments (tests/golden/client_invoke.json: deps, temp_instructions, history, run_args, tool
    overrides carrying ToolDefinitions), unkeyed, with the same handle."""
    import asyncio
    import datetime as dt
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import tools_def
    import tools_more
    from calfkit import _ids, agent_tool
    from calfkit.client import Client
    from calfkit.models import ToolContext, messages
    tools_more.ToolContext = ToolContext
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "client_invoke.json")))
    fixed = dt.datetime.fromisoformat(gold["frozen_clock"])

    class Conn:
        _connection = True

        def __init__(self):
            self.published = []

        async def publish(self, envelope, topic, correlation_id, **kw):
            self.published.append({"topic": topic, "correlation_id": correlation_id, "key": kw.get("key"), "payload": envelope.model_dump_json()})

    class Disp:
        def expect(self, correlation_id):
            return None

    history = [messages.ModelRequest(parts=[messages.UserPromptPart(content="earlier question")]),
               messages.ModelResponse(parts=[messages.TextPart(content="earlier answer")], timestamp=fixed)]
    overrides = [agent_tool(tools_def.TOOLS["get_weather"]), agent_tool(tools_more.with_defaults), agent_tool(tools_more.google_multiline)]
    for case in gold["cases"]:
        kw = dict(case["args"])
        if kw.get("message_history") == "HISTORY":
            kw["message_history"] = list(history)
        if kw.get("tool_overrides") == "OVERRIDES":
            kw["tool_overrides"] = list(overrides)
        counter = [0]

        def det():
            counter[0] += 1
            return f"{counter[0]:032x}"
        _ids.set_id_source(det)
        old_now = messages.now_utc
        try:
            conn = Conn()
            client = Client(conn, "calf-client-reply-test", Disp())
            handle = asyncio.run(client.invoke_node(**kw))
        finally:
            _ids.set_id_source(None)
        assert len(conn.published) == 1
        got = conn.published[0]
        # the user-prompt timestamps are "now": align them with the golden's frozen clock before comparing bytes
        want = case["publish"]
        import re
        stamp = fixed.isoformat().replace("+00:00", "Z")
        norm = lambda s: re.sub(r'"timestamp":"[^"]+"', '"timestamp":"%s"' % stamp, s)   # noqa: E731
        assert (got["topic"], got["correlation_id"], got["key"]) == (want["topic"], want["correlation_id"], want["key"]), case["name"]
        assert norm(got["payload"]) == norm(want["payload"]), case["name"]
        assert (handle.correlation_id, handle.topic, handle.reply_topic) == tuple(case["handle"][k] for k in ("correlation_id", "topic", "reply_topic"))
        del old_now


def test_host_tool_return_values_encode_like_the_reference():
    """ToolNodeDef._call_host's JSON for a tool's Python return value (dict, list, None, pydantic model, dataclass, datetime,
    Decimal / UUID / Enum / bytes, big ints, 17-digit floats) is exactly what the unmodified reference put into
    tool_results[id].return_value (tests/golden/tool_returns.json)."""
    import json
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import tools_more
    from hostsim import walk
    from calfkit import agent_tool
    from calfkit.models import ToolContext
    tools_more.ToolContext = ToolContext
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tool_returns.json")))["cases"]
    for case in gold:
        assert case["raises"] is None
        fn, args = tools_more.RETURNS[case["name"]]
        node = agent_tool(fn)
        rec = case["input"].encode()
        ok, cols = walk(rec)                       # the column spans the GPU walker hands to the host for this record
        assert ok
        got = node._call_host(json.dumps(args, ensure_ascii=False).encode(), memoryview(rec), cols.reshape(-1, 1), 0)
        payload = case["publishes"][0]["payload"]
        want_prefix = '"tool_results":{"call_1":{"return_value":'
        i = payload.index(want_prefix) + len(want_prefix)
        j = payload.index(',"content":null,"metadata":{"tool_call_id":"call_1"}', i)
        assert got.decode() == payload[i:j], case["name"]


def test_product_reply_projection_matches_reference_goldens():
    """calfkit.client.deserialize (the per-request projection behind InvocationHandle.result) on tests/golden/replies.json"""
    import json
    import os
    import pydantic_core
    from calfkit.client.deserialize import _UNSET, deserialize_to_node_result
    from calfkit.models import Envelope
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replies.json")))["cases"]
    for case in gold:
        for label, ot in (("auto", _UNSET), ("str", str), ("dict", dict)):
            exp = case["expect"][label]
            try:
                res = deserialize_to_node_result(Envelope.model_validate_json(case["input"]), ot)
                assert exp["ok"] and pydantic_core.to_json(res.output).decode() == exp["output_json"] and res.correlation_id == exp["correlation_id"]
            except AssertionError:
                raise
            except Exception as e:  # noqa: BLE001
                assert not exp["ok"] and type(e).__name__ == exp["error"], (case["name"], label, repr(e))


def test_agent_host_half_matches_reference_run():
    """The host half of the Agent node (Agent._llm_step: aggregation gate, history bookkeeping around the model call, invalid
    tool handling, routing decision) against the unmodified reference's Agent.run on the same inbound envelopes and the same
    scripted model answers (tests/golden/agent_run.json): same action, same post-LLM state bytes (the RetryPromptPart
    timestamps are "now" on both sides and are aligned before comparing), same pending batch."""
    import datetime as dt
    import json
    import os
    import re
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import tools_def
    from calfkit import Agent, agent_tool
    from calfkit.models import Envelope, messages
    from calfkit.nodes import FunctionModelClient
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agent_run.json")))["cases"]
    tools = [agent_tool(tools_def.TOOLS[n]) for n in ("get_weather", "get_temperature", "count_chars")]
    fixed = dt.datetime(2026, 1, 1, tzinfo=dt.timezone.utc)
    norm = lambda s: re.sub(r'"timestamp":"[^"]+"', '"timestamp":"T"', s)   # noqa: E731
    want_action = {"list[Call]": "fanout", "Call": "fanout", "TailCall": "tailcall", "ReturnCall": "return"}
    for case in gold:
        answer = case["model_answer"]

        def model(msgs, tool_defs, answer=answer):
            if isinstance(answer, str):
                return messages.ModelResponse(parts=[messages.TextPart(content=answer)], timestamp=fixed)
            return messages.ModelResponse(parts=[messages.ToolCallPart(tool_name=t, args=a, tool_call_id=c) for (t, a, c) in answer], timestamp=fixed)
        agent = Agent("planner", subscribe_topics="planner.input", publish_topic="planner.output", tools=tools,
                      model_client=FunctionModelClient(model), sequential_only_mode=case["sequential"])
        env = Envelope.model_validate_json(case["input"])
        action, state = agent._llm_step(env.context.deps.correlation_id, env.context.state, env.context.deps.provided_deps)
        assert action == want_action[case["action"]], (case["name"], action)
        assert norm(state.model_dump_json()) == norm(case["post_llm_state"]), case["name"]
        pend = sorted(next(iter(agent._pending_batches.values())).expected_tool_call_ids) if agent._pending_batches else None
        assert pend == case["pending_batch_ids"], case["name"]


def test_agent_per_request_registry_switch_and_restore():
    """A request that brings its own tools (overrides.override_agent_tools, reference agent.py:71-75) is routed with that
    registry: the Agent points the engine at it for that group of the batch and puts the node's own configuration back
    afterwards — checked here on a recording stand-in for the engine (the device plan itself is the same fan-out kernel)."""
    from calfkit import Agent, agent_tool
    from calfkit.nodes import FunctionModelClient

    @agent_tool
    def own(x: str) -> str:
        """own"""
        return x

    class Rec:
        def __init__(self):
            self.calls = []
            self.topic_names, self.topic_ids, self.num_partitions = {}, {}, 8

        def register_topics(self, names, num_partitions=0):
            uniq = list(dict.fromkeys(names))
            self.topic_names = dict(enumerate(uniq))
            self.topic_ids = {n: i for i, n in enumerate(uniq)}
            self.calls.append(("register_topics", tuple(uniq)))

        def set_tool_node(self, publish_topic, template):
            self.calls.append(("set_tool_node", publish_topic))

        def set_agent_node(self, name, cb, publish_topic, registry):
            assert all(t in self.topic_ids for t in registry.values())
            self.calls.append(("set_agent_node", tuple(sorted(registry.items()))))

        def gate_create(self, **kw):               # the device gate is created once per engine, not per registry switch
            self.gate_created = getattr(self, "gate_created", 0) + 1

    agent = Agent("planner", subscribe_topics="planner.input", publish_topic="planner.output", tools=[own],
                  model_client=FunctionModelClient(lambda m, t: None))
    eng = Rec()
    eng.register_topics(["planner.input", "planner.output", "tool.own.input"], num_partitions=8)
    agent.configure_engine(eng)
    ids_before = dict(eng.topic_ids)
    eng.calls.clear()
    saved = agent._use_registry(eng, {"other": "tool.other.input", "own": "tool.own.input"})
    assert saved == ["planner.input", "planner.output", "tool.own.input"]
    assert all(eng.topic_ids[k] == v for k, v in ids_before.items()) and "tool.other.input" in eng.topic_ids      # ids stay put
    assert eng.calls[-1] == ("set_agent_node", (("other", "tool.other.input"), ("own", "tool.own.input")))
    agent._restore_registry(eng, saved)
    assert eng.topic_ids == ids_before
    assert eng.calls[-1] == ("set_agent_node", (("own", "tool.own.input"),)) and ("set_tool_node", "planner.output") in eng.calls


def test_async_host_tools_are_awaited():
    """`async def` tools (the reference awaits the tool call, nodes/tool.py:64): driven to completion by the batch step,
    outside and inside a running event loop (Worker.run calls the step from its loop)."""
    import asyncio
    import numpy as np
    from calfkit import agent_tool

    @agent_tool
    async def slow_upper(text: str) -> str:
        """upper-case, asynchronously"""
        await asyncio.sleep(0.01)
        return text.upper()

    cols = np.zeros((1, 1), dtype=np.uint32)
    assert slow_upper._call_host(b'{"text":"abc"}', memoryview(b""), cols, 0) == b'"ABC"'

    async def inside_loop():
        return slow_upper._call_host(b'{"text":"xyz"}', memoryview(b""), cols, 0)
    assert asyncio.run(inside_loop()) == b'"XYZ"'


def test_a_raising_tool_fails_only_its_record():
    """user code that raises costs its own record (logged, nothing published for it), not the batch — as with the reference,
    where the handler raises for that one message and the consumer moves on"""
    import numpy as np
    from calfkit import agent_tool
    from calfkit.engine._lib import CK_ACT_HOST_TOOL, CK_ACT_SILENT, COL, NUM_COLS

    @agent_tool
    def picky(x: int) -> int:
        """fails on odd input"""
        if x % 2:
            raise ValueError("odd")
        return x * 10

    cols = np.zeros((NUM_COLS, 4), dtype=np.uint32)
    cols[COL["ACTION"]] = [CK_ACT_HOST_TOOL, CK_ACT_HOST_TOOL, CK_ACT_SILENT, CK_ACT_HOST_TOOL]
    args = [b'{"x":2}', b'{"x":3}', b"", b'{"x":4}']
    results, failed = picky._host_results(4, lambda i: args[i], lambda i: memoryview(b""), cols)
    assert results == [b"20", b"null", b"", b"40"] and failed == {1}


def test_agent_instructions_are_composed_like_the_reference():
    """reference _vendor/pydantic_ai/agent/__init__.py:1465-1487,638-650 with nodes/agent.py:54,126: literals (system prompt,
    temp_instructions) joined by a newline, then the @agent.instructions function outputs joined by a blank line; the request
    that carries the tool returns records the composed string."""
    from calfkit.models import State
    from calfkit.models.messages import ModelRequest, ModelResponse, TextPart, ToolCallPart, ToolReturn, UserPromptPart
    from calfkit.nodes import Agent, FunctionModelClient
    seen = {}

    class Client:
        def __call__(self, messages, instructions, tools, deps):
            seen["instructions"], seen["last"] = instructions, messages[-1]
            return ModelResponse(parts=[TextPart(content="ok")])
    agent = Agent("planner", system_prompt="You plan.", subscribe_topics="planner.input", model_client=Client())
    agent.instructions(lambda: "Dynamic A.")
    agent.instructions(lambda: None)
    agent.instructions(lambda: "Dynamic B.")
    st = State(message_history=[ModelRequest(parts=[UserPromptPart(content="hi")])], temp_instructions="Only today.")
    agent._llm_step("c" * 32, st, {})
    assert seen["instructions"] == "You plan.\nOnly today.\n\nDynamic A.\n\nDynamic B."
    # no temp instructions, no functions: just the system prompt; and the tool-return request records it
    agent2 = Agent("planner", system_prompt="You plan.", subscribe_topics="planner.input", model_client=Client())
    tc = ToolCallPart(tool_name="t", args={}, tool_call_id="id1")
    st2 = State(message_history=[ModelRequest(parts=[UserPromptPart(content="hi")]), ModelResponse(parts=[tc])],
                tool_calls={"id1": tc}, tool_results={"id1": ToolReturn(return_value="r", metadata={"tool_call_id": "id1"})})
    _a, new = agent2._llm_step("d" * 32, st2, {})
    assert seen["instructions"] == "You plan." and seen["last"].instructions == "You plan."
    assert new.message_history[-2].instructions == "You plan."
