"""Generates tests/golden/*.json by running the UNMODIFIED reference (/root/reference) through
oracle/ref_harness.py.  Run in the build container only (the reference does not travel):

    python tests/golden/make_golden.py

Every vector is (input bytes) -> what the reference's own code produced:
  codec.json       model_validate_json -> model_dump_json fixed points, re-canonicalisations and
                   pydantic error types                         (reference models/envelope.py:9-17)
  tool_node.json   ToolNodeDef.handler + _publish_action through a capture broker, plus the
                   handler-return publish to publish_topic       (nodes/base.py:70-164, nodes/tool.py:37-86,
                                                                  worker/worker.py:52-53)
  actions.json     _publish_action for Call / TailCall / ReturnCall / list[Call] / Silent with
                   injected deterministic frame ids              (nodes/base.py:70-147)
"""
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

# the product's generator builds bytes only; import it by path, without its `calfkit` package
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("ck_synth", os.path.join(ROOT, "calfkit-sdk_b200", "calfkit", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
sys.modules["ck_synth"] = synth
_spec.loader.exec_module(synth)
import tools_def  # noqa: E402
from pydantic import ValidationError  # noqa: E402

Envelope = ref.Envelope


def s(b: bytes):
    """JSON-storable form of input bytes: the text itself, or {"b64": ...} if not valid UTF-8."""
    try:
        return b.decode("utf-8")
    except UnicodeDecodeError:
        import base64
        return {"b64": base64.b64encode(b).decode()}


def codec_case(name, payload: bytes):
    try:
        env = Envelope.model_validate_json(payload)
        return {"name": name, "input": s(payload), "ok": True, "output": env.model_dump_json()}
    except ValidationError as e:
        errs = e.errors()
        return {"name": name, "input": s(payload), "ok": False, "n_errors": len(errs),
                "first_type": errs[0]["type"], "first_loc": [str(x) for x in errs[0]["loc"]]}


def run_tool_node(node, payload: bytes, header_corr=None):
    env = Envelope.model_validate_json(payload)
    corr = header_corr or env.context.deps.correlation_id
    br = rh.CaptureBroker()
    try:
        ret = asyncio.run(node.handler(env, corr, br))
    except Exception as e:  # the reference lets it propagate (FastStream would log + nack)
        return {"raises": type(e).__name__}
    pubs = [{"topic": t, "key": (k.decode() if k is not None else None), "correlation_id": c, "payload": p.decode()}
            for (t, k, c, p) in br.published]
    if node.publish_topic:   # worker/worker.py:52-53: the handler's return value is published too
        pubs.append({"topic": node.publish_topic, "key": None, "correlation_id": corr,
                     "payload": ret.model_dump_json()})
    return {"publishes": pubs}


def main():
    nodes = {name: ref.agent_tool(fn) for name, fn in tools_def.TOOLS.items()}

    # ---------------------------------------------------------------- tool_node.json
    cases = []

    def add(name, tool, payload, **kw):
        cases.append({"name": name, "tool": tool, "input": s(payload), **run_tool_node(nodes[tool], payload, **kw)})

    for i, rec in enumerate(synth.tool_events(24, seed=11)):
        add(f"1k_{i}", "get_weather", rec)
    for i, rec in enumerate(synth.tool_events(8, seed=12, size=None, full_history=True)):
        add(f"full_{i}", "get_weather", rec)
    for i, rec in enumerate(synth.tool_events(4, seed=13, size=None, tool_name="get_temperature")):
        add(f"dict_result_{i}", "get_temperature", rec)
    for i, rec in enumerate(synth.tool_events(2, seed=14, size=None, tool_name="count_chars")):
        add(f"int_result_{i}", "count_chars", rec)

    base = synth.tool_events(1, seed=15, size=None)[0].decode()
    cid = json.loads(base)["internal_workflow_state"]["call_stack"]["_internal_list"][-1]["input_args"][0]
    # tool_call_id that is not in tool_calls -> Silent: no callback publish, input echoed to publish_topic
    add("silent_missing_call", "get_weather", base.replace(f'["{cid}","weather_agent"]', '["nope","weather_agent"]').encode())
    # existing results: one for another id, one for the same id (dict assignment keeps position)
    other = '"other_id":{"return_value":1,"content":null,"metadata":null,"kind":"tool-return"}'
    add("existing_other_result", "get_weather", base.replace('"tool_results":{}', '"tool_results":{' + other + '}').encode())
    same = f'"{cid}":{{"message":"try again","kind":"model-retry"}}'
    add("existing_same_result_first", "get_weather",
        base.replace('"tool_results":{}', '"tool_results":{' + same + ',' + other + '}').encode())
    add("existing_same_result_last", "get_weather",
        base.replace('"tool_results":{}', '"tool_results":{' + other + ',' + same + '}').encode())
    # args given as a JSON string, as null (-> {}), and empty object
    # args given as a JSON string / null (args_as_dict: str -> from_json, falsy -> {}), canonical and
    # with an unknown key alongside (valid but not a fixed point: the key is dropped on re-emit)
    a0 = base.index('"args":{"location":')
    a1 = base.index('}', a0) + 1
    add("args_json_string", "get_weather", (base[:a0] + '"args":"{\\"location\\": \\"Rome\\"}"' + base[a1:]).encode())
    add("args_json_string_unknown_key", "get_weather", base.replace('"args":{"location":', '"args":"{\\"location\\": \\"Rome\\"}","zz":{"location":', 1).encode())
    na = synth.tool_events(1, seed=16, size=None, tool_name="no_args")[0].decode()
    n0 = na.index('"args":{"location":')
    n1 = na.index('}', n0) + 1
    add("no_args_tool_args_null", "no_args", (na[:n0] + '"args":null' + na[n1:]).encode())
    add("no_args_tool_args_empty", "no_args", (na[:n0] + '"args":{}' + na[n1:]).encode())
    # frame overrides replace state.overrides (prepare_context)
    ov = ('{"override_agent_tools":[{"node_id":"tool_x","subscribe_topics":["tool.x.input"],"publish_topic":null,'
          '"tool_schema":{"name":"x","parameters_json_schema":{"type":"object","properties":{}},"description":null,'
          '"outer_typed_dict_key":null,"strict":null,"sequential":false,"kind":"function","metadata":null,"timeout":null}}]}')
    last_null = base.rindex('"overrides":null')
    add("frame_overrides", "get_weather", (base[:last_null] + '"overrides":' + ov + base[last_null + len('"overrides":null'):]).encode())
    add("frame_overrides_none_list", "get_weather",
        (base[:last_null] + '"overrides":{"override_agent_tools":null}' + base[last_null + len('"overrides":null'):]).encode())
    first_null = base.index('"overrides":null')
    add("state_overrides_kept", "get_weather", (base[:first_null] + '"overrides":' + ov + base[first_null + len('"overrides":null'):]).encode())
    # Any subtrees with numbers / nesting / unicode, workflow metadata
    add("any_values", "get_weather", base.replace('"temp_instructions":null,"metadata":null',
        '"temp_instructions":"be brief","metadata":{"n":[1,-2,3.5,1e+22,1e-7,-0.0,12345678901234567890123],"s":"é\\n","o":{"a":null,"b":true}}').encode())
    add("wf_metadata", "get_weather", (base[:-len('"metadata":null}}')] + '"metadata":{"trace":["a",1]}}}').encode())
    # three frames deep
    deep = base.replace('"_internal_list":[', '"_internal_list":[' + synth.frame("root.input", "calf-client-reply-0", ["ARGS"], "0" * 32).replace('["ARGS"]', '["a",1,null,{"k":[1.5]}]') + ",")
    add("three_frames", "get_weather", deep.encode())
    # header correlation id differing from deps (key comes from the header value)
    add("header_corr_differs", "get_weather", base.encode(), header_corr="header-corr-id")
    # input_args null -> run() called without args -> TypeError propagates
    ia = base.rindex('"input_args":[')
    ia_end = base.index("]", ia)
    add("input_args_null", "get_weather", (base[:ia] + '"input_args":null' + base[ia_end + 1:]).encode())
    # non-canonical but valid input (whitespace, key order, missing defaults, unknown keys)
    obj = json.loads(base)
    obj["context"]["state"].pop("final_output_parts")
    obj["context"]["state"]["zzz_unknown"] = {"x": 1}
    obj["context"] = {"deps": obj["context"]["deps"], "state": obj["context"]["state"]}
    add("noncanonical_valid", "get_weather", json.dumps(obj, indent=1).encode())

    json.dump({"generated_by": "tests/golden/make_golden.py against /root/reference @ v0.2.5", "cases": cases},
              open(os.path.join(HERE, "tool_node.json"), "w"), ensure_ascii=False, indent=0)
    print("tool_node.json:", len(cases), "cases")

    # ---------------------------------------------------------------- codec.json
    C = []
    for i, rec in enumerate(synth.tool_events(6, seed=21)):
        C.append(codec_case(f"synth_1k_{i}", rec))
    for i, rec in enumerate(synth.tool_events(4, seed=22, size=None, full_history=True)):
        C.append(codec_case(f"synth_full_{i}", rec))
    for i, rec in enumerate(synth.fanout_events(2, seed=23, fanout=8)):
        C.append(codec_case(f"synth_fanout_{i}", rec))
    for i, rec in enumerate(synth.mixed_events(12, seed=24, lo=128, hi=20000, n_tools=16)):
        C.append(codec_case(f"synth_mixed_{i}", rec))
    BASE = ('{"context":{"state":{"tool_calls":{},"tool_results":{},"uncommitted_message":null,"message_history":[],'
            '"final_output_parts":[],"temp_instructions":null,"metadata":null,"overrides":null},"deps":{"correlation_id":"c",'
            '"provided_deps":{}}},"internal_workflow_state":{"call_stack":{"_internal_list":[]},"metadata":null}}')

    def rep(a, b):
        assert a in BASE, a
        return BASE.replace(a, b).encode()

    MD = '"metadata":null,"overrides"'
    bs = chr(92)
    edits = {
        "base": BASE.encode(),
        "minimal": b'{"context":{"state":{},"deps":{"correlation_id":"c","provided_deps":{}}},"internal_workflow_state":{"call_stack":{"_internal_list":[]}}}',
        "dupkey_last_wins": rep('"temp_instructions":null', '"temp_instructions":"a","temp_instructions":"b"'),
        "any_numbers": rep(MD, '"metadata":{"b":1, "a":[1.0,2.50,1e5,1E-7,-0.0, 1e400, 123456789012345678901234567890, 0.1e1, 1.5e300, 5e-324, 1e22,1e21,1e16,1e15, 123456789.123456789, 0.000001, 0.0000001, 0.00001, 0.0001, 123456789012345680.0, 1.7976931348623157e308, 2.2250738585072014e-308, 9007199254740993, 0.30000000000000004, 100, -7]},"overrides"'),
        "any_canonical_numbers": rep(MD, '"metadata":[1.0,2.5,100000.0,1e-7,-0.0,1e+22,1e+21,1e+16,1000000000000000.0,123456789.12345679,1e-6,0.00001,0.0001,-12,0,12345678901234567890123456789]' + ',"overrides"'),
        "str_field_int": rep('"correlation_id":"c"', '"correlation_id":5'),
        "opt_str_field_int": rep('"temp_instructions":null', '"temp_instructions":5'),
        "escapes_noncanonical": rep('"correlation_id":"c"', '"correlation_id":"aA' + bs + '/' + bs + 'u00e9' + bs + 'ud83d' + bs + 'ude00' + bs + 'b' + bs + 'f' + bs + 'n' + bs + 'r' + bs + 't' + bs + '"' + bs + bs + bs + 'u007f' + bs + 'u001F"'),
        "escapes_canonical": rep('"correlation_id":"c"', '"correlation_id":"é😀' + bs + 'b' + bs + 'f' + bs + 'n' + bs + 'r' + bs + 't' + bs + '"' + bs + bs + chr(0x7f) + bs + 'u001f' + bs + 'u0000/"'),
        "escape_upper_hex": rep('"correlation_id":"c"', '"correlation_id":"' + bs + 'u001F"'),
        "lone_surrogate": rep('"correlation_id":"c"', '"correlation_id":"' + bs + 'ud83d"'),
        "bad_escape": rep('"correlation_id":"c"', '"correlation_id":"' + bs + 'x"'),
        "ctrl_in_string": rep('"correlation_id":"c"', '"correlation_id":"a' + chr(1) + '"'),
        "tab_in_string": rep('"correlation_id":"c"', '"correlation_id":"a\tb"'),
        "trailing_ws": BASE.encode() + b" \n",
        "leading_ws": b"  " + BASE.encode(),
        "inner_ws": BASE.replace(":", ": ").replace(",", ", ").encode(),
        "trailing_garbage": BASE.encode() + b"x",
        "bom": b"\xef\xbb\xbf" + BASE.encode(),
        "empty": b"", "null": b"null", "list": b"[]", "string": b'"x"', "truncated": BASE[:-1].encode(),
        "nan": rep(MD, '"metadata":NaN,"overrides"'), "inf": rep(MD, '"metadata":Infinity,"overrides"'),
        "neg_inf": rep(MD, '"metadata":-Infinity,"overrides"'),
        "leading_zero": rep(MD, '"metadata":01,"overrides"'), "neg_zero_int": rep(MD, '"metadata":-0,"overrides"'),
        "dot_end": rep(MD, '"metadata":1.,"overrides"'), "dot_start": rep(MD, '"metadata":.5,"overrides"'),
        "plus": rep(MD, '"metadata":+1,"overrides"'), "cap_true": rep(MD, '"metadata":True,"overrides"'),
        "trailing_comma_obj": rep('"provided_deps":{}', '"provided_deps":{"a":1,}'),
        "trailing_comma_arr": rep('"message_history":[]', '"message_history":[,]'),
        "unknown_key_state": rep('"tool_calls":{}', '"zzz":[1,{"a":2}],"tool_calls":{}'),
        "unknown_key_envelope": BASE[:-1].encode() + b',"extra":1}',
        "missing_deps": rep(',"deps":{"correlation_id":"c","provided_deps":{}}', ''),
        "missing_provided_deps": rep(',"provided_deps":{}', ''),
        "overrides_none_list": rep('"overrides":null},"deps"', '"overrides":{"override_agent_tools":null}},"deps"'),
        "overrides_empty_obj": rep('"overrides":null},"deps"', '"overrides":{}},"deps"'),
        "overrides_defaults_filled": rep('"overrides":null},"deps"', '"overrides":{"override_agent_tools":[{"node_id":"n","subscribe_topics":["a"],"publish_topic":null,"tool_schema":{"name":"x"}}]}},"deps"'),
        "depth_100": rep(MD, '"metadata":' + '[' * 100 + ']' * 100 + ',"overrides"'),
        "depth_190": rep(MD, '"metadata":' + '[' * 190 + ']' * 190 + ',"overrides"'),
        "depth_300": rep(MD, '"metadata":' + '[' * 300 + ']' * 300 + ',"overrides"'),
        "dup_key_in_any": rep(MD, '"metadata":{"a":1,"a":2},"overrides"'),
        "dup_key_in_deps": rep('"provided_deps":{}', '"provided_deps":{"a":1,"b":2,"a":3}'),
        "tool_result_tagged_noncanon": rep('"tool_results":{}', '"tool_results":{"x":{"kind":"tool-return","return_value":1}}'),
        "tool_result_tagged_canon": rep('"tool_results":{}', '"tool_results":{"x":{"return_value":1,"content":null,"metadata":null,"kind":"tool-return"}}'),
        "tool_result_tagged_invalid_falls_to_any": rep('"tool_results":{}', '"tool_results":{"x":{"kind":"tool-return"}}'),
        "tool_result_untagged_any": rep('"tool_results":{}', '"tool_results":{"x":{"foo":[1,2]},"y":3,"z":"s","w":null}'),
        "tool_result_other_kind": rep('"tool_results":{}', '"tool_results":{"x":{"kind":"zzz","part_kind":"retry-prompt"}}'),
        "tool_result_retry_prompt": rep('"tool_results":{}', '"tool_results":{"x":{"content":"nope","tool_name":"t","tool_call_id":"i","timestamp":"2026-01-01T00:00:00Z","part_kind":"retry-prompt"}}'),
        "tool_result_model_retry": rep('"tool_results":{}', '"tool_results":{"x":{"message":"again","kind":"model-retry"}}'),
        "datetime_micro": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":"2026-01-01T00:00:00.123456Z","instructions":null,"kind":"request","run_id":null,"metadata":null}]'),
        "datetime_offset": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":"2026-01-01T02:00:00+02:00","instructions":null,"kind":"request","run_id":null,"metadata":null}]'),
        "datetime_naive": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":"2026-01-01T00:00:00","instructions":null,"kind":"request","run_id":null,"metadata":null}]'),
        "datetime_space": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":"2026-01-01 00:00:00Z","instructions":null,"kind":"request","run_id":null,"metadata":null}]'),
        "datetime_number": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":1767225600,"instructions":null,"kind":"request","run_id":null,"metadata":null}]'),
        "datetime_milli3": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":"2026-01-01T00:00:00.120Z","instructions":null,"kind":"request","run_id":null,"metadata":null}]'),
        "datetime_bad": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":"2026-13-01T00:00:00Z","instructions":null,"kind":"request","run_id":null,"metadata":null}]'),
        "bad_kind": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":null,"instructions":null,"kind":"nope","run_id":null,"metadata":null}]'),
        "missing_kind": rep('"message_history":[]', '"message_history":[{"parts":[]}]'),
        "usage_int_coercions": rep('"message_history":[]', '"message_history":[{"parts":[],"usage":{"input_tokens":"5","output_tokens":7.0,"details":{"x":"3"}},"timestamp":"2026-01-01T00:00:00Z","kind":"response"}]'),
        "usage_int_bad": rep('"message_history":[]', '"message_history":[{"parts":[],"usage":{"input_tokens":1.5},"timestamp":"2026-01-01T00:00:00Z","kind":"response"}]'),
        "response_vendor_alias": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":"2026-01-01T00:00:00Z","kind":"response","vendor_details":{"a":1},"vendor_id":"v"}]'),
        "finish_reason_bad": rep('"message_history":[]', '"message_history":[{"parts":[],"timestamp":"2026-01-01T00:00:00Z","kind":"response","finish_reason":"bogus"}]'),
        "frame_args_tuple_types": rep('"_internal_list":[]', '"_internal_list":[{"target_topic":"a","callback_topic":"b","input_args":[1,"x",null,{"k":[true]}],"frame_id":"f","overrides":null}]'),
        "frame_missing_frame_id": rep('"_internal_list":[]', '"_internal_list":[{"target_topic":"a","callback_topic":"b"}]'),
        "frame_args_string": rep('"_internal_list":[]', '"_internal_list":[{"target_topic":"a","callback_topic":"b","input_args":"xy","frame_id":"f","overrides":null}]'),
        "final_parts_all": rep('"final_output_parts":[]', '"final_output_parts":[{"kind":"text","text":"hi","metadata":null},{"kind":"data","data":{"a":[1]},"schema_":null,"metadata":null},{"kind":"file","media_type":"text/plain","uri":null,"data":null,"metadata":null},{"kind":"tool","tool_call_id":"i","kwargs":{},"tool_name":"t","metadata":null}]'),
        "final_parts_data_alias": rep('"final_output_parts":[]', '"final_output_parts":[{"kind":"data","data":1,"schema":{"a":1}}]'),
        "final_parts_bad_kind": rep('"final_output_parts":[]', '"final_output_parts":[{"kind":"nope"}]'),
        "user_content_list": rep('"message_history":[]', '"message_history":[{"parts":[{"content":["a","b"],"timestamp":"2026-01-01T00:00:00Z","name":null,"part_kind":"user-prompt"}],"timestamp":null,"instructions":null,"kind":"request","run_id":null,"metadata":null}]'),
        "bool_coercion": rep('"overrides":null},"deps"', '"overrides":{"override_agent_tools":[{"node_id":"n","subscribe_topics":["a"],"publish_topic":null,"tool_schema":{"name":"x","sequential":"true","strict":1,"timeout":5}}]}},"deps"'),
    }
    for name, payload in edits.items():
        C.append(codec_case(name, payload))
    for bad in [b"\xff", b"\xc3", b"\xe2\x82", b"\xc0\xaf", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xf0\x9f\x98\x80", b"\xf4\x8f\xbf\xbf", b"\xe0\x80\x80", b"\xef\xbf\xbf"]:
        C.append(codec_case("utf8_" + bad.hex(), BASE.encode().replace(b'"correlation_id":"c"', b'"correlation_id":"' + bad + b'"')))
    json.dump({"generated_by": "tests/golden/make_golden.py", "cases": C}, open(os.path.join(HERE, "codec.json"), "w"),
              ensure_ascii=False, indent=0)
    print("codec.json:", len(C), "cases;", sum(1 for c in C if c["ok"]), "ok")

    # ---------------------------------------------------------------- actions.json
    A = []
    models = ref.models

    class Scripted(ref.BaseNodeDef):
        def __init__(self, script):
            super().__init__(node_id="scripted", subscribe_topics=["scripted.input", "scripted.alt"], publish_topic="scripted.output")
            self.script = script

        async def run(self, ctx):
            return self.script(ctx)

    counter = [0]

    def det_uuid():
        counter[0] += 1
        return f"{counter[0]:032x}"

    rh.set_uuid_source(det_uuid)
    src = synth.fanout_events(1, seed=31, fanout=4)[0]
    src_env = json.loads(src)
    ids = list(src_env["context"]["state"]["tool_calls"].keys())

    def act(name, script, payload=src):
        counter[0] = 0
        env = Envelope.model_validate_json(payload)
        br = rh.CaptureBroker()
        node = Scripted(script)
        ret = asyncio.run(node.handler(env, env.context.deps.correlation_id, br))
        pubs = [{"topic": t, "key": k.decode() if k else None, "correlation_id": c, "payload": p.decode()} for (t, k, c, p) in br.published]
        pubs.append({"topic": "scripted.output", "key": None, "correlation_id": env.context.deps.correlation_id, "payload": ret.model_dump_json()})
        A.append({"name": name, "input": s(payload), "publishes": pubs})

    act("call", lambda ctx: models.Call("tool.tool_00.input", ctx.state, ids[0], "scripted"))
    act("call_no_args", lambda ctx: models.Call("other.input", ctx.state))
    act("tailcall", lambda ctx: models.TailCall("scripted.input", ctx.state))
    act("returncall", lambda ctx: models.ReturnCall(ctx.state))
    act("silent", lambda ctx: models.Silent())
    act("fanout", lambda ctx: [models.Call(f"tool.tool_{j:02d}.input", ctx.state.model_copy(deep=True), ids[j], "scripted") for j in range(4)])
    json.dump({"generated_by": "tests/golden/make_golden.py", "frame_id_source": "counter: f'{n:032x}' starting at 1 per case",
               "cases": A}, open(os.path.join(HERE, "actions.json"), "w"), ensure_ascii=False, indent=0)
    print("actions.json:", len(A), "cases")


if __name__ == "__main__":
    main()
