"""Generates tests/golden/replies.json by running the UNMODIFIED reference's client reply projection
(/root/reference/calfkit/client/deserialize.py:15-89) through oracle/ref_harness.py.  Build container only:

    python tests/golden/make_golden_replies.py

Each case: a reply envelope (bytes) -> for output_type in {unset (auto), str, dict}: the JSON of NodeResult.output
(pydantic_core.to_json) or the exception class the reference raised."""
import importlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

ref = rh.load_reference()
import pydantic_core  # noqa: E402

# calfkit/client/__init__.py pulls in the FastStream-based client (absent third-party dependency): register the
# package without running its __init__, then import the two unmodified modules this path consists of
import types  # noqa: E402
import calfkit  # noqa: E402  (the reference's, put on sys.path by the harness)
_pkg = types.ModuleType("calfkit.client")
_pkg.__path__ = [os.path.join(list(calfkit.__path__)[0], "client")]
sys.modules["calfkit.client"] = _pkg
des = importlib.import_module("calfkit.client.deserialize")
assert des.__file__.startswith("/root/reference/"), des.__file__
Envelope = ref.Envelope

BASE = ('{"context":{"state":{"tool_calls":{},"tool_results":{},"uncommitted_message":null,"message_history":[],'
        '"final_output_parts":%s,"temp_instructions":null,"metadata":null,"overrides":null},"deps":{"correlation_id":"%s",'
        '"provided_deps":{}}},"internal_workflow_state":{"call_stack":{"_internal_list":[]},"metadata":null}}')
TEXT = '{"kind":"text","text":%s,"metadata":null}'
DATA = '{"kind":"data","data":%s,"schema_":null,"metadata":null}'
FILE = '{"kind":"file","media_type":"text/plain","uri":null,"data":null,"metadata":null}'
TOOL = '{"kind":"tool","tool_call_id":"i","kwargs":{},"tool_name":"t","metadata":null}'
LONG = json.dumps("It's sunny in São Paulo — " + "x" * 300 + ' "quoted" \\ back\n\ttab', ensure_ascii=False)
parts = {
    "empty": [],
    "text_only": [TEXT % '"hello"'],
    "text_empty": [TEXT % '""'],
    "text_long_escapes": [TEXT % LONG],
    "data_only_obj": [DATA % '{"temp":21.5,"city":"Kraków","tags":["a","b"],"n":null}'],
    "data_null": [DATA % "null"],
    "data_scalar": [DATA % "42"],
    "data_string": [DATA % '"s"'],
    "data_big": [DATA % json.dumps({"rows": [{"i": i, "v": "y" * 20} for i in range(40)]}, separators=(",", ":"))],
    "text_then_data": [TEXT % '"t"', DATA % '{"a":1}'],
    "data_then_text": [DATA % '{"a":1}', TEXT % '"t"'],
    "two_texts": [TEXT % '"first"', TEXT % '"second"'],
    "two_datas": [DATA % "[1]", DATA % "[2]"],
    "file_tool_only": [FILE, TOOL],
    "file_then_text": [FILE, TEXT % '"after file"'],
    "tool_then_data_then_text": [TOOL, DATA % '{"k":"v"}', TEXT % '"x"'],
}
cases = []
for k, (name, ps) in enumerate(parts.items()):
    payload = (BASE % ("[" + ",".join(ps) + "]", f"{k:032x}")).encode()
    env = Envelope.model_validate_json(payload)
    assert env.model_dump_json().encode() == payload, name
    out = {}
    for label, ot in (("auto", des._UNSET), ("str", str), ("dict", dict)):
        try:
            res = des.deserialize_to_node_result(Envelope.model_validate_json(payload), ot)
            out[label] = {"ok": True, "output_json": pydantic_core.to_json(res.output).decode(), "correlation_id": res.correlation_id}
        except Exception as e:  # noqa: BLE001
            out[label] = {"ok": False, "error": type(e).__name__}
    cases.append({"name": name, "input": payload.decode(), "expect": out})
json.dump({"generated_by": "tests/golden/make_golden_replies.py", "cases": cases}, open(os.path.join(HERE, "replies.json"), "w"),
          ensure_ascii=False, indent=0)
print("replies.json:", len(cases), "cases")
for c in cases:
    print(c["name"], {k: (v.get("output_json", v.get("error"))[:30]) for k, v in c["expect"].items()})
