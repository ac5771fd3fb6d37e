"""Generates tests/golden/tool_schemas.json: what the UNMODIFIED reference's @agent_tool derives from a Python function —
node id, subscribe / publish topics (calfkit/nodes/tool.py:24-35,89-95) and the ToolDefinition sent to the model and carried
in OverridesState on the wire (calfkit/_vendor/pydantic_ai/tools.py:474-540).  Build container only:

    python tests/golden/make_golden_schemas.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

ref = rh.load_reference()
import tools_def  # noqa: E402
import tools_more  # noqa: E402
import importlib  # noqa: E402
tools_more.ToolContext = importlib.import_module("calfkit.models.tool_context").ToolContext     # the reference's context class
from pydantic import TypeAdapter  # noqa: E402

cases = []
for name, fn in {**tools_def.TOOLS, **tools_more.MORE}.items():
    node = ref.agent_tool(fn)
    td = node.tool_schema
    cases.append({"name": name, "node_id": node.node_id, "subscribe_topics": list(node.subscribe_topics), "publish_topic": node.publish_topic,
                  "tool_schema": json.loads(TypeAdapter(type(td)).dump_json(td))})
json.dump({"generated_by": "tests/golden/make_golden_schemas.py", "cases": cases}, open(os.path.join(HERE, "tool_schemas.json"), "w"),
          ensure_ascii=False, indent=1)
print("tool_schemas.json:", [c["name"] for c in cases])
