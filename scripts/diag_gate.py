"""diagnostic: the first golden aggregation case through submit -> fanout_plan -> gate_register, printing the columns"""
import json, sys
sys.path.insert(0, "tests"); sys.path.insert(0, "calfkit-sdk_b200")
import numpy as np
from test_gpu_gate import envelope
from conftest import golden
from calfkit.engine import BatchEngine
from calfkit.engine._lib import COL, LIB_PATH
from calfkit import synth
print("lib", LIB_PATH)
e = BatchEngine(0, max_records=4096, max_in_bytes=128 << 20, max_payloads=4096 * 70)
tools = {f"tool_{j:02d}": f"tool.tool_{j:02d}.input" for j in range(64)}
e.register_topics(list(tools.values()) + ["planner.input", "planner.output"], num_partitions=8)
e.set_tool_node("planner.output", None)
e.set_agent_node("planner", "planner.input", "planner.output", tools)
e.gate_create(max_entries=1024, arena_bytes=64 << 20)
for case in golden("aggregation.json")[:3]:
    env = envelope(case["base_state"], case["correlation_id"])
    b = synth.pack([env])
    e.submit(b.data, b.offsets)
    e.fanout_plan(1767225600000, 7, max_fanout=256)
    e.gate_register(min_pending=1)
    e.sync()
    cols = e.columns()
    print(case["name"], "live", e.gate_stats(), "status/action/tc/tr/trlen", [int(cols[COL[k]][0]) for k in ("STATUS", "ACTION", "TC_OFF", "TR_OFF", "TR_LEN")], len(env))
e.close()
