"""Generates tests/golden/aggregation.json by running the UNMODIFIED reference's aggregation gate
(/root/reference/calfkit/nodes/agent.py:57-68, BaseAgentNodeDef._parallel_state_aggregation) through
oracle/ref_harness.py.  Build container only:

    python tests/golden/make_golden_aggregation.py

The method only touches `self._pending_batches` and the context; the module's imports of the vendored agent loop and
of the provider client (network-facing, absent third-party dependencies) are satisfied with inert stand-ins so that
the module — and the method under test, unmodified — can be imported.

Each case: base state (the state at fan-out time), expected tool-call ids, and a sequence of returning states
(each the base state plus the results that particular tool node produced, as the reference's tool nodes return them)
-> after every arrival: "incomplete" or the merged state JSON the agent continues with."""
import importlib
import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness as rh  # noqa: E402

ref = rh.load_reference()
pai = sys.modules["calfkit._vendor.pydantic_ai"]
pai.Agent = type("Agent", (), {"__class_getitem__": classmethod(lambda cls, item: cls)})
pai.DeferredToolRequests = type("DeferredToolRequests", (), {})
for name, attrs in (("calfkit._vendor.pydantic_ai.output", {"OutputSpec": type("OutputSpec", (), {"__class_getitem__": classmethod(lambda c, i: c)})}),
                    ("calfkit._vendor.pydantic_ai.toolsets", {}),
                    ("calfkit._vendor.pydantic_ai.toolsets.external", {"ExternalToolset": object}),
                    ("calfkit.providers", {}), ("calfkit.providers.pydantic_ai", {}),
                    ("calfkit.providers.pydantic_ai.model_client", {"PydanticModelClient": object})):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__path__ = []
        sys.modules[name] = m
tools_mod = sys.modules["calfkit._vendor.pydantic_ai.tools"]
if not hasattr(tools_mod, "DeferredToolResults"):
    tools_mod.DeferredToolResults = object
sys.path.insert(0, rh.REF_ROOT)
agent_mod = importlib.import_module("calfkit.nodes.agent")
sys.path.remove(rh.REF_ROOT)
assert agent_mod.__file__.startswith("/root/reference/"), agent_mod.__file__
gate = agent_mod.BaseAgentNodeDef._parallel_state_aggregation
state_mod = importlib.import_module("calfkit.models.state")
State, PendingToolBatch = state_mod.State, state_mod.PendingToolBatch
session = importlib.import_module("calfkit.models.session_context")

import importlib.util  # noqa: E402
_spec = importlib.util.spec_from_file_location("ck_synth", os.path.join(ROOT, "calfkit-sdk_b200", "calfkit", "synth.py"))
synth = importlib.util.module_from_spec(_spec)
sys.modules["ck_synth"] = synth
_spec.loader.exec_module(synth)

rng = random.Random(17)
cases = []
for k, (fan, order_kind) in enumerate([(2, "in_order"), (3, "reverse"), (5, "shuffled"), (4, "duplicates"), (3, "foreign_first"), (1, "single"), (6, "shuffled")]):
    env = json.loads(synth.fanout_events(1, seed=50 + k, fanout=fan)[0])
    base_json = json.dumps(env["context"]["state"], separators=(",", ":"), ensure_ascii=False)
    corr = env["context"]["deps"]["correlation_id"]
    ids = list(env["context"]["state"]["tool_calls"].keys())
    arrivals = list(ids)
    if order_kind == "reverse":
        arrivals.reverse()
    elif order_kind == "shuffled":
        rng.shuffle(arrivals)
    elif order_kind == "duplicates":
        arrivals = [ids[0], ids[0], ids[1], ids[2], ids[1], ids[3]]
    fake = types.SimpleNamespace(_pending_batches={corr: PendingToolBatch(expected_tool_call_ids=frozenset(ids),
                                                                         base_state=State.model_validate_json(base_json))})
    steps = []
    if order_kind == "foreign_first":
        # a context of another correlation chain passes through untouched (no pending batch for it)
        other = State.model_validate_json(base_json)
        ctx = types.SimpleNamespace(state=other, deps=types.SimpleNamespace(correlation_id="f" * 32))
        gate(fake, ctx)
        steps.append({"correlation_id": "f" * 32, "incoming": other.model_dump_json(), "complete": True, "state": ctx.state.model_dump_json(),
                      "pending_left": sorted(fake._pending_batches)})
    for n_arr, tid in enumerate(arrivals):
        st = State.model_validate_json(base_json)
        # what the tool node returns: the state it was called with + its own result
        st.add_tool_result(tid, {"return_value": f"result of {tid} #{n_arr}", "content": None, "metadata": {"tool_call_id": tid}, "kind": "tool-return"})
        incoming = st.model_dump_json()
        st = State.model_validate_json(incoming)                        # as it arrives off the wire
        ctx = types.SimpleNamespace(state=st, deps=types.SimpleNamespace(correlation_id=corr))
        gate(fake, ctx)
        batch = fake._pending_batches.get(corr)
        complete = batch is None
        steps.append({"correlation_id": corr, "incoming": incoming, "complete": complete,
                      "state": ctx.state.model_dump_json() if complete else None, "pending_left": sorted(fake._pending_batches)})
        if complete:
            break
    cases.append({"name": f"{order_kind}_{fan}", "base_state": base_json, "correlation_id": corr, "expected_ids": ids, "steps": steps})
json.dump({"generated_by": "tests/golden/make_golden_aggregation.py", "cases": cases}, open(os.path.join(HERE, "aggregation.json"), "w"),
          ensure_ascii=False, indent=0)
print("aggregation.json:", len(cases), "cases;", [(c["name"], len(c["steps"]), c["steps"][-1]["complete"]) for c in cases])
