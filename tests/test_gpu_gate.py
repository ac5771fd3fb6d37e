"""SURVEY.md section 8 rows a10 / f4 — the aggregation gate on the device (csrc/ck_gate.cuh) against the reference's own gate:
tests/golden/aggregation.json was produced by the unmodified `_parallel_state_aggregation` (reference nodes/agent.py:57-68);
a 64-way fan-out with shuffled, duplicated and foreign arrivals is checked against the oracle port of the same function."""
import json
import random

import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu

FRAME = ('{"target_topic":"planner.input","callback_topic":"calf-client-reply-1","input_args":null,'
         '"frame_id":"00000000000000000000000000000001","overrides":null}')


def envelope(state_json: str, corr: str) -> bytes:
    return ('{"context":{"state":' + state_json + ',"deps":{"correlation_id":' + json.dumps(corr) + ',"provided_deps":{}}},'
            '"internal_workflow_state":{"call_stack":{"_internal_list":[' + FRAME + ']},"metadata":null}}').encode()


def state_of(env: bytes) -> str:
    s = env.decode()
    return s[len('{"context":{"state":'):s.index(',"deps":{"correlation_id":')]


@pytest.fixture()
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from calfkit.engine import BatchEngine
    e = BatchEngine(0, max_records=4096, max_in_bytes=128 << 20, max_payloads=4096 * 70)
    tools = {f"tool_{j:02d}": f"tool.tool_{j:02d}.input" for j in range(64)}
    e.register_topics(list(tools.values()) + ["planner.input", "planner.output"], num_partitions=8)
    e.set_tool_node("planner.output", None)
    e.set_agent_node("planner", "planner.input", "planner.output", tools)
    e.gate_create(max_entries=1024, arena_bytes=64 << 20)
    yield e
    e.close()


def register(engine, base_envs):
    from calfkit import synth
    b = synth.pack(base_envs)
    engine.submit(b.data, b.offsets)
    engine.fanout_plan(1767225600000, 7, max_fanout=256)
    engine.gate_register(min_pending=1)      # the golden cases include a one-id batch (the reference's rule is > 1)
    engine.sync()
    assert engine.gate_stats()["live"] == len(base_envs)


def arrive(engine, envs, stamp_base):
    from calfkit import synth
    from calfkit.engine._lib import COL
    b = synth.pack(envs)
    engine.submit(b.data, b.offsets)
    engine.gate_arrive(stamp_base)
    out = engine.fetch()
    return out, out.cols[COL["ACTION"]]


@pytest.mark.parametrize("case", golden("aggregation.json"), ids=lambda c: c["name"])
@pytest.mark.parametrize("batched", [False, True], ids=["one_by_one", "one_batch"])
def test_gate_matches_reference_goldens(engine, case, batched):
    from calfkit.engine._lib import CK_ACT_GATE_COMPLETE, CK_ACT_GATE_PASS, CK_ACT_SILENT
    register(engine, [envelope(case["base_state"], case["correlation_id"])])
    envs = [envelope(st["incoming"], st["correlation_id"]) for st in case["steps"]]
    results = []
    if batched:
        out, act = arrive(engine, envs, 100)
        results = [(int(act[i]), out.payload(i)) for i in range(len(envs))]
    else:
        for k, e in enumerate(envs):
            out, act = arrive(engine, [e], 100 + k)
            results.append((int(act[0]), out.payload(0)))
    for st, env, (action, payload) in zip(case["steps"], envs, results):
        if st["correlation_id"] != case["correlation_id"]:
            assert action == CK_ACT_GATE_PASS                   # no pending fan-out for that id: continues with its own state
        elif st["complete"]:
            assert action == CK_ACT_GATE_COMPLETE, case["name"]
            assert state_of(payload) == st["state"], case["name"]
            assert payload == envelope(st["state"], st["correlation_id"])
        else:
            assert action == CK_ACT_SILENT, case["name"]
            assert payload == env                               # handler return of a Silent: the inbound envelope -> publish_topic
    assert engine.gate_stats()["live"] == (0 if case["steps"][-1]["complete"] else 1)


@pytest.mark.parametrize("chunk", [1, 7, 1000], ids=["one_by_one", "chunks_of_7", "one_batch"])
def test_gate_64_way_shuffled_against_oracle(engine, chunk):
    """two concurrent 64-way fan-outs; arrivals shuffled together, with duplicates, a foreign correlation id and late
    arrivals after completion; the device (batched, stamps) must agree with the oracle's one-at-a-time gate."""
    from calfkit import synth
    from calfkit.engine._lib import CK_ACT_GATE_COMPLETE, CK_ACT_GATE_PASS, CK_ACT_SILENT
    from calfkit.models import State
    from calfkit.models.state import PendingToolBatch
    from oracle import port
    rng = random.Random(5)
    bases = synth.fanout_events(2, seed=77, fanout=64)
    register(engine, bases)
    batches, arrivals = {}, []
    for b in bases:
        env = json.loads(b)
        corr = env["context"]["deps"]["correlation_id"]
        st = State.model_validate(env["context"]["state"])
        ids = list(st.tool_calls)
        batches[corr] = PendingToolBatch(expected_tool_call_ids=frozenset(ids), base_state=st.model_copy(deep=True))
        for j, cid in enumerate(ids):
            s2 = st.model_copy(deep=True)
            value = {"return_value": f"It's sunny in city {j} é\n", "content": None, "metadata": {"tool_call_id": cid}, "kind": "tool-return"}
            s2.tool_results[cid] = value if j % 5 else {"raw": [j, 1.5, None], "note": "an untagged value"}
            arrivals.append((corr, s2.model_dump_json()))
        arrivals += rng.sample(arrivals[-64:], 5)                 # duplicates (at-least-once delivery)
    rng.shuffle(arrivals)
    arrivals.insert(3, ("f" * 32, arrivals[0][1]))             # a correlation id nobody is waiting for
    arrivals += [arrivals[10], arrivals[20]]                   # late arrivals: their fan-outs are complete by then
    envs = [envelope(s, c) for c, s in arrivals]
    got = []
    for a in range(0, len(envs), chunk):
        out, act = arrive(engine, envs[a:a + chunk], 1000 + a)
        got += [(int(act[i]), out.payload(i)) for i in range(len(envs[a:a + chunk]))]
    completes = 0
    for (corr, sjson), env, (action, payload) in zip(arrivals, envs, got):
        pending_before = corr in batches
        merged = port.aggregate(batches, State.model_validate_json(sjson), corr)
        if not pending_before:
            assert action == CK_ACT_GATE_PASS
        elif merged is None:
            assert action == CK_ACT_SILENT and payload == env
        else:
            completes += 1
            assert action == CK_ACT_GATE_COMPLETE
            assert payload == envelope(merged.model_dump_json(), corr)
    assert completes == 2 and not batches and engine.gate_stats()["live"] == 0
