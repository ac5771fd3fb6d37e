"""Pins the oracle port (oracle/port.py) to the golden vectors the UNMODIFIED reference produced
(tests/golden/make_golden.py), and the synthetic generator to the codec's fixed points."""
import pytest
from conftest import as_bytes, golden
from pydantic import ValidationError

from oracle import port
import tools_def


@pytest.mark.parametrize("case", golden("codec.json"), ids=lambda c: c["name"])
def test_codec_matches_reference(case):
    payload = as_bytes(case["input"])
    if case["ok"]:
        from calfkit import _ids
        _ids.set_id_source(lambda: f"{1:032x}")   # make_golden's uuid stub starts its counter at 1
        try:
            assert port.encode(port.decode(payload)).decode() == case["output"]
        finally:
            _ids.set_id_source(None)
    else:
        with pytest.raises(ValidationError) as ei:
            port.decode(payload)
        errs = ei.value.errors()
        assert errs[0]["type"] == case["first_type"]
        assert [str(x) for x in errs[0]["loc"]] == case["first_loc"]
        assert len(errs) == case["n_errors"]


@pytest.mark.parametrize("case", golden("tool_node.json"), ids=lambda c: c["name"])
def test_tool_node_matches_reference(case):
    node = port.ToolNode.of(tools_def.TOOLS[case["tool"]])
    payload = as_bytes(case["input"])
    if "raises" in case:
        with pytest.raises(Exception) as ei:
            port.tool_node_event(node, payload)
        assert type(ei.value).__name__ == case["raises"]
        return
    corr = case["publishes"][-1]["correlation_id"]
    got = port.tool_node_event(node, payload, correlation_id=corr)
    want = [(p["topic"], p["key"].encode() if p["key"] is not None else None, p["correlation_id"], p["payload"].encode())
            for p in case["publishes"]]
    assert got == want


def test_actions_match_reference():
    """_publish_action for every action kind, frame ids injected from the same counter."""
    from calfkit import _ids
    from calfkit.models import Call, ReturnCall, Silent, TailCall
    cases = {c["name"]: c for c in golden("actions.json")}
    src = as_bytes(cases["call"]["input"])
    ids = list(port.decode(src).context.state.tool_calls.keys())
    scripts = {
        "call": lambda ctx: Call("tool.tool_00.input", ctx.state, ids[0], "scripted"),
        "call_no_args": lambda ctx: Call("other.input", ctx.state),
        "tailcall": lambda ctx: TailCall("scripted.input", ctx.state),
        "returncall": lambda ctx: ReturnCall(ctx.state),
        "silent": lambda ctx: Silent(),
        "fanout": lambda ctx: [Call(f"tool.tool_{j:02d}.input", ctx.state.model_copy(deep=True), ids[j], "scripted") for j in range(4)],
    }
    for name, script in scripts.items():
        n = [0]

        def det():
            n[0] += 1
            return f"{n[0]:032x}"
        _ids.set_id_source(det)
        try:
            env = port.decode(as_bytes(cases[name]["input"]))
            corr = env.context.deps.correlation_id
            ctx = port.prepare_context(env)
            pubs, ret = port.publish_action("scripted.input", script(ctx), env, corr)
            got = [(t, k.decode(), c, port.encode(e).decode()) for (t, k, c, e) in pubs]
            got.append(("scripted.output", None, corr, port.encode(ret).decode()))
        finally:
            _ids.set_id_source(None)
        want = [(p["topic"], p["key"], p["correlation_id"], p["payload"]) for p in cases[name]["publishes"]]
        assert got == want, name


def test_synth_events_are_codec_fixed_points():
    from calfkit import synth
    recs = (synth.tool_events(50, seed=3) + synth.tool_events(20, seed=4, size=None, full_history=True)
            + synth.fanout_events(3, seed=5, fanout=16) + synth.mixed_events(30, seed=6, hi=30000, n_tools=32))
    for r in recs:
        assert port.encode(port.decode(r)) == r
    lens = [len(r) for r in synth.tool_events(200, seed=7)]
    assert min(lens) >= 1152 - 16 and max(lens) <= 1152 + 16
