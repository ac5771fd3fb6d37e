"""Seeded synthetic agent-event generator — shared by bench.py, smoke() and tests.

Events are emitted directly as canonical wire bytes (the fixed point of the reference's
`Envelope.model_dump_json()`) from string templates, so a million of them
can be built in seconds without constructing pydantic objects.  tests/test_oracle.py proves the
templates are fixed points of the reference codec (golden vectors) and of the oracle port.

Workloads
  tool_event_1k    config 2 headline: tool-stage envelope for one @agent_tool node in the metric's
                   1 KB class: tool_calls{1}, history = [user request], 2 frames, short ids.  The
                   smallest valid such record is ~1130 B, so records are padded (provided_deps) to
                   1152 +- 16 B — slightly MORE bytes per event than a nominal 1024, i.e. the
                   events/s figure is conservative.
  tool_event_full  same hop with the ModelResponse that carried the tool call also in history
                   (what the reference's Agent really sends, ~1.7 KB) — parity shape.
  fanout_event     config 3: post-LLM agent-stage envelope with F pending tool calls.
  mixed            config 5: sizes log-uniform 128 B .. 64 KB, many topics, UTF-8 and escapes.
"""
from __future__ import annotations

import json
from dataclasses import dataclass

import numpy as np

CITIES = ["Tokyo", "Paris", "São Paulo", "New York", "Kraków", "北京", "Reykjavík", "Nairobi",
          "San Francisco", "Zürich", "Ho Chi Minh City", "Москва", "Lima", "Oslo", "Cairo", "Quito"]
PROMPT_BITS = ["What's the weather in {c}?", "Is it raining in {c} right now?",
               "Weather report for {c}, please — include \"feels like\".",
               "Tell me:\n\tforecast for {c}\\today", "¿Qué tiempo hace en {c}? 🌦"]
TS = "2026-01-01T00:00:00Z"


def _hex(rng: np.random.Generator, n: int, width: int = 32) -> list[str]:
    words = rng.integers(0, 1 << 63, size=(n, (width + 15) // 16), dtype=np.int64)
    return ["".join(f"{int(w):016x}" for w in row)[:width] for row in words]


def jstr(s: str) -> str:
    """Canonical JSON string content incl. quotes (same escapes pydantic-core emits:
    \\" \\\\ \\n \\t \\r \\b \\f, other <0x20 as \\u00xx lowercase, everything else raw)."""
    return json.dumps(s, ensure_ascii=False)


def tool_call_part(tool_name: str, args_json: str, tool_call_id: str) -> str:
    return ('{"tool_name":' + jstr(tool_name) + ',"args":' + args_json + ',"tool_call_id":' + jstr(tool_call_id)
            + ',"id":null,"provider_name":null,"provider_details":null,"part_kind":"tool-call"}')


def user_request(prompt: str, ts: str = TS) -> str:
    return ('{"parts":[{"content":' + jstr(prompt) + ',"timestamp":"' + ts + '","name":null,"part_kind":"user-prompt"}],'
            '"timestamp":"' + ts + '","instructions":null,"kind":"request","run_id":null,"metadata":null}')


def model_response(parts_json: list[str], model_name: str = "function:synthetic", ts: str = TS) -> str:
    return ('{"parts":[' + ",".join(parts_json) + '],"usage":{"input_tokens":51,"cache_write_tokens":0,'
            '"cache_read_tokens":0,"output_tokens":7,"input_audio_tokens":0,"cache_audio_read_tokens":0,'
            '"output_audio_tokens":0,"details":{}},"model_name":' + jstr(model_name) + ',"name":null,"timestamp":"' + ts + '",'
            '"kind":"response","provider_name":null,"provider_url":null,"provider_details":null,'
            '"provider_response_id":null,"finish_reason":null,"run_id":null,"metadata":null}')


def frame(target: str, callback: str, input_args: list[str] | None, frame_id: str, overrides: str = "null") -> str:
    ia = "null" if input_args is None else "[" + ",".join(jstr(a) for a in input_args) + "]"
    return ('{"target_topic":' + jstr(target) + ',"callback_topic":' + jstr(callback) + ',"input_args":' + ia
            + ',"frame_id":"' + frame_id + '","overrides":' + overrides + '}')


def envelope(*, tool_calls: dict[str, str], tool_results: dict[str, str], uncommitted: str, history: list[str],
             final_parts: list[str], temp_instructions: str | None, state_metadata: str, state_overrides: str,
             correlation_id: str, provided_deps: str, frames: list[str], wf_metadata: str = "null") -> str:
    tc = ",".join(jstr(k) + ":" + v for k, v in tool_calls.items())
    tr = ",".join(jstr(k) + ":" + v for k, v in tool_results.items())
    ti = "null" if temp_instructions is None else jstr(temp_instructions)
    return ('{"context":{"state":{"tool_calls":{' + tc + '},"tool_results":{' + tr + '},"uncommitted_message":' + uncommitted
            + ',"message_history":[' + ",".join(history) + '],"final_output_parts":[' + ",".join(final_parts)
            + '],"temp_instructions":' + ti + ',"metadata":' + state_metadata + ',"overrides":' + state_overrides
            + '},"deps":{"correlation_id":' + jstr(correlation_id) + ',"provided_deps":' + provided_deps
            + '}},"internal_workflow_state":{"call_stack":{"_internal_list":[' + ",".join(frames)
            + ']},"metadata":' + wf_metadata + '}}')


@dataclass
class Batch:
    """Concatenated records + (n+1) int64 offsets: the layout ck_submit takes."""
    data: np.ndarray      # uint8 [total]
    offsets: np.ndarray   # int64 [n+1]

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    def record(self, i: int) -> bytes:
        return self.data[self.offsets[i]:self.offsets[i + 1]].tobytes()

    def records(self) -> list[bytes]:
        return [self.record(i) for i in range(self.n)]


def pack(records: list[bytes]) -> Batch:
    lens = np.fromiter((len(r) for r in records), dtype=np.int64, count=len(records))
    offsets = np.zeros(len(records) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    data = np.frombuffer(b"".join(records), dtype=np.uint8)
    return Batch(data=data, offsets=offsets)


def tool_events(n: int, seed: int = 0, *, size: int | None = 1152, jitter: int = 16, full_history: bool = False,
                tool_name: str = "get_weather", agent_topic: str = "weather_agent.input",
                agent_name: str = "weather_agent", n_tools: int = 1, compact: bool | None = None, n_agents: int = 1) -> list[bytes]:
    """config-2 records.  `size`=None leaves records unpadded; otherwise provided_deps carries a
    pad string so that len(record) = size + U[-jitter, +jitter] (clamped to the unpadded size).
    With n_tools > 1 records are spread over tools tool_00..tool_{n-1}; with n_agents > 1 their callback topics are spread
    over agent_000.input .. (config 5 routing: the route step resolves among that many destination topics)."""
    if compact is None:
        compact = size is not None and size <= 1200   # short ids: smallest valid record ~1130 B
    rng = np.random.default_rng(seed)
    corr = _hex(rng, n)
    tcid = _hex(rng, n)
    fid0 = _hex(rng, n)
    fid1 = _hex(rng, n)
    client = _hex(rng, n, 16)
    city_i = rng.integers(0, len(CITIES), size=n)
    prompt_i = rng.integers(0, 2 if compact else len(PROMPT_BITS), size=n)
    tool_i = rng.integers(0, n_tools, size=n)
    agent_i = rng.integers(0, n_agents, size=n) if n_agents > 1 else None
    agent_topic0 = agent_topic
    jit = rng.integers(-jitter, jitter + 1, size=n) if jitter else np.zeros(n, dtype=np.int64)
    out: list[bytes] = []
    for i in range(n):
        city = CITIES[city_i[i]]
        tname = tool_name if n_tools == 1 else f"tool_{int(tool_i[i]):02d}"
        agent_topic = agent_topic0 if agent_i is None else f"agent_{int(agent_i[i]):03d}.input"
        # OpenAI-style short ids in the 1 KB class, pydantic-ai generated ids otherwise
        call_id = ("call_" + tcid[i][:12]) if compact else ("pyd_ai_" + tcid[i])
        args = '{"location":' + jstr(city) + '}'
        tcp = tool_call_part(tname, args, call_id)
        hist = [user_request(PROMPT_BITS[prompt_i[i]].format(c=city))]
        if full_history:
            hist.append(model_response([tcp]))
        frames = [frame(agent_topic, "calf-client-reply-" + (client[i][:8] if compact else client[i]), None, fid0[i]),
                  frame(f"tool.{tname}.input", agent_topic, [call_id, agent_name], fid1[i])]

        def build(pad: str) -> bytes:
            deps = ('{"pad":"%s"}' % pad) if compact else ('{"tenant":"t-%04d","pad":"%s"}' % (i % 10000, pad))
            return envelope(tool_calls={call_id: tcp}, tool_results={}, uncommitted="null", history=hist,
                            final_parts=[], temp_instructions=None, state_metadata="null", state_overrides="null",
                            correlation_id=corr[i], provided_deps=deps, frames=frames).encode()

        rec = build("")
        if size is not None:
            want = size + int(jit[i])
            if want > len(rec):
                rec = build("x" * (want - len(rec)))
        out.append(rec)
    return out


def fanout_events(n: int, seed: int = 0, *, fanout: int = 64, agent_topic: str = "planner.input",
                  agent_name: str = "planner") -> list[bytes]:
    """config-3 records: agent-stage envelope after the LLM asked for `fanout` tool calls
    (tools tool_00..), i.e. the state `Agent.run` holds at reference nodes/agent.py:195."""
    rng = np.random.default_rng(seed)
    corr = _hex(rng, n)
    fid0 = _hex(rng, n)
    client = _hex(rng, n, 16)
    out: list[bytes] = []
    for i in range(n):
        ids = _hex(rng, fanout)
        parts, tcs = [], {}
        for j in range(fanout):
            cid = "pyd_ai_" + ids[j]
            p = tool_call_part(f"tool_{j:02d}", '{"location":' + jstr(CITIES[(i + j) % len(CITIES)]) + '}', cid)
            parts.append(p)
            tcs[cid] = p
        hist = [user_request(f"Compare the weather in {fanout} cities."), model_response(parts)]
        frames = [frame(agent_topic, "calf-client-reply-" + client[i], None, fid0[i])]
        out.append(envelope(tool_calls=tcs, tool_results={}, uncommitted="null", history=hist, final_parts=[],
                            temp_instructions=None, state_metadata="null", state_overrides="null",
                            correlation_id=corr[i], provided_deps="{}", frames=frames).encode())
    return out


def mixed_events(n: int, seed: int = 0, *, lo: int = 128, hi: int = 65536, n_tools: int = 256, n_agents: int = 1) -> list[bytes]:
    """config-5 records: log-uniform sizes; long records get long multi-turn histories with
    escapes and multi-byte UTF-8, short ones are stripped to the minimum envelope."""
    rng = np.random.default_rng(seed)
    sizes = np.exp(rng.uniform(np.log(lo), np.log(hi), size=n)).astype(np.int64)
    base = tool_events(n, seed + 1, size=None, n_tools=n_tools, n_agents=n_agents)
    out: list[bytes] = []
    filler = "The quick brown fox — «jumps» over\tthe lazy dog.\n\"quoted\" back\\slash ünïcödé 漢字 🙂 "
    for i in range(n):
        rec = base[i]
        if sizes[i] > len(rec) + 400:
            # grow history with extra request/response turns carrying text
            turns = []
            budget = int(sizes[i]) - len(rec)
            k = 0
            while budget > 0:
                take = min(budget, 1500 + 37 * (k % 11))
                txt = (filler * (take // len(filler) + 1))[:take]
                turns.append(user_request(txt))
                txt_part = ('{"content":' + jstr(txt[: take // 2]) + ',"id":null,"provider_name":null,'
                            '"provider_details":null,"part_kind":"text"}')
                turns.append(model_response([txt_part]))
                budget -= 2 * take + 900
                k += 1
            s = rec.decode()
            marker = '"message_history":['
            p = s.index(marker) + len(marker)
            rec = (s[:p] + ",".join(turns) + "," + s[p:]).encode()
        out.append(rec)
    return out
