#!/usr/bin/env python
"""Compiles the canonical-Envelope schema into the bytecode the device walker interprets
(tests/hostsim/ck_vm.cuh) and writes tests/hostsim/ck_vm_prog.h.

The program is the canonical key order of the reference's wire models (SURVEY.md Appendix A;
calfkit/models/*.py, _vendor/pydantic_ai/messages.py, tools.py) written as a sequence of
"match this literal / recognise this scalar / loop / alternative" steps.  One walker thread
interprets it per record; every literal below is a byte string pydantic's `model_dump_json()`
emits between two values.

    python tests/hostsim/gen_vm.py        # regenerates tests/hostsim/ck_vm_prog.h (committed)
"""
from __future__ import annotations

import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ck_vm_prog.h")
COMMON = os.path.join(HERE, "..", "..", "calfkit-sdk_b200", "csrc", "ck_common.h")

OPS = ["FAIL", "OK", "LIT", "ALT", "PEEKJ", "JMP", "CALL", "RET", "CATCH", "UNCATCH",
       "STR", "STRN", "ANYV", "NUM", "BOOL", "DT", "TAGSCANP",
       "MARK", "SPAN", "SSPAN", "SAVETN", "SPANTN", "DKEYP", "UKEYP", "UKEY_RESET", "FLAG_OR", "FLAG_TEST", "FLAG_CLR",
       "ARGS_BEGIN", "ARGS_LIST", "ARG_ANY", "ARG_STR", "FRAME_DONE", "FRAMES_END", "FIND_CALL", "FIND_RES", "ZERO", "ATEND"]
# ANYV arg: depth | mode << 8   (mode 0: any value, 1: must be an object, 2: object or null)
# NUM  arg: 1 = integer only, 2 = float only
OP = {n: i for i, n in enumerate(OPS)}


def cols() -> dict[str, int]:
    src = open(COMMON).read()
    block = src[src.index("CK_COL_STATUS = 0"):src.index("CK_NUM_COLS")]
    names = re.findall(r"CK_COL_([A-Z0-9_]+)", block)
    return {n: i for i, n in enumerate(names)}


COL = cols()


class Asm:
    def __init__(self):
        self.words: list[int | tuple[str, str]] = []     # ints or ("label", name) fix-ups
        self.labels: dict[str, int] = {}
        self._n = 0

    def new(self, stem: str) -> str:
        self._n += 1
        return f"{stem}_{self._n}"

    def label(self, name: str):
        assert name not in self.labels, name
        self.labels[name] = len(self.words)

    def emit(self, op: str, arg: int = 0):
        # pseudo-ops kept from the schema's point of view, lowered to the interpreter's merged cases
        if op == "ANY":
            op, arg = "ANYV", arg
        elif op == "OBJ":
            op, arg = "ANYV", arg | (1 << 8)
        elif op == "OBJN":
            op, arg = "ANYV", arg | (2 << 8)
        elif op == "INT":
            op, arg = "NUM", 1
        elif op == "FLT":
            op, arg = "NUM", 2
        elif op == "DKEY":
            self.emit("STR"); op = "DKEYP"
        elif op == "UKEY":
            self.emit("STR"); op = "UKEYP"
        elif op == "TAGSCAN":
            self.emit("MARK", 3); self.emit("ANYV", arg | (1 << 8)); op, arg = "TAGSCANP", 0
        elif op == "FRAME_ARGS":
            return self.frame_args(arg)
        assert 0 <= arg < (1 << 24), (op, arg)
        self.words.append(OP[op] | (arg << 8))

    def frame_args(self, depth: int):
        """input_args: null | [ any, ... ]  (captures the count and the first two elements)"""
        self.emit("ARGS_BEGIN")
        lst, done = self.new("ia_list"), self.new("ia_done")
        self.alt("null", lst); self.jmp(done)
        self.label(lst); self.lit("["); self.emit("ARGS_LIST")

        def elem():
            s, nx = self.new("ia_str"), self.new("ia_next")
            self.peekj('"', s)
            self.emit("MARK", 3); self.emit("ANYV", depth); self.emit("ARG_ANY"); self.jmp(nx)
            self.label(s); self.emit("STR"); self.emit("ARG_STR")
            self.label(nx)
        self.list_of(elem)
        self.label(done)

    def ref(self, name: str):
        self.words.append(("label", name))

    # ---- instructions ---------------------------------------------------------------------------
    def _litwords(self, s: str):
        b = s.encode()
        assert 0 < len(b) < 256, s
        for k in range(0, len(b), 8):
            chunk = b[k:k + 8].ljust(8, b"\0")
            v = int.from_bytes(chunk, "little")
            self.words.append(v & 0xFFFFFFFF)
            self.words.append(v >> 32)
        return len(b)

    def lit(self, s: str):
        self.emit("LIT", len(s.encode()))
        self._litwords(s)

    def alt(self, s: str, else_label: str):
        """if the literal is next: consume it and fall through, else jump (nothing consumed)"""
        self.emit("ALT", len(s.encode()))
        self.ref(else_label)
        self._litwords(s)

    def peekj(self, ch: str, target: str):
        self.emit("PEEKJ", ord(ch)); self.ref(target)

    def jmp(self, target: str):
        self.emit("JMP"); self.ref(target)

    def call(self, target: str, ddepth: int):
        self.emit("CALL", ddepth); self.ref(target)

    def catch(self, target: str):
        self.emit("CATCH"); self.ref(target)

    def flag_test(self, mask: int, target: str):
        self.emit("FLAG_TEST", mask); self.ref(target)

    def null_or(self, body):
        """value is `null` or whatever `body()` emits"""
        done, notnull = self.new("nn_done"), self.new("nn_val")
        self.alt("null", notnull); self.jmp(done)
        self.label(notnull); body()
        self.label(done)

    def list_of(self, elem):
        """[ elem (, elem)* ] or []   — the opening '[' is part of the preceding literal"""
        body, end = self.new("l_body"), self.new("l_end")
        self.peekj("]", end)
        self.label(body); elem()
        close = self.new("l_close")
        self.alt(",", close); self.jmp(body)
        self.label(close)
        self.label(end); self.lit("]")

    def span_field(self, col: str, body):
        self.emit("MARK", 0); body(); self.emit("SPAN", COL[col] | (0 << 12))

    def assemble(self) -> list[int]:
        out = []
        for w in self.words:
            if isinstance(w, tuple):
                out.append(self.labels[w[1]])
            else:
                out.append(w)
        return out


def build() -> tuple[list[int], dict[str, int]]:
    a = Asm()
    D = lambda k: k            # depth offsets are relative to the current depth base

    # ================================================================= main: Envelope
    a.label("main")
    a.lit('{"context":{"state":{"tool_calls":{')
    # ---- tool_calls: dict[str, ToolCallPart]
    tc_end, tc_body = a.new("tc_end"), a.new("tc_body")
    a.emit("MARK", 1)                                # R1 = pos (just after '{'); span starts one before
    a.peekj("}", tc_end)
    a.label(tc_body)
    a.emit("DKEY", 0)
    a.call("tcp", 5); a.lit('tool-call"}')
    nxt = a.new("tc_close")
    a.alt(",", nxt); a.jmp(tc_body)
    a.label(nxt)
    a.label(tc_end); a.lit("}")
    a.emit("SPAN", COL["TC_OFF"] | (1 << 12) | (1 << 14))     # span from R1-1

    a.lit(',"tool_results":{')
    tr_end, tr_body = a.new("tr_end"), a.new("tr_body")
    a.emit("MARK", 1)
    a.peekj("}", tr_end)
    a.label(tr_body)
    a.emit("DKEY", 1)
    a.call("trv", 5)
    nxt = a.new("tr_close")
    a.alt(",", nxt); a.jmp(tr_body)
    a.label(nxt)
    a.label(tr_end); a.lit("}")
    a.emit("SPAN", COL["TR_OFF"] | (1 << 12) | (1 << 14))

    a.lit(',"uncommitted_message":')
    a.span_field("UNC_OFF", lambda: a.null_or(lambda: a.call("msg", 4)))
    a.lit(',"message_history":[')
    a.emit("MARK", 1)
    a.list_of(lambda: a.call("msg", 5))
    a.emit("SPAN", COL["HIST_OFF"] | (1 << 12) | (1 << 14))
    a.lit(',"final_output_parts":[')
    a.emit("MARK", 1)
    a.list_of(lambda: a.call("cpart", 5))
    a.emit("SPAN", COL["FOP_OFF"] | (1 << 12) | (1 << 14))
    a.lit(',"temp_instructions":')
    a.span_field("TI_OFF", lambda: a.emit("STRN"))
    a.lit(',"metadata":')
    a.span_field("SMETA_OFF", lambda: a.emit("ANY", 4))
    a.lit(',"overrides":')
    a.span_field("SOV_OFF", lambda: a.call("ovr", 4))
    a.lit('},"deps":{"correlation_id":')
    a.emit("STR"); a.emit("SSPAN", COL["CORR_OFF"])
    a.lit(',"provided_deps":')
    a.span_field("PD_OFF", lambda: a.emit("OBJ", 4))
    a.lit('}},"internal_workflow_state":{"call_stack":{"_internal_list":[')
    # ---- frames
    a.emit("MARK", 1)
    fr_end, fr_body = a.new("fr_end"), a.new("fr_body")
    a.peekj("]", fr_end)
    a.label(fr_body)
    a.emit("MARK", 0)
    a.lit('{"target_topic":'); a.emit("STR"); a.emit("SSPAN", COL["TGT_OFF"])
    a.lit(',"callback_topic":'); a.emit("STR"); a.emit("SSPAN", COL["CB_OFF"])
    a.lit(',"input_args":'); a.emit("FRAME_ARGS", 6)
    a.lit(',"frame_id":'); a.emit("STR")
    a.lit(',"overrides":'); a.emit("MARK", 2); a.call("ovr", 6); a.emit("SPAN", COL["FOV_OFF"] | (2 << 12))
    a.lit("}")
    a.emit("SPAN", COL["TOP_OFF"] | (0 << 12))
    a.emit("FRAME_DONE")
    nxt = a.new("fr_close")
    a.alt(",", nxt); a.jmp(fr_body)
    a.label(nxt)
    a.label(fr_end); a.lit("]")
    a.emit("SPAN", COL["FRAMES_OFF"] | (1 << 12) | (1 << 14))
    a.emit("FRAMES_END")                              # NFRAMES, NARGS, ARGKINDS, ARG0/1 (+ zero frame columns if empty)
    a.lit('},"metadata":')
    a.span_field("WFMETA_OFF", lambda: a.emit("ANY", 3))
    a.lit("}}")
    a.emit("ATEND")                                   # pos must equal the record length
    # ---- resolve tool_calls[input_args[0]] / tool_results[input_args[0]]
    no_call, no_res = a.new("no_call"), a.new("no_res")
    a.emit("FIND_CALL"); a.ref(no_call)               # found: pos = value start, R0 = pos
    a.call("tcp", 5); a.lit('tool-call"}')
    a.emit("SPAN", COL["CALL_VAL_OFF"] | (0 << 12)); a.emit("SPANTN")
    a.jmp(a_after := a.new("after_call"))
    a.label(no_call); a.emit("ZERO", COL["CALL_VAL_OFF"] | (6 << 12))
    a.label(a_after)
    a.emit("FIND_RES"); a.ref(no_res)
    a.call("trv", 5)
    a.emit("SPAN", COL["RES_OFF"] | (0 << 12))
    a.emit("OK")
    a.label(no_res); a.emit("ZERO", COL["RES_OFF"] | (2 << 12))
    a.emit("OK")

    # ================================================================= tcp: ToolCallPart up to `,"part_kind":"`
    # (messages.py:1187-1283)  tool_name, args, tool_call_id, id, provider_name, provider_details
    a.label("tcp")
    a.lit('{"tool_name":'); a.emit("STR"); a.emit("SAVETN")
    a.lit(',"args":'); a.emit("MARK", 2)
    s_args, a_done = a.new("args_str"), a.new("args_done")
    a.peekj('"', s_args); a.emit("OBJN", 1); a.jmp(a_done)
    a.label(s_args); a.emit("STR")
    a.label(a_done); a.emit("MARK", 3)
    a.lit(',"tool_call_id":'); a.emit("STR")
    a.lit(',"id":'); a.emit("STRN")
    a.lit(',"provider_name":'); a.emit("STRN")
    a.lit(',"provider_details":'); a.emit("OBJN", 1)
    a.lit(',"part_kind":"')
    a.emit("RET")

    # ================================================================= trv: tool_results value
    # ToolReturn | ModelRetry | RetryPromptPart (tagged) | Any   (models/state.py:70, tools.py:189-210)
    a.label("trv")
    t_obj, t_gen = a.new("trv_obj"), a.new("trv_gen")
    a.peekj("{", t_obj)
    a.emit("ANY", 0); a.emit("RET")
    a.label(t_obj)
    a.catch(t_gen)
    l2, l3, l4 = a.new("trv2"), a.new("trv3"), a.new("trv4")
    a.alt('{"return_value":', l2)
    a.emit("ANY", 1); a.lit(',"content":'); a.emit("STRN"); a.lit(',"metadata":'); a.emit("ANY", 1)
    a.lit(',"kind":"tool-return"}'); a.emit("UNCATCH"); a.emit("RET")
    a.label(l2); a.alt('{"message":', l3)
    a.emit("STR"); a.lit(',"kind":"model-retry"}'); a.emit("UNCATCH"); a.emit("RET")
    a.label(l3); a.alt('{"content":', l4)
    a.emit("STR"); a.lit(',"tool_name":'); a.emit("STRN"); a.lit(',"tool_call_id":'); a.emit("STR")
    a.lit(',"timestamp":'); a.emit("DT"); a.lit(',"part_kind":"retry-prompt"}'); a.emit("UNCATCH"); a.emit("RET")
    a.label(l4); a.emit("UNCATCH")
    a.label(t_gen)                                    # not the canonical form of a tagged model
    a.emit("TAGSCAN", 0); a.emit("RET")

    # ================================================================= msg: ModelRequest | ModelResponse
    a.label("msg")
    a.emit("FLAG_CLR", 7)
    a.lit('{"parts":[')
    a.list_of(lambda: a.call("part", 2))
    resp = a.new("msg_resp")
    a.alt(',"timestamp":', resp)
    bad = a.new("msg_bad")
    a.flag_test(2, bad)                               # a response-side part inside a request
    a.null_or(lambda: a.emit("DT"))
    a.lit(',"instructions":'); a.emit("STRN")
    a.lit(',"kind":"request","run_id":'); a.emit("STRN")
    a.lit(',"metadata":'); a.emit("OBJN", 1); a.lit("}")
    a.emit("RET")
    a.label(resp)
    a.flag_test(1, bad)
    a.lit(',"usage":{"input_tokens":'); a.emit("INT")
    for k in ["cache_write_tokens", "cache_read_tokens", "output_tokens", "input_audio_tokens", "cache_audio_read_tokens",
              "output_audio_tokens"]:
        a.lit(f',"{k}":'); a.emit("INT")
    a.lit(',"details":{')
    u_end, u_body = a.new("u_end"), a.new("u_body")
    a.emit("UKEY_RESET")
    a.peekj("}", u_end)
    a.label(u_body); a.emit("UKEY"); a.emit("INT")
    nxt = a.new("u_close")
    a.alt(",", nxt); a.jmp(u_body)
    a.label(nxt)
    a.label(u_end); a.lit("}}")
    a.lit(',"model_name":'); a.emit("STRN")
    a.lit(',"name":'); a.emit("STRN")
    a.lit(',"timestamp":'); a.emit("DT")
    a.lit(',"kind":"response","provider_name":'); a.emit("STRN")
    a.lit(',"provider_url":'); a.emit("STRN")
    a.lit(',"provider_details":'); a.emit("OBJN", 1)
    a.lit(',"provider_response_id":'); a.emit("STRN")
    a.lit(',"finish_reason":')
    fr_done = a.new("fin_done")
    for i, v in enumerate(["null", '"stop"', '"length"', '"content_filter"', '"tool_call"']):
        nx = a.new("fin")
        a.alt(v, nx); a.jmp(fr_done); a.label(nx)
    a.lit('"error"')
    a.label(fr_done)
    a.lit(',"run_id":'); a.emit("STRN")
    a.lit(',"metadata":'); a.emit("OBJN", 1); a.lit("}")
    a.emit("RET")
    a.label(bad); a.emit("FAIL")

    # ================================================================= part: one message part
    # sets flag 1 for request-side parts, 2 for response-side parts
    a.label("part")
    a.emit("FLAG_CLR", 4)
    p_tool = a.new("part_toolname")
    a.alt('{"content":', p_tool)
    c_list, c_after = a.new("c_list"), a.new("c_after")
    a.peekj("[", c_list)
    a.emit("STR"); a.jmp(c_after)
    a.label(c_list); a.lit("["); a.list_of(lambda: a.emit("STR")); a.emit("FLAG_OR", 4)      # flag 4: content was list[str]
    a.label(c_after)
    no_ts = a.new("no_ts")
    a.alt(',"timestamp":', no_ts)
    a.emit("DT")
    usr = a.new("user_part")
    a.alt(',"dynamic_ref":', usr)
    a.flag_test(4, "fail")
    a.emit("STRN"); a.lit(',"name":'); a.emit("STRN"); a.lit(',"part_kind":"system-prompt"}')
    a.emit("FLAG_OR", 1); a.emit("RET")
    a.label(usr)
    a.lit(',"name":'); a.emit("STRN"); a.lit(',"part_kind":"user-prompt"}')
    a.emit("FLAG_OR", 1); a.emit("RET")
    a.label(no_ts)
    a.flag_test(4, "fail")
    no_retry = a.new("no_retry")
    a.alt(',"tool_name":', no_retry)
    a.emit("STRN"); a.lit(',"tool_call_id":'); a.emit("STR"); a.lit(',"timestamp":'); a.emit("DT")
    a.lit(',"part_kind":"retry-prompt"}'); a.emit("FLAG_OR", 1); a.emit("RET")
    a.label(no_retry)
    a.lit(',"id":'); a.emit("STRN")
    txt = a.new("text_part")
    a.alt(',"signature":', txt)
    a.emit("STRN"); a.lit(',"provider_name":'); a.emit("STRN"); a.lit(',"provider_details":'); a.emit("OBJN", 1)
    a.lit(',"part_kind":"thinking"}'); a.emit("FLAG_OR", 2); a.emit("RET")
    a.label(txt)
    a.lit(',"provider_name":'); a.emit("STRN"); a.lit(',"provider_details":'); a.emit("OBJN", 1)
    a.lit(',"part_kind":"text"}'); a.emit("FLAG_OR", 2); a.emit("RET")
    # tool_name-first parts
    a.label(p_tool)
    a.catch("part_tcp")                               # not `{"tool_name":S,"content":` -> it is a tool-call part
    a.lit('{"tool_name":'); a.emit("STR")
    a.lit(',"content":'); a.emit("UNCATCH")
    a.emit("ANY", 1); a.lit(',"tool_call_id":'); a.emit("STR"); a.lit(',"metadata":'); a.emit("ANY", 1)
    a.lit(',"timestamp":'); a.emit("DT")
    blt = a.new("builtin_ret")
    a.alt(',"part_kind":"tool-return"}', blt); a.emit("FLAG_OR", 1); a.emit("RET")
    a.label(blt)
    a.lit(',"provider_name":'); a.emit("STRN"); a.lit(',"provider_details":'); a.emit("OBJN", 1)
    a.lit(',"part_kind":"builtin-tool-return"}'); a.emit("FLAG_OR", 2); a.emit("RET")
    a.label("part_tcp")
    a.call("tcp", 0)
    b2 = a.new("builtin_call")
    a.alt('tool-call"}', b2); a.emit("FLAG_OR", 2); a.emit("RET")
    a.label(b2); a.lit('builtin-tool-call"}'); a.emit("FLAG_OR", 2); a.emit("RET")
    a.label("fail"); a.emit("FAIL")

    # ================================================================= cpart: final_output_parts element (payload.py:6-35)
    a.label("cpart")
    a.lit('{"kind":"')
    c2, c3, c4 = a.new("cp2"), a.new("cp3"), a.new("cp4")
    a.alt('text","text":', c2); a.emit("STR"); a.lit(',"metadata":'); a.emit("OBJN", 1); a.lit("}"); a.emit("RET")
    a.label(c2); a.alt('data","data":', c3); a.emit("ANY", 1); a.lit(',"schema_":null,"metadata":'); a.emit("OBJN", 1)
    a.lit("}"); a.emit("RET")
    a.label(c3); a.alt('file","media_type":', c4); a.emit("STR"); a.lit(',"uri":'); a.emit("STRN"); a.lit(',"data":')
    a.emit("STRN"); a.lit(',"metadata":'); a.emit("OBJN", 1); a.lit("}"); a.emit("RET")
    a.label(c4); a.lit('tool","tool_call_id":'); a.emit("STR"); a.lit(',"kwargs":'); a.emit("OBJ", 1)
    a.lit(',"tool_name":'); a.emit("STR"); a.lit(',"metadata":'); a.emit("OBJN", 1); a.lit("}"); a.emit("RET")

    # ================================================================= ovr: OverridesState | null
    a.label("ovr")
    o_obj = a.new("ovr_obj")
    a.alt("null", o_obj); a.emit("RET")
    a.label(o_obj)
    a.lit('{"override_agent_tools":')
    o_list = a.new("ovr_list")
    a.alt("null", o_list); a.lit("}"); a.emit("RET")
    a.label(o_list)
    a.lit("[")

    def tool_schema():
        a.lit('{"node_id":'); a.emit("STR"); a.lit(',"subscribe_topics":[')
        a.list_of(lambda: a.emit("STR"))
        a.lit(',"publish_topic":'); a.emit("STRN")
        a.lit(',"tool_schema":{"name":'); a.emit("STR")
        a.lit(',"parameters_json_schema":'); a.emit("OBJ", 4)
        a.lit(',"description":'); a.emit("STRN")
        a.lit(',"outer_typed_dict_key":'); a.emit("STRN")
        a.lit(',"strict":'); a.null_or(lambda: a.emit("BOOL"))
        a.lit(',"sequential":'); a.emit("BOOL")
        a.lit(',"kind":"')
        kd = a.new("kind_done")
        for v in ['function"', 'output"', 'external"']:
            nx = a.new("kind")
            a.alt(v, nx); a.jmp(kd); a.label(nx)
        a.lit('unapproved"')
        a.label(kd)
        a.lit(',"metadata":'); a.emit("OBJN", 4)
        a.lit(',"timeout":'); a.null_or(lambda: a.emit("FLT"))
        a.lit("}}")
    a.list_of(tool_schema)
    a.lit("}")
    a.emit("RET")

    return a.assemble(), a.labels


def main():
    import io
    words, labels = build()
    with io.StringIO() as f:
        f.write("// GENERATED by tests/hostsim/gen_vm.py — do not edit.\n")
        f.write("// Bytecode of the canonical-Envelope schema (one u32 per word; op = low 8 bits, arg = high 24).\n")
        f.write("#ifndef CK_VM_PROG_H\n#define CK_VM_PROG_H\n#include <stdint.h>\n\n")
        f.write("enum {\n" + "".join(f"    VOP_{n} = {i},\n" for n, i in OP.items()) + "};\n\n")
        f.write(f"#define CK_VM_PROG_WORDS {len(words)}\n#define CK_VM_ENTRY {labels['main']}\n\n")

        def le(b: bytes) -> str:
            return f"0x{int.from_bytes(b, 'little'):x}ull"
        for name, s in [("TOOLRET", b'"tool-return"'), ("MODELRETRY", b'"model-retry"'), ("RETRYPROMPT", b'"retry-prompt"')]:
            f.write(f"#define VM_TAG_{name}_A {le(s[:8])}\n#define VM_TAG_{name}_B {le(s[-8:])}\n#define VM_TAG_{name}_LEN {len(s)}\n")
        f.write(f"#define VM_KEY_KIND {le(b'kind')}\n#define VM_KEY_PARTKIN {le(b'part_kin')}\n\n")
        f.write("#define CK_VM_PROG_INIT { \\\n")
        for i in range(0, len(words), 8):
            f.write("    " + ", ".join(f"0x{w:08x}u" for w in words[i:i + 8]) + ", \\\n")
        f.write("}\n\nstatic const uint32_t ck_vm_prog_host[CK_VM_PROG_WORDS] = CK_VM_PROG_INIT;\n\n#endif\n")
        text = f.getvalue()
    old = open(OUT).read() if os.path.exists(OUT) else None
    if text != old:                      # an unchanged header keeps its mtime: no needless rebuild of the library
        with open(OUT, "w") as g:
            g.write(text)

    print(f"{OUT}: {len(words)} words, {len(labels)} labels")


if __name__ == "__main__":
    main()
