// Device canonicaliser: the general half of the codec.
//
// The fast path (ck_walk.cuh / ck_vm.cuh) only recognises records that are already byte-wise fixed points
// of the reference codec.  Everything else that `Envelope.model_validate_json` (reference
// calfkit/models/envelope.py:9-17, pydantic-core / jiter) accepts — whitespace, any key order, missing
// defaults, unknown keys, duplicate keys, \uXXXX / \/ escapes, exponent-form numbers, validation aliases —
// is handled here: one thread per record re-emits the record in the canonical form `model_dump_json()`
// would produce, after which it goes through the same splice kernels as any other record.  Verdicts:
//   CK_OK              canonical bytes written (byte-identical to dump(validate(input)))
//   CK_JSON_INVALID    jiter would reject the text (pydantic error type json_invalid)
//   CK_SCHEMA_INVALID  well-formed JSON that violates the Envelope schema (missing / *_type /
//                      union_tag_* / literal_error ...)
//   CK_UNSUPPORTED     constructs whose result cannot be decided / reproduced on the device yet
//                      (floats of more than 19 digits, fractional unix timestamps, exotic datetimes, default_factory fields
//                      that are absent, multi-modal content ...) — reported per record, never guessed.
// Soundness contract (fuzzed against pydantic through tests/hostsim): OK implies identical bytes,
// JSON_INVALID / SCHEMA_INVALID imply pydantic raises with that class.
#ifndef CK_CANON_CUH
#define CK_CANON_CUH

#include "ck_walk.cuh"
#include "ck_float.cuh"

#if defined(__CUDACC__)
#define CK_HDR __host__ __device__          // recursive (mutually) functions: no forced inlining
#else
#define CK_HDR
#endif
#define CJ_MAX_DEPTH 200                 // jiter: a value enclosed by more than 200 containers is rejected (probes in DESIGN.md)

struct CIn { const u8* p; u32 n; };
struct COut {
    u8* p; u32 cap, len; bool ovf;
    CK_HD void put(u8 b) { if (len < cap) p[len] = b; else ovf = true; len++; }
    CK_HD void puts(const char* s, u32 n) { for (u32 i = 0; i < n; i++) put((u8)s[i]); }
    CK_HD void copy(const CIn& in, u32 a, u32 b) { for (u32 i = a; i < b; i++) put(in.p[i]); }
};
#define CPUTS(o, lit) (o).puts(lit, (u32)(sizeof(lit) - 1))

CK_HD bool cj_ws(u8 c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
CK_HD u32 cj_skip_ws(const CIn& in, u32 pos) { while (pos < in.n && cj_ws(in.p[pos])) pos++; return pos; }
CK_HD int cj_hex(u8 c) { return (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1; }

// ---------------------------------------------------------------------------------------------------------
// Phase A: JSON syntax exactly as jiter accepts it.  Returns end position of the string / number, 0 = bad.
// ---------------------------------------------------------------------------------------------------------
CK_HD u32 cj_string_end(const CIn& in, u32 pos) {         // pos at the opening quote
    pos++;
    for (;;) {
        if (pos >= in.n) return 0;
        u8 c = in.p[pos];
        if (c == '"') return pos + 1;
        if (c < 0x20) return 0;
        if (c == '\\') {
            if (pos + 1 >= in.n) return 0;
            u8 e = in.p[pos + 1];
            if (e == '"' || e == '\\' || e == '/' || e == 'b' || e == 'f' || e == 'n' || e == 'r' || e == 't') { pos += 2; continue; }
            if (e != 'u' || pos + 5 >= in.n) return 0;
            int v = 0;
            for (u32 k = 2; k < 6; k++) { int h = cj_hex(in.p[pos + k]); if (h < 0) return 0; v = v * 16 + h; }
            pos += 6;
            if (v >= 0xDC00 && v <= 0xDFFF) return 0;                       // lone trailing surrogate
            if (v >= 0xD800 && v <= 0xDBFF) {                               // must be followed by \uDC00..\uDFFF
                if (pos + 5 >= in.n || in.p[pos] != '\\' || in.p[pos + 1] != 'u') return 0;
                int w = 0;
                for (u32 k = 2; k < 6; k++) { int h = cj_hex(in.p[pos + k]); if (h < 0) return 0; w = w * 16 + h; }
                if (w < 0xDC00 || w > 0xDFFF) return 0;
                pos += 6;
            }
            continue;
        }
        if (c < 0x80) { pos++; continue; }
        // UTF-8 sequence
        u32 need; u8 lo = 0x80, hi = 0xBF;
        if (c >= 0xC2 && c <= 0xDF) need = 1;
        else if (c == 0xE0) { need = 2; lo = 0xA0; }
        else if (c >= 0xE1 && c <= 0xEC) need = 2;
        else if (c == 0xED) { need = 2; hi = 0x9F; }
        else if (c >= 0xEE && c <= 0xEF) need = 2;
        else if (c == 0xF0) { need = 3; lo = 0x90; }
        else if (c >= 0xF1 && c <= 0xF3) need = 3;
        else if (c == 0xF4) { need = 3; hi = 0x8F; }
        else return 0;
        if (pos + need >= in.n) return 0;
        if (in.p[pos + 1] < lo || in.p[pos + 1] > hi) return 0;
        for (u32 k = 2; k <= need; k++) if (in.p[pos + k] < 0x80 || in.p[pos + k] > 0xBF) return 0;
        pos += need + 1;
    }
}

CK_HD bool cj_isdigit(u8 c) { return c >= '0' && c <= '9'; }
CK_HD u32 cj_number_end(const CIn& in, u32 pos) {          // JSON number grammar; 0 = bad
    u32 p = pos;
    if (p < in.n && in.p[p] == '-') p++;
    if (p >= in.n || !cj_isdigit(in.p[p])) return 0;
    if (in.p[p] == '0') { p++; if (p < in.n && cj_isdigit(in.p[p])) return 0; }
    else while (p < in.n && cj_isdigit(in.p[p])) p++;
    if (p < in.n && in.p[p] == '.') {
        p++;
        if (p >= in.n || !cj_isdigit(in.p[p])) return 0;
        while (p < in.n && cj_isdigit(in.p[p])) p++;
    }
    if (p < in.n && (in.p[p] == 'e' || in.p[p] == 'E')) {
        p++;
        if (p < in.n && (in.p[p] == '+' || in.p[p] == '-')) p++;
        if (p >= in.n || !cj_isdigit(in.p[p])) return 0;
        while (p < in.n && cj_isdigit(in.p[p])) p++;
    }
    return p;
}
CK_HD u32 cj_kw(const CIn& in, u32 pos, const char* kw, u32 L) {
    if (pos + L > in.n) return 0;
    for (u32 i = 0; i < L; i++) if (in.p[pos + i] != (u8)kw[i]) return 0;
    return pos + L;
}
// scalar token (not a container) -> end, 0 = bad
CK_HD u32 cj_scalar_end(const CIn& in, u32 pos) {
    u8 c = in.p[pos];
    if (c == '"') return cj_string_end(in, pos);
    if (c == 't') return cj_kw(in, pos, "true", 4);
    if (c == 'f') return cj_kw(in, pos, "false", 5);
    if (c == 'n') return cj_kw(in, pos, "null", 4);
    if (c == 'N') return cj_kw(in, pos, "NaN", 3);
    if (c == 'I') return cj_kw(in, pos, "Infinity", 8);
    if (c == '-' && pos + 1 < in.n && in.p[pos + 1] == 'I') return cj_kw(in, pos, "-Infinity", 9);
    return cj_number_end(in, pos);
}

CK_HD bool cj_validate(const CIn& in) {
    u32 kind[(CJ_MAX_DEPTH + 2 + 31) / 32];
    u32 depth = 0;                                  // number of containers enclosing the value at pos
    u32 pos = cj_skip_ws(in, 0);
    if (pos >= in.n) return false;
    for (;;) {
        // ---- a value starts at pos
        if (pos >= in.n) return false;
        if (depth > CJ_MAX_DEPTH) return false;
        u8 c = in.p[pos];
        bool opened = false;
        if (c == '{' || c == '[') {
            if (c == '{') kind[depth >> 5] |= 1u << (depth & 31); else kind[depth >> 5] &= ~(1u << (depth & 31));
            depth++;
            pos = cj_skip_ws(in, pos + 1);
            if (pos >= in.n) return false;
            if (in.p[pos] == (c == '{' ? '}' : ']')) { pos++; depth--; }
            else opened = true;
        } else {
            pos = cj_scalar_end(in, pos);
            if (!pos) return false;
        }
        // ---- after a value / after an opening bracket
        for (;;) {
            bool in_obj;
            if (opened) { in_obj = (kind[(depth - 1) >> 5] >> ((depth - 1) & 31)) & 1; opened = false; if (!in_obj) break; }
            else {
                pos = cj_skip_ws(in, pos);
                if (depth == 0) return pos == in.n;
                in_obj = (kind[(depth - 1) >> 5] >> ((depth - 1) & 31)) & 1;
                if (pos >= in.n) return false;
                u8 d = in.p[pos];
                if (d == (in_obj ? '}' : ']')) { pos++; depth--; continue; }
                if (d != ',') return false;
                pos = cj_skip_ws(in, pos + 1);
                if (!in_obj) break;
            }
            if (pos >= in.n || in.p[pos] != '"') return false;              // key must be a string
            pos = cj_string_end(in, pos);
            if (!pos) return false;
            pos = cj_skip_ws(in, pos);
            if (pos >= in.n || in.p[pos] != ':') return false;
            pos = cj_skip_ws(in, pos + 1);
            break;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Phase B helpers on syntactically valid text
// ---------------------------------------------------------------------------------------------------------
CK_HD u32 cj_skip_value(const CIn& in, u32 pos) {            // pos at the first byte of a value -> one past its end
    u32 depth = 0;
    for (;;) {
        u8 c = in.p[pos];
        if (c == '"') {
            pos++;
            while (in.p[pos] != '"') pos += (in.p[pos] == '\\') ? 2 : 1;
            pos++;
        } else if (c == '{' || c == '[') { depth++; pos++; }
        else if (c == '}' || c == ']') { depth--; pos++; }
        else if (c == ',' || c == ':' || cj_ws(c)) { pos++; continue; }
        else { while (pos < in.n && in.p[pos] != ',' && in.p[pos] != '}' && in.p[pos] != ']' && !cj_ws(in.p[pos])) pos++; }
        if (depth == 0) return pos;
    }
}

// decode the next code point of a string body at pos (not at the closing quote); advances pos
CK_HD u32 cj_next_cp(const CIn& in, u32& pos) {
    u8 c = in.p[pos];
    if (c == '\\') {
        u8 e = in.p[pos + 1];
        if (e != 'u') {
            pos += 2;
            return e == 'n' ? '\n' : e == 't' ? '\t' : e == 'r' ? '\r' : e == 'b' ? '\b' : e == 'f' ? '\f' : e;
        }
        u32 v = 0;
        for (u32 k = 2; k < 6; k++) v = v * 16 + (u32)cj_hex(in.p[pos + k]);
        pos += 6;
        if (v >= 0xD800 && v <= 0xDBFF) {
            u32 w = 0;
            for (u32 k = 2; k < 6; k++) w = w * 16 + (u32)cj_hex(in.p[pos + k]);
            pos += 6;
            v = 0x10000 + ((v - 0xD800) << 10) + (w - 0xDC00);
        }
        return v;
    }
    if (c < 0x80) { pos++; return c; }
    if (c < 0xE0) { u32 v = ((c & 0x1Fu) << 6) | (in.p[pos + 1] & 0x3Fu); pos += 2; return v; }
    if (c < 0xF0) { u32 v = ((c & 0x0Fu) << 12) | ((in.p[pos + 1] & 0x3Fu) << 6) | (in.p[pos + 2] & 0x3Fu); pos += 3; return v; }
    u32 v = ((c & 0x07u) << 18) | ((in.p[pos + 1] & 0x3Fu) << 12) | ((in.p[pos + 2] & 0x3Fu) << 6) | (in.p[pos + 3] & 0x3Fu);
    pos += 4;
    return v;
}

// string token at pos equals the ASCII literal (after unescaping)?
CK_HD bool cj_str_is(const CIn& in, u32 pos, const char* lit, u32 L) {
    pos++;
    for (u32 i = 0; i < L; i++) {
        if (in.p[pos] == '"') return false;
        if (cj_next_cp(in, pos) != (u32)(u8)lit[i]) return false;
    }
    return in.p[pos] == '"';
}
CK_HD u32 cj_str_hash(const CIn& in, u32 pos) {              // hash of the decoded string
    pos++;
    u32 h = 2166136261u;
    while (in.p[pos] != '"') h = (h ^ cj_next_cp(in, pos)) * 16777619u;
    return h;
}

// canonical spelling of a string token (SURVEY.md Appendix A scalar rules)
CK_HD void cj_emit_string(const CIn& in, u32 pos, COut& o) {
    o.put('"');
    pos++;
    while (in.p[pos] != '"') {
        u32 v = cj_next_cp(in, pos);
        if (v == '"') { o.put('\\'); o.put('"'); }
        else if (v == '\\') { o.put('\\'); o.put('\\'); }
        else if (v == '\n') { o.put('\\'); o.put('n'); }
        else if (v == '\t') { o.put('\\'); o.put('t'); }
        else if (v == '\r') { o.put('\\'); o.put('r'); }
        else if (v == '\b') { o.put('\\'); o.put('b'); }
        else if (v == '\f') { o.put('\\'); o.put('f'); }
        else if (v < 0x20) { const char* hx = "0123456789abcdef"; CPUTS(o, "\\u00"); o.put((u8)hx[v >> 4]); o.put((u8)hx[v & 15]); }
        else if (v < 0x80) o.put((u8)v);
        else if (v < 0x800) { o.put((u8)(0xC0 | (v >> 6))); o.put((u8)(0x80 | (v & 0x3F))); }
        else if (v < 0x10000) { o.put((u8)(0xE0 | (v >> 12))); o.put((u8)(0x80 | ((v >> 6) & 0x3F))); o.put((u8)(0x80 | (v & 0x3F))); }
        else { o.put((u8)(0xF0 | (v >> 18))); o.put((u8)(0x80 | ((v >> 12) & 0x3F))); o.put((u8)(0x80 | ((v >> 6) & 0x3F))); o.put((u8)(0x80 | (v & 0x3F))); }
    }
    o.put('"');
}

enum { CE_OK = 0, CE_SCHEMA = 1, CE_UNSUP = 2 };

// Number -> canonical text without binary floating point: for <= 15 significant digits the shortest
// round-trip spelling of the nearest double is the decimal itself (DBL_DIG), so only the layout changes:
// positional for 1e-5 <= |x| < 1e16 (always with a fractional part), else d[.ddd]e[+-]X (probe table in
// DESIGN.md).  Integers keep their digits ("-0" -> "0").  as_float: the field is float-typed (int -> N.0).
CK_HD int cj_emit_number(const CIn& in, u32 a, u32 b, COut& o, bool as_float) {
    bool neg = in.p[a] == '-';
    u32 p = neg ? a + 1 : a;
    bool is_float = false;
    for (u32 i = p; i < b; i++) if (in.p[i] == '.' || in.p[i] == 'e' || in.p[i] == 'E') is_float = true;
    if (!is_float && !as_float) {
        if (b - p > 4000) return CE_UNSUP;
        if (b - p == 1 && in.p[p] == '0') { o.put('0'); return CE_OK; }      // "-0" is the int 0
        o.copy(in, a, b);
        return CE_OK;
    }
    // significant digits D[0..nd) and decimal exponent such that value = 0.D * 10^e10
    u8 D[20]; u32 nd = 0; int e10 = 0; bool seen_nz = false, overflow_digits = false;
    u32 i = p;
    for (; i < b && cj_isdigit(in.p[i]); i++) {
        if (in.p[i] != '0') seen_nz = true;
        if (seen_nz) { if (nd < 20) D[nd++] = in.p[i]; else overflow_digits = true; e10++; }
    }
    if (i < b && in.p[i] == '.') {
        i++;
        for (; i < b && cj_isdigit(in.p[i]); i++) {
            if (in.p[i] != '0') seen_nz = true;
            if (seen_nz) { if (nd < 20) D[nd++] = in.p[i]; else overflow_digits = true; }
            else e10--;
        }
    }
    if (i < b && (in.p[i] == 'e' || in.p[i] == 'E')) {
        i++;
        bool eneg = false;
        if (in.p[i] == '+' || in.p[i] == '-') { eneg = in.p[i] == '-'; i++; }
        int ex = 0;
        for (; i < b; i++) { if (ex < 100000) ex = ex * 10 + (in.p[i] - '0'); }
        e10 += eneg ? -ex : ex;
    }
    while (nd > 0 && D[nd - 1] == '0') nd--;                                  // trailing zeros are not significant
    if (overflow_digits) return CE_UNSUP;
    if (nd == 0) { if (neg) o.put('-'); CPUTS(o, "0.0"); return CE_OK; }      // +-0.0
    // value = 0.D * 10^e10.  Beyond every double: infinity, which the reference dumps as null; below half of the smallest
    // subnormal: zero
    if (e10 >= 311) { CPUTS(o, "null"); return CE_OK; }
    if (e10 <= -326) { if (neg) o.put('-'); CPUTS(o, "0.0"); return CE_OK; }
    bool extreme = (e10 > 290 || e10 < -290);                                 // the <= 15 digit argument does not hold out here
    if (nd > 15 || extreme) {
        // kept if the digits are exactly what repr() of the nearest double prints, else replaced by that spelling
        // (csrc/ck_float.cuh: exact integer arithmetic, a search over the decimals next to the literal)
        if (nd > 19) return CE_UNSUP;
        u64 m = 0;
        for (u32 k = 0; k < nd; k++) m = m * 10 + (u64)(D[k] - '0');
        { u64 f_; int e_; int cls = ckf_nearest_double_ex(m, e10 - (int)nd, f_, e_);
          if (cls == CKF_INF) { CPUTS(o, "null"); return CE_OK; }
          if (cls == CKF_ZERO) { if (neg) o.put('-'); CPUTS(o, "0.0"); return CE_OK; }
          if (cls != CKF_FINITE) return CE_UNSUP; }
        if (!(nd <= 17 && ckf_is_repr(m, e10 - (int)nd))) {
            u64 ms; int ks;
            if (!ckf_shortest(m, e10 - (int)nd, ms, ks)) return CE_UNSUP;
            u32 n2 = 0; { u64 t = ms; while (t) { n2++; t /= 10; } }
            { u64 t = ms; for (u32 k = n2; k-- > 0;) { D[k] = (u8)('0' + t % 10); t /= 10; } }
            nd = n2; e10 = ks + (int)n2;
        }
    }
    if (neg) o.put('-');
    // value = D[0].D[1..] * 10^(e10-1)
    int x = e10 - 1;
    if (x >= -5 && x < 16) {
        if (e10 <= 0) { CPUTS(o, "0."); for (int k = 0; k < -e10; k++) o.put('0'); for (u32 k = 0; k < nd; k++) o.put(D[k]); }
        else {
            for (int k = 0; k < e10; k++) o.put(k < (int)nd ? D[k] : (u8)'0');
            o.put('.');
            if ((int)nd > e10) for (u32 k = (u32)e10; k < nd; k++) o.put(D[k]); else o.put('0');
        }
    } else {
        o.put(D[0]);
        if (nd > 1) { o.put('.'); for (u32 k = 1; k < nd; k++) o.put(D[k]); }
        o.put('e'); o.put(x < 0 ? '-' : '+');
        u32 ax = (u32)(x < 0 ? -x : x);
        if (ax >= 100) o.put((u8)('0' + ax / 100));
        if (ax >= 10) o.put((u8)('0' + (ax / 10) % 10));
        o.put((u8)('0' + ax % 10));
    }
    return CE_OK;
}

#define CJ_HS 256
struct CJ {                                   // one canonicalisation in flight
    CIn in; COut o;
    u32 depth;
    u32 hs[CJ_HS]; u32 hsp;                   // key hashes of the dict scopes being checked for duplicates (LIFO)
};

// decoded strings equal?
CK_HD bool cj_keys_equal(const CIn& in, u32 a, u32 b) {
    a++; b++;
    for (;;) {
        bool ea = in.p[a] == '"', eb = in.p[b] == '"';
        if (ea || eb) return ea && eb;
        if (cj_next_cp(in, a) != cj_next_cp(in, b)) return false;
    }
}

CK_HD int cj_emit_scalar(const CIn& in, u32 p, COut& o) {       // string / number / literal token at p
    u8 c = in.p[p];
    if (c == '"') { cj_emit_string(in, p, o); return CE_OK; }
    u32 e = cj_scalar_end(in, p);
    if (c == 't' || c == 'f' || c == 'n') { o.copy(in, p, e); return CE_OK; }
    if (c == 'N' || c == 'I' || (c == '-' && in.p[p + 1] == 'I')) { CPUTS(o, "null"); return CE_OK; }
    return cj_emit_number(in, p, e, o, false);
}

// generic value with duplicate keys somewhere inside: the reference holds Python dicts, so a repeated key keeps its FIRST
// position and takes the LAST value (dict assignment).  Recursive, bounded: <= CJ_DD_KEYS members per object with
// duplicates in play, <= CJ_DD_DEPTH levels; beyond that the record stays undecided.
#define CJ_DD_KEYS 32
#define CJ_DD_DEPTH 16
CK_HDR int cj_emit_any_dedup(CJ& cj, u32 pos, u32 depth) {
    const CIn& in = cj.in; COut& o = cj.o;
    if (depth > CJ_DD_DEPTH) return CE_UNSUP;
    u8 c = in.p[pos];
    if (c == '[') {
        o.put('[');
        u32 q = cj_skip_ws(in, pos + 1);
        bool first = true;
        while (in.p[q] != ']') {
            if (!first) o.put(',');
            first = false;
            int rc = cj_emit_any_dedup(cj, q, depth + 1); if (rc) return rc;
            q = cj_skip_ws(in, cj_skip_value(in, q));
            if (in.p[q] == ',') q = cj_skip_ws(in, q + 1);
        }
        o.put(']');
        return CE_OK;
    }
    if (c != '{') return cj_emit_scalar(in, pos, o);
    u32 kp[CJ_DD_KEYS], vp[CJ_DD_KEYS], n = 0;
    u32 q = cj_skip_ws(in, pos + 1);
    while (in.p[q] != '}') {
        u32 key = q;
        q = cj_skip_ws(in, cj_skip_value(in, q)); q++; q = cj_skip_ws(in, q);
        u32 j = 0;
        for (; j < n; j++) if (cj_keys_equal(in, kp[j], key)) break;
        if (j == n) { if (n >= CJ_DD_KEYS) return CE_UNSUP; kp[n] = key; n++; }
        vp[j] = q;
        q = cj_skip_ws(in, cj_skip_value(in, q));
        if (in.p[q] == ',') q = cj_skip_ws(in, q + 1);
    }
    o.put('{');
    for (u32 j = 0; j < n; j++) {
        if (j) o.put(',');
        cj_emit_string(in, kp[j], o);
        o.put(':');
        int rc = cj_emit_any_dedup(cj, vp[j], depth + 1); if (rc) return rc;
    }
    o.put('}');
    return CE_OK;
}

// generic value ("Any"): whitespace stripped, strings / numbers canonical, NaN / +-Infinity -> null.
CK_HDR int cj_emit_any(CJ& cj, u32 pos) {
    const CIn& in = cj.in; COut& o = cj.o;
    // iterative: the text is valid JSON, so emitting tokens in order minus whitespace reproduces the structure
    u32 end = cj_skip_value(in, pos);
    // duplicate-key check per object: O(k^2) on hashes of decoded keys, objects visited with a small stack
    {
        u32 p = pos;
        // walk every object in the subtree
        while (p < end) {
            u8 c = in.p[p];
            if (c == '"') { p = cj_skip_value(in, p); continue; }
            if (c == '{') {
                // collect this object's keys
                u32 q = cj_skip_ws(in, p + 1);
                u32 base = cj.hsp, nh = 0;
                while (in.p[q] != '}') {
                    u32 h = cj_str_hash(in, q);
                    for (u32 k = 0; k < nh; k++) if (cj.hs[base + k] == h) return cj_emit_any_dedup(cj, pos, 0);   // (or a hash collision: same result)
                    if (base + nh >= CJ_HS) return CE_UNSUP;
                    cj.hs[base + nh++] = h;
                    q = cj_skip_value(in, q);                 // key
                    q = cj_skip_ws(in, q); q++;                // ':'
                    q = cj_skip_ws(in, q);
                    q = cj_skip_value(in, q);                 // value
                    q = cj_skip_ws(in, q);
                    if (in.p[q] == ',') q = cj_skip_ws(in, q + 1);
                }
            }
            p++;
        }
    }
    u32 p = pos;
    while (p < end) {
        u8 c = in.p[p];
        if (cj_ws(c)) { p++; continue; }
        if (c == '"') { cj_emit_string(in, p, o); p = cj_skip_value(in, p); continue; }
        if (c == '{' || c == '}' || c == '[' || c == ']' || c == ',' || c == ':') { o.put(c); p++; continue; }
        u32 e = cj_scalar_end(in, p);
        if (c == 't' || c == 'f' || c == 'n') o.copy(in, p, e);
        else if (c == 'N' || c == 'I' || (c == '-' && in.p[p + 1] == 'I')) CPUTS(o, "null");
        else { int rc = cj_emit_number(in, p, e, o, false); if (rc) return rc; }
        p = e;
    }
    return CE_OK;
}

CK_HD bool cj_two(const u8* s, u32 k, u32& v) {
    if (!cj_isdigit(s[k]) || !cj_isdigit(s[k + 1])) return false;
    v = (u32)(s[k] - '0') * 10 + (u32)(s[k + 1] - '0');
    return true;
}
// datetime string -> canonical spelling; only layouts whose result is certain are accepted
// unix timestamp given as a plain integer (a JSON number, or a string holding one; an all-zero fraction is allowed): seconds,
// or milliseconds beyond the reference's watershed of 2e10 (speedate), always UTC.  Other numeric spellings (fractions,
// exponents: the reference goes through binary floating point there) stay undecided.
CK_HD int cj_emit_unix_datetime(const CIn& in, u32 a, u32 b, COut& o) {
    u32 p = a;
    bool neg = false;
    if (p < b && in.p[p] == '-') { neg = true; p++; }
    u32 d0 = p;
    long long v = 0;
    while (p < b && cj_isdigit(in.p[p])) { if (p - d0 >= 15) return CE_UNSUP; v = v * 10 + (in.p[p] - '0'); p++; }
    if (p == d0 || (in.p[d0] == '0' && p - d0 > 1)) return CE_UNSUP;
    if (p < b) {                                               // ".000": still the integer
        if (in.p[p] != '.' || p + 1 >= b) return CE_UNSUP;
        for (u32 q = p + 1; q < b; q++) if (in.p[q] != '0') return CE_UNSUP;
    }
    if (neg) v = -v;
    long long secs = v, micros = 0;
    if (v > 20000000000ll || v < -20000000000ll) {
        secs = v / 1000; long long ms = v % 1000;
        if (ms < 0) { ms += 1000; secs -= 1; }
        micros = ms * 1000;
    }
    long long days = secs / 86400, rem = secs % 86400;
    if (rem < 0) { rem += 86400; days -= 1; }
    // civil date from days since 1970-01-01 (proleptic Gregorian)
    long long z = days + 719468;
    long long era = (z >= 0 ? z : z - 146096) / 146097;
    long long doe = z - era * 146097;
    long long yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    long long y = yoe + era * 400;
    long long doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    long long mp = (5 * doy + 2) / 153;
    long long d = doy - (153 * mp + 2) / 5 + 1;
    long long mth = mp < 10 ? mp + 3 : mp - 9;
    if (mth <= 2) y += 1;
    if (y < 1 || y > 9999) return CE_UNSUP;
    u32 hh = (u32)(rem / 3600), mi = (u32)((rem % 3600) / 60), ss = (u32)(rem % 60);
    o.put('"');
    o.put((u8)('0' + (y / 1000) % 10)); o.put((u8)('0' + (y / 100) % 10)); o.put((u8)('0' + (y / 10) % 10)); o.put((u8)('0' + y % 10));
    o.put('-'); o.put((u8)('0' + mth / 10)); o.put((u8)('0' + mth % 10));
    o.put('-'); o.put((u8)('0' + d / 10)); o.put((u8)('0' + d % 10));
    o.put('T'); o.put((u8)('0' + hh / 10)); o.put((u8)('0' + hh % 10));
    o.put(':'); o.put((u8)('0' + mi / 10)); o.put((u8)('0' + mi % 10));
    o.put(':'); o.put((u8)('0' + ss / 10)); o.put((u8)('0' + ss % 10));
    if (micros) { o.put('.'); long long dv = 100000; for (int k = 0; k < 6; k++) { o.put((u8)('0' + (micros / dv) % 10)); dv /= 10; } }
    o.put('Z'); o.put('"');
    return CE_OK;
}

CK_HD int cj_emit_datetime(const CIn& in, u32 pos, COut& o) {
    if (in.p[pos] == '-' || cj_isdigit(in.p[pos])) return cj_emit_unix_datetime(in, pos, cj_scalar_end(in, pos), o);
    if (in.p[pos] != '"') return CE_UNSUP;
    u32 e = cj_skip_value(in, pos);
    u32 a = pos + 1, b = e - 1;
    for (u32 i = a; i < b; i++) if (in.p[i] == '\\' || in.p[i] >= 0x80) return CE_UNSUP;
    {   // a string holding a plain integer is taken as the number it spells
        u32 q = a; if (q < b && in.p[q] == '-') q++;
        u32 q0 = q; while (q < b && cj_isdigit(in.p[q])) q++;
        if (q == b && q > q0) return cj_emit_unix_datetime(in, a, b, o);
    }
    if (b - a < 19) return CE_UNSUP;
    const u8* s = in.p + a;
#define two(k, v) cj_two(s, (k), (v))
    u32 y1, y2, mo, d, h, mi, sec;
    if (!two(0, y1) || !two(2, y2) || s[4] != '-' || !two(5, mo) || s[7] != '-' || !two(8, d) || (s[10] != 'T' && s[10] != 't' && s[10] != ' ' && s[10] != '_') ||
        !two(11, h) || s[13] != ':' || !two(14, mi) || s[16] != ':' || !two(17, sec)) return CE_UNSUP;
    u32 y = y1 * 100 + y2;
    // the layout is RFC 3339's: a field out of its range is a parsing error in the reference (datetime_from_date_parsing), not
    // another spelling
    if (y < 1 || mo < 1 || mo > 12 || d < 1 || h > 23 || mi > 59 || sec > 59) return CE_SCHEMA;
    u32 dim = (mo == 2) ? (((y % 4 == 0 && y % 100 != 0) || y % 400 == 0) ? 29 : 28) : ((mo == 4 || mo == 6 || mo == 9 || mo == 11) ? 30 : 31);
    if (d > dim) return CE_SCHEMA;
    u32 k = 19, len = b - a;
    u8 frac[6] = {'0', '0', '0', '0', '0', '0'}; bool has_frac = false;
    if (k < len && (s[k] == '.' || s[k] == ',')) {
        k++;
        u32 nf = 0;
        while (k < len && cj_isdigit(s[k])) { if (nf < 6) frac[nf] = s[k]; nf++; k++; }
        if (nf == 0) return CE_SCHEMA;                        // "12:00:00.Z": a parsing error in the reference
        if (nf > 9) return CE_UNSUP;
        for (u32 j = 0; j < 6; j++) if (frac[j] != '0') has_frac = true;
    }
    // zone
    u8 zone = 0; u32 oh = 0, om = 0; u8 sign = '+';
    if (k == len) zone = 0;
    else if ((s[k] == 'Z' || s[k] == 'z') && k + 1 == len) zone = 1;
    else if ((s[k] == '+' || s[k] == '-') && ((k + 6 == len && s[k + 3] == ':') || k + 5 == len)) {      // +HH:MM or +HHMM
        sign = s[k];
        if (!two(k + 1, oh) || !two(k + (k + 6 == len ? 4 : 3), om)) return CE_UNSUP;
        if (oh > 23 || om > 59) return CE_SCHEMA;             // offset out of range: a parsing error in the reference
        zone = (oh == 0 && om == 0) ? 1 : 2;
    } else return CE_UNSUP;
#undef two
    o.put('"');
    for (u32 j = 0; j < 10; j++) o.put(s[j]);
    o.put('T');
    for (u32 j = 11; j < 19; j++) o.put(s[j]);
    if (has_frac) { o.put('.'); for (u32 j = 0; j < 6; j++) o.put(frac[j]); }
    if (zone == 1) o.put('Z');
    else if (zone == 2) { o.put(sign); o.put((u8)('0' + oh / 10)); o.put((u8)('0' + oh % 10)); o.put(':'); o.put((u8)('0' + om / 10)); o.put((u8)('0' + om % 10)); }
    o.put('"');
    return CE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Schema tables (canonical field order = declaration order of the reference models)
// ---------------------------------------------------------------------------------------------------------
enum {  // field types
    T_STR, T_STRN, T_ANY, T_OBJ, T_OBJN, T_INT, T_BOOL, T_BOOLN, T_FLTN, T_DT, T_DTN, T_CONST, T_ENUM, T_ENUMN,
    T_MODEL, T_MODELN, T_LIST_MODEL, T_LIST_STR,
    T_DICT_TCP, T_DICT_TRV, T_MSGN, T_LIST_MSG, T_LIST_CPART, T_ARGS, T_LIST_SCHEMA_N, T_LIST_REQPART, T_LIST_RESPPART,
    T_DICT_INT, T_STR_OR_LISTSTR, T_STR_ONLY_OR_UNSUP, T_ARGSUNION, T_STRN_OR_UNSUP
};
enum { D_REQ, D_NULL, D_LIST, D_DICT, D_FALSE, D_CONST, D_FACTORY, D_ZERO, D_SCHEMA_DEFAULT, D_USAGE, D_FUNCTION };
enum {  // models
    M_ENVELOPE, M_CONTEXT, M_STATE, M_DEPS, M_WF, M_STACK, M_FRAME, M_OVR, M_TOOLSCHEMA, M_TOOLDEF, M_REQUEST, M_RESPONSE,
    M_USAGE, M_SYSPROMPT, M_USERPROMPT, M_TOOLRETPART, M_RETRYPART, M_TEXTPART, M_THINKING, M_TOOLCALL, M_BUILTINCALL,
    M_BUILTINRET, M_TOOLRETURN, M_MODELRETRY, M_CTEXT, M_CFILE, M_CDATA, M_CTOOL, M_COUNT
};
enum { E_TOOLKIND, E_FINISH };
enum { K_REQUEST, K_RESPONSE, K_SYS, K_USER, K_TOOLRET, K_RETRY, K_TEXT, K_THINKING, K_TOOLCALL, K_BCALL, K_BRET, K_TR, K_MR, K_CTEXT,
       K_CFILE, K_CDATA, K_CTOOL };

struct CField { char name[24]; u8 nlen; u8 type; u8 arg; u8 dflt; char alt[16]; u8 altlen; u8 no_primary; };
struct CModel { u8 first, count; };

#if defined(__CUDACC__)
#define CJ_TABLE __device__ const
#else
#define CJ_TABLE static const
#endif
#define F(n, t, a, d) {n, (u8)(sizeof(n) - 1), t, a, d, "", 0, 0}
#define FA(n, t, a, d, alt) {n, (u8)(sizeof(n) - 1), t, a, d, alt, (u8)(sizeof(alt) - 1), 0}
#define FX(n, t, a, d, alt) {n, (u8)(sizeof(n) - 1), t, a, d, alt, (u8)(sizeof(alt) - 1), 1}

CJ_TABLE CField cj_fields[] = {
    /* M_ENVELOPE  0 */ F("context", T_MODEL, M_CONTEXT, D_REQ), F("internal_workflow_state", T_MODEL, M_WF, D_REQ),
    /* M_CONTEXT   2 */ F("state", T_MODEL, M_STATE, D_REQ), F("deps", T_MODEL, M_DEPS, D_REQ),
    /* M_STATE     4 */ F("tool_calls", T_DICT_TCP, 0, D_DICT), F("tool_results", T_DICT_TRV, 0, D_DICT),
                        F("uncommitted_message", T_MSGN, 0, D_NULL), F("message_history", T_LIST_MSG, 0, D_LIST),
                        F("final_output_parts", T_LIST_CPART, 0, D_LIST), F("temp_instructions", T_STRN, 0, D_NULL),
                        F("metadata", T_ANY, 0, D_NULL), F("overrides", T_MODELN, M_OVR, D_NULL),
    /* M_DEPS     12 */ F("correlation_id", T_STR, 0, D_REQ), F("provided_deps", T_OBJ, 0, D_REQ),
    /* M_WF       14 */ F("call_stack", T_MODEL, M_STACK, D_REQ), F("metadata", T_ANY, 0, D_NULL),
    /* M_STACK    16 */ F("_internal_list", T_LIST_MODEL, M_FRAME, D_LIST),
    /* M_FRAME    17 */ F("target_topic", T_STR, 0, D_REQ), F("callback_topic", T_STR, 0, D_REQ), F("input_args", T_ARGS, 0, D_NULL),
                        F("frame_id", T_STR, 0, D_FACTORY), F("overrides", T_MODELN, M_OVR, D_NULL),
    /* M_OVR      22 */ F("override_agent_tools", T_LIST_SCHEMA_N, 0, D_REQ),
    /* M_TOOLSCHEMA 23 */ F("node_id", T_STR, 0, D_REQ), F("subscribe_topics", T_LIST_STR, 0, D_REQ), F("publish_topic", T_STRN, 0, D_REQ),
                        F("tool_schema", T_MODEL, M_TOOLDEF, D_REQ),
    /* M_TOOLDEF  27 */ F("name", T_STR, 0, D_REQ), F("parameters_json_schema", T_OBJ, 0, D_SCHEMA_DEFAULT), F("description", T_STRN, 0, D_NULL),
                        F("outer_typed_dict_key", T_STRN, 0, D_NULL), F("strict", T_BOOLN, 0, D_NULL), F("sequential", T_BOOL, 0, D_FALSE),
                        F("kind", T_ENUM, E_TOOLKIND, D_FUNCTION), F("metadata", T_OBJN, 0, D_NULL), F("timeout", T_FLTN, 0, D_NULL),
    /* M_REQUEST  36 */ F("parts", T_LIST_REQPART, 0, D_REQ), F("timestamp", T_DTN, 0, D_NULL), F("instructions", T_STRN, 0, D_NULL),
                        F("kind", T_CONST, K_REQUEST, D_CONST), F("run_id", T_STRN, 0, D_NULL), F("metadata", T_OBJN, 0, D_NULL),
    /* M_RESPONSE 42 */ F("parts", T_LIST_RESPPART, 0, D_REQ), F("usage", T_MODEL, M_USAGE, D_USAGE), F("model_name", T_STRN, 0, D_NULL),
                        F("name", T_STRN, 0, D_NULL), F("timestamp", T_DT, 0, D_FACTORY), F("kind", T_CONST, K_RESPONSE, D_CONST),
                        F("provider_name", T_STRN, 0, D_NULL), F("provider_url", T_STRN, 0, D_NULL),
                        FA("provider_details", T_OBJN, 0, D_NULL, "vendor_details"), FA("provider_response_id", T_STRN, 0, D_NULL, "vendor_id"),
                        F("finish_reason", T_ENUMN, E_FINISH, D_NULL), F("run_id", T_STRN, 0, D_NULL), F("metadata", T_OBJN, 0, D_NULL),
    /* M_USAGE    55 */ F("input_tokens", T_INT, 0, D_ZERO), F("cache_write_tokens", T_INT, 0, D_ZERO), F("cache_read_tokens", T_INT, 0, D_ZERO),
                        F("output_tokens", T_INT, 0, D_ZERO), F("input_audio_tokens", T_INT, 0, D_ZERO), F("cache_audio_read_tokens", T_INT, 0, D_ZERO),
                        F("output_audio_tokens", T_INT, 0, D_ZERO), F("details", T_DICT_INT, 0, D_DICT),
    /* M_SYSPROMPT 63 */ F("content", T_STR, 0, D_REQ), F("timestamp", T_DT, 0, D_FACTORY), F("dynamic_ref", T_STRN, 0, D_NULL),
                        F("name", T_STRN, 0, D_NULL), F("part_kind", T_CONST, K_SYS, D_CONST),
    /* M_USERPROMPT 68 */ F("content", T_STR_OR_LISTSTR, 0, D_REQ), F("timestamp", T_DT, 0, D_FACTORY), F("name", T_STRN, 0, D_NULL),
                        F("part_kind", T_CONST, K_USER, D_CONST),
    /* M_TOOLRETPART 72 */ F("tool_name", T_STR, 0, D_REQ), F("content", T_ANY, 0, D_REQ), F("tool_call_id", T_STR, 0, D_FACTORY),
                        F("metadata", T_ANY, 0, D_NULL), F("timestamp", T_DT, 0, D_FACTORY), F("part_kind", T_CONST, K_TOOLRET, D_CONST),
    /* M_RETRYPART 78 */ F("content", T_STR_ONLY_OR_UNSUP, 0, D_REQ), F("tool_name", T_STRN, 0, D_NULL), F("tool_call_id", T_STR, 0, D_FACTORY),
                        F("timestamp", T_DT, 0, D_FACTORY), F("part_kind", T_CONST, K_RETRY, D_CONST),
    /* M_TEXTPART 83 */ F("content", T_STR, 0, D_REQ), F("id", T_STRN, 0, D_NULL), F("provider_name", T_STRN, 0, D_NULL),
                        F("provider_details", T_OBJN, 0, D_NULL), F("part_kind", T_CONST, K_TEXT, D_CONST),
    /* M_THINKING 88 */ F("content", T_STR, 0, D_REQ), F("id", T_STRN, 0, D_NULL), F("signature", T_STRN, 0, D_NULL), F("provider_name", T_STRN, 0, D_NULL),
                        F("provider_details", T_OBJN, 0, D_NULL), F("part_kind", T_CONST, K_THINKING, D_CONST),
    /* M_TOOLCALL 94 */ F("tool_name", T_STR, 0, D_REQ), F("args", T_ARGSUNION, 0, D_NULL), F("tool_call_id", T_STR, 0, D_FACTORY), F("id", T_STRN, 0, D_NULL),
                        F("provider_name", T_STRN, 0, D_NULL), F("provider_details", T_OBJN, 0, D_NULL), F("part_kind", T_CONST, K_TOOLCALL, D_CONST),
    /* M_BUILTINCALL 101 */ F("tool_name", T_STR, 0, D_REQ), F("args", T_ARGSUNION, 0, D_NULL), F("tool_call_id", T_STR, 0, D_FACTORY), F("id", T_STRN, 0, D_NULL),
                        F("provider_name", T_STRN, 0, D_NULL), F("provider_details", T_OBJN, 0, D_NULL), F("part_kind", T_CONST, K_BCALL, D_CONST),
    /* M_BUILTINRET 108 */ F("tool_name", T_STR, 0, D_REQ), F("content", T_ANY, 0, D_REQ), F("tool_call_id", T_STR, 0, D_FACTORY), F("metadata", T_ANY, 0, D_NULL),
                        F("timestamp", T_DT, 0, D_FACTORY), F("provider_name", T_STRN, 0, D_NULL), F("provider_details", T_OBJN, 0, D_NULL),
                        F("part_kind", T_CONST, K_BRET, D_CONST),
    /* M_TOOLRETURN 116 */ F("return_value", T_ANY, 0, D_REQ), F("content", T_STRN_OR_UNSUP, 0, D_NULL), F("metadata", T_ANY, 0, D_NULL),
                        F("kind", T_CONST, K_TR, D_CONST),
    /* M_MODELRETRY 120 */ F("message", T_STR, 0, D_REQ), F("kind", T_CONST, K_MR, D_REQ),
    /* M_CTEXT   122 */ F("kind", T_CONST, K_CTEXT, D_CONST), F("text", T_STR, 0, D_REQ), F("metadata", T_OBJN, 0, D_NULL),
    /* M_CFILE   125 */ F("kind", T_CONST, K_CFILE, D_CONST), F("media_type", T_STR, 0, D_REQ), F("uri", T_STRN, 0, D_NULL), F("data", T_STRN, 0, D_NULL),
                        F("metadata", T_OBJN, 0, D_NULL),
    /* M_CDATA   130 */ F("kind", T_CONST, K_CDATA, D_CONST), F("data", T_ANY, 0, D_REQ), FX("schema_", T_OBJN, 0, D_NULL, "schema"),
                        F("metadata", T_OBJN, 0, D_NULL),
    /* M_CTOOL   134 */ F("kind", T_CONST, K_CTOOL, D_CONST), F("tool_call_id", T_STR, 0, D_REQ), F("kwargs", T_OBJ, 0, D_REQ), F("tool_name", T_STR, 0, D_REQ),
                        F("metadata", T_OBJN, 0, D_NULL),
};
CJ_TABLE CModel cj_models[M_COUNT] = {
    {0, 2}, {2, 2}, {4, 8}, {12, 2}, {14, 2}, {16, 1}, {17, 5}, {22, 1}, {23, 4}, {27, 9}, {36, 6}, {42, 13}, {55, 8}, {63, 5}, {68, 4},
    {72, 6}, {78, 5}, {83, 5}, {88, 6}, {94, 7}, {101, 7}, {108, 8}, {116, 4}, {120, 2}, {122, 3}, {125, 5}, {130, 4}, {134, 5},
};
struct CConst { char s[24]; u8 len; };
CJ_TABLE CConst cj_consts[] = {
    {"request", 7}, {"response", 8}, {"system-prompt", 13}, {"user-prompt", 11}, {"tool-return", 11}, {"retry-prompt", 12}, {"text", 4},
    {"thinking", 8}, {"tool-call", 9}, {"builtin-tool-call", 17}, {"builtin-tool-return", 19}, {"tool-return", 11}, {"model-retry", 11},
    {"text", 4}, {"file", 4}, {"data", 4}, {"tool", 4},
};
CJ_TABLE CConst cj_enum_toolkind[] = {{"function", 8}, {"output", 6}, {"external", 8}, {"unapproved", 10}};
CJ_TABLE CConst cj_enum_finish[] = {{"stop", 4}, {"length", 6}, {"content_filter", 14}, {"tool_call", 9}, {"error", 5}};

#define CJ_MAXF 13

CK_HDR int cj_emit_model(CJ& c, u32 pos, u32 model);

CK_HD void cj_emit_const(COut& o, u32 k) { o.put('"'); o.puts(cj_consts[k].s, cj_consts[k].len); o.put('"'); }

// find member `name` (last occurrence wins) among the members of the object at pos ('{'); 0 = absent
CK_HD u32 cj_find_member(const CIn& in, u32 pos, const char* name, u32 nlen) {
    u32 q = cj_skip_ws(in, pos + 1), found = 0;
    while (in.p[q] != '}') {
        bool hit = cj_str_is(in, q, name, nlen);
        q = cj_skip_value(in, q); q = cj_skip_ws(in, q); q++; q = cj_skip_ws(in, q);
        if (hit) found = q;
        q = cj_skip_value(in, q); q = cj_skip_ws(in, q);
        if (in.p[q] == ',') q = cj_skip_ws(in, q + 1);
    }
    return found;
}

// tagged union element: object whose `tag` member selects the model
CK_HDR int cj_emit_tagged(CJ& c, u32 pos, const char* tag, u32 taglen, const u8* kinds, const u8* models, u32 nk, bool* unsup_file = nullptr) {
    const CIn& in = c.in;
    if (in.p[pos] != '{') return CE_SCHEMA;                                    // dict_type / model_type
    u32 v = cj_find_member(in, pos, tag, taglen);
    if (!v) return CE_SCHEMA;                                                  // union_tag_not_found
    if (in.p[v] != '"') return CE_SCHEMA;                                      // union_tag_invalid
    for (u32 k = 0; k < nk; k++)
        if (cj_str_is(in, v, cj_consts[kinds[k]].s, cj_consts[kinds[k]].len)) return cj_emit_model(c, pos, models[k]);
    if (unsup_file && cj_str_is(in, v, "file", 4)) return CE_UNSUP;           // response FilePart: out of scope
    return CE_SCHEMA;                                                          // union_tag_invalid
}

CK_HDR int cj_emit_list(CJ& c, u32 pos, u32 elem_type, u32 arg);

CK_HDR int cj_emit_value(CJ& c, u32 pos, u32 type, u32 arg) {
    const CIn& in = c.in; COut& o = c.o;
    u8 ch = in.p[pos];
    bool is_null = (ch == 'n');
    switch (type) {
    case T_STRN: if (is_null) { CPUTS(o, "null"); return CE_OK; } /* fall through */
    case T_STR: if (ch != '"') return CE_SCHEMA; cj_emit_string(in, pos, o); return CE_OK;
    case T_ANY: return cj_emit_any(c, pos);
    case T_OBJN: if (is_null) { CPUTS(o, "null"); return CE_OK; } /* fall through */
    case T_OBJ: if (ch != '{') return CE_SCHEMA; return cj_emit_any(c, pos);
    case T_INT: {
        if (ch == '"') {
            // lax mode: a string holding a plain canonical integer, "-?(0|[1-9][0-9]*)", is that integer; every other string
            // (signs, underscores, whitespace, "5.0" ...: pydantic has its own rules for them) stays undecided
            u32 a = pos + 1, b = a;
            if (in.p[b] == '-') b++;
            u32 d0 = b;
            while (cj_isdigit(in.p[b])) b++;
            if (in.p[b] != '"' || b == d0 || b - d0 > 18 || (in.p[d0] == '0' && b - d0 > 1) || (in.p[a] == '-' && in.p[d0] == '0')) {
                // a string with a character no integer spelling contains ("x", "abc"): int_parsing in the reference; anything made
                // of digits, signs, separators, exponent letters or blanks only stays undecided
                for (u32 q = a; in.p[q] != '"'; q++) {
                    u8 cc = in.p[q];
                    if (cc == '\\' || cc >= 0x80 || cc < 0x20) return CE_UNSUP;
                    bool maybe = cj_isdigit(cc) || cc == '+' || cc == '-' || cc == '_' || cc == '.' || cc == ' ' || cc == 'e' || cc == 'E';
                    if (!maybe) return CE_SCHEMA;
                }
                return CE_UNSUP;
            }
            o.copy(in, a, b);
            return CE_OK;
        }
        if (ch == 't' || ch == 'f') return CE_UNSUP;                           // bool -> int: undecided
        if (!(ch == '-' || cj_isdigit(ch))) return CE_SCHEMA;
        u32 e = cj_scalar_end(in, pos);
        for (u32 i = pos; i < e; i++) if (in.p[i] == '.' || in.p[i] == 'e' || in.p[i] == 'E') {
            // lax mode: a float literal with an all-zero fraction (7.0, 12.000) is the integer; other float spellings undecided
            u32 j = i;
            if (in.p[j] != '.' || j + 1 >= e || i - pos > 15) return CE_UNSUP;
            for (j = i + 1; j < e; j++) if (in.p[j] != '0') {
                // a non-zero fraction digit: with <= 15 significant digits in all (and no exponent) the double the reference
                // parses keeps its fractional part (distinct decimals of <= 15 digits are distinct doubles) -> int_from_float
                bool plain = true; u32 sig = 0; bool nz = false;
                for (u32 q = pos; q < e; q++) {
                    u8 c2 = in.p[q];
                    if (c2 == 'e' || c2 == 'E') plain = false;
                    if (cj_isdigit(c2)) { if (c2 != '0') nz = true; if (nz) sig++; }
                }
                return (plain && sig <= 15) ? CE_SCHEMA : CE_UNSUP;
            }
            if (in.p[pos] == '-' && i == pos + 2 && in.p[pos + 1] == '0') { o.put('0'); return CE_OK; }      // -0.0 -> 0
            return cj_emit_number(in, pos, i, o, false);
        }
        return cj_emit_number(in, pos, e, o, false);
    }
    case T_BOOLN: if (is_null) { CPUTS(o, "null"); return CE_OK; } /* fall through */
    case T_BOOL:
        if (ch == 't') { CPUTS(o, "true"); return CE_OK; }
        if (ch == 'f') { CPUTS(o, "false"); return CE_OK; }
        // lax mode, the unambiguous spellings only: 1 / 0 and "true" / "false"; the rest of pydantic's table stays undecided
        if ((ch == '1' || ch == '0') && cj_scalar_end(in, pos) == pos + 1) { if (ch == '1') CPUTS(o, "true"); else CPUTS(o, "false"); return CE_OK; }
        if (ch == '"' && cj_str_is(in, pos, "true", 4)) { CPUTS(o, "true"); return CE_OK; }
        if (ch == '"' && cj_str_is(in, pos, "false", 5)) { CPUTS(o, "false"); return CE_OK; }
        return (ch == '{' || ch == '[' || is_null) ? CE_SCHEMA : CE_UNSUP;
    case T_FLTN:
        if (is_null) { CPUTS(o, "null"); return CE_OK; }
        if (!(ch == '-' || cj_isdigit(ch))) return (ch == '{' || ch == '[') ? CE_SCHEMA : CE_UNSUP;
        return cj_emit_number(in, pos, cj_scalar_end(in, pos), o, true);
    case T_DTN: if (is_null) { CPUTS(o, "null"); return CE_OK; } /* fall through */
    case T_DT: if (ch == '{' || ch == '[' || is_null || ch == 't' || ch == 'f') return CE_SCHEMA; return cj_emit_datetime(in, pos, o);
    case T_CONST:
        if (ch != '"') return CE_SCHEMA;
        if (!cj_str_is(in, pos, cj_consts[arg].s, cj_consts[arg].len)) return CE_SCHEMA;
        cj_emit_const(o, arg);
        return CE_OK;
    case T_ENUMN: if (is_null) { CPUTS(o, "null"); return CE_OK; } /* fall through */
    case T_ENUM: {
        if (ch != '"') return CE_SCHEMA;
        const CConst* e = arg == E_TOOLKIND ? cj_enum_toolkind : cj_enum_finish;
        u32 ne = arg == E_TOOLKIND ? 4 : 5;
        for (u32 k = 0; k < ne; k++) if (cj_str_is(in, pos, e[k].s, e[k].len)) { o.put('"'); o.puts(e[k].s, e[k].len); o.put('"'); return CE_OK; }
        return CE_SCHEMA;
    }
    case T_MODELN: if (is_null) { CPUTS(o, "null"); return CE_OK; } /* fall through */
    case T_MODEL: if (ch != '{') return CE_SCHEMA; return cj_emit_model(c, pos, arg);
    case T_LIST_MODEL: case T_LIST_STR: case T_LIST_MSG: case T_LIST_CPART: case T_LIST_REQPART: case T_LIST_RESPPART:
        return cj_emit_list(c, pos, type, arg);
    case T_LIST_SCHEMA_N: if (is_null) { CPUTS(o, "null"); return CE_OK; } return cj_emit_list(c, pos, T_LIST_MODEL, M_TOOLSCHEMA);
    case T_MSGN:
        if (is_null) { CPUTS(o, "null"); return CE_OK; }
        { const u8 ks[2] = {K_REQUEST, K_RESPONSE}, ms[2] = {M_REQUEST, M_RESPONSE}; return cj_emit_tagged(c, pos, "kind", 4, ks, ms, 2); }
    case T_ARGS: {                                  // Sequence[Any] | None
        if (is_null) { CPUTS(o, "null"); return CE_OK; }
        if (ch != '[') return CE_SCHEMA;
        return cj_emit_any(c, pos);
    }
    case T_ARGSUNION:                               // str | dict[str, Any] | None
        if (is_null) { CPUTS(o, "null"); return CE_OK; }
        if (ch == '"') { cj_emit_string(in, pos, o); return CE_OK; }
        if (ch == '{') return cj_emit_any(c, pos);
        return CE_SCHEMA;
    case T_STR_OR_LISTSTR:
        if (ch == '"') { cj_emit_string(in, pos, o); return CE_OK; }
        if (ch == '[') return cj_emit_list(c, pos, T_LIST_STR, 1);      // arg 1: non-string elements are UNSUPPORTED (multi-modal)
        return CE_SCHEMA;
    case T_STR_ONLY_OR_UNSUP:
        if (ch == '"') { cj_emit_string(in, pos, o); return CE_OK; }
        return ch == '[' ? CE_UNSUP : CE_SCHEMA;                         // list[ErrorDetails]
    case T_STRN_OR_UNSUP:
        if (is_null) { CPUTS(o, "null"); return CE_OK; }
        if (ch == '"') { cj_emit_string(in, pos, o); return CE_OK; }
        return ch == '[' ? CE_UNSUP : CE_SCHEMA;
    case T_DICT_INT: case T_DICT_TCP: case T_DICT_TRV: {
        if (ch != '{') return CE_SCHEMA;
        o.put('{');
        u32 q = cj_skip_ws(in, pos + 1);
        u32 base = c.hsp, nh = 0;
        bool first = true;
        while (in.p[q] != '}') {
            u32 h = cj_str_hash(in, q);
            bool seen = false;
            u32 cont = 0;
            for (u32 k = 0; k < nh; k++) if (c.hs[base + k] == h) seen = true;
            if (seen) {
                // a key that occurred before: it keeps its first position and took its last value there (below)
                q = cj_skip_value(in, q); q = cj_skip_ws(in, q); q++; q = cj_skip_ws(in, q);
                q = cj_skip_value(in, q); q = cj_skip_ws(in, q);
                if (in.p[q] == ',') q = cj_skip_ws(in, q + 1);
                continue;
            }
            if (base + nh >= CJ_HS) return CE_UNSUP;
            c.hs[base + nh++] = h;
            c.hsp = base + nh;                                                    // nested scopes stack above this one
            if (!first) o.put(',');
            first = false;
            cj_emit_string(in, q, o);
            o.put(':');
            u32 key0 = q;
            q = cj_skip_value(in, q); q = cj_skip_ws(in, q); q++; q = cj_skip_ws(in, q);
            {   // dict assignment: the value of the LAST member with this key (a 32-bit hash hit is verified on the decoded bytes).
                // The reference validates every member before the assignment: a shadowed value that is not a valid instance
                // still fails the record (found by the mutation fuzz), so the shadowed ones are emitted into nothing first.
                u32 scan = cj_skip_ws(in, cj_skip_value(in, q)), lastv = 0, prevv = q;
                while (in.p[scan] == ',') {
                    scan = cj_skip_ws(in, scan + 1);
                    u32 k2 = scan;
                    scan = cj_skip_ws(in, cj_skip_value(in, scan)); scan++; scan = cj_skip_ws(in, scan);
                    if (cj_str_hash(in, k2) == h) {
                        if (!cj_keys_equal(in, key0, k2)) return CE_UNSUP;
                        if (type != T_DICT_TRV) {                                 // (tool_results values fall back to Any: always valid)
                            u32 save = o.len, sd = c.depth, sh = c.hsp; bool sovf = o.ovf;
                            int r0 = type == T_DICT_INT ? cj_emit_value(c, prevv, T_INT, 0) : cj_emit_value(c, prevv, T_MODEL, M_TOOLCALL);
                            o.len = save; o.ovf = sovf; c.depth = sd; c.hsp = sh;
                            if (r0) return r0;
                        }
                        lastv = scan; prevv = scan;
                    }
                    scan = cj_skip_ws(in, cj_skip_value(in, scan));
                }
                cont = q;
                if (lastv) q = lastv;
            }
            int rc;
            if (type == T_DICT_INT) rc = cj_emit_value(c, q, T_INT, 0);
            else if (type == T_DICT_TCP) rc = cj_emit_value(c, q, T_MODEL, M_TOOLCALL);
            else {
                // ToolReturn | ModelRetry | RetryPromptPart by `kind` (else `part_kind`), falling back to Any when the
                // tag is absent / foreign or the tagged model does not validate (smart union, models/state.py:70)
                rc = -1;
                if (in.p[q] == '{') {
                    u32 tv = cj_find_member(in, q, "kind", 4);
                    if (!tv) tv = cj_find_member(in, q, "part_kind", 9);
                    int m = -1;
                    if (tv && in.p[tv] == '"') {
                        if (cj_str_is(in, tv, "tool-return", 11)) m = M_TOOLRETURN;
                        else if (cj_str_is(in, tv, "model-retry", 11)) m = M_MODELRETRY;
                        else if (cj_str_is(in, tv, "retry-prompt", 12)) m = M_RETRYPART;
                    }
                    if (m >= 0) {
                        u32 save = o.len, sd = c.depth, sh = c.hsp;
                        bool sovf = o.ovf;
                        rc = cj_emit_model(c, q, (u32)m);
                        // tagged but not a valid instance: the reference keeps it as plain data (`| Any` of the smart union).
                        // The second walk accepts such a value only from here (trusting reader): it cannot re-derive that
                        // the tagged model does not validate.
                        if (rc == CE_SCHEMA) { o.len = save; o.ovf = sovf; c.depth = sd; c.hsp = sh; rc = -1; }
                    }
                }
                if (rc < 0) rc = cj_emit_any(c, q);
            }
            if (rc) return rc;
            q = cj_skip_value(in, cont); q = cj_skip_ws(in, q);
            if (in.p[q] == ',') q = cj_skip_ws(in, q + 1);
        }
        o.put('}');
        c.hsp = base;
        return CE_OK;
    }
    }
    return CE_UNSUP;
}

CK_HDR int cj_emit_list(CJ& c, u32 pos, u32 type, u32 arg) {
    const CIn& in = c.in; COut& o = c.o;
    if (in.p[pos] != '[') return CE_SCHEMA;                                     // list_type
    o.put('[');
    u32 q = cj_skip_ws(in, pos + 1);
    bool first = true;
    while (in.p[q] != ']') {
        if (!first) o.put(',');
        first = false;
        int rc;
        switch (type) {
        case T_LIST_STR:
            if (in.p[q] != '"') return arg ? CE_UNSUP : CE_SCHEMA;
            cj_emit_string(in, q, o); rc = CE_OK; break;
        case T_LIST_MODEL: rc = (in.p[q] == '{') ? cj_emit_model(c, q, arg) : CE_SCHEMA; break;
        case T_LIST_MSG: { const u8 ks[2] = {K_REQUEST, K_RESPONSE}, ms[2] = {M_REQUEST, M_RESPONSE}; rc = cj_emit_tagged(c, q, "kind", 4, ks, ms, 2); break; }
        case T_LIST_CPART: { const u8 ks[4] = {K_CTEXT, K_CFILE, K_CDATA, K_CTOOL}, ms[4] = {M_CTEXT, M_CFILE, M_CDATA, M_CTOOL};
                             rc = cj_emit_tagged(c, q, "kind", 4, ks, ms, 4); break; }
        case T_LIST_REQPART: { const u8 ks[4] = {K_SYS, K_USER, K_TOOLRET, K_RETRY}, ms[4] = {M_SYSPROMPT, M_USERPROMPT, M_TOOLRETPART, M_RETRYPART};
                               rc = cj_emit_tagged(c, q, "part_kind", 9, ks, ms, 4); break; }
        default: { const u8 ks[5] = {K_TEXT, K_TOOLCALL, K_BCALL, K_BRET, K_THINKING}, ms[5] = {M_TEXTPART, M_TOOLCALL, M_BUILTINCALL, M_BUILTINRET, M_THINKING};
                   bool f = true; rc = cj_emit_tagged(c, q, "part_kind", 9, ks, ms, 5, &f); break; }
        }
        if (rc) return rc;
        q = cj_skip_value(in, q); q = cj_skip_ws(in, q);
        if (in.p[q] == ',') q = cj_skip_ws(in, q + 1);
    }
    o.put(']');
    return CE_OK;
}

CK_HDR int cj_emit_model(CJ& c, u32 pos, u32 model) {
    const CIn& in = c.in; COut& o = c.o;
    if (++c.depth > 24) return CE_UNSUP;
    const CModel& M = cj_models[model];
    u32 fpos[CJ_MAXF];
    for (u32 k = 0; k < M.count; k++) fpos[k] = 0;
    // one pass over the members: last occurrence of every known key wins, unknown keys are ignored
    u32 q = cj_skip_ws(in, pos + 1);
    while (in.p[q] != '}') {
        u32 key = q;
        q = cj_skip_value(in, q); q = cj_skip_ws(in, q); q++; q = cj_skip_ws(in, q);
        for (u32 k = 0; k < M.count; k++) {
            const CField& f = cj_fields[M.first + k];
            bool hit = (!f.no_primary && cj_str_is(in, key, f.name, f.nlen));
            if (hit) { fpos[k] = q; break; }
        }
        q = cj_skip_value(in, q); q = cj_skip_ws(in, q);
        if (in.p[q] == ',') q = cj_skip_ws(in, q + 1);
    }
    // validation aliases (AliasChoices: the primary name wins when both are present)
    for (u32 k = 0; k < M.count; k++) {
        const CField& f = cj_fields[M.first + k];
        if (f.altlen && !fpos[k]) fpos[k] = cj_find_member(in, pos, f.alt, f.altlen);
    }
    o.put('{');
    for (u32 k = 0; k < M.count; k++) {
        const CField& f = cj_fields[M.first + k];
        if (k) o.put(',');
        o.put('"'); o.puts(f.name, f.nlen); o.put('"'); o.put(':');
        if (fpos[k]) {
            // DataPart.schema_ given through its alias survives one dump but not the next (the emitted key
            // `schema_` is ignored on validation): such a record has no stable canonical form
            // (DataPart.schema_ arrives through its alias `schema`: the reference dumps the value under `schema_`; that text
            // is no fixed point of the codec — a second validation would ignore the key — so only the trusting second walk
            // accepts it, from here)
            int rc = cj_emit_value(c, fpos[k], f.type, f.arg); if (rc) return rc; continue;
        }
        switch (f.dflt) {
        case D_REQ: return CE_SCHEMA;                                           // missing
        case D_NULL: CPUTS(o, "null"); break;
        case D_LIST: CPUTS(o, "[]"); break;
        case D_DICT: CPUTS(o, "{}"); break;
        case D_FALSE: CPUTS(o, "false"); break;
        case D_ZERO: CPUTS(o, "0"); break;
        case D_CONST: cj_emit_const(o, f.arg); break;
        case D_FUNCTION: CPUTS(o, "\"function\""); break;
        case D_SCHEMA_DEFAULT: CPUTS(o, "{\"type\":\"object\",\"properties\":{}}"); break;
        case D_USAGE: CPUTS(o, "{\"input_tokens\":0,\"cache_write_tokens\":0,\"cache_read_tokens\":0,\"output_tokens\":0,\"input_audio_tokens\":0,"
                               "\"cache_audio_read_tokens\":0,\"output_audio_tokens\":0,\"details\":{}}"); break;
        default: return CE_UNSUP;                                               // D_FACTORY: now() / generated ids
        }
    }
    o.put('}');
    c.depth--;
    return CE_OK;
}

// entry point: -> CK_OK / CK_JSON_INVALID / CK_SCHEMA_INVALID / CK_UNSUPPORTED; out_len = canonical length
CK_HD_NOINLINE u32 ck_canonicalise(const u8* rec, u32 len, u8* out, u32 cap, u32& out_len) {
    CJ c; c.in.p = rec; c.in.n = len; c.o.p = out; c.o.cap = cap; c.o.len = 0; c.o.ovf = false; c.depth = 0; c.hsp = 0;
    out_len = 0;
    if (!cj_validate(c.in)) return CK_JSON_INVALID;
    u32 pos = cj_skip_ws(c.in, 0);
    if (c.in.p[pos] != '{') return CK_SCHEMA_INVALID;                           // model_type
    int rc = cj_emit_model(c, pos, M_ENVELOPE);
    if (rc == CE_SCHEMA) return CK_SCHEMA_INVALID;
    out_len = c.o.len;
    if (rc == CE_UNSUP || (cap != 0 && c.o.ovf)) return CK_UNSUPPORTED;      // cap == 0: counting run, nothing is stored
    return CK_OK;
}

#endif  // CK_CANON_CUH
