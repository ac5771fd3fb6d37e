// TEST INFRASTRUCTURE ONLY.  Compiles the device walker source (csrc/ck_walk.cuh, which is written
// __host__ __device__) with g++ so the recogniser can be fuzzed against pydantic on the CPU, where
// millions of cases run in seconds.  Nothing in the product links or loads this file; the shipped
// libcalfkit_b200.so exposes no host implementation of the walker.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../calfkit-sdk_b200/csrc/ck_walk.cuh"

extern "C" int ck_host_walk(const uint8_t* buf, uint32_t len, uint32_t* cols /* CK_NUM_COLS */) {
    // 256-aligned copy with 64 bytes of slack each side, filled with '"' so that out-of-record reads
    // would be noticed if they were ever used
    size_t cap = (size_t)len + 512;
    std::vector<uint8_t> raw(cap + 256);
    uint8_t* base = (uint8_t*)(((uintptr_t)raw.data() + 255) & ~(uintptr_t)255);
    memset(base, '"', cap);
    uint8_t* rec = base + 64 + (len % 7);     // vary the alignment of the record start
    memcpy(rec, buf, len);
    Rd r; r.init(rec, len);
    memset(cols, 0, sizeof(uint32_t) * CK_NUM_COLS);
    WalkOut o; o.base = cols; o.stride = 1;
    AnyCtx cx; memset(&cx, 0, sizeof cx);
    u32 stop = 0;
    bool ok = ck_walk_envelope(r, o, cx, stop);
    cols[CK_COL_STATUS] = ok ? CK_OK : CK_NOT_CANONICAL;
    cols[CK_COL_ERR] = stop;
    return ok ? 1 : 0;
}
extern "C" int ck_host_num_cols() { return CK_NUM_COLS; }
