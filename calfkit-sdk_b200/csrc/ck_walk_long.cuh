// Long records: one warp per record, intra-record parallelism (see the block comment in ck_walk.cuh above ck_long_index).
//
// Why: a schema walk is sequential per record, so one thread per record takes (record length) x (~100-200 cycles per byte)
// no matter how idle the machine is — a 20 KB fan-out record 2.8 ms, a 64 KB history 5 ms (round 1: config 3 walk at 0.6 %
// of the HBM roofline, config 5 at 1.8 %).  Long records are long because of LISTS (64 tool calls, 64 parts, dozens of
// messages): their elements are independent given the boundaries.
//   ck_lx_build          warp-parallel structural pre-scan, 512 bytes per iteration (16 per lane): byte-class bit masks by
//                        SIMD-in-word compares, escaped quotes from the backslash runs, string mask by prefix XOR (within
//                        the lane by shifts, across lanes by shuffles), nesting depth by prefix sums of the bracket bits;
//                        the positions of commas and closers at depths 4 and 6 go to shared memory in order
//   ck_walk_long_kernel  the walker in lockstep over the record (URd); at the long lists the lanes take one element each
// The thread-per-record kernel hands every record of CK_LONG_MIN bytes or more to this one through a device-side list.
#ifndef CK_WALK_LONG_CUH
#define CK_WALK_LONG_CUH

#ifndef CK_LONG_MIN
#define CK_LONG_MIN 16384u
#endif
#define CK_LONG_WARPS 4

__device__ __forceinline__ u32 ck_nib(u32 m) { return (((m & 0x01010101u) * 0x01020408u) >> 24) & 0xFu; }     // 0xFF/0x00 per byte -> 4 bits
__device__ __forceinline__ u32 ck_mask16(const uint4& w, u32 c) {
    u32 cc = c * 0x01010101u;
    return ck_nib(__vcmpeq4(w.x, cc)) | (ck_nib(__vcmpeq4(w.y, cc)) << 4) | (ck_nib(__vcmpeq4(w.z, cc)) << 8) | (ck_nib(__vcmpeq4(w.w, cc)) << 12);
}
__device__ __forceinline__ u32 ck_mask16_or20(const uint4& w, u32 c) {          // bytes equal to c once bit 5 is set: '{' / '[' and '}' / ']'
    const u32 b = 0x20202020u;
    u32 cc = c * 0x01010101u;
    return ck_nib(__vcmpeq4(w.x | b, cc)) | (ck_nib(__vcmpeq4(w.y | b, cc)) << 4) | (ck_nib(__vcmpeq4(w.z | b, cc)) << 8) | (ck_nib(__vcmpeq4(w.w | b, cc)) << 12);
}

// [from, to): the bytes to scan (the whole record, or one container that starts at a structural character outside any
// string); depth0: containers open before `from`
__device__ __forceinline__ void ck_lx_build(const u8* __restrict__ g, u32 n, ck_long_index* __restrict__ lx, u32 from = 0, u32 to = 0xffffffffu, int depth0 = 0) {
    u32 lane = threadIdx.x & 31;
    u32 m0 = (u32)((uintptr_t)g & 15);
    const uint4* stream = (const uint4*)(g - m0);                   // 16-byte aligned; the buffers are padded on both sides of a record
    if (to > n) to = n;
    u32 m = m0 + from, total = m0 + to;                             // valid stream positions: [m, total)
    u32 n_sep[2] = {0, 0}, n_close[2] = {0, 0};
    u32 prev_bs = 0, str_carry = 0; int depth_carry = depth0;
    bool overflow = false;
    for (u32 t0 = m & ~511u; t0 < total; t0 += 512) {
        u32 p0 = t0 + 16 * lane;                                    // stream position of this lane's first byte
        uint4 w = make_uint4(0, 0, 0, 0);
        if (p0 < total) w = __ldg(stream + (p0 >> 4));
        // valid bytes: [m, total)
        u32 V = 0xFFFFu;
        if (p0 < m) V &= (m - p0 >= 16) ? 0u : (0xFFFFu << (m - p0));
        if (p0 + 16 > total) V &= (p0 >= total) ? 0u : (0xFFFFu >> (p0 + 16 - total));
        u32 Q = ck_mask16(w, '"') & V, B = ck_mask16(w, '\\') & V, K = ck_mask16(w, ',') & V;
        u32 O = ck_mask16_or20(w, '{') & V, C = ck_mask16_or20(w, '}') & V;
        // escaped quotes: preceded by a backslash run of length 1 or 3 (longer runs: the proposal may be wrong, the walk decides)
        u32 up = __shfl_up_sync(0xffffffffu, B >> 12, 1);
        u32 prev4 = lane ? up : prev_bs;
        prev_bs = __shfl_sync(0xffffffffu, B >> 12, 31);
        u32 B20 = (B << 4) | prev4;
        u32 e1 = B20 << 1, r2 = e1 & (B20 << 2), r3 = r2 & (B20 << 3), r4 = r3 & (B20 << 4);
        Q &= ~(((e1 & ~r2) | (r3 & ~r4)) >> 4);
        // string mask: exclusive prefix parity of the quote bits
        u32 S = Q; S ^= S << 1; S ^= S << 2; S ^= S << 4; S ^= S << 8; S &= 0xFFFFu;
        u32 par = __popc(Q) & 1u, inc = par;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (u32)o) inc ^= y; }
        u32 carry = (inc ^ par) ^ str_carry;                        // parity before this lane's first byte
        str_carry ^= __shfl_sync(0xffffffffu, inc, 31);
        u32 E = ((S << 1) & 0xFFFFu) ^ (carry ? 0xFFFFu : 0u);      // bit b: byte b lies inside a string
        O &= ~E; C &= ~E; K &= ~E;
        // nesting depth before each byte
        int delta = (int)__popc(O) - (int)__popc(C), dinc = delta;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, dinc, o); if (lane >= (u32)o) dinc += y; }
        int dbase = depth_carry + dinc - delta;
        depth_carry += __shfl_sync(0xffffffffu, dinc, 31);
        // this lane's entries for the four lists
        u32 cnt[4] = {0, 0, 0, 0};
        u32 pend = K | C;
        for (u32 x = pend; x; x &= x - 1) {
            u32 b = __ffs(x) - 1, below = (1u << b) - 1u;
            int d = dbase + (int)__popc(O & below) - (int)__popc(C & below);
            u32 li = d == 4 ? 0u : (d == 6 ? 1u : 2u);
            if (li < 2) cnt[li + (((C >> b) & 1u) ? 2u : 0u)]++;
        }
        u32 off[4];
#pragma unroll
        for (int l = 0; l < 4; l++) {
            u32 c = cnt[l], s = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { u32 y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= (u32)o) s += y; }
            u32 tot = __shfl_sync(0xffffffffu, s, 31);
            u32 basel = l < 2 ? n_sep[l] : n_close[l - 2];
            off[l] = basel + s - c;
            if (l < 2) { if (n_sep[l] + tot > CK_LX_SEP) overflow = true; n_sep[l] += tot; }
            else { if (n_close[l - 2] + tot > CK_LX_CLOSE) overflow = true; n_close[l - 2] += tot; }
        }
        if (!overflow) {
            for (u32 x = pend; x; x &= x - 1) {
                u32 b = __ffs(x) - 1, below = (1u << b) - 1u;
                int d = dbase + (int)__popc(O & below) - (int)__popc(C & below);
                u32 li = d == 4 ? 0u : (d == 6 ? 1u : 2u);
                if (li >= 2) continue;
                u32 pos = p0 + b - m0;
                if ((C >> b) & 1u) lx->close_[li][off[li + 2]++] = pos; else lx->sep[li][off[li]++] = pos;
            }
        }
    }
    if (lane == 0) {
        lx->n_sep[0] = n_sep[0]; lx->n_sep[1] = n_sep[1]; lx->n_close[0] = n_close[0]; lx->n_close[1] = n_close[1];
        lx->ok = overflow ? 0u : 1u;                                // too many entries: plain lockstep walk, no element parallelism
    }
    __syncwarp();
}

#ifndef CK_LONG_MINB
#define CK_LONG_MINB 6          // <= 80 registers: 24 warps per SM (measured on the mixed workload: 3.6 ms against 5.3 ms at 16 warps)
#endif
__global__ void __launch_bounds__(32 * CK_LONG_WARPS, CK_LONG_MINB)
ck_walk_long_kernel(ck_view v, u32 n, u32* __restrict__ cols, u32 stride) {
    u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 count = v.canon_ctl->pad;                                   // records the thread-per-record kernel handed over
    ck_long_index* lx = (ck_long_index*)ck_win_smem + w;
    for (u32 k = blockIdx.x * CK_LONG_WARPS + w; k < count; k += gridDim.x * CK_LONG_WARPS) {
        u32 i = v.long_list[k];
        u32 len; const u8* rec = ck_rec_in(v, i, len);
        ck_lx_build(rec, len, lx);
        WalkOut o; o.base = cols + i; o.stride = stride; o.active = (lane == 0);
        URd r; r.init(rec, len, 2);                                 // state 2 (bit 0 belongs to the match cores): lockstep, lists may fan out
        AnyCtx cx; cx.kfill = 0;
        u32 stop = 0;
        u32 status = ck_walk_envelope(r, o, cx, stop) ? CK_OK : CK_NOT_CANONICAL;
        o.set(CK_COL_STATUS, status);
        o.set(CK_COL_ERR, stop);
        if (lane == 0 && status == CK_NOT_CANONICAL && v.canon_ctl) {
            u32 kk = atomicAdd(&v.canon_ctl->count, 1u);
            v.canon_list[kk] = i;
        }
        __syncwarp();
    }
}

#endif  // CK_WALK_LONG_CUH
