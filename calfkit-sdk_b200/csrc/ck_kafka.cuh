// Kafka RecordBatch v2 framing on the device.
//
// Reference: done by aiokafka underneath broker.subscriber(...) / broker.publish(...) (calfkit/worker/worker.py:45-53,
// calfkit/nodes/base.py:82-87; faststream[kafka] -> aiokafka, third-party, absent from the image).  Format restated in
// oracle/kafka_batch.py from the published protocol (message format v2) and pinned there to known-answer vectors.
//
// decode (a fetch response's record set = concatenated frames, copied to HBM as it arrived):
//   ck_rb_crc_kernel     warp per batch: CRC32C (Castagnoli) of the covered bytes [21, end) — every lane runs a
//                        slicing-by-4 table CRC over its own contiguous chunk, the 32 partial CRCs are folded with
//                        GF(2) polynomial arithmetic (crc(A||B) = crc(A) * x^(8|B|) mod P  xor  crc(B));
//   ck_rb_split_kernel   thread per batch: hop from record to record (each starts with its own varint length)
//   ck_rb_fields_kernel  thread per record: zig-zag varints -> key span, value span, the `correlation_id` header span
// The walker then reads every value where it lies inside the raw buffer (ck_view.len): no per-record slicing on the host.
//
// encode (produce): the payloads of one topic-partition (an index list the producer side already has) become one frame:
//   ck_rb_size_kernel    thread per record: encoded size (varints depend on the record's offsetDelta = its rank)
//   (scan)               -> byte offset of every record inside the frame
//   ck_rb_write_kernel   warp per record: varint header, key, value, the two headers calfkit's publishes carry
//   ck_rb_header_kernel + ck_rb_crc_chunks/fold: the 61-byte batch header, CRC32C over the frame in 16 KB chunks
#ifndef CK_KAFKA_CUH
#define CK_KAFKA_CUH

#define CK_CRC_POLY 0x82F63B78u
#define CK_RB_HEADER 61u
#define CK_RB_CRC_FROM 21u

// ---- CRC32C -----------------------------------------------------------------------------------------------------------
// four 256-entry tables (slicing-by-4) built in shared memory by the block itself
__device__ __forceinline__ void ck_crc_tables(u32* t /* [4][256] */) {
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
        u32 c = i;
#pragma unroll
        for (int k = 0; k < 8; k++) c = (c >> 1) ^ ((c & 1u) ? CK_CRC_POLY : 0u);
        t[i] = c;
    }
    __syncthreads();
    for (u32 i = threadIdx.x; i < 256; i += blockDim.x) {
        u32 c = t[i];
#pragma unroll
        for (int k = 1; k < 4; k++) { c = (c >> 8) ^ t[c & 0xffu]; t[k * 256 + i] = c; }
    }
    __syncthreads();
}
// finalised CRC32C of p[0..n) (init / xor-out all ones)
__device__ __forceinline__ u32 ck_crc32c_span(const u32* __restrict__ t, const u8* __restrict__ p, u32 n) {
    u32 c = 0xffffffffu;
    while (n && ((uintptr_t)p & 3u)) { c = t[(c ^ *p++) & 0xffu] ^ (c >> 8); n--; }
    const u32* w = (const u32*)p;
    for (; n >= 4; n -= 4) {
        c ^= __ldg(w++);
        c = t[768 + (c & 0xffu)] ^ t[512 + ((c >> 8) & 0xffu)] ^ t[256 + ((c >> 16) & 0xffu)] ^ t[c >> 24];
    }
    p = (const u8*)w;
    while (n--) c = t[(c ^ *p++) & 0xffu] ^ (c >> 8);
    return ~c;
}
// a(x) * b(x) mod P(x), reflected representation (bit 31 = x^0)
__device__ __forceinline__ u32 ck_gf_mul(u32 a, u32 b) {
    u32 p = 0;
#pragma unroll 4
    for (u32 m = 0x80000000u; m; m >>= 1) {
        if (a & m) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ CK_CRC_POLY : b >> 1;
    }
    return p;
}
// x^(8 n) mod P
__device__ __forceinline__ u32 ck_gf_xpow8(unsigned long long n) {
    u32 p = 0x80000000u;                       // x^0
    u32 sq = 0x00800000u;                      // x^8
    while (n) { if (n & 1ull) p = ck_gf_mul(sq, p); sq = ck_gf_mul(sq, sq); n >>= 1; }
    return p;
}
__device__ __forceinline__ u32 ck_crc_combine(u32 crc_a, u32 crc_b, u32 xpow_len_b) { return ck_gf_mul(xpow_len_b, crc_a) ^ crc_b; }

// CRC32C of one span by one warp (all lanes return it): lane 0 takes the odd-sized head so that every other lane holds
// exactly `chunk` bytes and the tree fold shifts by multiples of one precomputed power
__device__ __forceinline__ u32 ck_crc32c_warp(const u32* __restrict__ t, const u8* __restrict__ p, u32 n, u32 lane) {
    u32 chunk = (n >> 5) & ~3u;
    u32 head = n - 31u * chunk;
    u32 c = lane == 0 ? ck_crc32c_span(t, p, head) : (chunk ? ck_crc32c_span(t, p + head + (lane - 1) * chunk, chunk) : 0u);
    if (!chunk) return __shfl_sync(0xffffffffu, c, 0);
    u32 xp = ck_gf_xpow8(chunk);               // x^(8 chunk); squared at every level: right blocks hold 1, 2, 4, 8, 16 chunks
#pragma unroll
    for (u32 s = 1; s < 32; s <<= 1) {
        u32 right = __shfl_down_sync(0xffffffffu, c, s);
        if ((lane & (2 * s - 1)) == 0) c = ck_crc_combine(c, right, xp);
        xp = ck_gf_mul(xp, xp);
    }
    return __shfl_sync(0xffffffffu, c, 0);
}

__device__ __forceinline__ u32 ck_be32(const u8* p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | p[3]; }
__device__ __forceinline__ void ck_put_be32(u8* p, u32 v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
__device__ __forceinline__ void ck_put_be64(u8* p, unsigned long long v) { ck_put_be32(p, (u32)(v >> 32)); ck_put_be32(p + 4, (u32)v); }

// ---- decode -------------------------------------------------------------------------------------------------------------
// batch_off[b] .. batch_off[b+1]: frame b inside buf (indexed by the host: the frames chain through their own length field)
__global__ void __launch_bounds__(256)
ck_rb_crc_kernel(const u8* __restrict__ buf, const long long* __restrict__ batch_off, u32 nb, u32* __restrict__ batch_bad) {
    __shared__ u32 t[1024];
    ck_crc_tables(t);
    u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nb) return;
    const u8* f = buf + batch_off[w];
    u32 flen = (u32)(batch_off[w + 1] - batch_off[w]);
    u32 bad = 0;
    if (flen < CK_RB_HEADER || f[16] != 2 || ((f[21] << 8 | f[22]) & 7)) bad = 1;      // not magic 2 / compressed: not handled here
    else {
        u32 c = ck_crc32c_warp(t, f + CK_RB_CRC_FROM, flen - CK_RB_CRC_FROM, lane);
        if (c != ck_be32(f + 17)) bad = 1;
    }
    if (lane == 0) batch_bad[w] = bad;
}

__device__ __forceinline__ bool ck_varint(const u8* __restrict__ p, u32& pos, u32 end, long long& out) {
    unsigned long long u = 0; u32 shift = 0;
    for (;;) {
        if (pos >= end || shift > 63) return false;
        u8 b = p[pos++];
        u |= (unsigned long long)(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
        shift += 7;
    }
    out = (long long)(u >> 1) ^ -(long long)(u & 1);
    return true;
}

// thread per batch: where does each record start.  rec_base[b] = records in the batches before b (host prefix sum of
// the recordsCount header fields); a batch whose records do not add up to its length is marked bad
__global__ void __launch_bounds__(128)
ck_rb_split_kernel(const u8* __restrict__ buf, const long long* __restrict__ batch_off, const u32* __restrict__ rec_base, u32 nb,
                   u32* __restrict__ batch_bad, long long* __restrict__ rec_pos, u32* __restrict__ rec_batch) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    u32 r0 = rec_base[b], cnt = rec_base[b + 1] - r0;
    long long base = batch_off[b];
    const u8* f = buf + base;
    u32 flen = (u32)(batch_off[b + 1] - base);
    u32 pos = CK_RB_HEADER;
    bool ok = !batch_bad[b];
    for (u32 k = 0; k < cnt; k++) {
        rec_batch[r0 + k] = b;
        if (!ok) { rec_pos[r0 + k] = -1; continue; }
        rec_pos[r0 + k] = base + pos;
        long long len;
        if (!ck_varint(f, pos, flen, len) || len < 0 || pos + (u32)len > flen) { ok = false; rec_pos[r0 + k] = -1; continue; }
        pos += (u32)len;
    }
    if (ok && pos != flen) ok = false;
    if (!ok) batch_bad[b] = 1;
}

// thread per record
__global__ void __launch_bounds__(128)
ck_rb_fields_kernel(const u8* __restrict__ buf, long long buf_len, const long long* __restrict__ rec_pos, const u32* __restrict__ rec_batch,
                    const u32* __restrict__ batch_bad, u32 n,
                    long long* __restrict__ val_off, u32* __restrict__ val_len, long long* __restrict__ key_off, int* __restrict__ key_len,
                    long long* __restrict__ corr_off, int* __restrict__ corr_len, u32* __restrict__ rec_bad) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    val_off[i] = 0; val_len[i] = 0; key_off[i] = 0; key_len[i] = -1; corr_off[i] = 0; corr_len[i] = -1;
    long long rp = rec_pos[i];
    if (rp < 0 || batch_bad[rec_batch[i]]) { rec_bad[i] = 1; return; }
    const u8* p = buf + rp;
    u32 end = (u32)min((long long)0x7fffffff, buf_len - rp), pos = 0;
    long long len, t;
    bool ok = ck_varint(p, pos, end, len);
    u32 rec_end = pos + (u32)len;
    if (ok) end = rec_end;
    pos += 1;                                                   // attributes
    ok = ok && ck_varint(p, pos, end, t) && ck_varint(p, pos, end, t);       // timestampDelta, offsetDelta
    ok = ok && ck_varint(p, pos, end, t);                       // key
    if (ok && t >= 0) { key_off[i] = rp + pos; key_len[i] = (int)t; pos += (u32)t; }
    ok = ok && pos <= end && ck_varint(p, pos, end, t);         // value
    if (ok && t >= 0) { val_off[i] = rp + pos; val_len[i] = (u32)t; pos += (u32)t; }
    long long nh = 0;
    ok = ok && pos <= end && ck_varint(p, pos, end, nh);
    for (long long hcount = 0; ok && hcount < nh; hcount++) {
        long long kl, vl;
        ok = ck_varint(p, pos, end, kl) && kl >= 0 && pos + (u32)kl <= end;
        if (!ok) break;
        u32 kpos = pos; pos += (u32)kl;
        ok = ck_varint(p, pos, end, vl);
        if (!ok) break;
        bool is_corr = (kl == 14);                              // "correlation_id" (FastStream's header)
        const char* want = "correlation_id";
        for (u32 b = 0; is_corr && b < 14; b++) is_corr = (p[kpos + b] == (u8)want[b]);
        if (vl >= 0) { if (is_corr) { corr_off[i] = rp + pos; corr_len[i] = (int)vl; } pos += (u32)vl; }
        ok = pos <= end;
    }
    ok = ok && pos == rec_end;
    rec_bad[i] = ok ? 0u : 1u;
    if (!ok) { val_len[i] = 0; }
}

__global__ void __launch_bounds__(256)
ck_rb_mark_bad_kernel(const u32* __restrict__ rec_bad, u32 n, u32* __restrict__ cols, u32 stride) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && rec_bad[i]) cols[(size_t)CK_COL_STATUS * stride + i] = CK_BAD_FRAME;
}

// ---- encode -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 ck_varint_size(long long v) {
    unsigned long long u = ((unsigned long long)v << 1) ^ (unsigned long long)(v >> 63);
    u32 n = 1;
    while (u >= 0x80) { u >>= 7; n++; }
    return n;
}
__device__ __forceinline__ u32 ck_put_varint(u8* p, long long v) {
    unsigned long long u = ((unsigned long long)v << 1) ^ (unsigned long long)(v >> 63);
    u32 n = 0;
    while (u >= 0x80) { p[n++] = (u8)(u | 0x80); u >>= 7; }
    p[n++] = (u8)u;
    return n;
}
// what calfkit's publishes carry: key = correlation id bytes when keyed, headers
// content-type: application/json and correlation_id (FastStream)
#define CK_RB_CT_HDR 30u             // varint(12) "content-type" varint(16) "application/json" = 1 + 12 + 1 + 16
struct ck_rb_rec { u32 body_len, key_len, corr_len, val_len; };

__device__ __forceinline__ ck_rb_rec ck_rb_measure(u32 rank, u32 val_len, u32 corr_len, bool keyed) {
    ck_rb_rec r; r.val_len = val_len; r.corr_len = corr_len; r.key_len = keyed ? corr_len : 0xffffffffu;
    u32 body = 1 + 1 + ck_varint_size(rank);                                    // attributes, timestampDelta 0, offsetDelta
    body += keyed ? ck_varint_size(corr_len) + corr_len : 1;                    // key (or null)
    body += ck_varint_size(val_len) + val_len;
    body += 1 + CK_RB_CT_HDR + (1 + 14 + ck_varint_size(corr_len) + corr_len);  // 2 headers
    r.body_len = body;
    return r;
}

// idx[k] = index into the publish table of the k-th record of this topic-partition (its offsetDelta = k)
__global__ void __launch_bounds__(256)
ck_rb_size_kernel(const ck_pub* __restrict__ pubs, const u32* __restrict__ idx, u32 n, const u32* __restrict__ pay_len,
                  const u32* __restrict__ cols, u32 stride, u32* __restrict__ sizes) {
    u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    ck_pub p = pubs[idx[k]];
    ck_rb_rec r = ck_rb_measure(k, pay_len[p.payload], cols[(size_t)CK_COL_CORR_LEN * stride + p.record], p.has_key != 0);
    sizes[k] = ck_varint_size(r.body_len) + r.body_len;
}

// warp per record; corr bytes are the raw JSON string content of deps.correlation_id (ids without escapes: the hashing
// rule of ck_murmur2_key applies to escaped ones — those are rejected here by the host side before encoding)
__global__ void __launch_bounds__(256)
ck_rb_write_kernel(ck_view vw, const ck_pub* __restrict__ pubs, const u32* __restrict__ idx, u32 n, const u32* __restrict__ pay_len,
                   const long long* __restrict__ out_off, const u8* __restrict__ out, const u32* __restrict__ cols, u32 stride,
                   const long long* __restrict__ rec_off, u8* __restrict__ frame) {
    u32 k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= n) return;
    ck_pub p = pubs[idx[k]];
    u32 rl; const u8* rec = ck_rec(vw, p.record, rl);
    u32 co = cols[(size_t)CK_COL_CORR_OFF * stride + p.record], cl = cols[(size_t)CK_COL_CORR_LEN * stride + p.record];
    ck_rb_rec r = ck_rb_measure(k, pay_len[p.payload], cl, p.has_key != 0);
    u8* d = frame + CK_RB_HEADER + rec_off[k];
    u32 pos = 0, vpos = 0, kpos = 0, cpos = 0;
    // every lane computes the same layout; lane 0 writes the small pieces, the warp copies key / value / header value
    u8 tmp[10];
    u32 m = ck_put_varint(tmp, r.body_len); if (lane == 0) for (u32 b = 0; b < m; b++) d[pos + b] = tmp[b]; pos += m;
    if (lane == 0) { d[pos] = 0; d[pos + 1] = 0; } pos += 2;                                  // attributes, timestampDelta = 0
    m = ck_put_varint(tmp, k); if (lane == 0) for (u32 b = 0; b < m; b++) d[pos + b] = tmp[b]; pos += m;
    if (p.has_key) { m = ck_put_varint(tmp, cl); if (lane == 0) for (u32 b = 0; b < m; b++) d[pos + b] = tmp[b]; pos += m; kpos = pos; pos += cl; }
    else { if (lane == 0) d[pos] = 1; pos += 1; }                                             // zig-zag(-1) = null key
    m = ck_put_varint(tmp, r.val_len); if (lane == 0) for (u32 b = 0; b < m; b++) d[pos + b] = tmp[b]; pos += m; vpos = pos; pos += r.val_len;
    if (lane == 0) {
        d[pos] = 4;                                                                           // 2 headers
        d[pos + 1] = 24; const char* a = "content-type"; for (u32 b = 0; b < 12; b++) d[pos + 2 + b] = (u8)a[b];
        d[pos + 14] = 32; const char* j = "application/json"; for (u32 b = 0; b < 16; b++) d[pos + 15 + b] = (u8)j[b];
        d[pos + 31] = 28; const char* c = "correlation_id"; for (u32 b = 0; b < 14; b++) d[pos + 32 + b] = (u8)c[b];
    }
    pos += 1 + CK_RB_CT_HDR + 15;
    m = ck_put_varint(tmp, cl); if (lane == 0) for (u32 b = 0; b < m; b++) d[pos + b] = tmp[b]; pos += m; cpos = pos;
    if (p.has_key) for (u32 b = lane; b < cl; b += 32) d[kpos + b] = rec[co + b];
    for (u32 b = lane; b < cl; b += 32) d[cpos + b] = rec[co + b];
    ck_warp_copy(d + vpos, out + out_off[p.payload], r.val_len, lane);
}

// frame header (everything but the crc), by one thread
__global__ void ck_rb_header_kernel(u8* __restrict__ frame, const long long* __restrict__ rec_off, u32 n, long long base_offset,
                                    long long timestamp_ms, long long* __restrict__ frame_len_out) {
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long total = CK_RB_HEADER + (unsigned long long)rec_off[n];
    ck_put_be64(frame, (unsigned long long)base_offset);
    ck_put_be32(frame + 8, (u32)(total - 12));
    ck_put_be32(frame + 12, 0);                              // partitionLeaderEpoch
    frame[16] = 2;
    frame[21] = 0; frame[22] = 0;                            // attributes: no compression, CreateTime
    ck_put_be32(frame + 23, n - 1);                          // lastOffsetDelta
    ck_put_be64(frame + 27, (unsigned long long)timestamp_ms);
    ck_put_be64(frame + 35, (unsigned long long)timestamp_ms);
    ck_put_be64(frame + 43, ~0ull);                          // producerId -1
    frame[51] = 0xff; frame[52] = 0xff;                      // producerEpoch -1
    ck_put_be32(frame + 53, 0xffffffffu);                    // baseSequence -1
    ck_put_be32(frame + 57, n);
    *frame_len_out = (long long)total;
}

// CRC32C of a large frame: one warp per 16 KB chunk of the covered bytes, then one warp folds the partials
#define CK_RB_CRC_CHUNK 16384u
__global__ void __launch_bounds__(256)
ck_rb_crc_chunks_kernel(const u8* __restrict__ frame, const long long* __restrict__ frame_len, u32* __restrict__ partial) {
    __shared__ u32 t[1024];
    ck_crc_tables(t);
    u32 w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    unsigned long long covered = (unsigned long long)*frame_len - CK_RB_CRC_FROM;
    unsigned long long start = (unsigned long long)w * CK_RB_CRC_CHUNK;
    if (start >= covered) return;
    u32 len = (u32)min((unsigned long long)CK_RB_CRC_CHUNK, covered - start);
    u32 c = ck_crc32c_warp(t, frame + CK_RB_CRC_FROM + start, len, lane);
    if (lane == 0) partial[w] = c;
}
__global__ void ck_rb_crc_fold_kernel(u8* __restrict__ frame, const long long* __restrict__ frame_len, const u32* __restrict__ partial) {
    if (blockIdx.x || threadIdx.x) return;
    unsigned long long covered = (unsigned long long)*frame_len - CK_RB_CRC_FROM;
    u32 nchunks = (u32)((covered + CK_RB_CRC_CHUNK - 1) / CK_RB_CRC_CHUNK);
    u32 xp = ck_gf_xpow8(CK_RB_CRC_CHUNK);
    u32 c = 0;                                               // crc of the empty prefix
    for (u32 k = 0; k < nchunks; k++) {
        unsigned long long start = (unsigned long long)k * CK_RB_CRC_CHUNK;
        u32 len = (u32)min((unsigned long long)CK_RB_CRC_CHUNK, covered - start);
        c = ck_crc_combine(c, partial[k], len == CK_RB_CRC_CHUNK ? xp : ck_gf_xpow8(len));
    }
    ck_put_be32(frame + 17, c);
}

#endif  // CK_KAFKA_CUH
