// Cross-partition forward over NVLink peer memory.
//
// Records shard by Kafka partition across the GPUs of a box; a keyed publish whose partition is owned by another rank
// (the reference would simply produce it to a topic-partition another worker consumes, nodes/base.py:82-87) is forwarded
// to that rank.  Round 1 packed the payloads and called NCCL three times with two host synchronisations per step; here
// pack and transfer are ONE kernel: every rank maps its peers' receive buffers (CUDA IPC over NVSwitch), and after the
// exchange plan (per-destination histogram -> scan -> stable scatter -> scan of padded lengths, all on the device) a warp
// per selected payload copies it from this rank's output buffer straight into the owner's receive region for this source,
// next to a 16-byte meta entry; a one-block kernel then writes the region headers.  No sizes travel through the host and no
// library collective is left on the path: the two barriers of a step (before the stores: every peer has consumed what it
// received last time; after: everybody's stores have landed) are flags in peer memory too (ck_xbarrier_kernel: one block,
// thread d stores the step number into peer d's flag word for this rank, then spins on its own flag word for rank d).
// Measured on 2 GPUs: with NCCL all-reduces as barriers the exchange never overlapped the other lane's kernels — NCCL's
// 640-thread blocks do not fit into the gaps the million-record kernels leave, even on a high-priority stream; a 32-thread
// block does.
//
// receive buffer of a rank: `world` regions of `region_stride` bytes, region s = what source rank s forwarded:
//   [0]  u64 step      [8] u32 count   [12] u32 overflow   [16] u64 payload bytes (16-byte padded)
//   [64] meta[max_fwd] {u32 len, i32 topic_id, i32 partition, u32 source publish index}
//   [64 + 16 * max_fwd] payload bytes, every payload starts 16-byte aligned
#ifndef CK_XSEND_CUH
#define CK_XSEND_CUH

struct ck_xregion_hdr { unsigned long long step; u32 count, overflow; unsigned long long nbytes; };
struct ck_xmeta { u32 len; int32_t topic_id; int32_t partition; u32 src_pub; };
#define CK_X_HDR 64u

struct ck_xpeers { u8* recv[CK_X_MAXWORLD]; };     // peer receive buffers as mapped into this process (own rank: own buffer)

// flag words live behind the regions of the receive buffer: [slot 0 | slot 1][source rank] u64
#define CK_X_FLAGS_BYTES (2 * CK_X_MAXWORLD * 8)
__global__ void ck_xbarrier_kernel(ck_xpeers peers, u32 rank, u32 world, unsigned long long flags_off, u32 slot, unsigned long long value,
                                   u32* __restrict__ timeout_flag) {
    u32 d = threadIdx.x;
    if (d >= world || d == rank) return;
    __threadfence_system();                                         // everything this rank stored before (payloads, headers) first
    *(volatile unsigned long long*)(peers.recv[d] + flags_off + ((size_t)slot * CK_X_MAXWORLD + rank) * 8) = value;
    volatile unsigned long long* mine = (volatile unsigned long long*)(peers.recv[rank] + flags_off + ((size_t)slot * CK_X_MAXWORLD + d) * 8);
    long long t0 = clock64();
    while (*mine < value) {
        if (clock64() - t0 > 8000000000ll) { *timeout_flag = 1; break; }       // ~4 s: a peer is gone; report, never hang the GPU
    }
    __threadfence_system();
}

__global__ void __launch_bounds__(256)
ck_xsend_kernel(const ck_pub* __restrict__ pubs, const u32* __restrict__ x_pub, const long long* __restrict__ x_src_off, const u32* __restrict__ x_len32,
                const long long* __restrict__ x_dst_off, const long long* __restrict__ base /* [world * nb + 1] */, u32 nb,
                const unsigned long long* __restrict__ total_sel, const u8* __restrict__ out, ck_xpeers peers, u32 rank, u32 world,
                unsigned long long region_stride, u32 max_fwd, unsigned long long data_cap, u32* __restrict__ overflow) {
    u32 lane = threadIdx.x & 31;
    u32 total = (u32)*total_sel;
    for (u32 slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slot < total; slot += (gridDim.x * blockDim.x) >> 5) {
    u32 pj = x_pub[slot];
    ck_pub p = pubs[pj];
    u32 d = (u32)p.partition % world;
    long long first = base[(size_t)d * nb];
    u32 local = (u32)(slot - first);
    unsigned long long boff = (unsigned long long)(x_dst_off[slot] - x_dst_off[first]);
    u32 len = x_len32[slot];
    if (local >= max_fwd || boff + ((len + 15u) & ~15u) > data_cap) { if (lane == 0) atomicAdd(overflow + d, 1u); continue; }
    u8* region = peers.recv[d] + (size_t)rank * region_stride;
    if (lane == 0) {
        ck_xmeta m; m.len = len; m.topic_id = p.topic_id; m.partition = p.partition; m.src_pub = pj;
        *(ck_xmeta*)(region + CK_X_HDR + (size_t)local * sizeof(ck_xmeta)) = m;
    }
    ck_warp_copy(region + CK_X_HDR + (size_t)max_fwd * sizeof(ck_xmeta) + boff, out + x_src_off[slot], len, lane);
    }
}

// region headers, after the payload kernel (stream order): thread d writes the header of this rank's region at peer d
__global__ void ck_xhdr_kernel(const long long* __restrict__ x_dst_off, const long long* __restrict__ base, u32 nb,
                               const unsigned long long* __restrict__ total_sel, ck_xpeers peers, u32 rank, u32 world,
                               unsigned long long region_stride, unsigned long long step, u32* __restrict__ overflow) {
    u32 d = threadIdx.x;
    if (d >= world || d == rank) return;
    long long first = base[(size_t)d * nb];
    long long next = d + 1 < world ? base[(size_t)(d + 1) * nb] : (long long)*total_sel;
    ck_xregion_hdr h;
    h.step = step; h.count = (u32)(next - first); h.overflow = overflow[d];
    h.nbytes = (unsigned long long)(x_dst_off[next] - x_dst_off[first]);
    if (h.overflow) { h.count = 0; h.nbytes = 0; }                 // a region that did not fit is reported, never half-delivered
    *(ck_xregion_hdr*)(peers.recv[d] + (size_t)rank * region_stride) = h;
    overflow[d] = 0;
}

#endif  // CK_XSEND_CUH
