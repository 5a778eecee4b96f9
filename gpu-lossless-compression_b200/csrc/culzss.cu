// CULZSS-compatible LZSS for sm_100a  (hot path 2).
//
// Bit-exact with the reference encoder/decoder (cuda-lzss-cluster): WINDOW_SIZE 128,
// MAX_CODED 128, MAX_UNCODED 2, 4096-byte packets (gpu_compress.h:62-69), the single-pass
// "streak scanner" match finder (gpu_compress.cu:104-168), greedy token selection with flag
// bytes (aftercomp, :462-566) and the per-buffer trailer (:624-657).
//
// What is different from the reference:
//   * match finding is exact but not literal: for position p the scanner visits scan index
//     t <- t + LCP(t) + 1 over the 127 window positions; window
//     positions whose first byte differs (LCP 0) are skipped in bulk through a per-value
//     occurrence bitmask of the sliding window kept in shared memory, so the work per position
//     is proportional to the number of visited candidates, not to the window size;
//   * the 2-bytes-per-input-byte token array never leaves the SM: selection (the CPU stage
//     `aftercomp`) and flag-byte packing run in the same kernel on the packet in shared memory;
//   * a second kernel concatenates the packets of a buffer and writes the trailer, replacing
//     aftercompression_wrapper's serial CPU loop and the 2 MiB D2H per MiB of input;
//   * the decoder runs one packet per LANE (the reference: one packet per single-thread CTA,
//     gpu_decompress.cu:120-244 with <<<lSize/4096, 1>>>) with the 128-byte window and the
//     output staging unified in one bank-private shared-memory ring per lane, 16-byte global
//     accesses, and input refills / output flushes at fixed points of the token loop.
#include "common.cuh"
#include "culzss_lane.cuh"
#include "../../include/b200lc.h"

namespace b200lc {
namespace lzss {

constexpr int kWindow = 128;
constexpr int kPacket = 4096;
constexpr int kChunks = kPacket / 128;
constexpr int kMaxPacketOut = kPacket + kPacket / 8;   // all literals: 4096 + 512 flag bytes
constexpr int kOccStride = 9;                          // 8 ring words + 1 pad (bank spread)

// ====================================================================================== encode
constexpr int kLevels = 7;                             // run-length planes cover L = 1..8
struct EncSmem {
    __align__(16) u8 pkt[kWindow + kPacket + 16];   // pkt[128 + q] = P[q]; pkt[0..127] = ' '
    union {
        struct {
            u32 occ[256 * kOccStride];              // per byte value: ring of 256 window slots
            uint4 ymask[128 + kLevels + 1];         // occurrence masks of the chunk's positions (+7 of the next)
        } a;                                        // match finding
        struct {
            __align__(16) u8 out[kMaxPacketOut + 16];
            u16 FB[kPacket / 8 + 1];                // output offset of the flag byte of group g
        } b;                                        // packing
        struct {
            short prev[kPacket + kWindow];          // fast mode: previous position with the same hash
        } f;
        struct {
            u16 J[kPacket];                         // first token position selected after pos's 32-block
            u8 E[kPacket / 32];                     // entry offset of the parse into each 32-block
        } s;                                        // selection
    } u;
    u8 tlen[kPacket];                               // token: match length, or 1 for a literal
    u8 toff[kPacket];                               // token: ring offset, or the literal byte
    u32 M[kPacket / 32];                            // bit p: token p is a match (len >= 3)
    u32 V[kPacket / 32];                            // bit p: token p is selected by the greedy parse
    u32 scan[4];
};

// four bytes at byte offset `at` of the packet buffer (two aligned words, funnel-shifted)
__device__ __forceinline__ u32 load4(const u8 *pkt, u32 at)
{
    const u32 *w = reinterpret_cast<const u32 *>(pkt + (at & ~3u));
    return __funnelshift_r(w[0], w[1], 8 * (at & 3));
}

// DEPTH == 0: the reference's match finder, bit-exact (parity mode).
// DEPTH  > 0: FAST MODE, NOT bit-exact with the reference encoder: the same token format, window
//   and packet layout (every stream decodes with the reference's DecodeKernel,
//   gpu_decompress.cu:164-242), but the match of a position is the longest among the DEPTH most
//   recent earlier positions whose first three bytes hash alike (shared-memory hash chain), not the
//   result of the reference's streak scanner over all 127 window positions.  A match never reaches
//   into its own output (source end <= current position), like the reference's.
// Packets [first, npackets).  sel != nullptr: run only if *sel == want (AUTO mode launches both
// parity kernels behind a probe; the one the probe did not pick returns at once).
template <int DEPTH>
__global__ void __launch_bounds__(128) culzss_encode_kernel(const u8 *__restrict__ in, u64 first, u64 npackets,
                                                            u8 *__restrict__ tmp_out,
                                                            u16 *__restrict__ pkt_size,
                                                            u8 *__restrict__ last_group_size,
                                                            const int *__restrict__ sel, int want)
{
    if (sel && *sel != want) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncSmem &sm = *reinterpret_cast<EncSmem *>(smem_raw);
    const u32 tx = threadIdx.x;
    const u32 lane = tx & 31, warp = tx >> 5;

    for (u64 pid = first + blockIdx.x; pid < npackets; pid += gridDim.x) {
        // ---------------------------------------------------------------- load packet, reset state
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(in + pid * kPacket);
            uint4 *dst = reinterpret_cast<uint4 *>(sm.pkt + kWindow);
            dst[tx] = src[tx];
            dst[tx + 128] = src[tx + 128];
            sm.pkt[tx] = ' ';
            if (DEPTH == 0) {
                for (u32 i = tx; i < 256 * kOccStride; i += 128) sm.u.a.occ[i] = 0;
            } else {
                uint4 *const T0 = reinterpret_cast<uint4 *>(smem_raw + ((sizeof(EncSmem) + 15) & ~size_t(15)));
                for (u32 i = tx; i < (1u << 10); i += 128) T0[i] = make_uint4(0, 0, 0, 0);
                sm.u.f.prev[tx] = (short)-32768;          // positions -128 .. -1: no predecessor
                if (tx < 4) reinterpret_cast<u32 *>(sm.pkt + kWindow + kPacket)[tx] = 0;   // defined padding
            }
        }
        __syncthreads();
        if (DEPTH > 0) {
            // Hash table behind EncSmem: per hash of three bytes EIGHT 16-bit slots, one per
            // (chunk parity, warp): slot [c & 1][w] holds the latest position (+ 129) that warp w
            // met in a chunk of parity c & 1.  A reader of chunk c sees its own chunk's lower warps
            // and the whole previous chunk with ONE 16-byte load; anything staler fails the window
            // test by construction, so the table is never cleared between chunks.
            uint4 *const T = reinterpret_cast<uint4 *>(smem_raw + ((sizeof(EncSmem) + 15) & ~size_t(15)));
            u16 *const T16 = reinterpret_cast<u16 *>(T);
            // the window in front of the packet is 128 spaces: the farthest one can lend the longest
            // run; position -128 = chunk -1 (parity 1), warp 0
            if (tx == 0) T16[(((0x202020u * 2654435761u) >> 22) << 3) + 4] = 1;
            __syncthreads();
            for (u32 c = 0; c < kChunks; ++c) {
                const u32 p = c * 128 + tx;
                const u32 v = sm.pkt[kWindow + p];
                const bool valid = p + 2 < (u32)kPacket;
                const u32 h = ((load4(sm.pkt, kWindow + p) & 0xffffffu) * 2654435761u) >> 22;
                const u32 peers = __match_any_sync(0xffffffffu, valid ? h : (0x80000000u | lane));
                if (valid && (peers >> lane) == 1u) T16[(h << 3) + ((c & 1) << 2) + warp] = (u16)(p + 129);
                __syncthreads();     // this chunk's positions are in the table
                // most recent earlier position with the same hash at distance 3..128 (a closer one
                // cannot lend three bytes): inside my warp from the peer mask, else from the table
                int cand = -32768;
                if (valid) {
                    const u32 lower = lane >= 3 ? (peers & ((1u << (lane - 2)) - 1u)) : 0u;
                    if (lower) {
                        cand = (int)(p - lane) + (31 - __clz(lower));
                    } else {
                        const uint4 tv = T[h];
                        const u32 lo = p + 1, hi = p + 126;          // position + 129 in [p - 128, p - 3]
                        const u32 lo2 = lo | (lo << 16), hi2 = hi | (hi << 16);
                        u32 m = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const u32 x = (&tv.x)[k];
                            m = __vmaxu2(m, x & __vcmpgeu2(x, lo2) & __vcmpleu2(x, hi2));
                        }
                        const u32 best = max(m & 0xffffu, m >> 16);
                        if (best) cand = (int)best - 129;
                    }
                }
                sm.u.f.prev[kWindow + p] = (short)cand;
                __syncthreads();     // links of this chunk visible
                u32 best_len = 1;
                int best_q = 0;
                if (valid) {
                    const u32 maxlen = min(127u, (u32)kPacket - p);
                    const u32 first = load4(sm.pkt, kWindow + p);
                    int q = cand;
#pragma unroll 1
                    for (int d = 0; d < DEPTH;) {
                        if (q < (int)p - kWindow) break;
                        const u32 dist = (u32)((int)p - q);
                        const u32 lim = min(dist, maxlen);
                        u32 L = 0;
                        u32 x = load4(sm.pkt, (u32)(kWindow + q)) ^ first;
                        while (true) {
                            if (x) { L += (u32)(__ffs(x) - 1) >> 3; break; }
                            L += 4;
                            if (L >= lim) break;
                            x = load4(sm.pkt, (u32)(kWindow + q) + L) ^ load4(sm.pkt, kWindow + p + L);
                        }
                        L = min(L, lim);
                        if (L > best_len) { best_len = L; best_q = q; }
                        // A match that ends only because it reached its own output (length ==
                        // distance) means the data repeats with that period: the source twice as far
                        // back lends twice as much (runs, periodic records).  Does not use up depth.
                        if (L == dist && L < maxlen && 2 * dist <= (u32)kWindow) {
                            q = (int)p - (int)(2 * dist);
                        } else {
                            q = sm.u.f.prev[kWindow + q];
                            ++d;
                        }
                    }
                }
                const bool is_match = best_len > 2;
                sm.tlen[p] = is_match ? (u8)best_len : (u8)1;
                sm.toff[p] = is_match ? (u8)((best_q + kWindow) & 255) : (u8)v;
                const u32 mm = __ballot_sync(0xffffffffu, is_match);
                if (lane == 0) sm.M[p >> 5] = mm;
                __syncthreads();     // tokens written
            }
        }
        // window slots 128..255 initially hold ' ' (gpu_compress.cu:208), chunk 0 enters slots 0..127
        if (DEPTH == 0) {
        if (tx < 4) sm.u.a.occ[0x20 * kOccStride + 4 + tx] = 0xffffffffu;
        __syncthreads();
        atomicOr(&sm.u.a.occ[sm.pkt[kWindow + tx] * kOccStride + (tx >> 5)], 1u << (tx & 31));
        __syncthreads();
        }

        for (u32 c = 0; DEPTH == 0 && c < kChunks; ++c) {
            const u32 p = c * 128 + tx;
            const u32 v = sm.pkt[kWindow + p];
            // ---- 128-bit occurrence mask over scan index t (t = 0 <-> position p-128)
            u32 Y[4];
            {
                const u32 sb = (p + 128) & 255;       // ring slot of position p - 128
                const u32 wo = sb >> 5, bo = sb & 31;
                const u32 *row = &sm.u.a.occ[v * kOccStride];
                u32 w[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) w[i] = row[(wo + i) & 7];
#pragma unroll
                for (int i = 0; i < 4; ++i) Y[i] = __funnelshift_r(w[i], w[i + 1], bo);
            }
            // scan length (gpu_compress.cu:120,149): 127, shrinking with tx in the last chunk
            const u32 n = (c == kChunks - 1) ? max(1u, 127u - tx) : 127u;
            // Publish the mask; masks of the first 7 positions of the next chunk come from
            // direct comparisons (their window is not in the ring yet).  Row r of ymask belongs
            // to position 128*c + r; bit t of a row: P[pos - 128 + t] == P[pos].
            sm.u.a.ymask[tx] = make_uint4(Y[0], Y[1], Y[2], Y[3]);
#pragma unroll
            for (int j = 0; j < kLevels; ++j) {
                const u32 pn = (c + 1) * 128 + j;
                const bool eq = (c + 1 < kChunks) && tx < 127 &&
                                sm.pkt[pn + tx] == sm.pkt[kWindow + pn];
                const u32 bal = __ballot_sync(0xffffffffu, eq);
                if (lane == 0) reinterpret_cast<u32 *>(&sm.u.a.ymask[128 + j])[warp] = bal;
            }
            __syncthreads();   // masks published; every thread has read the ring: slots may be recycled
            if (c + 1 < kChunks) {
                // clear the ring half that chunk c+1 is about to occupy (it holds chunk c-1)
                const u32 half = ((c + 1) & 1) * 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const u32 val = tx * 2 + (i >> 2);
                    sm.u.a.occ[val * kOccStride + half + (i & 3)] = 0;
                }
            }
            // Run-length planes: bit t of Y(p+k) is the comparison of byte k of the window string
            // at scan index t with byte k of the lookahead, so the AND over k = 0..j says L > j.
            // c0/c1/c2 hold min(L - 1, 7) bit-sliced per scan index.
            u32 c0[4], c1[4], c2[4];
            {
                uint4 r[kLevels];
#pragma unroll
                for (int k = 0; k < kLevels; ++k) r[k] = sm.u.a.ymask[tx + 1 + k];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const u32 a1 = Y[i] & (&r[0].x)[i];
                    const u32 a2 = a1 & (&r[1].x)[i];
                    const u32 a3 = a2 & (&r[2].x)[i];
                    const u32 a4 = a3 & (&r[3].x)[i];
                    const u32 a5 = a4 & (&r[4].x)[i];
                    const u32 a6 = a5 & (&r[5].x)[i];
                    const u32 a7 = a6 & (&r[6].x)[i];
                    c2[i] = a4;
                    c1[i] = (a2 & ~a4) | a6;
                    c0[i] = (a1 & ~a2) | (a3 & ~a4) | (a5 & ~a6) | a7;
                }
            }
            __syncthreads();   // ring half cleared; ymask rows consumed
            if (c + 1 < kChunks) {
                const u32 pn = p + 128;
                atomicOr(&sm.u.a.occ[sm.pkt[kWindow + pn] * kOccStride + ((pn & 255) >> 5)],
                         1u << (pn & 31));
            }

            // ---- streak scanner as an LCP walk over the candidate bits
            u32 best_len = 1, best_t = 0;
            u32 t = 0;
            const u8 *srcb = sm.pkt + p;             // srcb[t + k] = byte k of the window string at t
            const u8 *lab = sm.pkt + kWindow + p;    // lab[k]      = byte k of the lookahead
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                u32 w = Y[i];
                // keep scan indices in [t, n)
                const int lo = (int)t - 32 * i, hi = (int)n - 32 * i;
                if (lo >= 32 || hi <= 0) w = 0;
                else {
                    if (lo > 0) w &= ~((1u << lo) - 1u);
                    if (hi < 32) w &= (1u << hi) - 1u;
                }
                while (w) {
                    const u32 b = __ffs(w) - 1;
                    const u32 tt = 32 * i + b;
                    const u32 cap = n - tt;
                    const u32 cnt = ((c0[i] >> b) & 1u) | (((c1[i] >> b) & 1u) << 1) |
                                    (((c2[i] >> b) & 1u) << 2);
                    u32 L = 1 + cnt;
                    if (cnt == 7) {   // 8 or more: finish 4 bytes at a time (two unaligned words)
                        const u32 lim = cap < 128 ? cap : 128;
                        while (L + 4 <= lim) {
                            const u32 pa = p + tt + L, pb = kWindow + p + L;     // byte offsets in pkt
                            const u32 *wa = reinterpret_cast<const u32 *>(sm.pkt + (pa & ~3u));
                            const u32 *wb = reinterpret_cast<const u32 *>(sm.pkt + (pb & ~3u));
                            const u32 xa = __funnelshift_r(wa[0], wa[1], 8 * (pa & 3));
                            const u32 xb = __funnelshift_r(wb[0], wb[1], 8 * (pb & 3));
                            const u32 d = xa ^ xb;
                            if (d) { L += (__ffs(d) - 1) >> 3; goto lcp_done; }
                            L += 4;
                        }
                        while (L < cap && srcb[tt + L] == lab[L]) ++L;
                    lcp_done:;
                    }
                    L = min(L, cap);
                    if (L > best_len) { best_len = L; best_t = tt; }
                    t = tt + L + 1;
                    const int nlo = (int)t - 32 * i;
                    w = nlo >= 32 ? 0u : (w & ~((1u << nlo) - 1u));
                }
            }
            // gpu_compress.cu:251-274 / 313-342
            const bool is_match = best_len > 2;
            sm.tlen[p] = is_match ? (u8)best_len : (u8)1;
            sm.toff[p] = is_match ? (u8)((p + best_t) & 255) : (u8)v;
            const u32 mm = __ballot_sync(0xffffffffu, is_match);
            if (lane == 0) sm.M[p >> 5] = mm;
            __syncthreads();   // inserts of chunk c+1 visible; tokens of chunk c written
        }

        // ---------------------------------------------------------------- greedy selection (aftercomp)
        // The parse visits p -> p + step(p), step = 1 for a literal and len for a match
        // (gpu_compress.cu:500-517).  Three short phases instead of one 4096-step chain:
        //  S1  every thread resolves its 32 positions backwards: J[p] = first visited position
        //      beyond the block when the parse passes through p;
        //  S2  one thread hops block to block (<= 128 hops) and records where each block is entered;
        //  S3  every thread replays its block from the entry offset and sets the visited bits.
        {
            const u32 blk_end = 32 * (tx + 1);
            for (int i = 31; i >= 0; --i) {
                const u32 pos = 32 * tx + i;
                const u32 nxt = pos + sm.tlen[pos];
                sm.u.s.J[pos] = (u16)(nxt >= blk_end ? nxt : sm.u.s.J[nxt]);
            }
            sm.u.s.E[tx] = 0xff;
        }
        __syncthreads();
        if (tx == 0) {
            u32 p = 0;
            while (p < kPacket) {
                sm.u.s.E[p >> 5] = (u8)(p & 31);
                p = sm.u.s.J[p];
            }
        }
        __syncthreads();
        {
            const u32 e = sm.u.s.E[tx];
            u32 bits = 0;
            if (e != 0xff) {
                u32 pos = 32 * tx + e;
                const u32 blk_end = 32 * (tx + 1);
                while (pos < blk_end) {
                    bits |= 1u << (pos & 31);
                    pos += sm.tlen[pos];
                }
            }
            sm.V[tx] = bits;
        }
        __syncthreads();

        // ---------------------------------------------------------------- pack flag bytes + payload
        const u32 vt = sm.V[tx];
        const u32 mt = sm.M[tx] & vt;
        const u32 my = ((u32)(__popc(vt) + __popc(mt)) << 16) | (u32)__popc(vt);
        u32 incl = warp_incl_scan(my);
        if (lane == 31) sm.scan[warp] = incl;
        __syncthreads();
        u32 wbase = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const u32 s = sm.scan[i];
            if ((u32)i < warp) wbase += s;
        }
        const u32 excl = wbase + incl - my;
        const u32 tot = sm.scan[0] + sm.scan[1] + sm.scan[2] + sm.scan[3];
        const u32 ntok_total = tot & 0xffffu, pay_total = tot >> 16;
        const u32 out_size = pay_total + ((ntok_total + 7) >> 3);
        {
            u32 k = excl & 0xffffu, po = excl >> 16, bits = vt;
            while (bits) {
                const u32 b = __ffs(bits) - 1;
                bits &= bits - 1;
                const u32 pos = tx * 32 + b;
                const u32 o = po + (k >> 3) + 1;
                if ((k & 7) == 0) {
                    sm.u.b.FB[k >> 3] = (u16)(o - 1);
                    sm.u.b.out[o - 1] = 0;
                }
                if ((mt >> b) & 1) {
                    sm.u.b.out[o] = sm.tlen[pos];
                    sm.u.b.out[o + 1] = sm.toff[pos];
                    po += 2;
                } else {
                    sm.u.b.out[o] = sm.toff[pos];
                    po += 1;
                }
                ++k;
            }
        }
        __syncthreads();
        {
            u32 k = excl & 0xffffu, bits = vt;
            u32 *out32 = reinterpret_cast<u32 *>(sm.u.b.out);
            while (bits) {
                const u32 b = __ffs(bits) - 1;
                bits &= bits - 1;
                if (!((mt >> b) & 1)) {   // literal: flag bit set (gpu_compress.cu:503)
                    const u32 fb = sm.u.b.FB[k >> 3];
                    atomicOr(&out32[fb >> 2], (1u << (k & 7)) << (8 * (fb & 3)));
                }
                ++k;
            }
        }
        __syncthreads();
        // ---------------------------------------------------------------- write packet
        {
            uint4 *dst = reinterpret_cast<uint4 *>(tmp_out + pid * (u64)kMaxPacketOut);
            const uint4 *src = reinterpret_cast<const uint4 *>(sm.u.b.out);
            const u32 nvec = (out_size + 15) >> 4;
            for (u32 i = tx; i < nvec; i += 128) dst[i] = src[i];
            if (tx == 0) {
                pkt_size[pid] = (u16)out_size;
                last_group_size[pid] = (u8)(out_size - sm.u.b.FB[(ntok_total - 1) >> 3]);
            }
        }
        __syncthreads();
    }
}

// ====================================================================================== encode, lane mode
// One packet per LANE (csrc/culzss_lane.cuh), two match finders:
//   PARITY = false  FAST MODE (NON-PARITY): greedy parse through a lane-private 64-entry hash of
//                   three-byte prefixes;
//   PARITY = true   the reference's streak scanner, bit-exact, evaluated only at the positions the
//                   greedy selection visits (the reference and culzss_encode_kernel<0> above compute
//                   a match for all 4096 positions and drop the ones a longer match jumped over).
// A lane emits tokens and flag bytes as it goes and streams the result to the packet's slot -- no
// CTA barrier, no token arrays, no separate selection and packing passes.  All per-lane state sits in
// shared-memory columns laid out [word][lane] (one wavefront per access whatever the 32 packets
// of a warp are doing).  Input travels global -> registers (one 16-byte chunk ahead of its use,
// the line after next prefetched into L2) -> ring; output ring -> 16-byte stores.
constexpr int kLaneWarps = 2;      // 12 KiB of columns per warp: nine CTAs = 18 warps per SM
__device__ __forceinline__ void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

struct LaneDevIO {
    const u8 *src;      // the packet (16-byte aligned)
    u8 *dst;            // its output slot (16-byte aligned, kSlotBytes)
    __device__ __forceinline__ bool any(bool b) const { return __any_sync(0xffffffffu, b) != 0; }
    __device__ __forceinline__ void load(u32 off, lzss_lane::Chunk32 &c) const
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(src + off);
        const uint4 a = __ldg(g), b = __ldg(g + 1);
        c.w[0] = a.x; c.w[1] = a.y; c.w[2] = a.z; c.w[3] = a.w;
        c.w[4] = b.x; c.w[5] = b.y; c.w[6] = b.z; c.w[7] = b.w;
        // the line after next, once per 128-byte line
        if ((off & 127u) == 0 && off + 256u < lzss_lane::kPacket) prefetch_l2(src + off + 256u);
    }
    __device__ __forceinline__ u32 bytes4(u32 off) const      // rare: lookahead beyond the ring
    {
        u32 v = 0;
#pragma unroll
        for (u32 j = 0; j < 4; ++j)
            if (off + j < lzss_lane::kPacket) v |= (u32)__ldg(src + off + j) << (8 * j);
        return v;
    }
    __device__ __forceinline__ void store(u32 off, const u32 (&x)[4]) const
    {
        *reinterpret_cast<uint4 *>(dst + off) = make_uint4(x[0], x[1], x[2], x[3]);
    }
};

template <bool PARITY>
__global__ void __launch_bounds__(kLaneWarps * 32, 11) culzss_encode_lane_kernel(const u8 *__restrict__ in, u64 first,
                                                                            u64 npackets,
                                                                            u8 *__restrict__ tmp_out,
                                                                            u16 *__restrict__ pkt_size,
                                                                            u8 *__restrict__ last_group_size,
                                                                            const int *__restrict__ sel, int want)
{
    using namespace lzss_lane;
    if (sel && *sel != want) return;
    extern __shared__ __align__(16) u32 lane_cols[];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u64 pid = first + (u64)blockIdx.x * (kLaneWarps * 32) + threadIdx.x;
    const bool live = pid < npackets;
    const u64 pk = live ? pid : npackets - 1;         // idle lanes keep valid addresses
    LaneDevIO io{in + pk * kPacket, tmp_out + pk * (u64)kSlotBytes};
    if (live) prefetch_l2(io.src + 128);
    u32 last_group = 0;
    constexpr u32 kCol = PARITY ? kColumnWordsParity : kColumnWordsFast;
    const u32 size = encode_packet<32, PARITY>(lane_cols + warp * (kCol * 32) + lane, live, io, last_group);
    if (live) {
        pkt_size[pid] = (u16)size;
        last_group_size[pid] = (u8)last_group;
    }
}

// AUTO mode: the first `count` packets were coded by the CTA kernel (a probe whose output is kept);
// the packet-per-lane kernel only evaluates the positions the greedy selection visits, so it wins
// where matches are frequent and loses on nearly incompressible data, where every position is
// visited and its serial scan costs more than the CTA kernel's bit-parallel one (measured: quant
// codes 38.9 vs 17.2 GB/s, order-0 bytes of 6 bits/byte 11.3 vs 27.6).  Pick by the probe's ratio.
constexpr int kSelCta = 1, kSelLane = 2;
__global__ void __launch_bounds__(256) culzss_select_kernel(const u16 *__restrict__ pkt_size, u32 count,
                                                            int *__restrict__ sel)
{
    __shared__ u32 part[8];
    u32 sum = 0;
    for (u32 i = threadIdx.x; i < count; i += 256) sum += pkt_size[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 tot = 0;
        for (int i = 0; i < 8; ++i) tot += part[i];
        *sel = (u64)tot * 2 <= (u64)count * kPacket ? kSelLane : kSelCta;      // ratio >= 2
    }
}

// Per buffer: packet offsets (exclusive scan of the packet sizes), the trailer
// (gpu_compress.cu:624-657) and the "compression took more" decision, exactly like aftercomp's
// `if (j > finish)` test (:494-498): the test runs before every token, so it fires iff the
// output size before the final flush exceeds buf_length.
__global__ void __launch_bounds__(256) culzss_scan_kernel(const u16 *__restrict__ pkt_size,
                                                          const u8 *__restrict__ last_group_size,
                                                          u32 npk, u32 buf_length,
                                                          u8 *__restrict__ out, u64 out_stride,
                                                          u32 *__restrict__ pkt_off,
                                                          u32 *__restrict__ comp_len)
{
    __shared__ u32 sums[256];
    __shared__ u32 carry_s;
    const u32 b = blockIdx.x, tx = threadIdx.x;
    const u16 *sizes = pkt_size + (u64)b * npk;
    u8 *dst = out + (u64)b * out_stride;

    u32 local = 0;
    for (u32 i = tx; i < npk; i += 256) local += sizes[i];
    sums[tx] = local;
    __syncthreads();
    for (u32 s = 128; s > 0; s >>= 1) {
        if (tx < s) sums[tx] += sums[tx + s];
        __syncthreads();
    }
    const u32 total = sums[0];
    __syncthreads();
    const bool took_more = total - last_group_size[(u64)b * npk + npk - 1] > buf_length;
    const u32 clen = total + 2 * npk + 6;
    // clen == buf_length would be read back as a stored (raw) buffer by every decoder of the
    // container, the reference's included (culzss.c:241-242, deculzss.c:94-95): store it raw
    if (took_more || clen > out_stride || clen == buf_length) {
        if (tx == 0) comp_len[b] = 0;   // caller stores the buffer raw (culzss.c:177-183)
        return;
    }
    if (tx == 0) carry_s = 0;
    __syncthreads();
    for (u32 base = 0; base < npk; base += 256) {
        const u32 i = base + tx;
        const u32 sz = i < npk ? sizes[i] : 0;
        sums[tx] = sz;
        __syncthreads();
        for (u32 d = 1; d < 256; d <<= 1) {
            const u32 v = tx >= d ? sums[tx - d] : 0;
            __syncthreads();
            sums[tx] += v;
            __syncthreads();
        }
        const u32 carry = carry_s;
        __syncthreads();
        if (i < npk) {
            pkt_off[(u64)b * npk + i] = carry + sums[tx] - sz;
            dst[total + 2 * i] = (u8)(sz >> 8);      // trailer: packet size, big endian
            dst[total + 2 * i + 1] = (u8)sz;
        }
        if (tx == 255) carry_s = carry + sums[255];
        __syncthreads();
    }
    if (tx == 0) {
        u8 *t = dst + total + 2 * npk;
        t[0] = (u8)(buf_length >> 24);
        t[1] = (u8)(buf_length >> 16);
        t[2] = (u8)(buf_length >> 8);
        t[3] = (u8)buf_length;
        t[4] = 0;   // pad size (always 0 in the reference, gpu_compress.cu:646-654)
        t[5] = 0;
        comp_len[b] = clen;
    }
}

// One warp per packet: move it from its 16-byte aligned slot to its final, byte-aligned place.
// Destination words are 4-byte aligned; each is assembled from two aligned source words.
__global__ void __launch_bounds__(256) culzss_gather_kernel(const u8 *__restrict__ tmp_out,
                                                            const u16 *__restrict__ pkt_size,
                                                            const u32 *__restrict__ pkt_off,
                                                            const u32 *__restrict__ comp_len,
                                                            u32 npk, u64 npackets,
                                                            u8 *__restrict__ out, u64 out_stride)
{
    const u32 lane = threadIdx.x & 31;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 pid = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pid < npackets; pid += warps) {
        const u32 b = (u32)(pid / npk);
        if (comp_len[b] == 0) continue;
        const u8 *s = tmp_out + pid * (u64)kMaxPacketOut;
        u8 *d = out + (u64)b * out_stride + pkt_off[pid];
        const u32 n = pkt_size[pid];
        const u32 head = min(n, (4u - (u32)(reinterpret_cast<uintptr_t>(d) & 3)) & 3u);
        if (lane < head) d[lane] = s[lane];
        const u32 nw = (n - head) >> 2;
        const u32 *sw = reinterpret_cast<const u32 *>(s);
        u32 *dw = reinterpret_cast<u32 *>(d + head);
        const u32 sh = 8 * head;            // source byte offset of destination word 0 is `head`
        for (u32 k = lane; k < nw; k += 32) {
            const u32 lo = sw[k], hi = sw[k + 1];   // slot is padded: k + 1 stays inside it
            dw[k] = head ? __funnelshift_r(lo, hi, sh) : lo;
        }
        const u32 tail0 = head + 4 * nw;
        if (lane < n - tail0) d[tail0 + lane] = s[tail0 + lane];
    }
}

// ====================================================================================== decode
// Trailer parse (gpu_decompress.cu:258-294): one CTA per buffer -> per-packet start / size.
__global__ void __launch_bounds__(256) culzss_parse_kernel(const u8 *__restrict__ comp,
                                                           const u64 *__restrict__ comp_off,
                                                           u32 max_pk, u32 buf_length,
                                                           u32 *__restrict__ pk_start,
                                                           u32 *__restrict__ pk_size,
                                                           u32 *__restrict__ buf_npk)
{
    __shared__ u32 sums[256];
    __shared__ u32 carry_s;
    const u32 b = blockIdx.x, tx = threadIdx.x;
    const u8 *buf = comp + comp_off[b];
    const u64 len64 = comp_off[b + 1] - comp_off[b];
    const u32 len = (u32)len64;
    u32 npk = 0;
    bool raw = len == buf_length;     // stored uncompressed (culzss.c:241-242, deculzss.c:94-95)
    if (!raw && len >= 6) {
        const u32 orig = ((u32)buf[len - 6] << 24) | ((u32)buf[len - 5] << 16) |
                         ((u32)buf[len - 4] << 8) | (u32)buf[len - 3];
        if (orig % kPacket == 0 && orig / kPacket <= max_pk && 6 + 2 * (orig / kPacket) <= len)
            npk = orig / kPacket;
    }
    if (tx == 0) {
        buf_npk[b] = raw ? 0xffffffffu : npk;
        carry_s = 0;
    }
    __syncthreads();
    const u32 payload = len - 2 * npk - 6;
    for (u32 base = 0; base < npk; base += 256) {
        const u32 i = base + tx;
        u32 sz = 0;
        if (i < npk) sz = ((u32)buf[len - 2 * npk + 2 * i - 6] << 8) | (u32)buf[len - 2 * npk + 2 * i - 5];
        sums[tx] = sz;
        __syncthreads();
        for (u32 d = 1; d < 256; d <<= 1) {
            const u32 v = tx >= d ? sums[tx - d] : 0;
            __syncthreads();
            sums[tx] += v;
            __syncthreads();
        }
        const u32 carry = carry_s;
        __syncthreads();
        if (i < npk) {
            u32 st = carry + sums[tx] - sz;
            if (st > payload) { st = payload; sz = 0; }
            if (st + sz > payload) sz = payload - st;
            pk_start[(u64)b * max_pk + i] = st;
            pk_size[(u64)b * max_pk + i] = sz;
        }
        if (tx == 255) carry_s = carry + sums[255];
        __syncthreads();
    }
}

constexpr int kDecWarps = 4;

// Per lane: a 128-byte LZSS window and a 128-byte queue of compressed input, both laid out
// [word][lane] so that lane l only ever touches bank l: whatever offsets the 32 packets of a warp are
// at, a shared-memory access is ONE wavefront (round 1 kept a contiguous row per lane: 3-4
// wavefronts per access and 64 % of the shared-memory pipe).
struct DecSmem {
    u32 win[kDecWarps][32][32];
    u32 inq[kDecWarps][32][32];
};

// byte offset of ring byte o (0..127) inside a lane's [word][lane] column
__device__ __forceinline__ u32 ring_at(u32 o) { return ((o & 124u) << 5) | (o & 3u); }

// One packet per lane (gpu_decompress.cu:164-242 semantics: flag bits LSB first, 1 = literal,
// 0 = (len, off); window slot = output position mod 128, initialised with ' '; a match reads its
// whole source string before it writes).  Round-2 formulation: everything a lane does at a
// data-dependent TIME in round 1 (refilling its input, flushing its output: 1.6 and 2.6 active
// lanes on average) now happens at fixed points of the token loop for all lanes together:
//   * input: 16-byte chunks travel global -> registers -> queue twice per group of 8 tokens, one
//     chunk ahead of their use; tokens are cut from a 64-bit register that is topped up a word at
//     a time;
//   * output: whatever 32-byte sectors are complete leave at the top of every group;
//   * a match that does not overlap its own output is copied four bytes per step.
__global__ void __launch_bounds__(kDecWarps * 32) culzss_decode_kernel(
    const u8 *__restrict__ comp, const u64 *__restrict__ comp_off, u32 nbuf, u32 max_pk,
    u32 buf_length, const u32 *__restrict__ pk_start, const u32 *__restrict__ pk_size,
    const u32 *__restrict__ buf_npk, u8 *__restrict__ out)
{
    __shared__ DecSmem sm;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    u8 *const win = reinterpret_cast<u8 *>(&sm.win[warp][0][lane]);
    u8 *const inq = reinterpret_cast<u8 *>(&sm.inq[warp][0][lane]);
    const u64 total_slots = (u64)nbuf * max_pk;
    for (u64 slot = (u64)blockIdx.x * blockDim.x + tid; slot < total_slots;
         slot += (u64)gridDim.x * blockDim.x) {
        const u32 b = (u32)(slot / max_pk), i = (u32)(slot % max_pk);
        const u32 npk = buf_npk[b];
        if (npk == 0xffffffffu || i >= npk) continue;
        const u8 *src = comp + comp_off[b] + pk_start[slot];
        const u32 size = pk_size[slot];
        u8 *dst = out + (u64)b * buf_length + (u64)i * kPacket;

#pragma unroll
        for (int k = 0; k < 32; ++k) *reinterpret_cast<u32 *>(win + 128 * k) = 0x20202020u;

        // ---- input: `base` coordinates = bytes from the 16-byte aligned address in front of the packet
        const u32 mis = (u32)(reinterpret_cast<uintptr_t>(src) & 15);
        const u8 *base = src - mis;
        const u32 lim = mis + size;          // packet end
        u32 ld = 0;                          // bytes in the queue so far (multiple of 16)
        u32 rp = mis & ~3u;                  // next queue word to cut tokens from
        auto push = [&](const uint4 &v) {    // chunk [ld, ld + 16) into the queue
            u32 *q = reinterpret_cast<u32 *>(inq + ((ld & 112u) << 5));
            q[0] = v.x; q[32] = v.y; q[64] = v.z; q[96] = v.w;
            ld += 16;
        };
#pragma unroll 1
        for (int k = 0; k < 6 && ld < lim; ++k) push(__ldg(reinterpret_cast<const uint4 *>(base + ld)));
        uint4 pf = make_uint4(0, 0, 0, 0);
        bool pf_valid = ld < lim;
        if (pf_valid) pf = __ldg(reinterpret_cast<const uint4 *>(base + ld));
        auto top_up = [&]() {                // the chunk requested last time arrives; request the next
            if (pf_valid && ld - rp <= 112u) {
                push(pf);
                pf_valid = ld < lim;
                if (pf_valid) pf = __ldg(reinterpret_cast<const uint4 *>(base + ld));
            }
        };
        unsigned long long inreg;            // the next `nb` packet bytes, lowest first
        u32 nb;
        {
            const u32 w0 = *reinterpret_cast<const u32 *>(inq + ((rp & 124u) << 5));
            inreg = (unsigned long long)(w0 >> (8 * (mis & 3u)));
            nb = 4 - (mis & 3u);
            rp += 4;
        }
        auto refill = [&]() {                // keeps at least 5 bytes in inreg
            if (nb <= 4) {
                const u32 wq = *reinterpret_cast<const u32 *>(inq + ((rp & 124u) << 5));
                inreg |= (unsigned long long)wq << (8 * nb);
                nb += 4;
                rp += 4;
            }
        };

        u32 fp = 0;            // compressed bytes consumed
        u32 w = 0;             // bytes produced
        u32 flushed = 0;       // bytes already in global memory (multiple of 32)
        // Invariant: w - flushed <= 128 at all times (a window slot is only overwritten after it
        // has left), kept by flushing below 64 pending bytes before any copy of <= 64 bytes.
        auto flush_ready = [&]() {
            while (w - flushed >= 32 && flushed + 32 <= (u32)kPacket) {
                const u32 *r = reinterpret_cast<const u32 *>(win + ((flushed & 96u) << 5));
                uint4 *g4 = reinterpret_cast<uint4 *>(dst + flushed);
                g4[0] = make_uint4(r[0], r[32], r[64], r[96]);
                g4[1] = make_uint4(r[128], r[160], r[192], r[224]);
                flushed += 32;
            }
        };
        bool more = true;
        while (more) {
            top_up();
            flush_ready();
            refill();
            if (fp >= size) break;
            u32 flags = (u32)inreg & 0xffu;
            inreg >>= 8; --nb; ++fp;
#pragma unroll 1
            for (int t = 0; t < 8; ++t, flags >>= 1) {
                if (t == 4) top_up();
                refill();
                if (fp >= size) { more = false; break; }
                const u32 b0 = (u32)inreg & 0xffu, b1 = ((u32)inreg >> 8) & 0xffu;
                if (flags & 1) {
                    inreg >>= 8; --nb; ++fp;
                    if (w < (u32)kPacket) win[ring_at(w & 127u)] = (u8)b0;
                    ++w;
                    continue;
                }
                if (fp + 1 >= size) { more = false; break; }
                inreg >>= 16; nb -= 2; fp += 2;
                const u32 len = b0, off = b1;
                const u32 a = (w - off) & 127u;      // distance from source slot to write slot
                if ((a == 0 || a >= len) && w + len <= (u32)kPacket) {
                    // The source is not overwritten before it is read.  The copy runs over the
                    // DESTINATION words of the window: each is the funnel shift of two source words
                    // merged into the old word under a byte mask (first and last word of the string
                    // only) -- one shared-memory store per four bytes instead of four.  At most 64
                    // bytes between two looks at the flush condition.
                    for (u32 done = 0; done < len;) {
                        const u32 part = min(64u, len - done);
                        if (w - flushed >= 64) flush_ready();
                        const u32 d0 = w & 127u, s0 = (off + done) & 127u;
                        const u32 head = d0 & 3u;                       // bytes of the first word in front of the string
                        const u32 sh = 8 * ((s0 - d0) & 3u);
                        u32 sw = (s0 - head) & 124u;                    // source word that holds the byte for dest byte 0
                        u32 x0 = *reinterpret_cast<const u32 *>(win + (sw << 5));
                        const u32 nw = (head + part + 3) >> 2;
                        for (u32 j = 0; j < nw; ++j) {
                            sw = (sw + 4) & 124u;
                            const u32 x1 = *reinterpret_cast<const u32 *>(win + (sw << 5));
                            u32 x = __funnelshift_r(x0, x1, sh);
                            x0 = x1;
                            u32 *const dp = reinterpret_cast<u32 *>(win + (((d0 + 4 * j) & 124u) << 5));
                            const u32 lo = j == 0 ? head : 0u;
                            const u32 hi = min(4u, head + part - 4 * j);
                            if (lo != 0 || hi != 4) {
                                const u32 m = (0xffffffffu << (8 * lo)) & (0xffffffffu >> (8 * (4 - hi)));
                                x = (*dp & ~m) | (x & m);
                            }
                            *dp = x;
                        }
                        w += part;
                        done += part;
                    }
                } else {
                    // general case exactly as the reference (gpu_decompress.cu:220-236):
                    // read the whole string from the old window, then append it
                    u8 tmp[256];
                    for (u32 k = 0; k < len; ++k) tmp[k] = win[ring_at((off + k) & 127u)];
                    for (u32 k = 0; k < len; ++k) {
                        if (w < (u32)kPacket) win[ring_at(w & 127u)] = tmp[k];
                        ++w;
                        flush_ready();
                    }
                }
            }
        }
        // a well-formed packet ends with w == 4096 and everything flushed; otherwise write the tail
        flush_ready();
        const u32 wend = min(w, (u32)kPacket);
        for (u32 k = flushed; k < wend; ++k) dst[k] = win[ring_at(k & 127u)];
    }
}

// Raw (stored) buffers: plain copy, 16 bytes per thread and step when both sides are aligned.
__global__ void culzss_copy_raw_kernel(const u8 *__restrict__ comp, const u64 *__restrict__ comp_off,
                                       const u32 *__restrict__ buf_npk, u32 buf_length,
                                       u8 *__restrict__ out)
{
    const u32 b = blockIdx.y;
    if (buf_npk[b] != 0xffffffffu) return;
    const u8 *s = comp + comp_off[b];
    u8 *d = out + (u64)b * buf_length;
    const u32 tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
        uint4 *d4 = reinterpret_cast<uint4 *>(d);
        for (u32 k = tid; k < (buf_length >> 4); k += nthr) d4[k] = s4[k];
        for (u32 k = (buf_length & ~15u) + tid; k < buf_length; k += nthr) d[k] = s[k];
    } else {
        for (u32 k = tid; k < buf_length; k += nthr) d[k] = s[k];
    }
}

}  // namespace lzss
}  // namespace b200lc

using namespace b200lc;

// ====================================================================================== C ABI
extern "C" size_t b200lc_culzss_encode_scratch_bytes(size_t nbuf, size_t buf_length)
{
    const size_t npk = nbuf * (buf_length / lzss::kPacket);
    return npk * lzss::kMaxPacketOut + 16 + ((npk * 2 + 255) & ~size_t(255)) +
           ((npk + 255) & ~size_t(255)) + ((npk * 4 + 255) & ~size_t(255)) + 256;
}

// Parity mode switches to the packet-per-lane kernel from this many packets (160 MiB) on: a lane takes
// ~10 ms for a packet of quantisation codes whatever the batch size (one wave = 148 SMs x 22 warps x
// 32 lanes = 104k packets), the CTA kernel 0.24 us per packet -- they meet at ~41k packets.
constexpr size_t kLaneParityMinPackets = 40960;

static int encode_batch(const uint8_t *d_in, size_t nbuf, size_t buf_length, uint8_t *d_out, size_t out_stride,
                        uint32_t *d_comp_len, void *d_scratch, size_t scratch_bytes, int depth, cudaStream_t stream,
                        int parity_kernel = B200LC_CULZSS_KERNEL_AUTO)
{
    if (nbuf == 0) return B200LC_OK;
    if (!d_in || !d_out || !d_comp_len || !d_scratch) return B200LC_ERR_ARG;
    if (buf_length == 0 || buf_length % lzss::kPacket || buf_length > (1u << 30)) return B200LC_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_scratch) & 15))
        return B200LC_ERR_ARG;
    if (scratch_bytes < b200lc_culzss_encode_scratch_bytes(nbuf, buf_length)) return B200LC_ERR_SCRATCH;
    typedef void (*Kern)(const u8 *, u64, u64, u8 *, u16 *, u8 *, const int *, int);
    static const Kern kerns[4] = {lzss::culzss_encode_kernel<0>, lzss::culzss_encode_kernel<1>,
                                  lzss::culzss_encode_kernel<2>, lzss::culzss_encode_kernel<4>};
    int ki;
    switch (depth) {
        case B200LC_CULZSS_FAST_LANE: ki = -1; break;
        case 0: ki = 0; break;
        case 1: ki = 1; break;
        case 2: ki = 2; break;
        case 4: ki = 3; break;
        default: return B200LC_ERR_UNSUPPORTED;
    }
    const u32 npk_buf = (u32)(buf_length / lzss::kPacket);
    const u64 npk = (u64)nbuf * npk_buf;
    u8 *tmp = reinterpret_cast<u8 *>(d_scratch);
    u16 *sizes = reinterpret_cast<u16 *>(tmp + npk * lzss::kMaxPacketOut + 16);
    u8 *lastg = reinterpret_cast<u8 *>(sizes) + ((npk * 2 + 255) & ~u64(255));
    u32 *pkoff = reinterpret_cast<u32 *>(lastg + ((npk + 255) & ~u64(255)));
    int *sel = reinterpret_cast<int *>(reinterpret_cast<u8 *>(pkoff) + ((npk * 4 + 255) & ~u64(255)));   // the 256 spare bytes

    // Parity mode has two bit-identical kernels: a CTA per packet (any batch size) and a packet per
    // lane, which needs tens of thousands of packets in flight to fill the GPU and frequent matches
    // to pay off.  AUTO: small batches -> CTA kernel; large ones -> the CTA kernel codes a probe of
    // 512 packets, culzss_select_kernel looks at its ratio and both kernels are launched on the
    // rest, one of which returns at once (no host round trip: the call stays asynchronous).
    enum { kCta, kLane, kProbe } plan = ki < 0 ? kLane : kCta;
    if (ki == 0) {
        static int mode = -1;      // B200LC_CULZSS_PARITY_LANE = 0 never | 1 always | unset: probe
        if (mode < 0) {
            const char *e = getenv("B200LC_CULZSS_PARITY_LANE");
            mode = e ? (atoi(e) ? 1 : 0) : 2;
        }
        if (mode == 1) plan = kLane;
        if (mode == 2 && npk >= (u64)kLaneParityMinPackets) plan = kProbe;
        if (parity_kernel == B200LC_CULZSS_KERNEL_CTA) plan = kCta;
        if (parity_kernel == B200LC_CULZSS_KERNEL_LANE) plan = kLane;
    }
    const auto launch_lane = [&](u64 first, const int *sel_p) -> int {
        const bool parity = ki == 0;
        const size_t lsmem = (size_t)lzss::kLaneWarps * 32 * 4 *
                             (parity ? lzss_lane::kColumnWordsParity : lzss_lane::kColumnWordsFast);
        auto kern = parity ? lzss::culzss_encode_lane_kernel<true> : lzss::culzss_encode_lane_kernel<false>;
        static unsigned lane_attr_done[kMaxDevices][2] = {{0}};
        const int lslot = device_slot();
        if (lslot < 0 || lane_attr_done[lslot][parity] != context_epoch()) {
            B200LC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsmem));
            if (lslot >= 0) lane_attr_done[lslot][parity] = context_epoch();
        }
        const u32 per_cta = lzss::kLaneWarps * 32;
        kern<<<(u32)((npk - first + per_cta - 1) / per_cta), per_cta, lsmem, stream>>>(d_in, first, npk, tmp, sizes, lastg,
                                                                                     sel_p, lzss::kSelLane);
        B200LC_CUDA_TRY(cudaGetLastError());
        return B200LC_OK;
    };
    const auto launch_cta = [&](u64 first, u64 end, const int *sel_p) -> int {
        const int k = ki < 0 ? 0 : ki;
        // fast mode keeps its hash table (1024 x 16 bytes) behind EncSmem
        const size_t smem = ((sizeof(lzss::EncSmem) + 15) & ~size_t(15)) + (k ? (size_t(16) << 10) : 0);
        static unsigned attr_done[kMaxDevices][4] = {{0}};   // context epoch the attribute was set in
        const int slot = device_slot();
        if (slot < 0 || attr_done[slot][k] != context_epoch()) {
            B200LC_CUDA_TRY(cudaFuncSetAttribute(kerns[k], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (slot >= 0) attr_done[slot][k] = context_epoch();
        }
        const u32 grid = (u32)min(end - first, (u64)num_sms() * 64);
        kerns[k]<<<grid, 128, smem, stream>>>(d_in, first, end, tmp, sizes, lastg, sel_p, lzss::kSelCta);
        B200LC_CUDA_TRY(cudaGetLastError());
        return B200LC_OK;
    };
    int rc = B200LC_OK;
    if (plan == kLane) rc = launch_lane(0, nullptr);
    else if (plan == kCta) rc = launch_cta(0, npk, nullptr);
    else {
        const u64 probe = 512;
        rc = launch_cta(0, probe, nullptr);
        if (rc == B200LC_OK) {
            lzss::culzss_select_kernel<<<1, 256, 0, stream>>>(sizes, (u32)probe, sel);
            B200LC_CUDA_TRY(cudaGetLastError());
            rc = launch_lane(probe, sel);
        }
        if (rc == B200LC_OK) rc = launch_cta(probe, npk, sel);
    }
    if (rc != B200LC_OK) return rc;
    lzss::culzss_scan_kernel<<<(u32)nbuf, 256, 0, stream>>>(sizes, lastg, npk_buf, (u32)buf_length,
                                                           d_out, out_stride, pkoff, d_comp_len);
    B200LC_CUDA_TRY(cudaGetLastError());
    const u32 ggrid = (u32)min((npk + 7) / 8, (u64)num_sms() * 32);
    lzss::culzss_gather_kernel<<<ggrid, 256, 0, stream>>>(tmp, sizes, pkoff, d_comp_len, npk_buf, npk,
                                                         d_out, out_stride);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

extern "C" int b200lc_culzss_encode_batch(const uint8_t *d_in, size_t nbuf, size_t buf_length,
                                          uint8_t *d_out, size_t out_stride, uint32_t *d_comp_len,
                                          void *d_scratch, size_t scratch_bytes, void *stream_)
{
    return encode_batch(d_in, nbuf, buf_length, d_out, out_stride, d_comp_len, d_scratch, scratch_bytes, 0,
                        (cudaStream_t)stream_);
}

extern "C" int b200lc_culzss_encode_batch_ex(const uint8_t *d_in, size_t nbuf, size_t buf_length,
                                             uint8_t *d_out, size_t out_stride, uint32_t *d_comp_len,
                                             void *d_scratch, size_t scratch_bytes, int kernel, void *stream_)
{
    if (kernel < B200LC_CULZSS_KERNEL_AUTO || kernel > B200LC_CULZSS_KERNEL_LANE) return B200LC_ERR_ARG;
    return encode_batch(d_in, nbuf, buf_length, d_out, out_stride, d_comp_len, d_scratch, scratch_bytes, 0,
                        (cudaStream_t)stream_, kernel);
}

extern "C" int b200lc_culzss_encode_fast_batch(const uint8_t *d_in, size_t nbuf, size_t buf_length,
                                               uint8_t *d_out, size_t out_stride, uint32_t *d_comp_len,
                                               void *d_scratch, size_t scratch_bytes, int depth, void *stream_)
{
    if (depth == 0) return B200LC_ERR_ARG;
    return encode_batch(d_in, nbuf, buf_length, d_out, out_stride, d_comp_len, d_scratch, scratch_bytes, depth,
                        (cudaStream_t)stream_);
}

extern "C" size_t b200lc_culzss_decode_scratch_bytes(size_t nbuf, size_t buf_length)
{
    const size_t max_pk = buf_length / lzss::kPacket;
    return nbuf * max_pk * 8 + ((nbuf * 4 + 255) & ~size_t(255)) + 256;
}

extern "C" int b200lc_culzss_decode_batch(const uint8_t *d_comp, const uint64_t *d_comp_offsets,
                                          size_t nbuf, size_t buf_length, uint8_t *d_out,
                                          void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nbuf == 0) return B200LC_OK;
    if (!d_comp || !d_comp_offsets || !d_out || !d_scratch) return B200LC_ERR_ARG;
    if (buf_length == 0 || buf_length % lzss::kPacket || buf_length > (1u << 30)) return B200LC_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_out) & 15) || (reinterpret_cast<uintptr_t>(d_scratch) & 15))
        return B200LC_ERR_ARG;
    if (scratch_bytes < b200lc_culzss_decode_scratch_bytes(nbuf, buf_length)) return B200LC_ERR_SCRATCH;
    const u32 max_pk = (u32)(buf_length / lzss::kPacket);
    u32 *pk_start = reinterpret_cast<u32 *>(d_scratch);
    u32 *pk_size = pk_start + nbuf * max_pk;
    u32 *buf_npk = pk_size + nbuf * max_pk;
    lzss::culzss_parse_kernel<<<(u32)nbuf, 256, 0, stream>>>(d_comp, d_comp_offsets, max_pk,
                                                            (u32)buf_length, pk_start, pk_size, buf_npk);
    B200LC_CUDA_TRY(cudaGetLastError());
    const u64 slots = (u64)nbuf * max_pk;
    const u32 threads = lzss::kDecWarps * 32;
    const u32 grid = (u32)min((slots + threads - 1) / threads, (u64)num_sms() * 16);
    lzss::culzss_decode_kernel<<<grid, threads, 0, stream>>>(d_comp, d_comp_offsets, (u32)nbuf, max_pk,
                                                            (u32)buf_length, pk_start, pk_size,
                                                            buf_npk, d_out);
    B200LC_CUDA_TRY(cudaGetLastError());
    dim3 g2(32, (u32)nbuf);
    lzss::culzss_copy_raw_kernel<<<g2, 256, 0, stream>>>(d_comp, d_comp_offsets, buf_npk,
                                                        (u32)buf_length, d_out);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}
