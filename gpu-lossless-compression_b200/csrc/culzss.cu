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
//     output staging unified in one shared-memory ring per lane and 16-byte global accesses.
#include "common.cuh"
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

__global__ void __launch_bounds__(128) culzss_encode_kernel(const u8 *__restrict__ in, u64 npackets,
                                                            u8 *__restrict__ tmp_out,
                                                            u16 *__restrict__ pkt_size,
                                                            u8 *__restrict__ last_group_size)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    EncSmem &sm = *reinterpret_cast<EncSmem *>(smem_raw);
    const u32 tx = threadIdx.x;
    const u32 lane = tx & 31, warp = tx >> 5;

    for (u64 pid = blockIdx.x; pid < npackets; pid += gridDim.x) {
        // ---------------------------------------------------------------- load packet, reset state
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(in + pid * kPacket);
            uint4 *dst = reinterpret_cast<uint4 *>(sm.pkt + kWindow);
            dst[tx] = src[tx];
            dst[tx + 128] = src[tx + 128];
            sm.pkt[tx] = ' ';
            for (u32 i = tx; i < 256 * kOccStride; i += 128) sm.u.a.occ[i] = 0;
        }
        __syncthreads();
        // window slots 128..255 initially hold ' ' (gpu_compress.cu:208), chunk 0 enters slots 0..127
        if (tx < 4) sm.u.a.occ[0x20 * kOccStride + 4 + tx] = 0xffffffffu;
        __syncthreads();
        atomicOr(&sm.u.a.occ[sm.pkt[kWindow + tx] * kOccStride + (tx >> 5)], 1u << (tx & 31));
        __syncthreads();

        for (u32 c = 0; c < kChunks; ++c) {
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
    if (took_more || clen > out_stride) {
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
constexpr int kRowStride = 144;    // bytes per lane ring row (128 + pad, 16-byte aligned)
constexpr int kInChunk = 64;       // bytes of compressed input staged per lane refill
constexpr int kInStride = 80;

struct DecSmem {
    __align__(16) u8 ring[kDecWarps * 32 * kRowStride];
    __align__(16) u8 inb[kDecWarps * 32 * kInStride];
};

// One packet per lane.  The lane's 128-byte ring row is both the LZSS window
// (slot = output position mod 128, gpu_decompress.cu:164-242) and the staging of the output,
// flushed to global memory 64 bytes at a time.
__global__ void __launch_bounds__(kDecWarps * 32) culzss_decode_kernel(
    const u8 *__restrict__ comp, const u64 *__restrict__ comp_off, u32 nbuf, u32 max_pk,
    u32 buf_length, const u32 *__restrict__ pk_start, const u32 *__restrict__ pk_size,
    const u32 *__restrict__ buf_npk, u8 *__restrict__ out)
{
    __shared__ DecSmem sm;
    const u32 tid = threadIdx.x;
    u8 *row = sm.ring + tid * kRowStride;
    u8 *inb = sm.inb + tid * kInStride;
    const u64 total_slots = (u64)nbuf * max_pk;
    for (u64 slot = (u64)blockIdx.x * blockDim.x + tid; slot < total_slots;
         slot += (u64)gridDim.x * blockDim.x) {
        const u32 b = (u32)(slot / max_pk), i = (u32)(slot % max_pk);
        const u32 npk = buf_npk[b];
        if (npk == 0xffffffffu || i >= npk) continue;
        const u8 *src = comp + comp_off[b] + pk_start[slot];
        const u32 size = pk_size[slot];
        u8 *dst = out + (u64)b * buf_length + (u64)i * kPacket;

        for (int k = 0; k < 128; k += 16)
            *reinterpret_cast<uint4 *>(row + k) =
                make_uint4(0x20202020u, 0x20202020u, 0x20202020u, 0x20202020u);
        // compressed input is staged 64 bytes at a time through 16-byte aligned vector loads
        const u32 mis = (u32)(reinterpret_cast<uintptr_t>(src) & 15);
        const u8 *base = src - mis;
        const u32 lim = mis + size;          // packet end in `base` coordinates
        u32 loaded_lo = 0, loaded_hi = 0;    // staged range: base coords [lo, lo+64), packet coords < hi
        auto get = [&](u32 pos) -> u32 {     // byte `pos` of the packet, pos non-decreasing
            if (pos >= loaded_hi) {
                const u32 a = (pos + mis) & ~63u;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (a + 16 * q < lim)
                        *reinterpret_cast<uint4 *>(inb + 16 * q) =
                            *reinterpret_cast<const uint4 *>(base + a + 16 * q);
                loaded_lo = a;
                loaded_hi = a + kInChunk - mis;
            }
            return inb[pos + mis - loaded_lo];
        };
        u32 fp = 0;            // compressed bytes consumed
        u32 w = 0;             // bytes produced
        u32 flushed = 0;       // bytes already written to global memory (multiple of 64)
        // Invariant at every token boundary: w - flushed < 64, so a token of <= 64 bytes never
        // overwrites ring bytes that are not in global memory yet.
        auto flush_ready = [&]() {
            while (w - flushed >= 64 && flushed + 64 <= kPacket) {
                const uint4 *r4 = reinterpret_cast<const uint4 *>(row + (flushed & 127));
                uint4 *g4 = reinterpret_cast<uint4 *>(dst + flushed);
#pragma unroll
                for (int q = 0; q < 4; ++q) g4[q] = r4[q];
                flushed += 64;
            }
        };
        u32 flags = 0, flags_used = 7;
        while (true) {
            flags >>= 1;
            if (++flags_used == 8) {
                if (fp >= size) break;
                flags = get(fp++);
                flags_used = 0;
            }
            if (flags & 1) {
                if (fp >= size) break;
                const u32 cbyte = get(fp++);
                if (w < kPacket) row[w & 127] = (u8)cbyte;
                ++w;
            } else {
                if (fp >= size) break;
                const u32 len = get(fp++);
                if (fp >= size) break;
                const u32 off = get(fp++);
                const u32 a = (w - off) & 127;      // distance from source slot to write slot
                if ((a == 0 || a >= len) && len <= 64 && w + len <= kPacket) {
                    // source slots are not overwritten before they are read: forward copy, four
                    // bytes per step (two aligned ring words funnel-shifted to the source offset;
                    // a step only overwrites slots whose source bytes earlier steps have consumed)
                    const u32 *row32 = reinterpret_cast<const u32 *>(row);
                    for (u32 k = 0; k < len; k += 4) {
                        const u32 s = (off + k) & 127, i0 = s >> 2;
                        const u32 x = __funnelshift_r(row32[i0], row32[(i0 + 1) & 31], 8 * (s & 3));
                        const u32 d = w + k;
                        row[d & 127] = (u8)x;
                        if (k + 1 < len) row[(d + 1) & 127] = (u8)(x >> 8);
                        if (k + 2 < len) row[(d + 2) & 127] = (u8)(x >> 16);
                        if (k + 3 < len) row[(d + 3) & 127] = (u8)(x >> 24);
                    }
                    w += len;
                } else {
                    // general case exactly as the reference (gpu_decompress.cu:220-236):
                    // read the whole string from the old window, then append it
                    u8 tmp[256];
                    for (u32 k = 0; k < len; ++k) tmp[k] = row[(off + k) & 127];
                    for (u32 k = 0; k < len; ++k) {
                        if (w < kPacket) row[w & 127] = tmp[k];
                        ++w;
                        flush_ready();
                    }
                }
            }
            flush_ready();
        }
        // a well-formed packet ends with w == 4096 and everything flushed; otherwise write the tail
        const u32 wend = min(w, (u32)kPacket);
        for (u32 k = flushed; k < wend; ++k) dst[k] = row[k & 127];
    }
}

// Raw (stored) buffers: plain copy.
__global__ void culzss_copy_raw_kernel(const u8 *__restrict__ comp, const u64 *__restrict__ comp_off,
                                       const u32 *__restrict__ buf_npk, u32 buf_length,
                                       u8 *__restrict__ out)
{
    const u32 b = blockIdx.y;
    if (buf_npk[b] != 0xffffffffu) return;
    const u8 *s = comp + comp_off[b];
    u8 *d = out + (u64)b * buf_length;
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < buf_length; k += gridDim.x * blockDim.x)
        d[k] = s[k];
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

extern "C" int b200lc_culzss_encode_batch(const uint8_t *d_in, size_t nbuf, size_t buf_length,
                                          uint8_t *d_out, size_t out_stride, uint32_t *d_comp_len,
                                          void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nbuf == 0) return B200LC_OK;
    if (!d_in || !d_out || !d_comp_len || !d_scratch) return B200LC_ERR_ARG;
    if (buf_length == 0 || buf_length % lzss::kPacket || buf_length > (1u << 30)) return B200LC_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) || (reinterpret_cast<uintptr_t>(d_scratch) & 15))
        return B200LC_ERR_ARG;
    if (scratch_bytes < b200lc_culzss_encode_scratch_bytes(nbuf, buf_length)) return B200LC_ERR_SCRATCH;
    const u32 npk_buf = (u32)(buf_length / lzss::kPacket);
    const u64 npk = (u64)nbuf * npk_buf;
    u8 *tmp = reinterpret_cast<u8 *>(d_scratch);
    u16 *sizes = reinterpret_cast<u16 *>(tmp + npk * lzss::kMaxPacketOut + 16);
    u8 *lastg = reinterpret_cast<u8 *>(sizes) + ((npk * 2 + 255) & ~u64(255));
    u32 *pkoff = reinterpret_cast<u32 *>(lastg + ((npk + 255) & ~u64(255)));

    static unsigned attr_done[kMaxDevices] = {0};   // context epoch the attribute was set in
    const int slot = device_slot();
    if (slot < 0 || attr_done[slot] != context_epoch()) {
        B200LC_CUDA_TRY(cudaFuncSetAttribute(lzss::culzss_encode_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(lzss::EncSmem)));
        if (slot >= 0) attr_done[slot] = context_epoch();
    }
    const u32 grid = (u32)min(npk, (u64)num_sms() * 64);
    lzss::culzss_encode_kernel<<<grid, 128, sizeof(lzss::EncSmem), stream>>>(d_in, npk, tmp, sizes, lastg);
    B200LC_CUDA_TRY(cudaGetLastError());
    lzss::culzss_scan_kernel<<<(u32)nbuf, 256, 0, stream>>>(sizes, lastg, npk_buf, (u32)buf_length,
                                                           d_out, out_stride, pkoff, d_comp_len);
    B200LC_CUDA_TRY(cudaGetLastError());
    const u32 ggrid = (u32)min((npk + 7) / 8, (u64)num_sms() * 32);
    lzss::culzss_gather_kernel<<<ggrid, 256, 0, stream>>>(tmp, sizes, pkoff, d_comp_len, npk_buf, npk,
                                                         d_out, out_stride);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
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
