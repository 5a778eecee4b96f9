// CULZSS encoders with ONE PACKET PER LANE: match finding, greedy token selection, token emission
// and flag-byte packing in one serial walk over the packet.  Two match finders share the walk:
//
//   PARITY = true   bit-exact with the reference: FindMatch's streak scanner (gpu_compress.cu:104-168,
//                   SURVEY.md appendix A.2 "LCP walk") evaluated ONLY at the positions the greedy
//                   selection of aftercomp (:500-517) visits -- the reference computes a match for all
//                   4096 positions of a packet and then throws away those a longer match jumped over;
//   PARITY = false  FAST MODE, NON-PARITY: the reference's buffer / token format, window 128, packets
//                   of 4096 bytes -- every stream decodes with the reference's DecodeKernel,
//                   gpu_decompress.cu:164-242 -- but matches from a lane-private hash of three-byte
//                   prefixes instead of the reference's exhaustive scanner.
//
// The code a lane runs lives in this header, which also compiles for the host, so that the CPU test
// (tests/c/culzss_lane_host.cc, tests/test_culzss_lane_cpu.py) can run it without a GPU and hand its
// output to the oracle's decoder.
//
// Per lane (STRIDE = 32 words on the GPU: word w of a lane's column sits at column[w * 32], so a
// lane only ever touches its own shared-memory bank; STRIDE = 1 on the host):
//   ring   256 bytes   input bytes of positions [hi - 256, hi), index = position & 255; holds the
//                      128-byte window behind p and 45..128 bytes of lookahead (32-byte refills);
//                      a 65th word mirrors word 0 so that an unaligned 4-byte read never wraps
//   hash   2^kHashBits bytes   low 8 bits of the latest token start whose three bytes hash here
//   outq   32 bytes    compressed bytes [flushed, o), index = offset & 31; leaves 16 bytes at a time
//
// Everything a lane does at a data-dependent TIME is moved to warp-uniform points (encode_packet):
// every live lane codes exactly one token per step, so the eight tokens of a group, the flag byte
// and the output flush sit at the same place of the loop for all 32 packets; input is topped up for
// ALL lanes that have room as soon as ANY lane runs low, and a match is simply cut at the lookahead
// a lane has (a lane that just coded a long match asks for input at once, so that costs nothing on
// compressible data).
//
// The hash table needs no valid bits and no clearing between packets beyond the initial fill: a
// candidate is (p - entry) & 255 bytes back, it is used only if that distance is 3..128, and the
// bytes there are compared with the lookahead -- whatever position the entry once stood for, equal
// bytes inside the window are a legal match.
//
// Token format (gpu_compress.cu:500-517, gpu_decompress.cu:164-242): groups of 8 tokens behind a
// flag byte, bit k (LSB first) = 1: one literal byte; 0: {length, window slot of the source}.
// The decoder reads the whole source string before it writes, so a match is kept from reaching its
// own output (length <= distance).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define B200LC_LANE_HD __host__ __device__ __forceinline__
#else
#define B200LC_LANE_HD inline
#endif

namespace b200lc {
namespace lzss_lane {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;

constexpr u32 kPacket = 4096;
constexpr u32 kWindow = 128;
constexpr u32 kRingWords = 64;
constexpr u32 kHashBits = 6;
constexpr u32 kHashWords = (1u << kHashBits) / 4;
constexpr u32 kOutWords = 8;       // a group is <= 17 bytes and everything but < 16 bytes leaves after each group
constexpr u32 kColumnWordsFast = kRingWords + 1 + kHashWords + kOutWords;   // 89 words = 356 bytes per lane
constexpr u32 kColumnWordsParity = kRingWords + 1 + kOutWords;              // 73 words = 292 bytes per lane
constexpr u32 kMaxLen = 124;      // <= lookahead - 3 <= 125
constexpr u32 kSlotBytes = kPacket + kPacket / 8;                     // output slot of a packet

B200LC_LANE_HD u32 fsr(u32 lo, u32 hi, u32 s)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
B200LC_LANE_HD u32 ffs32(u32 x)
{
#if defined(__CUDA_ARCH__)
    return (u32)__ffs((int)x);
#else
    return (u32)__builtin_ffs((int)x);
#endif
}
// number of low bytes of d that are zero (0..4): d = xor of two little-endian 4-byte strings
B200LC_LANE_HD u32 same_bytes(u32 d)
{
#if defined(__CUDA_ARCH__)
    return (u32)__clz((int)__brev(d)) >> 3;
#else
    return d ? ((u32)__builtin_ctz(d) >> 3) : 4u;
#endif
}
B200LC_LANE_HD u32 min_u(u32 a, u32 b) { return a < b ? a : b; }
B200LC_LANE_HD u32 max_u(u32 a, u32 b) { return a > b ? a : b; }
// 0x80 in every byte of w that equals the corresponding byte of b (exact, no carries between bytes)
B200LC_LANE_HD u32 eq_bytes(u32 w, u32 b)
{
    const u32 x = w ^ b;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}

struct Chunk32 { u32 w[8]; };     // 32 bytes of input on their way from global memory to the ring

template <int STRIDE, bool PARITY>
struct Lane {
    u32 *ring, *hash, *outq;    // this lane's columns
    u32 p;                      // next position to code
    u32 hi;                     // input in the ring so far (multiple of 32)
    u32 o;                      // compressed bytes produced so far
    u32 flushed;                // compressed bytes stored to the slot (multiple of 16)
    u32 fpos;                   // offset of the open group's flag byte
    u32 flags;                  // flag bits of the open group
    u32 need;                   // lookahead below which this lane asks for input

    B200LC_LANE_HD u8 *byte_of(u32 *col, u32 i) const
    {
        return reinterpret_cast<u8 *>(col + (i >> 2) * STRIDE) + (i & 3u);
    }
    // four bytes at ring index i in 0..255 (two aligned words, funnel-shifted; word 64 mirrors word 0)
    B200LC_LANE_HD u32 ring4(u32 i) const
    {
        const u32 *w = ring + (i >> 2) * STRIDE;
        return fsr(w[0], w[STRIDE], 8 * i);
    }

    B200LC_LANE_HD void init(u32 *column, bool live)
    {
        ring = column;
        hash = column + (kRingWords + 1) * STRIDE;
        outq = hash + (PARITY ? 0u : kHashWords) * STRIDE;
        // positions -128 .. -1 are spaces (gpu_compress.cu:208); every hash entry points at -128
        for (u32 w = kRingWords / 2; w < kRingWords; ++w) ring[w * STRIDE] = 0x20202020u;
        ring[0] = ring[kRingWords * STRIDE] = 0;
        if (!PARITY)
            for (u32 w = 0; w < kHashWords; ++w) hash[w * STRIDE] = 0x80808080u;
        p = hi = live ? 0u : kPacket;
        o = 0; flushed = 0; fpos = 0; flags = 0; need = 48;
    }

    // input: 32 bytes at a time into ring slots that hold positions behind the window
    B200LC_LANE_HD bool has_room() const { return hi < kPacket && hi <= p + 96u; }
    B200LC_LANE_HD bool hungry() const { return has_room() && hi < p + need; }   // parity mode: p may have passed hi
    B200LC_LANE_HD void put_input(const Chunk32 &c)
    {
        u32 *q = ring + ((hi >> 2) & (kRingWords - 1)) * STRIDE;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 8; ++j) q[j * STRIDE] = c.w[j];
        if ((hi & 255u) == 0) ring[kRingWords * STRIDE] = c.w[0];      // the mirror of word 0
        hi += 32;
    }

    // output: whole 16-byte vectors of [flushed, o)
    B200LC_LANE_HD void take_output(u32 (&x)[4])
    {
        const u32 *q = outq + ((flushed >> 2) & (kOutWords - 1)) * STRIDE;
        x[0] = q[0]; x[1] = q[STRIDE]; x[2] = q[2 * STRIDE]; x[3] = q[3 * STRIDE];
        flushed += 16;
    }

    B200LC_LANE_HD void emit(u32 b) { *byte_of(outq, o & (4 * kOutWords - 1)) = (u8)b; ++o; }
    B200LC_LANE_HD void open_group() { fpos = o; ++o; flags = 0; }
    B200LC_LANE_HD void close_group() { *byte_of(outq, fpos & (4 * kOutWords - 1)) = (u8)flags; }

    // Token number t (0..7) of the open group.  Requires p < kPacket and >= 48 bytes of lookahead
    // (or all of the packet's input in the ring).
    B200LC_LANE_HD void step(u32 t)
    {
        const u32 x = ring4(p & 255u);
        const u32 h = ((x & 0xffffffu) * 2654435761u) >> (32 - kHashBits);
        u8 *const he = byte_of(hash, h);
        const u32 dist = (p - (u32)*he) & 255u;
        // bytes [p, p + lim + 3) must be in the ring (beyond the packet's end garbage is compared and
        // cut off by kPacket - p)
        const u32 avail = hi < kPacket ? hi - p - 3u : kPacket - p;
        const u32 lim = min_u(min_u(dist, kMaxLen), avail);
        u32 L = 0;
        if (dist - 3u <= kWindow - 3u && lim >= 3u) {
            const u32 q = (p - dist) & 255u;
            // the first eight bytes of both strings at once (three aligned words per side): most
            // matches end there and never see the loop
            const u32 *sw = ring + (q >> 2) * STRIDE;
            const u32 s0 = sw[0], s1 = sw[STRIDE];
            const u32 d0 = fsr(s0, s1, 8 * q) ^ x;
            if ((d0 & 0xffffffu) == 0) {
                u32 ws = (q >> 2) + 2, wp = ((p & 255u) >> 2) + 1;
                u32 s2 = ring[(ws & (kRingWords - 1)) * STRIDE];
                const u32 p1a = ring[(wp & (kRingWords - 1)) * STRIDE];
                u32 p2 = ring[((wp + 1) & (kRingWords - 1)) * STRIDE];
                const u32 d1 = fsr(s1, s2, 8 * q) ^ fsr(p1a, p2, 8 * p);
                const u32 l0 = same_bytes(d0);
                L = l0 == 4u ? 4u + same_bytes(d1) : l0;
                if (L == 8u && L < lim) {
                    // longer: a word at a time, one new aligned word per side and step
                    u32 d = 0;
                    ++wp;
                    do {
                        ++ws; ++wp;
                        const u32 a0 = s2, b0 = p2;
                        s2 = ring[(ws & (kRingWords - 1)) * STRIDE];
                        p2 = ring[(wp & (kRingWords - 1)) * STRIDE];
                        d = fsr(a0, s2, 8 * q) ^ fsr(b0, p2, 8 * p);
                        L += 4;
                    } while (d == 0 && L < lim);
                    L -= 4u - same_bytes(d);
                }
                L = min_u(L, lim);
            }
        }
        // one emission path for both token kinds (the branches above are the divergent part)
        const bool is_match = L >= 3u;
        emit(is_match ? L : (x & 0xffu));
        if (is_match) emit((p - dist) & 127u);
        else flags |= 1u << t;
        // Hash entry := p, except:  a source that runs right up to p (L == dist) means the data
        // repeats with period `dist` -- keeping the entry lets the next token reach twice as far back
        // and copy twice as much;  an entry closer than 3 becomes usable in a moment.
        const bool keep = is_match ? bool((L == dist) & (2u * dist <= kWindow)) : (dist < 3u);
        *he = (u8)(keep ? p - dist : p);      // p - dist is the old entry (mod 256): one store, no branch
        p += is_match ? L : 1u;
        need = L >= 32u ? 97u : 48u;      // after a long match: top up at once
    }

    // ------------------------------------------------------------------------------ parity mode
    // Token number t (0..7) of the open group, the reference's match:
    // scan index q = 0 .. n-1 stands for window position p - 128 + q (ring index (p + 128 + q) & 255);
    // walk q <- q + LCP(q) + 1 where LCP(q) is the common prefix of the window string at q and the
    // lookahead, cut so that q + LCP <= n; the first strictly longest wins; n = 127, shrinking with
    // the position inside the packet's last chunk (maxcheck, gpu_compress.cu:120,149).
    // Here: u = q + a counts ring bytes from the aligned word that holds q = 0.  Phase 1 (all lanes in
    // step, 33 independent loads): the bit mask of the window bytes that equal the first lookahead
    // byte -- only those can start a streak.  Phase 2: the walk over the set bits, skipping
    // LCP + 1 bytes after each.
    template <class IO>
    B200LC_LANE_HD void step_parity(u32 t, const IO &io)
    {
        const u32 n = p < kPacket - 128u ? 127u : max_u(1u, kPacket - 1u - p);
        const u32 x = ring4(p & 255u), x2 = ring4((p + 4u) & 255u);     // the first eight lookahead bytes
        const u32 b4 = (x & 0xffu) * 0x01010101u;
        const u32 i0 = (p + 128u) & 255u;
        const u32 a = i0 & 3u, wbase = i0 >> 2;
        const u32 end = n + a;                      // <= 130
        u32 M[5] = {0u, 0u, 0u, 0u, 0u};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (u32 k = 0; k < 33; ++k) {
            const u32 W = ring[((wbase + k) & (kRingWords - 1)) * STRIDE];
            // 0x80 flags of the four bytes -> four adjacent bits
            const u32 nib = (((eq_bytes(W, b4) >> 7) * 0x00204081u) >> 21) & 0xfu;
            M[k >> 3] |= nib << (4 * (k & 7));
        }
        M[0] &= 0xffffffffu << a;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (u32 w = 0; w < 5; ++w) {               // keep u < end
            const int top = (int)end - (int)(32 * w);
            if (top <= 0) M[w] = 0;
            else if (top < 32) M[w] &= (1u << top) - 1u;
        }
        u32 best = 1, best_u = a;
        u32 next = 0;                               // candidates below `next` were swallowed by a streak
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (u32 w = 0; w < 5; ++w) {
            u32 m = M[w];
            if (next > 32 * w) m = next - 32 * w >= 32u ? 0u : m & (0xffffffffu << (next - 32 * w));
            while (m) {
                const u32 bit = ffs32(m) - 1u;
                const u32 u = 32 * w + bit;
                const u32 cap = end - u;
                const u32 q = (4 * wbase + u) & 255u;             // ring index of the streak's first byte
                // eight bytes of the window string at once (three aligned words): most streaks end
                // there and never see the loops
                const u32 *sw = ring + (q >> 2) * STRIDE;
                const u32 s0 = sw[0], s1 = sw[STRIDE];
                u32 ws = (q >> 2) + 2;
                u32 s2 = ring[(ws & (kRingWords - 1)) * STRIDE];
                const u32 l0 = same_bytes(fsr(s0, s1, 8 * q) ^ x);
                u32 L = l0 == 4u ? 4u + same_bytes(fsr(s1, s2, 8 * q) ^ x2) : l0;
                if (L == 8u && L < cap) {
                    // Longer: a word at a time, one new aligned word per side and step while the
                    // lookahead is in the ring (capr), from global memory beyond it.
                    const u32 capr = hi >= kPacket ? cap : min_u(cap, hi - p - 3u);
                    u32 wp = ((p & 255u) >> 2) + 2;
                    u32 p1 = ring[(wp & (kRingWords - 1)) * STRIDE];
                    u32 d = 0;
                    while (d == 0 && L < capr) {
                        ++ws; ++wp;
                        const u32 a0 = s2, b0 = p1;
                        s2 = ring[(ws & (kRingWords - 1)) * STRIDE];
                        p1 = ring[(wp & (kRingWords - 1)) * STRIDE];
                        d = fsr(a0, s2, 8 * q) ^ fsr(b0, p1, 8 * p);
                        L += 4;
                    }
                    while (d == 0 && L < cap) {
                        ++ws;
                        const u32 a0 = s2;
                        s2 = ring[(ws & (kRingWords - 1)) * STRIDE];
                        d = fsr(a0, s2, 8 * q) ^ io.bytes4(p + L);
                        L += 4;
                    }
                    if (d) L -= 4u - same_bytes(d);
                }
                L = min_u(L, cap);
                if (L > best) { best = L; best_u = u; }
                next = u + L + 1;
                m = bit + L + 1 >= 32u ? 0u : m & (0xffffffffu << (bit + L + 1));
            }
        }
        // gpu_compress.cu:251-274: a match of 3 or more becomes {length, ring offset of its source}
        const bool is_match = best >= 3u;
        emit(is_match ? best : (x & 0xffu));
        if (is_match) emit((p + best_u - a) & 255u);
        else flags |= 1u << t;
        p += is_match ? best : 1u;
        need = best >= 32u ? 97u : 48u;
    }
};

// The packet loop.  IO supplies the things that differ between a GPU lane and the host:
//   bool any(bool)                         warp vote (host: identity)
//   void load(u32 offset, Chunk32 &)       32 input bytes at `offset` of the packet
//   u32 bytes4(u32 offset)                 4 input bytes at any offset (zero beyond the packet)
//   void store(u32 offset, const u32 (&)[4])   16 output bytes at `offset` of the packet's slot
// Returns the compressed size; last_group = bytes of the last group incl. its flag byte.
template <int STRIDE, bool PARITY, class IO>
B200LC_LANE_HD u32 encode_packet(u32 *column, bool live, IO &io, u32 &last_group)
{
    Lane<STRIDE, PARITY> ln;
    ln.init(column, live);
    Chunk32 pf;
    for (int j = 0; j < 8; ++j) pf.w[j] = 0;
    if (live) io.load(0, pf);
    for (;;) {
        const bool open = ln.p < kPacket;
        if (!io.any(open)) break;
        if (open) ln.open_group();
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (u32 t = 0; t < 8; ++t) {
            if (io.any(ln.hungry())) {
                while (ln.has_room()) {
                    ln.put_input(pf);
                    if (ln.hi < kPacket) io.load(ln.hi, pf);
                }
            }
            if (ln.p < kPacket) {
                if (PARITY) ln.step_parity(t, io);
                else ln.step(t);
            }
        }
        if (open) ln.close_group();
        // a group is at most 17 bytes: two vectors at most
        for (int k = 0; k < 2; ++k) {
            if (ln.o - ln.flushed >= 16u) {
                u32 x[4];
                const u32 at = ln.flushed;
                ln.take_output(x);
                io.store(at, x);
            }
        }
    }
    if (ln.flushed < ln.o) {      // the tail (< 16 bytes; the rest of the vector is padding in the slot)
        u32 x[4];
        const u32 at = ln.flushed;
        ln.take_output(x);
        io.store(at, x);
    }
    last_group = ln.o - ln.fpos;
    return ln.o;
}

}  // namespace lzss_lane
}  // namespace b200lc
