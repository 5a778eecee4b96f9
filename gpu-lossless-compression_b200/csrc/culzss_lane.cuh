// CULZSS fast mode, second formulation: ONE PACKET PER LANE, greedy hash parse, token emission and
// flag-byte packing in one serial walk (NON-PARITY: the reference's buffer / token format, window
// 128, packets of 4096 bytes -- every stream decodes with the reference's DecodeKernel,
// gpu_decompress.cu:164-242 -- but not the reference encoder's matches).
//
// The code a lane runs lives in this header, which also compiles for the host, so that the CPU test
// (tests/c/culzss_lane_host.cc, tests/test_culzss_lane_cpu.py) can run it without a GPU and hand its
// output to the oracle's decoder.
//
// Per lane (STRIDE = 32 words on the GPU: word w of a lane's column sits at column[w * 32], so a
// lane only ever touches its own shared-memory bank; STRIDE = 1 on the host):
//   ring   256 bytes   input bytes of positions [hi - 256, hi), index = position & 255; holds the
//                      128-byte window behind p and >= kLaneMaxLen + 3 bytes of lookahead
//   hash   2^kLaneHashBits bytes   low 8 bits of the latest token start whose three bytes hash here
//   outq   64 bytes    compressed bytes [flushed, o), index = offset & 63; leaves 16 bytes at a time
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
constexpr u32 kHashBits = 7;
constexpr u32 kHashWords = (1u << kHashBits) / 4;
constexpr u32 kOutWords = 16;
constexpr u32 kColumnWords = kRingWords + kHashWords + kOutWords;     // 112 words = 448 bytes per lane
constexpr u32 kMaxLen = 108;      // ring: 128 behind + (113..128) ahead in 16-byte refills
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

template <int STRIDE>
struct Lane {
    u32 *ring, *hash, *outq;    // this lane's columns
    const u8 *src;              // the packet (16-byte aligned)
    u8 *dst;                    // the packet's output slot (16-byte aligned, kSlotBytes)
    u32 p;                      // next position to code
    u32 hi;                     // input loaded so far (multiple of 16)
    u32 o;                      // compressed bytes produced so far
    u32 flushed;                // compressed bytes stored to dst (multiple of 16)
    u32 fpos;                   // offset of the open group's flag byte
    u32 flags, nt;              // flag bits and tokens of the open group

    B200LC_LANE_HD u8 *byte_of(u32 *col, u32 i) const
    {
        return reinterpret_cast<u8 *>(col + (i >> 2) * STRIDE) + (i & 3u);
    }
    // four bytes at ring index i (two aligned words, funnel-shifted)
    B200LC_LANE_HD u32 ring4(u32 i) const
    {
        const u32 w = i >> 2;
        const u32 a = ring[(w & (kRingWords - 1)) * STRIDE];
        const u32 b = ring[((w + 1) & (kRingWords - 1)) * STRIDE];
        return fsr(a, b, 8 * (i & 3u));
    }

    B200LC_LANE_HD void init(u32 *column, const u8 *src_, u8 *dst_)
    {
        ring = column;
        hash = column + kRingWords * STRIDE;
        outq = hash + kHashWords * STRIDE;
        src = src_;
        dst = dst_;
        // positions -128 .. -1 are spaces (gpu_compress.cu:208); every hash entry points at -128
        for (u32 w = kRingWords / 2; w < kRingWords; ++w) ring[w * STRIDE] = 0x20202020u;
        for (u32 w = 0; w < kHashWords; ++w) hash[w * STRIDE] = 0x80808080u;
        p = 0; hi = 0; o = 1; flushed = 0; fpos = 0; flags = 0; nt = 0;
    }

    // input: 16 bytes at a time while the chunk's ring slots hold positions behind the window
    B200LC_LANE_HD bool wants_input() const { return hi < kPacket && hi <= p + 112u; }
    B200LC_LANE_HD void put_input(u32 x0, u32 x1, u32 x2, u32 x3)
    {
        u32 *q = ring + ((hi >> 2) & (kRingWords - 1)) * STRIDE;
        q[0] = x0; q[STRIDE] = x1; q[2 * STRIDE] = x2; q[3 * STRIDE] = x3;
        hi += 16;
    }

    // output: 16 bytes leave once they lie in front of the open group's flag byte
    B200LC_LANE_HD bool has_output() const { return fpos - flushed >= 16u; }
    B200LC_LANE_HD void take_output(u32 &x0, u32 &x1, u32 &x2, u32 &x3)
    {
        const u32 *q = outq + ((flushed >> 2) & (kOutWords - 1)) * STRIDE;
        x0 = q[0]; x1 = q[STRIDE]; x2 = q[2 * STRIDE]; x3 = q[3 * STRIDE];
        flushed += 16;
    }

    B200LC_LANE_HD void emit(u32 b) { *byte_of(outq, o & (4 * kOutWords - 1)) = (u8)b; ++o; }
    B200LC_LANE_HD void close_group()
    {
        *byte_of(outq, fpos & (4 * kOutWords - 1)) = (u8)flags;
        flags = 0; nt = 0;
    }

    // One token.  Requires p < kPacket and !wants_input().
    B200LC_LANE_HD void step()
    {
        if (nt == 8) {          // open the next group: its flag byte is written when it closes
            close_group();
            fpos = o;
            ++o;
        }
        const u32 x = ring4(p & 255u);
        const u32 h = ((x & 0xffffffu) * 2654435761u) >> (32 - kHashBits);
        u8 *const he = byte_of(hash, h);
        const u32 dist = (p - (u32)*he) & 255u;
        const u32 lim = min_u(min_u(dist, kMaxLen), kPacket - p);
        u32 L = 0;
        if (dist - 3u <= kWindow - 3u && lim >= 3u) {
            const u32 q = p - dist;
            u32 d = ring4(q & 255u) ^ x;
            if ((d & 0xffffffu) == 0) {
                L = 4;
                while (d == 0 && L < lim) {
                    d = ring4((q + L) & 255u) ^ ring4((p + L) & 255u);
                    L += 4;
                }
                if (d) L -= 4u - ((ffs32(d) - 1u) >> 3);
                L = min_u(L, lim);
            }
        }
        // one emission path for both token kinds (the branches above are the divergent part)
        const bool is_match = L >= 3u;
        emit(is_match ? L : (x & 0xffu));
        if (is_match) emit((p - dist) & 127u);
        else flags |= 1u << nt;
        // Hash entry := p, except:  a source that runs right up to p (L == dist) means the data
        // repeats with period `dist` -- keeping the entry lets the next token reach twice as far back
        // and copy twice as much;  an entry closer than 3 becomes usable in a moment.
        const bool keep = is_match ? (L == dist && 2u * dist <= kWindow) : (dist < 3u);
        if (!keep) *he = (u8)p;
        p += is_match ? L : 1u;
        ++nt;
    }

    // After the last token: close the open group.  Compressed size = o.
    B200LC_LANE_HD void finish() { close_group(); }
    B200LC_LANE_HD u32 last_group_bytes() const { return o - fpos; }

    static B200LC_LANE_HD u32 min_u(u32 a, u32 b) { return a < b ? a : b; }
};

}  // namespace lzss_lane
}  // namespace b200lc
