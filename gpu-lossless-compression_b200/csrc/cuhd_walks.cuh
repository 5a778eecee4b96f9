// The per-thread decode walks and the shared-memory table entries of the CUHD decoder
// (cuhd_decode.cu), kept in a header that also compiles for the host so that the CPU test
// tests/c/cuhd_walks_host.cc can run them against a bit-serial decode without a GPU.
// Decode contract: SURVEY.md appendix A.1 (cuhd-icpp/src/cuhd_gpu_decoder.cu:16-143).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B200LC_HD __host__ __device__ __forceinline__
#else
#define B200LC_HD inline
#endif

namespace b200lc {
namespace cuhd {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;

// (hi:lo) << s, upper word; (hi:lo) >> s, lower word; s in 0..31
B200LC_HD u32 fsl(u32 lo, u32 hi, u32 s)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
#endif
}
B200LC_HD u32 fsr(u32 lo, u32 hi, u32 s)
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}
B200LC_HD u32 popc32(u32 x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return (u32)__builtin_popcount(x);
#endif
}
B200LC_HD u32 clz32(u32 x)
{
#if defined(__CUDA_ARCH__)
    return __clz(x);
#else
    return x ? (u32)__builtin_clz(x) : 32u;
#endif
}

// ---------------------------------------------------------------------------------- tables
// Length of the first codeword of window `i` (the next L stream bits).  A zero-length entry
// (unused prefix of an incomplete code) would stall the reference forever; it is mapped to
// length 1 here so that garbage input still terminates.
B200LC_HD u32 first_len(const u16 *lut, u32 i, u32 L)
{
    u32 len = lut[i] & 0xffu;
    if (len == 0 || len > L) len = 1;
    return len;
}

// {length, symbol} of the first codeword with the length made safe: low byte = first_len(),
// high byte = symbol (the layout of the caller's LUT itself).
B200LC_HD u16 len_sym_entry(const u16 *lut, u32 i, u32 L)
{
    return (u16)((lut[i] & 0xff00u) | first_len(lut, i, L));
}

// Write-pass entry: bits 0..7 first symbol, 8..15 second symbol, 16..23 bits consumed, bit 31 =
// second symbol present (the window holds two whole codewords).
B200LC_HD u32 write_entry(const u16 *lut, u32 i, u32 L)
{
    const u32 e0 = lut[i];
    const u32 len0 = first_len(lut, i, L);
    const u32 i1 = (i << len0) & ((1u << L) - 1);
    const u32 e1 = lut[i1];
    const u32 len1 = first_len(lut, i1, L);
    u32 entry = (e0 >> 8) | (len0 << 16);
    if (len0 + len1 <= L) entry = (e0 >> 8) | (e1 & 0xff00u) | ((len0 + len1) << 16) | 0x80000000u;
    return entry;
}

// Write-pass entry, layout 2 (opt-in): bits 0..7 first symbol, 8..15 second symbol, bit 16 =
// second symbol present, bits 24..31 = bits consumed -- `at += e >> 24` is one instruction
// (shift-and-add) where layout 1 needs shift, mask and add.
B200LC_HD u32 write_entry2(const u16 *lut, u32 i, u32 L)
{
    const u32 e = write_entry(lut, i, L);
    return (e & 0xffffu) | ((e >> 31) << 16) | (((e >> 16) & 0xffu) << 24);
}

// Pass-A entry that covers EVERY whole codeword inside the window: bits 12..15 = bits consumed
// (1..13), bit 12-o = a codeword starts at offset o (o = 1..12; offset 0 always starts one), so
// that `entry << 19` puts the start of offset o at bit 31-o and drops the bit count (its lowest
// bit lands on bit 31, which is set anyway).
// A codeword at offset o is known to be whole iff its length <= L - o: the window is
// zero-filled beyond L bits, and a prefix code is decided by the codeword's own bits.
B200LC_HD u16 multi_entry(const u16 *lut, u32 i, u32 L)
{
    const u32 mask = (1u << L) - 1;
    u32 o = first_len(lut, i, L);
    u32 e = 0;
    while (o < L) {
        const u32 len = first_len(lut, (i << o) & mask, L);
        if (o + len > L) break;
        e |= 1u << (12 - o);
        o += len;
    }
    return (u16)(e | (o << 12));
}

// ---------------------------------------------------------------------------------- walks
// Round 0: decode the subsequence from bit 0, remember every codeword start.
template <int S>
B200LC_HD void walk_record(const u32 (&u)[S + 1], const u8 *tab, u32 shift, u32 (&m)[S], u32 &end,
                           u32 &cnt)
{
    u32 at = 0, c = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        u32 mj = 0;
        while (at < 32) {
            mj |= 0x80000000u >> at;
            const u32 w = fsl(nxt, cur, at);
            at += tab[w >> shift];
        }
        m[j] = mj;
        c += popc32(mj);
        at -= 32;
    }
    end = at;
    cnt = c;
}

// Same result as walk_record from multi_entry() tables: one lookup advances over all whole
// codewords of the window (1.64 on the C2 code), i.e. 0.6x the dependent lookups.  Starts
// that fall into the next unit travel in `carry` and seed its mask; a step of the last unit may
// run past the first codeword of the next subsequence, so the exit state is the first start
// at or after the boundary, not where the walk stopped.
template <int S>
B200LC_HD void walk_record_multi(const u32 (&u)[S + 1], const u16 *mtab, u32 shift, u32 (&m)[S],
                                 u32 &end, u32 &cnt)
{
    u32 at = 0, c = 0, carry = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        u32 mj = carry;
        carry = 0;
        while (at < 32) {
            const u32 w = fsl(nxt, cur, at);
            const u32 e = mtab[w >> shift];
            const u32 E = 0x80000000u | (e << 19);
            mj |= E >> at;
            carry |= fsr(0u, E, at);
            at += e >> 12;
        }
        m[j] = mj;
        c += popc32(mj);
        at -= 32;
    }
    end = carry ? clz32(carry) : at;
    cnt = c;
}

// Decode from entry state `a` until the walk lands on a codeword start of the recorded path
// (then the rest of the subsequence is the recorded path: end = e0) or runs off the end.
// No early return: a lane that has merged idles through the remaining unit loops so that the
// warp reconverges after every unit (an early exit makes the lanes run the later loops one at
// a time -- measured: 52% of all issued instructions at 1 active thread).
template <int S>
B200LC_HD void walk_merge(const u32 (&u)[S + 1], const u32 (&m)[S], u32 a, u32 e0, const u8 *tab,
                          u32 shift, u32 &end, u32 &cnt)
{
    u32 at = a, k = 0, rest = 0;
    bool done = false;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1], mj = m[j];
        if (done) {
            rest += popc32(mj);
        } else {
            while (at < 32) {
                const u32 bit = 0x80000000u >> at;
                if (mj & bit) {
                    rest = popc32(mj & (bit | (bit - 1)));
                    done = true;
                    break;
                }
                const u32 w = fsl(nxt, cur, at);
                at += tab[w >> shift];
                ++k;
            }
            if (!done) at -= 32;
        }
    }
    end = done ? e0 : at;
    cnt = k + rest;
}

// Write pass: decode from the true entry state, symbol i of this subsequence goes to dst[i].
// Table entries carry TWO symbols when the window holds two whole codewords: bits 0..7 first
// symbol, 8..15 second symbol, 16..23 bits consumed, bit 31 = second symbol present.  A second
// symbol that starts beyond this subsequence is also the next subsequence's first symbol: it is
// stored twice with the same value at the same position (or beyond the tile, where nothing is
// copied out).  With CHECK, only tile-local positions in [lo, hi) are stored (staging-window
// overflow path).
template <int S, bool CHECK>
B200LC_HD void walk_write(const u32 (&u)[S + 1], const u32 *tab, u32 shift, u32 a, u8 *dst, u32 pos,
                          u32 lo, u32 hi)
{
    u32 at = a;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        while (at < 32) {
            const u32 w = fsl(nxt, cur, at);
            const u32 e = tab[w >> shift];
            if (!CHECK || (pos >= lo && pos < hi)) dst[pos] = (u8)e;
            if ((int)e < 0 && (!CHECK || (pos + 1 >= lo && pos + 1 < hi))) dst[pos + 1] = (u8)(e >> 8);
            pos += 1 + (e >> 31);
            at += (e >> 16) & 0xffu;
        }
        at -= 32;
    }
}

// Write pass on write_entry2() tables with a running store pointer instead of base + position.
template <int S, bool CHECK>
B200LC_HD void walk_write2(const u32 (&u)[S + 1], const u32 *tab, u32 shift, u32 a, u8 *dst, u32 pos,
                           u32 lo, u32 hi)
{
    u32 at = a;
    u8 *q = dst + pos;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        while (at < 32) {
            const u32 w = fsl(nxt, cur, at);
            const u32 e = tab[w >> shift];
            const bool two = (e & 0x10000u) != 0;
            if (CHECK) {
                const u32 at_pos = (u32)(q - dst);
                if (at_pos >= lo && at_pos < hi) q[0] = (u8)e;
                if (two && at_pos + 1 >= lo && at_pos + 1 < hi) q[1] = (u8)(e >> 8);
            } else {
                q[0] = (u8)e;
                if (two) q[1] = (u8)(e >> 8);
            }
            q += 1;
            if (two) q += 1;
            at += e >> 24;
        }
        at -= 32;
    }
}

// --------------------------------------------------------------------------- counting walk
// Round-2 pass A: no codeword-start masks.  One byte per window of LM >= L bits:
//   bits 0..3  bits consumed by the first n whole codewords of the window, n = 1..3
//   bits 6..7  n
// so that `acc += entry` advances the bit position (bits 0..5 of acc, < 64 inside a unit) and the
// codeword count (bits 6 and up) with ONE add, and the funnel shift takes its amount from the low
// five bits of the same register.  A codeword at offset o of the window is whole iff
// o + length <= LM (the bits beyond the window are unknown; a prefix code is decided by the
// codeword's own bits).  n is capped at 3 (two bits); the cap `max_n` = 1 gives the single-step
// table used near the end of a subsequence.
B200LC_HD u8 count_entry(const u16 *lut, u32 i, u32 L, u32 LM, u32 max_n)
{
    const u32 wmask = (1u << LM) - 1;
    u32 o = 0, n = 0;
    while (n < max_n) {
        const u32 idx = ((i << o) & wmask) >> (LM - L);
        const u32 len = first_len(lut, idx, L);
        if (n && o + len > LM) break;
        o += len;
        ++n;
        if (o >= LM) break;
    }
    return (u8)(o | (n << 6));
}

// Exit state and codeword count of the subsequence entered at bit `a` (0 <= a < 32).
// Contract (same as walk_record / walk_merge): cnt = codewords that START in [a, 32 S),
// end = first codeword start at or after 32 S, minus 32 S.  Multi-codeword lookups (mtab, window
// LM bits: shift_m = 32 - LM) are taken while every codeword they cover starts inside the
// subsequence, i.e. while the position is <= 32 S - LM; the last few bits are walked one codeword
// at a time (stab: count_entry(.., max_n = 1) on the L-bit window, shift = 32 - L).
template <int S>
B200LC_HD void walk_count(const u32 (&u)[S + 1], const u8 *mtab, u32 shift_m, const u8 *stab, u32 shift,
                          u32 a, u32 &end, u32 &cnt)
{
    u32 acc = a, c = 0;
#pragma unroll
    for (int j = 0; j < S - 1; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        while (!(acc & 32u)) {
            const u32 w = fsl(nxt, cur, acc);
            acc += mtab[w >> shift_m];
        }
        c += acc >> 6;
        acc = (acc & 63u) - 32u;
    }
    {
        const u32 cur = u[S - 1], nxt = u[S];
        const u32 lim = shift_m;                       // 32 - LM
        while ((acc & 63u) <= lim) {
            const u32 w = fsl(nxt, cur, acc);
            acc += mtab[w >> shift_m];
        }
        while (!(acc & 32u)) {
            const u32 w = fsl(nxt, cur, acc);
            acc += stab[w >> shift];
        }
        c += acc >> 6;
        acc = (acc & 63u) - 32u;
    }
    end = acc;
    cnt = c;
}

// ------------------------------------------------------------------- packed write pass
// Round-2 pass B.  One u32 per window of LW >= L bits:
//   bits  0..23  the symbols of the first n whole codewords of the window (n = 1..3, absent = 0)
//   bits 24..27  bits consumed by them          } the top byte is count_entry(.., LW, 3): one
//   bits 30..31  n                              } shift-and-add advances the position
// The lane packs its symbols into 32-bit words in a register and stores WORDS into the staging
// buffer (one conflict-prone shared-memory store per four symbols instead of one per symbol):
//   pow = 1 << 8 k, k = bytes pending in buf;   (hi : lo) = symbols * pow + buf   (one IMAD.WIDE)
//   pow' = rotl(pow, 8 n); it wraps (pow' < pow) exactly when the word is full: store lo, keep hi.
B200LC_HD u32 write_entry3(const u16 *lut, u32 i, u32 L, u32 LW)
{
    const u32 wmask = (1u << LW) - 1;
    u32 o = 0, n = 0, syms = 0;
    while (n < 3) {
        const u32 idx = ((i << o) & wmask) >> (LW - L);
        const u32 len = first_len(lut, idx, L);
        if (n && o + len > LW) break;
        syms |= (u32)(lut[idx] >> 8) << (8 * n);
        o += len;
        ++n;
        if (o >= LW) break;
    }
    return syms | (o << 24) | (n << 30);
}

B200LC_HD u32 rotl32(u32 x, u32 s) { return fsl(x, x, s); }

// Shared-memory accesses of the packed write walk through 32-bit shared addresses on the device:
// with generic pointers the compiler rebuilt the shared window base (S2R SR_CgaCtaId + LEA) in every
// unit loop and added base + offset in front of every staging store.
#if defined(__CUDA_ARCH__)
typedef u32 smem_addr;
__device__ __forceinline__ smem_addr smem_of(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ u32 smem_ld32(smem_addr a)
{
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void smem_st32(smem_addr a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
#else
typedef uintptr_t smem_addr;
inline smem_addr smem_of(const void *p) { return reinterpret_cast<uintptr_t>(p); }
inline u32 smem_ld32(smem_addr a) { return *reinterpret_cast<const u32 *>(a); }
inline void smem_st32(smem_addr a, u32 v) { *reinterpret_cast<u32 *>(a) = v; }
#endif

// Phase 1 of the packed write: the `cnt` symbols of the subsequence entered at bit `a` go to
// stage[d0 .. d0 + cnt) (stage 4-byte aligned).  Stores every word that fills up, INCLUDING the
// first one whose bytes below d0 are zero -- they belong to the lanes in front, which hold them
// back as pending bytes and store them in phase 2 (walk_write3_tail) after a warp barrier.
// Exactly cnt symbols are emitted: the last lookups of the subsequence are clamped, so that no
// word of the successor is touched.  buf0 = bytes already in front of d0 in its word that nobody
// will store again (bytes carried over from the previous staging round), else 0.  Returns the
// pending bytes (word (d0 + cnt) >> 2).
template <int S>
B200LC_HD u32 walk_write3(const u32 (&u)[S + 1], smem_addr tab, u32 shift_w, u32 a, u32 cnt, smem_addr stage,
                          u32 d0, u32 buf0)
{
    u32 acc = a, buf = buf0;
    u32 pow = 1u << (8 * (d0 & 3u));
    const smem_addr wa0 = stage + (d0 & ~3u);
    smem_addr wa = wa0;                                // address of the word being filled
#define B200LC_EMIT3(e, syms, r)                                           \
    {                                                                      \
        const unsigned long long prod = (unsigned long long)(syms) * pow;  \
        const u32 lo = (u32)prod | buf;                                    \
        const u32 pow2 = rotl32(pow, (r));                                 \
        buf = lo;                                                          \
        if (pow2 < pow) {                                                  \
            smem_st32(wa, lo);                                             \
            wa += 4;                                                       \
            buf = (u32)(prod >> 32);                                       \
        }                                                                  \
        pow = pow2;                                                        \
    }
#pragma unroll
    for (int j = 0; j < S - 1; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        while (!(acc & 32u)) {
            const u32 w = fsl(nxt, cur, acc);
            const u32 e = smem_ld32(tab + 4 * (smem_addr)(w >> shift_w));
            acc += e >> 24;
            B200LC_EMIT3(e, e & 0xffffffu, (e >> 27) & 0x18u)
        }
        acc = (acc & 63u) - 32u;
    }
    {
        const u32 cur = u[S - 1], nxt = u[S];
        const u32 lim = shift_w;                       // 32 - LW: every codeword of the window starts inside
        while ((acc & 63u) <= lim) {
            const u32 w = fsl(nxt, cur, acc);
            const u32 e = smem_ld32(tab + 4 * (smem_addr)(w >> shift_w));
            acc += e >> 24;
            B200LC_EMIT3(e, e & 0xffffffu, (e >> 27) & 0x18u)
        }
        const u32 k = (31u - clz32(pow)) >> 3;
        u32 rem = cnt - ((u32)(wa - wa0) + k - (d0 & 3u));
        while ((int)rem > 0 && !(acc & 32u)) {
            const u32 w = fsl(nxt, cur, acc);
            const u32 e = smem_ld32(tab + 4 * (smem_addr)(w >> shift_w));
            acc += e >> 24;
            u32 n = e >> 30;
            if (n > rem) n = rem;
            rem -= n;
            const u32 syms = e & (0xffffffu >> (24u - 8u * n));
            B200LC_EMIT3(e, syms, 8u * n)
        }
    }
#undef B200LC_EMIT3
    return buf;
}

// buf0 for walk_write3: stage[0, fill) are final bytes of earlier rounds.
B200LC_HD u32 walk_write3_head(const u8 *stage, u32 d0, u32 fill)
{
    const u32 w0 = d0 & ~3u;
    return w0 < fill ? *reinterpret_cast<const u32 *>(stage + w0) & ((1u << (8 * (d0 & 3u))) - 1u) : 0u;
}

// Phase 2: the pending bytes of word (d0 + cnt) >> 2 that are this lane's own.
B200LC_HD void walk_write3_tail(u8 *stage, u32 d0, u32 cnt, u32 buf)
{
    const u32 end = d0 + cnt, w0 = end & ~3u;
#pragma unroll
    for (u32 i = 0; i < 3; ++i) {
        const u32 p = w0 + i;
        if (p >= d0 && p < end) stage[p] = (u8)(buf >> (8 * i));
    }
}

// The same symbols one byte at a time, only positions in [lo, hi) stored: symbol i of the
// subsequence goes to dst[pos + i].  For the rare lane that straddles a staging window.
template <int S>
B200LC_HD void walk_write3_bytes(const u32 (&u)[S + 1], const u32 *tab3, u32 shift_w, u32 a, u32 cnt, u8 *dst,
                                 u32 pos, u32 lo, u32 hi)
{
    u32 acc = a, done = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        while (!(acc & 32u) && done < cnt) {
            const u32 w = fsl(nxt, cur, acc);
            const u32 e = tab3[w >> shift_w];
            acc += e >> 24;
            const u32 n = e >> 30;
            for (u32 t = 0; t < n && done < cnt; ++t, ++done) {
                const u32 p = pos + done;
                if (p >= lo && p < hi) dst[p] = (u8)(e >> (8 * t));
            }
        }
        acc = (acc & 63u) - 32u;
    }
}

// ------------------------------------------------------------------- decode-once building blocks
// (not used by the kernel yet.)  Pass A that also keeps the symbols of the
// path from bit 0 in a per-subsequence slot, so that the write pass becomes a copy:
//   symbols(entry state a) = the k symbols walk_merge_skip() steps over before it lands on the
//                            recorded path  +  slot[skip .. c0)
// with k + (c0 - skip) = the count walk_merge reports.  A subsequence with more than SLOT symbols
// on its recorded path reports overflow and is decoded by the second walk as today.
template <int S, int SLOT>
B200LC_HD bool walk_record_sym(const u32 (&u)[S + 1], const u16 *tab, u32 shift, u32 (&m)[S], u32 &end,
                               u32 &cnt, u8 *slot)
{
    u32 at = 0, c = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        u32 mj = 0;
        while (at < 32) {
            mj |= 0x80000000u >> at;
            const u32 w = fsl(nxt, cur, at);
            const u32 e = tab[w >> shift];          // {u8 len, u8 symbol} of the first codeword
            if (c < (u32)SLOT) slot[c] = (u8)(e >> 8);
            ++c;
            at += e & 0xffu;
        }
        m[j] = mj;
        at -= 32;
    }
    end = at;
    cnt = c;
    return c <= (u32)SLOT;
}

// walk_merge that also reports where the recorded path takes over: `k` symbols are decoded from
// entry state a before the walk lands on recorded start number `skip` (skip = c0 when it never
// lands: then all k symbols are the subsequence's own and the slot is not used).
template <int S>
B200LC_HD void walk_merge_skip(const u32 (&u)[S + 1], const u32 (&m)[S], u32 a, u32 e0, u32 c0,
                               const u8 *ltab, u32 shift, u32 &end, u32 &k_out, u32 &skip_out)
{
    u32 at = a, k = 0, before = 0;
    bool done = false;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1], mj = m[j];
        if (!done) {
            while (at < 32) {
                const u32 bit = 0x80000000u >> at;
                if (mj & bit) {
                    before += popc32(mj & ~(bit | (bit - 1)));   // recorded starts in front of the landing bit
                    done = true;
                    break;
                }
                const u32 w = fsl(nxt, cur, at);
                at += ltab[w >> shift];
                ++k;
            }
            if (!done) { at -= 32; before += popc32(mj); }
        }
    }
    end = done ? e0 : at;
    k_out = k;
    skip_out = done ? before : c0;
}

// The k symbols in front of the landing point, decoded again from entry state a (k is 1-3 on
// real data: the walks re-synchronise within a few codewords).
template <int S>
B200LC_HD void walk_emit(const u32 (&u)[S + 1], const u16 *tab, u32 shift, u32 a, u32 k, u8 *dst)
{
    u32 at = a, i = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        while (at < 32 && i < k) {
            const u32 w = fsl(nxt, cur, at);
            const u32 e = tab[w >> shift];
            dst[i++] = (u8)(e >> 8);
            at += e & 0xffu;
        }
        at -= 32;      // only meaningful while i < k
    }
}

}  // namespace cuhd
}  // namespace b200lc
