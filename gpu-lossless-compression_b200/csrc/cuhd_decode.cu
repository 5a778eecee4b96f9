// Self-synchronising canonical-Huffman decode for sm_100a  (hot path 3).
//
// Replaces the reference's 4 phases / 6 kernels / host do-while
// (cuhd-icpp/src/cuhd_gpu_decoder.cu:145-523) with ONE persistent kernel.  Round-2 formulation
// (round 1: CTA-wide sub-tiles, codeword-start masks and merge walks, 11 CTA barriers per 8 KiB;
// DESIGN.md section 3.1 keeps the history):
//
//   * subsequence = S units = 32 bytes = one lane; warp-step = 32 subsequences = 1 KiB;
//     segment = K consecutive warp-steps handled by ONE warp; piece = kWarps segments handled
//     by one CTA between two look-backs.  Pieces are handed out through a ticket counter.
//   * pass A (per warp, no CTA barrier): every lane walks its subsequence from bit 0 with
//     multi-codeword lookups that advance the bit position AND the codeword count with one add
//     (walk_count in cuhd_walks.cuh); entry states are then chained through the warp with
//     shuffles, and a lane whose entry state is not 0 walks once more from the true state.  This
//     replaces phase 1's "decode the next subsequence again until the sync point repeats"
//     (:188-231).  The exit state of a step is the exact entry state of the warp's next step.
//   * the entry state of a segment (and of a piece) is GUESSED from a walk over the subsequence
//     in front of it (a path from bit 0 re-synchronises within a subsequence with probability
//     > 0.999 on real data) and VERIFIED against the predecessor's exit state: inside the CTA
//     after pass A, between pieces by the look-back.  A wrong guess is repaired by re-running
//     pass A of that segment with the true state.  This replaces phase 2's host loop with D2H
//     flag copies (:458-495).
//   * pieces are chained by a warp-wide decoupled look-back over 16-byte descriptors
//     {assumed entry state, exit state, symbols}; a chain of aggregates is usable when every
//     link's assumed entry state equals its predecessor's exit state.  Replaces phase 3's three
//     passes over 16-byte sync points (:498-509).
//   * pass B (per warp): decode from the now known entry states with three-symbol table entries;
//     a lane packs its symbols into 32-bit words in a register and stores words into a per-warp
//     shared-memory staging buffer (walk_write3 in cuhd_walks.cuh: byte stores were 44 % of all
//     shared-memory wavefronts); the buffer leaves with aligned 16-byte streaming stores.
//     Phase 4 wrote single bytes to global memory (:101-105).
//
// HBM traffic: units in (pass B re-reads them through L2) + symbols out + 16 B per piece.
//
// Decode contract (bit-exact with the reference): out[i] = symbol of
// the i-th codeword when reading the stream MSB-first from bit 0 of unit 0 through the flat LUT
// `{u8 num_bits, u8 symbol}[1 << max_codeword_length]`; units past n_units read as zero.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "cuhd_walks.cuh"
#include "../../include/b200lc.h"

namespace b200lc {
namespace cuhd {

constexpr int S = 8;                      // units per subsequence
constexpr int kWarps = 16;                // warps per CTA = segments per piece (two CTAs per SM share
constexpr int kThreads = kWarps * 32;     //   the SM: the 32 KiB write table is paid twice, not four times)
constexpr u32 kStageBytes = 2048;         // per-warp output staging
constexpr u32 kWin = kStageBytes - 32;    // symbols staged per round (<= 15 carried bytes fit in front)
constexpr u32 kMultiBits = 13;            // window of the counting table and of the write table (>= L)

constexpr u64 kValid = 1ull << 63;
constexpr u64 kCountMask = (1ull << 56) - 1;

// One 16-byte descriptor per piece:
//   agg : bit 63 valid | bits 56..59 ASSUMED entry state A | bits 48..51 exit state X |
//         bits 0..31 symbols T -- X and T hold if the piece is entered in state A
//   incl: bit 63 valid | bits 56..59 true exit state | bits 0..55 symbols that start before the
//         next piece of the stream (inclusive over all earlier pieces)
struct __align__(16) PieceDesc {
    u64 agg;
    u64 incl;
};

// One independent bit stream (codes start at bit 0 of its first unit, entry state 0) and the
// range of pieces that decodes it.  A single-stream call carries its view inside the kernel
// parameters; a batch call (b200lc_cuhd_decode_batch) keeps one view per stream in global memory
// plus the stream number of every piece.
struct StreamView {
    const u32 *units;
    u64 n_units;
    u8 *out;
    u64 n_out;
    u32 first_piece;     // global number of the stream's first piece
    u32 aligned;         // units pointer is 16-byte aligned: lanes fetch their 32 bytes with two 16-byte loads
};

struct DecodeParams {
    StreamView one;      // the stream of a single-stream call
    const StreamView *streams;    // batch: views in global memory (nullptr: single stream)
    const u32 *piece_stream;      // batch: stream number of every piece
    const u16 *lut;      // {u8 num_bits, u8 symbol} little-endian pairs
    u32 max_len;         // L
    u32 multi_bits;      // LM: window of the counting table and of the write table, L <= LM <= 15
    PieceDesc *desc;
    u32 *ticket;
    u32 num_pieces;      // pieces [first_piece, num_pieces) are decoded by this launch
    u32 first_piece;
};

template <int K>
struct SmemLayout {
    // Two pieces are in flight per CTA (pass A of piece n runs before pass B of piece n - 1), and a
    // warp may run ahead of its CTA by most of an iteration: per-piece state is indexed by
    // turn % 2 where only the owning warp touches it (saved) and by turn % 3 where others read it.
    u16 saved[2][kWarps][K][32];            // pass A result per subsequence: entry state << 12 | symbols
    __align__(16) u8 stage[kWarps][kStageBytes];
    u32 wt[3][kWarps];                      // per segment: symbols,
    u8 wa[3][kWarps];                       //   assumed entry state,
    u8 wx[3][kWarps];                       //   exit state
    u64 base;
    u32 next_piece;
    u32 true_entry;
    u32 redo;
    StreamView view[3];                     // batch: stream of the piece of each turn
};

__device__ __forceinline__ void st_stream_v4(void *p, const uint4 &v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// The tables in shared memory.  The walks must be inlined into the kernel with table addresses the
// compiler can see are warp-uniform shared-memory addresses (LDS [R + UR + imm]: 6 instructions
// per lookup step); handing them to a __noinline__ function made them generic pointers (LD +
// 64-bit arithmetic, +19 % instructions measured), and offsets from a namespace-scope shared
// array still cost one IADD3 per step.  Hence ONE inlined call site for pass A in the kernel.
struct Tables {
    const u32 *wtab;     // write pass: <= 3 symbols per entry, window LM bits (write_entry3)
    smem_addr wtab_s;    //   the same as a 32-bit shared address (walk_write3)
    const u8 *mtab;      // counting: <= 3 codewords per entry, window LM bits
    const u8 *stab;      // counting: 1 codeword per entry, window L bits
    u32 shift, shift_m;  // 32 - L, 32 - LM
};

// My warp's segment of a piece, indexed with 32-bit numbers relative to its first unit: units
// seg_units[0 .. seg_rem) (zero beyond), seg_subs subsequences that hold units.
struct Segment {
    const u32 *seg_units;
    u32 seg_rem, seg_subs;
    bool aligned, stream_start;
};

template <int K>
__device__ __forceinline__ Segment make_segment(const StreamView &V, u32 lp, u32 warp)
{
    constexpr u32 kSegUnits = (u32)K * 32 * S;
    Segment g;
    const u64 first_unit = ((u64)lp * kWarps + warp) * (u64)kSegUnits;
    g.seg_units = V.units + first_unit;
    g.seg_rem = first_unit >= V.n_units ? 0u : (u32)min(V.n_units - first_unit, (u64)0x7fffff00u);
    g.seg_subs = min((u32)K * 32, (g.seg_rem + S - 1) / S);
    g.aligned = V.aligned != 0;
    g.stream_start = first_unit == 0;
    return g;
}

// segments of piece lp that hold stream units (the stream's last piece may have fewer)
template <int K>
__device__ __forceinline__ u32 real_segments(const StreamView &V, u32 lp)
{
    constexpr u32 kSegUnits = (u32)K * 32 * S;
    return (u32)min((u64)kWarps, (V.n_units - (u64)lp * (kWarps * kSegUnits) + (kSegUnits - 1)) / (u64)kSegUnits);
}

// L2 residency hints (round 1: 1.87x the stream was read from DRAM because the re-read of pass B
// missed an L2 that the output had streamed through): pass A fetches the units with evict_last,
// pass B -- their last use -- with evict_first; the output leaves with st.global.cs.
__device__ __forceinline__ u64 l2_policy_keep()
{
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ u64 l2_policy_done()
{
    u64 pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_hint(const uint4 *ptr, u64 pol)
{
    uint4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ u32 ldg_hint(const u32 *ptr, u64 pol)
{
    u32 v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(ptr), "l"(pol));
    return v;
}

// the S units of the segment's subsequence `sub` (may be -1: the one in front of the segment)
// plus one lookahead unit
__device__ __forceinline__ void load_units(const Segment &g, int sub, u32 (&u)[S + 1], u64 pol)
{
    const int first = sub * S;
    if (g.aligned && first + S + 1 <= (int)g.seg_rem) {
        const uint4 *src = reinterpret_cast<const uint4 *>(g.seg_units + first);
#pragma unroll
        for (int q = 0; q < S / 4; ++q) {
            const uint4 v = ldg_hint(src + q, pol);
            u[4 * q + 0] = v.x; u[4 * q + 1] = v.y; u[4 * q + 2] = v.z; u[4 * q + 3] = v.w;
        }
        u[S] = ldg_hint(g.seg_units + first + S, pol);
    } else {
#pragma unroll
        for (int j = 0; j <= S; ++j) u[j] = first + j < (int)g.seg_rem ? ldg_hint(g.seg_units + first + j, pol) : 0u;
    }
}

// Pass A of one segment entered in state `entry` (entry > 15: guess it).  Fills
// saved[subsequence] = entry state << 12 | symbols; returns assumed entry state of the segment |
// exit state << 8 | symbols << 32.
//
// Lane l owns the CHUNK of K consecutive subsequences [l K, (l + 1) K) and walks them one after
// the other: the exit state of one is the exact entry state of the next, so every subsequence is
// walked ONCE (lanes on consecutive subsequences had to walk from bit 0 first and again from the
// true state: 1.83 walks per subsequence on C2).  Only the entry state of a chunk is a guess --
// the path from bit 0 of the subsequence in front of it, one extra walk per K -- and it is
// verified against the exit state of the lane before; a lane that guessed wrong walks its chunk
// again (0.03 % of the guesses on C2).  The lanes read 32-byte sectors 32 K bytes apart: every
// sector is still fetched once.
template <int K>
__device__ __forceinline__ u64 segment_pass_a(const Segment &g, const Tables &tb, u32 entry, u16 *saved)
{
    const u8 *const mtab = tb.mtab, *const stab = tb.stab;
    const u32 lane = threadIdx.x & 31;
    if (g.seg_subs == 0) {
        const u32 e = entry > 15 ? 0u : entry;
        return (u64)(e | (e << 8));
    }
    const u32 first_sub = lane * K;
    const bool chunk_real = first_sub < g.seg_subs;
    u32 my_entry = 0;
    if (chunk_real && (lane > 0 || (entry > 15 && !g.stream_start))) {
        u32 u[S + 1], c;
        load_units(g, (int)first_sub - 1, u, l2_policy_keep());
        walk_count<S>(u, mtab, tb.shift_m, stab, tb.shift, 0u, my_entry, c);
    }
    if (lane == 0 && entry <= 15) my_entry = entry;
    const u32 seg_entry = __shfl_sync(0xffffffffu, my_entry, 0);
    u32 my_exit = 0, my_total = 0;
    bool active = chunk_real;
    for (;;) {
        if (active) {
            u32 e = my_entry;
            my_total = 0;
#pragma unroll 1
            for (u32 sub = first_sub; sub < first_sub + K && sub < g.seg_subs; ++sub) {
                u32 u[S + 1], ne, nc;
                load_units(g, (int)sub, u, l2_policy_keep());
                walk_count<S>(u, mtab, tb.shift_m, stab, tb.shift, e, ne, nc);
                saved[sub] = (u16)((e << 12) | nc);
                my_total += nc;
                e = ne;
            }
            my_exit = e;
        }
        // every chunk must have been entered in the state the chunk before it left
        const u32 prev_exit = __shfl_up_sync(0xffffffffu, my_exit, 1);
        active = chunk_real && lane > 0 && prev_exit != my_entry;
        if (!__any_sync(0xffffffffu, active)) break;
        if (active) my_entry = prev_exit;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_total += __shfl_xor_sync(0xffffffffu, my_total, d);
    const u32 exit_state = __shfl_sync(0xffffffffu, my_exit, (g.seg_subs - 1) / K);
    return (u64)(seg_entry | (exit_state << 8)) | ((u64)my_total << 32);
}

// Pass B of one segment: decode from the saved entry states, stage, store.  gstart = output
// index of the segment's first symbol.
template <int K>
__device__ __forceinline__ void segment_pass_b(const Segment &g, const Tables &tb, const u16 *saved,
                                               u8 *stage, smem_addr stage_s, u8 *out, u64 n_out, u64 gstart)
{
    const u32 lane = threadIdx.x & 31;
    if (g.seg_subs == 0 || gstart >= n_out) return;
    // stage[0] corresponds to the 16-byte aligned output address optr; of the staged bytes only
    // those at [lo_ok, hi_ok) relative to optr are mine and inside the output
    u32 lo_ok = (u32)((reinterpret_cast<uintptr_t>(out) + gstart) & 15u);
    u8 *optr = out + gstart - lo_ok;
    u32 hi_ok = (u32)min(n_out - gstart + lo_ok, (u64)0x7fffff00u);
    u32 fill = lo_ok;
    for (u32 step = 0; step * 32 < g.seg_subs; ++step) {
        if (fill >= hi_ok) break;      // the rest lies beyond the output
        const u32 sub = step * 32 + lane;
        u32 u[S + 1];
        if (sub < g.seg_subs) load_units(g, (int)sub, u, l2_policy_done());
        const u32 sv = sub < g.seg_subs ? (u32)saved[sub] : 0u;
        const u32 my_start = sv >> 12, my_cnt = sv & 0xfffu;
        const u32 incl = warp_incl_scan(my_cnt);
        const u32 pre = incl - my_cnt;
        const u32 total = __shfl_sync(0xffffffffu, incl, 31);
        for (u32 lo = 0; lo < total; lo += kWin) {
            const u32 hi = min(total, lo + kWin);
            // symbol at step-local position q goes to stage[fill - lo + q]
            const bool mine = my_cnt != 0 && pre < hi && pre + my_cnt > lo;
            const bool fast = mine && pre >= lo && pre + my_cnt <= hi;
            const u32 d0 = fill + pre - lo;
            u32 pend = 0;
            if (fast)
                pend = walk_write3<S>(u, tb.wtab_s, tb.shift_m, my_start, my_cnt, stage_s, d0,
                                      walk_write3_head(stage, d0, fill));
            __syncwarp();
            // bytes that share a word with a neighbour's first word, and the rare lane that
            // straddles the window, come after the word stores
            if (fast)
                walk_write3_tail(stage, d0, my_cnt, pend);
            else if (mine)
                walk_write3_bytes<S>(u, tb.wtab, tb.shift_m, my_start, my_cnt, stage + ((int)fill - (int)lo),
                                     pre, lo, hi);
            __syncwarp();
            fill += hi - lo;
            // whole 16-byte vectors of stage[0, fill) leave; the rest moves to the front
            const u32 nvec = fill >> 4;
            for (u32 i = lane; i < nvec; i += 32) {
                const uint4 v = *reinterpret_cast<const uint4 *>(stage + 16 * i);
                if (16 * i >= lo_ok && 16 * i + 16 <= hi_ok) {
                    st_stream_v4(optr + 16 * i, v);
                } else {
                    for (u32 b = 16 * i; b < 16 * i + 16; ++b)
                        if (b >= lo_ok && b < hi_ok) optr[b] = stage[b];
                }
            }
            const u32 tail = fill & 15u;
            u8 keep = 0;
            if (lane < tail) keep = stage[16 * nvec + lane];
            __syncwarp();
            if (lane < tail) stage[lane] = keep;
            __syncwarp();
            optr += 16 * nvec;
            lo_ok = lo_ok > 16 * nvec ? lo_ok - 16 * nvec : 0u;
            hi_ok = hi_ok > 16 * nvec ? hi_ok - 16 * nvec : 0u;
            fill = tail;
        }
    }
    // the segment's last partial vector
    for (u32 i = lane; i < fill; i += 32)
        if (i >= lo_ok && i < hi_ok) optr[i] = stage[i];
    __syncwarp();
}

// ---------------------------------------------------------------------------------- kernel
// Two CTAs of 16 warps per SM (64 registers per thread).
//
// Software pipeline over the pieces a CTA claims: iteration n runs pass A of piece n, publishes
// its aggregate, and only then chains piece n - 1 (look-back) and runs its pass B.  By that
// time the predecessors of piece n - 1 have long published theirs, so the look-back neither
// waits nor spins (measured without the pipeline: 13 % of all issued instructions were
// look-back polls and 16 % of the stall samples sat on the barrier behind them).
template <int K, bool BATCH>
__global__ void __launch_bounds__(kThreads, 2) cuhd_decode_kernel(const DecodeParams p)
{
    using Smem = SmemLayout<K>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    const u32 L = p.max_len, LM = p.multi_bits;
    u32 *wtab = reinterpret_cast<u32 *>(smem_raw + ((sizeof(Smem) + 127) & ~size_t(127)));
    u8 *mtab = reinterpret_cast<u8 *>(wtab + (size_t(1) << LM));
    u8 *stab = mtab + (size_t(1) << LM);

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31;
    const u32 warp = tid >> 5;

    // LUT -> shared-memory tables.  A zero-length entry (unused prefix of an incomplete code)
    // would stall the reference forever; first_len() maps it to length 1 so that garbage input
    // still terminates.
    for (u32 i = tid; i < (1u << L); i += kThreads) stab[i] = count_entry(p.lut, i, L, L, 1);
    for (u32 i = tid; i < (1u << LM); i += kThreads) {
        const u32 e = write_entry3(p.lut, i, L, LM);
        wtab[i] = e;
        mtab[i] = (u8)(e >> 24);       // == count_entry(p.lut, i, L, LM, 3)
    }
    Tables tb;
    const smem_addr smem_base = smem_of(smem_raw);
    tb.wtab = wtab; tb.mtab = mtab; tb.stab = stab;
    tb.wtab_s = smem_base + (smem_addr)(reinterpret_cast<unsigned char *>(wtab) - smem_raw);
    // keep the address in a register: the compiler otherwise rebuilds the shared window base
    // (S2R SR_CgaCtaId, MOV, LEA) in front of every 32-bit unit of the write walk
    asm volatile("mov.u32 %0, %0;" : "+r"(tb.wtab_s));
    tb.shift = 32 - L; tb.shift_m = 32 - LM;

    if (tid == 0) {
        const u32 t0 = p.first_piece + atomicAdd(p.ticket, 1u);
        sm.next_piece = t0;
        if (BATCH && t0 < p.num_pieces) sm.view[0] = p.streams[p.piece_stream[t0]];
    }
    __syncthreads();

    bool have_prev = false;
    u32 ppiece = 0;
    for (u32 turn = 0;; ++turn) {
        const u32 piece = sm.next_piece;
        const bool have_cur = piece < p.num_pieces;
        if (!have_cur && !have_prev) break;

        // Two evaluation jobs per iteration, ONE copy of the code (see Tables):
        //   job 0: pass A of the piece of this turn against guessed entry states, then publish;
        //   job 1: chain the piece of the previous turn (look-back); if it was entered in another
        //          state than guessed (rare), repair its pass A against the true state.
#pragma unroll 1
        for (u32 job = 0; job < 2; ++job) {
            if (job == 0 ? !have_cur : !have_prev) continue;
            const u32 t = turn - job;
            const u32 b3 = t % 3, b2 = t & 1;
            const u32 jpiece = job == 0 ? piece : ppiece;
            const StreamView &V = BATCH ? sm.view[b3] : p.one;
            const u32 lp = jpiece - V.first_piece;
            const u32 real_segs = real_segments<K>(V, lp);
            bool first = job == 0;
            u32 true_entry = 0;

            if (job == 1) {
                // ------------------------------------------------------------ look-back
                if (warp == 0) {
                    const u32 A = sm.wa[b3][0], X = sm.wx[b3][real_segs - 1];
                    u32 T = lane < (u32)kWarps ? sm.wt[b3][lane] : 0u;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) T += __shfl_xor_sync(0xffffffffu, T, d);
                    u32 astar = A;
                    u64 base = 0;
                    if (lp != 0) {
                        // warp-wide look-back: lane i inspects piece k - i; windows overlap by one
                        // so that every traversed piece sees the exit state of its predecessor.
                        int k = (int)jpiece - 1;
                        u64 acc = 0;
                        bool fresh = true;
                        u32 carried = 0;   // exit state assumed for the overlap piece by the previous window
                        while (true) {
                            const int idx = k - (int)lane;
                            const bool in = idx >= (int)V.first_piece;
                            u64 G = 0, I = 0;
                            if (in) {
                                I = ld_acquire_u64(&p.desc[idx].incl);
                                G = ld_acquire_u64(&p.desc[idx].agg);
                            }
                            const bool has_incl = (I & kValid) != 0;
                            const u32 incl_mask = __ballot_sync(0xffffffffu, has_incl);
                            const u32 pl = incl_mask ? (u32)__ffs(incl_mask) - 1 : 32u;
                            const bool ready = !in || has_incl || (G & kValid) != 0;
                            const u32 ready_mask = __ballot_sync(0xffffffffu, ready);
                            const u32 need = pl >= 32 ? 0xffffffffu : ((1u << pl) - 1);
                            if ((ready_mask & need) != need) {
                                __nanosleep(200);
                                continue;
                            }
                            const u32 prov = has_incl ? (u32)(I >> 56) & 0xfu : (u32)(G >> 48) & 0xfu;
                            if (!fresh && __shfl_sync(0xffffffffu, prov, 0) != carried) {
                                // the overlap piece left in another state than assumed: start over
                                k = (int)jpiece - 1;
                                acc = 0;
                                fresh = true;
                                continue;
                            }
                            const u32 pin = __shfl_down_sync(0xffffffffu, prov, 1);   // exit state of the predecessor
                            const u32 nl = min(pl, 31u);   // lanes [0, nl) are traversed in this window
                            bool link = true;
                            u32 c = 0;
                            if (lane < nl) {
                                link = ((u32)(G >> 56) & 0xfu) == pin;
                                c = (u32)G;
                            }
                            if (!__all_sync(0xffffffffu, link)) {
                                __nanosleep(400);   // a piece on the way was entered in another state than
                                continue;           // it assumed: it will publish its own inclusive state
                            }
                            u64 sum = c;
#pragma unroll
                            for (int dd = 16; dd > 0; dd >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, dd);
                            acc += sum;
                            if (fresh) {
                                astar = __shfl_sync(0xffffffffu, prov, 0);
                                fresh = false;
                            }
                            if (pl < 32) {
                                base = (__shfl_sync(0xffffffffu, I, pl) & kCountMask) + acc;
                                break;
                            }
                            carried = __shfl_sync(0xffffffffu, prov, 31);
                            k -= 31;
                        }
                        if (lane == 0 && astar == A)
                            st_release_u64(&p.desc[jpiece].incl, kValid | ((u64)X << 56) | (base + (u64)T));
                    }
                    if (lane == 0) {
                        sm.base = base;
                        sm.true_entry = astar;
                        sm.redo = astar != A;
                    }
                }
                __syncthreads();
                if (!sm.redo) continue;
                true_entry = sm.true_entry;
            }

            // ---------------------------------------------------------------- evaluation rounds
            // (Re)run pass A where a segment was entered in another state than its predecessor
            // left.  Normally (job 0): one pass A per warp, one check, no repair.
            {
                const Segment g = make_segment<K>(V, lp, warp);
                bool run = first;
                u32 my_entry = 0xffu;     // guess
                for (;;) {
                    if (run) {
                        const u64 r = segment_pass_a<K>(g, tb, my_entry, &sm.saved[b2][warp][0][0]);
                        if (lane == 0) {
                            sm.wa[b3][warp] = (u8)(r & 0xffu);
                            sm.wx[b3][warp] = (u8)((r >> 8) & 0xffu);
                            sm.wt[b3][warp] = (u32)(r >> 32);
                        }
                    }
                    __syncthreads();
                    if (first) {
                        first = false;
                        if (tid == 0) {      // claim the next piece (every thread has read sm.next_piece)
                            const u32 np = p.first_piece + atomicAdd(p.ticket, 1u);
                            sm.next_piece = np;
                            if (BATCH && np < p.num_pieces) sm.view[(t + 1) % 3] = p.streams[p.piece_stream[np]];
                        }
                    }
                    // first segment that was entered in another state than its predecessor left
                    const u32 entry0 = job == 1 ? true_entry : (u32)sm.wa[b3][0];
                    u32 wbad = kWarps, ebad = 0;
#pragma unroll
                    for (int w = kWarps - 1; w >= 0; --w) {
                        const u32 e_in = w == 0 ? entry0 : (u32)sm.wx[b3][w - 1];
                        if ((u32)w < real_segs && (u32)sm.wa[b3][w] != e_in) { wbad = (u32)w; ebad = e_in; }
                    }
                    if (wbad >= (u32)kWarps) break;
                    run = warp == wbad;
                    my_entry = ebad;
                    __syncthreads();         // everyone has read wa / wx before segment wbad rewrites them
                }
            }

            // ---------------------------------------------------------------- publish
            if (job == 0) {
                // the first piece of a stream is entered in state 0 exactly -> inclusive state;
                // any other piece -> aggregate under its guessed entry state
                if (warp == 0) {
                    const u32 A = sm.wa[b3][0], X = sm.wx[b3][real_segs - 1];
                    u32 T = lane < (u32)kWarps ? sm.wt[b3][lane] : 0u;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) T += __shfl_xor_sync(0xffffffffu, T, d);
                    if (lane == 0) {
                        if (lp == 0) st_release_u64(&p.desc[jpiece].incl, kValid | ((u64)X << 56) | (u64)T);
                        else st_release_u64(&p.desc[jpiece].agg, kValid | ((u64)A << 56) | ((u64)X << 48) | (u64)T);
                    }
                }
                // sm.next_piece was rewritten behind the barrier of the evaluation; the barrier
                // behind the look-back orders that against the read at the top of the loop --
                // except when there is no previous piece
                if (!have_prev) __syncthreads();
            } else if (tid == 0) {
                // the repaired piece's inclusive state comes late
                u64 T = 0;
                for (int w = 0; w < kWarps; ++w) T += sm.wt[b3][w];
                st_release_u64(&p.desc[jpiece].incl,
                               kValid | ((u64)sm.wx[b3][real_segs - 1] << 56) | (sm.base + T));
            }
        }

        // ================================================================ pass B of piece `turn - 1`
        if (have_prev) {
            const u32 pt = turn - 1, pb3 = pt % 3, pb2 = pt & 1;
            const StreamView &V = BATCH ? sm.view[pb3] : p.one;
            const u32 lp = ppiece - V.first_piece;
            u64 gstart = sm.base;                    // output index of my segment's first symbol
            for (u32 w = 0; w < warp; ++w) gstart += sm.wt[pb3][w];
            const Segment g = make_segment<K>(V, lp, warp);
            segment_pass_b<K>(g, tb, &sm.saved[pb2][warp][0][0], sm.stage[warp],
                              smem_base + (smem_addr)(sm.stage[warp] - smem_raw), V.out, V.n_out, gstart);
        }
        have_prev = have_cur;
        ppiece = piece;
    }
}

// ---------------------------------------------------------------------------------- host
// Kernel variants differ in K = warp-steps per segment, i.e. in the piece length
// (kWarps * K KiB of stream).  B200LC_CUHD_VARIANT pins one for tuning runs.
struct Variant {
    int K;
    void (*kern)(const DecodeParams);
    void (*kern_batch)(const DecodeParams);
    size_t smem_fixed;
};
#define B200LC_VARIANT(K_) \
    { K_, cuhd_decode_kernel<K_, false>, cuhd_decode_kernel<K_, true>, \
      ((sizeof(SmemLayout<K_>) + 127) & ~size_t(127)) }
static const Variant kVariants[] = {
    B200LC_VARIANT(8),    // 128 KiB pieces: default for long streams (16: 1.68 ms, 8: 1.60 ms, 4: 1.72 ms on C2)
    B200LC_VARIANT(4),    // shorter pieces for shorter streams (see pick_variant)
    B200LC_VARIANT(2),
    B200LC_VARIANT(1),
    B200LC_VARIANT(16),   // tuning points
    B200LC_VARIANT(32),
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);
constexpr int kSmallest = 3;   // index of the variant with the shortest pieces

static const Variant &variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("B200LC_CUHD_VARIANT");
        v = e ? atoi(e) : 0;
        if (v < 0 || v >= kNumVariants) v = 0;
    }
    return kVariants[v];
}

static u32 multi_bits(int L)
{
    static int lm = -1;
    if (lm < 0) {
        const char *e = getenv("B200LC_CUHD_MULTI_BITS");
        lm = e ? atoi(e) : (int)kMultiBits;
        if (lm < 1 || lm > 15) lm = (int)kMultiBits;
    }
    return (u32)(lm > L ? lm : L);
}

// dynamic shared memory: fixed layout + write table (4 B per entry) + counting tables (1 B each)
static size_t smem_bytes(const Variant &v, int L)
{
    return v.smem_fixed + (size_t(5) << multi_bits(L)) + (size_t(1) << L);
}

static u64 piece_units(const Variant &v) { return (u64)kWarps * v.K * 32 * S; }

static u32 pieces_for(const Variant &v, u64 n_units)
{
    return (u32)((n_units + piece_units(v) - 1) / piece_units(v));
}

// One-shot decodes pick the piece length by stream size: a piece is decoded by one CTA in two
// passes over K warp-steps, so a stream with fewer pieces than the GPU has CTA slots is
// latency-bound by its piece length.  Halve it until the pieces fill the machine.
static const Variant &pick_variant_for(u64 pieces_at_k1)
{
    if (getenv("B200LC_CUHD_VARIANT")) return variant();
    const u64 want = (u64)num_sms() * 2;
    for (int i = 0; i < kSmallest; ++i)
        if (pieces_at_k1 / (u64)kVariants[i].K >= want) return kVariants[i];
    return kVariants[kSmallest];
}
static const Variant &pick_variant(u64 n_units)
{
    return pick_variant_for(pieces_for(kVariants[kSmallest], n_units));
}

static StreamView make_view(const u32 *units, u64 n_units, u8 *out, u64 n_out, u32 first_piece)
{
    StreamView s;
    s.units = units;
    s.n_units = n_units;
    s.out = out;
    s.n_out = n_out;
    s.first_piece = first_piece;
    s.aligned = (reinterpret_cast<uintptr_t>(units) & 15) == 0 ? 1u : 0u;
    return s;
}

// The dynamic shared-memory limit is an attribute of the KERNEL (not of the table width it is
// launched with): raise it whenever a launch needs more than the largest value set so far on this
// device, and keep the occupancy that goes with each (kernel, table width).  Both are forgotten
// when the context epoch changes (cudaDeviceReset in resetGPU, culzss_api.cu).
struct KernelCache {
    unsigned epoch;
    size_t smem_set[kNumVariants][2];           // [variant][batch]
    int occ[kNumVariants][2][16];               // [variant][batch][L]
    size_t occ_smem[kNumVariants][2][16];
};
static KernelCache g_cache[kMaxDevices];

static int prepare_kernel(const Variant &v, bool batch, int L, size_t smem, int *occ_out)
{
    void (*const kern)(const DecodeParams) = batch ? v.kern_batch : v.kern;
    const int slot = device_slot();
    const int vi = (int)(&v - kVariants), bi = batch ? 1 : 0;
    KernelCache local = KernelCache();
    KernelCache &kc = slot >= 0 ? g_cache[slot] : local;
    if (kc.epoch != context_epoch()) {
        kc = KernelCache();
        kc.epoch = context_epoch();
    }
    if (smem > kc.smem_set[vi][bi]) {
        B200LC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kc.smem_set[vi][bi] = smem;
    }
    int occ = kc.occ_smem[vi][bi][L] == smem ? kc.occ[vi][bi][L] : 0;
    if (!occ) {
        B200LC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
        if (occ < 1) return B200LC_ERR_CUDA;
        kc.occ[vi][bi][L] = occ;
        kc.occ_smem[vi][bi][L] = smem;
    }
    *occ_out = occ;
    return B200LC_OK;
}

// batch: stream number of every piece (views are sorted by first_piece)
__global__ void piece_stream_kernel(const StreamView *__restrict__ views, u32 n_views, u32 n_pieces,
                                    u32 *__restrict__ piece_stream)
{
    const u32 piece = blockIdx.x * blockDim.x + threadIdx.x;
    if (piece >= n_pieces) return;
    u32 lo = 0, hi = n_views;             // last view with first_piece <= piece
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (views[mid].first_piece <= piece) lo = mid; else hi = mid;
    }
    piece_stream[piece] = lo;
}

}  // namespace cuhd
}  // namespace b200lc

using namespace b200lc;

extern "C" size_t b200lc_cuhd_decode_scratch_bytes(size_t n_units)
{
    // sized for the variant with the smallest pieces so that the answer does not depend on tuning
    return 128 + (size_t)cuhd::pieces_for(cuhd::kVariants[cuhd::kSmallest], n_units) * sizeof(cuhd::PieceDesc) + 128;
}

// Decodes pieces [first_piece, end_piece) of the stream.  The descriptors of earlier pieces must
// still be in d_scratch (first_piece == 0 clears them); units up to the end of the last piece
// + 4 must be resident.
static int decode_pieces(const cuhd::Variant &v, const uint32_t *d_units, size_t n_units,
                         uint8_t *d_out, size_t n_out, const void *d_table, int max_codeword_length,
                         void *d_scratch, size_t scratch_bytes, size_t first_piece, size_t end_piece,
                         cudaStream_t stream)
{
    if (max_codeword_length < 1 || max_codeword_length > 13) return B200LC_ERR_UNSUPPORTED;
    if (n_out == 0 || n_units == 0) return B200LC_OK;
    if (!d_units || !d_out || !d_table || !d_scratch) return B200LC_ERR_ARG;
    if (n_units >= (1ull << 40)) return B200LC_ERR_UNSUPPORTED;
    const size_t need = b200lc_cuhd_decode_scratch_bytes(n_units);
    if (scratch_bytes < need) return B200LC_ERR_SCRATCH;
    if (reinterpret_cast<uintptr_t>(d_scratch) & 127) return B200LC_ERR_ARG;

    const size_t smem = cuhd::smem_bytes(v, max_codeword_length);
    int occ = 0;
    const int rc = cuhd::prepare_kernel(v, false, max_codeword_length, smem, &occ);
    if (rc) return rc;
    cuhd::DecodeParams p;
    p.one = cuhd::make_view(d_units, n_units, d_out, n_out, 0);
    p.streams = nullptr;
    p.piece_stream = nullptr;
    p.lut = reinterpret_cast<const u16 *>(d_table);
    p.max_len = (u32)max_codeword_length;
    p.multi_bits = cuhd::multi_bits(max_codeword_length);
    p.ticket = reinterpret_cast<u32 *>(d_scratch);
    p.desc = reinterpret_cast<cuhd::PieceDesc *>(reinterpret_cast<char *>(d_scratch) + 128);
    const u32 all_pieces = cuhd::pieces_for(v, n_units);
    if (end_piece > all_pieces) end_piece = all_pieces;
    if (first_piece >= end_piece) return B200LC_OK;
    p.num_pieces = (u32)end_piece;
    p.first_piece = (u32)first_piece;

    if (first_piece == 0)
        B200LC_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, 128 + all_pieces * sizeof(cuhd::PieceDesc), stream));
    else
        B200LC_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, 128, stream));   // ticket only
    const u32 grid = (u32)min((u64)(end_piece - first_piece), (u64)num_sms() * (u64)occ);
    v.kern<<<grid, cuhd::kThreads, smem, stream>>>(p);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

extern "C" int b200lc_cuhd_decode(const uint32_t *d_units, size_t n_units, uint8_t *d_out,
                                  size_t n_out, const void *d_table, int max_codeword_length,
                                  void *d_scratch, size_t scratch_bytes, void *stream_)
{
    return decode_pieces(cuhd::pick_variant(n_units), d_units, n_units, d_out, n_out, d_table,
                         max_codeword_length, d_scratch, scratch_bytes, 0, ~size_t(0), (cudaStream_t)stream_);
}

extern "C" size_t b200lc_cuhd_decode_piece_units(void)
{
    return (size_t)cuhd::piece_units(cuhd::variant());
}

extern "C" int b200lc_cuhd_decode_pieces(const uint32_t *d_units, size_t n_units, uint8_t *d_out,
                                         size_t n_out, const void *d_table, int max_codeword_length,
                                         void *d_scratch, size_t scratch_bytes, size_t first_piece,
                                         size_t end_piece, void *stream_)
{
    return decode_pieces(cuhd::variant(), d_units, n_units, d_out, n_out, d_table, max_codeword_length,
                         d_scratch, scratch_bytes, first_piece, end_piece, (cudaStream_t)stream_);
}

// Asynchronous: copies the number of output symbols that are final once pieces [0, end_piece) are
// decoded (may exceed n_out on the last piece: padding bits) into *h_symbols (pinned host memory).
extern "C" int b200lc_cuhd_decode_progress_async(const void *d_scratch, size_t end_piece,
                                                 uint64_t *h_symbols, void *stream_)
{
    if (!d_scratch || !h_symbols || end_piece == 0) return B200LC_ERR_ARG;
    const cuhd::PieceDesc *desc =
        reinterpret_cast<const cuhd::PieceDesc *>(reinterpret_cast<const char *>(d_scratch) + 128);
    B200LC_CUDA_TRY(cudaMemcpyAsync(h_symbols, &desc[end_piece - 1].incl, sizeof(u64),
                                    cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
    return B200LC_OK;
}

// ------------------------------------------------------------------------------------ batch
// Many independent streams that share one code table, decoded by ONE launch: pieces of all
// streams are handed out through the same ticket counter and the look-back of a piece stops at
// the first piece of its stream (the same idea as the block segments of the radix sort).
namespace {
struct BatchPlan {
    const cuhd::Variant *v;
    std::vector<cuhd::StreamView> views;   // non-empty streams only
    u64 pieces;
    size_t desc_off, views_off, map_off, total;
};

static int plan_batch(const uint32_t *d_units, uint8_t *d_out, const b200lc_cuhd_stream *h, size_t n,
                      BatchPlan &bp)
{
    // piece length from the total size, like pick_variant does for one stream
    u64 small_pieces = 0;
    for (size_t i = 0; i < n; ++i)
        if (h[i].n_units && h[i].n_out) small_pieces += cuhd::pieces_for(cuhd::kVariants[cuhd::kSmallest], h[i].n_units);
    bp.v = &cuhd::pick_variant_for(small_pieces);
    bp.views.clear();
    bp.pieces = 0;
    for (size_t i = 0; i < n; ++i) {
        if (h[i].n_units == 0 || h[i].n_out == 0) continue;
        if (h[i].n_units >= (1ull << 40)) return B200LC_ERR_UNSUPPORTED;
        bp.views.push_back(cuhd::make_view(d_units + h[i].unit_offset, h[i].n_units,
                                           d_out + h[i].out_offset, h[i].n_out, (u32)bp.pieces));
        bp.pieces += cuhd::pieces_for(*bp.v, h[i].n_units);
    }
    if (bp.pieces >= (1ull << 31)) return B200LC_ERR_UNSUPPORTED;
    auto up = [](size_t x) { return (x + 127) & ~size_t(127); };
    bp.desc_off = 128;
    bp.views_off = bp.desc_off + up(bp.pieces * sizeof(cuhd::PieceDesc));
    bp.map_off = bp.views_off + up(bp.views.size() * sizeof(cuhd::StreamView));
    bp.total = bp.map_off + up(bp.pieces * 4);
    return B200LC_OK;
}
}  // namespace

extern "C" size_t b200lc_cuhd_decode_batch_scratch_bytes(const b200lc_cuhd_stream *h_streams, size_t n_streams)
{
    if (!h_streams) return 0;
    // sized for the variant with the shortest pieces so that the answer does not depend on tuning
    const cuhd::Variant &v = cuhd::kVariants[cuhd::kSmallest];
    u64 pieces = 0;
    for (size_t i = 0; i < n_streams; ++i) pieces += cuhd::pieces_for(v, h_streams[i].n_units);
    auto up = [](size_t x) { return (x + 127) & ~size_t(127); };
    return 128 + up(pieces * sizeof(cuhd::PieceDesc)) + up(n_streams * sizeof(cuhd::StreamView)) + up(pieces * 4) + 256;
}

extern "C" int b200lc_cuhd_decode_batch(const uint32_t *d_units, uint8_t *d_out,
                                        const b200lc_cuhd_stream *h_streams, size_t n_streams,
                                        const void *d_table, int max_codeword_length, void *d_scratch,
                                        size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (max_codeword_length < 1 || max_codeword_length > 13) return B200LC_ERR_UNSUPPORTED;
    if (n_streams == 0) return B200LC_OK;
    if (!d_units || !d_out || !h_streams || !d_table || !d_scratch) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_scratch) & 127) return B200LC_ERR_ARG;
    BatchPlan bp;
    int rc = plan_batch(d_units, d_out, h_streams, n_streams, bp);
    if (rc) return rc;
    if (bp.pieces == 0) return B200LC_OK;
    if (scratch_bytes < bp.total) return B200LC_ERR_SCRATCH;
    const cuhd::Variant &v = *bp.v;
    const size_t smem = cuhd::smem_bytes(v, max_codeword_length);
    int occ = 0;
    rc = cuhd::prepare_kernel(v, true, max_codeword_length, smem, &occ);
    if (rc) return rc;
    char *base = reinterpret_cast<char *>(d_scratch);
    cuhd::StreamView *d_views = reinterpret_cast<cuhd::StreamView *>(base + bp.views_off);
    u32 *d_map = reinterpret_cast<u32 *>(base + bp.map_off);
    B200LC_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, bp.views_off, stream));      // ticket + descriptors
    // the views come from pageable host memory: this copy returns once they have been staged
    B200LC_CUDA_TRY(cudaMemcpyAsync(d_views, bp.views.data(), bp.views.size() * sizeof(cuhd::StreamView),
                                    cudaMemcpyHostToDevice, stream));
    cuhd::piece_stream_kernel<<<(u32)((bp.pieces + 255) / 256), 256, 0, stream>>>(d_views, (u32)bp.views.size(),
                                                                                  (u32)bp.pieces, d_map);
    cuhd::DecodeParams p;
    p.one = bp.views[0];
    p.streams = d_views;
    p.piece_stream = d_map;
    p.lut = reinterpret_cast<const u16 *>(d_table);
    p.max_len = (u32)max_codeword_length;
    p.multi_bits = cuhd::multi_bits(max_codeword_length);
    p.ticket = reinterpret_cast<u32 *>(d_scratch);
    p.desc = reinterpret_cast<cuhd::PieceDesc *>(base + bp.desc_off);
    p.num_pieces = (u32)bp.pieces;
    p.first_piece = 0;
    const u32 grid = (u32)min((u64)bp.pieces, (u64)num_sms() * (u64)occ);
    v.kern_batch<<<grid, cuhd::kThreads, smem, stream>>>(p);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}
