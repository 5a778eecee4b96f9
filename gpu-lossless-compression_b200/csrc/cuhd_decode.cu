// Self-synchronising canonical-Huffman decode for sm_100a  (hot path 3).
//
// Replaces the reference's 4 phases / 6 kernels / host do-while
// (cuhd-icpp/src/cuhd_gpu_decoder.cu:145-523) with ONE persistent kernel:
//
//   * the compressed stream is cut into subsequences of S units (S*32 bits) and tiles of T
//     subsequences; each CTA pulls tiles from a ticket counter and stages the next tile's units
//     into shared memory with a 1-D TMA bulk copy (cp.async.bulk + mbarrier) while it works on
//     the current one;
//   * round 0: every thread decodes its subsequence from bit 0 and records the bit positions of
//     all codeword starts it saw (S registers of mask) -- this replaces phase 1's "decode the
//     whole next subsequence again and compare the last codeword" (:188-231) by an exact
//     merge test against the mask, so resynchronisation costs a few symbols, not a subsequence;
//   * the tile's effect on the decoder state is published as a FUNCTION of the entry state
//     (bit offset 0..L-1 of the first codeword): end state and symbol count for every entry
//     state that provably merges.  Tiles are chained with a decoupled look-back over those
//     function tables, which replaces phase 2's host loop + D2H flag copies (:458-495) and
//     phase 3's three passes over the 16-byte sync points (:498-509);
//   * the write pass re-decodes from the now known entry state into a shared-memory staging
//     buffer and leaves with 16-byte coalesced stores instead of phase 4's per-thread byte
//     stores (:101-105).
//
// HBM traffic is therefore units-in + symbols-out + 64 B of descriptor per tile; the
// reference's 20 B of sync-point state per 16 B of input is gone.
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

constexpr u32 kMaxStates = 16;       // entry states 0..L-1, L <= 13 supported
constexpr int kAltMaxSub = 8;        // entry-state walks give up after this many subsequences
constexpr u64 kInclValid = 1ull << 63;
constexpr u64 kCountMask = (1ull << 56) - 1;

// One 64-byte descriptor per tile, three independently readable parts:
//   incl : bit 63 valid | bits 56..59 exit state | bits 0..55 symbols that start before the next tile
//   agg  : bit 63 valid | bits 56..59 exit state E of every entry state that merges |
//          bits 40..55 mask of entry states that merge | bits 0..31 symbols for entry state 0
//   d[a] : symbols for entry state a minus symbols for entry state 0 (valid if mask bit a)
// By construction every merging entry state leaves the tile in the same state E, so a chain of
// tiles is traversable from aggregates alone whenever E of tile t-1 is in the mask of tile t.
struct __align__(64) TileDesc {
    u64 incl;
    u64 agg;
    short d[kMaxStates];
    u8 pad[16];
};
static_assert(sizeof(TileDesc) == 64, "descriptor is 64 bytes");
constexpr u64 kAggValid = 1ull << 63;

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
    u32 num_subtiles;
    u32 tma_tiles;       // leading sub-tiles (+ 4 lookahead units) that one TMA bulk copy can fetch
    u32 pad;
};

struct DecodeParams {
    StreamView one;      // the stream of a single-stream call
    const StreamView *streams;    // batch: views in global memory (nullptr: single stream)
    const u32 *piece_stream;      // batch: stream number of every piece
    const u16 *lut;      // {u8 num_bits, u8 symbol} little-endian pairs
    u32 max_len;         // L
    TileDesc *desc;
    u32 *ticket;
    u32 num_pieces;      // pieces [first_piece, num_pieces) are decoded by this launch
    u32 first_piece;
};

// ---------------------------------------------------------------------------------- kernel
// Work decomposition: subsequence = S units (one thread), sub-tile = T subsequences (one TMA
// transfer, T*S*4 bytes), piece = NSUB sub-tiles handled by one CTA between two look-backs.
template <int S, int T, int NSUB, int CAP>
struct SmemLayout {
    static constexpr int kTileUnits = T * S + 4;  // + one 16-byte lookahead
    u32 in[2][kTileUnits];
    union {
        __align__(16) u8 stage[CAP + 16];   // pass B: output staging
        u32 masks[T * S];                   // pass A, sub-tile 0: codeword-start masks of the paths from bit 0
    };
    u32 pre[T + 1];          // sub-tile 0 only: exclusive prefix of resolved symbol counts
    u16 saved[NSUB][T];      // pass A result per subsequence: entry state << 12 | symbol count
    u32 sub_total[NSUB];     // pass A symbols per sub-tile
    u8 sub_entry[NSUB];      // pass A entry / exit state per sub-tile
    u8 sub_exit[NSUB];
    u32 warp_sums[T / 32];
    u8 end[T];               // resolved exit state of every subsequence of the current sub-tile
    u8 wexit[2][T / 32];     // exit state of each warp's last lane, double-buffered by resolution round
    u8 onpath[T];            // sub-tile 0 only: resolved path ends on the recorded path
    u64 bar[2];
    u64 base;
    StreamView view[2];      // stream of the current piece / of the piece being prefetched (by piece parity)
    u32 next_piece;
    u32 total;
    u32 astar;
    u32 known;
};

// BATCH = false: one stream, its view comes from the kernel parameters (warp-uniform values the
// compiler keeps in uniform registers -- reading the same numbers from shared memory instead cost
// 8 % on C2); BATCH = true: the view of each piece's stream is fetched from global memory into
// shared memory by the thread that claims the piece.
// VAR: opt-in tuning switches (0 = the kernels every round-1 GPU test ran).
//   bit 0 (B200LC_CUHD_PASSA=multi): pass A walks with multi-symbol 16-bit entries
//         (walk_record_multi) from a third table behind the byte table;
//   bit 1 (B200LC_CUHD_WRITE=2): write-table layout 2 + running store pointer (walk_write2).
// Each switch alone compiles to the default's 56 registers; both together take 67 (3 CTAs/SM
// instead of 4) -- a register cap on that instantiation has to wait for a GPU to check it: every
// way of attaching one (minimum-blocks hint, __maxnreg__, body as a device function) also changed
// the code of the default kernel.
template <int S, int T, int NSUB, int CAP, bool BATCH, int VAR = 0>
__global__ void __launch_bounds__(T + 32) cuhd_decode_kernel(const DecodeParams p)
{
    using Smem = SmemLayout<S, T, NSUB, CAP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    u32 *tab = reinterpret_cast<u32 *>(smem_raw + ((sizeof(Smem) + 127) & ~size_t(127)));
    u8 *ltab = reinterpret_cast<u8 *>(tab + (size_t(1) << p.max_len));   // lengths only: 1 byte per entry
    // PAM only: a third table behind ltab, u16 per entry, all whole codewords of the window

    constexpr bool PAM = (VAR & 1) != 0, W2 = (VAR & 2) != 0;
    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31;
    const bool worker = tid < T;
    const u32 L = p.max_len;
    const u32 shift = 32 - L;
    constexpr u32 kTileUnits = Smem::kTileUnits;
    constexpr u32 kTileBytes = kTileUnits * 4;

    // LUT -> shared memory: two-symbol entries for the write pass (see walk_write) + a byte table
    // of first-codeword lengths for the counting passes.  A zero-length entry (unused prefix of an
    // incomplete code) would stall the reference forever; it is mapped to length 1 here so that
    // garbage input still terminates.
    for (u32 i = tid; i < (1u << L); i += blockDim.x) {
        const u32 e0 = p.lut[i];
        u32 len0 = e0 & 0xffu;
        if (len0 == 0 || len0 > L) len0 = 1;
        const u32 e1 = p.lut[(i << len0) & ((1u << L) - 1)];
        u32 len1 = e1 & 0xffu;
        if (len1 == 0 || len1 > L) len1 = 1;
        u32 entry = (e0 >> 8) | (len0 << 16);
        if (len0 + len1 <= L) entry = (e0 >> 8) | (e1 & 0xff00u) | ((len0 + len1) << 16) | 0x80000000u;
        if constexpr (W2) entry = (entry & 0xffffu) | ((entry >> 31) << 16) | (((entry >> 16) & 0xffu) << 24);
        tab[i] = entry;
        ltab[i] = (u8)len0;
        if constexpr (PAM) reinterpret_cast<u16 *>(ltab + (size_t(1) << L))[i] = multi_entry(p.lut, i, L);
    }
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // the stream a piece belongs to
    auto view_of = [&](u32 piece) -> StreamView {
        return p.streams[p.piece_stream[piece]];
    };
    // sub-tile index g (inside its stream) -> can it be fetched by one TMA bulk copy?
    auto tma_ok = [&](const StreamView &v, u32 g) -> bool {
        return g < v.tma_tiles;
    };
    auto issue_load = [&](const StreamView &v, u32 g, u32 buf) {  // one thread
        if (g < v.num_subtiles && tma_ok(v, g)) {
            mbar_expect_tx(&sm.bar[buf], kTileBytes);
            tma_load_1d(sm.in[buf], v.units + (u64)g * (T * S), kTileBytes, &sm.bar[buf]);
        }
    };

    if (tid == 0) {
        const u32 t0 = p.first_piece + atomicAdd(p.ticket, 1u);
        sm.next_piece = t0;
        if (t0 < p.num_pieces) {
            if (BATCH) {
                sm.view[0] = view_of(t0);
                issue_load(sm.view[0], (t0 - sm.view[0].first_piece) * NSUB, 0);
            } else {
                issue_load(p.one, t0 * NSUB, 0);
            }
        }
    }
    __syncthreads();
    u32 turn = 0;               // pieces this CTA has started: its view is sm.view[turn & 1]

    u32 step = 0;               // buffer = step & 1
    u32 phase0 = 0, phase1 = 0;

    // wait for (or synchronously load) sub-tile g into buffer (step & 1)
    auto acquire_input = [&](const StreamView &v, u32 g) {
        const u32 buf = step & 1;
        if (tma_ok(v, g)) {
            if (buf == 0) { mbar_wait(&sm.bar[0], phase0); phase0 ^= 1; }
            else          { mbar_wait(&sm.bar[1], phase1); phase1 ^= 1; }
        } else {
            const u64 first = (u64)g * (T * S);
            for (u32 i = tid; i < kTileUnits; i += blockDim.x) {
                const u64 idx = first + i;
                sm.in[buf][i] = idx < v.n_units ? v.units[idx] : 0u;
            }
            fence_proxy_async();
            __syncthreads();
        }
    };

    while (true) {
        const u32 piece = sm.next_piece;
        if (piece >= p.num_pieces) break;
        const StreamView &V = BATCH ? sm.view[turn & 1] : p.one;
        const u32 lp = piece - V.first_piece;          // piece number inside its stream
        const u32 g0 = lp * NSUB;
        const u32 nsub = min((u32)NSUB, V.num_subtiles - g0);

        u32 u[S + 1], m[S];
        u32 e0 = 0, c0 = 0;
        u32 my_start = 0, my_end = 0, my_cnt = 0;
        u32 pre = 0;

        auto load_units = [&](u32 buf) {
            const uint4 *src = reinterpret_cast<const uint4 *>(&sm.in[buf][tid * S]);
#pragma unroll
            for (int q = 0; q < S / 4; ++q) {
                const uint4 v = src[q];
                u[4 * q + 0] = v.x; u[4 * q + 1] = v.y; u[4 * q + 2] = v.z; u[4 * q + 3] = v.w;
            }
            u[S] = sm.in[buf][tid * S + S];
        };
        // Round 0 + chain resolution of the current sub-tile for entry state `entry`.
        // Leaves my_start / my_end / my_cnt per worker and sm.end[] = resolved exit states.
        // Only the first `real` subsequences hold stream units; the rest of the last sub-tile is
        // zero padding whose symbols lie beyond n_out.  Their entry states are not resolved: a
        // run of zeros never re-synchronises when the all-zero codeword is longer than one bit,
        // and the chain would crawl through it one subsequence per round.
        auto resolve_subtile = [&](u32 entry, bool keep_masks, u32 real) {
            if (worker) {
                if constexpr (PAM) walk_record_multi<S>(u, reinterpret_cast<const u16 *>(ltab + (size_t(1) << L)), shift, m, e0, c0);
                else walk_record<S>(u, ltab, shift, m, e0, c0);
                sm.end[tid] = (u8)e0;
                if (keep_masks) {
#pragma unroll
                    for (int j = 0; j < S; ++j) sm.masks[tid * S + j] = m[j];
                }
            }
            if (worker && lane == 31) sm.wexit[1][tid >> 5] = (u8)e0;
            __syncthreads();   // tentative exit states (paths from bit 0) visible
            my_start = 0;
            my_end = e0;
            my_cnt = c0;
            // Fixed point of "entry state = exit state of the predecessor".  Inside a warp the
            // chain is followed with shuffles (no CTA barrier); between warps through wexit[]:
            // in round r a warp starts from what its left neighbour's last lane published in
            // round r-1 and publishes its own exit state for round r+1.  Another CTA round runs
            // only if some warp published a different state than before -- rare, because nearly
            // every subsequence ends on the recorded path whatever its entry.
            u32 evaluated = 0;   // entry state for which my_end / my_cnt currently hold
            for (u32 round = 0;; ++round) {
                bool pub_changed = false;
                if (worker) {
                    const u32 rd = (round + 1) & 1, wr = round & 1, wid = tid >> 5;
                    const u32 warp_in = tid == 0 ? entry : (lane == 0 ? (u32)sm.wexit[rd][wid - 1] : 0u);
                    while (true) {
                        u32 sv = __shfl_up_sync(0xffffffffu, my_end, 1);
                        if (lane == 0) sv = warp_in;
                        my_start = sv;
                        const bool eval = sv != evaluated && tid < real;
                        bool changed = false;
                        if (eval) {
                            u32 ne, nc;
                            walk_merge<S>(u, m, sv, e0, ltab, shift, ne, nc);
                            changed = ne != my_end;
                            my_end = ne;
                            my_cnt = nc;
                            evaluated = sv;
                        }
                        if (!__any_sync(0xffffffffu, changed)) break;
                    }
                    sm.end[tid] = (u8)my_end;
                    if (lane == 31) {
                        pub_changed = my_end != (u32)sm.wexit[rd][wid];
                        sm.wexit[wr][wid] = (u8)my_end;
                    }
                }
                if (!__syncthreads_or(pub_changed)) break;
            }
        };
        // exclusive scan of my_cnt over the workers -> pre, sm.total
        auto block_scan = [&]() {
            u32 incl = 0;
            if (worker) {
                incl = warp_incl_scan(my_cnt);
                if (lane == 31) sm.warp_sums[tid >> 5] = incl;
            }
            __syncthreads();
            if (tid < 32) {
                u32 v = tid < T / 32 ? sm.warp_sums[tid] : 0u;
                const u32 s = warp_incl_scan(v);
                if (tid < T / 32) sm.warp_sums[tid] = s - v;
                if (tid == T / 32 - 1) sm.total = s;
            }
            __syncthreads();
            if (worker) pre = sm.warp_sums[tid >> 5] + incl - my_cnt;
        };
        // prefetch the sub-tile of the NEXT step into the other buffer (one thread)
        auto prefetch = [&](u32 pass, u32 c) {
            if (tid != 0) return;
            if (c + 1 < nsub) issue_load(V, g0 + c + 1, (step & 1) ^ 1);
            else if (pass == 0) issue_load(V, g0, (step & 1) ^ 1);
            else {
                const u32 np = p.first_piece + atomicAdd(p.ticket, 1u);
                sm.next_piece = np;
                if (np >= p.num_pieces) return;
                if (BATCH) {
                    StreamView &vn = sm.view[(turn + 1) & 1];
                    vn = view_of(np);
                    issue_load(vn, (np - vn.first_piece) * NSUB, (step & 1) ^ 1);
                } else {
                    issue_load(p.one, np * NSUB, (step & 1) ^ 1);
                }
            }
        };

        // number of subsequences of sub-tile g that start inside the stream
        const u64 total_subseq = (V.n_units + S - 1) / S;
        auto real_subseq = [&](u32 g) -> u32 {
            const u64 s0 = (u64)g * T;
            return s0 >= total_subseq ? 0u : (u32)min((u64)T, total_subseq - s0);
        };

        // ================================================================ pass A: states + counts
        u32 entry = 0;
        u32 piece_total = 0;
        int dl = 0;            // alt warp: symbols(entry state lane) - symbols(entry state 0)
        bool known = false;
        for (u32 c = 0; c < nsub; ++c) {
            acquire_input(V, g0 + c);
            prefetch(0, c);
            const u32 buf = step & 1;
            if (worker) load_units(buf);
            resolve_subtile(entry, c == 0, real_subseq(g0 + c));
            block_scan();
            if (worker) {
                sm.saved[c][tid] = (u16)((my_start << 12) | my_cnt);
                if (c == 0) {
                    sm.pre[tid] = pre;
                    sm.onpath[tid] = my_end == e0;
                    if (tid == T - 1) sm.pre[T] = pre + my_cnt;
                }
                if (tid == T - 1) {
                    sm.sub_entry[c] = (u8)entry;
                    sm.sub_exit[c] = (u8)my_end;
                    sm.sub_total[c] = sm.total;
                }
            }
            piece_total += sm.total;
            entry = sm.end[T - 1];
            if (c == 0) {
                __syncthreads();
                // The piece as a function of its entry state `lane`: walk from bit `lane` of the
                // first subsequence until the walk lands on the resolved path (usually within a
                // few symbols); from there on the piece behaves as for entry state 0.
                if (!worker) {
                    if (lane == 0) {
                        known = true;
                    } else if (lane < L) {
                        u32 at = lane, own = 0;
                        for (u32 sq = 0; sq < (u32)min(T, kAltMaxSub); ++sq) {
                            u32 au[S + 1], am[S];
#pragma unroll
                            for (int j = 0; j <= S; ++j) au[j] = sm.in[buf][sq * S + j];
                            const bool onp = sm.onpath[sq];
#pragma unroll
                            for (int j = 0; j < S; ++j) am[j] = onp ? sm.masks[sq * S + j] : 0u;
                            const u32 rend = sm.end[sq];
                            u32 ne, nc;
                            walk_merge<S>(au, am, at, rend, ltab, shift, ne, nc);
                            own += nc;
                            if (ne == rend) {
                                dl = (int)own - (int)sm.pre[sq + 1];
                                known = dl >= -32768 && dl <= 32767;
                                break;
                            }
                            at = ne;
                        }
                    }
                }
            }
            __syncthreads();   // buffer hand-over, sm.end / sm.total reuse
            ++step;
        }

        // ================================================================ publish + look-back
        if (!worker) {
            const u32 total0 = piece_total;
            const u32 exit0 = entry;
            const u32 known_mask = __ballot_sync(0xffffffffu, known) & 0xffffu;
            TileDesc *d = &p.desc[piece];
            u32 astar = 0;
            u64 base = 0;
            if (lp == 0) {
                if (lane == 0) st_release_u64(&d->incl, kInclValid | ((u64)exit0 << 56) | total0);
            } else {
                if (lane < kMaxStates) d->d[lane] = (short)(known ? dl : 0);
                __threadfence();
                __syncwarp();
                if (lane == 0)
                    st_release_u64(&d->agg, kAggValid | ((u64)exit0 << 56) |
                                                ((u64)known_mask << 40) | total0);

                // warp-wide look-back: lane i inspects piece k - i; windows overlap by one so
                // that every traversed piece sees the exit state of its predecessor.
                int k = (int)piece - 1;
                u64 acc = 0;
                bool first = true;
                u32 carried = 0;   // exit state assumed for the overlap piece by the previous window
                while (true) {
                    const int idx = k - (int)lane;
                    u64 A = 0, I = 0;
                    if (idx >= (int)V.first_piece) {
                        I = ld_acquire_u64(&p.desc[idx].incl);
                        A = ld_acquire_u64(&p.desc[idx].agg);
                    }
                    const bool has_incl = (I & kInclValid) != 0;
                    const u32 incl_mask = __ballot_sync(0xffffffffu, has_incl);
                    const u32 pl = incl_mask ? (u32)__ffs(incl_mask) - 1 : 32u;
                    const bool agg_ok = idx < (int)V.first_piece || (A & kAggValid) != 0;
                    const u32 agg_mask = __ballot_sync(0xffffffffu, agg_ok);
                    const u32 need = pl >= 32 ? 0xffffffffu : ((1u << pl) - 1);
                    if ((agg_mask & need) != need) {
                        __nanosleep(100);
                        continue;
                    }
                    const u32 prov = has_incl ? (u32)(I >> 56) & 0xfu : (u32)(A >> 56) & 0xfu;
                    if (!first && __shfl_sync(0xffffffffu, prov, 0) != carried) {
                        // the overlap piece left in another state than assumed: start over
                        k = (int)piece - 1;
                        acc = 0;
                        first = true;
                        continue;
                    }
                    const u32 in = __shfl_down_sync(0xffffffffu, prov, 1);
                    const u32 nl = min(pl, 31u);   // lanes [0, nl) are traversed in this window
                    bool ok = true;
                    u32 c = 0;
                    if (lane < nl) {
                        ok = ((u32)(A >> 40) >> in) & 1u;
                        if (ok) c = (u32)((int)(u32)A + (int)__ldcg(&p.desc[idx].d[in]));
                    }
                    if (!__all_sync(0xffffffffu, ok)) {
                        __nanosleep(200);   // a piece on the way must publish its own inclusive state
                        continue;
                    }
                    u64 sum = c;
#pragma unroll
                    for (int dd = 16; dd > 0; dd >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, dd);
                    acc += sum;
                    if (first) {
                        astar = __shfl_sync(0xffffffffu, prov, 0);
                        first = false;
                    }
                    if (pl < 32) {
                        base = (__shfl_sync(0xffffffffu, I, pl) & kCountMask) + acc;
                        break;
                    }
                    carried = __shfl_sync(0xffffffffu, prov, 31);
                    k -= 31;
                }
                // inclusive state of this piece, if its entry state merges
                const bool my_known = (known_mask >> astar) & 1u;
                const int my_d = __shfl_sync(0xffffffffu, dl, astar);
                if (lane == 0 && my_known)
                    st_release_u64(&d->incl, kInclValid | ((u64)exit0 << 56) |
                                                 (base + (u64)((int)total0 + my_d)));
                if (lane == 0) sm.known = my_known;
            }
            if (lane == 0) {
                sm.astar = astar;
                sm.base = base;
                if (lp == 0) sm.known = 1;
            }
        }
        __syncthreads();

        // ================================================================ pass B: decode + write
        u32 entry_true = sm.astar;
        u64 base_run = sm.base;
        const bool publish_late = !sm.known;
        for (u32 c = 0; c < nsub; ++c) {
            acquire_input(V, g0 + c);
            prefetch(1, c);
            const u32 buf = step & 1;
            if (worker) load_units(buf);
            u32 exit_state;
            if (entry_true != sm.sub_entry[c]) {
                // entry state differs from what pass A assumed (always for sub-tile 0 when the
                // piece's entry state is not 0): redo states + counts for this sub-tile
                resolve_subtile(entry_true, false, real_subseq(g0 + c));
                exit_state = sm.end[T - 1];
            } else {
                if (worker) {
                    const u32 sv = sm.saved[c][tid];
                    my_start = sv >> 12;
                    my_cnt = sv & 0xfffu;
                }
                exit_state = sm.sub_exit[c];
            }
            block_scan();
            const u32 total = sm.total;

            u64 tile_cnt = 0;
            if (base_run < V.n_out) tile_cnt = min((u64)total, V.n_out - base_run);
            for (u32 w0 = 0; w0 < tile_cnt; w0 += CAP) {
                const u32 wlen = (u32)min((u64)CAP, tile_cnt - w0);
                u8 *g = V.out + base_run + w0;
                const u32 sh = (u32)(reinterpret_cast<uintptr_t>(g) & 15u);
                if (worker) {
                    const u32 lo = w0, hi = w0 + wlen;
                    if (pre < hi && pre + my_cnt > lo) {
                        u8 *dst = sm.stage + ((int)sh - (int)w0);
                        if constexpr (W2) {
                            if (pre >= lo && pre + my_cnt <= hi)
                                walk_write2<S, false>(u, tab, shift, my_start, dst, pre, lo, hi);
                            else
                                walk_write2<S, true>(u, tab, shift, my_start, dst, pre, lo, hi);
                        } else if (pre >= lo && pre + my_cnt <= hi)
                            walk_write<S, false>(u, tab, shift, my_start, dst, pre, lo, hi);
                        else
                            walk_write<S, true>(u, tab, shift, my_start, dst, pre, lo, hi);
                    }
                }
                __syncthreads();
                const u32 head = min(wlen, (16u - sh) & 15u);
                const u32 nvec = (wlen - head) >> 4;
                const u32 tail0 = head + (nvec << 4);
                if (tid < head) g[tid] = sm.stage[sh + tid];
                const uint4 *sv = reinterpret_cast<const uint4 *>(sm.stage + sh + head);
                uint4 *gv = reinterpret_cast<uint4 *>(g + head);
                for (u32 i = tid; i < nvec; i += blockDim.x) gv[i] = sv[i];
                if (tid < wlen - tail0) g[tail0 + tid] = sm.stage[sh + tail0 + tid];
                __syncthreads();
            }
            base_run += total;
            entry_true = exit_state;
            __syncthreads();   // buffer hand-over, sm.total / sm.end reuse
            ++step;
        }
        if (publish_late && tid == 0)
            st_release_u64(&p.desc[piece].incl,
                           kInclValid | ((u64)entry_true << 56) | base_run);
        __syncthreads();
        ++turn;
    }
}

// ---------------------------------------------------------------------------------- host
// Kernel variants (subsequence units S, subsequences per sub-tile T, sub-tiles per piece NSUB,
// staging bytes CAP).  Variant 0 is the default; B200LC_CUHD_VARIANT selects another one for
// tuning runs.
struct Variant {
    int S, T, NSUB, CAP;
    void (*kern[4])(const DecodeParams);          // indexed by the opt-in switches VAR (0 = default)
    void (*kern_batch[4])(const DecodeParams);
    size_t smem_fixed;
};
#define B200LC_VARIANT(S_, T_, N_, C_) \
    { S_, T_, N_, C_, \
      { cuhd_decode_kernel<S_, T_, N_, C_, false, 0>, cuhd_decode_kernel<S_, T_, N_, C_, false, 1>, \
        cuhd_decode_kernel<S_, T_, N_, C_, false, 2>, cuhd_decode_kernel<S_, T_, N_, C_, false, 3> }, \
      { cuhd_decode_kernel<S_, T_, N_, C_, true, 0>, cuhd_decode_kernel<S_, T_, N_, C_, true, 1>, \
        cuhd_decode_kernel<S_, T_, N_, C_, true, 2>, cuhd_decode_kernel<S_, T_, N_, C_, true, 3> }, \
      ((sizeof(SmemLayout<S_, T_, N_, C_>) + 127) & ~size_t(127)) }
static const Variant kVariants[] = {
    B200LC_VARIANT(8, 256, 16, 16384),   // default for long streams: 407 GB/s of output on C2 (B200, round 1)
    B200LC_VARIANT(8, 256, 8, 16384),    // shorter pieces for shorter streams (see pick_variant)
    B200LC_VARIANT(8, 256, 4, 16384),
    B200LC_VARIANT(8, 256, 2, 16384),
    B200LC_VARIANT(8, 256, 1, 16384),
    B200LC_VARIANT(8, 256, 32, 16384),   // tuning points: 333
    B200LC_VARIANT(8, 128, 16, 8192),    // 375
    B200LC_VARIANT(4, 256, 32, 12288),   // 279
};
constexpr int kNumVariants = sizeof(kVariants) / sizeof(kVariants[0]);

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

// Opt-in tuning switches (template parameter VAR of the kernel): B200LC_CUHD_PASSA=multi -> bit 0,
// B200LC_CUHD_WRITE=2 -> bit 1.  Read once per process; the default is 0.
static int tuning_switches()
{
    static int v = -1;
    if (v < 0) {
        const char *a = getenv("B200LC_CUHD_PASSA"), *w = getenv("B200LC_CUHD_WRITE");
        v = ((a && a[0] == 'm') ? 1 : 0) | ((w && w[0] == '2') ? 2 : 0);
    }
    return v;
}
// dynamic shared memory: fixed layout + write table (4 B) + byte lengths (1 B) [+ multi table (2 B)]
static size_t smem_bytes(const Variant &v, int L, int var)
{
    return v.smem_fixed + (size_t((var & 1) ? 7 : 5) << L);
}

// One-shot decodes pick the piece length by stream size: a piece is decoded by one CTA, two
// passes over NSUB sub-tiles back to back (~5 us each), so a stream with fewer pieces than the
// GPU has CTA slots is latency-bound by its piece length.  Halve it until the pieces fill the
// machine (variants 0..4 differ only in NSUB).  B200LC_CUHD_VARIANT pins one variant.
static const Variant &pick_variant(u64 n_units)
{
    if (getenv("B200LC_CUHD_VARIANT")) return variant();
    const u64 want = (u64)num_sms() * 4;
    for (int i = 0; i < 4; ++i) {
        const Variant &v = kVariants[i];
        const u64 sub = ((n_units + v.S - 1) / v.S + v.T - 1) / v.T;
        if ((sub + v.NSUB - 1) / v.NSUB >= want) return v;
    }
    return kVariants[4];
}

static u32 subtiles_for(const Variant &v, u64 n_units)
{
    const u64 nsub = (n_units + v.S - 1) / v.S;
    return (u32)((nsub + v.T - 1) / v.T);
}
static u32 pieces_for(const Variant &v, u64 n_units)
{
    return (subtiles_for(v, n_units) + v.NSUB - 1) / v.NSUB;
}

static StreamView make_view(const Variant &v, const u32 *units, u64 n_units, u8 *out, u64 n_out, u32 first_piece)
{
    StreamView s;
    s.units = units;
    s.n_units = n_units;
    s.out = out;
    s.n_out = n_out;
    s.first_piece = first_piece;
    s.num_subtiles = subtiles_for(v, n_units);
    s.tma_tiles = 0;
    if ((reinterpret_cast<uintptr_t>(units) & 15) == 0 && n_units >= 4)
        s.tma_tiles = (u32)min((u64)s.num_subtiles, (u64)(n_units - 4) / (u64)(v.T * v.S));
    s.pad = 0;
    return s;
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
    size_t worst = 0;
    for (int i = 0; i < cuhd::kNumVariants; ++i) {
        const size_t n = cuhd::pieces_for(cuhd::kVariants[i], n_units);
        if (n > worst) worst = n;
    }
    return 128 + worst * sizeof(cuhd::TileDesc);
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

    const int var = cuhd::tuning_switches();
    void (*const kern)(const cuhd::DecodeParams) = v.kern[var];
    const size_t smem = cuhd::smem_bytes(v, max_codeword_length, var);
    static int occ_table[kMaxDevices][cuhd::kNumVariants][14] = {{{0}}};
    const int slot = device_slot();
    int occ = slot >= 0 ? occ_table[slot][&v - cuhd::kVariants][max_codeword_length] : 0;
    if (!occ) {
        B200LC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
        B200LC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, v.T + 32, smem));
        if (occ < 1) return B200LC_ERR_CUDA;
        if (slot >= 0) occ_table[slot][&v - cuhd::kVariants][max_codeword_length] = occ;
    }
    cuhd::DecodeParams p;
    p.one = cuhd::make_view(v, d_units, n_units, d_out, n_out, 0);
    p.streams = nullptr;
    p.piece_stream = nullptr;
    p.lut = reinterpret_cast<const u16 *>(d_table);
    p.max_len = (u32)max_codeword_length;
    p.ticket = reinterpret_cast<u32 *>(d_scratch);
    p.desc = reinterpret_cast<cuhd::TileDesc *>(reinterpret_cast<char *>(d_scratch) + 128);
    const u32 all_pieces = cuhd::pieces_for(v, n_units);
    if (end_piece > all_pieces) end_piece = all_pieces;
    if (first_piece >= end_piece) return B200LC_OK;
    p.num_pieces = (u32)end_piece;
    p.first_piece = (u32)first_piece;

    if (first_piece == 0)
        B200LC_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, 128 + all_pieces * sizeof(cuhd::TileDesc), stream));
    else
        B200LC_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, 128, stream));   // ticket only
    const u32 grid = (u32)min((u64)(end_piece - first_piece),
                              (u64)num_sms() * (u64)occ);
    kern<<<grid, v.T + 32, smem, stream>>>(p);
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
    const cuhd::Variant &v = cuhd::variant();
    return (size_t)v.S * v.T * v.NSUB;
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
    const cuhd::TileDesc *desc =
        reinterpret_cast<const cuhd::TileDesc *>(reinterpret_cast<const char *>(d_scratch) + 128);
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
    bp.v = &cuhd::kVariants[4];
    if (getenv("B200LC_CUHD_VARIANT")) bp.v = &cuhd::variant();
    else {
        const u64 want = (u64)num_sms() * 4;
        for (int k = 0; k < 4; ++k) {
            const cuhd::Variant &v = cuhd::kVariants[k];
            u64 pieces = 0;
            for (size_t i = 0; i < n; ++i)
                if (h[i].n_units && h[i].n_out) pieces += cuhd::pieces_for(v, h[i].n_units);
            if (pieces >= want) { bp.v = &v; break; }
        }
    }
    bp.views.clear();
    bp.pieces = 0;
    for (size_t i = 0; i < n; ++i) {
        if (h[i].n_units == 0 || h[i].n_out == 0) continue;
        if (h[i].n_units >= (1ull << 40)) return B200LC_ERR_UNSUPPORTED;
        bp.views.push_back(cuhd::make_view(*bp.v, d_units + h[i].unit_offset, h[i].n_units,
                                           d_out + h[i].out_offset, h[i].n_out, (u32)bp.pieces));
        bp.pieces += cuhd::pieces_for(*bp.v, h[i].n_units);
    }
    if (bp.pieces >= (1ull << 31)) return B200LC_ERR_UNSUPPORTED;
    auto up = [](size_t x) { return (x + 127) & ~size_t(127); };
    bp.desc_off = 128;
    bp.views_off = bp.desc_off + up(bp.pieces * sizeof(cuhd::TileDesc));
    bp.map_off = bp.views_off + up(bp.views.size() * sizeof(cuhd::StreamView));
    bp.total = bp.map_off + up(bp.pieces * 4);
    return B200LC_OK;
}
}  // namespace

extern "C" size_t b200lc_cuhd_decode_batch_scratch_bytes(const b200lc_cuhd_stream *h_streams, size_t n_streams)
{
    if (!h_streams) return 0;
    // sized for the variant with the shortest pieces so that the answer does not depend on tuning
    const cuhd::Variant &v = cuhd::kVariants[4];
    u64 pieces = 0;
    for (size_t i = 0; i < n_streams; ++i) pieces += cuhd::pieces_for(v, h_streams[i].n_units);
    auto up = [](size_t x) { return (x + 127) & ~size_t(127); };
    return 128 + up(pieces * sizeof(cuhd::TileDesc)) + up(n_streams * sizeof(cuhd::StreamView)) + up(pieces * 4) + 256;
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
    const int var = cuhd::tuning_switches();
    void (*const kern_batch)(const cuhd::DecodeParams) = v.kern_batch[var];
    const size_t smem = cuhd::smem_bytes(v, max_codeword_length, var);
    B200LC_CUDA_TRY(cudaFuncSetAttribute(kern_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    B200LC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern_batch, v.T + 32, smem));
    if (occ < 1) return B200LC_ERR_CUDA;
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
    p.ticket = reinterpret_cast<u32 *>(d_scratch);
    p.desc = reinterpret_cast<cuhd::TileDesc *>(base + bp.desc_off);
    p.num_pieces = (u32)bp.pieces;
    p.first_piece = 0;
    const u32 grid = (u32)min((u64)bp.pieces, (u64)num_sms() * (u64)occ);
    kern_batch<<<grid, v.T + 32, smem, stream>>>(p);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}
