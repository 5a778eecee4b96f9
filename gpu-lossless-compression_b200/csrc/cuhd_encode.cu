// GPU Huffman encoder producing CUHD-format streams + byte histogram  (SURVEY.md 8f N1).
//
// Replaces the sequential CPU packer llhuff::LLHuffmanEncoder::encode_memory
// (cuhd-icpp/encoder/src/llhuffman_encoder.cc:200-238; 1.27 s per 100 MiB in the reference's
// README) and the frequency count at :23-26.  Given the same dictionary the produced units are
// bit-identical to the reference's wherever the reference's output is defined (all units but the
// unused low bits of the last one); the final partial unit is
// zero-filled and always flushed.
//
// One persistent kernel, one pass over the input:
//   * tiles of kThreads*kSyms symbols are staged into shared memory by 1-D TMA bulk copies
//     (double buffered, ticket-ordered);
//   * each thread looks its 16 symbols up once (code|len kept in registers), a block scan gives
//     bit offsets inside the tile, a decoupled look-back over per-tile bit counts gives the
//     tile's global bit offset -- this is the "prefix-sum bit-pack" of the north star;
//   * the (<32) bits that precede the tile inside its first output word come from the
//     predecessor's published 31-bit tail, so every output word is written exactly once, with
//     plain 128-bit stores, and the output needs no zero-initialisation and no global atomics.
#include "common.cuh"
#include "../../include/b200lc.h"

namespace b200lc {
namespace cuhd_enc {

constexpr int kThreads = 512;
constexpr int kSyms = 16;                      // symbols per thread
constexpr int kTileSyms = kThreads * kSyms;    // 8192 symbols per sub-tile (one TMA transfer)
constexpr int kNSub = 16;                      // sub-tiles per piece (128 KiB, one look-back)
constexpr int kPieceSyms = kTileSyms * kNSub;
constexpr int kMaxLen = 13;
constexpr int kStageWords = kTileSyms * kMaxLen / 32 + 8;

// Piece descriptor: two independently published 64-bit words.
//   agg : bit 63 valid | bits 31..62 piece bit count | bits 0..30 last 31 bits of the piece
//   incl: bit 63 valid | bits 0..62 bit count of pieces 0..t
struct __align__(16) EncDesc {
    u64 agg;
    u64 incl;
};
constexpr u64 kValid = 1ull << 63;

struct EncParams {
    const u8 *in;
    u64 n;
    const u32 *code_of_symbol;  // [256]
    const u8 *len_of_symbol;    // [256]
    u32 *out;
    u64 out_cap_units;
    u64 *total_bits;            // device scalar, written by the last piece
    u32 *overflow;              // device flag, set if out_cap_units was too small
    EncDesc *desc;
    u32 *ticket;
    u32 num_pieces;
    // Blocks: the input is cut into blocks of block_syms symbols (one block = the whole input for
    // a plain call), each packed as an independent stream that starts at bit 0 of
    // out + block * unit_stride.  Pieces never straddle blocks; a block has pieces_per_block of them.
    u64 block_syms;
    u64 unit_stride;
    u32 pieces_per_block;
    u32 aligned;         // input base and block_syms are multiples of 16: TMA bulk copies allowed
    u64 *block_bits;     // [blocks] stream bits of every block (nullptr for a plain call)
    const u64 *base_bits;   // PLANNED: bit offset of every piece, known before the launch
};

// What a piece needs to know about its block.
struct BlockView {
    const u8 *in;        // first symbol of the block
    u64 n;               // symbols in the block
    u32 *out;
    u32 block, lp;       // block number, piece number inside the block
    u32 num_subtiles, tma_tiles, last_lp;
};
__device__ __forceinline__ BlockView block_view(const EncParams &p, u32 piece)
{
    BlockView v;
    v.block = piece / p.pieces_per_block;
    v.lp = piece - v.block * p.pieces_per_block;
    const u64 first = (u64)v.block * p.block_syms;
    v.in = p.in + first;
    v.n = min(p.block_syms, p.n - first);
    v.out = p.out + (u64)v.block * p.unit_stride;
    v.num_subtiles = (u32)((v.n + kTileSyms - 1) / kTileSyms);
    v.tma_tiles = p.aligned ? (u32)(v.n / kTileSyms) : 0u;
    v.last_lp = (v.num_subtiles + kNSub - 1) / kNSub - 1;
    return v;
}

struct EncSmem {
    __align__(16) u8 in[2][kTileSyms];
    __align__(16) u32 stage[kStageWords];
    u32 tab[256];      // code | len << 16
    u8 len8[256];
    u32 warp_sums[kThreads / 32];
    u64 bar[2];
    u64 base_bits;
    u64 piece_bits;
    u32 next_piece;
    u32 total;
    u32 prev_tail;
    u64 tail_acc[2];   // packed codes of the piece's last two threads (low 64 bits)
    u32 tail_bits[2];
};

// PLANNED: the bit offset of every piece is an input (computed from per-piece histograms and the
// code lengths, see b200lc_cuhd_encode_planned), so the counting pass and the look-back disappear.
// Without the predecessor's tail a piece cannot complete the word it shares with its neighbour:
// both sides OR their bits into that word, which the planning kernel has zeroed.
template <bool PLANNED>
__global__ void __launch_bounds__(kThreads) cuhd_encode_kernel(const EncParams p)
{
    __shared__ EncSmem sm;
    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31;

    if (tid < 256) {
        // the staging buffer and the two-symbols-per-step pack are sized for codes of <= 13 bits: a
        // longer entry in a caller-supplied dictionary is reported through the overflow flag and
        // clamped (the output of such a call is not a valid stream) instead of overrunning
        u32 len = p.len_of_symbol[tid];
        if (len > 13) {
            atomicExch(p.overflow, 1u);
            len = 13;
        }
        sm.tab[tid] = (p.code_of_symbol[tid] & ((1u << len) - 1u)) | (len << 16);
        sm.len8[tid] = (u8)len;
    }
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto tma_ok = [&](const BlockView &v, u32 g) -> bool {
        return g < v.tma_tiles;
    };
    auto issue_load = [&](const BlockView &v, u32 g, u32 buf) {
        if (g < v.num_subtiles && tma_ok(v, g)) {
            mbar_expect_tx(&sm.bar[buf], kTileSyms);
            tma_load_1d(sm.in[buf], v.in + (u64)g * kTileSyms, kTileSyms, &sm.bar[buf]);
        }
    };
    if (tid == 0) {
        const u32 t0 = atomicAdd(p.ticket, 1u);
        sm.next_piece = t0;
        if (t0 < p.num_pieces) {
            const BlockView v0 = block_view(p, t0);
            issue_load(v0, v0.lp * kNSub, 0);
        }
    }
    __syncthreads();

    u32 step = 0, phase0 = 0, phase1 = 0;
    auto acquire_input = [&](const BlockView &v, u32 g, u32 tile_n) {
        const u32 buf = step & 1;
        if (tma_ok(v, g)) {
            if (buf == 0) { mbar_wait(&sm.bar[0], phase0); phase0 ^= 1; }
            else          { mbar_wait(&sm.bar[1], phase1); phase1 ^= 1; }
        } else {
            const u64 first = (u64)g * kTileSyms;
            for (u32 i = tid; i < kTileSyms; i += kThreads)
                sm.in[buf][i] = i < tile_n ? v.in[first + i] : (u8)0;
            fence_proxy_async();
            __syncthreads();
        }
    };

    while (true) {
        const u32 piece = sm.next_piece;
        if (piece >= p.num_pieces) break;
        const BlockView V = block_view(p, piece);
        const u32 g0 = V.lp * kNSub;                 // sub-tile numbers are local to the block
        const u32 nsub = min((u32)kNSub, V.num_subtiles - g0);

        auto prefetch = [&](u32 pass, u32 c) {
            if (tid != 0) return;
            if (c + 1 < nsub) issue_load(V, g0 + c + 1, (step & 1) ^ 1);
            else if (pass == 0) issue_load(V, g0, (step & 1) ^ 1);
            else {
                const u32 np = atomicAdd(p.ticket, 1u);
                sm.next_piece = np;
                if (np >= p.num_pieces) return;
                const BlockView vn = block_view(p, np);
                issue_load(vn, vn.lp * kNSub, (step & 1) ^ 1);
            }
        };
        // code|len of the thread's 16 symbols; returns their bit count.  Full tiles (all but the
        // very last one) take the path without end-of-input selects.
        auto lookup = [&](u32 buf, u32 tile_n, u32 (&e)[kSyms]) -> u32 {
            const uint4 v = *reinterpret_cast<const uint4 *>(&sm.in[buf][tid * kSyms]);
            const u32 w[4] = {v.x, v.y, v.z, v.w};
            u32 bits = 0;
            if (tile_n == (u32)kTileSyms) {
#pragma unroll
                for (int i = 0; i < kSyms; ++i) {
                    const u32 x = sm.tab[(w[i >> 2] >> (8 * (i & 3))) & 0xffu];
                    e[i] = x;
                    bits += x >> 16;
                }
                return bits;
            }
            const u32 valid = tile_n > tid * kSyms ? min((u32)kSyms, tile_n - tid * kSyms) : 0u;
#pragma unroll
            for (int i = 0; i < kSyms; ++i) {
                const u32 s = (w[i >> 2] >> (8 * (i & 3))) & 0xffu;
                u32 x = sm.tab[s];
                if ((u32)i >= valid) x = 0;  // past the end of the input: zero-length code
                e[i] = x;
                bits += x >> 16;
            }
            return bits;
        };
        // bit count only (pass A): byte-wide length table, no scaling of the index
        auto count_bits = [&](u32 buf, u32 tile_n) -> u32 {
            const uint4 v = *reinterpret_cast<const uint4 *>(&sm.in[buf][tid * kSyms]);
            const u32 w[4] = {v.x, v.y, v.z, v.w};
            const u32 valid = tile_n > tid * kSyms ? min((u32)kSyms, tile_n - tid * kSyms) : 0u;
            u32 bits = 0;
            if (tile_n == (u32)kTileSyms) {
#pragma unroll
                for (int i = 0; i < kSyms; ++i) bits += sm.len8[(w[i >> 2] >> (8 * (i & 3))) & 0xffu];
            } else {
#pragma unroll
                for (int i = 0; i < kSyms; ++i)
                    if ((u32)i < valid) bits += sm.len8[(w[i >> 2] >> (8 * (i & 3))) & 0xffu];
            }
            return bits;
        };

        if (PLANNED) {
            if (tid == 0) {
                sm.base_bits = p.base_bits[piece];
                sm.prev_tail = 0;
            }
            __syncthreads();
        } else {
        // ================================================================ pass A: bit count
        u32 my_piece_bits = 0;   // per-thread partial over all sub-tiles (<= 16*16*13 bits)
        for (u32 c = 0; c < nsub; ++c) {
            const u64 first = (u64)(g0 + c) * kTileSyms;
            const u32 tile_n = (u32)min((u64)kTileSyms, V.n - first);
            acquire_input(V, g0 + c, tile_n);
            prefetch(0, c);
            const u32 bits = count_bits(step & 1, tile_n);
            my_piece_bits += bits;
            if (c == nsub - 1 && tid >= kThreads - 2) {
                // the piece's last 31 bits live in its last two threads (full pieces only)
                u32 e[kSyms];
                lookup(step & 1, tile_n, e);
                u64 acc = 0;
#pragma unroll
                for (int i = 0; i < kSyms; ++i) acc = (acc << (e[i] >> 16)) | (e[i] & 0xffffu);
                sm.tail_acc[tid - (kThreads - 2)] = acc;
                sm.tail_bits[tid - (kThreads - 2)] = bits;
            }
            __syncthreads();   // buffer hand-over
            ++step;
        }
        {   // block reduction of the per-thread partials
            u32 v = my_piece_bits;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if (lane == 0) sm.warp_sums[tid >> 5] = v;
            __syncthreads();
            if (tid < 32) {
                u32 t = tid < kThreads / 32 ? sm.warp_sums[tid] : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
                if (tid == 0) sm.piece_bits = t;
            }
            __syncthreads();
        }

        // ================================================================ publish + look-back (warp 0)
        if (tid < 32) {
            const u64 piece_bits = sm.piece_bits;
            if (lane == 0) {
                const u32 nb1 = sm.tail_bits[1];
                u64 acc = sm.tail_acc[1];
                if (nb1 < 31) acc |= sm.tail_acc[0] << nb1;
                const u32 tail31 = (u32)acc & 0x7fffffffu;
                st_release_u64(&p.desc[piece].agg, kValid | (piece_bits << 31) | tail31);
            }
            u64 base = 0;
            u32 prev_tail = 0;
            const int first_of_block = (int)(piece - V.lp);
            if (V.lp > 0) {
                // warp-wide look-back: lane i inspects piece (k - i); it stops at the block's first piece
                int k = (int)piece - 1;
                bool done = false;
                bool first_batch = true;
                while (!done) {
                    const int idx = k - (int)lane;
                    u64 a = 0, in = 0;
                    if (idx >= first_of_block) {
                        in = ld_acquire_u64(&p.desc[idx].incl);
                        if (!(in & kValid)) a = ld_acquire_u64(&p.desc[idx].agg);
                    }
                    const bool has_incl = idx >= first_of_block && (in & kValid);
                    const bool has_any = idx < first_of_block || has_incl || (a & kValid);
                    const u32 incl_mask = __ballot_sync(0xffffffffu, has_incl);
                    const u32 any_mask = __ballot_sync(0xffffffffu, has_any);
                    // usable prefix of lanes: all ready up to (and including) the first inclusive
                    const u32 first_incl = incl_mask ? (u32)__ffs(incl_mask) - 1 : 32u;
                    const u32 need = first_incl < 32 ? ((2u << first_incl) - 1) : 0xffffffffu;
                    if ((any_mask & need) != need) {
                        __nanosleep(100);
                        continue;  // somebody not ready yet: poll again
                    }
                    u64 contrib = 0;
                    if (lane < first_incl && idx >= first_of_block) contrib = (a & ~kValid) >> 31;
                    if (lane == first_incl) contrib = in & ~kValid;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, d);
                    base += contrib;
                    if (first_batch) {
                        // direct predecessor's tail: from its aggregate word (always published)
                        u64 pa = 0;
                        if (lane == 0) {
                            pa = a & kValid ? a : ld_acquire_u64(&p.desc[piece - 1].agg);
                            while (!(pa & kValid)) pa = ld_acquire_u64(&p.desc[piece - 1].agg);
                        }
                        prev_tail = (u32)__shfl_sync(0xffffffffu, pa, 0) & 0x7fffffffu;
                        first_batch = false;
                    }
                    if (first_incl < 32 || k - 32 < first_of_block) done = true;
                    k -= 32;
                }
            }
            if (lane == 0) {
                st_release_u64(&p.desc[piece].incl, kValid | (base + piece_bits));
                sm.base_bits = base;
                sm.prev_tail = prev_tail;
            }
        }
        __syncthreads();
        }   // !PLANNED

        // ================================================================ pass B: pack + write
        u64 base = sm.base_bits;                 // running global bit offset
        // bits of the first output word that belong to predecessors, left-aligned
        u32 carry = 0;
        {
            const u32 r0 = (u32)(base & 31);
            if (r0) carry = (sm.prev_tail & ((1u << r0) - 1)) << (32 - r0);
        }
        const bool last_piece = V.lp == V.last_lp;        // of its block
        for (u32 c = 0; c < nsub; ++c) {
            const u64 first = (u64)(g0 + c) * kTileSyms;
            const u32 tile_n = (u32)min((u64)kTileSyms, V.n - first);
            acquire_input(V, g0 + c, tile_n);
            prefetch(1, c);
            u32 e[kSyms];
            const u32 my_bits = lookup(step & 1, tile_n, e);

            // block scan of bit counts
            u32 incl = warp_incl_scan(my_bits);
            if (lane == 31) sm.warp_sums[tid >> 5] = incl;
            __syncthreads();
            if (tid < 32) {
                const u32 v = tid < kThreads / 32 ? sm.warp_sums[tid] : 0u;
                const u32 s = warp_incl_scan(v);
                if (tid < kThreads / 32) sm.warp_sums[tid] = s - v;
                if (tid == kThreads / 32 - 1) sm.total = s;
            }
            __syncthreads();
            const u32 pre = sm.warp_sums[tid >> 5] + incl - my_bits;
            const u32 total = sm.total;

            const u32 r = (u32)(base & 31);
            const u64 w0 = base >> 5;                  // global index of staging word `salign`
            const u32 salign = (u32)(w0 & 3);          // keep 16-byte phase of global and staging equal
            const bool last_tile = last_piece && c == nsub - 1;
            const u32 tile_bits = r + total;
            // words written now: all complete ones, plus the final partial one at the very end
            const u32 nwords = last_tile ? (tile_bits + 31) >> 5 : tile_bits >> 5;

            // Every complete word is stored (plain STS) by the one thread that holds its last bit,
            // with zeros where other threads' bits go; after a barrier the leading bits arrive by
            // atomicOr from the threads that hold them (one per thread: its trailing partial
            // word).  Only the tile's last, incomplete word is never stored, so it is zeroed.
            if (tid == 0) sm.stage[salign + (tile_bits >> 5)] = 0;
            const u32 b0 = r + pre;
            u32 wi = salign + (b0 >> 5);
            u32 nb = b0 & 31;          // bits already occupied in the current word
            u64 acc = 0;
#pragma unroll
            for (int i = 0; i < kSyms; i += 2) {   // two symbols at a time (<= 26 bits)
                const u32 lb = e[i + 1] >> 16;
                const u32 c2 = ((e[i] & 0xffffu) << lb) | (e[i + 1] & 0xffffu);
                const u32 l2 = (e[i] >> 16) + lb;
                acc = (acc << l2) | c2;
                nb += l2;
                if (nb >= 32) sm.stage[wi] = (u32)(acc >> (nb - 32));
                wi += nb >> 5;
                nb &= 31;
            }
            __syncthreads();
            if (nb && my_bits) atomicOr(&sm.stage[wi], (u32)acc << (32 - nb));
            if (tid == 0 && r) atomicOr(&sm.stage[salign], carry);
            __syncthreads();
            // the partial last word travels to the next sub-tile of this piece
            carry = sm.stage[salign + (tile_bits >> 5)];
            if (PLANNED && c == nsub - 1 && !last_piece && (tile_bits & 31) && tid == 0) {
                // ... or, at the end of a planned piece, is ORed into the word shared with the successor
                const u64 idx = w0 + (tile_bits >> 5);
                if (idx < p.out_cap_units) atomicOr(&V.out[idx], carry);
                else atomicExch(p.overflow, 1u);
            }

            if (w0 + nwords > p.out_cap_units) {
                if (tid == 0) atomicExch(p.overflow, 1u);
            } else {
                u32 *g = V.out + w0;
                // a planned piece shares its first word with its predecessor: OR instead of store
                const u32 skip = (PLANNED && c == 0 && r != 0 && nwords > 0) ? 1u : 0u;
                if (skip && tid == 0) atomicOr(g, sm.stage[salign]);
                const u32 sa = salign + skip, nw = nwords - skip;
                u32 *gs = g + skip;
                const u32 head = min(nw, (4u - (sa & 3u)) & 3u);
                const u32 nvec = (nw - head) >> 2;
                const u32 tail0 = head + (nvec << 2);
                if (tid < head) gs[tid] = sm.stage[sa + tid];
                const uint4 *sv = reinterpret_cast<const uint4 *>(&sm.stage[sa + head]);
                uint4 *gv = reinterpret_cast<uint4 *>(gs + head);
                for (u32 i = tid; i < nvec; i += kThreads) gv[i] = sv[i];
                if (tid < nw - tail0) gs[tail0 + tid] = sm.stage[sa + tail0 + tid];
                if (last_tile && tid == 0) {
                    if (p.block_bits) p.block_bits[V.block] = base + total;
                    else *p.total_bits = base + total;
                    if (w0 + nwords < p.out_cap_units) g[nwords] = 0;  // the reference's pad unit
                }
            }
            base += total;
            __syncthreads();   // staging + input buffer hand-over
            ++step;
        }
    }
}

// ---------------------------------------------------------------------------------- planning
// One 256-bin histogram per piece (kPieceSyms symbols), one CTA per piece.
__global__ void __launch_bounds__(256) piece_hist_kernel(const u8 *__restrict__ in, u64 n, u32 *__restrict__ piece_hist)
{
    __shared__ u32 sh[8][256];
    const u32 tid = threadIdx.x, warp = tid >> 5;
    for (u32 i = tid; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    const u64 first = (u64)blockIdx.x * kPieceSyms;
    const u32 cnt = (u32)min((u64)kPieceSyms, n - first);
    const uint4 *v = reinterpret_cast<const uint4 *>(in + first);      // piece starts are 16-byte aligned
    for (u32 i = tid; i < (cnt >> 4); i += 256) {
        const uint4 x = v[i];
        const u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            atomicAdd(&sh[warp][w[q] & 0xff], 1u);
            atomicAdd(&sh[warp][(w[q] >> 8) & 0xff], 1u);
            atomicAdd(&sh[warp][(w[q] >> 16) & 0xff], 1u);
            atomicAdd(&sh[warp][w[q] >> 24], 1u);
        }
    }
    for (u32 i = (cnt & ~15u) + tid; i < cnt; i += 256) atomicAdd(&sh[warp][in[first + i]], 1u);
    __syncthreads();
    u32 s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w][tid];
    piece_hist[(u64)blockIdx.x * 256 + tid] = s;
}

// column sums of the piece histograms -> the 64-bit histogram of the whole input
__global__ void __launch_bounds__(256) hist_reduce_kernel(const u32 *__restrict__ piece_hist, u32 pieces,
                                                          unsigned long long *__restrict__ hist)
{
    unsigned long long s = 0;
    for (u32 p = blockIdx.x; p < pieces; p += gridDim.x) s += piece_hist[(u64)p * 256 + threadIdx.x];
    if (s) atomicAdd(&hist[threadIdx.x], s);
}

// bits of every piece = <its histogram, code lengths>; one warp per piece
__global__ void __launch_bounds__(256) piece_bits_kernel(const u32 *__restrict__ piece_hist, u32 pieces,
                                                         const u8 *__restrict__ len_of_symbol,
                                                         u64 *__restrict__ piece_bits)
{
    const u32 piece = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (piece >= pieces) return;
    u32 s = 0;
    for (u32 k = lane; k < 256; k += 32) s += piece_hist[(u64)piece * 256 + k] * (u32)len_of_symbol[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) piece_bits[piece] = s;
}

// exclusive scan of the piece bits (in place), zeroing of every word two pieces share, capacity check
__global__ void __launch_bounds__(1024) plan_kernel(u64 *__restrict__ bits, u32 pieces, u32 *__restrict__ out,
                                                    u64 out_cap_units, u32 *__restrict__ overflow)
{
    __shared__ u64 ws[32];
    __shared__ u64 s_carry;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (u32 lo = 0; lo < pieces; lo += 1024) {
        const u32 i = lo + tid;
        const u64 v = i < pieces ? bits[i] : 0;
        u64 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u64 t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (u32)d) incl += t;
        }
        if (lane == 31) ws[warp] = incl;
        __syncthreads();
        u64 pre = s_carry;
        for (u32 w = 0; w < warp; ++w) pre += ws[w];
        const u64 base = pre + incl - v;
        if (i < pieces) {
            bits[i] = base;
            if (i > 0 && (base & 31)) {
                if ((base >> 5) < out_cap_units) out[base >> 5] = 0;
                else atomicExch(overflow, 1u);
            }
            if (i == pieces - 1 && ((base + v + 31) >> 5) > out_cap_units) atomicExch(overflow, 1u);
        }
        __syncthreads();
        if (tid == 1023) s_carry = base + v;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- histogram
constexpr int kHistThreads = 256;
__global__ void __launch_bounds__(kHistThreads) histogram_u8_kernel(const u8 *in, u64 n,
                                                                    unsigned long long *hist)
{
    __shared__ u32 sh[kHistThreads / 32][256];
    const u32 tid = threadIdx.x, warp = tid >> 5;
    for (u32 i = tid; i < (kHistThreads / 32) * 256; i += kHistThreads) (&sh[0][0])[i] = 0;
    __syncthreads();
    const u64 nvec = n >> 4;
    const uint4 *v = reinterpret_cast<const uint4 *>(in);
    // 16-byte aligned base assumed for the vector part (checked by the host)
    for (u64 i = (u64)blockIdx.x * kHistThreads + tid; i < nvec; i += (u64)gridDim.x * kHistThreads) {
        const uint4 x = v[i];
        const u32 w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            atomicAdd(&sh[warp][w[q] & 0xff], 1u);
            atomicAdd(&sh[warp][(w[q] >> 8) & 0xff], 1u);
            atomicAdd(&sh[warp][(w[q] >> 16) & 0xff], 1u);
            atomicAdd(&sh[warp][w[q] >> 24], 1u);
        }
    }
    if (blockIdx.x == 0)
        for (u64 i = (nvec << 4) + tid; i < n; i += kHistThreads) atomicAdd(&sh[warp][in[i]], 1u);
    __syncthreads();
    if (tid < 256) {
        u32 s = 0;
#pragma unroll
        for (int w = 0; w < kHistThreads / 32; ++w) s += sh[w][tid];
        if (s) atomicAdd(&hist[tid], (unsigned long long)s);
    }
}

static u32 subtiles_for(u64 n) { return (u32)((n + kTileSyms - 1) / kTileSyms); }
static u32 pieces_for(u64 n) { return (subtiles_for(n) + kNSub - 1) / kNSub; }

}  // namespace cuhd_enc
}  // namespace b200lc

using namespace b200lc;

extern "C" int b200lc_histogram_u8(const uint8_t *d_in, size_t n, uint64_t *d_hist, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_hist || (n && !d_in)) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_in) & 15) return B200LC_ERR_ARG;
    B200LC_CUDA_TRY(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint64_t), stream));
    if (n == 0) return B200LC_OK;
    // 32-bit per-block counters: bound the bytes one block can see below 2^32
    const u64 min_blocks = (n >> 31) + 1;
    const u64 want = (u64)num_sms() * 8;
    const u64 max_useful = (n / (16 * cuhd_enc::kHistThreads)) + 1;
    u64 grid = want < max_useful ? want : max_useful;
    if (grid < min_blocks) grid = min_blocks;
    cuhd_enc::histogram_u8_kernel<<<(u32)grid, cuhd_enc::kHistThreads, 0, stream>>>(
        d_in, n, reinterpret_cast<unsigned long long *>(d_hist));
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

extern "C" size_t b200lc_cuhd_encode_scratch_bytes(size_t n)
{
    return 256 + (size_t)cuhd_enc::pieces_for(n) * sizeof(cuhd_enc::EncDesc);
}

extern "C" int b200lc_cuhd_encode(const uint8_t *d_in, size_t n, const uint32_t *d_code_of_symbol,
                                  const uint8_t *d_len_of_symbol, uint32_t *d_units,
                                  size_t units_cap, uint64_t *d_total_bits, void *d_scratch,
                                  size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_code_of_symbol || !d_len_of_symbol || !d_units || !d_total_bits || !d_scratch)
        return B200LC_ERR_ARG;
    if (n && !d_in) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_scratch) & 127) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_units) & 15) return B200LC_ERR_ARG;
    const size_t need = b200lc_cuhd_encode_scratch_bytes(n);
    if (scratch_bytes < need) return B200LC_ERR_SCRATCH;
    B200LC_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, need, stream));
    if (n == 0) {
        B200LC_CUDA_TRY(cudaMemsetAsync(d_total_bits, 0, sizeof(uint64_t), stream));
        return B200LC_OK;
    }
    cuhd_enc::EncParams p;
    p.in = d_in;
    p.n = n;
    p.code_of_symbol = d_code_of_symbol;
    p.len_of_symbol = d_len_of_symbol;
    p.out = d_units;
    p.out_cap_units = units_cap;
    p.total_bits = d_total_bits;
    p.ticket = reinterpret_cast<u32 *>(d_scratch);
    p.overflow = reinterpret_cast<u32 *>(d_scratch) + 1;
    p.desc = reinterpret_cast<cuhd_enc::EncDesc *>(reinterpret_cast<char *>(d_scratch) + 256);
    p.num_pieces = cuhd_enc::pieces_for(n);
    p.block_syms = n;
    p.unit_stride = 0;
    p.pieces_per_block = p.num_pieces;
    p.aligned = (reinterpret_cast<uintptr_t>(d_in) & 15) == 0;
    p.block_bits = nullptr;
    p.base_bits = nullptr;

    static int occ_dev[kMaxDevices] = {0};
    const int slot = device_slot();
    int occ = slot >= 0 ? occ_dev[slot] : 0;
    if (!occ) {
        B200LC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
            &occ, cuhd_enc::cuhd_encode_kernel<false>, cuhd_enc::kThreads, 0));
        if (occ < 1) return B200LC_ERR_CUDA;
        if (slot >= 0) occ_dev[slot] = occ;
    }
    const u32 grid = (u32)min((u64)p.num_pieces, (u64)num_sms() * (u64)occ);
    cuhd_enc::cuhd_encode_kernel<false><<<grid, cuhd_enc::kThreads, 0, stream>>>(p);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

// Histogram that also keeps one 256-bin histogram per piece for b200lc_cuhd_encode_planned.
extern "C" size_t b200lc_cuhd_piece_hist_bytes(size_t n)
{
    return (size_t)cuhd_enc::pieces_for(n) * 256 * sizeof(u32);
}

extern "C" int b200lc_histogram_u8_pieces(const uint8_t *d_in, size_t n, uint64_t *d_hist,
                                          uint32_t *d_piece_hist, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_hist || !d_piece_hist || (n && !d_in)) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_in) & 15) return B200LC_ERR_ARG;
    B200LC_CUDA_TRY(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint64_t), stream));
    if (n == 0) return B200LC_OK;
    const u32 pieces = cuhd_enc::pieces_for(n);
    cuhd_enc::piece_hist_kernel<<<pieces, 256, 0, stream>>>(d_in, n, d_piece_hist);
    cuhd_enc::hist_reduce_kernel<<<min(pieces, 64u), 256, 0, stream>>>(
        d_piece_hist, pieces, reinterpret_cast<unsigned long long *>(d_hist));
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

// The same in two steps for callers whose input arrives in chunks (b200lc_cuhd_session_encode):
// pieces [first_piece, end_piece) of the n-symbol buffer, then the reduction of all of them.
extern "C" size_t b200lc_cuhd_piece_symbols(void) { return (size_t)cuhd_enc::kPieceSyms; }

extern "C" int b200lc_histogram_u8_pieces_part(const uint8_t *d_in, size_t n, size_t first_piece,
                                               size_t end_piece, uint32_t *d_piece_hist, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_piece_hist || !d_in) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_in) & 15) return B200LC_ERR_ARG;
    const size_t pieces = cuhd_enc::pieces_for(n);
    if (end_piece > pieces) end_piece = pieces;
    if (first_piece >= end_piece) return B200LC_OK;
    const size_t lo = first_piece * (size_t)cuhd_enc::kPieceSyms;
    cuhd_enc::piece_hist_kernel<<<(u32)(end_piece - first_piece), 256, 0, stream>>>(
        d_in + lo, n - lo, d_piece_hist + first_piece * 256);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

extern "C" int b200lc_histogram_u8_pieces_finish(const uint32_t *d_piece_hist, size_t n, uint64_t *d_hist,
                                                 void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_hist || !d_piece_hist) return B200LC_ERR_ARG;
    B200LC_CUDA_TRY(cudaMemsetAsync(d_hist, 0, 256 * sizeof(uint64_t), stream));
    if (n == 0) return B200LC_OK;
    const u32 pieces = cuhd_enc::pieces_for(n);
    cuhd_enc::hist_reduce_kernel<<<min(pieces, 64u), 256, 0, stream>>>(
        d_piece_hist, pieces, reinterpret_cast<unsigned long long *>(d_hist));
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

// b200lc_cuhd_encode with the piece histograms of b200lc_histogram_u8_pieces: same stream, one pass.
extern "C" int b200lc_cuhd_encode_planned(const uint8_t *d_in, size_t n, const uint32_t *d_code_of_symbol,
                                          const uint8_t *d_len_of_symbol, const uint32_t *d_piece_hist,
                                          uint32_t *d_units, size_t units_cap, uint64_t *d_total_bits,
                                          void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_code_of_symbol || !d_len_of_symbol || !d_piece_hist || !d_units || !d_total_bits || !d_scratch)
        return B200LC_ERR_ARG;
    if (n && !d_in) return B200LC_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(d_scratch) & 127) || (reinterpret_cast<uintptr_t>(d_units) & 15) ||
        (reinterpret_cast<uintptr_t>(d_in) & 15))
        return B200LC_ERR_ARG;
    const size_t need = b200lc_cuhd_encode_scratch_bytes(n);
    if (scratch_bytes < need) return B200LC_ERR_SCRATCH;
    B200LC_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, 256, stream));
    if (n == 0) {
        B200LC_CUDA_TRY(cudaMemsetAsync(d_total_bits, 0, sizeof(uint64_t), stream));
        return B200LC_OK;
    }
    cuhd_enc::EncParams p;
    p.in = d_in;
    p.n = n;
    p.code_of_symbol = d_code_of_symbol;
    p.len_of_symbol = d_len_of_symbol;
    p.out = d_units;
    p.out_cap_units = units_cap;
    p.total_bits = d_total_bits;
    p.ticket = reinterpret_cast<u32 *>(d_scratch);
    p.overflow = reinterpret_cast<u32 *>(d_scratch) + 1;
    p.desc = nullptr;
    p.num_pieces = cuhd_enc::pieces_for(n);
    p.block_syms = n;
    p.unit_stride = 0;
    p.pieces_per_block = p.num_pieces;
    p.aligned = 1;
    p.block_bits = nullptr;
    u64 *base_bits = reinterpret_cast<u64 *>(reinterpret_cast<char *>(d_scratch) + 256);
    p.base_bits = base_bits;
    cuhd_enc::piece_bits_kernel<<<(p.num_pieces + 7) / 8, 256, 0, stream>>>(d_piece_hist, p.num_pieces,
                                                                            d_len_of_symbol, base_bits);
    cuhd_enc::plan_kernel<<<1, 1024, 0, stream>>>(base_bits, p.num_pieces, d_units, units_cap, p.overflow);
    static int occ_dev[kMaxDevices] = {0};
    const int slot = device_slot();
    int occ = slot >= 0 ? occ_dev[slot] : 0;
    if (!occ) {
        B200LC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cuhd_enc::cuhd_encode_kernel<true>,
                                                                      cuhd_enc::kThreads, 0));
        if (occ < 1) return B200LC_ERR_CUDA;
        if (slot >= 0) occ_dev[slot] = occ;
    }
    const u32 grid = (u32)min((u64)p.num_pieces, (u64)num_sms() * (u64)occ);
    cuhd_enc::cuhd_encode_kernel<true><<<grid, cuhd_enc::kThreads, 0, stream>>>(p);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

// Blocks of block_symbols symbols packed as independent streams with one dictionary, one launch.
extern "C" size_t b200lc_cuhd_encode_blocks_scratch_bytes(size_t n, size_t block_symbols)
{
    if (block_symbols == 0) return 0;
    const size_t blocks = (n + block_symbols - 1) / block_symbols;
    return 256 + blocks * (size_t)cuhd_enc::pieces_for(block_symbols) * sizeof(cuhd_enc::EncDesc);
}

extern "C" int b200lc_cuhd_encode_blocks(const uint8_t *d_in, size_t n, size_t block_symbols,
                                         const uint32_t *d_code_of_symbol, const uint8_t *d_len_of_symbol,
                                         uint32_t *d_units, size_t unit_stride, uint64_t *d_block_bits,
                                         void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_code_of_symbol || !d_len_of_symbol || !d_units || !d_block_bits || !d_scratch) return B200LC_ERR_ARG;
    if (n == 0) return B200LC_OK;
    if (!d_in || block_symbols == 0 || (unit_stride & 3)) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_scratch) & 127) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_units) & 15) return B200LC_ERR_ARG;
    const size_t need = b200lc_cuhd_encode_blocks_scratch_bytes(n, block_symbols);
    if (scratch_bytes < need) return B200LC_ERR_SCRATCH;
    const u64 blocks = (n + block_symbols - 1) / block_symbols;
    const u64 ppb = cuhd_enc::pieces_for(block_symbols);
    if (blocks * ppb >= (1ull << 31)) return B200LC_ERR_UNSUPPORTED;
    B200LC_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, need, stream));
    cuhd_enc::EncParams p;
    p.in = d_in;
    p.n = n;
    p.code_of_symbol = d_code_of_symbol;
    p.len_of_symbol = d_len_of_symbol;
    p.out = d_units;
    p.out_cap_units = unit_stride;
    p.total_bits = nullptr;
    p.ticket = reinterpret_cast<u32 *>(d_scratch);
    p.overflow = reinterpret_cast<u32 *>(d_scratch) + 1;
    p.desc = reinterpret_cast<cuhd_enc::EncDesc *>(reinterpret_cast<char *>(d_scratch) + 256);
    p.num_pieces = (u32)(blocks * ppb);
    p.block_syms = block_symbols;
    p.unit_stride = unit_stride;
    p.pieces_per_block = (u32)ppb;
    p.aligned = (reinterpret_cast<uintptr_t>(d_in) & 15) == 0 && block_symbols % 16 == 0;
    p.block_bits = d_block_bits;
    p.base_bits = nullptr;
    int occ = 0;
    B200LC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cuhd_enc::cuhd_encode_kernel<false>,
                                                                  cuhd_enc::kThreads, 0));
    if (occ < 1) return B200LC_ERR_CUDA;
    const u32 grid = (u32)min((u64)p.num_pieces, (u64)num_sms() * (u64)occ);
    cuhd_enc::cuhd_encode_kernel<false><<<grid, cuhd_enc::kThreads, 0, stream>>>(p);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

// Overflow flag of the last b200lc_cuhd_encode call on this scratch buffer (synchronises).
extern "C" int b200lc_cuhd_encode_overflowed(const void *d_scratch, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    u32 flag = 0;
    B200LC_CUDA_TRY(cudaMemcpyAsync(&flag, reinterpret_cast<const u32 *>(d_scratch) + 1,
                                    sizeof(u32), cudaMemcpyDeviceToHost, stream));
    B200LC_CUDA_TRY(cudaStreamSynchronize(stream));
    return flag ? B200LC_ERR_OVERFLOW : B200LC_OK;
}
