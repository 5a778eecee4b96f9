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
// HBM traffic is therefore units-in + symbols-out + 128 B of descriptor per tile; the
// reference's 20 B of sync-point state per 16 B of input is gone.
//
// Decode contract (bit-exact with the reference): out[i] = symbol of
// the i-th codeword when reading the stream MSB-first from bit 0 of unit 0 through the flat LUT
// `{u8 num_bits, u8 symbol}[1 << max_codeword_length]`; units past n_units read as zero.
#include "common.cuh"
#include "../../include/b200lc.h"

namespace b200lc {
namespace cuhd {

constexpr u32 kMaxStates = 16;       // entry states 0..L-1, L <= 13 supported
constexpr u32 kUnknown = 0xFFu;
constexpr u64 kInclValid = 1ull << 63;
constexpr u64 kCountMask = (1ull << 56) - 1;

struct __align__(128) TileDesc {
    u64 incl;            // bit 63 valid | bits 56..59 end state | bits 0..55 symbols before next tile
    u32 agg_ready;       // 1 once end[]/cnt[] are valid
    u32 pad0;
    u8 end[kMaxStates];  // end state per entry state, kUnknown if it did not merge in this tile
    u32 cnt[kMaxStates]; // symbols starting in this tile per entry state
    u32 pad1[6];
};
static_assert(sizeof(TileDesc) == 128, "descriptor is one 128-byte line");

struct DecodeParams {
    const u32 *units;
    u64 n_units;
    const u16 *lut;      // {u8 num_bits, u8 symbol} little-endian pairs
    u32 max_len;         // L
    u8 *out;
    u64 n_out;
    TileDesc *desc;
    u32 *ticket;
    u32 num_tiles;
    u32 tma_ok_base;     // 1 if units pointer is 16-byte aligned
};

// ---------------------------------------------------------------------------------- walks
// Round 0: decode the subsequence from bit 0, remember every codeword start.
template <int S>
__device__ __forceinline__ void walk_record(const u32 (&u)[S + 1], const u32 *tab, u32 shift,
                                            u32 (&m)[S], u32 &end, u32 &cnt)
{
    u32 at = 0, c = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        u32 mj = 0;
        while (at < 32) {
            mj |= 0x80000000u >> at;
            const u32 w = __funnelshift_l(nxt, cur, at);
            at += tab[w >> shift] & 0xffu;
        }
        m[j] = mj;
        c += __popc(mj);
        at -= 32;
    }
    end = at;
    cnt = c;
}

// Decode from entry state `a` until the walk lands on a codeword start of the recorded path
// (then the rest of the subsequence is the recorded path: end = e0) or runs off the end.
template <int S>
__device__ __forceinline__ void walk_merge(const u32 (&u)[S + 1], const u32 (&m)[S], u32 a, u32 e0,
                                           const u32 *tab, u32 shift, u32 &end, u32 &cnt)
{
    u32 at = a, k = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        while (at < 32) {
            const u32 bit = 0x80000000u >> at;
            if (m[j] & bit) {
                u32 rest = __popc(m[j] & (bit | (bit - 1)));
#pragma unroll
                for (int jj = j + 1; jj < S; ++jj) rest += __popc(m[jj]);
                end = e0;
                cnt = k + rest;
                return;
            }
            const u32 w = __funnelshift_l(nxt, cur, at);
            at += tab[w >> shift] & 0xffu;
            ++k;
        }
        at -= 32;
    }
    end = at;
    cnt = k;
}

// Write pass: decode from the true entry state, symbol i of this subsequence goes to dst[i].
// With CHECK, only tile-local positions in [lo, hi) are stored (staging-window overflow path).
template <int S, bool CHECK>
__device__ __forceinline__ void walk_write(const u32 (&u)[S + 1], const u32 *tab, u32 shift, u32 a,
                                           u8 *dst, u32 pos, u32 lo, u32 hi)
{
    u32 at = a;
#pragma unroll
    for (int j = 0; j < S; ++j) {
        const u32 cur = u[j], nxt = u[j + 1];
        while (at < 32) {
            const u32 w = __funnelshift_l(nxt, cur, at);
            const u32 e = tab[w >> shift];
            if (!CHECK || (pos >= lo && pos < hi)) dst[pos] = (u8)(e >> 8);
            ++pos;
            at += e & 0xffu;
        }
        at -= 32;
    }
}

// ---------------------------------------------------------------------------------- kernel
template <int S, int T, int CAP>
struct SmemLayout {
    static constexpr int kTileUnits = T * S + 4;  // + one 16-byte lookahead
    u32 in[2][kTileUnits];
    __align__(16) u8 stage[CAP + 16];
    u32 warp_sums[T / 32];
    u32 mask0[S];
    u8 end[T];
    u64 bar[2];
    u64 base;
    u32 tile[2];
    u32 total;
    u32 astar;
    u32 e0_first;
    u32 cnt_first;
    u32 redo;
    int delta;
};

template <int S, int T, int CAP>
__global__ void __launch_bounds__(T + 32) cuhd_decode_kernel(const DecodeParams p)
{
    using Smem = SmemLayout<S, T, CAP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &sm = *reinterpret_cast<Smem *>(smem_raw);
    u32 *tab = reinterpret_cast<u32 *>(smem_raw + ((sizeof(Smem) + 127) & ~size_t(127)));

    const u32 tid = threadIdx.x;
    const u32 lane = tid & 31;
    const bool worker = tid < T;
    const u32 L = p.max_len;
    const u32 shift = 32 - L;
    constexpr u32 kTileUnits = Smem::kTileUnits;
    constexpr u32 kTileBytes = kTileUnits * 4;

    // LUT -> shared memory as 32-bit entries: bits 0..7 length, 8..15 symbol.  A zero-length
    // entry (unused prefix of an incomplete code) would stall the reference forever; it is
    // mapped to length 1 here so that garbage input still terminates.
    for (u32 i = tid; i < (1u << L); i += blockDim.x) {
        const u32 e = p.lut[i];
        u32 len = e & 0xffu;
        if (len == 0 || len > L) len = 1;
        tab[i] = len | (e & 0xff00u);
    }
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto tma_ok = [&](u32 tile) -> bool {
        return p.tma_ok_base && (u64)(tile + 1) * (T * S) + 4 <= p.n_units;
    };
    auto issue_load = [&](u32 tile, u32 buf) {  // one thread
        if (tile < p.num_tiles && tma_ok(tile)) {
            mbar_expect_tx(&sm.bar[buf], kTileBytes);
            tma_load_1d(sm.in[buf], p.units + (u64)tile * (T * S), kTileBytes, &sm.bar[buf]);
        }
    };

    if (tid == 0) {
        const u32 t0 = atomicAdd(p.ticket, 1u);
        sm.tile[0] = t0;
        issue_load(t0, 0);
    }
    __syncthreads();

    u32 cur = 0;
    u32 phase0 = 0, phase1 = 0;
    while (true) {
        const u32 tile = sm.tile[cur];
        if (tile >= p.num_tiles) break;

        // ---------------------------------------------------------------- stage input
        if (tma_ok(tile)) {
            if (cur == 0) { mbar_wait(&sm.bar[0], phase0); phase0 ^= 1; }
            else          { mbar_wait(&sm.bar[1], phase1); phase1 ^= 1; }
        } else {
            const u64 first = (u64)tile * (T * S);
            for (u32 i = tid; i < kTileUnits; i += blockDim.x) {
                const u64 idx = first + i;
                sm.in[cur][i] = idx < p.n_units ? p.units[idx] : 0u;
            }
            fence_proxy_async();
            __syncthreads();
        }
        if (tid == 0) {  // ticket + TMA prefetch of the next tile into the other buffer
            const u32 nt = atomicAdd(p.ticket, 1u);
            sm.tile[cur ^ 1] = nt;
            issue_load(nt, cur ^ 1);
        }

        // ---------------------------------------------------------------- round 0
        u32 u[S + 1], m[S];
        u32 e0 = 0, c0 = 0;
        if (worker) {
            const uint4 *src = reinterpret_cast<const uint4 *>(&sm.in[cur][tid * S]);
#pragma unroll
            for (int q = 0; q < S / 4; ++q) {
                const uint4 v = src[q];
                u[4 * q + 0] = v.x; u[4 * q + 1] = v.y; u[4 * q + 2] = v.z; u[4 * q + 3] = v.w;
            }
            u[S] = sm.in[cur][tid * S + S];
            walk_record<S>(u, tab, shift, m, e0, c0);
            sm.end[tid] = (u8)e0;
            if (tid == 0) {
#pragma unroll
                for (int j = 0; j < S; ++j) sm.mask0[j] = m[j];
                sm.e0_first = e0;
                sm.cnt_first = c0;
            }
        }
        __syncthreads();

        // ---------------------------------------------------------------- chain from state 0
        u32 my_start = 0, my_end = e0, my_cnt = c0;
        bool need_eval = false;
        if (worker && tid > 0) {
            my_start = sm.end[tid - 1];
            need_eval = my_start != 0;
        }
        // entry states 1..L-1 of the tile's first subsequence (alt warp, one lane per state)
        u32 alt_end = kUnknown, alt_cnt = 0;
        if (!worker && lane >= 1 && lane < L) {
            u32 au[S + 1], am[S];
#pragma unroll
            for (int j = 0; j <= S; ++j) au[j] = sm.in[cur][j];
#pragma unroll
            for (int j = 0; j < S; ++j) am[j] = sm.mask0[j];
            walk_merge<S>(au, am, lane, sm.e0_first, tab, shift, alt_end, alt_cnt);
        }
        __syncthreads();

        auto resolve = [&](bool eval) {
            while (true) {
                bool changed = false;
                if (eval) {
                    u32 ne = e0, nc = c0;
                    if (my_start != 0) walk_merge<S>(u, m, my_start, e0, tab, shift, ne, nc);
                    changed = ne != my_end;
                    my_end = ne;
                    my_cnt = nc;
                    if (changed) sm.end[tid] = (u8)ne;
                }
                if (!__syncthreads_or(changed)) break;
                eval = false;
                if (worker && tid > 0) {
                    const u32 ns = sm.end[tid - 1];
                    eval = ns != my_start;
                    my_start = ns;
                }
                __syncthreads();
            }
        };
        resolve(need_eval);

        // ---------------------------------------------------------------- block scan of counts
        u32 pre = 0;
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
        block_scan();

        // ---------------------------------------------------------------- publish + look-back
        if (!worker) {
            const u32 total0 = sm.total;
            const u32 tile_end0 = sm.end[T - 1];
            u32 f_end = kUnknown, f_cnt = 0;  // this tile as a function of entry state `lane`
            if (lane == 0) {
                f_end = tile_end0;
                f_cnt = total0;
            } else if (lane < L && alt_end == sm.e0_first) {
                f_end = tile_end0;
                f_cnt = total0 - sm.cnt_first + alt_cnt;
            }
            TileDesc *d = &p.desc[tile];
            u32 astar = 0;
            u64 base = 0;
            if (tile == 0) {
                if (lane == 0) st_release_u64(&d->incl, kInclValid | ((u64)f_end << 56) | f_cnt);
            } else {
                if (lane < kMaxStates) {
                    d->end[lane] = (u8)f_end;
                    d->cnt[lane] = f_cnt;
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) st_release_u32(&d->agg_ready, 1u);

                // comp = effect of tiles (k, tile) on the entry state of tile k+1
                u32 comp_end = lane < kMaxStates ? lane : kUnknown;
                u64 comp_cnt = 0;
                u32 k = tile - 1;
                while (true) {
                    const TileDesc *q = &p.desc[k];
                    const u64 incl = ld_acquire_u64(&q->incl);
                    if (incl & kInclValid) {
                        const u32 x = (u32)(incl >> 56) & 0xfu;
                        const u32 r_end = __shfl_sync(0xffffffffu, comp_end, x);
                        const u64 r_cnt = __shfl_sync(0xffffffffu, comp_cnt, x);
                        if (r_end != kUnknown) {
                            astar = r_end;
                            base = (incl & kCountMask) + r_cnt;
                            break;
                        }
                        // the needed entry did not merge somewhere in (k, tile): wait for the
                        // direct predecessor to finish its own slow path.
                        comp_end = lane < kMaxStates ? lane : kUnknown;
                        comp_cnt = 0;
                        k = tile - 1;
                        while (!(ld_acquire_u64(&p.desc[k].incl) & kInclValid)) __nanosleep(64);
                        continue;
                    }
                    if (!ld_acquire_u32(&q->agg_ready)) {
                        __nanosleep(32);
                        continue;
                    }
                    u32 e = kUnknown, c = 0;
                    if (lane < kMaxStates) {
                        e = __ldcg(&q->end[lane]);
                        c = __ldcg(&q->cnt[lane]);
                    }
                    const u32 src = e & 0xfu;
                    const u32 ne = __shfl_sync(0xffffffffu, comp_end, src);
                    const u64 nc = __shfl_sync(0xffffffffu, comp_cnt, src);
                    if (e == kUnknown || ne == kUnknown) {
                        comp_end = kUnknown;
                        comp_cnt = 0;
                    } else {
                        comp_end = ne;
                        comp_cnt = nc + c;
                    }
                    --k;  // k == 0 always carries a valid inclusive prefix, so this terminates
                }
                // inclusive prefix of this tile, if its entry state merged
                const u32 my_f_end = __shfl_sync(0xffffffffu, f_end, astar);
                const u32 my_f_cnt = __shfl_sync(0xffffffffu, f_cnt, astar);
                if (lane == 0 && my_f_end != kUnknown)
                    st_release_u64(&d->incl, kInclValid | ((u64)my_f_end << 56) | (base + my_f_cnt));
            }
            if (lane == 0) {
                sm.astar = astar;
                sm.base = base;
                sm.redo = 0;
                sm.delta = 0;
            }
        }
        __syncthreads();

        // ---------------------------------------------------------------- fix-up for entry state a*
        const u32 astar = sm.astar;
        const u64 base = sm.base;
        if (astar != 0) {
            if (tid == 0) {
                u32 ne, nc;
                walk_merge<S>(u, m, astar, e0, tab, shift, ne, nc);
                my_start = astar;
                if (ne == my_end) {
                    sm.delta = (int)nc - (int)my_cnt;
                    my_cnt = nc;
                } else {
                    sm.redo = 1;
                }
            }
            __syncthreads();
            if (sm.redo) {
                // rare: the true entry state does not merge inside the first subsequence
                resolve(tid == 0);
                block_scan();
                if (tid == 0)
                    st_release_u64(&p.desc[tile].incl,
                                   kInclValid | ((u64)sm.end[T - 1] << 56) | (base + sm.total));
            } else {
                if (worker && tid > 0) pre += sm.delta;
            }
        }
        const u32 total = sm.redo ? sm.total : (u32)((int)sm.total + sm.delta);

        // ---------------------------------------------------------------- write pass
        u64 tile_cnt = 0;
        if (base < p.n_out) tile_cnt = min((u64)total, p.n_out - base);
        for (u32 w0 = 0; w0 < tile_cnt; w0 += CAP) {
            const u32 wlen = (u32)min((u64)CAP, tile_cnt - w0);
            u8 *g = p.out + base + w0;
            const u32 sh = (u32)(reinterpret_cast<uintptr_t>(g) & 15u);
            if (worker) {
                const u32 lo = w0, hi = w0 + wlen;
                if (pre < hi && pre + my_cnt > lo) {
                    u8 *dst = sm.stage + ((int)sh - (int)w0);
                    if (pre >= lo && pre + my_cnt <= hi)
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
        __syncthreads();  // sm.redo/sm.delta/sm.tile reuse, input buffer hand-over
        cur ^= 1;
    }
}

// ---------------------------------------------------------------------------------- host
constexpr int kS = 4;
constexpr int kT = 256;
constexpr int kCap = 3 * kT * kS * 4;

static size_t smem_bytes(u32 L)
{
    return ((sizeof(SmemLayout<kS, kT, kCap>) + 127) & ~size_t(127)) + (size_t(4) << L);
}

static u32 tiles_for(u64 n_units)
{
    const u64 nsub = (n_units + kS - 1) / kS;
    return (u32)((nsub + kT - 1) / kT);
}

}  // namespace cuhd
}  // namespace b200lc

using namespace b200lc;

extern "C" size_t b200lc_cuhd_decode_scratch_bytes(size_t n_units)
{
    return 128 + (size_t)cuhd::tiles_for(n_units) * sizeof(cuhd::TileDesc);
}

extern "C" int b200lc_cuhd_decode(const uint32_t *d_units, size_t n_units, uint8_t *d_out,
                                  size_t n_out, const void *d_table, int max_codeword_length,
                                  void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (max_codeword_length < 1 || max_codeword_length > 13) return B200LC_ERR_UNSUPPORTED;
    if (n_out == 0 || n_units == 0) return B200LC_OK;
    if (!d_units || !d_out || !d_table || !d_scratch) return B200LC_ERR_ARG;
    if (n_units >= (1ull << 40)) return B200LC_ERR_UNSUPPORTED;
    const size_t need = b200lc_cuhd_decode_scratch_bytes(n_units);
    if (scratch_bytes < need) return B200LC_ERR_SCRATCH;
    if (reinterpret_cast<uintptr_t>(d_scratch) & 127) return B200LC_ERR_ARG;

    auto kern = cuhd::cuhd_decode_kernel<cuhd::kS, cuhd::kT, cuhd::kCap>;
    const size_t smem = cuhd::smem_bytes((u32)max_codeword_length);
    static int occ_cache[14] = {0};
    if (!occ_cache[max_codeword_length]) {
        B200LC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem));
        int occ = 0;
        B200LC_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, cuhd::kT + 32,
                                                                      smem));
        if (occ < 1) return B200LC_ERR_CUDA;
        occ_cache[max_codeword_length] = occ;
    }
    cuhd::DecodeParams p;
    p.units = d_units;
    p.n_units = n_units;
    p.lut = reinterpret_cast<const u16 *>(d_table);
    p.max_len = (u32)max_codeword_length;
    p.out = d_out;
    p.n_out = n_out;
    p.ticket = reinterpret_cast<u32 *>(d_scratch);
    p.desc = reinterpret_cast<cuhd::TileDesc *>(reinterpret_cast<char *>(d_scratch) + 128);
    p.num_tiles = cuhd::tiles_for(n_units);
    p.tma_ok_base = (reinterpret_cast<uintptr_t>(d_units) & 15) == 0;

    B200LC_CUDA_TRY(cudaMemsetAsync(d_scratch, 0, need, stream));
    const u32 grid = (u32)min((u64)p.num_tiles,
                              (u64)num_sms() * (u64)occ_cache[max_codeword_length]);
    kern<<<grid, cuhd::kT + 32, smem, stream>>>(p);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}
