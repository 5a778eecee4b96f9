// Move-to-front transform of many independent blocks  (hot path 1).
//
// Exact MTF with initial list 0..255 (gold: apps/cudpp_testrig/test_compress.cpp:93-125).
// The reference computes a 256-byte partial list per 64 input bytes and combines them with a
// four-kernel up-sweep / down-sweep (compress_kernel.cuh:1339-2023; 4 MiB of list scratch per
// MiB of data, 64-thread CTAs).  Here:
//   1. one CTA per block walks the block's 2 KiB segments in order and maintains the list state
//      at every segment boundary incrementally: the symbols present in a segment, ordered by
//      their last position (found with shared-memory atomicMax + a position bitmap), go to the
//      front; the others keep their order (ballot-based compaction).  Scratch = 256 B per 2 KiB.
//   2. one THREAD per segment then runs the sequential transform with constant work per byte:
//      the rank is a prefix count over a bitmap of last occurrences (see mtf_apply_kernel).
#include "common.cuh"
#include "../../include/b200lc.h"

namespace b200lc {
namespace mtf {

constexpr u32 kSeg = 2048;

// ---------------------------------------------------------------- 1. list state per segment
__global__ void __launch_bounds__(256) mtf_lists_kernel(const u8 *__restrict__ in, u32 n, u32 nseg,
                                                       u8 *__restrict__ lists)
{
    __shared__ u8 list[256], newlist[256];
    __shared__ int lp[256];
    __shared__ u32 bitmap[kSeg / 32];
    __shared__ u32 above[kSeg / 32];   // set bits in words with a higher index
    __shared__ u32 wcount[8];
    __shared__ u32 npresent_s;
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 blk = blockIdx.x;
    const u8 *src = in + (u64)blk * n;
    u8 *dst = lists + (u64)blk * nseg * 256;
    list[tid] = (u8)tid;
    __syncthreads();
    for (u32 s = 0; s < nseg; ++s) {
        dst[(u64)s * 256 + tid] = list[tid];
        if (s + 1 == nseg) break;
        lp[tid] = -1;
        if (tid < kSeg / 32) bitmap[tid] = 0;
        __syncthreads();
        {   // last position of every symbol inside segment s (a full segment: s is not the last one)
            const u32 base = s * kSeg + tid * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicMax(&lp[src[base + k]], (int)(tid * 8 + k));
        }
        __syncthreads();
        const int p = lp[tid];
        if (p >= 0) atomicOr(&bitmap[p >> 5], 1u << (p & 31));
        __syncthreads();
        if (tid < 64) {
            // above[w] = number of marked positions in words w+1 .. 63 (suffix sum over 64 words)
            const u32 c = __popc(bitmap[tid]);
            u32 incl = c;   // inclusive suffix scan inside each 32-word half
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const u32 t = __shfl_down_sync(0xffffffffu, incl, d);
                if (lane + d < 32) incl += t;
            }
            if (lane == 0) wcount[warp] = incl;
            __syncwarp();
            above[tid] = incl - c;
        }
        __syncthreads();
        if (tid < 32) above[tid] += wcount[1];
        if (tid == 0) npresent_s = wcount[0] + wcount[1];
        __syncthreads();
        const u32 npresent = npresent_s;
        if (p >= 0) {
            const u32 w = (u32)p >> 5, b = (u32)p & 31;
            const u32 higher = b == 31 ? 0u : __popc(bitmap[w] >> (b + 1));
            newlist[above[w] + higher] = (u8)tid;          // more recent symbols first
        }
        // symbols absent from the segment keep their relative order behind the present ones
        const u32 sym = list[tid];
        const bool absent = lp[sym] < 0;
        const u32 bal = __ballot_sync(0xffffffffu, absent);
        if (lane == 0) wcount[warp] = __popc(bal);
        __syncthreads();
        u32 before = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) before += (u32)w < warp ? wcount[w] : 0u;
        if (absent) newlist[npresent + before + __popc(bal & ((1u << lane) - 1))] = (u8)sym;
        __syncthreads();
        list[tid] = newlist[tid];
        __syncthreads();
    }
}

// ---------------------------------------------------------------- 2. sequential MTF per segment
// One THREAD per segment, constant work per input byte.  Round 1 searched the list itself (rank/4
// shared-memory read-modify-writes per byte): on data with large ranks a warp paid for the largest
// rank among its 32 lanes at every step (4.5 ms per 128 MiB of Zipf(1.3) bytes, 0.009 of the HBM
// roofline).  Round 2 never touches a list:
//     rank(i) = number of symbols whose LAST occurrence lies behind the last occurrence of in[i]
// Every symbol's last occurrence is one marked bit in a bitmap over "time": indices 0..255 are the
// initial list (front = 255), index 256 + i is position i of the segment.  Exactly 256 bits are
// marked at any time, so rank = 255 - (marked bits below last[c]), and that prefix count comes from
// three levels of counters (16-word groups, 4-word groups, words: at most 4 + 3 + 3 independent
// reads, no loop).  All per-lane state is laid out [index][lane]: one bank per lane, no conflicts.
constexpr int kApplyThreads = 64;
constexpr int kBitWords = (256 + kSeg) / 32;          // 72
constexpr int kQuadWords = (kBitWords / 4 + 3) / 4;   // 18 byte counters in 5 words
constexpr int kHexWords = ((kBitWords + 15) / 16 + 1) / 2;   // 5 u16 counters in 3 words
struct ApplyWarp {
    u32 last[128][32];                // u16 per symbol: index of its last occurrence
    u32 bits[kBitWords][32];
    u32 quad[kQuadWords][32];         // u8: marked bits per 4 words
    u32 hex[kHexWords][32];           // u16: marked bits per 16 words
};

__global__ void __launch_bounds__(kApplyThreads) mtf_apply_kernel(const u8 *__restrict__ in, u32 n,
                                                                  u32 nseg, u64 total_segs,
                                                                  const u8 *__restrict__ lists,
                                                                  u8 *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ApplyWarp &W = reinterpret_cast<ApplyWarp *>(smem_raw)[threadIdx.x >> 5];
    const u32 lane = threadIdx.x & 31;
    const u64 seg = (u64)blockIdx.x * kApplyThreads + threadIdx.x;
    if (seg >= total_segs) return;
    const u32 blk = (u32)(seg / nseg), s = (u32)(seg % nseg);
    u8 *const lastb = reinterpret_cast<u8 *>(&W.last[0][lane]);    // u16 of symbol c at lastb + (c >> 1) * 128 + (c & 1) * 2
    u8 *const quadb = reinterpret_cast<u8 *>(&W.quad[0][lane]);    // u8 of group q  at quadb + (q >> 2) * 128 + (q & 3)
    u8 *const hexb = reinterpret_cast<u8 *>(&W.hex[0][lane]);      // u16 of group h at hexb + (h >> 1) * 128 + (h & 1) * 2
    auto last_at = [&](u32 c) -> u16 * { return reinterpret_cast<u16 *>(lastb + ((c >> 1) << 7) + ((c & 1) << 1)); };
    auto quad_at = [&](u32 q) -> u8 * { return quadb + ((q >> 2) << 7) + (q & 3); };
    auto hex_at = [&](u32 h) -> u16 * { return reinterpret_cast<u16 *>(hexb + ((h >> 1) << 7) + ((h & 1) << 1)); };
    {
        const uint4 *lsrc = reinterpret_cast<const uint4 *>(lists + seg * 256);
#pragma unroll 1
        for (u32 q = 0; q < 16; ++q) {
            const uint4 v = lsrc[q];
            const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (u32 k = 0; k < 16; ++k) {
                const u32 sym = (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
                *last_at(sym) = (u16)(255u - (16 * q + k));
            }
        }
#pragma unroll 1
        for (u32 k = 0; k < (u32)kBitWords; ++k) W.bits[k][lane] = k < 8 ? 0xffffffffu : 0u;
#pragma unroll
        for (u32 k = 0; k < (u32)kQuadWords; ++k) W.quad[k][lane] = k == 0 ? 0x00008080u : 0u;   // groups 0, 1: 128 bits each
#pragma unroll
        for (u32 k = 0; k < (u32)kHexWords; ++k) W.hex[k][lane] = k == 0 ? 256u : 0u;
    }
    const u64 base = (u64)blk * n + (u64)s * kSeg;
    const u32 len = min(kSeg, n - s * kSeg);
    const u8 *src = in + base;
    u8 *dst = out + base;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    auto step = [&](u32 c, u32 i) -> u32 {
        u16 *const lp = last_at(c);
        const u32 pi = *lp;
        const u32 wpi = pi >> 5, bpi = pi & 31u, qd = wpi >> 2, h = wpi >> 4;
        u32 below = 0;
#pragma unroll
        for (u32 j = 0; j < 4; ++j) below += j < h ? (u32)*hex_at(j) : 0u;
#pragma unroll
        for (u32 j = 0; j < 3; ++j) below += 4 * h + j < qd ? (u32)*quad_at(4 * h + j) : 0u;
#pragma unroll
        for (u32 j = 0; j < 3; ++j) below += 4 * qd + j < wpi ? (u32)__popc(W.bits[4 * qd + j][lane]) : 0u;
        const u32 wold = W.bits[wpi][lane];
        below += __popc(wold & ((1u << bpi) - 1u));
        // the mark moves from pi to 256 + i
        const u32 cur = 256u + i, wc = cur >> 5;
        W.bits[wpi][lane] = wold & ~(1u << bpi);
        *quad_at(qd) -= 1;
        *hex_at(h) -= 1;
        W.bits[wc][lane] |= 1u << (cur & 31u);
        *quad_at(wc >> 2) += 1;
        *hex_at(wc >> 4) += 1;
        *lp = (u16)cur;
        return 255u - below;
    };
    u32 i = 0;
    if (aligned) {
        for (; i + 16 <= len; i += 16) {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + i);
            u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                u32 o = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) o |= step((w[q] >> (8 * k)) & 0xffu, i + 4 * q + k) << (8 * k);
                w[q] = o;
            }
            *reinterpret_cast<uint4 *>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    for (; i < len; ++i) dst[i] = (u8)step(src[i], i);
}

}  // namespace mtf
}  // namespace b200lc

using namespace b200lc;

extern "C" size_t b200lc_mtf_scratch_bytes(size_t nblocks, size_t n)
{
    const size_t nseg = (n + mtf::kSeg - 1) / mtf::kSeg;
    return nblocks * nseg * 256 + 256;
}

extern "C" int b200lc_mtf_batch(const uint8_t *d_in, size_t nblocks, size_t n, uint8_t *d_out,
                                void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nblocks == 0 || n == 0) return B200LC_OK;
    if (!d_in || !d_out || !d_scratch) return B200LC_ERR_ARG;
    if (n >= (1ull << 31) || (reinterpret_cast<uintptr_t>(d_scratch) & 15)) return B200LC_ERR_ARG;
    if (scratch_bytes < b200lc_mtf_scratch_bytes(nblocks, n)) return B200LC_ERR_SCRATCH;
    const u32 nseg = (u32)((n + mtf::kSeg - 1) / mtf::kSeg);
    u8 *lists = reinterpret_cast<u8 *>(d_scratch);
    mtf::mtf_lists_kernel<<<(u32)nblocks, 256, 0, stream>>>(d_in, (u32)n, nseg, lists);
    B200LC_CUDA_TRY(cudaGetLastError());
    const u64 total = (u64)nblocks * nseg;
    const u32 grid = (u32)((total + mtf::kApplyThreads - 1) / mtf::kApplyThreads);
    const size_t apply_smem = sizeof(mtf::ApplyWarp) * (mtf::kApplyThreads / 32);
    static unsigned attr_done[kMaxDevices] = {0};   // context epoch the attribute was set in
    const int slot = device_slot();
    if (slot < 0 || attr_done[slot] != context_epoch()) {
        B200LC_CUDA_TRY(cudaFuncSetAttribute(mtf::mtf_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)apply_smem));
        if (slot >= 0) attr_done[slot] = context_epoch();
    }
    mtf::mtf_apply_kernel<<<grid, mtf::kApplyThreads, apply_smem, stream>>>(d_in, (u32)n, nseg, total, lists, d_out);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}
