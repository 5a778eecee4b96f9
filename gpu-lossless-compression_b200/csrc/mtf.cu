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
//   2. one THREAD per segment then runs the sequential transform on its own list, kept in
//      shared memory in transposed layout (list[j][thread]) so that lanes hit different banks.
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
constexpr int kApplyThreads = 64;

// The list of a thread is 64 little-endian words (byte 0 of word 0 = front) stored transposed,
// W[w * 64 + t].  One pass per input byte: every word in front of the match is shifted up by one
// byte while it is being searched (4 list entries per shared-memory access, __vcmpeq4 finds the
// match), so the cost is rank/4 iterations instead of 2*rank byte moves.
__global__ void __launch_bounds__(kApplyThreads) mtf_apply_kernel(const u8 *__restrict__ in, u32 n,
                                                                  u32 nseg, u64 total_segs,
                                                                  const u8 *__restrict__ lists,
                                                                  u8 *__restrict__ out)
{
    __shared__ u32 W[64 * kApplyThreads];
    const u32 t = threadIdx.x;
    const u64 seg = (u64)blockIdx.x * kApplyThreads + t;
    if (seg >= total_segs) return;
    const u32 blk = (u32)(seg / nseg), s = (u32)(seg % nseg);
    {
        const uint4 *lsrc = reinterpret_cast<const uint4 *>(lists + seg * 256);
#pragma unroll 4
        for (int q = 0; q < 16; ++q) {
            const uint4 v = lsrc[q];
            W[(4 * q + 0) * kApplyThreads + t] = v.x;
            W[(4 * q + 1) * kApplyThreads + t] = v.y;
            W[(4 * q + 2) * kApplyThreads + t] = v.z;
            W[(4 * q + 3) * kApplyThreads + t] = v.w;
        }
    }
    const u64 base = (u64)blk * n + (u64)s * kSeg;
    const u32 len = min(kSeg, n - s * kSeg);
    const u8 *src = in + base;
    u8 *dst = out + base;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    u32 front = W[t];                      // word 0 lives in a register
    auto step = [&](u32 c) -> u32 {
        const u32 cc = c * 0x01010101u;
        u32 m = __vcmpeq4(front, cc);
        if (m) {                           // rank 0..3: registers only
            const u32 b = (__ffs(m) - 1) >> 3;
            const u32 lomask = b == 3 ? 0xffffffffu : ((1u << (8 * (b + 1))) - 1u);
            front = (front & ~lomask) | (((front << 8) | c) & lomask);
            return b;
        }
        u32 carry = front >> 24;
        front = (front << 8) | c;
        u32 w = 1;
        while (true) {
            const u32 x = W[w * kApplyThreads + t];
            m = __vcmpeq4(x, cc);
            if (m) {
                const u32 b = (__ffs(m) - 1) >> 3;
                const u32 lomask = b == 3 ? 0xffffffffu : ((1u << (8 * (b + 1))) - 1u);
                W[w * kApplyThreads + t] = (x & ~lomask) | (((x << 8) | carry) & lomask);
                return 4 * w + b;
            }
            W[w * kApplyThreads + t] = (x << 8) | carry;
            carry = x >> 24;
            ++w;
        }
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
                for (int k = 0; k < 4; ++k) o |= step((w[q] >> (8 * k)) & 0xffu) << (8 * k);
                w[q] = o;
            }
            *reinterpret_cast<uint4 *>(dst + i) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    for (; i < len; ++i) dst[i] = (u8)step(src[i]);
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
    mtf::mtf_apply_kernel<<<grid, mtf::kApplyThreads, 0, stream>>>(d_in, (u32)n, nseg, total, lists, d_out);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}
