// cudppCompress-compatible Huffman stage for many independent blocks
// (hot path 1, SURVEY.md 8a rows c5-c8, stream format appendix A.3).
//
// Bit-exact with the reference: 257-symbol alphabet (256 bytes + EOF with count 1), the explicit
// tree of huffman_build_tree_kernel with its FindMinimumCount tie-breaking and "move min1 to the
// next free slot" step (compress_kernel.cuh:2306-2392, compress_cta.cuh:550-571), codes from the
// left=0 / right=1 walk (:2416-2496), 4096-symbol blocks packed MSB-first into 32-bit words,
// stream = [nWords][words...] per block with word offsets (:2616-2618,2645-2706,2726-2747).
//
// What is different: the tree is built by one WARP per input block (minimum search = 9 slots per
// lane + shuffle reduction) instead of one thread of a <<<1,128>>> launch; block encoding shifts
// whole codes into 64-bit accumulators after a prefix sum of code lengths (the reference emits
// one bit at a time with atomicOr and scans with thread 0); block offsets come from a scan
// instead of every block summing all previous sizes (O(B^2), :2736-2741); no host round trip
// (the reference copies nCodesPacked back to the host between kernels, compress_app.cu:106).
#include "common.cuh"
#include "../../include/b200lc.h"

namespace b200lc {
namespace chuff {

constexpr int kSyms = 257;
constexpr int kEof = 256;
constexpr int kNodes = 2 * kSyms - 1;      // 513
constexpr int kBlockChars = 4096;          // HUFF_THREADS_PER_BLOCK * HUFF_WORK_PER_THREAD
constexpr int kBlockWordsMax = 1536;       // HUFF_CODE_BYTES (cudpp_globals.h:65-66)
constexpr int kTreeShorts = 3 * kNodes + 1;

// ---------------------------------------------------------------- histogram (u32[256] per block)
__global__ void __launch_bounds__(256) hist_kernel(const u8 *__restrict__ in, u32 n, u32 chunks,
                                                   u32 *__restrict__ hist)
{
    __shared__ u32 sh[8][256];
    const u32 tid = threadIdx.x, warp = tid >> 5;
    const u32 blk = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    for (u32 i = tid; i < 8 * 256; i += 256) (&sh[0][0])[i] = 0;
    __syncthreads();
    const u32 per = (n + chunks - 1) / chunks;
    const u32 lo = chunk * per, hi = min(n, lo + per);
    const u8 *src = in + (u64)blk * n;
    for (u32 i = lo + tid; i < hi; i += 256) atomicAdd(&sh[warp][src[i]], 1u);
    __syncthreads();
    u32 s = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sh[w][tid];
    if (s) atomicAdd(&hist[(u64)blk * 256 + tid], s);
}

// ---------------------------------------------------------------- tree + codes, one warp per block
struct TreeSmem {
    u32 count[kNodes];
    short level[kNodes];
    short left[kNodes], right[kNodes], parent[kNodes], value[kNodes];
    u8 ignore[kNodes];
};
constexpr int kTreeWarps = 2;

__device__ __forceinline__ int find_min(const TreeSmem &t, int n, u32 lane)
{
    // lexicographic minimum of (count, level, index) over the slots that are not ignored
    u32 bc = 0xffffffffu, bl = 0xffffu, bi = 0xffffu;
    for (int i = (int)lane; i < n; i += 32) {
        if (t.ignore[i]) continue;
        const u32 c = t.count[i], l = (u32)t.level[i];
        if (c < bc || (c == bc && l < bl)) { bc = c; bl = l; bi = (u32)i; }   // i increases: first index wins
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const u32 oc = __shfl_xor_sync(0xffffffffu, bc, d);
        const u32 ol = __shfl_xor_sync(0xffffffffu, bl, d);
        const u32 oi = __shfl_xor_sync(0xffffffffu, bi, d);
        if (oc < bc || (oc == bc && (ol < bl || (ol == bl && oi < bi)))) { bc = oc; bl = ol; bi = oi; }
    }
    return bi == 0xffffu ? -1 : (int)bi;
}

__global__ void __launch_bounds__(kTreeWarps * 32) tree_kernel(const u32 *__restrict__ hist, u32 nblocks,
                                                              u32 *__restrict__ codes,      // [nblocks][257]
                                                              u8 *__restrict__ lens,        // [nblocks][257]
                                                              short *__restrict__ tree_out, // optional, see cudpp_tree.cuh
                                                              u32 *__restrict__ error)
{
    __shared__ TreeSmem trees[kTreeWarps];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 blk = blockIdx.x * kTreeWarps + warp;
    if (blk >= nblocks) return;
    TreeSmem &t = trees[warp];
    const u32 *h = hist + (u64)blk * 256;
    for (int j = (int)lane; j < kNodes; j += 32) {
        t.count[j] = 0;
        t.level[j] = 0;
        t.left[j] = t.right[j] = t.parent[j] = -1;
        t.value[j] = (short)(j < kSyms ? j : 0);
        t.ignore[j] = 1;
    }
    __syncwarp();
    // leaves in increasing symbol order (compress_kernel.cuh:2306-2318): warp-wide compaction
    int n = 0;
    for (int base = 0; base < kSyms + 31; base += 32) {
        const int j = base + (int)lane;
        const u32 c = j < 256 ? h[j] : (j == kEof ? 1u : 0u);
        const u32 bal = __ballot_sync(0xffffffffu, c > 0);
        if (c > 0) {
            const int slot = n + __popc(bal & ((1u << lane) - 1));
            t.count[slot] = c;
            t.ignore[slot] = 0;
            t.value[slot] = (short)j;
        }
        n += __popc(bal);
    }
    __syncwarp();
    int next_free = n, head = -1;
    for (;;) {
        const int min1 = find_min(t, n, lane);
        head = min1;
        if (min1 < 0) break;
        __syncwarp();
        if (lane == 0) t.ignore[min1] = 1;
        __syncwarp();
        const int min2 = find_min(t, n, lane);
        if (min2 < 0) break;
        __syncwarp();
        if (lane == 0) {
            const int i = next_free;      // first slot >= n never used (its count is still 0)
            t.count[i] = t.count[min1];
            t.level[i] = t.level[min1];
            t.value[i] = t.value[min1];
            t.left[i] = t.left[min1];
            t.right[i] = t.right[min1];
            t.ignore[i] = 1;
            t.parent[i] = (short)min1;
            if (t.left[i] >= 0) t.parent[t.left[i]] = (short)i;
            if (t.right[i] >= 0) t.parent[t.right[i]] = (short)i;
            t.left[min1] = (short)i;
            t.ignore[min2] = 1;
            t.value[min1] = -1;           // composite
            t.ignore[min1] = 0;
            t.count[min1] += t.count[min2];
            t.level[min1] = (short)(max((int)t.level[min1], (int)t.level[min2]) + 1);
            t.right[min1] = (short)min2;
            t.parent[min2] = (short)min1;
            t.parent[min1] = -1;
        }
        ++next_free;
        __syncwarp();
    }
    if (tree_out) {   // explicit tree for the decoder: left[513], right[513], value[513], head
        short *o = tree_out + (u64)blk * kTreeShorts;
        for (int j = (int)lane; j < kNodes; j += 32) {
            o[j] = t.left[j];
            o[kNodes + j] = t.right[j];
            o[2 * kNodes + j] = t.value[j];
        }
        if (lane == 0) o[3 * kNodes] = (short)head;
    }
    // codes: depth-first walk, left = 0, right = 1 (compress_kernel.cuh:2416-2496)
    u32 *c_out = codes + (u64)blk * kSyms;
    u8 *l_out = lens + (u64)blk * kSyms;
    for (int j = (int)lane; j < kSyms; j += 32) { c_out[j] = 0; l_out[j] = 0; }
    __syncwarp();
    if (lane == 0) {
        int cur = head, depth = 0;
        u64 path = 0;
        bool too_long = false;
        for (;;) {
            while (t.left[cur] != -1) { path <<= 1; cur = t.left[cur]; ++depth; }
            if (t.value[cur] != -1) {
                if (depth > 32) too_long = true;
                c_out[t.value[cur]] = (u32)path;
                l_out[t.value[cur]] = (u8)depth;
            }
            while (t.parent[cur] != -1) {
                if (cur != t.right[t.parent[cur]]) { path |= 1; cur = t.right[t.parent[cur]]; break; }
                --depth; path >>= 1; cur = t.parent[cur];
            }
            if (t.parent[cur] == -1) break;
        }
        if (too_long) atomicExch(error, 1u);
    }
}

// ---------------------------------------------------------------- bits per 4096-symbol block
__global__ void __launch_bounds__(128) bits_kernel(const u8 *__restrict__ in, u32 n, u32 nhb,
                                                   const u8 *__restrict__ lens,
                                                   u32 *__restrict__ nwords, u32 *__restrict__ error)
{
    __shared__ u8 sl[256];
    __shared__ u32 wsum[4];
    const u32 tid = threadIdx.x;
    const u32 blk = blockIdx.x / nhb, hb = blockIdx.x % nhb;
    sl[tid] = lens[(u64)blk * kSyms + tid];
    sl[tid + 128] = lens[(u64)blk * kSyms + tid + 128];
    __syncthreads();
    const u32 lo = hb * kBlockChars + tid * 32;
    const u8 *src = in + (u64)blk * n;
    u32 bits = 0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i)
        if (lo + i < n) bits += sl[src[lo + i]];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, d);
    if ((tid & 31) == 0) wsum[tid >> 5] = bits;
    __syncthreads();
    if (tid == 0) {
        const u32 total = wsum[0] + wsum[1] + wsum[2] + wsum[3];
        const u32 nw = (total + 31) >> 5;
        nwords[blockIdx.x] = nw;
        if (nw > kBlockWordsMax) atomicExch(error, 2u);   // the reference would overrun encoded.code[]
    }
}

// ---------------------------------------------------------------- offsets: one CTA per input block
__global__ void __launch_bounds__(256) offsets_kernel(const u32 *__restrict__ nwords, u32 nhb,
                                                      u32 *__restrict__ offsets,      // [nblocks][nhb]
                                                      u32 *__restrict__ total_words,  // [nblocks]
                                                      u32 *__restrict__ out, u64 out_stride_words)
{
    __shared__ u32 sums[256];
    __shared__ u32 carry_s;
    const u32 blk = blockIdx.x, tx = threadIdx.x;
    if (tx == 0) carry_s = 0;
    __syncthreads();
    for (u32 base = 0; base < nhb; base += 256) {
        const u32 i = base + tx;
        const u32 v = i < nhb ? 1 + nwords[(u64)blk * nhb + i] : 0;
        sums[tx] = v;
        __syncthreads();
        for (u32 d = 1; d < 256; d <<= 1) {
            const u32 a = tx >= d ? sums[tx - d] : 0;
            __syncthreads();
            sums[tx] += a;
            __syncthreads();
        }
        const u32 carry = carry_s;
        __syncthreads();
        if (i < nhb) {
            const u32 off = carry + sums[tx] - v;
            offsets[(u64)blk * nhb + i] = off;
            if (off < out_stride_words) out[(u64)blk * out_stride_words + off] = v - 1;   // the [nWords] cell
        }
        if (tx == 255) carry_s = carry + sums[255];
        __syncthreads();
    }
    if (tx == 0) total_words[blk] = carry_s;
}

// ---------------------------------------------------------------- encode one 4096-symbol block per CTA
__global__ void __launch_bounds__(128) encode_kernel(const u8 *__restrict__ in, u32 n, u32 nhb,
                                                     const u32 *__restrict__ codes,
                                                     const u8 *__restrict__ lens,
                                                     const u32 *__restrict__ offsets,
                                                     const u32 *__restrict__ nwords,
                                                     u32 *__restrict__ out, u64 out_stride_words,
                                                     u32 *__restrict__ error)
{
    __shared__ u32 scode[256];
    __shared__ u8 slen[256];
    __shared__ u32 stage[kBlockWordsMax + 2];
    __shared__ u32 wsum[4];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 blk = blockIdx.x / nhb, hb = blockIdx.x % nhb;
    const u32 nw = nwords[blockIdx.x];
    const u32 off = offsets[blockIdx.x];
    if (nw > kBlockWordsMax || (u64)off + 1 + nw > out_stride_words) {
        if (tid == 0) atomicExch(error, 3u);
        return;
    }
    for (u32 i = tid; i < 256; i += 128) {
        scode[i] = codes[(u64)blk * kSyms + i];
        slen[i] = lens[(u64)blk * kSyms + i];
    }
    for (u32 i = tid; i < nw + 1; i += 128) stage[i] = 0;
    __syncthreads();
    const u32 lo = hb * kBlockChars + tid * 32;
    const u8 *src = in + (u64)blk * n + lo;
    const u32 valid = lo < n ? min(32u, n - lo) : 0u;
    u32 my_bits = 0;
    for (u32 i = 0; i < valid; ++i) my_bits += slen[src[i]];
    u32 incl = warp_incl_scan(my_bits);
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    u32 pre = incl - my_bits;
#pragma unroll
    for (int w = 0; w < 4; ++w) pre += (u32)w < warp ? wsum[w] : 0u;
    {
        u32 wi = pre >> 5;
        u32 nb = pre & 31;
        u64 acc = 0;
        bool first_word = true;
        for (u32 i = 0; i < valid; ++i) {
            const u32 s = src[i];
            const u32 len = slen[s];
            acc = (acc << len) | scode[s];
            nb += len;
            if (nb >= 32) {
                const u32 word = (u32)(acc >> (nb - 32));
                if (first_word) { atomicOr(&stage[wi], word); first_word = false; }
                else stage[wi] = word;
                ++wi;
                nb -= 32;
            }
        }
        if (nb && my_bits) atomicOr(&stage[wi], (u32)(acc << (32 - nb)));
    }
    __syncthreads();
    u32 *dst = out + (u64)blk * out_stride_words + off + 1;
    for (u32 i = tid; i < nw; i += 128) dst[i] = stage[i];
}

// Used by the decoder (cudpp_decode.cu): same tree, same tie-breaks, from the stored histogram.
cudaError_t launch_tree(const u32 *hist, u32 nblocks, u32 *codes, u8 *lens, short *tree_out, u32 *error,
                        cudaStream_t stream)
{
    tree_kernel<<<(nblocks + kTreeWarps - 1) / kTreeWarps, kTreeWarps * 32, 0, stream>>>(hist, nblocks, codes,
                                                                                        lens, tree_out, error);
    return cudaGetLastError();
}

}  // namespace chuff
}  // namespace b200lc

using namespace b200lc;

extern "C" size_t b200lc_cudpp_huffman_scratch_bytes(size_t nblocks, size_t n)
{
    const size_t nhb = (n + chuff::kBlockChars - 1) / chuff::kBlockChars;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    return up(nblocks * chuff::kSyms * 4) + up(nblocks * chuff::kSyms) + up(nblocks * nhb * 4) + 256;
}

// d_mtf: nblocks blocks of n symbols.  Outputs per block b: d_hist[b*256..], d_offsets[b*nhb..],
// d_total_words[b], stream at d_out + b * out_stride_words.  *d_error (device u32) becomes
// non-zero on a code longer than 32 bits / a block beyond 1536 words / out_stride too small.
extern "C" int b200lc_cudpp_huffman_batch(const uint8_t *d_mtf, size_t nblocks, size_t n,
                                          uint32_t *d_hist, uint32_t *d_offsets,
                                          uint32_t *d_total_words, uint32_t *d_out,
                                          size_t out_stride_words, uint32_t *d_error,
                                          void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nblocks == 0 || n == 0) return B200LC_OK;
    if (!d_mtf || !d_hist || !d_offsets || !d_total_words || !d_out || !d_error || !d_scratch)
        return B200LC_ERR_ARG;
    if (n >= (1ull << 31) || (reinterpret_cast<uintptr_t>(d_scratch) & 255)) return B200LC_ERR_ARG;
    if (scratch_bytes < b200lc_cudpp_huffman_scratch_bytes(nblocks, n)) return B200LC_ERR_SCRATCH;
    const u32 nhb = (u32)((n + chuff::kBlockChars - 1) / chuff::kBlockChars);
    if ((u64)nblocks * nhb >= (1ull << 31)) return B200LC_ERR_UNSUPPORTED;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    char *s = reinterpret_cast<char *>(d_scratch);
    u32 *codes = reinterpret_cast<u32 *>(s);
    u8 *lens = reinterpret_cast<u8 *>(s + up(nblocks * chuff::kSyms * 4));
    u32 *nwords = reinterpret_cast<u32 *>(s + up(nblocks * chuff::kSyms * 4) + up(nblocks * chuff::kSyms));

    B200LC_CUDA_TRY(cudaMemsetAsync(d_hist, 0, nblocks * 256 * sizeof(u32), stream));
    B200LC_CUDA_TRY(cudaMemsetAsync(d_error, 0, sizeof(u32), stream));
    const u32 chunks = (u32)max((size_t)1, min((size_t)64, n / 16384));
    chuff::hist_kernel<<<(u32)(nblocks * chunks), 256, 0, stream>>>(d_mtf, (u32)n, chunks, d_hist);
    B200LC_CUDA_TRY(cudaGetLastError());
    B200LC_CUDA_TRY(chuff::launch_tree(d_hist, (u32)nblocks, codes, lens, nullptr, d_error, stream));
    chuff::bits_kernel<<<(u32)(nblocks * nhb), 128, 0, stream>>>(d_mtf, (u32)n, nhb, lens, nwords, d_error);
    B200LC_CUDA_TRY(cudaGetLastError());
    chuff::offsets_kernel<<<(u32)nblocks, 256, 0, stream>>>(nwords, nhb, d_offsets, d_total_words, d_out,
                                                           out_stride_words);
    B200LC_CUDA_TRY(cudaGetLastError());
    chuff::encode_kernel<<<(u32)(nblocks * nhb), 128, 0, stream>>>(d_mtf, (u32)n, nhb, codes, lens,
                                                                  d_offsets, nwords, d_out,
                                                                  out_stride_words, d_error);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}
