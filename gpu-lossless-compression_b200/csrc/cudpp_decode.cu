// Decoder for the cudppCompress stream, batched over independent blocks.
//
// The reference has no GPU decoder: its only decoder is the CPU gold of the test rig
// (apps/cudpp_testrig/test_compress.cpp:192-364: bit-by-bit tree walk per 4096-symbol block,
// sequential inverse MTF, inverse BWT through a radix sort of (byte, index) and a sequential
// n-step pointer walk).  This file does the three inverse stages on the GPU:
//   1. Huffman: the tree is rebuilt from the stored histogram by the encoder's tree kernel (same
//      tie-breaks), one THREAD decodes one 4096-symbol block through a 10-bit first-level table in
//      shared memory (longer codes continue bit by bit in the tree);
//   2. inverse MTF: one thread per 2 KiB segment runs the transform on the IDENTITY list and emits
//      "position ids" plus the segment's permutation; a chain kernel composes the permutations
//      into the list at every segment start; a map kernel turns ids into symbols;
//   3. inverse BWT: T = stable sort of each block's bytes -> index (one 8-bit pass of the
//      segmented one-sweep sort in devprims.cu), packed with the first-column byte into one word
//      per row; the n-step walk
//      idx <- T[idx] is cut at "splitter" rows (every gap-th row + the start row): every splitter
//      walks to the next one in parallel, one thread per block ranks the <= 4097 splitters in
//      shared memory, every splitter re-walks its piece writing output bytes (sparse ruling set
//      list ranking).  A walk that closes its cycle before n steps (periodic block) is extended
//      by out[i] = out[i mod period].
#include "common.cuh"
#include "../../include/b200lc.h"
#include "devprims.cuh"

namespace b200lc {
namespace chuff {
constexpr int kSyms = 257;
constexpr int kEof = 256;
constexpr int kNodes = 2 * kSyms - 1;
constexpr int kBlockChars = 4096;
constexpr int kTreeShorts = 3 * kNodes + 1;
cudaError_t launch_tree(const u32 *hist, u32 nblocks, u32 *codes, u8 *lens, short *tree_out, u32 *error,
                        cudaStream_t stream);
}  // namespace chuff

namespace cdec {

constexpr u32 kSeg = 2048;
constexpr int kLutBits = 10;
constexpr int kHdecThreads = 128;

// ---------------------------------------------------------------- 1. Huffman, thread per 4096-symbol block
__global__ void __launch_bounds__(kHdecThreads) hdec_kernel(const u32 *__restrict__ comp, u64 comp_stride_words,
                                                            const u32 *__restrict__ offsets, u32 n, u32 nhb,
                                                            u32 groups, const short *__restrict__ trees,
                                                            u8 *__restrict__ out, u32 *__restrict__ error)
{
    __shared__ u32 lut[1 << kLutBits];
    __shared__ short left[chuff::kNodes], right[chuff::kNodes], value[chuff::kNodes];
    const u32 tid = threadIdx.x;
    const u32 blk = blockIdx.x / groups, g = blockIdx.x % groups;
    const short *tr = trees + (u64)blk * chuff::kTreeShorts;
    for (u32 j = tid; j < (u32)chuff::kNodes; j += kHdecThreads) {
        left[j] = tr[j];
        right[j] = tr[chuff::kNodes + j];
        value[j] = tr[2 * chuff::kNodes + j];
    }
    const int head = tr[3 * chuff::kNodes];
    __syncthreads();
    for (u32 idx = tid; idx < (1u << kLutBits); idx += kHdecThreads) {
        int cur = head;
        u32 e = 0xffffffffu;   // walks off the tree: corrupt stream
        for (int k = 0; k < kLutBits && cur >= 0; ++k) {
            cur = ((idx >> (kLutBits - 1 - k)) & 1) ? right[cur] : left[cur];
            if (cur >= 0 && value[cur] != -1) { e = (u32)value[cur] | ((u32)(k + 1) << 16); break; }
        }
        if (e == 0xffffffffu && cur >= 0) e = (u32)cur;   // internal node, length field 0
        lut[idx] = e;
    }
    __syncthreads();
    const u32 hb = g * kHdecThreads + tid;
    if (hb >= nhb) return;
    const u32 off = offsets[(u64)blk * nhb + hb];
    if ((u64)off + 1 > comp_stride_words) { atomicExch(error, 4u); return; }
    const u32 *w = comp + (u64)blk * comp_stride_words + off;
    u32 nw = w[0];
    if ((u64)off + 1 + nw > comp_stride_words) { atomicExch(error, 4u); return; }
    ++w;
    const u32 want = min((u32)chuff::kBlockChars, n - hb * chuff::kBlockChars);
    u8 *dst = out + (u64)blk * n + (u64)hb * chuff::kBlockChars;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 3) == 0;
    u64 buf = 0;      // next bit of the stream = bit 63
    int avail = 0;
    u32 wi = 0, found = 0, word = 0;
    u64 consumed = 0;
    const u64 total_bits = (u64)nw * 32;
    bool bad = false;
    while (found < want) {
        if (avail < 32) {
            const u32 x = wi < nw ? __ldg(w + wi) : 0u;
            ++wi;
            buf |= (u64)x << (32 - avail);
            avail += 32;
        }
        u32 e = lut[(u32)(buf >> (64 - kLutBits))];
        u32 len = e >> 16;
        u32 sym = e & 0xffffu;
        if (len == 0) {   // code longer than the table: continue in the tree (avail >= 32 >= code length)
            if (e == 0xffffffffu) { bad = true; break; }
            int cur = (int)sym;
            len = kLutBits;
            for (;;) {
                if (len >= 32) { bad = true; break; }
                cur = ((buf >> (63 - len)) & 1) ? right[cur] : left[cur];
                ++len;
                if (cur < 0) { bad = true; break; }
                if (value[cur] != -1) { sym = (u32)value[cur]; break; }
            }
            if (bad) break;
        } else if (len > 32) { bad = true; break; }
        consumed += len;
        if (sym == (u32)chuff::kEof || consumed > total_bits) { bad = true; break; }
        buf <<= len;
        avail -= (int)len;
        if (aligned) {
            word |= sym << (8 * (found & 3));
            if ((found & 3) == 3) { *reinterpret_cast<u32 *>(dst + found - 3) = word; word = 0; }
        } else {
            dst[found] = (u8)sym;
        }
        ++found;
    }
    if (aligned && (found & 3))
        for (u32 k = 0; k < (found & 3); ++k) dst[(found & ~3u) + k] = (u8)(word >> (8 * k));
    if (bad) atomicExch(error, 5u);
}

// ---------------------------------------------------------------- 2. inverse MTF
constexpr int kPidThreads = 64;

// Thread per segment: inverse MTF from the identity list.  in/out may alias.
__global__ void __launch_bounds__(kPidThreads) imtf_pid_kernel(const u8 *in, u32 n, u32 nseg, u64 total_segs,
                                                               u8 *out, u8 *__restrict__ perms)
{
    __shared__ u32 W[64 * kPidThreads];
    const u32 t = threadIdx.x;
    const u64 seg = (u64)blockIdx.x * kPidThreads + t;
    if (seg >= total_segs) return;
    const u32 blk = (u32)(seg / nseg), s = (u32)(seg % nseg);
#pragma unroll 8
    for (u32 w = 1; w < 64; ++w) W[w * kPidThreads + t] = 0x03020100u + 0x04040404u * w;
    u32 front = 0x03020100u;
    const u64 base = (u64)blk * n + (u64)s * kSeg;
    const u32 len = min(kSeg, n - s * kSeg);
    const u8 *src = in + base;
    u8 *dst = out + base;
    const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    auto step = [&](u32 r) -> u32 {
        const u32 wq = r >> 2, b = r & 3;
        const u32 lomask = b == 3 ? 0xffffffffu : ((1u << (8 * (b + 1))) - 1u);
        if (wq == 0) {
            const u32 c = (front >> (8 * b)) & 0xffu;
            front = (front & ~lomask) | (((front << 8) | c) & lomask);
            return c;
        }
        const u32 x = W[wq * kPidThreads + t];
        const u32 c = (x >> (8 * b)) & 0xffu;
        u32 carry = front >> 24;
        front = (front << 8) | c;
        for (u32 w = 1; w < wq; ++w) {
            const u32 y = W[w * kPidThreads + t];
            W[w * kPidThreads + t] = (y << 8) | carry;
            carry = y >> 24;
        }
        W[wq * kPidThreads + t] = (x & ~lomask) | (((x << 8) | carry) & lomask);
        return c;
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
    u32 *p = reinterpret_cast<u32 *>(perms + seg * 256);
    p[0] = front;
#pragma unroll 8
    for (u32 w = 1; w < 64; ++w) p[w] = W[w * kPidThreads + t];
}

// CTA per block: list at the start of every segment, written over the segment's permutation.
__global__ void __launch_bounds__(256) imtf_chain_kernel(u8 *__restrict__ perms, u32 nseg)
{
    __shared__ u8 list[256];
    const u32 tid = threadIdx.x;
    u8 *p = perms + (u64)blockIdx.x * nseg * 256;
    list[tid] = (u8)tid;
    u32 q = p[tid];
    __syncthreads();
    for (u32 s = 0; s < nseg; ++s) {
        const u32 cur = list[tid];
        p[(u64)s * 256 + tid] = (u8)cur;
        if (s + 1 == nseg) break;
        const u32 nq = s + 2 < nseg ? p[(u64)(s + 1) * 256 + tid] : 0u;   // prefetch the next permutation
        const u32 nv = list[q];
        __syncthreads();
        list[tid] = (u8)nv;
        q = nq;
        __syncthreads();
    }
}

// CTA per segment: ids -> symbols (in place) and the (block, byte) sort keys of the inverse BWT.
__global__ void __launch_bounds__(128) imtf_map_kernel(u8 *__restrict__ data, u32 n, u32 nseg,
                                                       const u8 *__restrict__ lists, u32 *__restrict__ keys,
                                                       u32 *__restrict__ vals)
{
    __shared__ u8 list[256];
    const u32 tid = threadIdx.x;
    const u32 blk = blockIdx.x / nseg, s = blockIdx.x % nseg;
    list[tid] = lists[(u64)blockIdx.x * 256 + tid];
    list[tid + 128] = lists[(u64)blockIdx.x * 256 + tid + 128];
    __syncthreads();
    const u32 lo = s * kSeg;
    const u32 len = min(kSeg, n - lo);
    const u64 base = (u64)blk * n + lo;
    for (u32 i = tid; i < len; i += 128) {
        const u32 c = list[data[base + i]];
        data[base + i] = (u8)c;
        if (keys) {
            keys[base + i] = (blk << 8) | c;
            vals[base + i] = lo + i;
        }
    }
}

// ---------------------------------------------------------------- 3. inverse BWT
__global__ void __launch_bounds__(256) ibwt_keys_kernel(const u8 *__restrict__ bwt, u64 N, u32 n,
                                                        u32 *__restrict__ keys, u32 *__restrict__ vals)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const u32 blk = (u32)(i / n);
    keys[i] = (blk << 8) | bwt[i];
    vals[i] = (u32)(i - (u64)blk * n);
}

// One entry per row of the sorted first column: the row the walk goes to next and the row's first
// byte.  u32 entries (24-bit row + byte) serve the batches of small blocks; one block of more than
// 2^24 rows (libbsc: 25 MiB by default) uses u64 entries (32-bit row, byte above it).
__device__ __forceinline__ u32 entry_next(u32 e) { return e & 0xffffffu; }
__device__ __forceinline__ u32 entry_byte(u32 e) { return e >> 24; }
__device__ __forceinline__ u32 entry_next(u64 e) { return (u32)e; }
__device__ __forceinline__ u32 entry_byte(u64 e) { return (u32)(e >> 32); }
__device__ __forceinline__ void entry_make(u32 &e, u32 row, u32 key) { e = row | (key << 24); }
__device__ __forceinline__ void entry_make(u64 &e, u32 row, u32 key) { e = (u64)row | ((u64)(key & 0xffu) << 32); }

template <typename E>
__global__ void __launch_bounds__(256) ibwt_pack_kernel(const u32 *__restrict__ keys, const u32 *__restrict__ vals,
                                                        u64 N, E *__restrict__ packed)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    entry_make(packed[i], vals[i], keys[i]);
}

// ---- one block WITHOUT an end marker in its alphabet (libbsc, bsc_bwt_decode): the order of equal
// bytes in the first and in the last column only agrees when the text ends with a unique smallest
// symbol.  The cudppCompress blocks carry one (their final 0 byte); for libbsc's blocks the marker
// '$' is virtual: row 0 of an (n + 1)-row problem.  `u` = what bsc_bwt_encode wrote (U[0] = T[n-1],
// then the last column without the row of suffix 0), `primary` = its return value = the row of
// suffix 0 among the n + 1 rows, where the last column holds '$'.
__global__ void __launch_bounds__(256) ibwt_sentinel_keys_kernel(const u8 *__restrict__ u, u32 n, u32 primary,
                                                                 u32 *__restrict__ keys, u32 *__restrict__ vals)
{
    const u32 j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    keys[j] = u[j];
    vals[j] = j < primary ? j : j + 1;        // row of u[j] in the (n + 1)-row last column
}

// entries of rows 1..n from the sorted bytes; row 0 = '$', whose last-column copy sits in row `primary`
template <typename E>
__global__ void __launch_bounds__(256) ibwt_sentinel_pack_kernel(const u32 *__restrict__ keys,
                                                                 const u32 *__restrict__ vals, u32 n, u32 primary,
                                                                 E *__restrict__ packed)
{
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) entry_make(packed[i + 1], vals[i], keys[i]);
    if (i == 0) entry_make(packed[0], primary, 0u);
}

struct Splitters {
    u32 gap_log2;   // regular splitters at rows k << gap_log2
    u32 ns;         // number of regular splitters; id ns = the start row
};

__device__ __forceinline__ bool is_splitter(u32 j, u32 start, u32 gap_mask) { return (j & gap_mask) == 0 || j == start; }

template <typename E>
__global__ void __launch_bounds__(128) ibwt_walk1_kernel(const E *__restrict__ packed, u32 n, Splitters sp,
                                                         const int *__restrict__ bwt_index, u32 nblocks,
                                                         u32 *__restrict__ succ, u32 *__restrict__ plen,
                                                         u32 *__restrict__ error)
{
    const u64 gid = (u64)blockIdx.x * 128 + threadIdx.x;
    const u32 per = sp.ns + 1;
    if (gid >= (u64)nblocks * per) return;
    const u32 blk = (u32)(gid / per), k = (u32)(gid % per);
    u32 start = (u32)bwt_index[blk];
    if (start >= n) { if (k == 0) atomicExch(error, 6u); start = 0; }
    const u32 gap_mask = (1u << sp.gap_log2) - 1;
    const E *T = packed + (u64)blk * n;
    u32 j = k < sp.ns ? k << sp.gap_log2 : start;
    u32 steps = 0;
    do {
        j = entry_next(T[j]);
        ++steps;
    } while (!is_splitter(j, start, gap_mask) && steps < n);
    succ[gid] = j == start ? sp.ns : j >> sp.gap_log2;
    plen[gid] = steps;
}

constexpr u32 kMaxSplitters = 4096;          // per block of a batch
constexpr u32 kMaxSplittersOne = 24576;      // one long block (libbsc): more walkers, 192 KiB of shared memory

// CTA per block: positions of the splitters along the walk that starts at the start row.
// Dynamic shared memory: 2 x (ns + 1) words.
__global__ void __launch_bounds__(256) ibwt_rank_kernel(const u32 *__restrict__ succ, const u32 *__restrict__ plen,
                                                        u32 n, Splitters sp, u32 *__restrict__ task_pos,
                                                        u32 *__restrict__ period)
{
    extern __shared__ u32 s_rank[];
    const u32 per = sp.ns + 1;
    u32 *s_succ = s_rank, *s_len = s_rank + per;
    const u64 base = (u64)blockIdx.x * per;
    for (u32 i = threadIdx.x; i < per; i += 256) {
        s_succ[i] = succ[base + i];
        s_len[i] = plen[base + i];
        task_pos[base + i] = 0xffffffffu;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 k = sp.ns, pos = 0;
        do {
            task_pos[base + k] = pos;
            pos += s_len[k];
            k = s_succ[k];
        } while (pos < n && k != sp.ns);
        period[blockIdx.x] = pos < n ? pos : 0u;
    }
}

template <typename E>
__global__ void __launch_bounds__(128) ibwt_walk2_kernel(const E *__restrict__ packed, u32 n, Splitters sp,
                                                         const int *__restrict__ bwt_index, u32 nblocks,
                                                         const u32 *__restrict__ task_pos,
                                                         const u32 *__restrict__ plen, u8 *__restrict__ out)
{
    const u64 gid = (u64)blockIdx.x * 128 + threadIdx.x;
    const u32 per = sp.ns + 1;
    if (gid >= (u64)nblocks * per) return;
    const u32 blk = (u32)(gid / per), k = (u32)(gid % per);
    const u32 pos = task_pos[gid];
    if (pos == 0xffffffffu) return;
    u32 start = (u32)bwt_index[blk];
    if (start >= n) start = 0;
    const E *T = packed + (u64)blk * n;
    u8 *dst = out + (u64)blk * n;
    u32 j = k < sp.ns ? k << sp.gap_log2 : start;
    const u32 cnt = min(plen[gid], n - pos);
    for (u32 t = 0; t < cnt; ++t) {
        const E e = T[j];
        dst[pos + t] = (u8)entry_byte(e);
        j = entry_next(e);
    }
}

__global__ void __launch_bounds__(256) ibwt_extend_kernel(u8 *__restrict__ out, u64 N, u32 n,
                                                          const u32 *__restrict__ period)
{
    const u64 i = (u64)blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const u32 blk = (u32)(i / n);
    const u32 m = period[blk];
    if (m == 0) return;
    const u32 local = (u32)(i - (u64)blk * n);
    if (local >= m) out[i] = out[(u64)blk * n + local % m];
}

static Splitters splitters_for(u32 n, u32 max_splitters = kMaxSplitters)
{
    Splitters sp;
    sp.gap_log2 = 8;
    while (((u64)n + (1ull << sp.gap_log2) - 1) >> sp.gap_log2 > max_splitters) ++sp.gap_log2;
    sp.ns = (u32)(((u64)n + (1ull << sp.gap_log2) - 1) >> sp.gap_log2);
    return sp;
}

struct IbwtLayout {
    size_t keys_a, keys_b, vals_a, vals_b, succ, plen, task, period, sort_temp, sort_bytes, packed64, total;
};

// rows fit the 24-bit field of a u32 entry
static bool narrow_rows(u32 n) { return n < (1u << 24); }

static IbwtLayout ibwt_layout(u64 nblocks, u32 n)
{
    IbwtLayout L;
    const u64 N = nblocks * n;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    L.sort_bytes = prims::sort_scratch_bytes(N, n);
    // a single block may be walked with the larger splitter budget (inverse_bwt_sentinel)
    const Splitters sp = splitters_for(n, nblocks == 1 ? kMaxSplittersOne : kMaxSplitters);
    const size_t per = (size_t)nblocks * (sp.ns + 1) * 4;
    size_t o = 0;
    L.keys_a = o; o += up(N * 4);
    L.keys_b = o; o += up(N * 4);
    L.vals_a = o; o += up(N * 4);
    L.vals_b = o; o += up(N * 4);
    L.succ = o; o += up(per);
    L.plen = o; o += up(per);
    L.task = o; o += up(per);
    L.period = o; o += up(nblocks * 4);
    L.sort_temp = o; o += up(L.sort_bytes);
    L.packed64 = o;
    if (!narrow_rows(n)) o += up(N * 8);
    L.total = o;
    return L;
}

// bwt bytes (nblocks x n) -> original blocks.  If keys_ready, keys_a / vals_a were already filled
// by imtf_map_kernel.
static int inverse_bwt(const u8 *d_bwt, const int *d_index, u64 nblocks, u32 n, u8 *d_out, u32 *d_error,
                       char *scratch, const IbwtLayout &L, bool keys_ready, cudaStream_t stream)
{
    const u64 N = nblocks * n;
    u32 *keys_a = reinterpret_cast<u32 *>(scratch + L.keys_a), *keys_b = reinterpret_cast<u32 *>(scratch + L.keys_b);
    u32 *vals_a = reinterpret_cast<u32 *>(scratch + L.vals_a), *vals_b = reinterpret_cast<u32 *>(scratch + L.vals_b);
    u32 *succ = reinterpret_cast<u32 *>(scratch + L.succ), *plen = reinterpret_cast<u32 *>(scratch + L.plen);
    u32 *task = reinterpret_cast<u32 *>(scratch + L.task), *period = reinterpret_cast<u32 *>(scratch + L.period);
    const u32 grid_n = (u32)((N + 255) / 256);
    if (!keys_ready) {
        ibwt_keys_kernel<<<grid_n, 256, 0, stream>>>(d_bwt, N, n, keys_a, vals_a);
        B200LC_CUDA_TRY(cudaGetLastError());
    }
    // every block is a sort segment, so the block number in bits 8.. of the key is not sorted on:
    // one 8-bit pass = a stable counting sort of each block's bytes
    int in_b = 0;
    const int rc = prims::sort_pairs<u32>(keys_a, keys_b, vals_a, vals_b, N, n, 0, 8, scratch + L.sort_temp,
                                          L.sort_bytes, stream, &in_b);
    if (rc) return rc;
    const Splitters sp = splitters_for(n);
    const u64 tasks = nblocks * (sp.ns + 1);
    const u32 grid_t = (u32)((tasks + 127) / 128);
    const u32 *skeys = in_b ? keys_b : keys_a, *svals = in_b ? vals_b : vals_a;
    const auto walks = [&](auto *packed) -> int {
        ibwt_pack_kernel<<<grid_n, 256, 0, stream>>>(skeys, svals, N, packed);
        B200LC_CUDA_TRY(cudaGetLastError());
        ibwt_walk1_kernel<<<grid_t, 128, 0, stream>>>(packed, n, sp, d_index, (u32)nblocks, succ, plen, d_error);
        B200LC_CUDA_TRY(cudaGetLastError());
        ibwt_rank_kernel<<<(u32)nblocks, 256, 2 * (sp.ns + 1) * sizeof(u32), stream>>>(succ, plen, n, sp, task, period);
        B200LC_CUDA_TRY(cudaGetLastError());
        ibwt_walk2_kernel<<<grid_t, 128, 0, stream>>>(packed, n, sp, d_index, (u32)nblocks, task, plen, d_out);
        B200LC_CUDA_TRY(cudaGetLastError());
        return B200LC_OK;
    };
    // u32 entries are written over the key buffer the sort left free; u64 entries have their own
    const int wrc = narrow_rows(n) ? walks(in_b ? keys_a : keys_b)
                                   : walks(reinterpret_cast<u64 *>(scratch + L.packed64));
    if (wrc) return wrc;
    ibwt_extend_kernel<<<grid_n, 256, 0, stream>>>(d_out, N, n, period);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

// u[0..n) + primary (device int at d_primary, host copy `primary`) -> n text bytes at d_out, which
// needs n + 1 bytes (the walk also emits the virtual end marker).  L = ibwt_layout(1, n + 1).
static int inverse_bwt_sentinel(const u8 *d_u, u32 n, u32 primary, const int *d_primary, u8 *d_out, u32 *d_error,
                                char *scratch, const IbwtLayout &L, cudaStream_t stream)
{
    u32 *keys_a = reinterpret_cast<u32 *>(scratch + L.keys_a), *keys_b = reinterpret_cast<u32 *>(scratch + L.keys_b);
    u32 *vals_a = reinterpret_cast<u32 *>(scratch + L.vals_a), *vals_b = reinterpret_cast<u32 *>(scratch + L.vals_b);
    u32 *succ = reinterpret_cast<u32 *>(scratch + L.succ), *plen = reinterpret_cast<u32 *>(scratch + L.plen);
    u32 *task = reinterpret_cast<u32 *>(scratch + L.task), *period = reinterpret_cast<u32 *>(scratch + L.period);
    const u32 rows = n + 1;
    const u32 grid_n = (n + 255) / 256;
    ibwt_sentinel_keys_kernel<<<grid_n, 256, 0, stream>>>(d_u, n, primary, keys_a, vals_a);
    B200LC_CUDA_TRY(cudaGetLastError());
    int in_b = 0;
    const int rc = prims::sort_pairs<u32>(keys_a, keys_b, vals_a, vals_b, n, n, 0, 8, scratch + L.sort_temp,
                                          L.sort_bytes, stream, &in_b);
    if (rc) return rc;
    const Splitters sp = splitters_for(rows, kMaxSplittersOne);
    const u32 grid_t = (sp.ns + 1 + 127) / 128;
    const size_t rank_smem = 2 * (size_t)(sp.ns + 1) * sizeof(u32);
    if (rank_smem > 48 * 1024) {
        static unsigned attr_done[kMaxDevices] = {0};
        const int slot = device_slot();
        if (slot < 0 || attr_done[slot] != context_epoch()) {
            B200LC_CUDA_TRY(cudaFuncSetAttribute(ibwt_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 (int)(2 * (kMaxSplittersOne + 1) * sizeof(u32))));
            if (slot >= 0) attr_done[slot] = context_epoch();
        }
    }
    const u32 *skeys = in_b ? keys_b : keys_a, *svals = in_b ? vals_b : vals_a;
    const auto walks = [&](auto *packed) -> int {
        ibwt_sentinel_pack_kernel<<<grid_n, 256, 0, stream>>>(skeys, svals, n, primary, packed);
        B200LC_CUDA_TRY(cudaGetLastError());
        ibwt_walk1_kernel<<<grid_t, 128, 0, stream>>>(packed, rows, sp, d_primary, 1u, succ, plen, d_error);
        B200LC_CUDA_TRY(cudaGetLastError());
        ibwt_rank_kernel<<<1, 256, rank_smem, stream>>>(succ, plen, rows, sp, task, period);
        B200LC_CUDA_TRY(cudaGetLastError());
        ibwt_walk2_kernel<<<grid_t, 128, 0, stream>>>(packed, rows, sp, d_primary, 1u, task, plen, d_out);
        B200LC_CUDA_TRY(cudaGetLastError());
        return B200LC_OK;
    };
    // the u32 entries need rows * 4 bytes: the free key buffer was sized for n + 1 elements
    return narrow_rows(rows) ? walks(in_b ? keys_a : keys_b) : walks(reinterpret_cast<u64 *>(scratch + L.packed64));
}

// blocks of up to 2^24 - 1 rows in any number; longer blocks (u64 entries) one at a time
static bool shape_ok(size_t nblocks, size_t n)
{
    if ((u64)nblocks * n > prims::kSortMaxElems || nblocks > (1ull << 23)) return false;
    return n < (1ull << 24) || nblocks == 1;
}

}  // namespace cdec
}  // namespace b200lc

using namespace b200lc;

static size_t up256(size_t x) { return (x + 255) & ~size_t(255); }

extern "C" size_t b200lc_inverse_mtf_scratch_bytes(size_t nblocks, size_t n)
{
    const size_t nseg = (n + cdec::kSeg - 1) / cdec::kSeg;
    return up256(nblocks * nseg * 256);
}

// ranks (nblocks x n) -> symbols; d_in and d_out may be the same buffer.
extern "C" int b200lc_inverse_mtf_batch(const uint8_t *d_in, size_t nblocks, size_t n, uint8_t *d_out,
                                        void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nblocks == 0 || n == 0) return B200LC_OK;
    if (!d_in || !d_out || !d_scratch) return B200LC_ERR_ARG;
    if (n >= (1ull << 31) || (reinterpret_cast<uintptr_t>(d_scratch) & 15)) return B200LC_ERR_ARG;
    if (scratch_bytes < b200lc_inverse_mtf_scratch_bytes(nblocks, n)) return B200LC_ERR_SCRATCH;
    const u32 nseg = (u32)((n + cdec::kSeg - 1) / cdec::kSeg);
    const u64 total = (u64)nblocks * nseg;
    if (total >= (1ull << 31)) return B200LC_ERR_UNSUPPORTED;
    u8 *perms = reinterpret_cast<u8 *>(d_scratch);
    cdec::imtf_pid_kernel<<<(u32)((total + cdec::kPidThreads - 1) / cdec::kPidThreads), cdec::kPidThreads, 0,
                            stream>>>(d_in, (u32)n, nseg, total, d_out, perms);
    B200LC_CUDA_TRY(cudaGetLastError());
    cdec::imtf_chain_kernel<<<(u32)nblocks, 256, 0, stream>>>(perms, nseg);
    B200LC_CUDA_TRY(cudaGetLastError());
    cdec::imtf_map_kernel<<<(u32)total, 128, 0, stream>>>(d_out, (u32)n, nseg, perms, nullptr, nullptr);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

extern "C" size_t b200lc_inverse_bwt_scratch_bytes(size_t nblocks, size_t n)
{
    if (nblocks == 0 || n == 0 || !cdec::shape_ok(nblocks, n)) return 256;
    return cdec::ibwt_layout(nblocks, (u32)n).total;
}

// d_bwt (nblocks x n) + d_bwt_index[nblocks] -> original blocks.  d_out must not alias d_bwt.
extern "C" int b200lc_inverse_bwt_batch(const uint8_t *d_bwt, const int *d_bwt_index, size_t nblocks, size_t n,
                                        uint8_t *d_out, uint32_t *d_error, void *d_scratch,
                                        size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nblocks == 0 || n == 0) return B200LC_OK;
    if (!d_bwt || !d_bwt_index || !d_out || !d_error || !d_scratch) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_scratch) & 255) return B200LC_ERR_ARG;
    if (!cdec::shape_ok(nblocks, n)) return B200LC_ERR_UNSUPPORTED;
    const cdec::IbwtLayout L = cdec::ibwt_layout(nblocks, (u32)n);
    if (scratch_bytes < L.total) return B200LC_ERR_SCRATCH;
    B200LC_CUDA_TRY(cudaMemsetAsync(d_error, 0, sizeof(u32), stream));
    return cdec::inverse_bwt(d_bwt, d_bwt_index, nblocks, (u32)n, d_out, d_error,
                             reinterpret_cast<char *>(d_scratch), L, false, stream);
}

// One block in libbsc's convention (no end marker in the alphabet): d_u = what bsc_bwt_encode wrote,
// primary = its return value (1..n).  d_out needs n + 1 bytes; the first n are the block.
extern "C" size_t b200lc_inverse_bwt_primary_scratch_bytes(size_t n)
{
    if (n == 0 || n + 1 > prims::kSortMaxElems) return 256;
    return cdec::ibwt_layout(1, (u32)n + 1).total + 256;
}

extern "C" int b200lc_inverse_bwt_primary(const uint8_t *d_u, size_t n, int primary, uint8_t *d_out,
                                          uint32_t *d_error, void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return B200LC_OK;
    if (!d_u || !d_out || !d_error || !d_scratch) return B200LC_ERR_ARG;
    if (primary <= 0 || (size_t)primary > n) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_scratch) & 255) return B200LC_ERR_ARG;
    if (n + 1 > prims::kSortMaxElems) return B200LC_ERR_UNSUPPORTED;
    const cdec::IbwtLayout L = cdec::ibwt_layout(1, (u32)n + 1);
    if (scratch_bytes < L.total + 256) return B200LC_ERR_SCRATCH;
    int *d_primary = reinterpret_cast<int *>(reinterpret_cast<char *>(d_scratch) + L.total);
    B200LC_CUDA_TRY(cudaMemsetAsync(d_error, 0, sizeof(u32), stream));
    B200LC_CUDA_TRY(cudaMemcpyAsync(d_primary, &primary, sizeof(int), cudaMemcpyHostToDevice, stream));
    const int rc = cdec::inverse_bwt_sentinel(d_u, (u32)n, (u32)primary, d_primary, d_out, d_error,
                                              reinterpret_cast<char *>(d_scratch), L, stream);
    // `primary` is a stack variable: the copy above must have happened before we return
    B200LC_CUDA_TRY(cudaStreamSynchronize(stream));
    return rc;
}

extern "C" size_t b200lc_cudpp_decompress_scratch_bytes(size_t nblocks, size_t n)
{
    if (nblocks == 0 || n == 0 || !cdec::shape_ok(nblocks, n)) return 256;
    const size_t nseg = (n + cdec::kSeg - 1) / cdec::kSeg;
    return up256(nblocks * n)                                   // ranks -> ids -> BWT bytes
           + up256(nblocks * nseg * 256)                        // permutations / lists
           + up256(nblocks * chuff::kTreeShorts * 2)            // trees
           + up256(nblocks * chuff::kSyms * 4) + up256(nblocks * chuff::kSyms)   // codes, lens (tree kernel)
           + cdec::ibwt_layout(nblocks, (u32)n).total;
}

// Inverse of b200lc_cudpp_compress_batch / cudppCompress: per block b the stream at
// d_comp + b * comp_stride_words with word offsets d_offsets[b * nhb ..], histogram d_hist[b*256..]
// and d_bwt_index[b]  ->  n bytes at d_out + b * n.  *d_error != 0 afterwards = corrupt stream.
extern "C" int b200lc_cudpp_decompress_batch(const int *d_bwt_index, const uint32_t *d_hist,
                                             const uint32_t *d_offsets, const uint32_t *d_comp,
                                             size_t comp_stride_words, size_t nblocks, size_t n,
                                             uint8_t *d_out, uint32_t *d_error, void *d_scratch,
                                             size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    if (nblocks == 0 || n == 0) return B200LC_OK;
    if (!d_bwt_index || !d_hist || !d_offsets || !d_comp || !d_out || !d_error || !d_scratch) return B200LC_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(d_scratch) & 255) return B200LC_ERR_ARG;
    if (!cdec::shape_ok(nblocks, n)) return B200LC_ERR_UNSUPPORTED;
    if (scratch_bytes < b200lc_cudpp_decompress_scratch_bytes(nblocks, n)) return B200LC_ERR_SCRATCH;
    const u32 nseg = (u32)((n + cdec::kSeg - 1) / cdec::kSeg);
    const u32 nhb = (u32)((n + chuff::kBlockChars - 1) / chuff::kBlockChars);
    const u64 total_segs = (u64)nblocks * nseg;
    if (total_segs >= (1ull << 31) || (u64)nblocks * nhb >= (1ull << 31)) return B200LC_ERR_UNSUPPORTED;
    char *s = reinterpret_cast<char *>(d_scratch);
    u8 *data = reinterpret_cast<u8 *>(s);              s += up256(nblocks * n);
    u8 *perms = reinterpret_cast<u8 *>(s);             s += up256(nblocks * nseg * 256);
    short *trees = reinterpret_cast<short *>(s);       s += up256(nblocks * chuff::kTreeShorts * 2);
    u32 *codes = reinterpret_cast<u32 *>(s);           s += up256(nblocks * chuff::kSyms * 4);
    u8 *lens = reinterpret_cast<u8 *>(s);              s += up256(nblocks * chuff::kSyms);
    const cdec::IbwtLayout L = cdec::ibwt_layout(nblocks, (u32)n);

    B200LC_CUDA_TRY(cudaMemsetAsync(d_error, 0, sizeof(u32), stream));
    B200LC_CUDA_TRY(chuff::launch_tree(d_hist, (u32)nblocks, codes, lens, trees, d_error, stream));
    const u32 groups = (nhb + cdec::kHdecThreads - 1) / cdec::kHdecThreads;
    cdec::hdec_kernel<<<(u32)(nblocks * groups), cdec::kHdecThreads, 0, stream>>>(
        d_comp, comp_stride_words, d_offsets, (u32)n, nhb, groups, trees, data, d_error);
    B200LC_CUDA_TRY(cudaGetLastError());
    cdec::imtf_pid_kernel<<<(u32)((total_segs + cdec::kPidThreads - 1) / cdec::kPidThreads), cdec::kPidThreads, 0,
                            stream>>>(data, (u32)n, nseg, total_segs, data, perms);
    B200LC_CUDA_TRY(cudaGetLastError());
    cdec::imtf_chain_kernel<<<(u32)nblocks, 256, 0, stream>>>(perms, nseg);
    B200LC_CUDA_TRY(cudaGetLastError());
    cdec::imtf_map_kernel<<<(u32)total_segs, 128, 0, stream>>>(data, (u32)n, nseg, perms,
                                                              reinterpret_cast<u32 *>(s + L.keys_a),
                                                              reinterpret_cast<u32 *>(s + L.vals_a));
    B200LC_CUDA_TRY(cudaGetLastError());
    return cdec::inverse_bwt(data, d_bwt_index, nblocks, (u32)n, d_out, d_error, s, L, true, stream);
}
