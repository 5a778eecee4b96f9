// Burrows-Wheeler transform of many independent blocks  (hot path 1).
//
// Definition (bit-exact with cudppBurrowsWheelerTransform):
// SA = suffix array of the block with an implicit terminator smaller than every byte
// (sa_kernel.cuh:47-60: T[i] = in[i] + 1 followed by zeros); bwt[i] = in[SA[i] - 1], or
// in[n-1] where SA[i] == 0, and index = that row (compress_kernel.cuh:55-74).
//
// The reference builds ONE suffix array at a time with a recursive skew/DC3 (13 small kernels,
// 4 radix sorts and a merge per level, cudaMalloc/cudaFree around every sort,
// sa_app.cu:61-101,125-298).  Here all blocks of a batch are sorted together by prefix
// doubling: round r sorts the 64-bit keys (block, rank_h[i], rank_h[i+h]) of every suffix of
// every block with one device-wide radix sort, renames the groups and doubles h, until every
// group is a singleton.  Random-like data needs 1-3 rounds, text 4-6.
// The sorts and scans are the hand-written segmented one-sweep radix sort and look-back scans
// of devprims.cu: in the first round every block is a segment of its own, so the block number
// costs no key bits (6 passes over 48 bits of characters).
#include "common.cuh"
#include "devprims.cuh"
#include "../../include/b200lc.h"

namespace b200lc {
namespace bwt {

constexpr int kThreads = 256;

// Every element-wise kernel below gives CTA b the contiguous elements [b * kChunk, (b + 1) * kChunk):
// CTAs are scheduled in order, so at any moment the whole GPU works inside a window of a few
// blocks and the random accesses of the suffix-array kernels (rank[sa[j]], rank[g + h],
// in[pos - 1]) stay in L2.  (A grid-stride loop lets the CTAs drift apart: measured 32 B read +
// 32 B written in DRAM per scattered 4-byte rank.)
constexpr u32 kChunk = 16 * kThreads;
#define B200LC_FOR_CHUNK(var, count)                                                            \
    for (u32 var = blockIdx.x * kChunk + threadIdx.x, var##_end = min((u32)(count), (blockIdx.x + 1) * kChunk); \
         var < var##_end; var += kThreads)
static inline u32 chunk_grid(u64 count) { return (u32)((count + kChunk - 1) / kChunk); }

// First round: key = the first kFirstChars raw bytes of the suffix, zero-padded past the end of
// the block; every block is one sort segment.  The (at most kFirstChars - 1) suffixes that are
// shorter than the key are ordered by the sort's stability instead of by key bits: the initial
// order inside a block is by DESCENDING start position, so among equal keys the shorter suffix --
// which is the smaller one, its implicit terminator being smaller than byte 0 -- comes first.
// Such a suffix is always a group of its own (mark_heads_kernel).
constexpr u32 kFirstChars = 6;
constexpr u32 kFirstKeyBits = 8 * kFirstChars;
__global__ void __launch_bounds__(kThreads) init_keys_kernel(const u8 *__restrict__ in, u64 N, u32 n,
                                                            u64 *__restrict__ keys,
                                                            u32 *__restrict__ start)
{
    B200LC_FOR_CHUNK(g, N) {
        const u32 blk = g / n;
        const u32 base = blk * n;
        const u32 i = n - 1 - (g - base);
        u64 k = 0;
#pragma unroll
        for (u32 c = 0; c < kFirstChars; ++c) k = (k << 8) | (i + c < n ? (u64)in[base + i + c] : 0);
        keys[g] = k;
        start[g] = (u32)(base + i);
    }
}

// heads[j] = j at the first element of every group of equal first-round keys, else 0; counts
// the non-head elements (0 means every suffix is alone in its group).  A block start and a
// suffix shorter than the key always open a group, and so does the element after such a suffix.
__global__ void __launch_bounds__(kThreads) mark_heads_kernel(const u64 *__restrict__ keys,
                                                             const u32 *__restrict__ sa, u64 N, u32 n,
                                                             u32 *__restrict__ heads,
                                                             unsigned long long *__restrict__ dup)
{
    u32 local = 0;
    B200LC_FOR_CHUNK(j, N) {
        bool head = j % n == 0 || keys[j] != keys[j - 1];
        if (!head) {
            const u32 a = sa[j] % n, b = sa[j - 1] % n;
            head = a + kFirstChars > n || b + kFirstChars > n;
        }
        heads[j] = head ? (u32)j : 0u;
        local += !head;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) local += __shfl_xor_sync(0xffffffffu, local, d);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(dup, (unsigned long long)local);
}

// After the first (full) sort: rank of a suffix = 1 + position of its group's head inside the
// block; a position is "unresolved" when its group has more than one member.
__global__ void __launch_bounds__(kThreads) first_rank_kernel(const u32 *__restrict__ head_of,
                                                             const u32 *__restrict__ sa, u64 N, u32 n,
                                                             u32 *__restrict__ rank,
                                                             u32 *__restrict__ uflag)
{
    B200LC_FOR_CHUNK(j, N) {
        const u32 blk = j / n;
        const u32 hd = head_of[j];
        rank[sa[j]] = hd - blk * n + 1;
        const bool single = hd == (u32)j && (j + 1 == N || head_of[j + 1] == (u32)(j + 1));
        uflag[j] = single ? 0u : 1u;
    }
}

// Refinement round: only the members of unresolved groups are re-sorted.  Key = (position of the
// group's head in the suffix array, rank of the suffix h characters further on).
__global__ void __launch_bounds__(kThreads) compact_keys_kernel(const u32 *__restrict__ uflag,
                                                               const u32 *__restrict__ cidx,
                                                               const u32 *__restrict__ head_of,
                                                               const u32 *__restrict__ sa,
                                                               const u32 *__restrict__ rank, u64 N, u32 n,
                                                               u32 h, u32 rank_bits, u64 *__restrict__ ckey,
                                                               u32 *__restrict__ cval)
{
    B200LC_FOR_CHUNK(j, N) {
        if (!uflag[j]) continue;
        const u32 g = sa[j];
        const u32 blk = g / n, i = g - blk * n;
        const u64 r2 = (i + h < n) ? rank[g + h] : 0;
        const u32 c = cidx[j];
        ckey[c] = ((u64)head_of[j] << rank_bits) | r2;
        cval[c] = g;
    }
}

// Sorted compact element c goes to suffix-array position p = head + (c - compact index of head).
__global__ void __launch_bounds__(kThreads) place_kernel(const u64 *__restrict__ skey,
                                                        const u32 *__restrict__ sval,
                                                        const u32 *__restrict__ cidx, u32 M, u32 rank_bits,
                                                        u32 *__restrict__ sa, u32 *__restrict__ newhead)
{
    B200LC_FOR_CHUNK(c, M) {
        const u64 k = skey[c];
        const u32 hd = (u32)(k >> rank_bits);
        const u32 first = cidx[hd];
        const u32 p = hd + (c - first);
        sa[p] = sval[c];
        newhead[c] = (c == first || k != skey[c - 1]) ? p : 0u;
    }
}

__global__ void __launch_bounds__(kThreads) rerank_kernel(const u64 *__restrict__ skey,
                                                         const u32 *__restrict__ sval,
                                                         const u32 *__restrict__ cidx,
                                                         const u32 *__restrict__ nh, u32 M, u32 n,
                                                         u32 rank_bits, u32 *__restrict__ head_of,
                                                         u32 *__restrict__ rank, u32 *__restrict__ uflag)
{
    B200LC_FOR_CHUNK(c, M) {
        const u32 hd = (u32)(skey[c] >> rank_bits);
        const u32 p = hd + (c - cidx[hd]);
        const u32 my_head = nh[c];
        head_of[p] = my_head;
        rank[sval[c]] = my_head - (p / n) * n + 1;
        bool next_is_head = true;
        if (c + 1 < M) {
            const u32 hd2 = (u32)(skey[c + 1] >> rank_bits);
            const u32 p2 = hd2 + (c + 1 - cidx[hd2]);
            next_is_head = nh[c + 1] == p2;
        }
        uflag[p] = (my_head == p && next_is_head) ? 0u : 1u;
    }
}

__global__ void __launch_bounds__(kThreads) bwt_gather_kernel(const u8 *__restrict__ in,
                                                             const u32 *__restrict__ sa, u64 N, u32 n,
                                                             u8 *__restrict__ out,
                                                             int *__restrict__ index,
                                                             u32 *__restrict__ sa_out)
{
    B200LC_FOR_CHUNK(j, N) {
        const u32 blk = j / n;
        const u32 base = blk * n;
        const u32 pos = sa[j] - base;             // suffix start inside the block
        if (sa_out) sa_out[j] = pos;
        if (out) {
            if (pos == 0) {
                out[j] = in[base + n - 1];
                index[blk] = (int)(j - base);
            } else {
                out[j] = in[base + pos - 1];
            }
        }
    }
}

struct Layout {
    size_t keys_a, keys_b, vals_a, vals_b, vals_c, rank, heads, uflag, cidx, counter, prim_temp, total;
    size_t prim_bytes;
};

static Layout layout(u64 N, u32 n)
{
    Layout L;
    L.prim_bytes = prims::sort_scratch_bytes(N, n);
    const size_t whole = prims::sort_scratch_bytes(N, N), scan = prims::scan_scratch_bytes(N);
    if (whole > L.prim_bytes) L.prim_bytes = whole;
    if (scan > L.prim_bytes) L.prim_bytes = scan;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t o = 0;
    L.keys_a = o; o += up(N * 8);
    L.keys_b = o; o += up(N * 8);
    L.vals_a = o; o += up(N * 4);
    L.vals_b = o; o += up(N * 4);
    L.vals_c = o; o += up(N * 4);
    L.rank = o; o += up(N * 4);
    L.heads = o; o += up(N * 4);
    L.uflag = o; o += up(N * 4);
    L.cidx = o; o += up(N * 4);
    L.counter = o; o += 256;
    L.prim_temp = o; o += up(L.prim_bytes);
    L.total = o;
    return L;
}

// Sorted suffix order of every block (global indices) somewhere in scratch; returns 0 or an error.
static int suffix_sort(const u8 *d_in, u64 nblocks, u32 n, char *scratch, const Layout &L,
                       cudaStream_t stream, const u32 **sa_global)
{
    const u64 N = nblocks * n;
    u64 *keys_a = reinterpret_cast<u64 *>(scratch + L.keys_a);
    u64 *keys_b = reinterpret_cast<u64 *>(scratch + L.keys_b);
    u32 *vals_a = reinterpret_cast<u32 *>(scratch + L.vals_a);
    u32 *vals_b = reinterpret_cast<u32 *>(scratch + L.vals_b);
    u32 *rank = reinterpret_cast<u32 *>(scratch + L.rank);
    u32 *heads = reinterpret_cast<u32 *>(scratch + L.heads);
    u32 *uflag = reinterpret_cast<u32 *>(scratch + L.uflag);
    u32 *cidx = reinterpret_cast<u32 *>(scratch + L.cidx);
    unsigned long long *dup = reinterpret_cast<unsigned long long *>(scratch + L.counter);
    void *ptemp = scratch + L.prim_temp;
    const size_t pbytes = L.prim_bytes;
    const u32 grid = chunk_grid(N);
    int pos_bits = 1;
    while ((1ull << pos_bits) < N) ++pos_bits;
    u32 rank_bits = 1;                       // ranks run from 0 (past the end) to n
    while ((1ull << rank_bits) <= n) ++rank_bits;

    // ---- round 1: every suffix, by its first kFirstChars characters, block by block
    init_keys_kernel<<<grid, kThreads, 0, stream>>>(d_in, N, n, keys_a, vals_a);
    B200LC_CUDA_TRY(cudaGetLastError());
    int in_b = 0;
    int rc = prims::sort_pairs<u64>(keys_a, keys_b, vals_a, vals_b, N, n, 0, (int)kFirstKeyBits, ptemp,
                                    pbytes, stream, &in_b);
    if (rc) return rc;
    const u64 *skeys = in_b ? keys_b : keys_a;
    u32 *sa = in_b ? vals_b : vals_a;             // the suffix array lives here from now on
    u32 *fvals = in_b ? vals_a : vals_b;          // free 4-byte buffer
    *sa_global = sa;
    B200LC_CUDA_TRY(cudaMemsetAsync(dup, 0, sizeof(*dup), stream));
    mark_heads_kernel<<<grid, kThreads, 0, stream>>>(skeys, sa, N, n, heads, dup);
    B200LC_CUDA_TRY(cudaGetLastError());
    unsigned long long h_dup = 0;
    B200LC_CUDA_TRY(cudaMemcpyAsync(&h_dup, dup, sizeof(h_dup), cudaMemcpyDeviceToHost, stream));
    B200LC_CUDA_TRY(cudaStreamSynchronize(stream));
    if (h_dup == 0) return B200LC_OK;
    // every element learns the position of its group's head (running maximum), then its rank
    rc = prims::inclusive_max_u32(heads, heads, N, ptemp, pbytes, stream);
    if (rc) return rc;
    first_rank_kernel<<<grid, kThreads, 0, stream>>>(heads, sa, N, n, rank, uflag);
    B200LC_CUDA_TRY(cudaGetLastError());

    // ---- refinement rounds: only members of unresolved groups.  Both key buffers and the value
    // buffer that does not hold the suffix array are free now; vals_c is the second value buffer.
    u32 *vals_c = reinterpret_cast<u32 *>(scratch + L.vals_c);
    for (u32 h = kFirstChars; h < n; h <<= 1) {
        rc = prims::exclusive_sum_u32(uflag, cidx, N, ptemp, pbytes, stream);
        if (rc) return rc;
        u32 last[2] = {0, 0};
        B200LC_CUDA_TRY(cudaMemcpyAsync(&last[0], cidx + (N - 1), 4, cudaMemcpyDeviceToHost, stream));
        B200LC_CUDA_TRY(cudaMemcpyAsync(&last[1], uflag + (N - 1), 4, cudaMemcpyDeviceToHost, stream));
        B200LC_CUDA_TRY(cudaStreamSynchronize(stream));
        const u32 M = last[0] + last[1];
        if (M == 0) break;
        compact_keys_kernel<<<grid, kThreads, 0, stream>>>(uflag, cidx, heads, sa, rank, N, n, h, rank_bits, keys_a, fvals);
        B200LC_CUDA_TRY(cudaGetLastError());
        rc = prims::sort_pairs<u64>(keys_a, keys_b, fvals, vals_c, M, M, 0, (int)rank_bits + pos_bits, ptemp,
                                    pbytes, stream, &in_b);
        if (rc) return rc;
        const u64 *skey = in_b ? keys_b : keys_a;
        const u32 *sval = in_b ? vals_c : fvals;
        u32 *newhead = reinterpret_cast<u32 *>(in_b ? keys_a : keys_b);   // the key buffer the sort left free
        const u32 mgrid = chunk_grid(M);
        place_kernel<<<mgrid, kThreads, 0, stream>>>(skey, sval, cidx, M, rank_bits, sa, newhead);
        B200LC_CUDA_TRY(cudaGetLastError());
        rc = prims::inclusive_max_u32(newhead, newhead, M, ptemp, pbytes, stream);
        if (rc) return rc;
        rerank_kernel<<<mgrid, kThreads, 0, stream>>>(skey, sval, cidx, newhead, M, n, rank_bits, heads, rank, uflag);
        B200LC_CUDA_TRY(cudaGetLastError());
    }
    return B200LC_OK;
}

}  // namespace bwt
}  // namespace b200lc

using namespace b200lc;

extern "C" size_t b200lc_bwt_scratch_bytes(size_t nblocks, size_t n)
{
    return bwt::layout((u64)nblocks * n, (u32)n).total;
}

static int bwt_check(const void *d_in, size_t nblocks, size_t n, void *d_scratch, size_t scratch_bytes)
{
    if (!d_in || !d_scratch) return B200LC_ERR_ARG;
    if (n == 0 || nblocks == 0) return B200LC_ERR_UNSUPPORTED;
    if ((u64)nblocks * n > prims::kSortMaxElems) return B200LC_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(d_scratch) & 255) return B200LC_ERR_ARG;
    if (scratch_bytes < b200lc_bwt_scratch_bytes(nblocks, n)) return B200LC_ERR_SCRATCH;
    return B200LC_OK;
}

// Synchronises the stream (the doubling loop reads one counter per round).
extern "C" int b200lc_bwt_batch(const uint8_t *d_in, size_t nblocks, size_t n, uint8_t *d_out,
                                int *d_index, void *d_scratch, size_t scratch_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = bwt_check(d_in, nblocks, n, d_scratch, scratch_bytes);
    if (rc) return rc;
    if (!d_out || !d_index) return B200LC_ERR_ARG;
    const bwt::Layout L = bwt::layout((u64)nblocks * n, (u32)n);
    const u32 *sa = nullptr;
    rc = bwt::suffix_sort(d_in, nblocks, (u32)n, reinterpret_cast<char *>(d_scratch), L, stream, &sa);
    if (rc) return rc;
    const u64 N = (u64)nblocks * n;
    const u32 grid = bwt::chunk_grid(N);
    bwt::bwt_gather_kernel<<<grid, bwt::kThreads, 0, stream>>>(d_in, sa, N, (u32)n, d_out, d_index, nullptr);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

extern "C" int b200lc_suffix_array_batch(const uint8_t *d_in, size_t nblocks, size_t n,
                                         uint32_t *d_sa, void *d_scratch, size_t scratch_bytes,
                                         void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = bwt_check(d_in, nblocks, n, d_scratch, scratch_bytes);
    if (rc) return rc;
    if (!d_sa) return B200LC_ERR_ARG;
    const bwt::Layout L = bwt::layout((u64)nblocks * n, (u32)n);
    const u32 *sa = nullptr;
    rc = bwt::suffix_sort(d_in, nblocks, (u32)n, reinterpret_cast<char *>(d_scratch), L, stream, &sa);
    if (rc) return rc;
    const u64 N = (u64)nblocks * n;
    const u32 grid = bwt::chunk_grid(N);
    bwt::bwt_gather_kernel<<<grid, bwt::kThreads, 0, stream>>>(d_in, sa, N, (u32)n, nullptr, nullptr, d_sa);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}
