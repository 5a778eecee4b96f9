// bzip2's Huffman stage on the GPU  (second half).
//
// Replaces sendMTFValues (cuda-bzip2-ipdpsw/compress.c:252-606) with BZ2_hbMakeCodeLengths /
// BZ2_hbAssignCodes (huffman.c:63-153), which one CPU thread runs per block:
//   init_kernel      2..6 tables by nMTF, initial tables = symbol ranges of equal total frequency
//   4 x { select_kernel   one thread per 50-symbol group: cost under every table from shared
//                         memory, FIRST cheapest table -> selector, the group's symbols counted
//                         into that table's frequencies (shared-memory histogram per CTA)
//         lengths_kernel  one thread per table: the reference's heap construction with its
//                         tie-breaks and its halve-and-retry rule for the 17-bit limit }
//   header_kernel    canonical codes; selector move-to-front ranks (one thread: a 6-entry list);
//                    symbol map, table count, selector count and the delta-coded code lengths
//                    written bit by bit (a few thousand bits)
//   pack kernels     selectors (unary) and symbols: bit lengths -> exclusive sum (devprims) ->
//                    every item ORs its bits into a zeroed array of MSB-first 32-bit units
//   bytes_kernel     units -> byte stream
// The bit string is what the reference hands to bsW between the block header and the block's
// end; the caller appends it to its own bit stream.
#include <mutex>

#include "common.cuh"
#include "devprims.cuh"
#include "../../include/b200lc.h"
#include "../../include/bzip2_gpu.h"

namespace b200lc {
namespace bzhuff {

constexpr int kMaxAlpha = 258, kTables = 6, kGroup = 50, kLimit = 17;

struct Ctx {
    int n_mtf, alpha, groups, nsel;
    int freq[kMaxAlpha];
    u8 in_use[256];
    u8 len[kTables][kMaxAlpha];
    int code[kTables][kMaxAlpha];
    int rfreq[kTables][kMaxAlpha];
    unsigned long long sel_start, table_start, data_start, total_bits;
};

__device__ __forceinline__ void put_bits(u32 *units, unsigned long long pos, u32 nb, u32 value, bool atomic)
{
    if (nb == 0) return;
    const u64 v = (u64)value << (64 - nb - (u32)(pos & 31));
    const u32 hi = (u32)(v >> 32), lo = (u32)v;
    u32 *p = units + (pos >> 5);
    if (atomic) {
        if (hi) atomicOr(p, hi);
        if (lo) atomicOr(p + 1, lo);
    } else {
        p[0] |= hi;
        if (lo) p[1] |= lo;
    }
}

__global__ void init_kernel(Ctx *c)
{
    if (threadIdx.x != 0) return;
    const int n = c->n_mtf, alpha = c->alpha;
    const int groups = n < 200 ? 2 : n < 600 ? 3 : n < 1200 ? 4 : n < 2400 ? 5 : 6;     // compress.c:274-279
    c->groups = groups;
    c->nsel = (n + kGroup - 1) / kGroup;
    for (int t = 0; t < kTables; ++t)
        for (int v = 0; v < alpha; ++v) c->len[t][v] = 15;
    int part = groups, remaining = n, first = 0;                                        // :282-319
    while (part > 0) {
        const int target = remaining / part;
        int last = first - 1, acc = 0;
        while (acc < target && last < alpha - 1) acc += c->freq[++last];
        if (last > first && part != groups && part != 1 && ((groups - part) % 2 == 1)) acc -= c->freq[last--];
        for (int v = 0; v < alpha; ++v) c->len[part - 1][v] = (v >= first && v <= last) ? 0 : 15;
        --part;
        first = last + 1;
        remaining -= acc;
    }
}

__global__ void __launch_bounds__(128) select_kernel(Ctx *c, const u16 *__restrict__ mtfv, u8 *__restrict__ selector)
{
    __shared__ u8 s_len[kTables][kMaxAlpha + 2];
    __shared__ int s_freq[kTables][kMaxAlpha];
    const int groups = c->groups, alpha = c->alpha, n = c->n_mtf;
    for (int i = threadIdx.x; i < kTables * kMaxAlpha; i += 128) {
        const int t = i / kMaxAlpha, v = i - t * kMaxAlpha;
        s_len[t][v] = c->len[t][v];
        s_freq[t][v] = 0;
    }
    __syncthreads();
    const int g = blockIdx.x * 128 + threadIdx.x;
    if (g < c->nsel) {
        const int lo = g * kGroup, hi = min(lo + kGroup, n);
        u32 cost[kTables];
#pragma unroll
        for (int t = 0; t < kTables; ++t) cost[t] = 0;
        for (int i = lo; i < hi; ++i) {
            const u32 sym = mtfv[i];
#pragma unroll
            for (int t = 0; t < kTables; ++t) cost[t] += s_len[t][sym];
        }
        int best = 0;
        u32 best_cost = 0xffffffffu;
#pragma unroll
        for (int t = 0; t < kTables; ++t)
            if (t < groups && (cost[t] & 0xffffu) < best_cost) { best_cost = cost[t] & 0xffffu; best = t; }   // UInt16 cost[] (:259)
        selector[g] = (u8)best;
        for (int i = lo; i < hi; ++i) atomicAdd(&s_freq[best][mtfv[i]], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < groups * kMaxAlpha; i += 128) {
        const int t = i / kMaxAlpha, v = i - t * kMaxAlpha;
        if (v < alpha && s_freq[t][v]) atomicAdd(&c->rfreq[t][v], s_freq[t][v]);
    }
}

// huffman.c:63-131, one thread per table
__global__ void lengths_kernel(Ctx *c)
{
    const int t = threadIdx.x;
    if (t >= c->groups) return;
    const int alpha = c->alpha;
    int heap[kMaxAlpha + 2], weight[kMaxAlpha * 2], parent[kMaxAlpha * 2];
    for (int i = 0; i < alpha; ++i) {
        const int f = c->rfreq[t][i];
        weight[i + 1] = (f == 0 ? 1 : f) << 8;
    }
    while (true) {
        int nodes = alpha, count = 0;
        heap[0] = 0; weight[0] = 0; parent[0] = -2;
        for (int i = 1; i <= alpha; ++i) {
            parent[i] = -1;
            int z = ++count;
            heap[z] = i;
            const int node = i;
            while (weight[node] < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
            heap[z] = node;
        }
        while (count > 1) {
            int pick[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                pick[r] = heap[1];
                heap[1] = heap[count--];
                int z = 1;
                const int node = heap[1];
                while (true) {
                    int child = z << 1;
                    if (child > count) break;
                    if (child < count && weight[heap[child + 1]] < weight[heap[child]]) ++child;
                    if (weight[node] < weight[heap[child]]) break;
                    heap[z] = heap[child];
                    z = child;
                }
                heap[z] = node;
            }
            ++nodes;
            parent[pick[0]] = parent[pick[1]] = nodes;
            const int wa = weight[pick[0]], wb = weight[pick[1]];
            weight[nodes] = (int)(((u32)wa & 0xffffff00u) + ((u32)wb & 0xffffff00u)) | (1 + max(wa & 0xff, wb & 0xff));
            parent[nodes] = -1;
            int z = ++count;
            heap[z] = nodes;
            while (weight[nodes] < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
            heap[z] = nodes;
        }
        bool too_long = false;
        for (int i = 1; i <= alpha; ++i) {
            int depth = 0;
            for (int k = i; parent[k] >= 0; k = parent[k]) ++depth;
            c->len[t][i - 1] = (u8)depth;
            too_long |= depth > kLimit;
        }
        if (!too_long) break;
        for (int i = 1; i <= alpha; ++i) weight[i] = (1 + (weight[i] >> 8) / 2) << 8;
    }
    for (int i = 0; i < alpha; ++i) c->rfreq[t][i] = 0;      // ready for the next round
}

__global__ void header_kernel(Ctx *c, const u8 *__restrict__ selector, u32 *__restrict__ sel_bits, u32 *units)
{
    const int t = threadIdx.x;
    const int alpha = c->alpha, groups = c->groups;
    if (t < groups) {                                        // huffman.c:134-153
        int lo = 32, hi = 0, next = 0;
        for (int i = 0; i < alpha; ++i) { const int l = c->len[t][i]; hi = max(hi, l); lo = min(lo, l); }
        for (int n = lo; n <= hi; ++n) {
            for (int i = 0; i < alpha; ++i)
                if (c->len[t][i] == n) c->code[t][i] = next++;
            next <<= 1;
        }
    }
    __syncwarp();
    if (t != 0) return;
    unsigned long long pos = 0;
    // symbol map (:498-516)
    u32 rows = 0;
    for (int r = 0; r < 16; ++r)
        for (int k = 0; k < 16; ++k)
            if (c->in_use[r * 16 + k]) rows |= 1u << r;
    for (int r = 0; r < 16; ++r) { put_bits(units, pos, 1, (rows >> r) & 1u, false); ++pos; }
    for (int r = 0; r < 16; ++r)
        if ((rows >> r) & 1u)
            for (int k = 0; k < 16; ++k) { put_bits(units, pos, 1, c->in_use[r * 16 + k] ? 1u : 0u, false); ++pos; }
    put_bits(units, pos, 5, (u32)groups, false); pos += 5;            // :524 (this fork: 5 bits)
    put_bits(units, pos, 17, (u32)c->nsel, false); pos += 17;          // :527 (this fork: 17 bits)
    c->sel_start = pos;
    // selector move-to-front ranks (:458-474); their unary codes are packed by sel_pack_kernel
    {
        u8 order[kTables];
        for (int k = 0; k < groups; ++k) order[k] = (u8)k;
        unsigned long long total = 0;
        for (int g = 0; g < c->nsel; ++g) {
            const u8 want = selector[g];
            int r = 0;
            u8 carry = order[0];
            while (carry != want) {
                ++r;
                const u8 nxt = order[r];
                order[r] = carry;
                carry = nxt;
            }
            order[0] = carry;
            sel_bits[g] = (u32)r + 1;
            total += (u32)r + 1;
        }
        pos += total;
    }
    c->table_start = pos;
    // code lengths, delta coded (:538-546)
    for (int tt = 0; tt < groups; ++tt) {
        int cur = c->len[tt][0];
        put_bits(units, pos, 5, (u32)cur, false); pos += 5;
        for (int i = 0; i < alpha; ++i) {
            const int want = c->len[tt][i];
            while (cur < want) { put_bits(units, pos, 2, 2, false); pos += 2; ++cur; }
            while (cur > want) { put_bits(units, pos, 2, 3, false); pos += 2; --cur; }
            put_bits(units, pos, 1, 0, false); ++pos;
        }
    }
    c->data_start = pos;
}

__global__ void __launch_bounds__(256) sel_pack_kernel(const Ctx *c, const u32 *__restrict__ sel_bits,
                                                       const u32 *__restrict__ sel_off, u32 *units)
{
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= c->nsel) return;
    const u32 nb = sel_bits[g];                                        // rank + 1: rank ones, then a zero (:529-532)
    put_bits(units, c->sel_start + sel_off[g], nb, ((1u << (nb - 1)) - 1u) << 1, true);
}

__global__ void __launch_bounds__(256) data_len_kernel(const Ctx *c, const u16 *__restrict__ mtfv,
                                                       const u8 *__restrict__ selector, u32 *__restrict__ nbits)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= c->n_mtf) return;
    nbits[i] = c->len[selector[i / kGroup]][mtfv[i]];
}

__global__ void __launch_bounds__(256) data_pack_kernel(Ctx *c, const u16 *__restrict__ mtfv,
                                                        const u8 *__restrict__ selector, const u32 *__restrict__ off,
                                                        u32 *units)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= c->n_mtf) return;
    const int t = selector[i / kGroup];
    const u32 sym = mtfv[i];
    const u32 nb = c->len[t][sym];
    const unsigned long long pos = c->data_start + off[i];
    put_bits(units, pos, nb, (u32)c->code[t][sym], true);                  // :556-600
    if (i == c->n_mtf - 1) c->total_bits = pos + nb;
}

__global__ void __launch_bounds__(256) bytes_kernel(const u32 *__restrict__ units, u32 nunits, u32 *__restrict__ out)
{
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i < nunits) out[i] = __byte_perm(units[i], 0, 0x0123);
}

struct Work {
    Ctx *d_ctx = nullptr;
    u16 *d_mtfv = nullptr;
    u8 *d_selector = nullptr;
    u32 *d_sel_bits = nullptr, *d_sym_bits = nullptr, *d_units = nullptr, *d_bytes = nullptr;
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0, cap = 0;
    void release()
    {
        cudaFree(d_ctx); cudaFree(d_mtfv); cudaFree(d_selector); cudaFree(d_sel_bits); cudaFree(d_sym_bits);
        cudaFree(d_units); cudaFree(d_bytes); cudaFree(d_scratch);
        *this = Work();
    }
};
static Work g_work;
static std::mutex g_lock;

static size_t unit_capacity(size_t n) { return (n * kLimit + 8 * n / kGroup + 65536) / 32 + 4; }

static int ensure(size_t n)
{
    if (g_work.cap >= n) return B200LC_OK;
    g_work.release();
    Work &w = g_work;
    w.scratch_bytes = prims::scan_scratch_bytes(n) + 256;
    B200LC_CUDA_TRY(cudaMalloc(&w.d_ctx, sizeof(Ctx)));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_mtfv, n * 2));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_selector, n / kGroup + 2));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_sel_bits, (n / kGroup + 2) * 4));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_sym_bits, n * 4));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_units, unit_capacity(n) * 4));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_bytes, unit_capacity(n) * 4));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_scratch, w.scratch_bytes));
    w.cap = n;
    return B200LC_OK;
}

}  // namespace bzhuff
}  // namespace b200lc

using namespace b200lc;

extern "C" int b200lc_bzip2_send_mtf_values(const unsigned short *mtfv, int n_mtf, const int *mtf_freq,
                                            const unsigned char *in_use, int n_in_use, unsigned char *bits,
                                            size_t bits_cap, unsigned long long *n_bits,
                                            unsigned char *len_out, unsigned char *selector_out)
{
    using namespace bzhuff;
    if (!mtfv || !mtf_freq || !in_use || !bits || !n_bits || n_mtf <= 0) return B200LC_ERR_ARG;
    if (n_in_use < 1 || n_in_use > 256) return B200LC_ERR_ARG;
    const size_t n = (size_t)n_mtf;
    std::lock_guard<std::mutex> guard(g_lock);
    int rc = ensure(n);
    if (rc) return rc;
    Work &w = g_work;
    static Ctx h;                         // guarded by g_lock
    memset(&h, 0, sizeof(h));
    h.n_mtf = n_mtf;
    h.alpha = n_in_use + 2;
    memcpy(h.freq, mtf_freq, (size_t)h.alpha * sizeof(int));
    memcpy(h.in_use, in_use, 256);
    const size_t nunits = unit_capacity(n);
    B200LC_CUDA_TRY(cudaMemcpy(w.d_ctx, &h, sizeof(h), cudaMemcpyHostToDevice));
    B200LC_CUDA_TRY(cudaMemcpy(w.d_mtfv, mtfv, n * 2, cudaMemcpyHostToDevice));
    B200LC_CUDA_TRY(cudaMemset(w.d_units, 0, nunits * 4));
    const int nsel = (n_mtf + kGroup - 1) / kGroup;
    init_kernel<<<1, 32>>>(w.d_ctx);
    for (int round = 0; round < 4; ++round) {                          // BZ_N_ITERS (:324)
        select_kernel<<<(nsel + 127) / 128, 128>>>(w.d_ctx, w.d_mtfv, w.d_selector);
        lengths_kernel<<<1, 32>>>(w.d_ctx);
    }
    header_kernel<<<1, 32>>>(w.d_ctx, w.d_selector, w.d_sel_bits, w.d_units);
    B200LC_CUDA_TRY(cudaGetLastError());
    rc = prims::exclusive_sum_u32(w.d_sel_bits, w.d_sym_bits, (u64)nsel, w.d_scratch, w.scratch_bytes, nullptr);
    if (rc) return rc;
    sel_pack_kernel<<<(nsel + 255) / 256, 256>>>(w.d_ctx, w.d_sel_bits, w.d_sym_bits, w.d_units);
    data_len_kernel<<<(n_mtf + 255) / 256, 256>>>(w.d_ctx, w.d_mtfv, w.d_selector, w.d_sym_bits);
    rc = prims::exclusive_sum_u32(w.d_sym_bits, w.d_sym_bits, n, w.d_scratch, w.scratch_bytes, nullptr);
    if (rc) return rc;
    data_pack_kernel<<<(n_mtf + 255) / 256, 256>>>(w.d_ctx, w.d_mtfv, w.d_selector, w.d_sym_bits, w.d_units);
    bytes_kernel<<<(u32)((nunits + 255) / 256), 256>>>(w.d_units, (u32)nunits, w.d_bytes);
    B200LC_CUDA_TRY(cudaGetLastError());
    B200LC_CUDA_TRY(cudaMemcpy(&h, w.d_ctx, sizeof(h), cudaMemcpyDeviceToHost));
    const size_t nbytes = (size_t)((h.total_bits + 7) / 8);
    if (nbytes > bits_cap || nbytes > nunits * 4) return B200LC_ERR_OVERFLOW;
    B200LC_CUDA_TRY(cudaMemcpy(bits, w.d_bytes, nbytes, cudaMemcpyDeviceToHost));
    *n_bits = h.total_bits;
    if (len_out) memcpy(len_out, h.len, sizeof(h.len));
    if (selector_out) B200LC_CUDA_TRY(cudaMemcpy(selector_out, w.d_selector, (size_t)nsel, cudaMemcpyDeviceToHost));
    return B200LC_OK;
}
