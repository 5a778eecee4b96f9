// Segmented one-sweep radix sort and single-pass scans for sm_100a (see devprims.cuh).
//
// Sort: least-significant-digit radix sort, 8 bits per pass.
//   * one histogram kernel counts the digits of EVERY pass per segment in one read of the keys;
//     a tiny kernel turns the counts into global start offsets;
//   * one kernel per pass ("one sweep"): a CTA takes a tile of 4096 pairs through a ticket,
//     stages it into shared memory with two 1-D TMA bulk copies, ranks the keys warp by warp
//     (match.any on the digit: peers of a digit share one counter update, which keeps the sort
//     stable), publishes the tile's 256 digit counts right away, reorders the tile in shared
//     memory, and only then resolves the counts of all earlier tiles of its segment by a
//     decoupled look-back (one thread per digit; by now their counts have had time to land)
//     and writes every digit run to its final place with coalesced stores.
//     Keys and values cross HBM once per pass in each direction.
//   * segments (independent BWT blocks) never exchange elements: tiles do not straddle segment
//     boundaries and the look-back stops at the first tile of a segment, so the block number
//     needs no key bits and no passes.
// Scans: 4096 elements per CTA, tile prefix by decoupled look-back over 64-bit status words.
#include "devprims.cuh"

namespace b200lc {
namespace prims {

// ------------------------------------------------------------------------------------ sort
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;   // 4096 pairs per CTA
constexpr int kHistChunk = 16384;                       // keys per CTA of the histogram kernel
constexpr int kMaxPasses = 8;
constexpr u32 kFlagAgg = 1u << 30;    // status word: tile's own digit count
constexpr u32 kFlagIncl = 2u << 30;   //              count of this and all earlier tiles of the segment
constexpr u32 kValueMask = (1u << 30) - 1;

template <typename K>
__device__ __forceinline__ u32 digit_of(K key, int shift, u32 mask)
{
    return (u32)(key >> shift) & mask;
}

// exclusive prefix over the 256 threads of a CTA (ws: kSortWarps words of shared memory)
__device__ __forceinline__ u32 block_excl_sum_256(u32 v, u32 *ws)
{
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 incl = warp_incl_scan(v);
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    u32 pre = 0;
#pragma unroll
    for (u32 w = 0; w < (u32)kSortWarps; ++w)
        if (w < warp) pre += ws[w];
    __syncthreads();
    return pre + incl - v;
}

template <typename K>
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const K *__restrict__ keys, u64 n,
                                                                 u64 seg_len, u32 chunks_per_seg,
                                                                 int begin_bit, int end_bit, int passes,
                                                                 u32 *__restrict__ hist)
{
    __shared__ u32 sh[kMaxPasses][256];
    const u32 tid = threadIdx.x;
    const u32 seg = blockIdx.x / chunks_per_seg, chunk = blockIdx.x - seg * chunks_per_seg;
    const u64 seg_begin = (u64)seg * seg_len;
    const u64 seg_end = min(seg_begin + seg_len, n);
    const u64 lo = seg_begin + (u64)chunk * kHistChunk;
    if (lo >= seg_end) return;
    const u32 cnt = (u32)min((u64)kHistChunk, seg_end - lo);
    for (u32 i = tid; i < (u32)passes * 256; i += kSortThreads) (&sh[0][0])[i] = 0;
    __syncthreads();
    const u32 full = cnt & ~31u;
    for (u32 j = tid; j < full; j += kSortThreads) {       // whole warps active
        const K k = keys[lo + j];
        for (int p = 0; p < passes; ++p) {
            const int shift = begin_bit + 8 * p;
            const u32 mask = (1u << min(8, end_bit - shift)) - 1;
            const u32 dg = digit_of(k, shift, mask);
            int same;
            __match_all_sync(0xffffffffu, dg, &same);
            if (same) {
                if ((tid & 31) == 0) atomicAdd(&sh[p][dg], 32u);
            } else {
                atomicAdd(&sh[p][dg], 1u);
            }
        }
    }
    for (u32 j = full + tid; j < cnt; j += kSortThreads) {
        const K k = keys[lo + j];
        for (int p = 0; p < passes; ++p) {
            const int shift = begin_bit + 8 * p;
            const u32 mask = (1u << min(8, end_bit - shift)) - 1;
            atomicAdd(&sh[p][digit_of(k, shift, mask)], 1u);
        }
    }
    __syncthreads();
    for (u32 i = tid; i < (u32)passes * 256; i += kSortThreads) {
        const u32 c = (&sh[0][0])[i];
        if (c) atomicAdd(&hist[(size_t)seg * passes * 256 + i], c);
    }
}

// counts -> start offset of every (segment, pass, digit) in the output array
__global__ void __launch_bounds__(kSortThreads) sort_offsets_kernel(u32 *__restrict__ hist, u64 seg_len,
                                                                    int passes)
{
    __shared__ u32 ws[kSortWarps];
    const u32 seg = blockIdx.x / passes;
    u32 *h = hist + (size_t)blockIdx.x * 256;
    const u32 c = h[threadIdx.x];
    const u32 ex = block_excl_sum_256(c, ws);
    h[threadIdx.x] = (u32)((u64)seg * seg_len) + ex;
}

template <typename K>
struct SortArgs {
    const K *kin;
    K *kout;
    const u32 *vin;
    u32 *vout;
    u64 n, seg_len;
    u32 tiles_per_seg;
    u32 *status;        // [tiles][256]
    u32 *ticket;
    const u32 *base;    // [segments][passes][256]
    int passes, pass, shift;
    u32 mask;
    int tma_ok;
};

template <typename K>
struct SortSmem {
    alignas(128) K keys[kSortTile];
    alignas(128) u32 vals[kSortTile];
    u16 warp_cnt[kSortWarps][256];   // per-warp digit counts (<= 512), later offsets inside the tile (< 4096)
    u32 digit_start[256];
    u32 gbase[256];
    u32 ws[kSortWarps];
    alignas(8) u64 bar;
    u32 tile;
};

template <typename K>
__global__ void __launch_bounds__(kSortThreads, 4) sort_onesweep_kernel(const SortArgs<K> a)
{
    extern __shared__ __align__(128) unsigned char sort_smem_raw[];
    SortSmem<K> &sm = *reinterpret_cast<SortSmem<K> *>(sort_smem_raw);
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid == 0) {
        sm.tile = atomicAdd(a.ticket, 1u);
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
    }
    for (u32 i = lane; i < 128; i += 32) reinterpret_cast<u32 *>(sm.warp_cnt[warp])[i] = 0;
    __syncthreads();

    const u32 tile = sm.tile;
    const u32 seg = tile / a.tiles_per_seg, tseg = tile - seg * a.tiles_per_seg;
    const u64 seg_begin = (u64)seg * a.seg_len;
    const u64 seg_end = min(seg_begin + a.seg_len, a.n);
    const u64 first = seg_begin + (u64)tseg * kSortTile;
    const u32 tile_n = (u32)min((u64)kSortTile, seg_end - first);

    // ---- stage the tile: TMA bulk copies for whole aligned tiles, guarded loads otherwise
    if (a.tma_ok && tile_n == (u32)kSortTile && (first & 3) == 0) {
        if (tid == 0) {
            mbar_expect_tx(&sm.bar, (u32)(kSortTile * (sizeof(K) + 4)));
            tma_load_1d(sm.keys, a.kin + first, (u32)(kSortTile * sizeof(K)), &sm.bar);
            tma_load_1d(sm.vals, a.vin + first, (u32)(kSortTile * 4), &sm.bar);
        }
        mbar_wait(&sm.bar, 0);
    } else {
        for (u32 j = tid; j < (u32)kSortTile; j += kSortThreads) {
            const bool ok = j < tile_n;
            sm.keys[j] = ok ? a.kin[first + j] : ~K(0);    // padding sorts behind every real key
            sm.vals[j] = ok ? a.vin[first + j] : 0u;
        }
        __syncthreads();
    }

    // ---- rank: item i of lane l of warp w is element w*512 + i*32 + l of the tile
    K key[kSortItems];
    u32 pos[kSortItems];
    const u32 my0 = warp * (32 * kSortItems) + lane;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) key[i] = sm.keys[my0 + i * 32];
    u16 *wc = sm.warp_cnt[warp];
    const u32 lt = (1u << lane) - 1;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const u32 dg = digit_of(key[i], a.shift, a.mask);
        const u32 peers = __match_any_sync(0xffffffffu, dg);
        const u32 leader = 31u - (u32)__clz(peers);
        u32 old = 0;
        if (lane == leader) {
            old = wc[dg];
            wc[dg] = (u16)(old + (u32)__popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        pos[i] = old + (u32)__popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();   // every key is in registers, warp counters are final

    // ---- thread d owns digit d: warp offsets, tile count; the count is published right away so
    // that later tiles can start summing while this one reorders its data
    const u32 d = tid;
    u32 count = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
        const u32 t = sm.warp_cnt[w][d];
        sm.warp_cnt[w][d] = (u16)count;
        count += t;
    }
    u32 *st = a.status + (size_t)tile * 256;
    const bool seg_first = tseg == 0;
    st_relaxed_u32(&st[d], (seg_first ? kFlagIncl : kFlagAgg) | count);
    const u32 gstart = a.base[((size_t)seg * a.passes + a.pass) * 256 + d];
    const u32 start = block_excl_sum_256(count, sm.ws);
    sm.digit_start[d] = start;
    __syncthreads();

    // ---- reorder inside shared memory
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const u32 dg = digit_of(key[i], a.shift, a.mask);
        const u32 p = sm.digit_start[dg] + wc[dg] + pos[i];
        sm.keys[p] = key[i];
        pos[i] = p;
    }
    u32 val[kSortItems];
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) val[i] = sm.vals[my0 + i * 32];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) sm.vals[pos[i]] = val[i];

    // ---- look-back over the earlier tiles of the segment (their counts have had time to land)
    u32 excl = 0;
    if (!seg_first) {
        const u32 *prev = st + d - 256;
        while (true) {
            u32 s;
            do {
                s = ld_relaxed_u32(prev);
            } while ((s >> 30) == 0);
            excl += s & kValueMask;
            if (s & kFlagIncl) break;
            prev -= 256;
        }
        st_relaxed_u32(&st[d], kFlagIncl | (excl + count));
    }
    sm.gbase[d] = gstart + excl - start;
    __syncthreads();

    // ---- digit runs leave with consecutive addresses
    for (u32 j = tid; j < tile_n; j += kSortThreads) {
        const K k = sm.keys[j];
        const u32 dst = sm.gbase[digit_of(k, a.shift, a.mask)] + j;
        a.kout[dst] = k;
        a.vout[dst] = sm.vals[j];
    }
}

struct SortLayout {
    u64 nseg, tiles;
    u32 tiles_per_seg;
    size_t status, hist, total;
};

static SortLayout sort_layout(u64 n, u64 seg_len)
{
    SortLayout L;
    if (seg_len == 0 || seg_len > n) seg_len = n ? n : 1;
    L.nseg = n ? (n + seg_len - 1) / seg_len : 1;
    L.tiles_per_seg = (u32)((seg_len + kSortTile - 1) / kSortTile);
    const u64 last = n - (L.nseg - 1) * seg_len;
    L.tiles = (L.nseg - 1) * L.tiles_per_seg + (last + kSortTile - 1) / kSortTile;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    L.status = 256;
    L.hist = L.status + up(L.tiles * 1024);
    L.total = L.hist + up(L.nseg * kMaxPasses * 1024);
    return L;
}

size_t sort_scratch_bytes(u64 n, u64 seg_len) { return sort_layout(n, seg_len).total; }

template <typename K>
int sort_pairs(K *keys_a, K *keys_b, u32 *vals_a, u32 *vals_b, u64 n, u64 seg_len, int begin_bit,
               int end_bit, void *scratch, size_t scratch_bytes, cudaStream_t stream, int *result_in_b)
{
    *result_in_b = 0;
    if (n == 0 || end_bit <= begin_bit) return B200LC_OK;
    if (n > kSortMaxElems) return B200LC_ERR_UNSUPPORTED;
    if (begin_bit < 0 || end_bit > (int)(8 * sizeof(K))) return B200LC_ERR_ARG;
    if (!keys_a || !keys_b || !vals_a || !vals_b || !scratch) return B200LC_ERR_ARG;
    if (seg_len == 0 || seg_len > n) seg_len = n;
    const int passes = (end_bit - begin_bit + 7) / 8;
    const SortLayout L = sort_layout(n, seg_len);
    if (scratch_bytes < L.total) return B200LC_ERR_SCRATCH;
    if (L.tiles >= (1ull << 31) || L.nseg * passes >= (1ull << 31)) return B200LC_ERR_UNSUPPORTED;
    char *base = reinterpret_cast<char *>(scratch);
    u32 *ticket = reinterpret_cast<u32 *>(base);
    u32 *status = reinterpret_cast<u32 *>(base + L.status);
    u32 *hist = reinterpret_cast<u32 *>(base + L.hist);

    static unsigned attr_done[kMaxDevices] = {0};   // context epoch the attribute was set in
    const int slot = device_slot();
    const size_t smem = sizeof(SortSmem<K>);
    if (slot < 0 || attr_done[slot] != context_epoch()) {
        B200LC_CUDA_TRY(cudaFuncSetAttribute(sort_onesweep_kernel<K>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (slot >= 0) attr_done[slot] = context_epoch();
    }

    B200LC_CUDA_TRY(cudaMemsetAsync(hist, 0, L.nseg * passes * 1024, stream));
    const u32 chunks_per_seg = (u32)((seg_len + kHistChunk - 1) / kHistChunk);
    sort_hist_kernel<K><<<(u32)(L.nseg * chunks_per_seg), kSortThreads, 0, stream>>>(
        keys_a, n, seg_len, chunks_per_seg, begin_bit, end_bit, passes, hist);
    B200LC_CUDA_TRY(cudaGetLastError());
    sort_offsets_kernel<<<(u32)(L.nseg * passes), kSortThreads, 0, stream>>>(hist, seg_len, passes);
    B200LC_CUDA_TRY(cudaGetLastError());

    SortArgs<K> a;
    a.n = n;
    a.seg_len = seg_len;
    a.tiles_per_seg = L.tiles_per_seg;
    a.status = status;
    a.ticket = ticket;
    a.base = hist;
    a.passes = passes;
    K *kin = keys_a, *kout = keys_b;
    u32 *vin = vals_a, *vout = vals_b;
    for (int p = 0; p < passes; ++p) {
        a.kin = kin; a.kout = kout; a.vin = vin; a.vout = vout;
        a.pass = p;
        a.shift = begin_bit + 8 * p;
        a.mask = (1u << (end_bit - a.shift < 8 ? end_bit - a.shift : 8)) - 1;
        a.tma_ok = ((reinterpret_cast<uintptr_t>(kin) | reinterpret_cast<uintptr_t>(vin)) & 15) == 0 &&
                   (seg_len % 4 == 0 || L.nseg == 1);
        B200LC_CUDA_TRY(cudaMemsetAsync(base, 0, L.status + L.tiles * 1024, stream));
        sort_onesweep_kernel<K><<<(u32)L.tiles, kSortThreads, smem, stream>>>(a);
        B200LC_CUDA_TRY(cudaGetLastError());
        K *tk = kin; kin = kout; kout = tk;
        u32 *tv = vin; vin = vout; vout = tv;
    }
    *result_in_b = passes & 1;
    return B200LC_OK;
}

template int sort_pairs<u32>(u32 *, u32 *, u32 *, u32 *, u64, u64, int, int, void *, size_t, cudaStream_t, int *);
template int sort_pairs<u64>(u64 *, u64 *, u32 *, u32 *, u64, u64, int, int, void *, size_t, cudaStream_t, int *);

// ------------------------------------------------------------------------------------ scans
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;
constexpr u64 kScanAgg = 1ull << 32;
constexpr u64 kScanIncl = 2ull << 32;

template <bool IS_MAX>
__device__ __forceinline__ u32 scan_op(u32 a, u32 b)
{
    return IS_MAX ? (a > b ? a : b) : a + b;
}

// IS_MAX: inclusive running maximum; otherwise exclusive sum.  0 is the identity of both.
template <bool IS_MAX>
__global__ void __launch_bounds__(kScanThreads) scan_kernel(const u32 *in, u32 *out, u64 n,
                                                            u64 *status, u32 *ticket)
{
    __shared__ u32 s_tile, s_prefix;
    __shared__ u32 ws[kScanThreads / 32];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = s_tile;
    const u64 first = (u64)tile * kScanTile;
    const u32 tile_n = (u32)min((u64)kScanTile, n - first);
    const u32 off = tid * kScanItems;
    const bool vec = tile_n == (u32)kScanTile &&
                     ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;

    u32 x[kScanItems];
    if (vec) {
        const uint4 *src = reinterpret_cast<const uint4 *>(in + first + off);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; ++q) {
            const uint4 v = src[q];
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i) x[i] = off + i < tile_n ? in[first + off + i] : 0u;
    }
    u32 agg = x[0];
#pragma unroll
    for (int i = 1; i < kScanItems; ++i) agg = scan_op<IS_MAX>(agg, x[i]);

    // inclusive scan of the thread aggregates inside the warp, then across warps
    u32 incl = agg;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, incl, dd);
        if (lane >= (u32)dd) incl = scan_op<IS_MAX>(incl, t);
    }
    u32 before = __shfl_up_sync(0xffffffffu, incl, 1);   // exclusive inside the warp
    if (lane == 0) before = 0;
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    u32 pre = 0, total = 0;
#pragma unroll
    for (u32 w = 0; w < (u32)(kScanThreads / 32); ++w) {
        const u32 v = ws[w];
        if (w < warp) pre = scan_op<IS_MAX>(pre, v);
        total = scan_op<IS_MAX>(total, v);
    }

    if (tid == 0) {
        u32 acc = 0;
        if (tile == 0) {
            st_relaxed_u64(&status[0], kScanIncl | total);
        } else {
            st_relaxed_u64(&status[tile], kScanAgg | total);
            const u64 *prev = &status[tile - 1];
            while (true) {
                u64 s;
                do {
                    s = ld_relaxed_u64(prev);
                } while ((s >> 32) == 0);
                acc = scan_op<IS_MAX>(acc, (u32)s);
                if (s & kScanIncl) break;
                --prev;
            }
            st_relaxed_u64(&status[tile], kScanIncl | scan_op<IS_MAX>(acc, total));
        }
        s_prefix = acc;
    }
    __syncthreads();
    u32 run = scan_op<IS_MAX>(s_prefix, scan_op<IS_MAX>(pre, before));
    u32 y[kScanItems];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (IS_MAX) {
            run = scan_op<true>(run, x[i]);
            y[i] = run;
        } else {
            y[i] = run;
            run += x[i];
        }
    }
    if (vec) {
        uint4 *dst = reinterpret_cast<uint4 *>(out + first + off);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; ++q)
            dst[q] = make_uint4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < kScanItems; ++i)
            if (off + i < tile_n) out[first + off + i] = y[i];
    }
}

size_t scan_scratch_bytes(u64 n)
{
    const u64 tiles = (n + kScanTile - 1) / kScanTile;
    return 256 + ((tiles * 8 + 255) & ~size_t(255));
}

template <bool IS_MAX>
static int run_scan(const u32 *in, u32 *out, u64 n, void *scratch, size_t scratch_bytes, cudaStream_t stream)
{
    if (n == 0) return B200LC_OK;
    if (!in || !out || !scratch) return B200LC_ERR_ARG;
    const size_t need = scan_scratch_bytes(n);
    if (scratch_bytes < need) return B200LC_ERR_SCRATCH;
    const u64 tiles = (n + kScanTile - 1) / kScanTile;
    if (tiles >= (1ull << 31)) return B200LC_ERR_UNSUPPORTED;
    B200LC_CUDA_TRY(cudaMemsetAsync(scratch, 0, 256 + tiles * 8, stream));
    u32 *ticket = reinterpret_cast<u32 *>(scratch);
    u64 *status = reinterpret_cast<u64 *>(reinterpret_cast<char *>(scratch) + 256);
    scan_kernel<IS_MAX><<<(u32)tiles, kScanThreads, 0, stream>>>(in, out, n, status, ticket);
    B200LC_CUDA_TRY(cudaGetLastError());
    return B200LC_OK;
}

int exclusive_sum_u32(const u32 *in, u32 *out, u64 n, void *scratch, size_t scratch_bytes, cudaStream_t stream)
{
    return run_scan<false>(in, out, n, scratch, scratch_bytes, stream);
}
int inclusive_max_u32(const u32 *in, u32 *out, u64 n, void *scratch, size_t scratch_bytes, cudaStream_t stream)
{
    return run_scan<true>(in, out, n, scratch, scratch_bytes, stream);
}

}  // namespace prims
}  // namespace b200lc

// ------------------------------------------------------------------------------------ C ABI
using namespace b200lc;

extern "C" size_t b200lc_sort_scratch_bytes(size_t n, size_t seg_len)
{
    return prims::sort_scratch_bytes(n, seg_len);
}

extern "C" int b200lc_sort_pairs_u64(uint64_t *d_keys_a, uint64_t *d_keys_b, uint32_t *d_vals_a,
                                     uint32_t *d_vals_b, size_t n, size_t seg_len, int begin_bit,
                                     int end_bit, void *d_scratch, size_t scratch_bytes, void *stream,
                                     int *result_in_b)
{
    if (!result_in_b) return B200LC_ERR_ARG;
    return prims::sort_pairs<u64>(d_keys_a, d_keys_b, d_vals_a, d_vals_b, n, seg_len, begin_bit, end_bit,
                                  d_scratch, scratch_bytes, (cudaStream_t)stream, result_in_b);
}

extern "C" int b200lc_sort_pairs_u32(uint32_t *d_keys_a, uint32_t *d_keys_b, uint32_t *d_vals_a,
                                     uint32_t *d_vals_b, size_t n, size_t seg_len, int begin_bit,
                                     int end_bit, void *d_scratch, size_t scratch_bytes, void *stream,
                                     int *result_in_b)
{
    if (!result_in_b) return B200LC_ERR_ARG;
    return prims::sort_pairs<u32>(d_keys_a, d_keys_b, d_vals_a, d_vals_b, n, seg_len, begin_bit, end_bit,
                                  d_scratch, scratch_bytes, (cudaStream_t)stream, result_in_b);
}

extern "C" size_t b200lc_scan_scratch_bytes(size_t n) { return prims::scan_scratch_bytes(n); }

extern "C" int b200lc_exclusive_sum_u32(const uint32_t *d_in, uint32_t *d_out, size_t n, void *d_scratch,
                                        size_t scratch_bytes, void *stream)
{
    return prims::exclusive_sum_u32(d_in, d_out, n, d_scratch, scratch_bytes, (cudaStream_t)stream);
}

extern "C" int b200lc_inclusive_max_u32(const uint32_t *d_in, uint32_t *d_out, size_t n, void *d_scratch,
                                        size_t scratch_bytes, void *stream)
{
    return prims::inclusive_max_u32(d_in, d_out, n, d_scratch, scratch_bytes, (cudaStream_t)stream);
}
