// bzip2's MTF + zero-run (RUNA/RUNB) stage on the GPU  (first half).
//
// Replaces generateMTFValues (cuda-bzip2-ipdpsw/compress.c:122-246), which one CPU thread runs
// per block after the GPU sort: last column of the sorted rotations in the block's dense
// alphabet -> move-to-front ranks -> runs of rank 0 written in bijective base 2 with the symbols
// RUNA/RUNB, every other rank r as r + 1, then EOB; plus the frequency table the Huffman stage
// starts from.
//
//   1. last_column_kernel   ll[i] = unseqToSeq[block[ptr[i] - 1]]           (compress.c:169-170)
//   2. MTF                  the batched kernel of mtf.cu; its initial list 0..255 gives the same
//                           ranks as bzip2's 0..nInUse-1 because the unused values never occur
//                           and therefore never move in front of a used one
//   3. RLE as scan passes   running maximum of "index of the last non-zero rank" gives every
//                           zero its run length; the run's LAST zero owns the run's digits
//                           (floor(log2(len + 1)) of them), every non-zero rank owns one symbol;
//                           an exclusive sum of those counts is the write offset
//   4. emit_kernel          writes the symbols and counts them (shared-memory histogram)
#include <mutex>

#include "common.cuh"
#include "devprims.cuh"
#include "../../include/b200lc.h"
#include "../../include/bzip2_gpu.h"

namespace b200lc {
namespace bzmtf {

constexpr u32 kRunA = 0, kRunB = 1;   // BZ_RUNA / BZ_RUNB (bzlib_private.h)

__global__ void __launch_bounds__(256) last_column_kernel(const u8 *__restrict__ block, const u32 *__restrict__ ptr,
                                                          u32 n, const u8 *__restrict__ seq, u8 *__restrict__ ll)
{
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const u32 p = ptr[i];
    ll[i] = seq[block[p ? p - 1 : n - 1]];
}

// mark[i] = i + 1 where rank[i] != 0, else 0  (running maximum = 1 + index of the last non-zero)
__global__ void __launch_bounds__(256) mark_kernel(const u8 *__restrict__ rank, u32 n, u32 *__restrict__ mark)
{
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) mark[i] = rank[i] ? i + 1 : 0u;
}

__device__ __forceinline__ u32 run_digits(u32 len) { return 31u - (u32)__clz(len + 1u); }

__global__ void __launch_bounds__(256) count_kernel(const u8 *__restrict__ rank, const u32 *__restrict__ lastnz,
                                                    u32 n, u32 *__restrict__ cnt)
{
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    u32 c = 1;
    if (rank[i] == 0) {
        const bool run_end = i + 1 == n || rank[i + 1] != 0;
        c = run_end ? run_digits(i + 1 - lastnz[i]) : 0u;
    }
    cnt[i] = c;
}

__global__ void __launch_bounds__(256) emit_kernel(const u8 *__restrict__ rank, const u32 *__restrict__ lastnz,
                                                   const u32 *__restrict__ off, u32 n, u32 eob,
                                                   u16 *__restrict__ mtfv, int *__restrict__ freq,
                                                   int *__restrict__ n_mtf)
{
    __shared__ int sh[258];
    for (u32 k = threadIdx.x; k < 258; k += 256) sh[k] = 0;
    __syncthreads();
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const u32 r = rank[i];
        u32 o = off[i];
        if (r) {
            mtfv[o] = (u16)(r + 1);
            atomicAdd(&sh[r + 1], 1);
            ++o;
        } else if (i + 1 == n || rank[i + 1] != 0) {
            u32 z = i + 1 - lastnz[i] - 1;               // compress.c:178-190
            while (true) {
                const u32 sym = (z & 1) ? kRunB : kRunA;
                mtfv[o++] = (u16)sym;
                atomicAdd(&sh[sym], 1);
                if (z < 2) break;
                z = (z - 2) >> 1;
            }
        }
        if (i + 1 == n) {
            mtfv[o] = (u16)eob;                          // compress.c:243
            atomicAdd(&sh[eob], 1);
            *n_mtf = (int)(o + 1);
        }
    }
    __syncthreads();
    for (u32 k = threadIdx.x; k < 258; k += 256)
        if (sh[k]) atomicAdd(&freq[k], sh[k]);
}

struct Work {
    u8 *d_block = nullptr, *d_ll = nullptr, *d_rank = nullptr, *d_seq = nullptr;
    u32 *d_ptr = nullptr, *d_a = nullptr, *d_b = nullptr;
    u16 *d_mtfv = nullptr;
    int *d_small = nullptr;     // [0..257] freq, [258] nMTF
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0, cap = 0;
    void release()
    {
        cudaFree(d_block); cudaFree(d_ll); cudaFree(d_rank); cudaFree(d_seq); cudaFree(d_ptr); cudaFree(d_a);
        cudaFree(d_b); cudaFree(d_mtfv); cudaFree(d_small); cudaFree(d_scratch);
        *this = Work();
    }
};
static Work g_work;
static std::mutex g_lock;   // worker threads of the reference compress blocks concurrently (compress.c:898-930)

static int ensure(size_t n)
{
    if (g_work.cap >= n) return B200LC_OK;
    g_work.release();
    Work &w = g_work;
    w.scratch_bytes = b200lc_mtf_scratch_bytes(1, n);
    const size_t scan = prims::scan_scratch_bytes(n);
    if (scan > w.scratch_bytes) w.scratch_bytes = scan;
    w.scratch_bytes += 256;
    B200LC_CUDA_TRY(cudaMalloc(&w.d_block, n));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_ll, n));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_rank, n));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_seq, 256));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_ptr, n * 4));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_a, n * 4));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_b, n * 4));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_mtfv, (n + 1) * 2));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_small, 260 * sizeof(int)));
    B200LC_CUDA_TRY(cudaMalloc(&w.d_scratch, w.scratch_bytes));
    w.cap = n;
    return B200LC_OK;
}

}  // namespace bzmtf
}  // namespace b200lc

using namespace b200lc;

extern "C" int b200lc_bzip2_mtf_rle(const unsigned char *block, const unsigned int *ptr, int nblock,
                                    const unsigned char *in_use, unsigned short *mtfv, int *n_mtf,
                                    int *mtf_freq, int *n_in_use)
{
    if (!block || !ptr || !in_use || !mtfv || !n_mtf || !mtf_freq || nblock <= 0) return B200LC_ERR_ARG;
    const u32 n = (u32)nblock;
    u8 seq[256];
    int used = 0;
    for (int i = 0; i < 256; ++i) seq[i] = in_use[i] ? (u8)used++ : (u8)0;      // makeMaps_e, compress.c:109-118
    if (n_in_use) *n_in_use = used;
    if (used == 0) return B200LC_ERR_ARG;
    const u32 eob = (u32)used + 1;

    std::lock_guard<std::mutex> guard(bzmtf::g_lock);
    int rc = bzmtf::ensure(n);
    if (rc) return rc;
    bzmtf::Work &w = bzmtf::g_work;
    B200LC_CUDA_TRY(cudaMemcpy(w.d_block, block, n, cudaMemcpyHostToDevice));
    B200LC_CUDA_TRY(cudaMemcpy(w.d_ptr, ptr, (size_t)n * 4, cudaMemcpyHostToDevice));
    B200LC_CUDA_TRY(cudaMemcpy(w.d_seq, seq, 256, cudaMemcpyHostToDevice));
    B200LC_CUDA_TRY(cudaMemset(w.d_small, 0, 260 * sizeof(int)));
    const u32 grid = (n + 255) / 256;
    bzmtf::last_column_kernel<<<grid, 256>>>(w.d_block, w.d_ptr, n, w.d_seq, w.d_ll);
    B200LC_CUDA_TRY(cudaGetLastError());
    rc = b200lc_mtf_batch(w.d_ll, 1, n, w.d_rank, w.d_scratch, w.scratch_bytes, nullptr);
    if (rc) return rc;
    bzmtf::mark_kernel<<<grid, 256>>>(w.d_rank, n, w.d_a);
    rc = prims::inclusive_max_u32(w.d_a, w.d_a, n, w.d_scratch, w.scratch_bytes, nullptr);   // d_a = lastnz
    if (rc) return rc;
    bzmtf::count_kernel<<<grid, 256>>>(w.d_rank, w.d_a, n, w.d_b);
    rc = prims::exclusive_sum_u32(w.d_b, w.d_b, n, w.d_scratch, w.scratch_bytes, nullptr);   // d_b = offsets
    if (rc) return rc;
    bzmtf::emit_kernel<<<grid, 256>>>(w.d_rank, w.d_a, w.d_b, n, eob, w.d_mtfv, w.d_small, w.d_small + 258);
    B200LC_CUDA_TRY(cudaGetLastError());
    int small[260];
    B200LC_CUDA_TRY(cudaMemcpy(small, w.d_small, sizeof(small), cudaMemcpyDeviceToHost));
    const int count = small[258];
    if (count <= 0 || (u32)count > n + 1) return B200LC_ERR_OVERFLOW;
    B200LC_CUDA_TRY(cudaMemcpy(mtfv, w.d_mtfv, (size_t)count * 2, cudaMemcpyDeviceToHost));
    for (u32 k = 0; k <= eob; ++k) mtf_freq[k] = small[k];                      // compress.c:160
    *n_mtf = count;
    return B200LC_OK;
}
