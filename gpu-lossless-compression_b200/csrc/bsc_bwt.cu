// libbsc's BWT stage: bsc_bwt_encode = divbwt on the GPU.
//
// The reference computes the transform with divsufsort's induced sorting on the CPU
// (cuda-bsc/libbsc/bwt/bwt.cpp:43-52, divsufsort/divsufsort.c:1869-1907; its CUDA code only covers
// the ST5-8 sort transform, st/st2.cu).  Here the block goes through the batched prefix-doubling
// suffix sorter of bwt.cu as one segment; the output contract (U[0] = T[n-1], the row of suffix 0
// dropped, primary index = that row + 1, secondary indexes = rows of the suffixes at multiples of
// the step) is produced by two small kernels from the suffix array.
#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "devprims.cuh"
#include "../../include/b200lc.h"
#include "../../include/libbsc_gpu.h"

namespace b200lc {
namespace bsc {

constexpr int kBadParameter = -1, kGpuError = -7, kGpuNotSupported = -8, kGpuNoMemory = -9;

// last[j] = byte in front of the j-th smallest suffix (T[n-1] for suffix 0, whose row is *row0);
// secondary indexes on the way.
__global__ void __launch_bounds__(256) last_column_kernel(const u8 *__restrict__ in, const u32 *__restrict__ sa,
                                                          u32 n, u32 mod, u8 *__restrict__ last,
                                                          int *__restrict__ row0, int *__restrict__ indexes)
{
    const u32 j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const u32 s = sa[j];
    if (s == 0) {
        last[j] = in[n - 1];
        *row0 = (int)j;
    } else {
        last[j] = in[s - 1];
        if (indexes && (s & mod) == 0) indexes[s / (mod + 1) - 1] = (int)j;
    }
}

// U = [last[row0]] + last without row row0
__global__ void __launch_bounds__(256) drop_row_kernel(const u8 *__restrict__ last, const int *__restrict__ row0,
                                                       u32 n, u8 *__restrict__ U)
{
    const u32 j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const u32 r = (u32)*row0;
    U[j < r ? j + 1 : (j == r ? 0 : j)] = last[j];
}

struct Work {
    u8 *d_in = nullptr, *d_last = nullptr, *d_out = nullptr;
    u32 *d_sa = nullptr;
    int *d_small = nullptr;      // [0] = row of suffix 0, [1..256] = secondary indexes
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0, cap = 0;
    void release()
    {
        cudaFree(d_in); cudaFree(d_last); cudaFree(d_out); cudaFree(d_sa); cudaFree(d_small);
        cudaFree(d_scratch);
        *this = Work();
    }
};
static Work g_work;
static std::mutex g_lock;

static int ensure(size_t n)
{
    if (g_work.cap >= n) return 0;
    g_work.release();
    g_work.scratch_bytes = std::max(b200lc_bwt_scratch_bytes(1, n), b200lc_inverse_bwt_primary_scratch_bytes(n)) + 256;
    if (cudaMalloc(&g_work.d_in, n) != cudaSuccess || cudaMalloc(&g_work.d_last, n) != cudaSuccess ||
        cudaMalloc(&g_work.d_out, n + 16) != cudaSuccess || cudaMalloc(&g_work.d_sa, n * 4) != cudaSuccess ||
        cudaMalloc(&g_work.d_small, 257 * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&g_work.d_scratch, g_work.scratch_bytes) != cudaSuccess) {
        cudaGetLastError();
        g_work.release();
        return kGpuNoMemory;
    }
    g_work.cap = n;
    return 0;
}

static int encode(unsigned char *T, int n, unsigned char *num_indexes, int *indexes)
{
    if (T == nullptr || n < 0) return kBadParameter;
    if (n <= 1) return n;                                  // divsufsort.c:1877 (U aliases T)
    if ((u64)n > prims::kSortMaxElems) return kGpuNotSupported;
    std::lock_guard<std::mutex> guard(g_lock);
    int rc = ensure((size_t)n);
    if (rc) return rc;
    Work &w = g_work;
    u32 mod = (u32)n / 8;                                   // divsufsort.c:1750-1754
    mod |= mod >> 1; mod |= mod >> 2; mod |= mod >> 4; mod |= mod >> 8; mod |= mod >> 16; mod >>= 1;
    const u32 nidx = ((u32)n - 1) / (mod + 1);
    const bool want_idx = num_indexes != nullptr && indexes != nullptr;
    if (cudaMemcpy(w.d_in, T, (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) return kGpuError;
    rc = b200lc_suffix_array_batch(w.d_in, 1, (size_t)n, w.d_sa, w.d_scratch, w.scratch_bytes, nullptr);
    if (rc) return rc == B200LC_ERR_UNSUPPORTED ? kGpuNotSupported : kGpuError;
    const u32 grid = ((u32)n + 255) / 256;
    last_column_kernel<<<grid, 256>>>(w.d_in, w.d_sa, (u32)n, mod, w.d_last, w.d_small,
                                      want_idx ? w.d_small + 1 : nullptr);
    drop_row_kernel<<<grid, 256>>>(w.d_last, w.d_small, (u32)n, w.d_out);
    if (cudaGetLastError() != cudaSuccess) return kGpuError;
    int small[257];
    if (cudaMemcpy(T, w.d_out, (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess) return kGpuError;
    if (cudaMemcpy(small, w.d_small, (1 + (want_idx ? nidx : 0)) * sizeof(int), cudaMemcpyDeviceToHost) !=
        cudaSuccess)
        return kGpuError;
    if (want_idx) {
        *num_indexes = (unsigned char)nidx;
        for (u32 t = 0; t < nidx; ++t) indexes[t] = small[1 + t];
    }
    return small[0] + 1;
}

// bsc_bwt_decode (bwt.cpp:359-397): the reference rebuilds the text with a serial (or, with the
// secondary indexes, 8-way parallel) LF walk on the CPU.  Here the block goes through the inverse
// BWT of the cudppCompress decoder (csrc/cudpp_decode.cu: one 8-bit sort pass = the LF mapping,
// then the walk cut at <= 4096 splitter rows that are walked in parallel, ranked, and walked again
// writing output) with a VIRTUAL end marker -- libbsc's alphabet has none, and without one equal
// bytes do not come in the same order in the first and the last column; blocks of 2^24 rows and
// more use 64-bit row entries.  The secondary indexes are not needed.
static int decode(unsigned char *T, int n, int index)
{
    if (T == nullptr || n < 0 || index <= 0 || index > n) return kBadParameter;     // bwt.cpp:361-364
    if (n <= 1) return 0;
    if ((u64)n > prims::kSortMaxElems) return kGpuNotSupported;
    std::lock_guard<std::mutex> guard(g_lock);
    int rc = ensure((size_t)n);
    if (rc) return rc;
    Work &w = g_work;
    u32 *d_error = reinterpret_cast<u32 *>(w.d_small + 1);
    if (cudaMemcpy(w.d_in, T, (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) return kGpuError;
    // d_out holds n + 1 bytes: the walk also emits the virtual end marker
    rc = b200lc_inverse_bwt_primary(w.d_in, (size_t)n, index, w.d_out, d_error, w.d_scratch, w.scratch_bytes, nullptr);
    if (rc) return rc == B200LC_ERR_UNSUPPORTED ? kGpuNotSupported : kGpuError;
    u32 h_error = 0;
    if (cudaMemcpy(T, w.d_out, (size_t)n, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&h_error, d_error, sizeof(u32), cudaMemcpyDeviceToHost) != cudaSuccess)
        return kGpuError;
    return h_error ? kGpuError : 0;
}

}  // namespace bsc
}  // namespace b200lc

extern "C" int bsc_bwt_decode(unsigned char *T, int n, int index, unsigned char, int *, int)
{
    return b200lc::bsc::decode(T, n, index);
}

extern "C" int bsc_bwt_encode(unsigned char *T, int n, unsigned char *num_indexes, int *indexes, int)
{
    return b200lc::bsc::encode(T, n, num_indexes, indexes);
}

extern "C" void b200lc_bsc_release(void)
{
    std::lock_guard<std::mutex> guard(b200lc::bsc::g_lock);
    b200lc::bsc::g_work.release();
}
