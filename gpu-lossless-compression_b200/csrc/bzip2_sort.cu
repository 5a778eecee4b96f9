// cuda-bzip2's GPU block sort (hot path 1, bzip2 flavour) on top of the
// batched suffix sorter.
//
// The reference (cuda-bzip2-ipdpsw/gpuBWTSort.cu:202-484) sorts the mod-3 != 0 rotations with up
// to 17 rounds of 8-byte-key thrust sorts plus comparator sorts with doubling depth, then the
// mod-3 == 0 rotations by (char, rank), and leaves a DC3-style merge to one CPU thread
// (compress.c:609-710).  Cyclic rotation order of B equals the order of the first n suffixes of
// BB, so here one prefix-doubling suffix sort of the doubled block gives the complete rotation
// order; the reference's three output arrays are filtered views of it.
#include "devprims.cuh"

#include "common.cuh"
#include "../../include/b200lc.h"
#include "../../include/bzip2_gpu.h"

namespace b200lc {
namespace bz {

__global__ void classify_kernel(const u32 *__restrict__ sa, u32 n, u32 *__restrict__ flag_first,
                                u32 *__restrict__ flag_second)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2 * n) return;
    const u32 p = sa[j];
    u32 f = 0, s = 0;
    if (p < n) {
        const bool first = (p % 3 != 0) || (n % 3 == 1 && p == n - 1);   // gpuBWTSort.cu:227-256
        f = first;
        s = !first;
    }
    flag_first[j] = f;
    flag_second[j] = s;
}

__global__ void scatter_kernel(const u32 *__restrict__ sa, u32 n, const u32 *__restrict__ flag_first,
                               const u32 *__restrict__ pos_first, const u32 *__restrict__ pos_second,
                               u32 *__restrict__ order_first, u32 *__restrict__ order_second,
                               u32 *__restrict__ rank, u32 *__restrict__ ptr)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2 * n) return;
    const u32 p = sa[j];
    if (p >= n) return;
    const u32 a = pos_first[j], b = pos_second[j];
    ptr[a + b] = p;                      // position among all rotations
    if (flag_first[j]) {
        order_first[a] = p;
        rank[p] = a;
    } else {
        order_second[b] = p;
        rank[p] = 0;
    }
}

// longest common prefix of cyclically adjacent first-sort rotations -> number of characters
// needed to tell every pair apart
__global__ void depth_kernel(const u8 *__restrict__ block, u32 n, const u32 *__restrict__ order_first,
                             u32 f, u32 *__restrict__ max_depth)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k + 1 >= f) return;
    u32 a = order_first[k], b = order_first[k + 1], l = 0;
    while (l < n && block[a] == block[b]) {
        ++l;
        if (++a == n) a = 0;
        if (++b == n) b = 0;
    }
    atomicMax(max_depth, l + 1);
}

struct Work {
    u8 *d_dbl = nullptr;
    u32 *d_sa = nullptr, *d_ff = nullptr, *d_fs = nullptr, *d_pf = nullptr, *d_ps = nullptr;
    u32 *d_of = nullptr, *d_os = nullptr, *d_rank = nullptr, *d_ptr = nullptr, *d_depth = nullptr;
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0, cap = 0;
    void release()
    {
        cudaFree(d_dbl); cudaFree(d_sa); cudaFree(d_ff); cudaFree(d_fs); cudaFree(d_pf); cudaFree(d_ps);
        cudaFree(d_of); cudaFree(d_os); cudaFree(d_rank); cudaFree(d_ptr); cudaFree(d_depth);
        cudaFree(d_scratch);
        *this = Work();
    }
};
static Work g_work;   // called from one thread at a time (OpenMP thread 0, compress.c:898-930)

static int ensure(size_t n)
{
    if (g_work.cap >= n) return B200LC_OK;
    g_work.release();
    const size_t m = 2 * n;
    g_work.scratch_bytes = b200lc_bwt_scratch_bytes(1, m);
    const size_t scan_bytes = prims::scan_scratch_bytes(m);
    if (g_work.scratch_bytes < scan_bytes) g_work.scratch_bytes = scan_bytes;
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_dbl, m));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_sa, m * 4));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_ff, m * 4));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_fs, m * 4));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_pf, m * 4));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_ps, m * 4));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_of, n * 4));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_os, n * 4));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_rank, n * 4));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_ptr, n * 4));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_depth, 256));
    B200LC_CUDA_TRY(cudaMalloc(&g_work.d_scratch, g_work.scratch_bytes + 256));
    g_work.cap = n;
    return B200LC_OK;
}

// Device-side sort of one block; results stay in g_work.  Returns f (first sort length) or < 0.
static int sort_block(const u8 *block, u32 n, bool want_depth, int *depth_out)
{
    if (n == 0 || 2ull * n >= (1u << 21)) return B200LC_ERR_UNSUPPORTED;
    int rc = ensure(n);
    if (rc) return rc;
    Work &w = g_work;
    const u32 m = 2 * n;
    B200LC_CUDA_TRY(cudaMemcpy(w.d_dbl, block, n, cudaMemcpyHostToDevice));
    B200LC_CUDA_TRY(cudaMemcpy(w.d_dbl + n, w.d_dbl, n, cudaMemcpyDeviceToDevice));
    rc = b200lc_suffix_array_batch(w.d_dbl, 1, m, w.d_sa, w.d_scratch, w.scratch_bytes + 256, nullptr);
    if (rc) return rc;
    const u32 grid = (m + 255) / 256;
    classify_kernel<<<grid, 256>>>(w.d_sa, n, w.d_ff, w.d_fs);
    rc = prims::exclusive_sum_u32(w.d_ff, w.d_pf, m, w.d_scratch, w.scratch_bytes, nullptr);
    if (rc) return rc;
    rc = prims::exclusive_sum_u32(w.d_fs, w.d_ps, m, w.d_scratch, w.scratch_bytes, nullptr);
    if (rc) return rc;
    scatter_kernel<<<grid, 256>>>(w.d_sa, n, w.d_ff, w.d_pf, w.d_ps, w.d_of, w.d_os, w.d_rank, w.d_ptr);
    B200LC_CUDA_TRY(cudaGetLastError());
    int f = 2 * (int)((n - 1) / 3) + (int)((n - 1) % 3);     // gpuBWTSort.cu:227
    if (n % 3 == 1) ++f;
    if (want_depth) {
        B200LC_CUDA_TRY(cudaMemset(w.d_depth, 0, 4));
        if (f > 1) depth_kernel<<<(f + 255) / 256, 256>>>(w.d_dbl, n, w.d_of, (u32)f, w.d_depth);
        u32 D = 0;
        B200LC_CUDA_TRY(cudaMemcpy(&D, w.d_depth, 4, cudaMemcpyDeviceToHost));
        // the reference's schedule: 4-byte steps up to depth 64 (+4), then comparator sorts at
        // offset o = 64, 128, 256, ... each covering [o, 3o)  (gpuBWTSort.cu:290-349, 355-418)
        int depth;
        if (D <= 68) depth = D <= 4 ? 0 : (int)((D + 3) / 4) * 4 - 4;
        else {
            depth = 64;
            while ((u64)D > 3ull * (u64)depth && depth < (int)n) depth *= 2;
        }
        *depth_out = depth;
    }
    B200LC_CUDA_TRY(cudaDeviceSynchronize());
    return f;
}

}  // namespace bz
}  // namespace b200lc

using namespace b200lc;

int gpuBlockSort(unsigned char *block, unsigned int *, unsigned int *orderFirstSort,
                 unsigned int *orderSecondSort, unsigned int *orderFirstSortRank, int blockSize,
                 int *sortingDepth)
{
    int depth = 0;
    const int f = bz::sort_block(block, (u32)blockSize, sortingDepth != nullptr, &depth);
    if (f < 0) {
        fprintf(stderr, "b200lc: gpuBlockSort failed (%d)\n", f);
        return f;
    }
    const u32 n = (u32)blockSize;
    cudaMemcpy(orderFirstSort, bz::g_work.d_of, (size_t)f * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(orderSecondSort, bz::g_work.d_os, (size_t)(n - f) * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(orderFirstSortRank, bz::g_work.d_rank, (size_t)n * 4, cudaMemcpyDeviceToHost);
    if (sortingDepth) *sortingDepth = depth;
    return f;
}

void gpuSetDevice(int devId) { cudaSetDevice(devId); }

extern "C" int b200lc_bzip2_block_sort(unsigned char *block, unsigned int *orderFirstSort,
                                       unsigned int *orderSecondSort,
                                       unsigned int *orderFirstSortRank, int blockSize,
                                       int *sortingDepth)
{
    return gpuBlockSort(block, nullptr, orderFirstSort, orderSecondSort, orderFirstSortRank, blockSize,
                        sortingDepth);
}

extern "C" int b200lc_bzip2_rotation_order(const unsigned char *block, int blockSize,
                                           unsigned int *ptr, int *origPtr)
{
    if (!block || !ptr || !origPtr || blockSize <= 0) return B200LC_ERR_ARG;
    int depth = 0;
    const int f = bz::sort_block(block, (u32)blockSize, false, &depth);
    if (f < 0) return f;
    B200LC_CUDA_TRY(cudaMemcpy(ptr, bz::g_work.d_ptr, (size_t)blockSize * 4, cudaMemcpyDeviceToHost));
    *origPtr = -1;
    for (int i = 0; i < blockSize; ++i)
        if (ptr[i] == 0) { *origPtr = i; break; }
    return B200LC_OK;
}
