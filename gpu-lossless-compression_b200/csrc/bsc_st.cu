// libbsc's Sort Transform of order 5..8 on the GPU: bsc_st_encode_cuda with
// the reference's name and contract (cuda-bsc/libbsc/st/st.cuh:64-72, st/st2.cu:113-428; called
// from bsc_st_encode, st/st.cpp:1011-1017, when libbsc is built with LIBBSC_SORT_TRANSFORM_SUPPORT
// and LIBBSC_CUDA_SUPPORT).
//
// ST-k sorts the positions of the CYCLIC text by the k bytes that follow them, equal contexts in
// text order, and outputs the byte in front of each position; the index is the sorted rank of
// position 0.  The reference packs (byte in front, 7 context bytes) into 64-bit keys and sorts the
// context bits with b40c (ST8: key = 8 context bytes, the byte in front as the value).  Here: ONE
// formulation for k = 5..8 on the library's own segmented one-sweep radix sort (devprims.cu) --
// key = the 8 bytes T[i..i+7], MSB first, value = i, k stable 8-bit passes over the top k key
// bytes; then one gather: out[j] = T[value[j] - 1], and the thread that meets value 0 reports j.
#include <mutex>

#include "common.cuh"
#include "devprims.cuh"
#include "../../include/b200lc.h"
#include "../../include/libbsc_gpu.h"

namespace b200lc {
namespace bsc_st {

constexpr int kNoError = 0, kBadParameter = -1, kGpuError = -7, kGpuNotSupported = -8, kGpuNoMemory = -9;

// key[i] = T[i .. i+7] (cyclic), most significant byte first; val[i] = i.  The text is read through
// the read-only path a byte at a time: a 25 MiB block is L2-resident and every byte is used 8 times.
__global__ void __launch_bounds__(256) st_keys_kernel(const u8 *__restrict__ T, u32 n, u64 *__restrict__ keys,
                                                      u32 *__restrict__ vals)
{
    const u32 i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    u64 key = 0;
    u32 at = i;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        key = (key << 8) | __ldg(T + at);
        at = at + 1 == n ? 0 : at + 1;
    }
    keys[i] = key;
    vals[i] = i;
}

__global__ void __launch_bounds__(256) st_gather_kernel(const u8 *__restrict__ T, const u32 *__restrict__ order,
                                                        u32 n, u8 *__restrict__ out, int *__restrict__ index)
{
    const u32 j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const u32 i = order[j];
    out[j] = __ldg(T + (i == 0 ? n - 1 : i - 1));
    if (i == 0) *index = (int)j;
}

struct Work {
    u8 *d_in = nullptr, *d_out = nullptr;
    u64 *keys_a = nullptr, *keys_b = nullptr;
    u32 *vals_a = nullptr, *vals_b = nullptr;
    int *d_index = nullptr;
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0, cap = 0;
    unsigned epoch = 0;
    void release()
    {
        cudaFree(d_in); cudaFree(d_out); cudaFree(keys_a); cudaFree(keys_b); cudaFree(vals_a); cudaFree(vals_b);
        cudaFree(d_index); cudaFree(d_scratch);
        *this = Work();
    }
};
static Work g_work;
static std::mutex g_lock;

static int ensure(size_t n)
{
    if (g_work.epoch != context_epoch()) {      // the context was reset: the old pointers are gone
        g_work = Work();
        g_work.epoch = context_epoch();
    }
    if (g_work.cap >= n) return 0;
    g_work.release();
    g_work.epoch = context_epoch();
    g_work.scratch_bytes = prims::sort_scratch_bytes(n, n) + 256;
    if (cudaMalloc(&g_work.d_in, n) != cudaSuccess || cudaMalloc(&g_work.d_out, n) != cudaSuccess ||
        cudaMalloc(&g_work.keys_a, n * 8) != cudaSuccess || cudaMalloc(&g_work.keys_b, n * 8) != cudaSuccess ||
        cudaMalloc(&g_work.vals_a, n * 4) != cudaSuccess || cudaMalloc(&g_work.vals_b, n * 4) != cudaSuccess ||
        cudaMalloc(&g_work.d_index, sizeof(int)) != cudaSuccess ||
        cudaMalloc(&g_work.d_scratch, g_work.scratch_bytes) != cudaSuccess) {
        cudaGetLastError();
        g_work.release();
        return kGpuNoMemory;
    }
    g_work.cap = n;
    return 0;
}

static int encode(unsigned char *T, int n, int k)
{
    if (T == nullptr || n < 0) return kBadParameter;          // st2.cu:371-373
    if (k < 5 || k > 8) return kBadParameter;
    if (n <= 1) return 0;
    if ((u64)n > prims::kSortMaxElems) return kGpuNotSupported;
    std::lock_guard<std::mutex> guard(g_lock);                 // the reference serialises too (st2.cu:72-76)
    int rc = ensure((size_t)n);
    if (rc) return rc;
    Work &w = g_work;
    cudaStream_t stream = nullptr;
    if (cudaMemcpyAsync(w.d_in, T, (size_t)n, cudaMemcpyHostToDevice, stream) != cudaSuccess) return kGpuError;
    const u32 grid = ((u32)n + 255) / 256;
    st_keys_kernel<<<grid, 256, 0, stream>>>(w.d_in, (u32)n, w.keys_a, w.vals_a);
    if (cudaGetLastError() != cudaSuccess) return kGpuError;
    int in_b = 0;
    rc = prims::sort_pairs<u64>(w.keys_a, w.keys_b, w.vals_a, w.vals_b, (u64)n, (u64)n, 64 - 8 * k, 64, w.d_scratch,
                                w.scratch_bytes, stream, &in_b);
    if (rc) return kGpuError;
    st_gather_kernel<<<grid, 256, 0, stream>>>(w.d_in, in_b ? w.vals_b : w.vals_a, (u32)n, w.d_out, w.d_index);
    if (cudaGetLastError() != cudaSuccess) return kGpuError;
    int index = -1;
    if (cudaMemcpyAsync(T, w.d_out, (size_t)n, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaMemcpyAsync(&index, w.d_index, sizeof(int), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)
        return kGpuError;
    return index >= 0 && index < n ? index : kGpuError;
}

}  // namespace bsc_st
}  // namespace b200lc

extern "C" int bsc_st_cuda_init(int) { return b200lc::bsc_st::kNoError; }

extern "C" int bsc_st_encode_cuda(unsigned char *T, int n, int k, int)
{
    return b200lc::bsc_st::encode(T, n, k);
}

extern "C" void b200lc_bsc_st_release(void)
{
    std::lock_guard<std::mutex> guard(b200lc::bsc_st::g_lock);
    b200lc::bsc_st::g_work.release();
}
