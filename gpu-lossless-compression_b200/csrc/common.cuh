// Shared device/host helpers for the sm_100a kernels in this directory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b200lc {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

#define B200LC_NUM_SMS_FALLBACK 148

// Error codes shared by every C-ABI entry point (include/b200lc.h).
enum {
    B200LC_OK = 0,
    B200LC_ERR_ARG = -1,
    B200LC_ERR_CUDA = -2,
    B200LC_ERR_SCRATCH = -3,
    B200LC_ERR_UNSUPPORTED = -4,
    B200LC_ERR_OVERFLOW = -5,
};

#define B200LC_CUDA_TRY(expr)                                                         \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            fprintf(stderr, "b200lc: %s failed at %s:%d: %s\n", #expr, __FILE__,      \
                    __LINE__, cudaGetErrorString(_e));                                \
            return B200LC_ERR_CUDA;                                                   \
        }                                                                             \
    } while (0)

// Kernel attributes, occupancy and the SM count belong to the CURRENT device, and one process may
// drive several devices (cudaSetDevice between calls): every cached value is kept per device.
constexpr int kMaxDevices = 64;
inline int device_slot()        // index into a per-device cache, -1 = do not cache
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
    return dev;
}

// cudaDeviceReset (resetGPU in culzss_api.cu) destroys the context and with it every kernel
// attribute: caches of "attribute already set" remember the epoch they were filled in.
unsigned &context_epoch();      // defined in version.cpp; bumped by resetGPU

inline int num_sms()
{
    static int cached[kMaxDevices] = {0};
    const int slot = device_slot();
    if (slot >= 0 && cached[slot]) return cached[slot];
    int n = 0;
    if (slot < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, slot) != cudaSuccess || n <= 0)
        return B200LC_NUM_SMS_FALLBACK;
    cached[slot] = n;
    return n;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- shared-memory addresses
__device__ __forceinline__ u32 smem_u32(const void *p)
{
    return (u32)__cvta_generic_to_shared(p);
}

// ---------------------------------------------------------------- mbarrier + TMA bulk copy
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, u32 bytes,
                                            u64 *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 1-D TMA bulk copy shared -> global (bulk-group completion).
__device__ __forceinline__ void tma_store_1d(void *dst_gmem, const void *src_smem, u32 bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// Make generic-proxy shared-memory writes visible to the async proxy (TMA).
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- acquire / release globals
__device__ __forceinline__ u32 ld_acquire_u32(const u32 *p)
{
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(u32 *p, u32 v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u64 ld_acquire_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(u64 *p, u64 v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(u64 *p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u32 ld_relaxed_u32(const u32 *p)
{
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(u32 *p, u32 v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ u32 warp_incl_scan(u32 v)
{
    const u32 lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (u32)d) v += t;
    }
    return v;
}
#endif  // __CUDACC__

}  // namespace b200lc
