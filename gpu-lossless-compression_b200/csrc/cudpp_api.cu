// CUDPP-named boundary (include/cudpp.h) and the batched BWT -> MTF -> Huffman pipeline.
#include <new>

#include "common.cuh"
#include "../../include/b200lc.h"
#include "../../include/cudpp.h"

using namespace b200lc;

// ============================================================================ batched pipeline
extern "C" size_t b200lc_cudpp_compress_scratch_bytes(size_t nblocks, size_t n)
{
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    size_t a = b200lc_bwt_scratch_bytes(nblocks, n);
    size_t b = b200lc_mtf_scratch_bytes(nblocks, n);
    size_t c = b200lc_cudpp_huffman_scratch_bytes(nblocks, n);
    size_t stage = a;                 // the three stages run one after the other
    if (b > stage) stage = b;
    if (c > stage) stage = c;
    return up(stage) + 2 * up(nblocks * n) + 256;
}

// nblocks independent blocks of n bytes at d_in -> per block b: d_bwt_index[b], d_hist[b*256..],
// d_offsets[b*(n/4096)..], d_total_words[b], stream at d_out + b * out_stride_words.
// Synchronises the stream (BWT stage).  *d_error as in b200lc_cudpp_huffman_batch.
extern "C" int b200lc_cudpp_compress_batch(const uint8_t *d_in, size_t nblocks, size_t n,
                                           int *d_bwt_index, uint32_t *d_hist,
                                           uint32_t *d_offsets, uint32_t *d_total_words,
                                           uint32_t *d_out, size_t out_stride_words,
                                           uint32_t *d_error, void *d_scratch,
                                           size_t scratch_bytes, void *stream)
{
    if (nblocks == 0 || n == 0) return B200LC_OK;
    if (!d_scratch || (reinterpret_cast<uintptr_t>(d_scratch) & 255)) return B200LC_ERR_ARG;
    if (scratch_bytes < b200lc_cudpp_compress_scratch_bytes(nblocks, n)) return B200LC_ERR_SCRATCH;
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    char *s = reinterpret_cast<char *>(d_scratch);
    const size_t stage_bytes = scratch_bytes - 2 * up(nblocks * n) - 256;
    uint8_t *bwt = reinterpret_cast<uint8_t *>(s + up(stage_bytes));
    uint8_t *mtf = bwt + up(nblocks * n);
    int rc = b200lc_bwt_batch(d_in, nblocks, n, bwt, d_bwt_index, s, stage_bytes, stream);
    if (rc) return rc;
    rc = b200lc_mtf_batch(bwt, nblocks, n, mtf, s, stage_bytes, stream);
    if (rc) return rc;
    return b200lc_cudpp_huffman_batch(mtf, nblocks, n, d_hist, d_offsets, d_total_words, d_out,
                                      out_stride_words, d_error, s, stage_bytes, stream);
}

// ============================================================================ CUDPP objects
namespace {
struct Plan {
    unsigned magic;
    CUDPPConfiguration config;
    size_t n;
    void *scratch;
    size_t scratch_bytes;
    unsigned int *d_total;     // 1 word
    unsigned int *d_error;     // 1 word
};
struct Manager { unsigned magic; };
constexpr unsigned kPlanMagic = 0xB200C0DEu, kMgrMagic = 0xB2000001u;

Plan *plan_of(CUDPPHandle h)
{
    if (h == 0 || h == CUDPP_INVALID_HANDLE) return nullptr;
    Plan *p = reinterpret_cast<Plan *>(h);
    return p->magic == kPlanMagic ? p : nullptr;
}
CUDPPResult check_device()
{
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
        return CUDPP_ERROR_UNKNOWN;
    return major < 2 ? CUDPP_ERROR_ILLEGAL_CONFIGURATION : CUDPP_SUCCESS;   // cudpp.cpp:776-787
}
}  // namespace

extern "C" CUDPPResult cudppCreate(CUDPPHandle *theCudpp)
{
    if (!theCudpp) return CUDPP_ERROR_INVALID_HANDLE;
    Manager *m = new (std::nothrow) Manager();
    if (!m) return CUDPP_ERROR_INSUFFICIENT_RESOURCES;
    m->magic = kMgrMagic;
    *theCudpp = reinterpret_cast<CUDPPHandle>(m);
    return CUDPP_SUCCESS;
}

extern "C" CUDPPResult cudppDestroy(CUDPPHandle theCudpp)
{
    Manager *m = reinterpret_cast<Manager *>(theCudpp);
    if (!theCudpp || theCudpp == CUDPP_INVALID_HANDLE || m->magic != kMgrMagic)
        return CUDPP_ERROR_INVALID_HANDLE;
    m->magic = 0;
    delete m;
    return CUDPP_SUCCESS;
}

extern "C" CUDPPResult cudppPlan(const CUDPPHandle cudppHandle, CUDPPHandle *planHandle,
                                 CUDPPConfiguration config, size_t n, size_t, size_t)
{
    if (!planHandle) return CUDPP_ERROR_INVALID_HANDLE;
    *planHandle = CUDPP_INVALID_HANDLE;
    Manager *m = reinterpret_cast<Manager *>(cudppHandle);
    if (!cudppHandle || cudppHandle == CUDPP_INVALID_HANDLE || m->magic != kMgrMagic)
        return CUDPP_ERROR_INVALID_HANDLE;
    if (config.algorithm != CUDPP_COMPRESS && config.algorithm != CUDPP_BWT &&
        config.algorithm != CUDPP_MTF && config.algorithm != CUDPP_SA && config.algorithm != CUDPP_SORT_RADIX)
        return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    if (n == 0) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    if (config.algorithm == CUDPP_SORT_RADIX) {
        // the sort the reference's own test decodes its BWT with (test_compress.cpp:318-344):
        // unsigned char or unsigned int keys, with or without unsigned int values
        if (config.datatype != CUDPP_UCHAR && config.datatype != CUDPP_UINT) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
        if (!(config.options & (CUDPP_OPTION_KEYS_ONLY | CUDPP_OPTION_KEY_VALUE_PAIRS)) || n >= (size_t(1) << 30))
            return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    } else if (config.datatype != CUDPP_UCHAR) {
        return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    }
    Plan *p = new (std::nothrow) Plan();
    if (!p) return CUDPP_ERROR_INSUFFICIENT_RESOURCES;
    p->magic = kPlanMagic;
    p->config = config;
    p->n = n;
    switch (config.algorithm) {
        case CUDPP_COMPRESS: p->scratch_bytes = b200lc_cudpp_compress_scratch_bytes(1, n); break;
        case CUDPP_MTF: p->scratch_bytes = b200lc_mtf_scratch_bytes(1, n); break;
        case CUDPP_SORT_RADIX:   // sort scratch + keys A/B + values A/B
            p->scratch_bytes = ((b200lc_sort_scratch_bytes(n, 0) + 255) & ~size_t(255)) + 4 * ((n * 4 + 255) & ~size_t(255));
            break;
        default: p->scratch_bytes = b200lc_bwt_scratch_bytes(1, n); break;
    }
    if (cudaMalloc(&p->scratch, p->scratch_bytes + 512) != cudaSuccess) {
        delete p;
        return CUDPP_ERROR_INSUFFICIENT_RESOURCES;
    }
    p->d_total = reinterpret_cast<unsigned int *>(reinterpret_cast<char *>(p->scratch) + p->scratch_bytes);
    p->d_error = p->d_total + 64;
    *planHandle = reinterpret_cast<CUDPPHandle>(p);
    return CUDPP_SUCCESS;
}

extern "C" CUDPPResult cudppDestroyPlan(CUDPPHandle plan)
{
    Plan *p = plan_of(plan);
    if (!p) return CUDPP_ERROR_INVALID_HANDLE;
    cudaFree(p->scratch);
    p->magic = 0;
    delete p;
    return CUDPP_SUCCESS;
}

extern "C" CUDPPResult cudppCompress(CUDPPHandle planHandle, unsigned char *d_uncompressed,
                                     int *d_bwtIndex, unsigned int *, unsigned int *d_hist,
                                     unsigned int *d_encodeOffset, unsigned int *d_compressedSize,
                                     unsigned int *d_compressed, size_t numElements)
{
    CUDPPResult dev = check_device();
    if (dev != CUDPP_SUCCESS) return dev;
    Plan *p = plan_of(planHandle);
    if (!p) return CUDPP_ERROR_INVALID_HANDLE;
    if (p->config.algorithm != CUDPP_COMPRESS) return CUDPP_ERROR_INVALID_PLAN;
    if (p->config.datatype != CUDPP_UCHAR) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    if (numElements == 0 || numElements > p->n) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    const size_t nhb = (numElements + 4095) / 4096;
    const size_t cap_words = nhb * (1536 + 1);   // what the reference test allocates (test_compress.cpp:713-718)
    int rc = b200lc_cudpp_compress_batch(d_uncompressed, 1, numElements, d_bwtIndex, d_hist,
                                         d_encodeOffset, d_compressedSize, d_compressed, cap_words,
                                         p->d_error, p->scratch, p->scratch_bytes, nullptr);
    if (rc != B200LC_OK) return rc == B200LC_ERR_CUDA ? CUDPP_ERROR_UNKNOWN : CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    unsigned int err = 0;
    if (cudaMemcpy(&err, p->d_error, sizeof(err), cudaMemcpyDeviceToHost) != cudaSuccess)
        return CUDPP_ERROR_UNKNOWN;
    // a code longer than 32 bits or a block beyond 1536 words: the reference overruns its
    // buffers here; report instead
    return err ? CUDPP_ERROR_INSUFFICIENT_RESOURCES : CUDPP_SUCCESS;
}

extern "C" CUDPPResult cudppBurrowsWheelerTransform(CUDPPHandle planHandle, unsigned char *d_in,
                                                    unsigned char *d_out, int *d_index,
                                                    size_t numElements)
{
    CUDPPResult dev = check_device();
    if (dev != CUDPP_SUCCESS) return dev;
    Plan *p = plan_of(planHandle);
    if (!p) return CUDPP_ERROR_INVALID_HANDLE;
    if (p->config.algorithm != CUDPP_BWT) return CUDPP_ERROR_INVALID_PLAN;
    if (p->config.datatype != CUDPP_UCHAR) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    if (numElements == 0 || numElements > p->n) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    int rc = b200lc_bwt_batch(d_in, 1, numElements, d_out, d_index, p->scratch, p->scratch_bytes, nullptr);
    return rc == B200LC_OK ? CUDPP_SUCCESS : CUDPP_ERROR_UNKNOWN;
}

extern "C" CUDPPResult cudppMoveToFrontTransform(CUDPPHandle planHandle, unsigned char *d_in,
                                                 unsigned char *d_out, size_t numElements)
{
    CUDPPResult dev = check_device();
    if (dev != CUDPP_SUCCESS) return dev;
    Plan *p = plan_of(planHandle);
    if (!p) return CUDPP_ERROR_INVALID_HANDLE;
    if (p->config.algorithm != CUDPP_MTF) return CUDPP_ERROR_INVALID_PLAN;
    if (p->config.datatype != CUDPP_UCHAR) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    if (numElements > p->n) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    int rc = b200lc_mtf_batch(d_in, 1, numElements, d_out, p->scratch, p->scratch_bytes, nullptr);
    return rc == B200LC_OK ? CUDPP_SUCCESS : CUDPP_ERROR_UNKNOWN;
}

extern "C" CUDPPResult cudppSuffixArray(CUDPPHandle planHandle, unsigned char *d_str,
                                        unsigned int *d_keys_sa, size_t numElements)
{
    CUDPPResult dev = check_device();
    if (dev != CUDPP_SUCCESS) return dev;
    Plan *p = plan_of(planHandle);
    if (!p) return CUDPP_ERROR_INVALID_HANDLE;
    if (p->config.algorithm != CUDPP_SA) return CUDPP_ERROR_INVALID_PLAN;
    if (p->config.datatype != CUDPP_UCHAR) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    if (numElements == 0 || numElements > p->n) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    int rc = b200lc_suffix_array_batch(d_str, 1, numElements, d_keys_sa, p->scratch, p->scratch_bytes, nullptr);
    return rc == B200LC_OK ? CUDPP_SUCCESS : CUDPP_ERROR_UNKNOWN;
}

// ============================================================================ cudppRadixSort
namespace {
__global__ void widen_keys_kernel(const unsigned char *__restrict__ in, u32 *__restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
__global__ void narrow_keys_kernel(const u32 *__restrict__ in, unsigned char *__restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned char)in[i];
}
}  // namespace

// Stable ascending radix sort of numElements keys (and their values) in place, on the default
// stream (cudpp-inpar/include/cudpp.h:256-259; dispatch cudpp.cpp `cudppRadixSort`).  The keys
// travel through the one-sweep sort of devprims.cu as 32-bit words (8 or 32 key bits sorted).
extern "C" CUDPPResult cudppRadixSort(const CUDPPHandle planHandle, void *d_keys, void *d_values,
                                      size_t numElements)
{
    CUDPPResult dev = check_device();
    if (dev != CUDPP_SUCCESS) return dev;
    Plan *p = plan_of(planHandle);
    if (!p) return CUDPP_ERROR_INVALID_HANDLE;
    if (p->config.algorithm != CUDPP_SORT_RADIX) return CUDPP_ERROR_INVALID_PLAN;
    const bool pairs = (p->config.options & CUDPP_OPTION_KEY_VALUE_PAIRS) != 0;
    if (numElements > p->n || !d_keys || (pairs && !d_values)) return CUDPP_ERROR_ILLEGAL_CONFIGURATION;
    if (numElements == 0) return CUDPP_SUCCESS;
    const size_t n = numElements;
    const size_t sort_bytes = (b200lc_sort_scratch_bytes(p->n, 0) + 255) & ~size_t(255);
    const size_t arr = (p->n * 4 + 255) & ~size_t(255);
    char *base = reinterpret_cast<char *>(p->scratch);
    u32 *ka = reinterpret_cast<u32 *>(base + sort_bytes), *kb = reinterpret_cast<u32 *>(base + sort_bytes + arr);
    u32 *va = reinterpret_cast<u32 *>(base + sort_bytes + 2 * arr), *vb = reinterpret_cast<u32 *>(base + sort_bytes + 3 * arr);
    const bool bytes = p->config.datatype == CUDPP_UCHAR;
    const u32 grid = (u32)((n + 255) / 256);
    if (bytes) widen_keys_kernel<<<grid, 256>>>(static_cast<const unsigned char *>(d_keys), ka, n);
    else if (cudaMemcpyAsync(ka, d_keys, n * 4, cudaMemcpyDeviceToDevice, 0) != cudaSuccess) return CUDPP_ERROR_UNKNOWN;
    if (pairs) {
        if (cudaMemcpyAsync(va, d_values, n * 4, cudaMemcpyDeviceToDevice, 0) != cudaSuccess) return CUDPP_ERROR_UNKNOWN;
    } else if (cudaMemsetAsync(va, 0, n * 4, 0) != cudaSuccess) {
        return CUDPP_ERROR_UNKNOWN;
    }
    int in_b = 0;
    const int rc = b200lc_sort_pairs_u32(ka, kb, va, vb, n, 0, 0, bytes ? 8 : 32, base, sort_bytes, nullptr, &in_b);
    if (rc != B200LC_OK) return CUDPP_ERROR_UNKNOWN;
    const u32 *ks = in_b ? kb : ka, *vs = in_b ? vb : va;
    if (bytes) narrow_keys_kernel<<<grid, 256>>>(ks, static_cast<unsigned char *>(d_keys), n);
    else if (cudaMemcpyAsync(d_keys, ks, n * 4, cudaMemcpyDeviceToDevice, 0) != cudaSuccess) return CUDPP_ERROR_UNKNOWN;
    if (pairs && cudaMemcpyAsync(d_values, vs, n * 4, cudaMemcpyDeviceToDevice, 0) != cudaSuccess) return CUDPP_ERROR_UNKNOWN;
    return cudaGetLastError() == cudaSuccess ? CUDPP_SUCCESS : CUDPP_ERROR_UNKNOWN;
}
