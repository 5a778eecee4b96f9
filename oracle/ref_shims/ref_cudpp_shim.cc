// oracle/ref_shims/ref_cudpp_shim.cc -- TEST INFRASTRUCTURE.
//
// extern "C" door onto the reference's own CPU gold code for hot path 1, compiled together with
// the unmodified reference sources where they lie (oracle/Makefile):
//   /root/reference/cudpp-inpar/apps/cudpp_testrig/test_compress.cpp   (included below: computeBwtGold,
//        computeMtfGold, huffman_build_tree_cpu, computeCompressGold = the reference's decoder)
//   /root/reference/cudpp-inpar/apps/cudpp_testrig/sa_gold.cpp         (computeSaGold, CPU DC3)
// Output: oracle/_ref/libref_cudpp.so.
//
// computeCompressGold inverts the BWT with cudppRadixSort on device buffers
// (test_compress.cpp:318-344).  There is no GPU in the build container, so the handful of CUDA
// runtime / cudpp calls it makes are given CPU stand-ins here: host malloc/memcpy and a stable
// key-value sort.  Everything else (tree rebuild, bit walk, inverse MTF, BWT walk) is the
// reference's code.  The test functions in the same file are never called.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

// (StopWatch / command-line helpers come from cudpp_testrig_options.cpp, which defines
// CUDPP_APP_COMMON_IMPL itself)
#include <cuda_runtime_api.h>
#include "cudpp.h"

extern "C" {
cudaError_t cudaMalloc(void **p, size_t n) { *p = std::malloc(n); return cudaSuccess; }
cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void *d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t) { return "cpu stand-in"; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaThreadSynchronize(void) { return cudaSuccess; }
CUDPPResult cudppCreate(CUDPPHandle *h) { *h = 1; return CUDPP_SUCCESS; }
CUDPPResult cudppDestroy(CUDPPHandle) { return CUDPP_SUCCESS; }
CUDPPResult cudppPlan(CUDPPHandle, CUDPPHandle *p, CUDPPConfiguration, size_t, size_t, size_t) { *p = 1; return CUDPP_SUCCESS; }
CUDPPResult cudppDestroyPlan(CUDPPHandle) { return CUDPP_SUCCESS; }
CUDPPResult cudppRadixSort(CUDPPHandle, void *keys, void *values, size_t n)
{
    unsigned char *k = (unsigned char *)keys;
    unsigned int *v = (unsigned int *)values;
    std::vector<unsigned int> idx(n);
    for (size_t i = 0; i < n; i++) idx[i] = (unsigned int)i;
    std::stable_sort(idx.begin(), idx.end(), [&](unsigned a, unsigned b) { return k[a] < k[b]; });
    std::vector<unsigned char> k2(n);
    std::vector<unsigned int> v2(n);
    for (size_t i = 0; i < n; i++) { k2[i] = k[idx[i]]; v2[i] = v[idx[i]]; }
    std::memcpy(k, k2.data(), n);
    std::memcpy(v, v2.data(), n * sizeof(unsigned int));
    return CUDPP_SUCCESS;
}
CUDPPResult cudppCompress(CUDPPHandle, unsigned char *, int *, unsigned int *, unsigned int *, unsigned int *, unsigned int *, unsigned int *, size_t) { abort(); }
CUDPPResult cudppBurrowsWheelerTransform(CUDPPHandle, unsigned char *, unsigned char *, int *, size_t) { abort(); }
CUDPPResult cudppMoveToFrontTransform(CUDPPHandle, unsigned char *, unsigned char *, size_t) { abort(); }
}

// StopWatch is header-implemented behind CUDPP_APP_COMMON_IMPL (stopwatch.h:254); the command-line
// helpers come from cudpp_testrig_options.cpp, which defines the macro itself.
#define CUDPP_APP_COMMON_IMPL
#include "stopwatch.h"
#undef CUDPP_APP_COMMON_IMPL
cudaDeviceProp devProps;   // global of cudpp_testrig.cpp, referenced by the test functions

#include "test_compress.cpp"

extern "C" {

void ref_cudpp_sa(const unsigned char *in, unsigned int *sa /*[n+3]*/, size_t n)
{
    computeSaGold(const_cast<unsigned char *>(in), sa, n);
}

void ref_cudpp_bwt(const unsigned char *in, unsigned char *out, int *index, unsigned int n)
{
    int idx = -1;
    computeBwtGold(const_cast<unsigned char *>(in), out, idx, n);
    *index = idx;
}

void ref_cudpp_mtf(const unsigned char *in, unsigned char *out, unsigned int n)
{
    computeMtfGold(out, in, n);
}

// Tree exactly as the reference builds it from a 256-bin histogram (+EOF): returns the arrays
// left/right/parent/value[513] and the head node (test_compress.cpp:201-236,127-190).
void ref_cudpp_tree(const unsigned int *hist256, int *left, int *right, int *parent, int *value, int *head)
{
    my_huffman_node_t *t = new my_huffman_node_t[NUM_CHARS * 2 - 1];
    unsigned int h[NUM_CHARS];
    for (int j = 0; j < 256; ++j) h[j] = hist256[j];
    h[EOF_CHAR] = 1;
    unsigned int nNodes = 0;
    for (int j = 0; j < NUM_CHARS * 2 - 1; j++) {
        t[j].iter = (unsigned int)j;
        t[j].value = j < NUM_CHARS ? j : 0;
        t[j].ignore = true;
        t[j].count = 0;
        t[j].level = 0;
        t[j].left = t[j].right = t[j].parent = -1;
    }
    for (int j = 0; j < NUM_CHARS; j++)
        if (h[j] > 0) {
            t[nNodes].count = h[j];
            t[nNodes].ignore = 0;
            t[nNodes].value = j;
            nNodes++;
        }
    int hd = -1;
    huffman_build_tree_cpu(t, nNodes, hd);
    for (int j = 0; j < NUM_CHARS * 2 - 1; j++) {
        left[j] = t[j].left; right[j] = t[j].right; parent[j] = t[j].parent; value[j] = t[j].value;
    }
    *head = hd;
    delete[] t;
}

// The reference decoder (n must be 1048576: it walks exactly 256 blocks of 4096 symbols).
void ref_cudpp_decompress(unsigned char *out, int bwtIndex, unsigned int *hist257,
                          unsigned int *encodeOffset, size_t compressedSize,
                          unsigned int *compressed, size_t n)
{
    computeCompressGold(out, bwtIndex, hist257, encodeOffset, compressedSize, compressed, n);
}

}  // extern "C"
