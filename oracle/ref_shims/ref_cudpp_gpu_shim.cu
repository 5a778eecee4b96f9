// oracle/ref_shims/ref_cudpp_gpu_shim.cu -- TEST INFRASTRUCTURE.
//
// Runs the reference's own Huffman kernels of cudppCompress on the GPU (compiled for sm_100a from
// the unmodified kernel header where it lies, oracle/Makefile -> oracle/_ref/libref_cudpp_gpu.so):
//   /root/reference/cudpp-inpar/src/cudpp/kernel/compress_kernel.cuh
//        huffman_build_histogram_kernel (:2037-2121), huffman_build_tree_kernel (:2199-2512),
//        huffman_kernel_en (:2524-2708), huffman_datapack_kernel (:2716-2750)
// with the launch sequence, grid sizes and scratch buffers of huffmanEncoding() and
// allocCompressStorage() (app/compress_app.cu:65-125, 409-439), which cannot be linked here because
// they hang off the CUDPP plan classes.  Input: the MTF bytes of one block on the device; outputs
// = cudppCompress' d_hist, d_encodeOffset, d_compressedSize, d_compressed.  The parity test
// compares them word for word with libb200lc.so (tests/test_cudpp_gpu.py).
#include <cstdio>
#include <cuda_runtime.h>

#include "cudpp_globals.h"
#include "kernel/compress_kernel.cuh"

#define REF_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "ref_cudpp_gpu: %s\n", cudaGetErrorString(e_)); return 1; } } while (0)

extern "C" int ref_cudpp_huffman_gpu(const unsigned char *d_mtf, size_t numElements, unsigned int *d_hist,
                                     unsigned int *d_encodeOffset, unsigned int *d_compressedSize,
                                     unsigned int *d_compressed)
{
    const size_t hist_span = (size_t)HUFF_WORK_PER_THREAD_HIST * HUFF_THREADS_PER_BLOCK_HIST;
    // the reference's own expression (compress_app.cu:76-77), right only for multiples of 32768
    const size_t histBlocks = (numElements % hist_span == 0) ? numElements / hist_span : numElements % hist_span + 1;
    const size_t tThreads = (numElements % HUFF_WORK_PER_THREAD == 0) ? numElements / HUFF_WORK_PER_THREAD
                                                                      : numElements / HUFF_WORK_PER_THREAD + 1;
    const size_t nBlocks = (tThreads % HUFF_THREADS_PER_BLOCK == 0) ? tThreads / HUFF_THREADS_PER_BLOCK
                                                                    : tThreads / HUFF_THREADS_PER_BLOCK + 1;
    const size_t numBitsAlloc = (size_t)HUFF_NUM_CHARS * (HUFF_NUM_CHARS + 1) / 2;
    const size_t numCharsAlloc = (numBitsAlloc % 8 == 0) ? numBitsAlloc / 8 : numBitsAlloc / 8 + 1;
    unsigned char *codes = 0, *lengths = 0;
    unsigned int *locations = 0, *histograms = 0, *nCodesPacked_d = 0;
    encoded *enc = 0;
    REF_TRY(cudaMalloc((void **)&codes, numCharsAlloc));
    REF_TRY(cudaMalloc((void **)&locations, HUFF_NUM_CHARS * sizeof(size_t)));
    REF_TRY(cudaMalloc((void **)&lengths, HUFF_NUM_CHARS));
    REF_TRY(cudaMalloc((void **)&histograms, histBlocks * 256 * sizeof(size_t)));
    REF_TRY(cudaMalloc((void **)&nCodesPacked_d, sizeof(size_t)));
    REF_TRY(cudaMalloc((void **)&enc, sizeof(encoded) * nBlocks));
    REF_TRY(cudaMemset(nCodesPacked_d, 0, sizeof(size_t)));

    huffman_build_histogram_kernel<<<dim3((unsigned)histBlocks), dim3(HUFF_THREADS_PER_BLOCK_HIST)>>>(
        (unsigned int *)d_mtf, histograms, numElements);
    huffman_build_tree_kernel<<<dim3(1), dim3(128)>>>(d_mtf, codes, locations, lengths, histograms, d_hist,
                                                     nCodesPacked_d, d_compressedSize, histBlocks, numElements);
    size_t nCodesPacked = 0;
    REF_TRY(cudaMemcpy(&nCodesPacked, nCodesPacked_d, sizeof(size_t), cudaMemcpyDeviceToHost));
    huffman_kernel_en<<<dim3((unsigned)nBlocks), dim3(HUFF_THREADS_PER_BLOCK), nCodesPacked * sizeof(unsigned char)>>>(
        (uchar4 *)d_mtf, codes, locations, lengths, enc, nCodesPacked, tThreads);
    huffman_datapack_kernel<<<dim3((unsigned)nBlocks), dim3(HUFF_THREADS_PER_BLOCK)>>>(enc, d_compressed, d_compressedSize,
                                                                                     d_encodeOffset);
    REF_TRY(cudaDeviceSynchronize());
    cudaFree(codes); cudaFree(locations); cudaFree(lengths); cudaFree(histograms); cudaFree(nCodesPacked_d); cudaFree(enc);
    return 0;
}
