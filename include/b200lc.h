/*
 * b200lc.h -- C ABI of libb200lc.so, the Blackwell (sm_100a) kernel suite behind the entry points
 * of the compressors vendored in dingwentao/GPU-lossless-compression.
 *
 * Two kinds of symbols are exported:
 *   1. b200lc_*  : explicit-stream, explicit-scratch, device-pointer entry points.  No hidden
 *                  allocation, no host synchronisation unless stated.  These are what the
 *                  reference-named shims below are built on and what bench.py times.
 *   2. the reference's own names (include/cudpp.h, include/culzss_gpu.h, include/cuhd_c.h):
 *                  same names, argument meaning and error behaviour as the reference functions
 *                  they replace, so a reference caller links against libb200lc.so unchanged.
 *
 * All reference citations are relative to /root/reference (dingwentao/GPU-lossless-compression).
 * Every function returns B200LC_OK (0) or a negative B200LC_ERR_* code unless stated otherwise.
 * There is no CPU fallback anywhere: a missing GPU is an error.
 */
#ifndef B200LC_H_
#define B200LC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200LC_OK 0
#define B200LC_ERR_ARG (-1)         /* null / misaligned pointer, bad size */
#define B200LC_ERR_CUDA (-2)        /* a CUDA runtime call failed (message on stderr) */
#define B200LC_ERR_SCRATCH (-3)     /* scratch buffer smaller than *_scratch_bytes() */
#define B200LC_ERR_UNSUPPORTED (-4) /* parameter outside the supported range */
#define B200LC_ERR_OVERFLOW (-5)    /* output capacity exceeded */

/* Version / build identification: "b200lc <git-describe or 'dev'> sm_100a". */
const char *b200lc_version(void);

/* ------------------------------------------------------------------------------------------
 * Hot path 3: canonical-Huffman decode of a CUHD stream.
 * Replaces cuhd::CUHDGPUDecoder::decode (cuhd-icpp/src/cuhd_gpu_decoder.cu:422-523,
 * declared cuhd-icpp/include/cuhd_gpu_decoder.h:24-32).
 *
 *   d_units      compressed stream, uint32 units, codes MSB-first inside each unit
 *                (cuhd-icpp/encoder/src/llhuffman_encoder.cc:200-238); may or may not include
 *                the reference's trailing zero pad unit (cuhd_input_buffer.cc:20-27): units past
 *                n_units read as zero.
 *   n_units      number of units readable at d_units
 *   d_out        n_out decoded symbols (uint8)
 *   d_table      flat LUT of (1 << max_codeword_length) entries {uint8 num_bits, uint8 symbol}
 *                = cuhd::CUHDCodetableItemSingle (cuhd-icpp/include/cuhd_codetable.h:20-23)
 *   max_codeword_length   11 in the reference (cuhd_constants.h:15); 1..13 accepted
 *   d_scratch    >= b200lc_cuhd_decode_scratch_bytes(n_units) bytes, 128-byte aligned; replaces
 *                CUHDGPUDecoderMemory (cuhd_gpu_decoder_memory.cc:25-48; 20 B per 16 B of input)
 *                by 128 B per 4 KiB of input
 *   stream       cudaStream_t (NULL = default stream).  Asynchronous: returns after enqueue.
 * All pointers are device pointers.
 */
size_t b200lc_cuhd_decode_scratch_bytes(size_t n_units);
int b200lc_cuhd_decode(const uint32_t *d_units, size_t n_units, uint8_t *d_out, size_t n_out,
                       const void *d_table, int max_codeword_length, void *d_scratch,
                       size_t scratch_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200LC_H_ */
