/*
 * culzss_gpu.h -- the CULZSS entry points of libb200lc.so under the reference's own names.
 *
 * Same names, argument order and protocol as cuda-lzss-cluster/gpu_compress.h:121-133 and
 * gpu_decompress.h:109-112, so culzss.c / deculzss.c link against libb200lc.so unchanged
 * (call sites: culzss.c:85-86,108,170,176 and deculzss.c:78-79,98).  The three reference headers
 * disagree on decompression_kernel_wrapper's parameter list (gpu_decompress.h:109,
 * gpu_compress.h:122, culzss.h:85); the 6-argument definition (gpu_decompress.cu:247) is the one
 * exported, as in the reference binary.
 *
 * Differences that a caller cannot observe through this interface:
 *   - compression_kernel_wrapper runs match finding, token selection, packing and the trailer on
 *     the GPU and copies the finished buffer (not 2 bytes of tokens per input byte) into
 *     `bufferout`; aftercompression_wrapper only moves it into `buffer`.  The content of
 *     `bufferout` between the two calls is private to the library.
 *   - errors are reported on stderr and by the return value instead of exit(EXIT_FAILURE).
 */
#ifndef B200LC_CULZSS_GPU_H_
#define B200LC_CULZSS_GPU_H_

#ifdef __cplusplus
extern "C" {
#endif

/* gpu_compress.cu:352-423 */
unsigned char *initGPUmem(int buf_length);
unsigned char *initCPUmem(int buf_length);
void deleteGPUmem(unsigned char *mem_d);
void deleteCPUmem(unsigned char *mem_d);
void initGPU(void);
void resetGPU(void);
int streams_in_GPU(void);
int onestream_finish_GPU(int index);
void deleteGPUStreams(void);

/* gpu_compress.cu:426-460.  buffer: pinned host input (buf_length bytes, multiple of 4096);
 * bufferout: pinned host, >= 2 * buf_length bytes; in_d >= buf_length and out_d >= 2 * buf_length
 * device bytes; asynchronous on the stream group `index` (0..3); returns 1. */
int compression_kernel_wrapper(unsigned char *buffer, int buf_length, unsigned char *bufferout,
                               int compression_type, int wsize, int numthre, int nstreams,
                               int index, unsigned char *in_d, unsigned char *out_d);
/* gpu_compress.cu:569-673.  After onestream_finish_GPU(index): overwrites `buffer` with the
 * compressed buffer (packets + trailer), *comp_length = its size, returns 1; returns 0 when the
 * reference would report "compression took more" (caller keeps the raw buffer). */
int aftercompression_wrapper(unsigned char *buffer, int buf_length, unsigned char *bufferout,
                             int *comp_length);

/* gpu_decompress.cu:95-118, :247-358.  Decodes in place: `buffer` holds buf_length compressed
 * bytes and receives *decomp_length decoded bytes (it must be large enough for them). */
unsigned char *deinitGPUmem(int buf_length);
void dedeleteGPUmem(unsigned char *mem_d);
void deinitGPU(void);
int decompression_kernel_wrapper(unsigned char *buffer, int buf_length, int *decomp_length,
                                 int compression_type, int wsize, int numthre);

#ifdef __cplusplus
}
#endif
#endif /* B200LC_CULZSS_GPU_H_ */
