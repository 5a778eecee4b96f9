/*
 * libbsc_gpu.h -- the BWT stage of libbsc on the B200.
 *
 * libb200lc.so exports bsc_bwt_encode with the reference's name, C linkage, argument meaning and
 * return codes (cuda-bsc/libbsc/bwt/bwt.h:38-48, bwt.cpp:43-52), so the reference's libbsc built
 * without its own definition (see oracle/Makefile, target bsc_b200) runs bsc_compress with the
 * suffix sort on the GPU and everything else (LZP, QLFC coder, container, CLI) unchanged:
 *
 *   T            in: n input bytes; out: U[0] = T[n-1] followed by the last-column bytes of all
 *                rotations except the one of suffix 0 (divsufsort.c:1894-1896).  HOST pointer.
 *   num_indexes  out (may be NULL): number of secondary indexes, (n-1) / step with
 *                step = 2^k derived from n/8 (divsufsort.c:1750-1756)
 *   indexes      out (may be NULL): indexes[t-1] = sorted row of suffix t*step
 *   features     LIBBSC_FEATURE_* bit mask; ignored (the call always runs on the current device)
 *   returns      primary index (row of suffix 0, plus 1) or LIBBSC_BAD_PARAMETER (-1),
 *                LIBBSC_GPU_ERROR (-7), LIBBSC_GPU_NOT_ENOUGH_MEMORY (-9), LIBBSC_GPU_NOT_SUPPORTED
 *                (-8: n >= 2^30)
 * Synchronous.  Thread-safe: concurrent callers (bsc.cpp:206 compresses blocks from OpenMP
 * threads) are serialised on one device work area, like the reference's own CUDA section
 * (cuda-bsc/libbsc/st/st2.cu:72-76).
 */
#ifndef B200LC_LIBBSC_GPU_H_
#define B200LC_LIBBSC_GPU_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

int bsc_bwt_encode(unsigned char *T, int n, unsigned char *num_indexes, int *indexes, int features);

/*
 * The block container of libbsc under the reference's names (cuda-bsc/libbsc/libbsc.h:96-163;
 * behaviour of libbsc/libbsc/libbsc.cpp:61-95,226-352,354-628): same arguments, same 28-byte header
 * {blockSize, dataSize, mode, index, adler32(data), adler32(payload), adler32(header)}, same
 * LIBBSC_* return codes.  HOST pointers; thread-safe across blocks.
 *   bsc_compress     block sort = bsc_bwt_encode on the GPU; blockSorter must be
 *                    LIBBSC_BLOCKSORTER_BWT (1) -- the sort transforms ST3..ST8 are a compile-time
 *                    option of libbsc that is off by default and not built here -> BAD_PARAMETER.
 *                    LZP and the QLFC coder are the stages registered with b200lc_bsc_set_stages();
 *                    without a coder the block is stored (mode 0), as the reference does for a
 *                    block that does not shrink.
 *   bsc_store        stored block (mode 0).
 *   bsc_block_info   validates a header and reports block / data size.
 *   bsc_decompress   stored blocks directly; compressed blocks through the registered
 *                    coder_decompress / bwt_decode / lzp_decompress, LIBBSC_NOT_SUPPORTED (-4)
 *                    when they are missing.
 */
int bsc_init(int features);
int bsc_init_full(int features, void *(*malloc_fn)(size_t size), void *(*zero_malloc_fn)(size_t size),
                  void (*free_fn)(void *address));
int bsc_compress(const unsigned char *input, unsigned char *output, int n, int lzpHashSize, int lzpMinLen,
                 int blockSorter, int coder, int features);
int bsc_store(const unsigned char *input, unsigned char *output, int n, int features);
int bsc_block_info(const unsigned char *blockHeader, int headerSize, int *pBlockSize, int *pDataSize,
                   int features);
int bsc_decompress(const unsigned char *input, int inputSize, unsigned char *output, int outputSize,
                   int features);

/*
 * The CPU stages of libbsc that stay with the host program (signatures of lzp.h:50,62,
 * coder.h:56,66, bwt.h:61).  Any member may be NULL.  The table is copied.
 */
typedef struct b200lc_bsc_stages {
    int (*coder_compress)(const unsigned char *input, unsigned char *output, int n, int coder, int features);
    int (*coder_decompress)(const unsigned char *input, unsigned char *output, int coder, int features);
    int (*lzp_compress)(const unsigned char *input, unsigned char *output, int n, int hashSize, int minLen,
                        int features);
    int (*lzp_decompress)(const unsigned char *input, unsigned char *output, int n, int hashSize, int minLen,
                          int features);
    int (*bwt_decode)(unsigned char *T, int n, int index, unsigned char num_indexes, int *indexes, int features);
} b200lc_bsc_stages;
void b200lc_bsc_set_stages(const b200lc_bsc_stages *stages);

/* Frees the device work area kept between calls. */
void b200lc_bsc_release(void);

#ifdef __cplusplus
}
#endif
#endif /* B200LC_LIBBSC_GPU_H_ */
