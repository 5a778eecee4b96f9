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
 * The inverse, with the reference's name and contract (bwt.h:50-61, bwt.cpp:359-397): T holds what
 * bsc_bwt_encode wrote, `index` its return value; on return T is the original block.  The secondary
 * indexes (the reference's 8-way CPU parallelism) are accepted and not needed: the walk is cut at
 * up to 4096 splitter rows on the GPU.  Returns LIBBSC_NO_ERROR (0), LIBBSC_BAD_PARAMETER (-1: index
 * outside 1..n), LIBBSC_GPU_ERROR (-7), LIBBSC_GPU_NOT_SUPPORTED (-8), LIBBSC_GPU_NOT_ENOUGH_MEMORY (-9).
 * HOST pointer, synchronous, serialised on the same work area as bsc_bwt_encode.
 */
int bsc_bwt_decode(unsigned char *T, int n, int index, unsigned char num_indexes, int *indexes, int features);

/*
 * libbsc's Sort Transform of order k = 5..8 on the GPU under the reference's names
 * (cuda-bsc/libbsc/st/st.cuh:56-72; st/st2.cu:367-428): what bsc_st_encode calls when libbsc is built
 * with LIBBSC_SORT_TRANSFORM_SUPPORT and LIBBSC_CUDA_SUPPORT (st/st.cpp:1011-1017), so such a build
 * links against libb200lc.so instead of its own st2.cu + b40c.
 *   T        in: n input bytes; out: for the positions of the cyclic text sorted by the k bytes that
 *            follow them (equal contexts in text order) the byte in front of each.  HOST pointer.
 *   returns  the sorted rank of position 0 (the index bsc_st_decode needs), or LIBBSC_BAD_PARAMETER
 *            (-1: k outside 5..8), LIBBSC_GPU_ERROR (-7), LIBBSC_GPU_NOT_SUPPORTED (-8: n >= 2^30),
 *            LIBBSC_GPU_NOT_ENOUGH_MEMORY (-9); 0 for n <= 1.
 * For k = 5, 6 the bytes and the index equal the reference's CPU bsc_st_encode; k = 7, 8 exist on the
 * GPU only in the reference too (st.cpp:1016,1026) and are checked through its CPU bsc_st_decode.
 * Synchronous, serialised on one device work area (b200lc_bsc_st_release frees it).
 */
int bsc_st_cuda_init(int features);
int bsc_st_encode_cuda(unsigned char *T, int n, int k, int features);
void b200lc_bsc_st_release(void);

/*
 * The block container of libbsc under the reference's names (cuda-bsc/libbsc/libbsc.h:96-163;
 * behaviour of libbsc/libbsc/libbsc.cpp:61-95,226-352,354-628): same arguments, same 28-byte header
 * {blockSize, dataSize, mode, index, adler32(data), adler32(payload), adler32(header)}, same
 * LIBBSC_* return codes.  HOST pointers; thread-safe across blocks.
 *   bsc_compress     block sort on the GPU: blockSorter = LIBBSC_BLOCKSORTER_BWT (1) -> bsc_bwt_encode,
 *                    LIBBSC_BLOCKSORTER_ST5..ST8 (5..8) -> bsc_st_encode_cuda (the sort transforms
 *                    that have a GPU path in the reference; a compile-time option of libbsc that
 *                    its default build leaves off); ST3 / ST4 (CPU-only there) -> BAD_PARAMETER.
 *                    LZP and the QLFC coder are the stages registered with b200lc_bsc_set_stages();
 *                    without a coder the block is stored (mode 0), as the reference does for a
 *                    block that does not shrink.
 *   bsc_store        stored block (mode 0).
 *   bsc_block_info   validates a header and reports block / data size.
 *   bsc_decompress   stored blocks directly; compressed blocks through the registered
 *                    coder_decompress / bwt_decode (ST blocks: st_decode) / lzp_decompress,
 *                    LIBBSC_NOT_SUPPORTED (-4) when they are missing.
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
 * coder.h:56,66, bwt.h:61, st.h:83).  Any member may be NULL.  The table is copied.
 */
typedef struct b200lc_bsc_stages {
    int (*coder_compress)(const unsigned char *input, unsigned char *output, int n, int coder, int features);
    int (*coder_decompress)(const unsigned char *input, unsigned char *output, int coder, int features);
    int (*lzp_compress)(const unsigned char *input, unsigned char *output, int n, int hashSize, int minLen,
                        int features);
    int (*lzp_decompress)(const unsigned char *input, unsigned char *output, int n, int hashSize, int minLen,
                          int features);
    int (*bwt_decode)(unsigned char *T, int n, int index, unsigned char num_indexes, int *indexes, int features);
    /* st.h:83, needed to decompress ST5..ST8 blocks only (added in round 2, last member: a table
     * initialised with five members leaves it NULL) */
    int (*st_decode)(unsigned char *T, int n, int k, int index, int features);
} b200lc_bsc_stages;
void b200lc_bsc_set_stages(const b200lc_bsc_stages *stages);

/* Frees the device work area kept between calls. */
void b200lc_bsc_release(void);

#ifdef __cplusplus
}
#endif
#endif /* B200LC_LIBBSC_GPU_H_ */
