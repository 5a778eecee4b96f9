/*
 * bzip2_gpu.h -- the GPU block-sort entry points of cuda-bzip2 served by libb200lc.so.
 *
 * gpuBlockSort / gpuSetDevice keep the reference's C++ linkage and signature
 * (cuda-bzip2-ipdpsw/bzlib_private.h:527-531; the reference Makefile compiles every .c with
 * g++, Makefile:18, so the symbols are C++-mangled): compress.c:726,880,1099,1150 link against
 * libb200lc.so unchanged instead of gpuBWTSort.o.
 *
 * Contract (gpuBWTSort.cu:202-484), all pointers HOST memory:
 *   block[0..blockSize)            RLE1-coded block
 *   orderFirstSort[0..f)           start positions i with i % 3 != 0 (plus i = n-1 when
 *                                  n % 3 == 1) in cyclic-rotation order
 *   orderFirstSortRank[0..n)       rank of position i inside orderFirstSort, 0 for the others
 *   orderSecondSort[0..n-f)        the remaining positions (i % 3 == 0) ordered by
 *                                  (block[i], rank[i+1]) = their rotation order
 *   *sortingDepth                  depth of the reference's refinement schedule at which all
 *                                  first-sort ties were resolved (only used for statistics,
 *                                  compress.c:1036-1037)
 *   return value                   f
 * `order` is not written (the reference does not write it either).
 * The rotation order comes from one suffix sort of the doubled block
 * (b200lc_suffix_array_batch); blocks up to 1,048,575 bytes (bzip2: <= 900,000).
 */
#ifndef B200LC_BZIP2_GPU_H_
#define B200LC_BZIP2_GPU_H_

#include <stddef.h>

#ifdef __cplusplus
int gpuBlockSort(unsigned char *block, unsigned int *order, unsigned int *orderFirstSort,
                 unsigned int *orderSecondSort, unsigned int *orderFirstSortRank, int blockSize,
                 int *sortingDepth);
void gpuSetDevice(int devId);
extern "C" {
#endif

/* Same as gpuBlockSort with C linkage. */
int b200lc_bzip2_block_sort(unsigned char *block, unsigned int *orderFirstSort,
                            unsigned int *orderSecondSort, unsigned int *orderFirstSortRank,
                            int blockSize, int *sortingDepth);
/* Whole rotation order in one array: ptr[0..n) and origPtr (row of rotation 0) -- what
 * merge_two_sort_arrays (compress.c:609-710) computes from the three arrays above.  Host
 * pointers.  Returns 0 or a negative B200LC_ERR_* code. */
int b200lc_bzip2_rotation_order(const unsigned char *block, int blockSize, unsigned int *ptr,
                                int *origPtr);

/* bzip2's MTF + zero-run stage: what generateMTFValues computes
 * (compress.c:122-246).  HOST pointers, synchronous, thread-safe (serialised):
 *   block[nblock], ptr[nblock]   the block and its sorted rotation order (s->block, s->ptr)
 *   in_use[256]                  s->inUse (Bool = unsigned char)
 *   mtfv[nblock + 1]             out: RUNA/RUNB digits, rank + 1 symbols, EOB (s->mtfv; may alias
 *                                ptr like in the reference: ptr is consumed before mtfv is written)
 *   *n_mtf                       out: s->nMTF
 *   mtf_freq[nInUse + 2]         out: s->mtfFreq[0 .. EOB]
 *   *n_in_use                    out (may be NULL): s->nInUse as makeMaps_e computes it
 * Returns 0 or a negative B200LC_ERR_* code. */
int b200lc_bzip2_mtf_rle(const unsigned char *block, const unsigned int *ptr, int nblock,
                         const unsigned char *in_use, unsigned short *mtfv, int *n_mtf,
                         int *mtf_freq, int *n_in_use);

/* bzip2's Huffman stage: the bit string sendMTFValues hands to bsW
 * (compress.c:252-606; huffman.c:63-153): symbol map, 5-bit table count and 17-bit selector count
 * (this fork's field widths, compress.c:524-527), unary move-to-front coded selectors, delta coded
 * code lengths, the symbols.  HOST pointers, synchronous, thread-safe (serialised):
 *   mtfv[n_mtf], mtf_freq[n_in_use + 2], in_use[256], n_in_use    outputs of the MTF + RLE stage
 *   bits[bits_cap]     out: the bit string, MSB first; *n_bits its length in bits
 *   len_out            out (may be NULL): s->len, unsigned char [6][258]
 *   selector_out       out (may be NULL): s->selector[ceil(n_mtf / 50)]
 * Returns 0 or a negative B200LC_ERR_* code (B200LC_ERR_OVERFLOW: bits_cap too small;
 * n_mtf * 17 / 8 + n_mtf / 50 + 8192 bytes always suffice). */
int b200lc_bzip2_send_mtf_values(const unsigned short *mtfv, int n_mtf, const int *mtf_freq,
                                 const unsigned char *in_use, int n_in_use, unsigned char *bits,
                                 size_t bits_cap, unsigned long long *n_bits, unsigned char *len_out,
                                 unsigned char *selector_out);

#ifdef __cplusplus
}
#endif
#endif /* B200LC_BZIP2_GPU_H_ */
