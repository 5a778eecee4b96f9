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

#ifdef __cplusplus
extern "C" {
#endif

int bsc_bwt_encode(unsigned char *T, int n, unsigned char *num_indexes, int *indexes, int features);

/* Frees the device work area kept between calls. */
void b200lc_bsc_release(void);

#ifdef __cplusplus
}
#endif
#endif /* B200LC_LIBBSC_GPU_H_ */
