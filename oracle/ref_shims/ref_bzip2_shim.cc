// oracle/ref_shims/ref_bzip2_shim.cc -- TEST INFRASTRUCTURE.
//
// extern "C" door onto the reference's complete cuda-bzip2 library, compiled from the sources
// where they lie (oracle/Makefile) in two flavours:
//   oracle/_ref/libref_bzip2.so       reference CPU stages + the reference's own gpuBWTSort.cu
//   oracle/_ref/libref_bzip2_b200.so  the same reference CPU objects, gpuBlockSort/gpuSetDevice
//                                     resolved from libb200lc.so  (the drop-in link)
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bzlib.h"
#include "bzlib_private.h"

extern "C" {

// Whole-buffer compression through the reference's streaming API with an explicit number of
// additional CPU worker threads (bzlib.h:106-112; 0 = every block goes through gpuBlockSort).
// The reference writes the concatenated block streams with fprintf to strm->handle
// (bzlib.c:506-551), not to next_out, and then calls exit(1) (bzlib.c:606): this function
// DOES NOT RETURN on success -- the caller runs it in a child process and reads `path`
// afterwards (exit() flushes the FILE).
int ref_bzip2_compress_to_file(const char *path, char *source, unsigned int sourceLen,
                               int blockSize100k, int numThreads)
{
    bz_stream strm;
    std::memset(&strm, 0, sizeof(strm));
    FILE *f = std::fopen(path, "wb");
    if (!f) return BZ_IO_ERROR;
    int ret = BZ2_bzCompressInit(&strm, blockSize100k, 0, 30, numThreads);
    if (ret != BZ_OK) { std::fclose(f); return ret; }
    strm.handle = f;
    static char sink[1 << 16];
    strm.next_in = source;
    strm.avail_in = sourceLen;
    do {
        strm.next_out = sink;
        strm.avail_out = sizeof(sink);
        ret = BZ2_bzCompress(&strm, BZ_FINISH);
    } while (ret == BZ_FINISH_OK);
    BZ2_bzCompressEnd(&strm);
    std::fclose(f);
    return ret == BZ_STREAM_END ? BZ_OK : ret;
}

int ref_bzip2_decompress(char *dest, unsigned int *destLen, char *source, unsigned int sourceLen)
{
    return BZ2_bzBuffToBuffDecompress(dest, destLen, source, sourceLen, 0, 0);
}

int ref_bzip2_gpuBlockSort(unsigned char *block, unsigned int *orderFirstSort,
                           unsigned int *orderSecondSort, unsigned int *orderFirstSortRank,
                           int blockSize, int *sortingDepth)
{
    return gpuBlockSort(block, nullptr, orderFirstSort, orderSecondSort, orderFirstSortRank, blockSize,
                        sortingDepth);
}

}  // extern "C"
