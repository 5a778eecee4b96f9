// oracle/ref_shims/ref_bzip2_b200mtf_shim.cc -- TEST INFRASTRUCTURE.
//
// Drop-in proof for row N2 (first half): the body a maintainer would give generateMTFValues
// (cuda-bzip2-ipdpsw/compress.c:122-246) to run the MTF + RUNA/RUNB stage on the GPU.
// oracle/Makefile compiles the reference's compress.c with its own definition renamed (sed into a
// temporary file that is deleted after the compile), so the call in BZ2_compressBlock binds here.
#include <cstdio>
#include <cstdlib>

#include "bzlib_private.h"
#include "bzip2_gpu.h"

void generateMTFValues(EState *s)
{
    const int rc = b200lc_bzip2_mtf_rle(s->block, s->ptr, s->nblock, s->inUse, s->mtfv, &s->nMTF,
                                        s->mtfFreq, &s->nInUse);
    if (rc) {
        std::fprintf(stderr, "b200lc_bzip2_mtf_rle failed (%d)\n", rc);
        std::exit(3);
    }
    int k = 0;                                    // makeMaps_e (compress.c:109-118)
    for (int i = 0; i < 256; ++i)
        if (s->inUse[i]) s->unseqToSeq[i] = (UChar)k++;
}
