// oracle/ref_shims/ref_bzip2_mtf_shim.cc -- TEST INFRASTRUCTURE.
//
// The reference's generateMTFValues is `static` (cuda-bzip2-ipdpsw/compress.c:122-123), so the
// only way to run it unmodified is to compile compress.c into this translation unit (oracle/Makefile
// passes -I <reference>/cuda-bzip2-ipdpsw; nothing is copied) and call it on a minimal EState.
#include <cstdlib>
#include <cstring>

#include "compress.c"

extern "C" int ref_bzip2_generate_mtf(const unsigned char *block, const unsigned int *ptr, int nblock,
                                      const unsigned char *in_use, unsigned short *mtfv, int *mtf_freq,
                                      int *n_in_use)
{
    EState *s = (EState *)std::calloc(1, sizeof(EState));
    if (!s) return -1;
    s->block = (UChar *)block;
    s->ptr = (UInt32 *)ptr;
    s->mtfv = (UInt16 *)mtfv;
    s->nblock = nblock;
    std::memcpy(s->inUse, in_use, 256);
    generateMTFValues(s);
    std::memcpy(mtf_freq, s->mtfFreq, (size_t)(s->nInUse + 2) * sizeof(int));
    *n_in_use = s->nInUse;
    const int n = s->nMTF;
    std::free(s);
    return n;
}
