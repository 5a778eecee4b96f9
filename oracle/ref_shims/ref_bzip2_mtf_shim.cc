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

// The reference's static sendMTFValues (compress.c:252-606) on a minimal EState: the bit stream it
// writes with bsW goes to `bits`; *nbits counts the bits before the final flush.
extern "C" int ref_bzip2_send_mtf(const unsigned short *mtfv, int nMTF, const int *mtf_freq,
                                  const unsigned char *in_use, int n_in_use, unsigned char *bits,
                                  unsigned long long *nbits, unsigned char *len_out, unsigned char *selector_out)
{
    EState *s = (EState *)std::calloc(1, sizeof(EState));
    if (!s) return -1;
    s->mtfv = (UInt16 *)mtfv;
    s->nMTF = nMTF;
    s->nInUse = n_in_use;
    std::memcpy(s->inUse, in_use, 256);
    std::memcpy(s->mtfFreq, mtf_freq, (size_t)(n_in_use + 2) * sizeof(int));
    s->zbits = bits;
    s->numZ = 0;
    BZ2_bsInitWrite(s);
    sendMTFValues(s);
    *nbits = (unsigned long long)s->numZ * 8ull + (unsigned long long)s->bsLive;
    bsFinishWrite(s);
    std::memcpy(len_out, s->len, sizeof(s->len));
    std::memcpy(selector_out, s->selector, (size_t)((nMTF + BZ_G_SIZE - 1) / BZ_G_SIZE));
    std::free(s);
    return 0;
}
