// The libbsc block container behind the reference's own entry points (libbsc row):
//   bsc_init / bsc_init_full / bsc_compress / bsc_store / bsc_block_info / bsc_decompress
//   (cuda-bsc/libbsc/libbsc.h:96-163; behaviour of cuda-bsc/libbsc/libbsc/libbsc.cpp:61-95,
//   226-352 and 354-628: header layout, mode word, fall back to a stored block, error codes).
//
// What runs where: the block sort of bsc_compress is bsc_bwt_encode on the GPU (bsc_bwt.cu).  The
// stages on either side of it -- LZP, the QLFC entropy coder and the inverse BWT -- are CPU code in
// the reference and are NOT rebuilt here (SURVEY.md 8f row N4 keeps them on the CPU); the host
// program hands them over once with b200lc_bsc_set_stages(), normally the functions of its own
// libbsc.a (lzp.h:50,62, coder.h:56,66, bwt.h:61).  Without a coder every block is stored
// (mode 0), exactly what the reference does for a block that does not shrink
// (libbsc.cpp:316-319); a compressed block met without the decode stages reports
// LIBBSC_NOT_SUPPORTED.  No CPU stand-in is compiled into this library.
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/libbsc_gpu.h"

namespace {

constexpr int kHeader = 28;               // LIBBSC_HEADER_SIZE
constexpr int kNoError = 0, kBadParameter = -1, kNotEnoughMemory = -2, kNotSupported = -4,
              kUnexpectedEob = -5, kDataCorrupt = -6;
constexpr int kSorterBwt = 1, kSorterSt5 = 5, kSorterSt8 = 8, kCoderStatic = 1, kCoderAdaptive = 2;   // libbsc.h:64-73

std::mutex g_mutex;
b200lc_bsc_stages g_stages = {nullptr, nullptr, nullptr, nullptr, nullptr};
void *(*g_malloc)(size_t) = nullptr;
void (*g_free)(void *) = nullptr;

b200lc_bsc_stages stages()
{
    std::lock_guard<std::mutex> lock(g_mutex);
    return g_stages;
}

void *block_alloc(size_t n)
{
    void *(*m)(size_t);
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        m = g_malloc;
    }
    return m ? m(n) : malloc(n);
}

void block_free(void *p)
{
    void (*f)(void *);
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        f = g_free;
    }
    if (f) f(p); else free(p);
}

// Adler-32 as libbsc computes it (adler32.cpp:46-76): sums modulo 65521, low word = 1 + sum of
// bytes, high word = sum of the running low words.
unsigned adler32(const unsigned char *p, size_t n)
{
    unsigned long long a = 1, b = 0;
    while (n) {
        size_t run = n < 4096 ? n : 4096;      // b < 2^16 + 4096 * (2^16 + 4096 * 255): no overflow
        n -= run;
        while (run--) { a += *p++; b += a; }
        a %= 65521u;
        b %= 65521u;
    }
    return (unsigned)(a | (b << 16));
}

void put32(unsigned char *p, unsigned v) { memcpy(p, &v, 4); }
int get32(const unsigned char *p) { int v; memcpy(&v, p, 4); return v; }

// mode word -> fields; returns false when the word is not one bsc_compress can have written
struct Mode { int sorter, coder, lzp_min, lzp_hash; };
bool parse_mode(int mode, Mode &m)
{
    m.sorter = mode & 0x1f;
    m.coder = (mode >> 5) & 7;
    m.lzp_min = (mode >> 8) & 0xff;
    m.lzp_hash = (mode >> 16) & 0xff;
    int again = 0;
    if (m.sorter == kSorterBwt || (m.sorter >= kSorterSt5 && m.sorter <= kSorterSt8)) again = m.sorter;
    else if (m.sorter > 0) return false;      // ST3 / ST4 blocks: CPU-only sort transforms, not built
    if (m.coder == kCoderStatic || m.coder == kCoderAdaptive) again += m.coder << 5;
    else if (m.coder > 0) return false;
    if (m.lzp_min || m.lzp_hash) {
        if (m.lzp_min < 4 || m.lzp_min > 255 || m.lzp_hash < 10 || m.lzp_hash > 28) return false;
        again += (m.lzp_min << 8) + (m.lzp_hash << 16);
    }
    return again == mode;
}

}  // namespace

extern "C" void b200lc_bsc_set_stages(const b200lc_bsc_stages *s)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (s) g_stages = *s;
    else g_stages = b200lc_bsc_stages{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
}

extern "C" int bsc_init_full(int features, void *(*malloc_fn)(size_t), void *(*zero_malloc_fn)(size_t),
                             void (*free_fn)(void *))
{
    (void)features;
    // all three or none, like bsc_platform_init (platform.cpp)
    if ((malloc_fn || zero_malloc_fn || free_fn) && !(malloc_fn && zero_malloc_fn && free_fn)) return kBadParameter;
    std::lock_guard<std::mutex> lock(g_mutex);
    g_malloc = malloc_fn;
    g_free = free_fn;
    return kNoError;
}

extern "C" int bsc_init(int features) { return bsc_init_full(features, nullptr, nullptr, nullptr); }

extern "C" int bsc_store(const unsigned char *input, unsigned char *output, int n, int features)
{
    (void)features;
    if (!input || !output || n < 0) return kBadParameter;
    const unsigned sum = adler32(input, (size_t)n);
    memmove(output + kHeader, input, (size_t)n);
    put32(output + 0, (unsigned)(n + kHeader));
    put32(output + 4, (unsigned)n);
    put32(output + 8, 0);
    put32(output + 12, 0);
    put32(output + 16, sum);
    put32(output + 20, sum);
    put32(output + 24, adler32(output, 24));
    return n + kHeader;
}

extern "C" int bsc_compress(const unsigned char *input, unsigned char *output, int n, int lzpHashSize,
                            int lzpMinLen, int blockSorter, int coder, int features)
{
    if (!input || !output) return kBadParameter;
    // BWT, or the sort transforms that have a GPU path in the reference (ST5..ST8, st.cpp:1011-1017);
    // ST3 / ST4 are CPU-only there and not built here
    if (blockSorter != kSorterBwt && (blockSorter < kSorterSt5 || blockSorter > kSorterSt8)) return kBadParameter;
    if (coder != kCoderStatic && coder != kCoderAdaptive) return kBadParameter;
    int mode = blockSorter + (coder << 5);
    if (lzpMinLen != 0 || lzpHashSize != 0) {
        if (lzpMinLen < 4 || lzpMinLen > 255) return kBadParameter;
        if (lzpHashSize < 10 || lzpHashSize > 28) return kBadParameter;
        mode += (lzpMinLen << 8) + (lzpHashSize << 16);
    }
    if (n < 0 || n > 1073741824) return kBadParameter;
    const b200lc_bsc_stages st = stages();

    // the block is worked on in place in `output`; an aliased input needs its own copy for the
    // checksum and the stored fall-back
    unsigned char *copy = nullptr;
    if (input == output) {
        copy = (unsigned char *)block_alloc((size_t)n + 1);
        if (!copy) return kNotEnoughMemory;
        memcpy(copy, input, (size_t)n);
        input = copy;
    }
    auto finish = [&](int rc) { if (copy) block_free(copy); return rc; };

    if (n <= kHeader || !st.coder_compress) return finish(bsc_store(input, output, n, features));

    int lz_size = n;
    if (mode != (mode & 0xff)) {
        lz_size = st.lzp_compress ? st.lzp_compress(input, output, n, lzpHashSize, lzpMinLen, features) : kNotSupported;
        if (lz_size < kNoError) mode &= 0xff;                  // LZP did not help: plain block
    }
    if (mode == (mode & 0xff)) {
        lz_size = n;
        memcpy(output, input, (size_t)n);
    }

    if (lz_size <= kHeader) {                                  // libbsc.cpp:290-294
        blockSorter = kSorterBwt;
        mode = (mode & ~0x1f) | kSorterBwt;
    }
    int indexes[256];
    unsigned char num_indexes = 0;
    const int index = blockSorter == kSorterBwt
                          ? bsc_bwt_encode(output, lz_size, &num_indexes, indexes, features)          // GPU
                          : bsc_st_encode_cuda(output, lz_size, blockSorter, features);               // GPU
    if (n < 64 * 1024) num_indexes = 0;
    if (index < kNoError) return finish(index);

    unsigned char *buffer = (unsigned char *)block_alloc((size_t)lz_size + 4096);
    if (!buffer) return finish(kNotEnoughMemory);
    int result = st.coder_compress(output, buffer, lz_size, coder, features);
    if (result >= kNoError) memcpy(output + kHeader, buffer, (size_t)result);
    block_free(buffer);
    if (result < kNoError || result + 1 + 4 * (int)num_indexes >= n) return finish(bsc_store(input, output, n, features));

    if (num_indexes) memcpy(output + kHeader + result, indexes, 4 * (size_t)num_indexes);
    output[kHeader + result + 4 * (int)num_indexes] = num_indexes;
    result += 1 + 4 * (int)num_indexes;
    put32(output + 0, (unsigned)(result + kHeader));
    put32(output + 4, (unsigned)n);
    put32(output + 8, (unsigned)mode);
    put32(output + 12, (unsigned)index);
    put32(output + 16, adler32(input, (size_t)n));
    put32(output + 20, adler32(output + kHeader, (size_t)result));
    put32(output + 24, adler32(output, 24));
    return finish(result + kHeader);
}

extern "C" int bsc_block_info(const unsigned char *blockHeader, int headerSize, int *pBlockSize, int *pDataSize,
                              int features)
{
    (void)features;
    if (!blockHeader || headerSize < kHeader) return kUnexpectedEob;
    if ((unsigned)get32(blockHeader + 24) != adler32(blockHeader, 24)) return kDataCorrupt;
    const int block_size = get32(blockHeader + 0), data_size = get32(blockHeader + 4);
    const int mode = get32(blockHeader + 8), index = get32(blockHeader + 12);
    Mode m;
    if (!parse_mode(mode, m)) return kDataCorrupt;
    if (block_size < kHeader || block_size > kHeader + data_size) return kDataCorrupt;
    if (index < 0 || index > data_size) return kDataCorrupt;
    if (pBlockSize) *pBlockSize = block_size;
    if (pDataSize) *pDataSize = data_size;
    return kNoError;
}

extern "C" int bsc_decompress(const unsigned char *input, int inputSize, unsigned char *output, int outputSize,
                              int features)
{
    int block_size = 0, data_size = 0;
    const int info = bsc_block_info(input, inputSize, &block_size, &data_size, features);
    if (info != kNoError) return info;
    if (!output) return kBadParameter;
    if (inputSize < block_size || outputSize < data_size) return kUnexpectedEob;
    if ((unsigned)get32(input + 20) != adler32(input + kHeader, (size_t)(block_size - kHeader))) return kDataCorrupt;
    const int mode = get32(input + 8);
    if (mode == 0) {
        memmove(output, input + kHeader, (size_t)data_size);
        return kNoError;
    }
    const b200lc_bsc_stages st = stages();
    Mode m;
    parse_mode(mode, m);
    // the inverse BWT runs on the GPU unless the host program registered its own; the inverse sort
    // transform exists on the CPU only (st.cpp:1506-1548)
    if (!st.coder_decompress || (m.sorter != kSorterBwt && !st.st_decode)) return kNotSupported;
    if (mode != (mode & 0xff) && !st.lzp_decompress) return kNotSupported;

    // the stages read the block while they write the output: an aliased block is copied first
    unsigned char *copy = nullptr;
    if (input == output) {
        copy = (unsigned char *)block_alloc((size_t)block_size);
        if (!copy) return kNotEnoughMemory;
        memcpy(copy, input, (size_t)block_size);
        input = copy;
    }
    auto finish = [&](int rc) { if (copy) block_free(copy); return rc; };

    const int index = get32(input + 12);
    const unsigned want = (unsigned)get32(input + 16);
    int indexes[256];
    const unsigned char num_indexes = input[block_size - 1];
    if (kHeader + 1 + 4 * (int)num_indexes > block_size) return finish(kDataCorrupt);
    if (num_indexes) memcpy(indexes, input + block_size - 1 - 4 * (int)num_indexes, 4 * (size_t)num_indexes);

    const int lz_size = st.coder_decompress(input + kHeader, output, m.coder, features);
    if (lz_size < kNoError) return finish(lz_size);
    if (lz_size > outputSize) return finish(kDataCorrupt);
    int result = m.sorter != kSorterBwt ? st.st_decode(output, lz_size, m.sorter, index, features)
                 : st.bwt_decode    ? st.bwt_decode(output, lz_size, index, num_indexes, indexes, features)
                                    : bsc_bwt_decode(output, lz_size, index, num_indexes, indexes, features);
    if (result < kNoError) return finish(result);
    int produced = lz_size;
    if (mode != (mode & 0xff)) {
        unsigned char *buffer = (unsigned char *)block_alloc((size_t)lz_size + 1);
        if (!buffer) return finish(kNotEnoughMemory);
        memcpy(buffer, output, (size_t)lz_size);
        produced = st.lzp_decompress(buffer, output, lz_size, m.lzp_hash, m.lzp_min, features);
        block_free(buffer);
        if (produced < kNoError) return finish(produced);
    }
    if (produced != data_size || adler32(output, (size_t)data_size) != want) return finish(kDataCorrupt);
    return finish(kNoError);
}
