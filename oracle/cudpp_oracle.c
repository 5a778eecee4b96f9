/*
 * oracle/cudpp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of cudppCompress = BWT -> MTF -> Huffman (hot path 1 of SURVEY.md
 * section 8, rows c1-c8; compress_app.cu:507-526).
 *
 * Parity pin (tests/test_oracle_cudpp.py):
 *   - BWT, MTF, tree shape: oracle/_ref/libref_cudpp.so = the reference testrig's gold code
 *     (computeSaGold / computeBwtGold / computeMtfGold / huffman_build_tree_cpu);
 *   - the compressed words: the reference's own decoder computeCompressGold from the same
 *     library must reproduce the input from this oracle's stream (the reference tests pin the
 *     stream only through that round trip, test_compress.cpp:744-797);
 *   - committed fixtures tests/golden/cudpp_*.npz.
 *
 * Reference lines restated (paths relative to /root/reference/cudpp-inpar/):
 *   BWT definition       src/cudpp/kernel/sa_kernel.cuh:47-60, kernel/compress_kernel.cuh:55-74,
 *                        apps/cudpp_testrig/test_compress.cpp:79-91
 *   MTF                  apps/cudpp_testrig/test_compress.cpp:93-125
 *   tree build           src/cudpp/kernel/compress_kernel.cuh:2306-2392, cta/compress_cta.cuh:550-571
 *   code walk            src/cudpp/kernel/compress_kernel.cuh:2403-2496
 *   block encode / pack  src/cudpp/kernel/compress_kernel.cuh:2524-2750
 *   decoder              apps/cudpp_testrig/test_compress.cpp:192-364
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NUM_CHARS 257
#define EOF_CHAR 256
#define COMPOSITE (-1)
#define BLOCK_CHARS 4096
#define BLOCK_WORDS_MAX 1536 /* HUFF_CODE_BYTES, cudpp_globals.h:65-66 */

/* ------------------------------------------------------------------------------------ BWT */
/* Suffix array of in[0..n) with an implicit terminator smaller than every byte (the reference
 * sorts T[i] = in[i] + 1 followed by zeros).  Prefix doubling with radix passes. */
static void radix_pass(const uint32_t *src, uint32_t *dst, const uint32_t *key, uint32_t n,
                       uint32_t K, uint32_t *cnt)
{
    memset(cnt, 0, (size_t)(K + 1) * sizeof(uint32_t));
    for (uint32_t i = 0; i < n; ++i) cnt[key[src[i]]]++;
    uint32_t sum = 0;
    for (uint32_t k = 0; k <= K; ++k) { uint32_t t = cnt[k]; cnt[k] = sum; sum += t; }
    for (uint32_t i = 0; i < n; ++i) dst[cnt[key[src[i]]]++] = src[i];
}

void cudpp_oracle_sa(const uint8_t *in, uint32_t n, uint32_t *sa)
{
    if (n == 0) return;
    uint32_t *rank = (uint32_t *)malloc((size_t)n * 4), *key2 = (uint32_t *)malloc((size_t)n * 4);
    uint32_t *tmp = (uint32_t *)malloc((size_t)n * 4), *nr = (uint32_t *)malloc((size_t)n * 4);
    uint32_t K = n > 256 ? n : 256;
    uint32_t *cnt = (uint32_t *)malloc((size_t)(K + 2) * 4);
    for (uint32_t i = 0; i < n; ++i) { rank[i] = in[i] + 1u; sa[i] = i; }
    radix_pass(sa, tmp, rank, n, 256, cnt);
    memcpy(sa, tmp, (size_t)n * 4);
    /* ranks after 1 char: 1 + index of group head */
    nr[sa[0]] = 1;
    for (uint32_t j = 1; j < n; ++j) nr[sa[j]] = in[sa[j]] == in[sa[j - 1]] ? nr[sa[j - 1]] : j + 1;
    memcpy(rank, nr, (size_t)n * 4);
    for (uint32_t h = 1;; h <<= 1) {
        int unique = 1;
        for (uint32_t j = 1; j < n && unique; ++j) if (rank[sa[j]] == rank[sa[j - 1]]) unique = 0;
        if (unique || h >= n) break;
        for (uint32_t i = 0; i < n; ++i) key2[i] = i + h < n ? rank[i + h] : 0;
        for (uint32_t i = 0; i < n; ++i) tmp[i] = i;
        radix_pass(tmp, sa, key2, n, K, cnt);      /* by second key */
        radix_pass(sa, tmp, rank, n, K, cnt);      /* stable by first key */
        memcpy(sa, tmp, (size_t)n * 4);
        nr[sa[0]] = 1;
        for (uint32_t j = 1; j < n; ++j) {
            uint32_t a = sa[j], b = sa[j - 1];
            nr[a] = (rank[a] == rank[b] && key2[a] == key2[b]) ? nr[b] : j + 1;
        }
        memcpy(rank, nr, (size_t)n * 4);
    }
    free(rank); free(key2); free(tmp); free(nr); free(cnt);
}

/* bwt[i] = in[SA[i] - 1], or in[n-1] where SA[i] == 0, whose row is *index
 * (compress_kernel.cuh:66-72). */
void cudpp_oracle_bwt(const uint8_t *in, uint32_t n, uint8_t *out, int *index)
{
    uint32_t *sa = (uint32_t *)malloc((size_t)(n ? n : 1) * 4);
    cudpp_oracle_sa(in, n, sa);
    for (uint32_t i = 0; i < n; ++i) {
        if (sa[i] == 0) { *index = (int)i; out[i] = in[n - 1]; }
        else out[i] = in[sa[i] - 1];
    }
    free(sa);
}

/* ------------------------------------------------------------------------------------ MTF */
void cudpp_oracle_mtf(const uint8_t *in, uint32_t n, uint8_t *out)
{
    uint8_t list[256];
    for (int i = 0; i < 256; ++i) list[i] = (uint8_t)i;
    for (uint32_t i = 0; i < n; ++i) {
        unsigned j = 0;
        while (list[j] != in[i]) ++j;
        out[i] = (uint8_t)j;
        for (; j > 0; --j) list[j] = list[j - 1];
        list[0] = in[i];
    }
}

/* -------------------------------------------------------------------------------- Huffman */
typedef struct {
    int value;
    unsigned count;
    int ignore, level;
    int left, right, parent;
} node_t;

static int find_min(const node_t *t, int elements)
{
    int cur = -1, curLevel = INT_MAX;
    unsigned curCount = 0;
    int have = 0;
    for (int i = 0; i < elements; ++i) {
        if (t[i].ignore) continue;
        /* the reference starts from currentCount = INT_MAX and compares count < (unsigned)INT_MAX */
        if (!have) {
            if (t[i].count < (unsigned)INT_MAX ||
                (t[i].count == (unsigned)INT_MAX && t[i].level < INT_MAX)) {
                cur = i; curCount = t[i].count; curLevel = t[i].level; have = 1;
            }
        } else if (t[i].count < curCount || (t[i].count == curCount && t[i].level < curLevel)) {
            cur = i; curCount = t[i].count; curLevel = t[i].level;
        }
    }
    return cur;
}

/* Tree from hist[0..255] (+ EOF with count 1).  Arrays of 513 entries; returns head. */
int cudpp_oracle_tree(const uint32_t *hist, int *left, int *right, int *parent, int *value)
{
    node_t t[NUM_CHARS * 2 - 1];
    int n = 0;
    for (int j = 0; j < NUM_CHARS * 2 - 1; ++j) {
        t[j].value = j < NUM_CHARS ? j : 0;
        t[j].count = 0; t[j].ignore = 1; t[j].level = 0;
        t[j].left = t[j].right = t[j].parent = -1;
    }
    for (int j = 0; j < NUM_CHARS; ++j) {
        unsigned c = j == EOF_CHAR ? 1u : hist[j];
        if (c > 0) { t[n].count = c; t[n].ignore = 0; t[n].value = j; ++n; }
    }
    int min1 = -1, min2 = -1;
    for (;;) {
        min1 = find_min(t, n);
        if (min1 < 0) break;
        t[min1].ignore = 1;
        min2 = find_min(t, n);
        if (min2 < 0) break;
        t[min1].ignore = 0;
        int placed = 0;
        for (int i = n; i < NUM_CHARS * 2 - 1; ++i) {
            if (t[i].count == 0) {
                t[i] = t[min1];
                t[i].ignore = 1;
                t[i].parent = min1;
                if (t[i].left >= 0) t[t[i].left].parent = i;
                if (t[i].right >= 0) t[t[i].right].parent = i;
                t[min1].left = i;
                placed = 1;
                break;
            }
        }
        if (!placed) break;
        t[min2].ignore = 1;
        t[min1].value = COMPOSITE;
        t[min1].ignore = 0;
        t[min1].count += t[min2].count;
        t[min1].level = (t[min1].level > t[min2].level ? t[min1].level : t[min2].level) + 1;
        t[min1].right = min2;
        t[min2].parent = min1;
        t[min1].parent = -1;
    }
    for (int j = 0; j < NUM_CHARS * 2 - 1; ++j) {
        left[j] = t[j].left; right[j] = t[j].right; parent[j] = t[j].parent; value[j] = t[j].value;
    }
    return min1;
}

/* Codes by the walk of compress_kernel.cuh:2416-2496: left edge 0, right edge 1.
 * code[s] right-aligned in 64 bits, len[s] = depth; len 0 = symbol absent. */
void cudpp_oracle_codes(const int *left, const int *right, const int *parent, const int *value,
                        int head, uint64_t *code, uint8_t *len)
{
    memset(code, 0, NUM_CHARS * sizeof(uint64_t));
    memset(len, 0, NUM_CHARS);
    int cur = head, depth = 0;
    uint64_t path = 0;
    for (;;) {
        while (left[cur] != -1) { path <<= 1; cur = left[cur]; ++depth; }
        if (value[cur] != COMPOSITE) { code[value[cur]] = path; len[value[cur]] = (uint8_t)depth; }
        while (parent[cur] != -1) {
            if (cur != right[parent[cur]]) { path |= 1; cur = right[parent[cur]]; break; }
            --depth; path >>= 1; cur = parent[cur];
        }
        if (parent[cur] == -1) break;
    }
}

/* Full Huffman stage on MTF output `mtf` (n symbols): hist[256], per-4096-symbol block stream
 * [nWords][words...], offsets[b], total words.  Returns 0, or -1 if a block exceeds the
 * reference's 1536-word capacity (where the reference would overrun its buffer). */
int cudpp_oracle_huffman(const uint8_t *mtf, uint32_t n, uint32_t *hist, uint32_t *offsets,
                         uint32_t *total_words, uint32_t *out, uint32_t out_cap)
{
    int left[513], right[513], parent[513], value[513];
    uint64_t code[NUM_CHARS];
    uint8_t len[NUM_CHARS];
    memset(hist, 0, 256 * sizeof(uint32_t));
    for (uint32_t i = 0; i < n; ++i) hist[mtf[i]]++;
    int head = cudpp_oracle_tree(hist, left, right, parent, value);
    cudpp_oracle_codes(left, right, parent, value, head, code, len);
    uint32_t nblocks = (n + BLOCK_CHARS - 1) / BLOCK_CHARS, w = 0;
    for (uint32_t b = 0; b < nblocks; ++b) {
        uint32_t lo = b * BLOCK_CHARS, hi = lo + BLOCK_CHARS < n ? lo + BLOCK_CHARS : n;
        uint64_t bits = 0;
        for (uint32_t i = lo; i < hi; ++i) bits += len[mtf[i]];
        uint32_t nw = (uint32_t)((bits + 31) / 32);
        if (nw > BLOCK_WORDS_MAX) return -1;
        if (w + 1 + nw > out_cap) return -2;
        offsets[b] = w;
        out[w] = nw;
        memset(out + w + 1, 0, (size_t)nw * 4);
        uint64_t pos = 0;
        for (uint32_t i = lo; i < hi; ++i) {
            unsigned L = len[mtf[i]];
            uint64_t c = code[mtf[i]];
            for (unsigned k = 0; k < L; ++k, ++pos)
                if ((c >> (L - 1 - k)) & 1) out[w + 1 + (pos >> 5)] |= 0x80000000u >> (pos & 31);
        }
        w += 1 + nw;
    }
    *total_words = w;
    return 0;
}

/* cudppCompress end to end. */
int cudpp_oracle_compress(const uint8_t *in, uint32_t n, int *bwt_index, uint32_t *hist,
                          uint32_t *offsets, uint32_t *total_words, uint32_t *out, uint32_t out_cap)
{
    uint8_t *a = (uint8_t *)malloc(n ? n : 1), *b = (uint8_t *)malloc(n ? n : 1);
    cudpp_oracle_bwt(in, n, a, bwt_index);
    cudpp_oracle_mtf(a, n, b);
    int rc = cudpp_oracle_huffman(b, n, hist, offsets, total_words, out, out_cap);
    free(a); free(b);
    return rc;
}

/* Decoder for any n (generalises computeCompressGold, test_compress.cpp:192-364, which is
 * hard-wired to 256 blocks): tree from the histogram, bit walk per block, inverse MTF, inverse
 * BWT through a stable counting sort of (byte, index). */
int cudpp_oracle_decompress(uint8_t *out, uint32_t n, int bwt_index, const uint32_t *hist,
                            const uint32_t *offsets, const uint32_t *comp)
{
    int left[513], right[513], parent[513], value[513];
    int head = cudpp_oracle_tree(hist, left, right, parent, value);
    uint8_t *mtf = (uint8_t *)malloc(n ? n : 1), *bwt = (uint8_t *)malloc(n ? n : 1);
    uint32_t nblocks = (n + BLOCK_CHARS - 1) / BLOCK_CHARS;
    for (uint32_t b = 0; b < nblocks; ++b) {
        uint32_t want = b * BLOCK_CHARS + BLOCK_CHARS <= n ? BLOCK_CHARS : n - b * BLOCK_CHARS;
        const uint32_t *w = comp + offsets[b] + 1;
        uint64_t pos = 0;
        int cur = head;
        uint32_t found = 0;
        while (found < want) {
            int bit = (w[pos >> 5] >> (31 - (pos & 31))) & 1;
            ++pos;
            cur = bit ? right[cur] : left[cur];
            if (cur < 0) { free(mtf); free(bwt); return -1; }
            if (value[cur] != COMPOSITE) {
                if (value[cur] == EOF_CHAR) break;
                mtf[b * BLOCK_CHARS + found++] = (uint8_t)value[cur];
                cur = head;
            }
        }
    }
    uint8_t list[256];
    for (int i = 0; i < 256; ++i) list[i] = (uint8_t)i;
    for (uint32_t i = 0; i < n; ++i) {
        uint8_t r = mtf[i], c = list[r];
        bwt[i] = c;
        for (unsigned j = r; j > 0; --j) list[j] = list[j - 1];
        list[0] = c;
    }
    uint32_t cnt[257] = {0};
    uint32_t *next = (uint32_t *)malloc((size_t)(n ? n : 1) * 4);
    for (uint32_t i = 0; i < n; ++i) cnt[bwt[i] + 1]++;
    for (int c = 0; c < 256; ++c) cnt[c + 1] += cnt[c];
    for (uint32_t i = 0; i < n; ++i) next[cnt[bwt[i]]++] = i;
    uint32_t idx = (uint32_t)bwt_index;
    for (uint32_t i = 0; i < n; ++i) { idx = next[idx]; out[i] = bwt[idx]; }
    free(next); free(mtf); free(bwt);
    return 0;
}
