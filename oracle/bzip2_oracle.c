/*
 * oracle/bzip2_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of cuda-bzip2's block-sort contract:
 * the three arrays gpuBlockSort returns (cuda-bzip2-ipdpsw/gpuBWTSort.cu:202-484, contract in
 * SURVEY.md appendix A.4) and the CPU merge that turns them into bzip2's ptr[] / origPtr
 * (compress.c:609-710).
 *
 * Parity pin (tests/test_bzip2_gpu.py, on the GPU box): oracle/_ref/libref_bzip2.so = the
 * reference's complete libbz2 (its own gpuBWTSort.cu for sm_100a + unmodified CPU stages); the
 * arrays of its gpuBlockSort equal the ones computed here, and .bz2 streams produced by the
 * reference library are byte-identical whether it is linked with its own gpuBWTSort.o or with
 * libb200lc.so (oracle/_ref/libref_bzip2_b200.so).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void cudpp_oracle_sa(const uint8_t *in, uint32_t n, uint32_t *sa);

/* Cyclic rotation order of block[0..n): order of the first n suffixes of block+block. */
void bzip2_oracle_rotation_order(const uint8_t *block, uint32_t n, uint32_t *ptr)
{
    uint8_t *dbl = (uint8_t *)malloc((size_t)2 * n);
    uint32_t *sa = (uint32_t *)malloc((size_t)2 * n * 4);
    memcpy(dbl, block, n);
    memcpy(dbl + n, block, n);
    cudpp_oracle_sa(dbl, 2 * n, sa);
    uint32_t k = 0;
    for (uint32_t j = 0; j < 2 * n; ++j) if (sa[j] < n) ptr[k++] = sa[j];
    free(dbl);
    free(sa);
}

/* gpuBlockSort's outputs; returns firstSortLength. */
int bzip2_oracle_block_sort(const uint8_t *block, uint32_t n, uint32_t *orderFirst,
                            uint32_t *orderSecond, uint32_t *rank)
{
    uint32_t *ptr = (uint32_t *)malloc((size_t)n * 4);
    bzip2_oracle_rotation_order(block, n, ptr);
    uint32_t f = 0, s = 0;
    memset(rank, 0, (size_t)n * 4);
    for (uint32_t j = 0; j < n; ++j) {
        uint32_t p = ptr[j];
        int first = (p % 3 != 0) || (n % 3 == 1 && p == n - 1);
        if (first) { rank[p] = f; orderFirst[f++] = p; }
        else orderSecond[s++] = p;
    }
    free(ptr);
    return (int)f;
}

/* compress.c:609-710 restated; returns origPtr. */
int bzip2_oracle_merge(const uint8_t *block, int n, int f, const uint32_t *first,
                       const uint32_t *second, const uint32_t *rank, uint32_t *order)
{
    int origPtr = -1, k = 0, a = 0, b = 0, sl = n - f;
    for (k = 0; k < n && a < f && b < sl; ++k) {
        int i1 = (int)first[a], i2 = (int)second[b];
        int take_first;
        if (block[i1] != block[i2]) {
            take_first = block[i1] < block[i2];
            order[k] = take_first ? first[a++] : second[b++];
            continue;      /* note: the reference does not test for origPtr on this path */
        }
        if (i1 == n - 1) {
            if (block[0] == block[i2 + 1]) {
                if (i2 == n - 2) {
                    if (block[1] == block[0]) take_first = rank[2] < rank[1];
                    else take_first = block[1] < block[0];
                } else take_first = rank[1] < rank[i2 + 2];
            } else take_first = block[0] < block[i2 + 1];
        } else if (i1 % 3 == 1) {
            take_first = rank[i1 + 1] < rank[i2 + 1];
        } else {   /* i1 % 3 == 2 */
            if (block[i1 + 1] == block[i2 + 1]) {
                if (i1 + 2 == n || i2 + 2 == n) {
                    int x = (i1 + 2) % n, y = (i2 + 2) % n, found = 0;
                    take_first = 0;
                    while (x % 3 == 0 || y % 3 == 0) {
                        if (block[x] != block[y]) { take_first = block[x] < block[y]; found = 1; break; }
                        x = (x + 1) % n;
                        y = (y + 1) % n;
                    }
                    if (found) { order[k] = take_first ? first[a++] : second[b++]; continue; }
                    take_first = rank[x] < rank[y];
                } else take_first = rank[i1 + 2] < rank[i2 + 2];
            } else take_first = block[i1 + 1] < block[i2 + 1];
        }
        order[k] = take_first ? first[a++] : second[b++];
        if (order[k] == 0) origPtr = k;
    }
    while (a < f) { order[k] = first[a++]; if (order[k] == 0) origPtr = k; ++k; }
    while (b < sl) { order[k] = second[b++]; if (order[k] == 0) origPtr = k; ++k; }
    return origPtr;
}

/* MTF + zero-run coding of a sorted block: restatement of generateMTFValues
 * (cuda-bzip2-ipdpsw/compress.c:122-246; makeMaps_e :109-118).  The last column of the sorted
 * rotations, mapped to the block's dense alphabet, goes through a move-to-front list; runs of
 * rank 0 are written as their length in bijective base 2 (digits RUNA = 0 / RUNB = 1, least
 * significant first), every other rank r as r + 1, and EOB = nInUse + 1 closes the block.
 * Returns nMTF; freq[0 .. EOB] are the symbol counts; *n_in_use the alphabet size. */
int bzip2_oracle_mtf_rle(const uint8_t *block, const uint32_t *ptr, int n, const uint8_t *in_use,
                         uint16_t *mtfv, int *freq, int *n_in_use)
{
    uint8_t seq[256], list[256];
    int used = 0;
    for (int i = 0; i < 256; ++i) if (in_use[i]) seq[i] = (uint8_t)used++;
    *n_in_use = used;
    const int eob = used + 1;
    for (int i = 0; i <= eob; ++i) freq[i] = 0;
    for (int i = 0; i < used; ++i) list[i] = (uint8_t)i;
    int wr = 0;
    long run = 0;
    for (int i = 0; i <= n; ++i) {
        int r = 0;
        if (i < n) {
            const uint32_t p = ptr[i];
            const uint8_t c = seq[block[p ? p - 1 : (uint32_t)n - 1]];
            while (list[r] != c) ++r;
            if (r == 0) { ++run; continue; }
            for (int k = r; k > 0; --k) list[k] = list[k - 1];
            list[0] = c;
        }
        /* a non-zero rank or the end of the block flushes the pending run */
        if (run > 0) {
            long z = run - 1;
            for (;;) {
                const int sym = (int)(z & 1);
                mtfv[wr++] = (uint16_t)sym;
                freq[sym]++;
                if (z < 2) break;
                z = (z - 2) / 2;
            }
            run = 0;
        }
        if (i < n) { mtfv[wr++] = (uint16_t)(r + 1); freq[r + 1]++; }
    }
    mtfv[wr++] = (uint16_t)eob;
    freq[eob]++;
    return wr;
}
