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

/* ------------------------------------------------------------------------------------------
 * Huffman stage of a bzip2 block: restatement of sendMTFValues
 * (cuda-bzip2-ipdpsw/compress.c:252-606) with BZ2_hbMakeCodeLengths / BZ2_hbAssignCodes
 * (huffman.c:63-153), written as separate steps:
 *   tables      2..6 by nMTF (:274-279); initial tables = consecutive symbol ranges of roughly
 *               equal total frequency, cost 0 inside / 15 outside (:282-319)
 *   4 rounds    every 50-symbol group picks the FIRST cheapest table, the table's frequencies
 *               are re-counted from its groups and its code lengths rebuilt (limit 17) (:324-444)
 *   lengths     a min-heap on weight = freq << 8 | depth merges the two lightest nodes (ties by
 *               heap order); if a length exceeds the limit all weights are halved (+1) and the
 *               build repeats (huffman.c:63-153)
 *   stream      used-symbol map (16 + 16 per non-empty row), 5 bits nGroups and 17 bits
 *               nSelectors (this fork widened both fields, :524-527), selectors move-to-front
 *               coded in unary, code lengths delta coded, then the symbols (:458-603)
 * Bits are written MSB first into `bits` (zeroed by the caller); *nbits = their number. */
typedef struct { uint8_t *buf; uint64_t n; } bz_bits_t;

static void bz_put(bz_bits_t *w, int nb, uint32_t v)
{
    for (int k = nb - 1; k >= 0; --k, ++w->n)
        if ((v >> k) & 1u) w->buf[w->n >> 3] |= (uint8_t)(0x80u >> (w->n & 7));
}

static void bz_sift_up(int *heap, const int *weight, int z)
{
    const int node = heap[z];
    while (weight[node] < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
    heap[z] = node;
}

static void bz_sift_down(int *heap, const int *weight, int count, int z)
{
    const int node = heap[z];
    for (;;) {
        int child = z << 1;
        if (child > count) break;
        if (child < count && weight[heap[child + 1]] < weight[heap[child]]) ++child;
        if (weight[node] < weight[heap[child]]) break;
        heap[z] = heap[child];
        z = child;
    }
    heap[z] = node;
}

void bzip2_oracle_code_lengths(uint8_t *len, const int *freq, int alpha, int limit)
{
    int heap[260], weight[516], parent[516];
    for (int i = 0; i < alpha; ++i) weight[i + 1] = (freq[i] == 0 ? 1 : freq[i]) << 8;
    for (;;) {
        int nodes = alpha, count = 0;
        heap[0] = 0; weight[0] = 0; parent[0] = -2;
        for (int i = 1; i <= alpha; ++i) { parent[i] = -1; heap[++count] = i; bz_sift_up(heap, weight, count); }
        while (count > 1) {
            const int a = heap[1]; heap[1] = heap[count--]; bz_sift_down(heap, weight, count, 1);
            const int b = heap[1]; heap[1] = heap[count--]; bz_sift_down(heap, weight, count, 1);
            ++nodes;
            parent[a] = parent[b] = nodes;
            const int da = weight[a] & 0xff, db = weight[b] & 0xff;
            weight[nodes] = (int)(((uint32_t)weight[a] & 0xffffff00u) + ((uint32_t)weight[b] & 0xffffff00u)) |
                            (1 + (da > db ? da : db));
            parent[nodes] = -1;
            heap[++count] = nodes;
            bz_sift_up(heap, weight, count);
        }
        int too_long = 0;
        for (int i = 1; i <= alpha; ++i) {
            int depth = 0;
            for (int k = i; parent[k] >= 0; k = parent[k]) ++depth;
            len[i - 1] = (uint8_t)depth;
            if (depth > limit) too_long = 1;
        }
        if (!too_long) return;
        for (int i = 1; i <= alpha; ++i) weight[i] = (1 + (weight[i] >> 8) / 2) << 8;
    }
}

int bzip2_oracle_send_mtf(const uint16_t *mtfv, int nMTF, const int *mtfFreq, const uint8_t *in_use,
                          int n_in_use, uint8_t *bits, uint64_t *nbits, uint8_t *len_out /*[6][258]*/,
                          uint8_t *selector_out, int *n_groups_out, int *n_selectors_out)
{
    enum { G = 50, MAXA = 258 };
    const int alpha = n_in_use + 2;
    static uint8_t len[6][MAXA];
    static int code[6][MAXA], rfreq[6][MAXA];
    if (nMTF <= 0 || alpha > MAXA) return -1;
    const int groups = nMTF < 200 ? 2 : nMTF < 600 ? 3 : nMTF < 1200 ? 4 : nMTF < 2400 ? 5 : 6;
    const int nsel = (nMTF + G - 1) / G;
    uint8_t *selector = (uint8_t *)malloc((size_t)nsel);
    for (int t = 0; t < 6; ++t) for (int v = 0; v < alpha; ++v) len[t][v] = 15;

    /* initial tables */
    {
        int part = groups, remaining = nMTF, first = 0;
        while (part > 0) {
            const int target = remaining / part;
            int last = first - 1, acc = 0;
            while (acc < target && last < alpha - 1) acc += mtfFreq[++last];
            if (last > first && part != groups && part != 1 && ((groups - part) % 2 == 1)) acc -= mtfFreq[last--];
            for (int v = 0; v < alpha; ++v) len[part - 1][v] = (v >= first && v <= last) ? 0 : 15;
            --part;
            first = last + 1;
            remaining -= acc;
        }
    }
    /* refinement rounds */
    for (int round = 0; round < 4; ++round) {
        for (int t = 0; t < groups; ++t) for (int v = 0; v < alpha; ++v) rfreq[t][v] = 0;
        for (int g = 0; g < nsel; ++g) {
            const int lo = g * G, hi = lo + G < nMTF ? lo + G : nMTF;
            int best = -1;
            uint32_t best_cost = 999999999u;
            for (int t = 0; t < groups; ++t) {
                uint16_t c = 0;                               /* UInt16 cost[] (:259) */
                for (int i = lo; i < hi; ++i) c = (uint16_t)(c + len[t][mtfv[i]]);
                if (c < best_cost) { best_cost = c; best = t; }
            }
            selector[g] = (uint8_t)best;
            for (int i = lo; i < hi; ++i) rfreq[best][mtfv[i]]++;
        }
        for (int t = 0; t < groups; ++t) bzip2_oracle_code_lengths(len[t], rfreq[t], alpha, 17);
    }
    /* canonical codes per table (huffman.c:134-153) */
    for (int t = 0; t < groups; ++t) {
        int lo = 32, hi = 0, next = 0;
        for (int i = 0; i < alpha; ++i) { if (len[t][i] > hi) hi = len[t][i]; if (len[t][i] < lo) lo = len[t][i]; }
        for (int n = lo; n <= hi; ++n) {
            for (int i = 0; i < alpha; ++i) if (len[t][i] == n) code[t][i] = next++;
            next <<= 1;
        }
    }
    bz_bits_t w = { bits, 0 };
    /* symbol map */
    int row_used[16];
    for (int r = 0; r < 16; ++r) { row_used[r] = 0; for (int c = 0; c < 16; ++c) if (in_use[r * 16 + c]) row_used[r] = 1; }
    for (int r = 0; r < 16; ++r) bz_put(&w, 1, (uint32_t)row_used[r]);
    for (int r = 0; r < 16; ++r) if (row_used[r]) for (int c = 0; c < 16; ++c) bz_put(&w, 1, in_use[r * 16 + c] ? 1u : 0u);
    bz_put(&w, 5, (uint32_t)groups);
    bz_put(&w, 17, (uint32_t)nsel);
    /* selectors: move-to-front rank in unary */
    {
        uint8_t order[6];
        for (int t = 0; t < groups; ++t) order[t] = (uint8_t)t;
        for (int g = 0; g < nsel; ++g) {
            int r = 0;
            while (order[r] != selector[g]) ++r;
            for (int k = r; k > 0; --k) order[k] = order[k - 1];
            order[0] = selector[g];
            for (int k = 0; k < r; ++k) bz_put(&w, 1, 1);
            bz_put(&w, 1, 0);
        }
    }
    /* code lengths, delta coded */
    for (int t = 0; t < groups; ++t) {
        int cur = len[t][0];
        bz_put(&w, 5, (uint32_t)cur);
        for (int i = 0; i < alpha; ++i) {
            while (cur < len[t][i]) { bz_put(&w, 2, 2); ++cur; }
            while (cur > len[t][i]) { bz_put(&w, 2, 3); --cur; }
            bz_put(&w, 1, 0);
        }
    }
    /* symbols */
    for (int i = 0; i < nMTF; ++i) {
        const int t = selector[i / G];
        bz_put(&w, len[t][mtfv[i]], (uint32_t)code[t][mtfv[i]]);
    }
    *nbits = w.n;
    if (len_out) memcpy(len_out, len, sizeof(len));
    if (selector_out) memcpy(selector_out, selector, (size_t)nsel);
    if (n_groups_out) *n_groups_out = groups;
    if (n_selectors_out) *n_selectors_out = nsel;
    free(selector);
    return 0;
}
