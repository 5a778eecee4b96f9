/* oracle/bsc_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of libbsc's BWT stage contract, bsc_bwt_encode = divbwt
 * (cuda-bsc/libbsc/bwt/bwt.cpp:43-52, cuda-bsc/libbsc/bwt/divsufsort/divsufsort.c:1742-1833,
 * 1869-1907), written from the definition instead of the induced-sorting implementation:
 *
 *   SA      = suffix array of T[0..n) with the usual "a proper prefix sorts first" order;
 *   U[0]    = T[n-1];  the other n-1 output bytes are T[SA[j]-1] for j = 0..n-1, skipping the
 *             row of suffix 0 (divsufsort.c:1894-1896);
 *   return  = (row of suffix 0) + 1                                  (divsufsort.c:1897);
 *   step    = mod + 1 with mod = (n/8 rounded up to 2^k - 1) >> 1    (divsufsort.c:1750-1754);
 *   *num_indexes = (n-1) / step                                      (divsufsort.c:1756);
 *   indexes[t-1] = row of suffix t*step, t = 1..num_indexes          (divsufsort.c:1773,1799,1811,1821).
 *
 * Pinned against the reference's own bsc_bwt_encode (oracle/_ref/libref_bsc.so) in
 * tests/test_oracle_bsc.py. */
#include <stdint.h>
#include <stdlib.h>

void cudpp_oracle_sa(const uint8_t *in, uint32_t n, uint32_t *sa);   /* cudpp_oracle.c */

int bsc_oracle_bwt_encode(const uint8_t *T, int n, uint8_t *U, uint8_t *num_indexes, int *indexes)
{
    if (!T || !U || n < 0) return -1;
    if (n <= 1) { if (n == 1) U[0] = T[0]; return n; }
    uint32_t *sa = (uint32_t *)malloc((size_t)n * 4);
    if (!sa) return -2;
    cudpp_oracle_sa(T, (uint32_t)n, sa);
    int mod = n / 8;
    mod |= mod >> 1; mod |= mod >> 2; mod |= mod >> 4; mod |= mod >> 8; mod |= mod >> 16; mod >>= 1;
    if (num_indexes) *num_indexes = (uint8_t)((n - 1) / (mod + 1));
    int pidx = -1, o = 1;
    U[0] = T[n - 1];
    for (int j = 0; j < n; ++j) {
        const uint32_t s = sa[j];
        if (s == 0) { pidx = j; continue; }
        U[o++] = T[s - 1];
        if (num_indexes && indexes && (s & (uint32_t)mod) == 0) indexes[s / (uint32_t)(mod + 1) - 1] = j;
    }
    free(sa);
    return pidx + 1;
}

/* Sort Transform of order k (ST5..ST8), the definition behind libbsc's bsc_st_encode
 * (cuda-bsc/libbsc/st/st.h:64-72) as its CUDA path computes it (st/st2.cu:113-428): position i's
 * context is the k bytes T[i .. i+k-1] of the CYCLIC text; positions are sorted by context, equal
 * contexts keep text order (a stable radix sort of keys laid out in text order, st2.cu:236-250);
 * the output is the byte IN FRONT of each position in that order (T[i-1] cyclic, the top key byte,
 * st2.cu:192); the returned index is the sorted rank of position 0 (the first sorted entry equal to
 * position 0's key, st2.cu:255-262, and position 0 is the first among equal contexts).
 * Pinned against the reference's CPU bsc_st_encode (k = 5, 6) and bsc_st_decode (k = 5..8) from
 * oracle/_ref/libref_bsc.so in tests/test_oracle_bsc.py. */
static const uint8_t *st_text;
static int st_n, st_k;
static int st_cmp(const void *a, const void *b)
{
    const int i = *(const int *)a, j = *(const int *)b;
    for (int d = 0; d < st_k; ++d) {
        const uint8_t x = st_text[(i + d) % st_n], y = st_text[(j + d) % st_n];
        if (x != y) return x < y ? -1 : 1;
    }
    return i < j ? -1 : (i > j ? 1 : 0);
}
int bsc_oracle_st_encode(const uint8_t *T, int n, int k, uint8_t *out)
{
    if (!T || !out || n < 0 || k < 3 || k > 8) return -1;
    if (n <= 1) { if (n == 1) out[0] = T[0]; return 0; }
    int *order = (int *)malloc((size_t)n * sizeof(int));
    if (!order) return -2;
    for (int i = 0; i < n; ++i) order[i] = i;
    st_text = T; st_n = n; st_k = k;
    qsort(order, (size_t)n, sizeof(int), st_cmp);
    int index = -1;
    for (int j = 0; j < n; ++j) {
        out[j] = T[(order[j] + n - 1) % n];
        if (order[j] == 0) index = j;
    }
    free(order);
    return index;
}
