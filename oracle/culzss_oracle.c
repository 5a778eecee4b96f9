/*
 * oracle/culzss_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of CULZSS (hot path 2 of SURVEY.md section 8, rows b1-b8).  It follows
 * the reference's own data structures (256-byte ring, 256-byte lookahead, 128 "threads" per
 * chunk) rather than the closed form the product kernels use, so the two are independent.
 *
 * Parity pin (tests/test_oracle_culzss.py, tests/test_culzss_gpu.py):
 *   - aftercomp / trailer / container: against oracle/_ref/libref_culzss.so = the reference's
 *     gpu_compress.cu + gpu_decompress.cu compiled by oracle/Makefile (CPU functions run here);
 *   - token generation and decode: against the reference's EncodeKernel / DecodeKernel from the
 *     same library, executed on the B200 in the -m gpu tests;
 *   - committed fixtures tests/golden/culzss_*.bin.
 *
 * Reference lines restated (paths relative to /root/reference/cuda-lzss-cluster/):
 *   FindMatch            gpu_compress.cu:104-168
 *   EncodeKernel         gpu_compress.cu:182-350
 *   aftercomp            gpu_compress.cu:462-566
 *   aftercompression_wrapper (trailer)  gpu_compress.cu:569-673
 *   DecodeKernel         gpu_decompress.cu:120-244
 *   decompression_kernel_wrapper (trailer parse)  gpu_decompress.cu:247-358
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WINDOW_SIZE 128
#define MAX_UNCODED 2
#define MAX_CODED 128
#define PCKTSIZE 4096
#define RING (WINDOW_SIZE + MAX_CODED)

typedef struct { int offset, length; } match_t;

/* gpu_compress.cu:104-168, one "thread". */
static match_t find_match(int windowHead, int uncodedHead, const uint8_t *ring,
                          const uint8_t *la, int tx, int lastcheck)
{
    match_t m;
    int i = windowHead, j = 0, matching = 0, loop = 0;
    int maxcheck = MAX_CODED - tx * lastcheck;
    m.length = 1;
    m.offset = 1;
    while (loop < WINDOW_SIZE) {
        if (ring[i] == la[(uncodedHead + j) % RING]) {
            j++;
            matching = 1;
        } else {
            if (matching && j > m.length) {
                int t = i - j;
                if (t < 0) t += RING;
                m.length = j;
                m.offset = t;
            }
            j = 0;
            matching = 0;
        }
        i = (i + 1) % RING;
        loop++;
        if (loop >= maxcheck - 1) loop = WINDOW_SIZE;
    }
    if (j > m.length && matching) {
        int t = i - j;
        if (t < 0) t += RING;
        m.length = j;
        m.offset = t;
    }
    return m;
}

/* gpu_compress.cu:251-274 / :319-342: match -> two output bytes. */
static void emit(match_t m, uint8_t literal, uint8_t *out2)
{
    if (m.length >= MAX_CODED) m.length = MAX_CODED - 1;
    if (m.length <= MAX_UNCODED) {
        out2[0] = 1;
        out2[1] = literal;
    } else {
        out2[0] = (uint8_t)m.length;
        out2[1] = (uint8_t)m.offset;
    }
}

/* EncodeKernel for one 4096-byte packet: 8192 bytes of (len, off) / (1, literal) pairs.
 * The 128 threads of a chunk are simulated one after the other between the barriers of
 * gpu_compress.cu:182-350. */
void culzss_oracle_packet_tokens(const uint8_t *pkt, uint8_t *tokens)
{
    uint8_t ring[RING], la[RING];
    match_t md[MAX_CODED];
    int filepoint = 0, wfilepoint = 0, lastcheck = 0;
    int head = 0; /* windowHead - tx == uncodedHead - tx for every thread */

    for (int tx = 0; tx < MAX_CODED; ++tx) ring[tx] = ' ';
    for (int tx = 0; tx < MAX_CODED; ++tx) la[tx] = pkt[tx];
    filepoint += MAX_CODED;
    for (int tx = 0; tx < MAX_CODED; ++tx) ring[(tx + WINDOW_SIZE) % RING] = la[tx];
    for (int tx = 0; tx < MAX_CODED; ++tx) la[MAX_CODED + tx] = pkt[filepoint + tx];
    filepoint += MAX_CODED;
    for (int tx = 0; tx < MAX_CODED; ++tx) md[tx] = find_match(tx, tx, ring, la, tx, 0);

    while (filepoint <= PCKTSIZE && !lastcheck) {
        for (int tx = 0; tx < MAX_CODED; ++tx)
            emit(md[tx], la[(head + tx) % RING], tokens + wfilepoint + 2 * tx);
        wfilepoint += MAX_CODED * 2;
        head = (head + MAX_CODED) % RING;
        if (filepoint < PCKTSIZE) {
            for (int tx = 0; tx < MAX_CODED; ++tx)
                la[(head + tx + MAX_CODED) % RING] = pkt[filepoint + tx];
            filepoint += MAX_CODED;
            for (int tx = 0; tx < MAX_CODED; ++tx)
                ring[(head + tx + WINDOW_SIZE) % RING] = la[(head + tx) % RING];
        } else {
            lastcheck++;
            for (int tx = 0; tx < MAX_CODED; ++tx) ring[(head + tx + MAX_CODED) % RING] = '^';
        }
        for (int tx = 0; tx < MAX_CODED; ++tx)
            md[tx] = find_match((head + tx) % RING, (head + tx) % RING, ring, la, tx, lastcheck);
    }
    for (int tx = 0; tx < MAX_CODED; ++tx) {
        if (lastcheck == 1 && md[tx].length > MAX_CODED - tx) md[tx].length = MAX_CODED - tx;
        emit(md[tx], la[(head + tx) % RING], tokens + wfilepoint + 2 * tx);
    }
}

/* compression_kernel_wrapper's device work for a whole buffer (buf_length % 4096 == 0):
 * tokens[2 * buf_length]. */
void culzss_oracle_buffer_tokens(const uint8_t *buffer, int buf_length, uint8_t *tokens)
{
    for (int p = 0; p < buf_length / PCKTSIZE; ++p)
        culzss_oracle_packet_tokens(buffer + (size_t)p * PCKTSIZE, tokens + (size_t)p * PCKTSIZE * 2);
}

/* aftercomp (NWORKERS = 1) + aftercompression_wrapper: greedy token selection, flag bytes,
 * per-packet sizes, trailer.  `out` needs buf_length + buf_length/8 + 1024 bytes (the reference
 * writes in place over its input buffer and may run a little past buf_length before it notices).
 * Returns 1 and *comp_length, or 0 when the reference reports "compression took more". */
int culzss_oracle_aftercomp(const uint8_t *tokens, int buf_length, uint8_t *out, int *comp_length)
{
    int i = 0, j = 0, k = 0, tempj = 0, holdcount = 0;
    const int finish = buf_length;
    uint8_t flags = 0, flagPos = 1, hold[16];
    int npk = buf_length / PCKTSIZE;
    int *header = (int *)malloc(sizeof(int) * (size_t)(npk > 0 ? npk : 1));

    while (i < finish * 2) {
        if (j > finish) {
            free(header);
            return 0;
        }
        int t = tokens[i];
        if (t == 1) {
            flags |= flagPos;
            hold[holdcount++] = tokens[i + 1];
            i += 2;
        } else {
            hold[holdcount++] = (uint8_t)t;
            hold[holdcount++] = tokens[i + 1];
            i += t * 2;
        }
        if (flagPos == 0x80) {
            out[j++] = flags;
            for (int m = 0; m < holdcount; ++m) out[j++] = hold[m];
            flags = 0;
            flagPos = 1;
            holdcount = 0;
        } else {
            flagPos <<= 1;
        }
        if (i % (PCKTSIZE * 2) == 0 && i > 0) {
            if (holdcount > 0) {
                out[j++] = flags;
                for (int m = 0; m < holdcount; ++m) out[j++] = hold[m];
                holdcount = 0;
            }
            flags = 0;
            flagPos = 1;
            header[k++] = j - tempj;
            tempj = j;
        }
    }
    for (int p = 0; p < npk; ++p) {
        out[j++] = (uint8_t)(header[p] >> 8);
        out[j++] = (uint8_t)header[p];
    }
    out[j++] = (uint8_t)(buf_length >> 24);
    out[j++] = (uint8_t)(buf_length >> 16);
    out[j++] = (uint8_t)(buf_length >> 8);
    out[j++] = (uint8_t)buf_length;
    out[j++] = 0; /* pad size, big endian u16 */
    out[j++] = 0;
    *comp_length = j;
    free(header);
    return 1;
}

/* DecodeKernel for one packet: `in` holds `size` compressed bytes; writes up to cap bytes.
 * Returns the number of bytes produced. */
int culzss_oracle_decode_packet(const uint8_t *in, int size, uint8_t *out, int cap)
{
    uint8_t window[WINDOW_SIZE], tmp[256];
    int nextChar = 0, filepoint = 0, w = 0;
    unsigned flags = 0, flagsUsed = 7;
    memset(window, ' ', sizeof(window));
    for (;;) {
        flags >>= 1;
        flagsUsed++;
        if (flagsUsed == 8) {
            if (filepoint >= size) break;
            flags = in[filepoint++];
            flagsUsed = 0;
        }
        if (flags & 1) {
            if (filepoint >= size) break;
            uint8_t c = in[filepoint++];
            if (w < cap) out[w] = c;
            w++;
            window[nextChar] = c;
            nextChar = (nextChar + 1) % WINDOW_SIZE;
        } else {
            if (filepoint >= size) break;
            int len = in[filepoint++];
            if (filepoint >= size) break;
            int off = in[filepoint++];
            for (int i = 0; i < len; ++i) {
                uint8_t c = window[(off + i) % WINDOW_SIZE];
                if (w < cap) out[w] = c;
                w++;
                tmp[i] = c;
            }
            for (int i = 0; i < len; ++i) window[(nextChar + i) % WINDOW_SIZE] = tmp[i];
            nextChar = (nextChar + len) % WINDOW_SIZE;
        }
    }
    return w;
}

/* decompression_kernel_wrapper: parse the trailer, decode every packet.  Returns 1 and
 * *decomp_length = origsize - padsize, or 0 on a malformed trailer. */
int culzss_oracle_decode_buffer(const uint8_t *buf, int buf_length, uint8_t *out, int out_cap,
                                int *decomp_length)
{
    if (buf_length < 6) return 0;
    int origsize = (buf[buf_length - 6] << 24) ^ (buf[buf_length - 5] << 16) ^
                   (buf[buf_length - 4] << 8) ^ buf[buf_length - 3];
    int padsize = (buf[buf_length - 2] << 8) ^ buf[buf_length - 1];
    if (origsize <= 0 || origsize % PCKTSIZE || origsize > out_cap) return 0;
    int npk = origsize / PCKTSIZE;
    if (buf_length < 6 + 2 * npk) return 0;
    int start = 0;
    for (int p = 0; p < npk; ++p) {
        int sz = (buf[buf_length - 2 * npk + 2 * p - 6] << 8) ^ buf[buf_length - 2 * npk + 2 * p + 1 - 6];
        if (start + sz > buf_length - 2 * npk - 6) return 0;
        culzss_oracle_decode_packet(buf + start, sz, out + (size_t)p * PCKTSIZE, PCKTSIZE);
        start += sz;
    }
    *decomp_length = origsize - padsize;
    return 1;
}
