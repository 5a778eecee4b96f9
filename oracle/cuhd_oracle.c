/*
 * oracle/cuhd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the CUHD / llhuff algorithms (hot path 3 of SURVEY.md section 8).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (gpu-lossless-compression_b200/csrc) never does.
 *
 * Parity pin: tests/test_oracle_cuhd.py checks every function below against
 *   (1) oracle/_ref/libref_cuhd.so = the reference's own llhuffman_encoder.cc + cuhd_codetable.cc
 *       compiled where they lie under /root/reference (recipe: oracle/Makefile), and
 *   (2) the committed fixtures tests/golden/cuhd_*.bin generated from (1) by
 *       tools/make_golden.py.
 *
 * Reference lines restated (paths relative to /root/reference/):
 *   - decode contract: cuhd-icpp/src/cuhd_gpu_decoder.cu:16-143 (window/next sliding decode,
 *     LUT index = next MAX_CODEWORD_LENGTH bits MSB-first, advance by num_bits)
 *   - bit packer:      cuhd-icpp/encoder/src/llhuffman_encoder.cc:200-238 (encode_memory)
 *   - canonical codes: cuhd-icpp/encoder/src/llhuffman_encoder.cc:160-198 (get_encoder_table)
 *   - flat LUT:        cuhd-icpp/encoder/src/llhuffman_encoder.cc:240-262 (get_decoder_table)
 *   - pad unit:        cuhd-icpp/src/cuhd_input_buffer.cc:13-32
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* LUT entry layout = cuhd::CUHDCodetableItemSingle {uint8 num_bits; uint8 symbol}
 * (cuhd-icpp/include/cuhd_codetable.h:20-23). */
typedef struct { uint8_t num_bits; uint8_t symbol; } cuhd_lut_item;

/* Serial decode: out[0..n_out) = symbols read from bit 0 of unit 0.  Bits past the end of the
 * buffer read as zero (the reference appends one zero unit for the same purpose).
 * Returns number of symbols written (== n_out unless a zero-length LUT entry is hit). */
size_t cuhd_oracle_decode(const uint32_t *units, size_t n_units, const cuhd_lut_item *lut,
                          int max_len, uint8_t *out, size_t n_out)
{
    uint64_t bitpos = 0;
    for (size_t i = 0; i < n_out; ++i) {
        size_t u = (size_t)(bitpos >> 5);
        unsigned at = (unsigned)(bitpos & 31);
        uint64_t w0 = u < n_units ? units[u] : 0;
        uint64_t w1 = u + 1 < n_units ? units[u + 1] : 0;
        uint64_t both = (w0 << 32) | w1;
        uint32_t idx = (uint32_t)((both << at) >> (64 - max_len));
        cuhd_lut_item hit = lut[idx];
        if (hit.num_bits == 0) return i;
        out[i] = hit.symbol;
        bitpos += hit.num_bits;
    }
    return n_out;
}

/* Canonical codes from (symbol, length) pairs already sorted by non-decreasing length
 * (llhuffman_encoder.cc:183-195: code = (code + 1) << (next_len - cur_len)). */
void cuhd_oracle_canonical(const uint8_t *symbols, const uint8_t *lengths, size_t k,
                           uint32_t *code_of_symbol /*[256]*/, uint8_t *len_of_symbol /*[256]*/)
{
    memset(code_of_symbol, 0, 256 * sizeof(uint32_t));
    memset(len_of_symbol, 0, 256);
    uint32_t code = 0;
    for (size_t i = 0; i < k; ++i) {
        code_of_symbol[symbols[i]] = code;
        len_of_symbol[symbols[i]] = lengths[i];
        unsigned next_len = (i + 1 < k) ? lengths[i + 1] : lengths[i];
        code = (code + 1) << (next_len - lengths[i]);
    }
}

/* Flat LUT (llhuffman_encoder.cc:249-259): a code of length L fills 2^(max_len-L) entries. */
void cuhd_oracle_build_lut(const uint32_t *code_of_symbol, const uint8_t *len_of_symbol,
                           int max_len, cuhd_lut_item *lut /*[1<<max_len]*/)
{
    memset(lut, 0, sizeof(cuhd_lut_item) << max_len);
    for (int s = 0; s < 256; ++s) {
        unsigned L = len_of_symbol[s];
        if (!L) continue;
        unsigned shift = (unsigned)max_len - L;
        for (uint32_t j = 0; j < (1u << shift); ++j) {
            lut[(code_of_symbol[s] << shift) + j].num_bits = (uint8_t)L;
            lut[(code_of_symbol[s] << shift) + j].symbol = (uint8_t)s;
        }
    }
}

/* Size in units of the packed stream (llhuffman_encoder.cc:166-180), without the pad unit. */
size_t cuhd_oracle_compressed_units(const uint64_t *hist /*[256]*/, const uint8_t *len_of_symbol)
{
    uint64_t bits = 0;
    for (int s = 0; s < 256; ++s) bits += hist[s] * len_of_symbol[s];
    return (size_t)((bits + 31) / 32);
}

/* MSB-first packer.  Where the reference is well defined this produces the same units as
 * encode_memory (llhuffman_encoder.cc:200-238).  Two documented divergences, both confined to
 * the final unit: the reference leaves the unused low bits of the last
 * unit unspecified (shift by >= width) and never flushes a final partial unit that holds only
 * the tail of a codeword split across the last unit boundary; this restatement zero-fills the
 * unused bits and always flushes.  *defined_units receives the number of leading units that
 * the reference is guaranteed to have written with defined contents. */
size_t cuhd_oracle_encode(const uint8_t *in, size_t n, const uint32_t *code_of_symbol,
                          const uint8_t *len_of_symbol, uint32_t *out, size_t out_units,
                          size_t *defined_units)
{
    uint64_t acc = 0;      /* bits accumulate at the low end */
    unsigned have = 0;     /* number of valid bits in acc */
    size_t o = 0;
    for (size_t i = 0; i < n; ++i) {
        unsigned L = len_of_symbol[in[i]];
        acc = (acc << L) | code_of_symbol[in[i]];
        have += L;
        if (have >= 32) {
            if (o < out_units) out[o] = (uint32_t)(acc >> (have - 32));
            ++o;
            have -= 32;
            acc &= (have ? ((1ull << have) - 1) : 0);
        }
    }
    if (defined_units) *defined_units = o;
    if (have) {
        if (o < out_units) out[o] = (uint32_t)(acc << (32 - have));
        ++o;
    }
    return o;
}
