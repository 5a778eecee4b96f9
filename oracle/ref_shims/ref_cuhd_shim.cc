// oracle/ref_shims/ref_cuhd_shim.cc -- TEST INFRASTRUCTURE.
//
// Thin extern "C" door onto the reference's own CPU code for hot path 3, compiled together with
// the unmodified reference sources where they lie (see oracle/Makefile):
//   /root/reference/cuhd-icpp/encoder/src/llhuffman_encoder.cc   (lengths, codes, packer, LUT)
//   /root/reference/cuhd-icpp/src/cuhd_codetable.cc              (LUT container)
// Output: oracle/_ref/libref_cuhd.so.  Nothing from the reference is copied into this repo; this
// file only calls the reference's public functions (llhuffman_encoder.h:32-44).
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#include <llhuff.h>
#include <cuhd_codetable.h>

extern "C" {

// Runs get_symbol_lengths -> get_encoder_table -> get_decoder_table -> encode_memory on `data`.
//   code_of_symbol[256], len_of_symbol[256]: encoder dictionary (len 0 = symbol absent)
//   lut: (1 << MAX_CODEWORD_LENGTH) entries of {u8 num_bits, u8 symbol}
//   units_out: capacity `units_cap` u32; *n_units receives table.compressed_size (no pad unit)
// Buffer is zero-initialised before encode_memory so that units the reference never writes
// (llhuffman_encoder.cc:213 loop exit) are recognisable.
// Returns 0 on success, -1 if the reference refused the input, -2 if units_cap is too small.
int ref_cuhd_encode(const uint8_t *data, size_t n, uint32_t *code_of_symbol,
                    uint8_t *len_of_symbol, uint8_t *lut, uint32_t *units_out, size_t units_cap,
                    size_t *n_units)
{
    auto lengths = llhuff::LLHuffmanEncoder::get_symbol_lengths(
        const_cast<uint8_t *>(data), n);
    if (!lengths) return -1;
    auto enc = llhuff::LLHuffmanEncoder::get_encoder_table(lengths);
    auto dec = llhuff::LLHuffmanEncoder::get_decoder_table(enc);
    std::memset(code_of_symbol, 0, 256 * sizeof(uint32_t));
    std::memset(len_of_symbol, 0, 256);
    for (auto &kv : enc->dict) {
        code_of_symbol[kv.first] = kv.second.codeword;
        len_of_symbol[kv.first] = (uint8_t)kv.second.length;
    }
    std::memcpy(lut, dec->get(), dec->get_size() * sizeof(cuhd::CUHDCodetableItemSingle));
    *n_units = enc->compressed_size;
    if (enc->compressed_size > units_cap) return -2;
    std::memset(units_out, 0, units_cap * sizeof(uint32_t));
    llhuff::LLHuffmanEncoder::encode_memory(units_out, enc->compressed_size,
                                            const_cast<uint8_t *>(data), n, enc);
    return 0;
}

// encode_memory alone with a caller-supplied dictionary (for timing the CPU encode leg).
int ref_cuhd_encode_with_table(const uint8_t *data, size_t n, const uint32_t *code_of_symbol,
                               const uint8_t *len_of_symbol, uint32_t *units_out,
                               size_t n_units)
{
    auto enc = std::make_shared<llhuff::LLHuffmanEncoderTable>();
    enc->compressed_size = n_units;
    for (int s = 0; s < 256; ++s)
        if (len_of_symbol[s])
            enc->dict[(uint8_t)s] = {code_of_symbol[s], (size_t)len_of_symbol[s]};
    llhuff::LLHuffmanEncoder::encode_memory(units_out, n_units, const_cast<uint8_t *>(data), n,
                                            enc);
    return 0;
}

int ref_cuhd_max_codeword_length(void) { return MAX_CODEWORD_LENGTH; }

}  // extern "C"
