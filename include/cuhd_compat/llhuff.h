/*
 * llhuff.h (b200lc) -- source-compatible mirror of cuhd-icpp's length-limited Huffman encoder
 * interface (encoder/include/llhuffman_encoder.h:20-44, llhuffman_encoder_table.h:17-32), so the
 * reference's demo.cc compiles unchanged.  The four static functions keep their signatures:
 *   get_symbol_lengths  -> exact-integer package-merge (b200lc_cuhd_build_table) instead of the
 *                          float one (llhuffman_encoder.cc:53-128)
 *   get_encoder_table   -> canonical codes over (length, symbol) order (:160-198)
 *   encode_memory       -> the GPU packer (b200lc_cuhd_encode) instead of the serial CPU loop
 *                          (:200-238); the final partial unit is flushed and zero-filled
 *   get_decoder_table   -> the flat LUT (:240-262)
 */
#ifndef B200LC_LLHUFF_COMPAT_H_
#define B200LC_LLHUFF_COMPAT_H_

#include <algorithm>
#include <memory>
#include <unordered_map>
#include <vector>

#include "cuhd.h"

namespace llhuff {

struct LLHuffmanEncoderTable {
    struct LLHuffmanEncoderTableItem {
        UNIT_TYPE codeword;
        size_t length;
    };
    size_t compressed_size;   // units, without the pad unit
    std::unordered_map<SYMBOL_TYPE, LLHuffmanEncoderTableItem> dict;
};

class LLHuffmanEncoder {
   public:
    struct Symbol {
        size_t length;
        size_t count;
        SYMBOL_TYPE symbol;
    };

    static std::shared_ptr<std::vector<Symbol>> get_symbol_lengths(SYMBOL_TYPE* data, size_t size) {
        std::uint64_t hist[256] = {0};
        for (size_t i = 0; i < size; ++i) ++hist[data[i]];
        std::uint32_t code[256];
        std::uint8_t len[256];
        if (b200lc_cuhd_build_table(hist, MAX_CODEWORD_LENGTH, code, len, nullptr) != B200LC_OK)
            return nullptr;
        auto out = std::make_shared<std::vector<Symbol>>();
        for (int s = 0; s < 256; ++s)
            if (len[s]) out->push_back(Symbol{len[s], static_cast<size_t>(hist[s]), static_cast<SYMBOL_TYPE>(s)});
        std::stable_sort(out->begin(), out->end(),
                         [](const Symbol& a, const Symbol& b) { return a.length < b.length; });
        return out;
    }

    static std::shared_ptr<LLHuffmanEncoderTable> get_encoder_table(
        std::shared_ptr<std::vector<Symbol>> symbol_lengths) {
        auto table = std::make_shared<LLHuffmanEncoderTable>();
        size_t bits = 0;
        for (auto& s : *symbol_lengths) bits += s.length * s.count;
        table->compressed_size = (bits + 8 * sizeof(UNIT_TYPE) - 1) / (8 * sizeof(UNIT_TYPE));
        UNIT_TYPE code = 0;
        for (size_t i = 0; i < symbol_lengths->size(); ++i) {
            const size_t cur = (*symbol_lengths)[i].length;
            table->dict[(*symbol_lengths)[i].symbol] = {code, cur};
            const size_t next = i + 1 < symbol_lengths->size() ? (*symbol_lengths)[i + 1].length : cur;
            code = (code + 1) << (next - cur);
        }
        return table;
    }

    static void encode_memory(UNIT_TYPE* out, size_t size_out, SYMBOL_TYPE* in, size_t size_in,
                              std::shared_ptr<LLHuffmanEncoderTable> encoder_table) {
        std::uint32_t code[256] = {0};
        std::uint8_t len[256] = {0};
        for (auto& kv : encoder_table->dict) {
            code[kv.first] = kv.second.codeword;
            len[kv.first] = static_cast<std::uint8_t>(kv.second.length);
        }
        std::uint8_t *d_in = nullptr, *d_len = nullptr, *d_scratch = nullptr;
        std::uint32_t *d_code = nullptr, *d_units = nullptr;
        std::uint64_t* d_bits = nullptr;
        const size_t cap = size_out + 2;
        const size_t sb = b200lc_cuhd_encode_scratch_bytes(size_in);
        cudaMalloc(&d_in, size_in + 16);
        cudaMalloc(&d_len, 256);
        cudaMalloc(&d_code, 256 * sizeof(std::uint32_t));
        cudaMalloc(&d_units, cap * sizeof(std::uint32_t) + 16);
        cudaMalloc(&d_bits, sizeof(std::uint64_t));
        cudaMalloc(&d_scratch, sb);
        CUERR
        cudaMemcpy(d_in, in, size_in, cudaMemcpyHostToDevice);
        cudaMemcpy(d_len, len, 256, cudaMemcpyHostToDevice);
        cudaMemcpy(d_code, code, sizeof(code), cudaMemcpyHostToDevice);
        const int rc = b200lc_cuhd_encode(d_in, size_in, d_code, d_len, d_units, cap, d_bits, d_scratch,
                                          sb, nullptr);
        if (rc != B200LC_OK || b200lc_cuhd_encode_overflowed(d_scratch, nullptr) != B200LC_OK) {
            std::cout << "b200lc_cuhd_encode failed" << std::endl;
            exit(1);
        }
        cudaMemcpy(out, d_units, size_out * sizeof(UNIT_TYPE), cudaMemcpyDeviceToHost);
        CUERR
        cudaFree(d_in); cudaFree(d_len); cudaFree(d_code); cudaFree(d_units); cudaFree(d_bits);
        cudaFree(d_scratch);
    }

    static std::shared_ptr<cuhd::CUHDCodetable> get_decoder_table(
        std::shared_ptr<LLHuffmanEncoderTable> enc_table) {
        auto table = std::make_shared<cuhd::CUHDCodetable>(enc_table->dict.size());
        cuhd::CUHDCodetableItemSingle* dec = table->get();
        for (auto& kv : enc_table->dict) {
            const size_t shift = MAX_CODEWORD_LENGTH - kv.second.length;
            for (size_t j = 0; j < (size_t(1) << shift); ++j)
                dec[(static_cast<size_t>(kv.second.codeword) << shift) + j] = {
                    static_cast<BIT_COUNT_TYPE>(kv.second.length), kv.first};
        }
        return table;
    }
};

}  // namespace llhuff
#endif /* B200LC_LLHUFF_COMPAT_H_ */
