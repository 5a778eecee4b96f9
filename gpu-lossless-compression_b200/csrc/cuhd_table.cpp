// Host-side code table construction for the CUHD stream format.
//
// Functional counterpart of llhuff::LLHuffmanEncoder::get_symbol_lengths / get_encoder_table /
// get_decoder_table (cuhd-icpp/encoder/src/llhuffman_encoder.cc:18-198,240-262): optimal
// length-limited prefix code (package-merge), canonical code assignment
// `code = (code + 1) << (next_len - cur_len)` over symbols sorted by length, flat LUT with
// 2^(L-len) entries per symbol.
//
// Differences from the reference, on purpose:
//   * weights are the exact integer counts, not float(count)/float(size) (:67-70), so the code
//     is optimal for every histogram and independent of rounding;
//   * ties inside one length class are broken by symbol value, not by std::unordered_map
//     iteration order (:143-155), so the table is reproducible across C++ runtimes.
// Any table produced here is a valid input to the reference decoder and to b200lc_cuhd_decode;
// streams encoded with a reference-built table are bit-identical to the reference's (the table
// is an input of the packer).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/b200lc.h"

namespace {

struct Item {
    uint64_t weight;
    std::vector<uint8_t> count;  // how many coins of each symbol (indexed by rank) are inside
};

}  // namespace

extern "C" int b200lc_cuhd_build_table(const uint64_t *hist, int max_len, uint32_t *code_of_symbol,
                                       uint8_t *len_of_symbol, uint8_t *lut)
{
    if (!hist || !code_of_symbol || !len_of_symbol) return B200LC_ERR_ARG;
    if (max_len < 1 || max_len > 13) return B200LC_ERR_UNSUPPORTED;
    std::memset(code_of_symbol, 0, 256 * sizeof(uint32_t));
    std::memset(len_of_symbol, 0, 256);

    std::vector<int> syms;
    for (int s = 0; s < 256; ++s)
        if (hist[s]) syms.push_back(s);
    const size_t k = syms.size();
    if (k == 0) return B200LC_ERR_ARG;
    if (k > (size_t(1) << max_len)) return B200LC_ERR_UNSUPPORTED;  // llhuffman_encoder.cc:30-32

    if (k == 1) {
        len_of_symbol[syms[0]] = 1;  // llhuffman_encoder.cc:38-46: single code "0"
    } else {
        // leaves sorted by (count, symbol)
        std::stable_sort(syms.begin(), syms.end(),
                         [&](int a, int b) { return hist[a] < hist[b]; });
        std::vector<Item> leaves(k);
        for (size_t i = 0; i < k; ++i) {
            leaves[i].weight = hist[syms[i]];
            leaves[i].count.assign(k, 0);
            leaves[i].count[i] = 1;
        }
        std::vector<Item> row = leaves;
        for (int level = 1; level < max_len; ++level) {
            std::vector<Item> packaged;
            for (size_t j = 0; j + 1 < row.size(); j += 2) {
                Item p;
                p.weight = row[j].weight + row[j + 1].weight;
                p.count = row[j].count;
                for (size_t q = 0; q < k; ++q) p.count[q] += row[j + 1].count[q];
                packaged.push_back(std::move(p));
            }
            std::vector<Item> merged;
            merged.reserve(packaged.size() + k);
            size_t a = 0, b = 0;
            while (a < leaves.size() || b < packaged.size()) {
                const bool take_leaf =
                    b >= packaged.size() ||
                    (a < leaves.size() && leaves[a].weight <= packaged[b].weight);
                if (take_leaf) merged.push_back(leaves[a++]);
                else merged.push_back(std::move(packaged[b++]));
            }
            row.swap(merged);
        }
        const size_t take = 2 * (k - 1);
        if (row.size() < take) return B200LC_ERR_UNSUPPORTED;
        std::vector<unsigned> len(k, 0);
        for (size_t i = 0; i < take; ++i)
            for (size_t q = 0; q < k; ++q) len[q] += row[i].count[q];
        for (size_t q = 0; q < k; ++q) len_of_symbol[syms[q]] = (uint8_t)len[q];
    }

    // canonical codes over (length, symbol) order
    std::vector<int> order;
    for (int s = 0; s < 256; ++s)
        if (len_of_symbol[s]) order.push_back(s);
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return len_of_symbol[a] < len_of_symbol[b]; });
    uint32_t code = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        const unsigned cur = len_of_symbol[order[i]];
        code_of_symbol[order[i]] = code;
        const unsigned next = i + 1 < order.size() ? len_of_symbol[order[i + 1]] : cur;
        code = (code + 1) << (next - cur);
    }

    if (lut) {
        std::memset(lut, 0, size_t(2) << max_len);
        for (int s = 0; s < 256; ++s) {
            const unsigned L = len_of_symbol[s];
            if (!L) continue;
            const unsigned shift = (unsigned)max_len - L;
            const uint32_t first = code_of_symbol[s] << shift;
            for (uint32_t j = 0; j < (1u << shift); ++j) {
                lut[2 * (first + j)] = (uint8_t)L;
                lut[2 * (first + j) + 1] = (uint8_t)s;
            }
        }
    }
    return B200LC_OK;
}

// Number of stream units (without the pad unit) for a histogram under a code
// (llhuffman_encoder.cc:166-180).
extern "C" size_t b200lc_cuhd_compressed_units(const uint64_t *hist, const uint8_t *len_of_symbol)
{
    uint64_t bits = 0;
    for (int s = 0; s < 256; ++s) bits += hist[s] * len_of_symbol[s];
    return (size_t)((bits + 31) / 32);
}
