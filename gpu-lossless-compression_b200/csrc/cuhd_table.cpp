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
        // Package-merge without materialising the packages' contents: row l keeps, per item, its
        // weight and whether it is a leaf.  Leaves enter every row in sorted order and packages
        // pair up the previous row in order, so "the first t items of row l" is described by
        // (number of leaves a_l, number of packages p_l) and pulls the first 2 p_l items of row
        // l-1.  Symbol q (rank in sorted order) gets one more bit for every row with q < a_l.
        // Ties: a leaf goes before a package of equal weight.
        std::vector<uint64_t> w(k);
        for (size_t i = 0; i < k; ++i) w[i] = hist[syms[i]];
        std::vector<std::vector<uint64_t>> weight(max_len);
        std::vector<std::vector<uint8_t>> is_leaf(max_len);
        weight[0] = w;
        is_leaf[0].assign(k, 1);
        for (int level = 1; level < max_len; ++level) {
            const std::vector<uint64_t> &prev = weight[level - 1];
            const size_t np = prev.size() / 2;
            std::vector<uint64_t> &cur = weight[level];
            std::vector<uint8_t> &leaf = is_leaf[level];
            cur.reserve(np + k);
            leaf.reserve(np + k);
            size_t a = 0, b = 0;
            while (a < k || b < np) {
                const uint64_t pw = b < np ? prev[2 * b] + prev[2 * b + 1] : 0;
                const bool take_leaf = b >= np || (a < k && w[a] <= pw);
                if (take_leaf) { cur.push_back(w[a++]); leaf.push_back(1); }
                else { cur.push_back(pw); leaf.push_back(0); ++b; }
            }
        }
        size_t take = 2 * (k - 1);
        if (weight[max_len - 1].size() < take) return B200LC_ERR_UNSUPPORTED;
        std::vector<unsigned> len(k, 0);
        for (int level = max_len - 1; level >= 0 && take > 0; --level) {
            size_t leaves_taken = 0;
            for (size_t i = 0; i < take; ++i) leaves_taken += is_leaf[level][i];
            for (size_t q = 0; q < leaves_taken; ++q) ++len[q];
            take = 2 * (take - leaves_taken);
        }
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
