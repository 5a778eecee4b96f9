// CPU check of the CUHD decoder's per-thread walks (gpu-lossless-compression_b200/csrc/cuhd_walks.cuh,
// compiled for the host) against a bit-serial decode through the flat LUT -- the decode contract
// of cuhd-icpp/src/cuhd_gpu_decoder.cu:16-143.  Test infrastructure.
//   walk_record / walk_record_multi : codeword-start masks, exit state, symbol count from bit 0
//   walk_merge                      : entry state a -> (exit state, count) via the recorded path
//   walk_record_sym / walk_merge_skip / walk_emit : decode-once building blocks
//   walk_write / walk_write2        : symbols from the true entry state (two-symbol entries, both layouts)
// Usage: cuhd_walks_host [seed] [rounds]; exit code 0 = all equal.
//        cuhd_walks_host stream <lut.bin> <units.bin> <L>   (statistics of a real stream)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <vector>

#include "cuhd_walks.cuh"

using namespace b200lc::cuhd;

static const int S = 8;

// random complete prefix code with lengths <= L over `nsym` symbols -> flat LUT {len, sym}
static std::vector<u16> random_lut(std::mt19937_64 &rng, u32 L, u32 nsym, bool complete)
{
    // split leaves of a binary tree at random until nsym leaves, depth limited to L
    std::vector<u32> depth(1, 0);
    while (depth.size() < nsym) {
        std::vector<size_t> cand;
        for (size_t i = 0; i < depth.size(); ++i)
            if (depth[i] < L) cand.push_back(i);
        if (cand.empty()) break;
        // bias towards splitting shallow leaves sometimes, deep leaves other times
        size_t pick = cand[rng() % cand.size()];
        if (rng() % 3 == 0) pick = *std::min_element(cand.begin(), cand.end(), [&](size_t a, size_t b) { return depth[a] < depth[b]; });
        const u32 d = depth[pick] + 1;
        depth[pick] = d;
        depth.push_back(d);
    }
    std::sort(depth.begin(), depth.end());
    std::vector<u16> lut(size_t(1) << L, 0);   // 0 = unused prefix
    u32 code = 0, prev = depth[0];
    std::vector<u32> syms(depth.size());
    for (size_t i = 0; i < syms.size(); ++i) syms[i] = (u32)(rng() & 0xff);
    for (size_t i = 0; i < depth.size(); ++i) {
        const u32 len = depth[i];
        code <<= (len - prev);
        prev = len;
        if (!(complete == false && i + 1 == depth.size() && depth.size() > 1)) {   // incomplete: drop the last code
            if (len == 0) {   // single leaf: the encoder gives it the code 0 of length 1
                for (size_t j = 0; j < lut.size() / 2; ++j) lut[j] = (u16)(1 | (syms[i] << 8));
            } else {
                const u32 first = code << (L - len);
                for (u32 j = 0; j < (1u << (L - len)); ++j) lut[first + j] = (u16)(len | (syms[i] << 8));
            }
        }
        ++code;
    }
    return lut;
}

struct Serial {
    std::vector<u32> starts;   // bit positions of codeword starts in [a, 32*S)
    std::vector<u8> syms;
    u32 end;                   // first start at or after 32*S, minus 32*S
};

static u32 window(const u32 *u, u32 bit, u32 L)
{
    const u32 j = bit >> 5, o = bit & 31;
    const u32 w = o ? (u[j] << o) | (u[j + 1] >> (32 - o)) : u[j];
    return w >> (32 - L);
}

static Serial serial_decode(const u32 *u, const std::vector<u16> &lut, u32 L, u32 a)
{
    Serial r;
    u32 at = a;
    while (at < 32u * S) {
        const u32 i = window(u, at, L);
        u32 len = lut[i] & 0xff;
        if (len == 0 || len > L) len = 1;
        r.starts.push_back(at);
        r.syms.push_back((u8)(lut[i] >> 8));
        at += len;
    }
    r.end = at - 32u * S;
    return r;
}

// Statistics of a real stream (LUT file: u16[1 << L], units file: u32[]), every subsequence entered
// in its true state: how many symbols the merge walk decodes before the recorded path takes over,
// how many subsequences never land on it, how full the symbol slots get.  Also checks the
// decode-once composition against the serial decode of the whole stream.
static int stream_stats(const char *lut_path, const char *units_path, u32 L)
{
    std::vector<u16> lut(size_t(1) << L);
    FILE *f = fopen(lut_path, "rb");
    if (!f || fread(lut.data(), 2, lut.size(), f) != lut.size()) { printf("cannot read %s\n", lut_path); return 2; }
    fclose(f);
    f = fopen(units_path, "rb");
    if (!f) { printf("cannot read %s\n", units_path); return 2; }
    std::vector<u32> units;
    u32 buf[4096];
    size_t got;
    while ((got = fread(buf, 4, 4096, f)) > 0) units.insert(units.end(), buf, buf + got);
    fclose(f);
    const size_t nsub = units.size() / S;
    units.resize((nsub + 1) * S + 2, 0);
    const u32 shift = 32 - L;
    std::vector<u8> ltab(size_t(1) << L);
    std::vector<u16> stab(size_t(1) << L);
    for (u32 i = 0; i < (1u << L); ++i) {
        ltab[i] = (u8)first_len(lut.data(), i, L);
        stab[i] = len_sym_entry(lut.data(), i, L);
    }
    u32 entry = 0;
    unsigned long long k_sum = 0, never = 0, symbols = 0, over64 = 0, over48 = 0, k_max = 0, k_hist[6] = {0};
    for (size_t q = 0; q < nsub; ++q) {
        u32 un[S + 1];
        memcpy(un, &units[q * S], sizeof(un));
        u32 m[S], e0, c0, ne, k, skip;
        u8 slot[32 * S];
        walk_record_sym<S, 32 * S>(un, stab.data(), shift, m, e0, c0, slot);
        walk_merge_skip<S>(un, m, entry, e0, c0, ltab.data(), shift, ne, k, skip);
        const Serial sa = serial_decode(un, lut, L, entry);
        std::vector<u8> out(k + 8);
        walk_emit<S>(un, stab.data(), shift, entry, k, out.data());
        out.resize(k);
        out.insert(out.end(), slot + skip, slot + c0);
        if (out != sa.syms || ne != sa.end) { printf("decode-once mismatch at subsequence %zu\n", q); return 1; }
        k_sum += k; symbols += out.size();
        if (skip == c0 && k) ++never;
        if (c0 > 64) ++over64;
        if (c0 > 48) ++over48;
        if (k > k_max) k_max = k;
        ++k_hist[k < 5 ? k : 5];
        entry = ne;
    }
    printf("subsequences %zu, symbols %llu (%.1f per subsequence)\n", nsub, symbols, (double)symbols / nsub);
    printf("own symbols before the recorded path takes over: mean %.3f, max %llu; k=0..4,5+: %.3f %.3f %.3f %.3f %.3f %.3f\n",
           (double)k_sum / nsub, k_max, (double)k_hist[0] / nsub, (double)k_hist[1] / nsub, (double)k_hist[2] / nsub,
           (double)k_hist[3] / nsub, (double)k_hist[4] / nsub, (double)k_hist[5] / nsub);
    printf("never land on the recorded path: %.5f; recorded path longer than 64 symbols: %.5f, than 48: %.5f\n",
           (double)never / nsub, (double)over64 / nsub, (double)over48 / nsub);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc > 4 && !strcmp(argv[1], "stream")) return stream_stats(argv[2], argv[3], (u32)atoi(argv[4]));
    const unsigned long long seed = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    const int rounds = argc > 2 ? atoi(argv[2]) : 300;
    std::mt19937_64 rng(seed);
    long checked = 0, multi_steps = 0, single_steps = 0, fix_symbols = 0, fix_cases = 0;
    for (int round = 0; round < rounds; ++round) {
        const u32 L = 1 + (u32)(rng() % 13);                        // 1..13
        const u32 max_sym = std::min<u32>(256, 1u << L);
        const u32 nsym = 1 + (u32)(rng() % max_sym);
        const bool complete = rng() % 4 != 0;
        const std::vector<u16> lut = random_lut(rng, L, nsym, complete);
        const u32 shift = 32 - L;
        std::vector<u8> ltab(size_t(1) << L);
        std::vector<u16> mtab(size_t(1) << L), stab(size_t(1) << L);
        std::vector<u32> wtab(size_t(1) << L), wtab2(size_t(1) << L);
        for (u32 i = 0; i < (1u << L); ++i) {
            ltab[i] = (u8)first_len(lut.data(), i, L);
            mtab[i] = multi_entry(lut.data(), i, L);
            stab[i] = len_sym_entry(lut.data(), i, L);
            wtab[i] = write_entry(lut.data(), i, L);
            wtab2[i] = write_entry2(lut.data(), i, L);
        }
        const u32 LMs[3] = {L, std::max<u32>(L, 13), 15};
        std::vector<u8> ctab[3], stab1(size_t(1) << L);
        for (int t = 0; t < 3; ++t) {
            ctab[t].resize(size_t(1) << LMs[t]);
            for (u32 i = 0; i < (1u << LMs[t]); ++i) ctab[t][i] = count_entry(lut.data(), i, L, LMs[t], 3);
        }
        for (u32 i = 0; i < (1u << L); ++i) stab1[i] = count_entry(lut.data(), i, L, L, 1);
        for (int sub = 0; sub < 64; ++sub) {
            u32 u[S + 2];
            const int kind = (int)(rng() % 5);
            for (int j = 0; j < S + 2; ++j) {
                u32 x = (u32)rng();
                if (kind == 1) x &= (u32)rng() & (u32)rng();        // sparse ones: long runs of the all-zero code
                if (kind == 2) x |= (u32)rng() | (u32)rng();        // dense ones
                if (kind == 3) x = 0;
                if (kind == 4) x = 0xffffffffu;
                u[j] = x;
            }
            u32 un[S + 1];
            memcpy(un, u, sizeof(un));
            // --- walk_record vs serial from bit 0
            const Serial s0 = serial_decode(u, lut, L, 0);
            u32 m[S], e0 = 0, c0 = 0;
            walk_record<S>(un, ltab.data(), shift, m, e0, c0);
            u32 want[S] = {0};
            for (u32 b : s0.starts) want[b >> 5] |= 0x80000000u >> (b & 31);
            if (memcmp(m, want, sizeof(m)) || e0 != s0.end || c0 != s0.starts.size()) {
                printf("walk_record mismatch: seed %llu round %d sub %d L %u\n", seed, round, sub, L);
                return 1;
            }
            single_steps += c0;
            // --- walk_record_multi == walk_record
            u32 mm[S], e1 = 0, c1 = 0;
            walk_record_multi<S>(un, mtab.data(), shift, mm, e1, c1);
            if (memcmp(mm, m, sizeof(m)) || e1 != e0 || c1 != c0) {
                printf("walk_record_multi mismatch: seed %llu round %d sub %d L %u (end %u/%u cnt %u/%u)\n",
                       seed, round, sub, L, e1, e0, c1, c0);
                return 1;
            }
            for (u32 at = 0; at < 32u * S;) { at += mtab[window(u, at, L)] >> 12; ++multi_steps; }
            // --- decode-once building blocks: the recorded path with its symbols
            u32 ms[S], e2 = 0, c2 = 0;
            u8 slot[32 * S];
            const bool fits = walk_record_sym<S, 32 * S>(un, stab.data(), shift, ms, e2, c2, slot);
            if (!fits || memcmp(ms, m, sizeof(m)) || e2 != e0 || c2 != c0 || memcmp(slot, s0.syms.data(), c0)) {
                printf("walk_record_sym mismatch: seed %llu round %d sub %d L %u\n", seed, round, sub, L);
                return 1;
            }
            u8 small_slot[16];
            u32 ms2[S];
            if (walk_record_sym<S, 16>(un, stab.data(), shift, ms2, e2, c2, small_slot) != (c0 <= 16)) {
                printf("walk_record_sym overflow flag wrong: seed %llu round %d sub %d\n", seed, round, sub);
                return 1;
            }
            // --- walk_merge / walk_write for every entry state
            for (u32 a = 0; a < L; ++a) {
                const Serial sa = serial_decode(u, lut, L, a);
                u32 ne = 0, nc = 0;
                // round-2 counting walk (no masks), every table width
                for (int t = 0; t < 3; ++t) {
                    u32 we = 99, wc = 99;
                    walk_count<S>(un, ctab[t].data(), 32 - LMs[t], stab1.data(), shift, a, we, wc);
                    if (we != sa.end || wc != sa.starts.size()) {
                        printf("walk_count mismatch: seed %llu round %d sub %d L %u LM %u a %u (end %u/%u cnt %u/%zu)\n",
                               seed, round, sub, L, LMs[t], a, we, sa.end, wc, sa.starts.size());
                        return 1;
                    }
                }
                walk_merge<S>(un, m, a, e0, ltab.data(), shift, ne, nc);
                if (ne != sa.end || nc != sa.starts.size()) {
                    printf("walk_merge mismatch: seed %llu round %d sub %d L %u a %u (end %u/%u cnt %u/%zu)\n",
                           seed, round, sub, L, a, ne, sa.end, nc, sa.starts.size());
                    return 1;
                }
                // decode once: k own symbols, then the recorded ones from `skip` on
                {
                    u32 ne2 = 0, k = 0, skip = 0;
                    walk_merge_skip<S>(un, m, a, e0, c0, ltab.data(), shift, ne2, k, skip);
                    std::vector<u8> got(k + 8, 0xEE);
                    walk_emit<S>(un, stab.data(), shift, a, k, got.data());
                    got.resize(k);
                    got.insert(got.end(), slot + skip, slot + c0);
                    if (ne2 != sa.end || got != sa.syms) {
                        printf("decode-once mismatch: seed %llu round %d sub %d L %u a %u (k %u skip %u c0 %u, %zu/%zu symbols)\n",
                               seed, round, sub, L, a, k, skip, c0, got.size(), sa.syms.size());
                        return 1;
                    }
                    fix_symbols += k;
                    fix_cases += 1;
                }
                // walk_write may store one symbol past the subsequence (the successor's first)
                std::vector<u8> dst(sa.syms.size() + 40, 0xEE), chk(sa.syms.size() + 40, 0xEE);
                walk_write<S, false>(un, wtab.data(), shift, a, dst.data() + 8, 0, 0, 0);
                if (memcmp(dst.data() + 8, sa.syms.data(), sa.syms.size())) {
                    printf("walk_write mismatch: seed %llu round %d sub %d L %u a %u\n", seed, round, sub, L, a);
                    return 1;
                }
                for (int q = 0; q < 8; ++q)
                    if (dst[q] != 0xEE) { printf("walk_write wrote before dst\n"); return 1; }
                // layout 2 stores exactly the same bytes
                std::vector<u8> dst2(dst.size(), 0xEE), chk2(dst.size(), 0xEE);
                walk_write2<S, false>(un, wtab2.data(), shift, a, dst2.data() + 8, 0, 0, 0);
                if (dst2 != dst) {
                    printf("walk_write2 mismatch: seed %llu round %d sub %d L %u a %u\n", seed, round, sub, L, a);
                    return 1;
                }
                // CHECK variant: only positions in [lo, hi) are stored
                const u32 lo = sa.syms.size() / 3, hi = std::max<u32>(lo, (u32)(2 * sa.syms.size() / 3));
                walk_write<S, true>(un, wtab.data(), shift, a, chk.data() + 8, 0, lo, hi);
                walk_write2<S, true>(un, wtab2.data(), shift, a, chk2.data() + 8, 0, lo, hi);
                if (chk2 != chk) {
                    printf("walk_write2<CHECK> mismatch: seed %llu round %d sub %d L %u a %u\n", seed, round, sub, L, a);
                    return 1;
                }
                for (u32 q = 0; q < sa.syms.size() + 32; ++q) {
                    const u8 expect = (q >= lo && q < hi && q < sa.syms.size()) ? sa.syms[q] : (u8)0xEE;
                    if (q >= sa.syms.size() && q >= lo && q < hi) continue;   // the duplicate of the successor's first symbol
                    if (chk[8 + q] != expect) {
                        printf("walk_write<CHECK> mismatch: seed %llu round %d sub %d L %u a %u q %u\n",
                               seed, round, sub, L, a, q);
                        return 1;
                    }
                }
                ++checked;
            }
        }
    }
    // --- packed write pass (walk_write3): a warp of 32 consecutive subsequences staged in two
    // phases, lanes visited in random order inside each phase (no ordering between lanes may be
    // assumed on the GPU), staging window and carried bytes as in segment_pass_b
    long warps = 0;
    for (int round = 0; round < rounds; ++round) {
        const u32 L = 1 + (u32)(rng() % 13);
        const u32 LW = std::max<u32>(L, (u32)(rng() % 3 == 0 ? L : 11 + rng() % 4));
        const u32 nsym = 1 + (u32)(rng() % std::min<u32>(256, 1u << L));
        const std::vector<u16> lut = random_lut(rng, L, nsym, rng() % 4 != 0);
        std::vector<u32> tab3(size_t(1) << LW);
        std::vector<u8> ctab(size_t(1) << LW), stab1(size_t(1) << L);
        for (u32 i = 0; i < (1u << LW); ++i) {
            tab3[i] = write_entry3(lut.data(), i, L, LW);
            ctab[i] = count_entry(lut.data(), i, L, LW, 3);
            if ((tab3[i] >> 24) != ctab[i]) { printf("write_entry3 top byte != count_entry\n"); return 1; }
        }
        for (u32 i = 0; i < (1u << L); ++i) stab1[i] = count_entry(lut.data(), i, L, L, 1);
        for (int rep = 0; rep < 4; ++rep) {
            const int nl = 1 + (int)(rng() % 32);
            std::vector<u32> units((size_t)nl * S + 2);
            const int kind = (int)(rng() % 4);
            for (auto &x : units) {
                x = (u32)rng();
                if (kind == 1) x &= (u32)rng() & (u32)rng();
                if (kind == 2) x |= (u32)rng() | (u32)rng();
            }
            // serial decode of the whole range from a random entry state
            u32 entry[33], cnt[33], pre[33];
            std::vector<u8> all;
            entry[0] = (u32)(rng() % L);
            pre[0] = 0;
            for (int l = 0; l < nl; ++l) {
                const Serial sa = serial_decode(&units[(size_t)l * S], lut, L, entry[l]);
                cnt[l] = (u32)sa.syms.size();
                all.insert(all.end(), sa.syms.begin(), sa.syms.end());
                entry[l + 1] = sa.end;
                pre[l + 1] = pre[l] + cnt[l];
                u32 un[S + 1], we, wc;
                memcpy(un, &units[(size_t)l * S], sizeof(un));
                walk_count<S>(un, ctab.data(), 32 - LW, stab1.data(), 32 - L, entry[l], we, wc);
                if (we != sa.end || wc != cnt[l]) { printf("walk_count (LW table) mismatch\n"); return 1; }
            }
            const u32 total = pre[nl];
            const u32 win = 64 + (u32)(rng() % 2500);
            u32 fill = (u32)(rng() % 16);
            std::vector<u8> out;
            std::vector<u8> stage(16 + win + 64, 0xEE);
            std::vector<u8> carried(fill);
            for (auto &c : carried) c = (u8)rng();
            memcpy(stage.data(), carried.data(), fill);
            for (u32 lo = 0; lo < total; lo += win) {
                const u32 hi = std::min(total, lo + win);
                int order[32];
                for (int l = 0; l < nl; ++l) order[l] = l;
                u32 pend[32] = {0};
                bool fast[32] = {false}, mine[32] = {false};
                std::shuffle(order, order + nl, rng);
                for (int t = 0; t < nl; ++t) {
                    const int l = order[t];
                    mine[l] = pre[l] < hi && pre[l] + cnt[l] > lo;
                    fast[l] = mine[l] && pre[l] >= lo && pre[l] + cnt[l] <= hi;
                    if (!fast[l]) continue;
                    u32 un[S + 1];
                    memcpy(un, &units[(size_t)l * S], sizeof(un));
                    const u32 d0 = fill + pre[l] - lo;
                    pend[l] = walk_write3<S>(un, smem_of(tab3.data()), 32 - LW, entry[l], cnt[l], smem_of(stage.data()), d0,
                                             walk_write3_head(stage.data(), d0, fill));
                }
                std::shuffle(order, order + nl, rng);
                for (int t = 0; t < nl; ++t) {
                    const int l = order[t];
                    u32 un[S + 1];
                    memcpy(un, &units[(size_t)l * S], sizeof(un));
                    if (fast[l]) walk_write3_tail(stage.data(), fill + pre[l] - lo, cnt[l], pend[l]);
                    else if (mine[l])
                        walk_write3_bytes<S>(un, tab3.data(), 32 - LW, entry[l], cnt[l],
                                             stage.data() + ((int)fill - (int)lo), pre[l], lo, hi);
                }
                fill += hi - lo;
                for (u32 q = fill; q < stage.size(); ++q)
                    if (q >= ((fill + 3) & ~3u) && stage[q] != 0xEE) { printf("packed write stored beyond the window\n"); return 1; }
                const u32 nvec = fill >> 4;
                out.insert(out.end(), stage.begin(), stage.begin() + 16 * nvec);
                const u32 tail = fill & 15u;
                std::vector<u8> keep(stage.begin() + 16 * nvec, stage.begin() + 16 * nvec + tail);
                std::fill(stage.begin(), stage.end(), (u8)0xEE);
                memcpy(stage.data(), keep.data(), tail);
                fill = tail;
            }
            out.insert(out.end(), stage.begin(), stage.begin() + fill);
            std::vector<u8> want(carried);
            want.insert(want.end(), all.begin(), all.end());
            if (out != want) {
                size_t q = 0;
                while (q < out.size() && q < want.size() && out[q] == want[q]) ++q;
                printf("packed write mismatch: seed %llu round %d L %u LW %u lanes %d win %u at byte %zu of %zu\n",
                       seed, round, L, LW, nl, win, q, want.size());
                return 1;
            }
            ++warps;
        }
    }
    printf("ok: %ld (subsequence, entry state) cases; %ld packed-write warps; %.2f symbols per multi-symbol lookup; "
           "%.2f symbols decoded before the recorded path takes over\n", checked, warps,
           multi_steps ? (double)single_steps / (double)multi_steps : 0.0,
           fix_cases ? (double)fix_symbols / (double)fix_cases : 0.0);
    return 0;
}
