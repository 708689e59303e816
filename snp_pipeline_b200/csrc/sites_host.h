// sites_host.h -- host-side construction of the site table (sites.cuh) from snplist / exclude positions.
// Plain C++ (no CUDA): used by api.cu to build the arrays it uploads and by the cpu_sim test harness.
#pragma once
#include <algorithm>
#include <string>
#include <string.h>
#include <vector>
#include "sites.cuh"
#include "line_quick3.cuh"

namespace snpgpu {

struct HostNameAt {                         // (a functor, not a lambda: nvcc compiles this header for api.cu too)
    const char *nm;
    SNP_HD uint8_t operator()(uint32_t i) const { return (uint8_t)nm[i]; }
};

struct HostSites {
    int32_t n_contigs = 0;
    size_t  n_unique = 0;
    std::vector<uint32_t> names4;
    std::vector<int32_t>  off4, len1, name_off;
    std::vector<uint8_t>  names;
    std::vector<int64_t>  bit_base, max_pos;
    std::vector<uint32_t> bits, rank;
    std::vector<uint8_t>  flags;
    std::vector<SiteWord> words;
    std::vector<int32_t>  snp_unique;       // unique-site index of snplist entry k
    std::vector<uint32_t> q3rows;           // per contig: line_quick3.cuh's name rows

    SiteTable view() const {
        SiteTable t;
        t.n_contigs = n_contigs; t.n_unique = (int32_t)n_unique;
        t.names4 = names4.data(); t.off4 = off4.data(); t.len1 = len1.data();
        t.names = names.data(); t.name_off = name_off.data();
        t.bit_base = bit_base.data(); t.max_pos = max_pos.data();
        t.bits = bits.data(); t.rank = rank.data(); t.flags = flags.data(); t.words = words.data();
        t.q3rows = q3rows.data();
        return t;
    }
};

// Returns 0, or 1 = bad argument, 4 = outside the supported domain, 18 = too large; *why gets the reason.
inline int build_host_sites(const char *contig_names, const int32_t *name_off, int32_t n_contigs,
                            const int32_t *snp_contig, const int64_t *snp_pos, size_t n_snp,
                            const int32_t *exc_contig, const int64_t *exc_pos, size_t n_exc, HostSites *out,
                            const char **why) {
    *why = "";
    // contig names are whitespace-split tokens in the reference: non-empty printable ASCII
    for (int c = 0; c < n_contigs; c++) {
        if (name_off[c + 1] <= name_off[c]) { *why = "empty contig name"; return 1; }
        for (int32_t i = name_off[c]; i < name_off[c + 1]; i++) {
            unsigned ch = (unsigned char)contig_names[i];
            if (ch <= 0x20u || ch >= 0x7fu) { *why = "contig name outside printable ASCII"; return 4; }
        }
    }
    struct Key { int32_t c; int64_t p; };
    std::vector<Key> keys;
    keys.reserve(n_snp + n_exc);
    auto ok = [&](int32_t c, int64_t p) { return c >= 0 && c < n_contigs && p >= 0 && p < ((int64_t)1 << 31); };
    for (size_t i = 0; i < n_snp; i++) {
        if (!ok(snp_contig[i], snp_pos[i])) { *why = "snplist entry outside contig table or position outside [0, 2^31)"; return 4; }
        keys.push_back({snp_contig[i], snp_pos[i]});
    }
    for (size_t i = 0; i < n_exc; i++) {
        if (!ok(exc_contig[i], exc_pos[i])) { *why = "exclude entry outside contig table or position outside [0, 2^31)"; return 4; }
        keys.push_back({exc_contig[i], exc_pos[i]});
    }
    auto less = [](const Key &a, const Key &b) { return a.c != b.c ? a.c < b.c : a.p < b.p; };
    std::sort(keys.begin(), keys.end(), less);
    keys.erase(std::unique(keys.begin(), keys.end(), [](const Key &a, const Key &b) { return a.c == b.c && a.p == b.p; }),
               keys.end());
    HostSites &h = *out;
    h.n_contigs = n_contigs;
    h.n_unique = keys.size();
    const size_t nc1 = (size_t)std::max(n_contigs, 1);
    h.max_pos.assign(nc1, -1);
    h.bit_base.assign(nc1, 0);
    for (const Key &k : keys) h.max_pos[k.c] = std::max(h.max_pos[k.c], k.p);
    int64_t total_bits = 0;
    for (int c = 0; c < n_contigs; c++) {
        h.bit_base[c] = total_bits;
        total_bits += (h.max_pos[c] + 1 + 31) / 32 * 32;
    }
    if (total_bits > ((int64_t)1 << 33)) { *why = "site bitmap above 1 GiB"; return 18; }
    const size_t n_words = (size_t)(total_bits / 32) + 1;
    h.bits.assign(n_words, 0u);
    h.rank.assign(n_words, 0u);
    for (const Key &k : keys) { int64_t b = h.bit_base[k.c] + k.p; h.bits[b >> 5] |= 1u << (b & 31); }
    uint32_t run = 0;
    for (size_t w = 0; w < n_words; w++) { h.rank[w] = run; run += (uint32_t)__builtin_popcount(h.bits[w]); }
    h.flags.assign(h.n_unique + 1, 0);
    h.snp_unique.assign(n_snp + 1, 0);
    auto find = [&](int32_t c, int64_t p) {
        Key k{c, p};
        return (size_t)(std::lower_bound(keys.begin(), keys.end(), k, less) - keys.begin());
    };
    for (size_t i = 0; i < n_snp; i++) { size_t u = find(snp_contig[i], snp_pos[i]); h.flags[u] |= SITE_SNP; h.snp_unique[i] = (int32_t)u; }
    for (size_t i = 0; i < n_exc; i++) h.flags[find(exc_contig[i], exc_pos[i])] |= SITE_EXCLUDED;
    h.words.assign(n_words, SiteWord{0u, 0u, 0u, 0u});
    for (size_t w = 0; w < n_words; w++) { h.words[w].any = h.bits[w]; h.words[w].rank = h.rank[w]; }
    for (size_t u = 0; u < keys.size(); u++) {
        const int64_t b = h.bit_base[keys[u].c] + keys[u].p;
        if (h.flags[u] & SITE_SNP) h.words[b >> 5].snp |= 1u << (b & 31);
        if (h.flags[u] & SITE_EXCLUDED) h.words[b >> 5].exc |= 1u << (b & 31);
    }
    // padded contig entries: name + '\t', zero-filled to whole words
    h.off4.assign((size_t)n_contigs + 1, 0);
    h.len1.assign(nc1, 0);
    h.names4.clear();
    for (int c = 0; c < n_contigs; c++) {
        int32_t L = name_off[c + 1] - name_off[c];
        h.len1[c] = L + 1;
        h.off4[c] = (int32_t)h.names4.size();
        std::string e(contig_names + name_off[c], (size_t)L);
        e.push_back('\t');
        while (e.size() % 4) e.push_back('\0');
        for (size_t i = 0; i < e.size(); i += 4) { uint32_t w; memcpy(&w, e.data() + i, 4); h.names4.push_back(w); }
    }
    h.q3rows.assign(nc1 * SITE_Q3ROWS_WORDS, 0u);
    for (int c = 0; c < n_contigs; c++) {
        const uint32_t L = (uint32_t)(name_off[c + 1] - name_off[c]) + 1u;
        if (((L + 6u) >> 2) > Q3_NAMEW) continue;           // (a name too long for the first tier: never cached)
        const char *nm = contig_names + name_off[c];
        const HostNameAt name_at{nm};
        for (uint32_t k = 0; k < SITE_Q3ROWS_WORDS; k++) h.q3rows[(size_t)c * SITE_Q3ROWS_WORDS + k] = q3_rows_word(name_at, L, k);
    }
    h.off4[n_contigs] = (int32_t)h.names4.size();
    h.names4.push_back(0); h.names4.push_back(0);
    h.name_off.assign(name_off ? name_off : nullptr, name_off ? name_off + n_contigs + 1 : nullptr);
    if (h.name_off.empty()) h.name_off.push_back(0);
    const size_t names_len = n_contigs ? (size_t)name_off[n_contigs] : 0;
    h.names.assign((const uint8_t *)contig_names, (const uint8_t *)contig_names + names_len);
    h.names.push_back(0);
    return 0;
}

}  // namespace snpgpu
