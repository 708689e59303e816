// sites.cuh -- the site table: snplist.txt positions plus the exclude VCF's positions, as the pileup kernel
// probes them (call_consensus.py:117-125, :133, :147-151; the set test of pileup.py:425-427).
//
// Layout (all arrays in the memory of whoever dereferences them -- HBM for the kernel, host for cpu_sim):
//   one bit per (contig, position) from 0 up to the largest site position of that contig, contigs laid end
//   to end at bit_base[c] (a multiple of 32); rank[w] = number of set bits before bitmap word w, so a hit
//   resolves to the index u of the site among the sorted unique sites; flags[u] bit0 = in the snplist,
//   bit1 = in the exclude file.  5 Mbp of positions is 625 KB of bitmap + 625 KB of ranks: L2-resident.
#pragma once
#include "hd.cuh"

namespace snpgpu {

enum : uint8_t { SITE_SNP = 1, SITE_EXCLUDED = 2 };

// what the first-tier parser wants to know about 32 positions, in one 16-byte load: which are sites, which of those are
// in the snplist / in the exclude file, and how many sites lie in front of the word
struct SiteWord { uint32_t any, snp, exc, rank; };

struct SiteTable {
    int32_t         n_contigs;
    int32_t         n_unique;
    const uint32_t *names4;     // per contig: name bytes followed by '\t', zero-padded to whole words
    const int32_t  *off4;       // word offset of contig c's entry in names4 (n_contigs + 1)
    const int32_t  *len1;       // name length + 1 (the tab)
    const uint8_t  *names;      // the same names, packed, for byte-wise comparison
    const int32_t  *name_off;   // n_contigs + 1
    const int64_t  *bit_base;   // per contig, multiple of 32
    const int64_t  *max_pos;    // per contig, -1 when the contig holds no site
    const uint32_t *bits;
    const uint32_t *rank;
    const uint8_t  *flags;      // n_unique
    const SiteWord *words;      // the same facts packed per bitmap word (bits / flags / rank remain for the other tiers)
    const uint32_t *q3rows;     // per contig SITE_Q3ROWS_WORDS words: the name rows / masks of line_quick3.cuh's Q3Contig
};
constexpr uint32_t SITE_Q3ROWS_WORDS = 128;

// index of (contig, pos) among the unique sites, or -1
SNP_HD int32_t site_find(const SiteTable &t, int cid, int64_t pos) {
    if (cid < 0 || pos < 0 || pos > t.max_pos[cid]) return -1;
    int64_t bit = t.bit_base[cid] + pos;
    uint32_t w = t.bits[bit >> 5];
    uint32_t b = (uint32_t)bit & 31u;
    if (!((w >> b) & 1u)) return -1;
    return (int32_t)(t.rank[bit >> 5] + (uint32_t)popc32(w & ((1u << b) - 1u)));
}

// contig index of the token s[0..n), or -1 (byte-wise; the general path)
SNP_HD int contig_find(const SiteTable &t, const uint8_t *s, int64_t n) {
    for (int c = 0; c < t.n_contigs; c++) {
        int32_t o = t.name_off[c];
        if ((int64_t)(t.name_off[c + 1] - o) != n) continue;
        bool same = true;
        for (int64_t i = 0; i < n; i++) if (t.names[o + i] != s[i]) { same = false; break; }
        if (same) return c;
    }
    return -1;
}

}  // namespace snpgpu
