// synth_line.cuh -- one line of the synthetic pileup (bench / tests only; SURVEY.md section 8d's input spec), host and device.
//
// Every random draw is a counter-based hash of (seed, sample, position, slot), so any sample can be regenerated anywhere:
// synth.cu writes a sample's text straight into HBM, oracle/synth_host.cpp writes the very same bytes on the host (the
// input of bench.py --impl reference, which must not touch the GPU library).
// Mix (rates from the bundled lambda-virus pileups): depth ~ Binomial(64, 3/8) scaled to mean_depth and clipped
// to [0, 60]; '.'/',' by strand; 0.7 % substitutions; 0.8 % N/n; 0.15 % '*'; '^'+MAPQ and '$' at 1/130 per
// read; one indel token on 0.1 % of the lines; zero-depth lines at 0.08 %; qualities Phred 13..39.
#pragma once
#include <string.h>
#include "hd.cuh"
#include "../../include/snpgpu.h"

namespace snpgpu {

struct SynthArgs {
    uint64_t seed;
    uint32_t sample, genome_len, mean_depth;
    uint32_t pool_thresh;      // a position is a pool site iff hash32(pool) < pool_thresh
    uint32_t carry_thresh;     // a sample carries a pool site iff hash32(carry) < carry_thresh
    uint32_t indel_thresh;     // a line carries an indel token iff its 24-bit draw < indel_thresh (0.1 % by default)
    int      name_len;
    char     name[64];
};

SNP_HD uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
SNP_HD uint64_t draw(uint64_t seed, uint64_t sample, uint64_t pos, uint64_t slot) {
    return mix64(mix64(seed ^ (sample * 0xd1342543de82ef95ull)) ^ mix64(pos * 0x2545f4914f6cdd1dull + slot));
}
// properties shared by every sample: reference base, pool membership, alternate allele
SNP_HD unsigned synth_ref_base(uint64_t seed, uint32_t pos) { return "ACGT"[draw(seed, 0xffffffffull, pos, 1) & 3u]; }
SNP_HD bool synth_is_pool(uint64_t seed, uint32_t pos, uint32_t pool_thresh) {
    return (uint32_t)(draw(seed, 0xffffffffull, pos, 2) >> 32) < pool_thresh;
}
SNP_HD unsigned synth_alt_base(uint64_t seed, uint32_t pos) {
    unsigned r = (unsigned)(draw(seed, 0xffffffffull, pos, 1) & 3u);
    unsigned k = (unsigned)((draw(seed, 0xffffffffull, pos, 3) >> 8) % 3u) + 1u;
    return "ACGT"[(r + k) & 3u];
}
SNP_HD bool synth_carries(const SynthArgs &a, uint32_t pos) {
    if (!synth_is_pool(a.seed, pos, a.pool_thresh)) return false;
    // half of the carriage is clade-structured (shared by samples of the same sample % 8), half private
    uint64_t clade = draw(a.seed, 0xfffffff0ull + (a.sample & 7u), pos, 4);
    uint64_t own = draw(a.seed, a.sample, pos, 5);
    return (uint32_t)(clade >> 32) < a.carry_thresh / 2u || (uint32_t)(own >> 32) < a.carry_thresh / 2u;
}

// Writes the line for `pos` to out (when non-null) and returns its length.
template <bool WRITE>
SNP_HD uint32_t synth_line(const SynthArgs &a, uint32_t pos, uint8_t *out) {
    uint32_t n = 0;
#define PUT(ch) do { const uint8_t put_ch_ = (uint8_t)(ch); if (WRITE) out[n] = put_ch_; n++; } while (0)
    for (int i = 0; i < a.name_len; i++) PUT(a.name[i]);
    PUT('\t');
    char dig[12];
    int nd = 0;
    uint32_t v = pos;
    do { dig[nd++] = (char)('0' + v % 10u); v /= 10u; } while (v);
    while (nd) PUT(dig[--nd]);
    PUT('\t');
    const unsigned ref = synth_ref_base(a.seed, pos);
    PUT(ref);
    PUT('\t');
    const uint64_t h0 = draw(a.seed, a.sample, pos, 8);
    uint32_t depth;
    {
        uint64_t r1 = draw(a.seed, a.sample, pos, 9), r2 = draw(a.seed, a.sample, pos, 10), r3 = draw(a.seed, a.sample, pos, 11);
        uint32_t b = (uint32_t)popc32((uint32_t)(r1 & (r2 | r3))) + (uint32_t)popc32((uint32_t)((r1 & (r2 | r3)) >> 32));
        depth = (b * a.mean_depth + 12u) / 24u;
        if (depth > 60u) depth = 60u;
    }
    if ((uint32_t)(h0 & 0xffffffu) < 13422u) depth = 0;                 // 0.08 % of 2^24
    if (depth == 0) {
        PUT('0'); PUT('\t'); PUT('*'); PUT('\t'); PUT('*'); PUT('\n');
        return n;
    }
    if (depth >= 10u) PUT('0' + depth / 10u);
    PUT('0' + depth % 10u);
    PUT('\t');
    const bool variant = synth_carries(a, pos);
    const unsigned alt = synth_alt_base(a.seed, pos);
    const bool has_indel = (uint32_t)((h0 >> 24) & 0xffffffu) < a.indel_thresh;
    const uint32_t indel_at = (uint32_t)((h0 >> 48) % depth);
    for (uint32_t r = 0; r < depth; r++) {
        const uint64_t h = draw(a.seed, a.sample, pos, 16 + r);
        const bool fwd = (h & 1u) != 0;
        const uint32_t u = (uint32_t)((h >> 1) & 0xfffffu);               // 20 bits
        if (((h >> 21) & 0x7fffu) < 252u) {                               // 1/130 of 2^15
            PUT('^');
            PUT("KIUS!~]"[(h >> 36) % 7u]);
        }
        unsigned c;
        const uint32_t ecls = (uint32_t)(mix64(h) & 0xfffffu);            // independent 20-bit draw: error class
        if (variant && u < 1017118u) c = fwd ? alt : (alt | 0x20u);      // 97 % of the reads show the allele
        else if (ecls < 1573u) c = '*';                                   // 0.15 %
        else if (ecls < 1573u + 7340u) {                                  // 0.7 % substitution
            unsigned o = "ACGT"[((ref == 'A' ? 0u : ref == 'C' ? 1u : ref == 'G' ? 2u : 3u) + 1u + (unsigned)((h >> 40) % 3u)) & 3u];
            c = fwd ? o : (o | 0x20u);
        } else if (ecls < 1573u + 7340u + 8389u) c = fwd ? 'N' : 'n';     // 0.8 %
        else c = fwd ? '.' : ',';
        PUT(c);
        if (has_indel && r == indel_at) {
            const uint32_t len = 1u + (uint32_t)((h >> 44) % 13u);
            PUT((h >> 43) & 1u ? '+' : '-');
            if (len >= 10u) PUT('1');
            PUT('0' + len % 10u);
            for (uint32_t k = 0; k < len; k++) {
                unsigned o = "ACGT"[(draw(a.seed, a.sample, pos, 200 + k) >> 5) & 3u];
                PUT(fwd ? o : (o | 0x20u));
            }
        }
        if (((h >> 50) & 0x3fffu) < 126u) PUT('$');                       // 1/130 of 2^14
    }
    PUT('\t');
    for (uint32_t r = 0; r < depth; r++) {
        const uint64_t h = draw(a.seed, a.sample, pos, 100 + r);
        PUT('.' + (unsigned)(h % 27u));                                    // Phred 13..39
    }
    PUT('\n');
#undef PUT
    return n;
}

inline SynthArgs synth_make_args(const snpgpu_synth_spec &spec, const char *contig_name) {
    SynthArgs a;
    memset(&a, 0, sizeof(a));
    a.seed = spec.seed; a.sample = spec.sample; a.genome_len = spec.genome_len;
    a.mean_depth = spec.mean_depth ? spec.mean_depth : 24u;
    double pf = spec.genome_len ? (double)spec.n_pool_sites / (double)spec.genome_len : 0.0;
    if (pf > 1.0) pf = 1.0;
    a.pool_thresh = (uint32_t)(pf * 4294967295.0);
    double cp = spec.site_carry_prob;
    if (cp < 0) cp = 0;
    if (cp > 1) cp = 1;
    a.carry_thresh = (uint32_t)(cp * 4294967295.0);
    double ir = spec.indel_line_rate > 0 ? spec.indel_line_rate : 0.001;       // lines with one indel token
    if (ir > 1) ir = 1;
    a.indel_thresh = (uint32_t)(ir * 16777216.0);
    if (contig_name) {
        size_t L = strlen(contig_name);
        if (L > 63) L = 63;
        memcpy(a.name, contig_name, L);
        a.name_len = (int)L;
    }
    return a;
}


}  // namespace snpgpu
