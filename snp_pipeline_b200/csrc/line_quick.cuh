// line_quick.cuh -- the first-tier parser of the pileup kernel: decides the lines whose call does not depend
// on WHICH other symbols a read carries.
//
// After pileup.py:276-325 has stripped "^x", "$" and indel tokens, pileup.py:255-266 counts '.'/',' as the
// reference base and every other byte as itself.  When the '.'/',' outnumber all other surviving bytes TOGETHER
// and none of those bytes is the reference letter written out, the reference base is the strict winner of
// pileup.py:260-266 whatever the others are, and the caller (pileup.py:550-588) only needs
//     good_depth = surviving bytes,  consensus depth = '.' + ',',  forward = '.',  reverse = ','.
// That is >= 98 % of the lines of a real pileup.  This parser computes exactly those four numbers with SWAR
// masks, four bytes per step and no symbol-dependent branch, validates every byte of the line on the way
// (separators, digits, printable quality string as long as the stripped bases -- the zip() of pileup.py:248),
// and returns ST_DETAIL for anything else: another winner possible, an indel token, a "^^" chain, a trailing
// '^', an odd separator, "\r\n", depth 0, a contig other than the hinted one ...  ST_DETAIL has no side
// effects; the caller hands the line to line_fast.cuh (exact tallies) and, from there, to line_general.cuh.
//
// Precondition (the kernel checks it per tile): no byte >= 0x80 in the staged window, so byte lanes never
// carry into each other.  buf is 4-byte aligned with '\n' sentinels behind `limit`.
#pragma once
#include "line_fast.cuh"

namespace snpgpu {

enum : int { ST_DETAIL = 67 };

struct QuickLine {
    int32_t  site;         // unique-site index or -1
    uint32_t end;          // offset of the line terminator in buf
    uint8_t  base;         // consensus character before the '-' substitutions of call_consensus.py:169-176
    uint8_t  fail;         // FAIL_* mask (without FAIL_REGION)
};

// acc + 128 * (number of bytes of `flags` that are 0x80); flags holds 0x80 / 0x00 bytes only
SNP_HD uint32_t flag_sum(uint32_t flags, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp4a(flags, 0x01010101u, acc);
#else
    return acc + 128u * (uint32_t)__builtin_popcount(flags);
#endif
}

SNP_HD uint32_t funnel_l8(uint32_t lo, uint32_t hi) {       // (hi << 8) | (lo >> 24)
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, 8);
#else
    return (hi << 8) | (lo >> 24);
#endif
}

// sequential little-endian words from an arbitrary byte offset of a 4-byte aligned buffer: one load + one
// funnel shift per word (reads at most 7 bytes past the last byte asked for)
struct WordReader {
    const uint32_t *p;
    uint32_t lo, sh;
    SNP_HD void seek(const uint8_t *buf, uint32_t i) {
        p = reinterpret_cast<const uint32_t *>(buf + (i & ~3u));
        sh = (i & 3u) * 8u;
        lo = p[0];
    }
    SNP_HD uint32_t next() {
        const uint32_t hi = p[1];
        const uint32_t w = funnel_r(lo, hi, sh);
        lo = hi;
        ++p;
        return w;
    }
};

// The contig a warp currently expects, staged where its lanes reach it without a trip to global memory
// (the name in shared memory, the rest in registers).  len1 == 0: nothing cached, every line is declined.
struct ContigCache {
    const uint32_t *name4;     // name bytes + '\t', zero padded to whole words
    uint32_t len1;             // name length + 1
    int32_t  cid;              // index in the site table
    int64_t  max_pos;          // largest site position on the contig, -1 when it holds none
    int64_t  bit_base;
};

SNP_HD void contig_cache_load(const SiteTable &t, int cid, uint32_t *name4_store, uint32_t store_words, ContigCache *cc) {
    cc->name4 = name4_store; cc->len1 = 0; cc->cid = cid; cc->max_pos = -1; cc->bit_base = 0;
    if (cid < 0 || cid >= t.n_contigs) return;
    const uint32_t L = (uint32_t)t.len1[cid], nw = (L + 3u) >> 2;
    if (nw > store_words) return;
    for (uint32_t j = 0; j < nw; j++) name4_store[j] = t.names4[t.off4[cid] + j];
    cc->len1 = L; cc->max_pos = t.max_pos[cid]; cc->bit_base = t.bit_base[cid];
}

// One line starting at buf[s]; bytes at and after buf[limit] are '\n' sentinels.  cc: the contig the caller
// expects.  Returns ST_OK (out filled), ST_SKIP (not at a wanted site, filter mode, pileup.py:423-427) or
// ST_DETAIL.
SNP_HD int quick_line(const uint8_t *buf, uint32_t s, uint32_t limit, const SiteTable &sites, const ContigCache &cc,
                      const CallParams &p, bool all_positions, QuickLine *out) {
    const uint32_t H = 0x80808080u, K = 0x7f7f7f7fu;
    // ---- columns 1-4: contig, position, reference base, raw depth ---------------------------------
    if (cc.len1 == 0) return ST_DETAIL;
    {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(buf + (s & ~3u));
        const uint32_t sh = (s & 3u) * 8u, nw = cc.len1 >> 2, rem = cc.len1 & 3u;
        uint32_t lo = q[0], diff = 0, j = 0;
        for (; j < nw; j++) {
            const uint32_t hi = q[j + 1];
            diff |= funnel_r(lo, hi, sh) ^ cc.name4[j];
            lo = hi;
        }
        if (rem) diff |= (funnel_r(lo, q[j + 1], sh) ^ cc.name4[j]) & ((1u << (8u * rem)) - 1u);
        if (diff) return ST_DETAIL;
    }
    uint32_t i = s + cc.len1;
    uint32_t pos = 0;
    bool bad = !fast_digits(buf, i, pos);
    bad |= buf[i] != '\t';
    i++;
    if (bad) return ST_DETAIL;
    int32_t site = -1;
    if ((int64_t)pos <= cc.max_pos) {                  // site_find (sites.cuh) on the cached contig
        const int64_t bit = cc.bit_base + (int64_t)pos;
        const uint32_t bw = sites.bits[bit >> 5], bb = (uint32_t)bit & 31u;
        if ((bw >> bb) & 1u) site = (int32_t)(sites.rank[bit >> 5] + (uint32_t)popc32(bw & ((1u << bb) - 1u)));
    }
    if (!all_positions && site < 0) return ST_SKIP;
    const unsigned ref = buf[i];
    bad = ((ref | 0x20u) - 'a') >= 26u;                // a letter: '.'/',' stand for REF / ref (pileup.py:255-256)
    bad |= buf[i + 1] != '\t';
    i += 2;
    uint32_t raw_depth = 0;
    bad |= !fast_digits(buf, i, raw_depth);
    bad |= buf[i] != '\t';
    if (bad || raw_depth == 0) return ST_DETAIL;       // depth 0: pileup.py:226-234, left to the detailed parser
    i++;
    // ---- column 5: bases ----------------------------------------------------------------------------
    WordReader rd;
    rd.seek(buf, i);
    const uint32_t refb = (ref | 0x20u) * 0x01010101u;
    uint32_t a_rem = 0, a_dc = 0, a_dot = 0;           // 128 x (removed bytes, kept '.'/',', kept '.')
    uint32_t anomaly = 0, prevcar = 0, nfull = 0;
    uint32_t w, c;
    for (;;) {
        w = rd.next();
        c = w + 0x5f5f5f5fu;                           // bit 7 clear <-> byte < 0x21
        if (~c & H) break;
        const uint32_t in = (w + 0x55555555u) & ~(w + 0x51515151u) & H;        // + , - .
        const uint32_t s7 = w << 7;                                            // bit 0 -> bit 7: '+' and '-'
        const uint32_t car = ~((w ^ 0x5e5e5e5eu) + K) & H;                     // '^'
        const uint32_t dol = ~((w ^ 0x24242424u) + K) & H;                     // '$'
        const uint32_t part = funnel_l8(prevcar, car);                         // the byte after a '^'
        const uint32_t refm = ~(((w | 0x20202020u) ^ refb) + K) & H;           // the reference letter, either case
        anomaly |= (in & s7 & ~part) | (car & part) | refm;
        const uint32_t dck = in & ~s7 & ~part;
        a_rem = flag_sum(car | part | dol, a_rem);
        a_dc = flag_sum(dck, a_dc);
        a_dot = flag_sum(dck & (w << 6), a_dot);                               // bit 1 -> bit 7: '.' not ','
        prevcar = car;
        nfull++;
    }
    uint32_t bases_len;
    {   // the word that holds the separator: the same, restricted to the bytes in front of it
        const uint32_t low = ~c & H;
        const uint32_t first = low & (0u - low);
        const uint32_t valid = (first - 1u) & H;
        const uint32_t j = (uint32_t)ctz32(first) >> 3;
        const uint32_t in = (w + 0x55555555u) & ~(w + 0x51515151u) & valid;
        const uint32_t s7 = w << 7;
        const uint32_t car = ~((w ^ 0x5e5e5e5eu) + K) & valid;
        const uint32_t dol = ~((w ^ 0x24242424u) + K) & valid;
        const uint32_t part = funnel_l8(prevcar, car);                         // may reach the separator itself
        const uint32_t refm = ~(((w | 0x20202020u) ^ refb) + K) & valid;
        anomaly |= (in & s7 & ~part) | (car & part) | refm | (part & first);   // "^" + separator: trailing '^'
        const uint32_t dck = in & ~s7 & ~part;
        a_rem = flag_sum((car | part | dol) & valid, a_rem);
        a_dc = flag_sum(dck, a_dc);
        a_dot = flag_sum(dck & (w << 6), a_dot);
        if (((w >> (8u * j)) & 0xffu) != '\t') anomaly |= H;                   // the column ends in a tab
        bases_len = 4u * nfull + j;
    }
    if (anomaly || bases_len == 0) return ST_DETAIL;
    const uint32_t nb = bases_len - (a_rem >> 7);      // length of the stripped string
    const uint32_t q0 = i + bases_len + 1u;
    if (nb < 1u || q0 + nb > limit) return ST_DETAIL;
    // ---- column 6: as many printable bytes as bases survived, then the line end (pileup.py:248-250) ----
    rd.seek(buf, q0);
    uint32_t acc = H;
#pragma unroll 2
    for (uint32_t k = nb >> 2; k > 0; k--) acc &= rd.next() + 0x5f5f5f5fu;
    w = rd.next();
    const uint32_t r = nb & 3u;
    const uint32_t expect = 0x80u << (8u * r);
    const uint32_t low = ~(w + 0x5f5f5f5fu) & H;
    if ((acc & H) != H || (low & (expect | (expect - 1u))) != expect || ((w >> (8u * r)) & 0xffu) != '\n')
        return ST_DETAIL;
    // ---- call (pileup.py:550-588): the reference base wins outright ---------------------------------
    const uint32_t dc = a_dc >> 7, dot = a_dot >> 7;
    if (dc <= nb - dc) return ST_DETAIL;
    out->site = site;
    out->end = q0 + nb;
    out->base = (uint8_t)ref;
    out->fail = filter_mask(nb, dc, dot, dc - dot, p);
    return ST_OK;
}

}  // namespace snpgpu
