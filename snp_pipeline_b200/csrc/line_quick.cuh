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
// skips indel tokens of up to 999 bases in a second look at the column (only lines that showed a '+' / '-'),
// and returns ST_DETAIL for anything else: another winner possible, a "^^" chain, a trailing '^', a malformed
// or longer indel token, an odd separator, "\r\n", depth 0, a contig other than the hinted one ...  ST_DETAIL
// has no side effects; the caller hands the line to line_fast.cuh (exact tallies) and, from there, to line_general.cuh.
//
// Precondition (the kernel checks it per tile): no byte >= 0x80 in the staged window, so byte lanes never
// carry into each other.  buf is 4-byte aligned with '\n' sentinels behind `limit`.
#pragma once
#include <type_traits>
#include "line_fast.cuh"

namespace snpgpu {

enum : int { ST_DETAIL = 67 };

struct QuickLine {
    int32_t  site;         // unique-site index or -1
    uint32_t end;          // offset of the line terminator in buf
    uint8_t  base;         // consensus character before the '-' substitutions of call_consensus.py:169-176
    uint8_t  fail;         // FAIL_* mask (without FAIL_REGION)
    uint8_t  flags;        // SITE_SNP / SITE_EXCLUDED of the site
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
// (name and byte masks in shared memory, the rest in registers).  len1 == 0: nothing cached, every line is declined.
constexpr uint32_t QUICK_PAD = 32;   // '\n' sentinels the caller keeps behind `limit` (word over-reads land there)

struct ContigCache {
    const uint32_t *name4;     // name bytes + '\t', zero padded to blocks of four words (16-byte aligned)
    const uint32_t *mask4;     // per word: 0xff in the bytes that belong to name + tab
    uint32_t len1;             // name length + 1
    uint32_t nblk;             // four-word blocks that cover name + tab
    int32_t  cid;              // index in the site table
    int64_t  max_pos;          // largest site position on the contig, -1 when it holds none
    int64_t  bit_base;
};

// name4_store / mask4_store: store_words words each (a multiple of 4), 16-byte aligned
SNP_HD void contig_cache_load(const SiteTable &t, int cid, uint32_t *name4_store, uint32_t *mask4_store,
                              uint32_t store_words, ContigCache *cc) {
    cc->name4 = name4_store; cc->mask4 = mask4_store; cc->len1 = 0; cc->nblk = 0; cc->cid = cid; cc->max_pos = -1;
    cc->bit_base = 0;
    if (cid < 0 || cid >= t.n_contigs) return;
    const uint32_t L = (uint32_t)t.len1[cid], nw = (L + 3u) >> 2;
    if (nw > store_words) return;
    for (uint32_t j = 0; j < store_words; j++) {
        name4_store[j] = j < nw ? t.names4[t.off4[cid] + j] : 0u;
        mask4_store[j] = 4u * j + 4u <= L ? 0xffffffffu : (4u * j < L ? (1u << (8u * (L - 4u * j))) - 1u : 0u);
    }
    cc->len1 = L; cc->nblk = (nw + 3u) >> 2; cc->max_pos = t.max_pos[cid]; cc->bit_base = t.bit_base[cid];
}

// the register half of contig_cache_load: for a second warp that shares the first one's name4 / mask4 words
SNP_HD void contig_cache_attach(const SiteTable &t, int cid, const uint32_t *name4_store, const uint32_t *mask4_store,
                                uint32_t store_words, ContigCache *cc) {
    cc->name4 = name4_store; cc->mask4 = mask4_store; cc->len1 = 0; cc->nblk = 0; cc->cid = cid; cc->max_pos = -1;
    cc->bit_base = 0;
    if (cid < 0 || cid >= t.n_contigs) return;
    const uint32_t L = (uint32_t)t.len1[cid], nw = (L + 3u) >> 2;
    if (nw > store_words) return;
    cc->len1 = L; cc->nblk = (nw + 3u) >> 2; cc->max_pos = t.max_pos[cid]; cc->bit_base = t.bit_base[cid];
}

SNP_HD SiteWord load_site_word(const SiteWord *p) {
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    return SiteWord{v.x, v.y, v.z, v.w};
#else
    return *p;
#endif
}

struct Word4 { uint32_t x, y, z, w; };
SNP_HD Word4 load_word4(const uint32_t *p) {                  // p 16-byte aligned
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    return Word4{v.x, v.y, v.z, v.w};
#else
    return Word4{p[0], p[1], p[2], p[3]};
#endif
}

// value of four decimal digit bytes (0..9 each, most significant in the lowest byte)
SNP_HD uint32_t digits4_value(uint32_t w) {
    const uint32_t p = (w * 10u + (w >> 8)) & 0x00ff00ffu;    // byte 0: d0 d1, byte 2: d2 d3
    return (p & 0xffffu) * 100u + (p >> 16);
}

// 1..7 decimal digits at buf[i] (eight or more: fast_digits() takes over); leaves i on the first non-digit.
// All bytes < 0x80.  Reads the 12 bytes from buf[i & ~3].
SNP_HD bool quick_digits8(const uint8_t *buf, uint32_t &i, uint32_t &v) {
    const uint32_t H = 0x80808080u;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(buf + (i & ~3u));
    const uint32_t sh = (i & 3u) * 8u;
    const uint32_t a = p[0], b = p[1], c = p[2];
    const uint32_t lo = funnel_r(a, b, sh) ^ 0x30303030u, hi = funnel_r(b, c, sh) ^ 0x30303030u;   // digits -> 0..9
    const uint32_t ndl = (lo + 0x76767676u) & H, ndh = (hi + 0x76767676u) & H;                       // bit 7: not a digit
    if ((ndl | ndh) == 0u) return fast_digits(buf, i, v);
    const bool in_lo = ndl != 0u;
    const uint32_t n = in_lo ? (uint32_t)ctz32(ndl) >> 3 : 4u + ((uint32_t)ctz32(ndh) >> 3);
    const uint32_t s8 = (n & 3u) * 8u;
    // the digits right-aligned in eight bytes: x_hi = bytes n-4 .. n-1, x_lo = bytes n-8 .. n-5 (zero in front of byte 0)
    const uint32_t x_hi = funnel_r(in_lo ? 0u : lo, in_lo ? lo : hi, s8);
    const uint32_t x_lo = in_lo ? 0u : funnel_r(0u, lo, s8);
    v = digits4_value(x_lo) * 10000u + digits4_value(x_hi);
    i += n;
    return n > 0u;
}

// the same for 1..3 digits (four or more: fast_digits()); reads the 8 bytes from buf[i & ~3]
SNP_HD bool quick_digits4(const uint8_t *buf, uint32_t &i, uint32_t &v) {
    const uint32_t w = load_u32(buf, i) ^ 0x30303030u;
    const uint32_t nd = (w + 0x76767676u) & 0x80808080u;
    if (nd == 0u) return fast_digits(buf, i, v);
    const uint32_t n = (uint32_t)ctz32(nd) >> 3;
    v = digits4_value(funnel_r(0u, w, n * 8u));
    i += n;
    return n > 0u;
}

// One line starting at buf[s]; bytes at and after buf[limit] are '\n' sentinels.  cc: the contig the caller
// expects.  Returns ST_OK (out filled), ST_SKIP (not at a wanted site, filter mode, pileup.py:423-427) or
// ST_DETAIL.
SNP_HD int quick_line(const uint8_t *buf, uint32_t s, uint32_t limit, const SiteTable &sites, const ContigCache &cc,
                      const CallParams &p, bool all_positions, QuickLine *out) {
    const uint32_t H = 0x80808080u, K = 0x7f7f7f7fu;
    // ---- columns 1-4: contig, position, reference base, raw depth ---------------------------------
    if (cc.len1 == 0 || s + 16u * cc.nblk + 8u > limit + QUICK_PAD) return ST_DETAIL;
    {   // name + tab, four words at a time; bytes behind them are masked out
        const uint32_t *q = reinterpret_cast<const uint32_t *>(buf + (s & ~3u));
        const uint32_t sh = (s & 3u) * 8u;
        uint32_t lo = q[0], diff = 0;
        for (uint32_t b = 0; b < cc.nblk; b++) {
            const uint32_t h0 = q[4u * b + 1u], h1 = q[4u * b + 2u], h2 = q[4u * b + 3u], h3 = q[4u * b + 4u];
            const Word4 nm = load_word4(cc.name4 + 4u * b), mk = load_word4(cc.mask4 + 4u * b);
            diff |= ((funnel_r(lo, h0, sh) ^ nm.x) & mk.x) | ((funnel_r(h0, h1, sh) ^ nm.y) & mk.y);
            diff |= ((funnel_r(h1, h2, sh) ^ nm.z) & mk.z) | ((funnel_r(h2, h3, sh) ^ nm.w) & mk.w);
            lo = h3;
        }
        if (diff) return ST_DETAIL;
    }
    uint32_t i = s + cc.len1;
    uint32_t pos = 0;
    bool bad = !quick_digits8(buf, i, pos);
    bad |= buf[i] != '\t';
    i++;
    if (bad) return ST_DETAIL;
    // site_find (sites.cuh) on the cached contig: the packed word of the position's 32-position group is asked for here
    // (one 16-byte load: site bits, snplist / exclude bits, rank); in all-positions mode it is only looked at when the
    // line is done, so its latency hides behind the rest of the parse and nothing else has to be loaded for the site
    const int64_t bit = cc.bit_base + (int64_t)pos;
    const uint32_t bb = (uint32_t)bit & 31u;
    SiteWord sw{0u, 0u, 0u, 0u};
    if (all_positions) {
        if ((int64_t)pos <= cc.max_pos) sw = load_site_word(sites.words + (bit >> 5));
    } else {                                           // filter mode: the plain bitmap decides, nearly always "skip"
        const uint32_t bw = (int64_t)pos <= cc.max_pos ? sites.bits[bit >> 5] : 0u;
        if (!((bw >> bb) & 1u)) return ST_SKIP;
        sw = load_site_word(sites.words + (bit >> 5));
    }
    const unsigned ref = buf[i];
    bad = ((ref | 0x20u) - 'a') >= 26u;                // a letter: '.'/',' stand for REF / ref (pileup.py:255-256)
    bad |= buf[i + 1] != '\t';
    i += 2;
    uint32_t raw_depth = 0;
    bad |= !quick_digits4(buf, i, raw_depth);
    bad |= buf[i] != '\t';
    if (bad || raw_depth == 0) return ST_DETAIL;       // depth 0: pileup.py:226-234, left to the detailed parser
    i++;
    // ---- column 5: bases ----------------------------------------------------------------------------
    WordReader rd;
    rd.seek(buf, i);
    const uint32_t refb = (ref | 0x20u) * 0x01010101u;
    uint32_t a_rem, a_dc, a_dot;                       // 128 x (removed bytes, kept '.'/',', kept '.')
    uint32_t anomaly = 0, signs = 0, prevcar = 0;
    uint32_t wpos;                                     // offset of the word being looked at
    uint32_t bases_len = 0;
    // An indel token [+-]<n><n letters> (pileup.py:315-320) whose sign is the lowest flag of `sign` in the word at wpos:
    // the bytes in front of it are counted like any others, the token is skipped byte-wise (1..3 digits, a sequence of
    // letters / '*' as line_fast.cuh takes it) and the word reader starts again behind it.  false: not for this tier.
    auto skip_indel = [&](uint32_t w, uint32_t sign, uint32_t in, uint32_t s7, uint32_t car, uint32_t dol, uint32_t part,
                          uint32_t refm) -> bool {
        const uint32_t fs = sign & (0u - sign);
        const uint32_t v2 = (fs - 1u) & H;                                     // the bytes in front of the sign
        const uint32_t car2 = car & v2, part2 = part & v2, dck2 = in & ~s7 & ~part & v2;
        anomaly |= (car2 & part2) | (refm & v2);
        a_rem = flag_sum(car2 | part2 | (dol & v2), a_rem);
        a_dc = flag_sum(dck2, a_dc);
        a_dot = flag_sum(dck2 & (w << 6), a_dot);
        uint32_t k = wpos + ((uint32_t)ctz32(fs) >> 3) + 1u, n = 0, nd = 0;
        while (nd < 3u && (unsigned)buf[k] - '0' < 10u) { n = n * 10u + ((unsigned)buf[k] - '0'); k++; nd++; }
        if (nd == 0u || (unsigned)buf[k] - '0' < 10u || k + n > limit) return false;   // a bare sign, a long number
        for (uint32_t x = 0; x < n; x++) {
            const unsigned ch = buf[k + x];
            if ((ch | 0x20u) - 'a' >= 26u && ch != '*') return false;
        }
        a_rem += 128u * (1u + nd + n);
        wpos = k + n;
        rd.seek(buf, wpos);
        prevcar = 0;
        return true;
    };
    // One look at the column.  The first look only notes '+' / '-' (no branch per word: the loop is latency-bound); a
    // line that showed one and nothing else odd is looked at again with the tokens skipped.  false: not for this tier.
    auto look = [&](auto with_indels) -> bool {
        constexpr bool INDEL = decltype(with_indels)::value;
        rd.seek(buf, i);
        a_rem = a_dc = a_dot = 0;
        anomaly = signs = prevcar = 0;
        wpos = i;
        uint32_t w, c;
        for (;;) {
            w = rd.next();
            c = w + 0x5f5f5f5fu;                       // bit 7 clear <-> byte < 0x21
            if (~c & H) break;
            const uint32_t in = (w + 0x55555555u) & ~(w + 0x51515151u) & H;    // + , - .
            const uint32_t s7 = w << 7;                                        // bit 0 -> bit 7: '+' and '-'
            const uint32_t car = ~((w ^ 0x5e5e5e5eu) + K) & H;                 // '^'
            const uint32_t dol = ~((w ^ 0x24242424u) + K) & H;                 // '$'
            const uint32_t part = funnel_l8(prevcar, car);                     // the byte after a '^'
            const uint32_t refm = ~(((w | 0x20202020u) ^ refb) + K) & H;       // the reference letter, either case
            const uint32_t sign = in & s7 & ~part;                             // '+' / '-' that is not a '^' partner
            if constexpr (INDEL) {
                if (sign) {
                    if (!skip_indel(w, sign, in, s7, car, dol, part, refm)) return false;
                    continue;
                }
            } else {
                signs |= sign;
            }
            anomaly |= (car & part) | refm;
            const uint32_t dck = in & ~s7 & ~part;
            a_rem = flag_sum(car | part | dol, a_rem);
            a_dc = flag_sum(dck, a_dc);
            a_dot = flag_sum(dck & (w << 6), a_dot);                           // bit 1 -> bit 7: '.' not ','
            prevcar = car;
            wpos += 4u;
        }
        // the word that holds the separator: the same, restricted to the bytes in front of it
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
        const uint32_t sign = in & s7 & ~part;
        // (with tokens skipped, a sign here is the 1-in-4 case "+1A" + separator in one word, or malformed: next tier)
        if constexpr (INDEL) anomaly |= sign; else signs |= sign;
        anomaly |= (car & part) | refm | (part & first);                       // "^" + separator: trailing '^'
        const uint32_t dck = in & ~s7 & ~part;
        a_rem = flag_sum((car | part | dol) & valid, a_rem);
        a_dc = flag_sum(dck, a_dc);
        a_dot = flag_sum(dck & (w << 6), a_dot);
        if (((w >> (8u * j)) & 0xffu) != '\t') anomaly |= H;                   // the column ends in a tab
        bases_len = wpos + j - i;
        return true;
    };
    look(std::false_type{});
    if (signs && !look(std::true_type{})) return ST_DETAIL;   // (what else the first look flagged may lie in a token)
    if (anomaly || bases_len == 0) return ST_DETAIL;
    const uint32_t nb = bases_len - (a_rem >> 7);      // length of the stripped string
    const uint32_t q0 = i + bases_len + 1u;
    if (nb < 1u || q0 + nb > limit) return ST_DETAIL;
    // ---- column 6: as many printable bytes as bases survived, then the line end (pileup.py:248-250) ----
    rd.seek(buf, q0);
    uint32_t acc = H;
#pragma unroll 2
    for (uint32_t k = nb >> 2; k > 0; k--) acc &= rd.next() + 0x5f5f5f5fu;
    const uint32_t w = rd.next();
    const uint32_t r = nb & 3u;
    const uint32_t expect = 0x80u << (8u * r);
    const uint32_t low = ~(w + 0x5f5f5f5fu) & H;
    if ((acc & H) != H || (low & (expect | (expect - 1u))) != expect || ((w >> (8u * r)) & 0xffu) != '\n')
        return ST_DETAIL;
    // ---- call (pileup.py:550-588): the reference base wins outright ---------------------------------
    const uint32_t dc = a_dc >> 7, dot = a_dot >> 7;
    if (dc <= nb - dc) return ST_DETAIL;
    out->site = ((sw.any >> bb) & 1u) ? (int32_t)(sw.rank + (uint32_t)popc32(sw.any & ((1u << bb) - 1u))) : -1;
    out->flags = (uint8_t)((((sw.snp >> bb) & 1u) ? SITE_SNP : 0u) | (((sw.exc >> bb) & 1u) ? SITE_EXCLUDED : 0u));
    out->end = q0 + nb;
    out->base = (uint8_t)ref;
    out->fail = filter_mask(nb, dc, dot, dc - dot, p);
    return ST_OK;
}

}  // namespace snpgpu
