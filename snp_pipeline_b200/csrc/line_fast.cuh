// line_fast.cuh -- the fast parser for one pileup line of the shape samtools mpileup writes:
//
//     <known contig> \t <digits> \t <one of ACGTNacgtn> \t <digits> \t <bases> \t <quals> [\r]
//
// with <bases> over the alphabet  . , A C G T N a c g t n *  plus the markers  ^x  $  and  [+-]<n><n letters>.
// Every byte of the line is validated; the moment anything else shows up (another separator, an odd symbol,
// a short quality string, an indel token of unusual shape, a contig that is not in the site table ...) the
// function returns ST_FALLBACK without side effects and the caller hands the line to line_general.cuh,
// which is exact for any input.  For lines it accepts, the result is what the reference computes:
//   pileup.py:209-274   Record._init_from_split_line   (tallies, strand split, most common base)
//   pileup.py:276-325   Record._strip_unwanted_base_patterns
//   pileup.py:492-590   ConsensusCaller.call_consensus
// One thread walks one line.  The text sits in a 4-byte aligned staging buffer (shared memory on the device)
// and is fetched as aligned 32-bit words recombined with a funnel shift -- four bases per load.
#pragma once
#include "hd.cuh"
#include "sites.cuh"

namespace snpgpu {

struct FastLine {
    int64_t  pos;
    int32_t  cid;          // contig index in the site table
    int32_t  site;         // unique-site index or -1
    uint8_t  ref;
    uint8_t  base;         // consensus character before the '-' substitutions of call_consensus.py:169-176
    uint8_t  fail;         // FAIL_* mask (without FAIL_REGION)
};

// Does buf[s ..] start with contig entry c (name + tab)?
SNP_HD bool match_contig(const uint8_t *buf, uint32_t s, const SiteTable &t, int c) {
    const uint32_t *nm = t.names4 + t.off4[c];
    uint32_t L = (uint32_t)t.len1[c];
    const uint32_t *p = reinterpret_cast<const uint32_t *>(buf + (s & ~3u));
    uint32_t sh = (s & 3u) * 8u;
    uint32_t lo = p[0];
    uint32_t diff = 0;
    uint32_t nw = L >> 2, j = 0;
    for (; j < nw; j++) {
        uint32_t hi = p[j + 1];
        diff |= funnel_r(lo, hi, sh) ^ nm[j];
        lo = hi;
    }
    uint32_t rem = L & 3u;
    if (rem) diff |= (funnel_r(lo, p[j + 1], sh) ^ nm[j]) & ((1u << (8u * rem)) - 1u);
    return diff == 0;
}

// 1..9 decimal digits at buf[i]; leaves i on the first non-digit
SNP_HD bool fast_digits(const uint8_t *buf, uint32_t &i, uint32_t &v) {
    uint32_t x = 0;
    int nd = 0;
    for (;;) {
        unsigned d = (unsigned)buf[i] - '0';
        if (d > 9u) break;
        if (++nd > 9) return false;
        x = x * 10u + d;
        i++;
    }
    v = x;
    return nd > 0;
}

// tallies over the fast alphabet; letters in byte order A C G N T
struct FastTally {
    uint32_t dot, comma, star;
    uint32_t f[5], r[5];
};

// One surviving base (pileup.py:255-274).  false = not in the fast alphabet.
SNP_HD bool tally_symbol(unsigned c, FastTally &t) {
    switch (c) {
        case 'A': t.f[0]++; return true;
        case 'C': t.f[1]++; return true;
        case 'G': t.f[2]++; return true;
        case 'N': t.f[3]++; return true;
        case 'T': t.f[4]++; return true;
        case 'a': t.r[0]++; return true;
        case 'c': t.r[1]++; return true;
        case 'g': t.r[2]++; return true;
        case 'n': t.r[3]++; return true;
        case 't': t.r[4]++; return true;
        case '*': t.star++; return true;
        default: return false;
    }
}

// Are all n bytes at buf[i ..] printable ASCII (0x21..0x7f)?  (quality column: no separator, nothing that
// Python would decode or split differently, nothing below phred 0)
SNP_HD bool all_printable(const uint8_t *buf, uint32_t i, uint32_t n) {
    uint32_t bad = 0;
    while (n >= 4u) {
        uint32_t w = load_u32(buf, i);
        bad |= ~(((w & 0x7f7f7f7fu) + 0x5f5f5f5fu) & ~w);
        i += 4u; n -= 4u;
    }
    bad &= 0x80808080u;
    if (n) {
        uint32_t w = load_u32(buf, i);
        uint32_t m = (0x80808080u >> (8u * (4u - n)));
        bad |= ~(((w & 0x7f7f7f7fu) + 0x5f5f5f5fu) & ~w) & m;
    }
    return bad == 0;
}

// first tab at or after buf[i], or limit if none before limit
SNP_HD uint32_t find_tab(const uint8_t *buf, uint32_t i, uint32_t limit) {
    while (i < limit) {
        uint32_t t = load_u32(buf, i) ^ 0x09090909u;
        uint32_t z = (t - 0x01010101u) & ~t & 0x80808080u;
        if (z) {
            uint32_t p = i + ((uint32_t)ctz32(z) >> 3);
            return p < limit ? p : limit;
        }
        i += 4u;
    }
    return limit;
}

// The line occupies buf[s, e): e is the position of its terminator ('\n', or the end of the text, where the
// caller keeps a '\n' sentinel).  all_positions: parse whatever the position (--vcfAllPos, pileup.py:419-421);
// otherwise lines away from the site table return ST_SKIP after the key columns (pileup.py:423-427).
// hint: the contig the previous line of this thread matched (in/out).
template <bool HAS_QUAL>
SNP_HD int fast_line(const uint8_t *buf, uint32_t s, uint32_t e, const SiteTable &sites, int &hint,
                     const CallParams &p, bool all_positions, FastLine *out) {
    if (e > s && buf[e - 1] == '\r') e--;            // "\r\n": rstrip() drops the '\r' (pileup.py:206)
    // ---- column 1: contig ------------------------------------------------------------------------
    int cid = hint;
    if (sites.n_contigs == 0) return ST_FALLBACK;
    if (!match_contig(buf, s, sites, cid)) {
        cid = -1;
        for (int c = 0; c < sites.n_contigs; c++) {
            if (c != hint && match_contig(buf, s, sites, c)) { cid = c; break; }
        }
        if (cid < 0) return ST_FALLBACK;
        hint = cid;
    }
    uint32_t i = s + (uint32_t)sites.len1[cid];
    // ---- column 2: position ----------------------------------------------------------------------
    uint32_t pos;
    if (!fast_digits(buf, i, pos)) return ST_FALLBACK;
    if (buf[i] != '\t') return ST_FALLBACK;
    i++;
    int32_t site = site_find(sites, cid, (int64_t)pos);
    if (!all_positions && site < 0) return ST_SKIP;
    // ---- column 3: reference base ------------------------------------------------------------------
    unsigned ref = buf[i];
    unsigned U = ref & 0xdfu;
    int ui;                                            // index of REF.upper() in A C G N T
    switch (U) {
        case 'A': ui = 0; break;
        case 'C': ui = 1; break;
        case 'G': ui = 2; break;
        case 'N': ui = 3; break;
        case 'T': ui = 4; break;
        default: return ST_FALLBACK;
    }
    if ((ref | 0x20u) - 'a' >= 26u) return ST_FALLBACK;
    if (buf[i + 1] != '\t') return ST_FALLBACK;
    i += 2;
    // ---- column 4: raw depth -----------------------------------------------------------------------
    uint32_t raw_depth;
    if (!fast_digits(buf, i, raw_depth)) return ST_FALLBACK;
    out->pos = (int64_t)pos; out->cid = cid; out->site = site; out->ref = (uint8_t)ref;
    if (raw_depth == 0) {                              // pileup.py:226-234 -> ('-', RawDpth)
        if (buf[i] != '\t' && i != e) return ST_FALLBACK;
        out->base = '-'; out->fail = FAIL_RAWDPTH;
        return ST_OK;
    }
    if (buf[i] != '\t') return ST_FALLBACK;
    i++;
    // ---- column 5 (bases) against column 6 (qualities) ---------------------------------------------
    uint32_t qs = 0, nq = 0;
    int thr = 33 + p.min_base_qual;
    if (HAS_QUAL) {
        uint32_t tab = find_tab(buf, i, e);
        if (tab >= e) return ST_FALLBACK;
        qs = tab + 1;
        nq = e - qs;
        if (nq < 1 || !all_printable(buf, qs, nq)) return ST_FALLBACK;
    }
    FastTally t;
    t.dot = t.comma = t.star = 0;
    for (int k = 0; k < 5; k++) { t.f[k] = 0; t.r[k] = 0; }
    uint32_t nb = 0;                                   // length of the stripped string so far
    for (;;) {
        uint32_t w = load_u32(buf, i);
        uint32_t consumed = 4;
        int action = 0;                                // 1 = tab reached, 2 = indel token at i + consumed
        bool skip = false;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            unsigned c = (w >> (8 * j)) & 0xffu;
            if (skip) {                                // the byte after '^' (pileup.py:312)
                if (c - 0x21u > 0x5du) return ST_FALLBACK;
                skip = false;
                continue;
            }
            if (c != '.' && c != ',') {
                if (c == '\t') { consumed = j; action = 1; break; }
                if (c == '$') continue;
                if (c == '^') { skip = true; continue; }
                if (c == '+' || c == '-') { consumed = j; action = 2; break; }
            }
            // c survives the strip: it pairs with quality byte nb (pileup.py:248-250)
            bool good = true;
            if (HAS_QUAL) {
                if (nb >= nq) return ST_FALLBACK;      // zip() would truncate
                good = (int)buf[qs + nb] >= thr;
            }
            nb++;
            if (c == '.') { if (good) t.dot++; }
            else if (c == ',') { if (good) t.comma++; }
            else if (good) { if (!tally_symbol(c, t)) return ST_FALLBACK; }
            else { FastTally scratch = t; if (!tally_symbol(c, scratch)) return ST_FALLBACK; }
        }
        i += consumed;
        if (action == 1) break;
        if (action == 2) {
            // [+-]<digits><that many letters>  (pileup.py:315-320), the shape samtools writes
            uint32_t k = i + 1, n;
            if (!fast_digits(buf, k, n) || n > 4096u) return ST_FALLBACK;
            for (uint32_t x = 0; x < n; x++) {
                unsigned c = buf[k + x];
                if ((c | 0x20u) - 'a' >= 26u && c != '*') return ST_FALLBACK;
            }
            i = k + n;
        } else if (skip) {                             // '^' was the 4th byte: its partner starts the next word
            unsigned c = buf[i];
            if (c - 0x21u > 0x5du) return ST_FALLBACK;
            i++;
        }
    }
    if (i >= e) return ST_FALLBACK;                    // the tab must lie inside the line
    if (!HAS_QUAL) {
        qs = i + 1;
        nq = e - qs;
        if (nq < 1 || !all_printable(buf, qs, nq)) return ST_FALLBACK;
    }
    if (nb != nq) return ST_FALLBACK;                  // zip() truncation -> general path
    // ---- rank and call (pileup.py:259-266, 550-588) ------------------------------------------------
    uint32_t f[5], r[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        f[k] = t.f[k] + (k == ui ? t.dot : 0u);
        r[k] = t.r[k] + (k == ui ? t.comma : 0u);
    }
    uint32_t good = t.star;
#pragma unroll
    for (int k = 0; k < 5; k++) good += f[k] + r[k];
    if (good < 1) { out->base = '-'; out->fail = FAIL_RAWDPTH; return ST_OK; }
    // candidates in byte order  * A C G N T ; strict '>' keeps the smallest byte among equals
    uint32_t best = t.star, bf = t.star, br = 0;
    unsigned wsym = '*';
    const unsigned sym[5] = {'A', 'C', 'G', 'N', 'T'};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        uint32_t tot = f[k] + r[k];
        if (tot > best) { best = tot; bf = f[k]; br = r[k]; wsym = sym[k]; }
    }
    out->fail = filter_mask(good, best, bf, br, p);
    out->base = (uint8_t)((wsym == U) ? ref : wsym);
    return ST_OK;
}

}  // namespace snpgpu
