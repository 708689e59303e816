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

// ---- tallies of one line ---------------------------------------------------------------------------------
// '.' and ',' have their own counters.  The first other letter of the line (normally the alternate allele) is
// counted in place by strand; the few remaining symbols are queued, eight to a 64-bit register, and tallied at
// the end of the line -- so the per-byte loop stays free of symbol-dependent branches and the lanes of a warp,
// each on its own line, run it in lock step.
struct LineTally {
    uint32_t dot, comma;          // '.' / ','  (pileup.py:255-256 turns them into REF / ref)
    uint32_t osym;                // upper-case letter counted in place (0 = none yet)
    uint32_t ofwd, orev;
    uint64_t pend;                // queued symbols, one byte each
    uint32_t npend;
    FastTally t;                  // everything that went through the queue
    bool bad;
};

SNP_HD void tally_flush(LineTally &L) {
    while (L.npend) {
        unsigned c = (unsigned)(L.pend & 0xffu);
        L.pend >>= 8;
        L.npend--;
        if (!tally_symbol(c, L.t)) L.bad = true;
    }
}

// c survives the strip and passes the quality gate: count it
SNP_HD void tally_other(unsigned c, LineTally &L) {
    const unsigned u = c & 0xdfu;
    const unsigned v = u - 'A';
    const bool acgtn = v < 20u && ((0x82045u >> v) & 1u) && ((c | 0x20u) - 'a' < 26u);
    if (acgtn && (L.osym == 0u || L.osym == u)) {
        L.osym = u;
        if (c & 0x20u) L.orev++; else L.ofwd++;
    } else {
        if (L.npend == 8u) tally_flush(L);
        L.pend |= (uint64_t)c << (8u * L.npend);
        L.npend++;
    }
}

// The line occupies buf[s, e): e is the position of its terminator ('\n', or the end of the text, where the
// caller keeps a '\n' sentinel).  all_positions: parse whatever the position (--vcfAllPos, pileup.py:419-421);
// otherwise lines away from the site table return ST_SKIP after the key columns (pileup.py:423-427).
// hint: the contig the previous line of this thread matched (in/out).
// Control flow is kept single-exit on purpose: on the GPU the 32 lanes of a warp each walk their own line, and
// they only stay converged if nobody leaves a loop early.
template <bool HAS_QUAL>
SNP_HD int fast_line(const uint8_t *buf, uint32_t s, uint32_t e, const SiteTable &sites, int &hint,
                     const CallParams &p, bool all_positions, FastLine *out) {
    if (e > s && buf[e - 1] == '\r') e--;            // "\r\n": rstrip() drops the '\r' (pileup.py:206)
    // ---- column 1: contig ------------------------------------------------------------------------
    if (sites.n_contigs == 0) return ST_FALLBACK;
    int cid = hint;
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
    uint32_t pos = 0;
    bool bad = !fast_digits(buf, i, pos);
    bad |= buf[i] != '\t';
    i++;
    if (bad) return ST_FALLBACK;
    const int32_t site = site_find(sites, cid, (int64_t)pos);
    if (!all_positions && site < 0) return ST_SKIP;
    // ---- column 3: reference base ------------------------------------------------------------------
    const unsigned ref = buf[i];
    const unsigned U = ref & 0xdfu;
    const unsigned rv = U - 'A';
    bad |= !(rv < 20u && ((0x82045u >> rv) & 1u) && ((ref | 0x20u) - 'a' < 26u));     // A C G N T, either case
    const int ui = U == 'A' ? 0 : U == 'C' ? 1 : U == 'G' ? 2 : U == 'N' ? 3 : 4;      // index in A C G N T
    bad |= buf[i + 1] != '\t';
    i += 2;
    // ---- column 4: raw depth -----------------------------------------------------------------------
    uint32_t raw_depth = 0;
    bad |= !fast_digits(buf, i, raw_depth);
    if (bad) return ST_FALLBACK;
    out->pos = (int64_t)pos; out->cid = cid; out->site = site; out->ref = (uint8_t)ref;
    if (raw_depth == 0) {                              // pileup.py:226-234 -> ('-', RawDpth)
        if (buf[i] != '\t' && i != e) return ST_FALLBACK;
        for (uint32_t k = i; k < e; k++)                // the rest is ignored, but a lone CR would end the line
            if (buf[k] == '\r') return ST_FALLBACK;
        out->base = '-'; out->fail = FAIL_RAWDPTH;
        return ST_OK;
    }
    if (buf[i] != '\t') return ST_FALLBACK;
    i++;
    // ---- column 5 (bases) against column 6 (qualities) ---------------------------------------------
    uint32_t qs = 0, nq = 0;
    const int thr = 33 + p.min_base_qual;
    if (HAS_QUAL) {
        const uint32_t tab = find_tab(buf, i, e);
        if (tab >= e) return ST_FALLBACK;
        qs = tab + 1;
        nq = e - qs;
        if (nq < 1 || !all_printable(buf, qs, nq)) return ST_FALLBACK;
    }
    LineTally L;
    L.dot = L.comma = 0; L.osym = 0; L.ofwd = L.orev = 0; L.pend = 0; L.npend = 0; L.bad = false;
    L.t.dot = L.t.comma = L.t.star = 0;
    for (int k = 0; k < 5; k++) { L.t.f[k] = 0; L.t.r[k] = 0; }
    uint32_t nb = 0;                                   // length of the stripped string so far
    bool skip = false;                                 // the next byte is the partner of a '^' (pileup.py:312)
    bool done = false;
    while (!done) {
        const uint32_t w = load_u32(buf, i);
        bool bytewise = HAS_QUAL;
        if (!HAS_QUAL) {
            // ---- four bases at a time, SWAR (every byte of the tile is < 0x80: the kernel sends tiles with
            //      high bytes to the exact path, so byte lanes never carry into each other) -----------------
            const uint32_t H = 0x80808080u, K = 0x7f7f7f7fu;
            const uint32_t low = ~(w + 0x5f5f5f5fu) & H;                          // < 0x21: tab, terminator, odd
            const uint32_t dc = ~(((w ^ 0x2c2c2c2cu) & 0xfdfdfdfdu) + K) & H;     // ',' or '.'
            const uint32_t dotm = dc & (w << 6);                                  // '.' has bit 1 set, ',' has not
            uint32_t car = ~((w ^ 0x5e5e5e5eu) + K) & H;                          // '^'
            const uint32_t dol = ~((w ^ 0x24242424u) + K) & H;                    // '$'
            const uint32_t sgn = ~(((w ^ 0x2b2b2b2bu) & 0xf9f9f9f9u) + K) & H;    // '+' '-' (also ')' '/')
            const uint32_t stopm = low | sgn;
            const uint32_t first = stopm & (0u - stopm);                          // bit 7 of the first stop byte
            const uint32_t valid = (first - 1u) & H;                              // bytes before it (all if none)
            car &= valid;
            const uint32_t part0 = skip ? 0x80u : 0u;                             // byte 0 is the partner of a '^'
            uint32_t t = car & ~part0;                                            // true carets: resolve the chain
            t = car & ~(part0 | (t << 8));
            t = car & ~(part0 | (t << 8));
            t = car & ~(part0 | (t << 8));
            const uint32_t part = part0 | (t << 8);
            if (first & sgn) {
                bytewise = true;                                                  // indel token (or a "^+"): rare
            } else {
                if (part & first) L.bad = true;                                   // "^" + separator: trailing lone '^'
                const uint32_t kept = valid & ~(t | part | dol);
                nb += (uint32_t)popc32(kept);
                L.dot += (uint32_t)popc32(dotm & kept);
                L.comma += (uint32_t)popc32((dc ^ dotm) & kept);
                uint32_t oth = kept & ~dc;
                while (oth) {
                    const uint32_t bit = (uint32_t)ctz32(oth);
                    oth &= oth - 1u;
                    tally_other((w >> (bit - 7u)) & 0xffu, L);
                }
                if (stopm) {
                    i += (uint32_t)ctz32(first) >> 3;
                    done = true;                                                  // buf[i] is checked to be the tab below
                    skip = false;
                } else {
                    i += 4u;
                    skip = (t >> 31) != 0u;
                }
            }
        }
        if (bytewise) {
            uint32_t adv = 4;
            bool stop = false, indel = false;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned c = (w >> (8 * j)) & 0xffu;
                const bool live = !stop;
                const bool partner = live && skip;
                const bool ex = live && !skip;         // an ordinary position of the bases string
                if (partner) { L.bad |= c < 0x21u; skip = false; }
                const bool is_tab = ex && c == '\t';
                const bool is_sign = ex && (c == '+' || c == '-');
                const bool is_caret = ex && c == '^';
                if (is_tab || is_sign) { stop = true; adv = (uint32_t)j; indel = is_sign; done = is_tab; }
                if (is_caret) skip = true;
                const bool keep = ex && !(is_tab || is_sign || is_caret || c == '$');
                if (ex && c < 0x21u && c != '\t') { L.bad = true; }   // a terminator or odd separator: not ours
                if (keep) {
                    bool good = true;
                    if (HAS_QUAL) {                    // pairs with quality byte nb (pileup.py:248-250)
                        if (nb >= nq) { L.bad = true; good = false; } // zip() would truncate
                        else good = (int)buf[qs + nb] >= thr;
                    }
                    nb++;
                    if (c == '.') L.dot += good ? 1u : 0u;
                    else if (c == ',') L.comma += good ? 1u : 0u;
                    else if (good) tally_other(c, L);
                    else { FastTally scratch = L.t; if (!tally_symbol(c, scratch)) L.bad = true; }
                }
            }
            i += adv;
            if (indel) {
                // [+-]<digits><that many letters>  (pileup.py:315-320), the shape samtools writes
                uint32_t k = i + 1, n = 0;
                if (!fast_digits(buf, k, n) || n > 4096u) { L.bad = true; n = 0; }
                for (uint32_t x = 0; x < n; x++) {
                    const unsigned c = buf[k + x];
                    if ((c | 0x20u) - 'a' >= 26u && c != '*') { L.bad = true; break; }
                }
                i = k + n;
            }
        }
        if (L.bad || i > e) { L.bad = true; done = true; }
    }
    if (buf[i] != '\t') L.bad = true;                  // the bases column must end in a tab
    tally_flush(L);
    if (L.bad || i >= e) return ST_FALLBACK;           // the tab must lie inside the line
    if (!HAS_QUAL) {
        qs = i + 1;
        nq = e - qs;
        if (nq < 1 || !all_printable(buf, qs, nq)) return ST_FALLBACK;
    }
    if (nb != nq) return ST_FALLBACK;                  // zip() truncation -> general path
    // ---- rank and call (pileup.py:259-266, 550-588) ------------------------------------------------
    const int oi = L.osym == 'A' ? 0 : L.osym == 'C' ? 1 : L.osym == 'G' ? 2 : L.osym == 'N' ? 3 : 4;
    uint32_t f[5], r[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        f[k] = L.t.f[k] + (k == ui ? L.dot : 0u) + ((L.osym && k == oi) ? L.ofwd : 0u);
        r[k] = L.t.r[k] + (k == ui ? L.comma : 0u) + ((L.osym && k == oi) ? L.orev : 0u);
    }
    uint32_t good = L.t.star;
#pragma unroll
    for (int k = 0; k < 5; k++) good += f[k] + r[k];
    if (good < 1) { out->base = '-'; out->fail = FAIL_RAWDPTH; return ST_OK; }
    // candidates in byte order  * A C G N T ; strict '>' keeps the smallest byte among equals
    uint32_t best = L.t.star, bf = L.t.star, br = 0;
    unsigned wsym = '*';
    const unsigned sym[5] = {'A', 'C', 'G', 'N', 'T'};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t tot = f[k] + r[k];
        if (tot > best) { best = tot; bf = f[k]; br = r[k]; wsym = sym[k]; }
    }
    out->fail = filter_mask(good, best, bf, br, p);
    out->base = (uint8_t)((wsym == U) ? ref : wsym);
    return ST_OK;
}

}  // namespace snpgpu
