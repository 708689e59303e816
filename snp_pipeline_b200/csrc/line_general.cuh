// line_general.cuh -- exact (any-input) parser for one pileup line.
//
// This is the path every line takes that the fast path (line_fast.cuh) declines: odd alphabets, indel
// tokens of unusual shape, short quality strings, extra columns, odd integers, lines longer than a staging
// slab.  It follows the reference's semantics to the letter:
//   pileup.py:206        line.rstrip().split()
//   pileup.py:209-274    Record._init_from_split_line
//   pileup.py:276-325    Record._strip_unwanted_base_patterns (two regex passes + '$' removal)
//   pileup.py:492-590    ConsensusCaller.call_consensus
// One thread walks one line, reading the text where it lies (global memory on the device).
#pragma once
#include "hd.cuh"

namespace snpgpu {

struct Tok { int64_t off; int64_t len; };

// line.rstrip().split(): records up to max_tok tokens, returns how many there are in total.
SNP_HD int split_tokens(const uint8_t *s, int64_t n, Tok *t, int max_tok) {
    int nt = 0;
    int64_t i = 0;
    while (i < n) {
        while (i < n && py_space(s[i])) i++;
        if (i >= n) break;
        int64_t b = i;
        while (i < n && !py_space(s[i])) i++;
        if (nt < max_tok) { t[nt].off = b; t[nt].len = i - b; }
        nt++;
    }
    return nt;
}

// Python int(token) for an ASCII token: optional sign, digits, single underscores between digits.
SNP_HD int py_int(const uint8_t *s, int64_t n, int64_t *out) {
    int64_t i = 0;
    bool neg = false;
    if (n > 0 && (s[0] == '+' || s[0] == '-')) { neg = s[0] == '-'; i = 1; }
    if (i >= n) return ST_VALUE;
    uint64_t v = 0;
    bool prev_digit = false;
    for (; i < n; i++) {
        unsigned c = s[i];
        if (is_digit(c)) {
            unsigned d = c - '0';
            if (v > (0x7fffffffffffffffULL - d) / 10) return ST_DOMAIN;
            v = v * 10 + d;
            prev_digit = true;
        } else if (c == '_' && prev_digit && i + 1 < n && is_digit(s[i + 1])) {
            prev_digit = false;
        } else {
            return ST_VALUE;
        }
    }
    *out = neg ? -(int64_t)v : (int64_t)v;
    return ST_OK;
}

// ---- walking the bases column ---------------------------------------------------------------------
// Both walkers hand every surviving base (after pileup.py:276-325) to `sink(c, k)`, k being its index in
// the stripped string, i.e. the index of the quality byte it pairs with (pileup.py:249).

// first index >= i that does not start a "^x" pair (pileup.py:312, regex \^. applied left to right)
SNP_HD int64_t skip_carets(const uint8_t *b, int64_t m, int64_t i) {
    while (i + 1 < m && b[i] == '^') i += 2;
    return i;
}

// Left-to-right walk.  Equivalent to the reference's right-to-left splice (pileup.py:315-320) unless an
// indel token starts inside the bytes another indel token removes; returns false (and stops) when it meets
// that shape, so the caller can redo the line with walk_spliced.
template <class Sink>
SNP_HD bool walk_streaming(const uint8_t *b, int64_t m, Sink &sink) {
    int64_t k = 0, i = 0;
    while (i < m) {
        unsigned c = b[i];
        if (c == '^' && i + 1 < m) { i += 2; continue; }
        if (c == '+' || c == '-') {
            int64_t j = skip_carets(b, m, i + 1);
            if (j < m && is_digit(b[j])) {
                uint64_t n = 0;
                for (;;) {
                    j = skip_carets(b, m, j);
                    if (j < m && is_digit(b[j])) {
                        if (n < (1ULL << 40)) n = n * 10 + (b[j] - '0');
                        j++;
                    } else {
                        break;
                    }
                }
                for (uint64_t skipped = 0; skipped < n; skipped++) {
                    j = skip_carets(b, m, j);
                    if (j >= m) break;
                    unsigned ch = b[j];
                    if (ch == '+' || ch == '-') {
                        int64_t j2 = skip_carets(b, m, j + 1);
                        if (j2 < m && is_digit(b[j2])) return false;
                    }
                    j++;
                }
                i = j;
                continue;
            }
        }
        i++;
        if (c == '$') continue;
        if (!sink(c, k)) return true;
        k++;
    }
    return true;
}

// Exact splice in a scratch buffer of m bytes.  Pass 1 compacts away the "^x" pairs; pass 2 runs right to
// left keeping the survivors on a stack that grows down from the end of the buffer: an ordinary byte is
// pushed; at a sign followed (in the caret-free string) by digits, the digits -- still on top of the
// stack -- are popped and read as N, then N more bytes are popped.  That is exactly
// "for match in reversed(matches): s = s[:start] + s[end+N:]".  Returns the stripped string's span.
SNP_HD void splice_exact(const uint8_t *b, int64_t m, uint8_t *scratch, int64_t *out_begin, int64_t *out_end) {
    int64_t n = 0;
    for (int64_t i = 0; i < m;) {
        if (b[i] == '^' && i + 1 < m) i += 2;
        else scratch[n++] = b[i++];
    }
    int64_t top = n;          // stack = scratch[top, n)
    int64_t run = 0;          // length of the run of original digits immediately right of i
    for (int64_t i = n - 1; i >= 0; i--) {
        unsigned c = scratch[i];          // still the original byte: top > i until we push it
        if ((c == '+' || c == '-') && run > 0) {
            uint64_t num = 0;
            for (int64_t d = 0; d < run; d++) {
                if (num < (1ULL << 40)) num = num * 10 + (scratch[top + d] - '0');
            }
            top += run;
            uint64_t avail = (uint64_t)(n - top);
            top += (int64_t)(num < avail ? num : avail);
            run = 0;
            continue;
        }
        scratch[--top] = (uint8_t)c;
        run = is_digit(c) ? run + 1 : 0;
    }
    *out_begin = top;
    *out_end = n;
}

template <class Sink>
SNP_HD void walk_stripped(const uint8_t *s, int64_t begin, int64_t end, Sink &sink) {
    int64_t k = 0;
    for (int64_t i = begin; i < end; i++) {
        unsigned c = s[i];
        if (c == '$') continue;
        if (!sink(c, k)) return;
        k++;
    }
}

// pass 1: totals per upper-cased symbol (pileup.py:259) and good_depth (pileup.py:253)
struct TotalSink {
    const uint8_t *q; int64_t nq; int thr; unsigned U, L;
    uint32_t *total; uint32_t good;
    SNP_HD bool operator()(unsigned c, int64_t k) {
        if (k >= nq) return false;                      // zip() ran out of quality bytes
        if ((int)q[k] < thr) return true;               // pileup.py:250
        if (c == '.') c = U;                            // pileup.py:255
        if (c == ',') c = L;                            // pileup.py:256
        total[up8(c) & 127]++;
        good++;
        return true;
    }
};
// pass 2: strand counts of one symbol (pileup.py:269-274)
struct StrandSink {
    const uint8_t *q; int64_t nq; int thr; unsigned U, L; unsigned want;
    uint32_t fwd, rev;
    SNP_HD bool operator()(unsigned c, int64_t k) {
        if (k >= nq) return false;
        if ((int)q[k] < thr) return true;
        if (c == '.') c = U;
        if (c == ',') c = L;
        if (c <= 'Z' && c == want) fwd++;
        if (c >= 'a' && up8(c) == want) rev++;
        return true;
    }
};

struct LineCall {
    int      status;       // ST_OK, ST_VALUE/INDEX/UNPACK/DOMAIN, or ST_NEED_ARENA
    int64_t  pos;
    int64_t  chrom_off, chrom_len;   // first column, relative to the line start
    int64_t  bases_len;    // length of the bases column (scratch bytes walk needs when ST_NEED_ARENA)
    uint8_t  ref;
    uint8_t  base;         // consensus character before the '-' substitutions of call_consensus.py:169-176
    uint8_t  fail;         // FAIL_* mask (without FAIL_REGION)
};

// Columns 1-2 only: what pileup.Reader.__iter__ does to a line before the position filter
// (pileup.py:423-427).  n excludes the line terminator.
SNP_HD void general_key(const uint8_t *line, int64_t n, LineCall *r) {
    r->status = ST_OK;
    for (int64_t i = 0; i < n; i++) if (line[i] >= 0x80) { r->status = ST_DOMAIN; return; }
    Tok t[2];
    int nt = split_tokens(line, n, t, 2);
    if (nt < 2) { r->status = ST_UNPACK; return; }
    r->chrom_off = t[0].off; r->chrom_len = t[0].len;
    r->status = py_int(line + t[1].off, t[1].len, &r->pos);
}

// The whole record (pileup.py:209-274) and the consensus call (pileup.py:492-590).  scratch: nullable;
// when the line needs the exact splice and scratch is null (or smaller than the bases column) the call
// returns ST_NEED_ARENA with bases_len set and must be repeated with a buffer of that size.
SNP_HD_NOINLINE void general_line(const uint8_t *line, int64_t n, const CallParams &p, uint8_t *scratch,
                                  int64_t scratch_len, LineCall *r) {
    r->status = ST_OK; r->pos = 0; r->ref = 0; r->base = '-'; r->fail = FAIL_RAWDPTH; r->bases_len = 0;
    for (int64_t i = 0; i < n; i++) if (line[i] >= 0x80) { r->status = ST_DOMAIN; return; }
    Tok t[6];
    int nt = split_tokens(line, n, t, 6);
    if (nt < 2) { r->status = ST_INDEX; return; }
    r->chrom_off = t[0].off; r->chrom_len = t[0].len;
    int st = py_int(line + t[1].off, t[1].len, &r->pos);
    if (st) { r->status = st; return; }
    if (nt < 4) { r->status = ST_INDEX; return; }
    int64_t raw_depth;
    st = py_int(line + t[3].off, t[3].len, &raw_depth);
    if (st) { r->status = st; return; }
    if (t[2].len != 1) { r->status = ST_DOMAIN; return; }
    r->ref = line[t[2].off];
    if (raw_depth == 0 || nt < 5) return;                     // empty record -> ('-', RawDpth)
    if (nt < 6) { r->status = ST_INDEX; return; }

    const uint8_t *b = line + t[4].off;
    int64_t m = t[4].len;
    r->bases_len = m;
    const uint8_t *q = line + t[5].off;
    int64_t nq = t[5].len;
    int thr = 33 + p.min_base_qual;
    unsigned U = up8(r->ref), L = low8(r->ref);

    uint32_t total[128];
    for (int i = 0; i < 128; i++) total[i] = 0;
    TotalSink s1{q, nq, thr, U, L, total, 0};
    bool spliced = false;
    int64_t sb = 0, se = 0;
    if (!walk_streaming(b, m, s1)) {
        if (!scratch || scratch_len < m) { r->status = ST_NEED_ARENA; return; }
        for (int i = 0; i < 128; i++) total[i] = 0;
        s1.good = 0;
        splice_exact(b, m, scratch, &sb, &se);
        walk_stripped(scratch, sb, se, s1);
        spliced = true;
    }
    if (s1.good < 1) return;                                   // most_common_good_bases is None
    unsigned w = 0;
    uint32_t best = 0;
    for (unsigned c = 0; c < 128; c++) if (total[c] > best) { best = total[c]; w = c; }   // (-count, byte)
    StrandSink s2{q, nq, thr, U, L, w, 0, 0};
    if (spliced) walk_stripped(scratch, sb, se, s2);
    else walk_streaming(b, m, s2);
    r->fail = filter_mask(s1.good, best, s2.fwd, s2.rev, p);
    r->base = (uint8_t)((w == U) ? r->ref : w);                // pileup.py:586-588
}

// ---- full tallies of one line, for the per-sample consensus VCF (vcf_writer.py:295-379) ----------------------
// What pileup.Record exposes to SingleSampleWriter._make_vcf_record_from_pileup: raw depth, the reference base's
// depths, and every other symbol that survived -- the ALT alleles -- with total / forward / reverse depth, in
// most_common_good_bases order (pileup.py:260-266: count descending, then byte ascending).
struct LineTallyOut {
    int      status;        // ST_OK, ST_VALUE/INDEX/UNPACK/DOMAIN, or ST_NEED_ARENA
    int64_t  pos, raw_depth;
    int64_t  chrom_off, chrom_len;
    int64_t  bases_len;
    uint8_t  ref, base, fail;
    uint8_t  has_depth;     // 0: most_common_good_bases is None (no good depth)
    uint8_t  first_is_ref;  // most_common_good_bases[0] == REF.upper()
    uint32_t rd, rdf, rdr;  // base_good_depth / forward / reverse of REF.upper()
    uint32_t n_alt;
};

// pass: totals, forward and reverse counts per upper-cased symbol (pileup.py:259, 269-274)
struct FullSink {
    const uint8_t *q; int64_t nq; int thr; unsigned U, L;
    uint32_t *total, *fwd, *rev; uint32_t good;
    SNP_HD bool operator()(unsigned c, int64_t k) {
        if (k >= nq) return false;
        if ((int)q[k] < thr) return true;
        if (c == '.') c = U;
        if (c == ',') c = L;
        total[up8(c) & 127]++;
        if (c <= 'Z') fwd[c & 127]++;
        if (c >= 'a') rev[up8(c) & 127]++;
        good++;
        return true;
    }
};

// tot/fwd/rev: caller-provided arrays of 128 counters each; on return they hold the per-symbol depths, from which
// the caller lists the n_alt ALT alleles with next_alt().
SNP_HD_NOINLINE void general_tally(const uint8_t *line, int64_t n, const CallParams &p, uint8_t *scratch,
                                   int64_t scratch_len, uint32_t *tot, uint32_t *fwd, uint32_t *rev, LineTallyOut *r) {
    r->status = ST_OK; r->pos = 0; r->raw_depth = 0; r->ref = 0; r->base = '-'; r->fail = FAIL_RAWDPTH; r->bases_len = 0;
    r->has_depth = 0; r->first_is_ref = 0; r->rd = r->rdf = r->rdr = 0; r->n_alt = 0; r->chrom_off = r->chrom_len = 0;
    for (int64_t i = 0; i < n; i++) if (line[i] >= 0x80) { r->status = ST_DOMAIN; return; }
    Tok t[6];
    int nt = split_tokens(line, n, t, 6);
    if (nt < 2) { r->status = ST_INDEX; return; }
    r->chrom_off = t[0].off; r->chrom_len = t[0].len;
    int st = py_int(line + t[1].off, t[1].len, &r->pos);
    if (st) { r->status = st; return; }
    if (nt < 4) { r->status = ST_INDEX; return; }
    st = py_int(line + t[3].off, t[3].len, &r->raw_depth);
    if (st) { r->status = st; return; }
    if (t[2].len != 1) { r->status = ST_DOMAIN; return; }
    r->ref = line[t[2].off];
    if (r->raw_depth == 0 || nt < 5) return;                  // empty record (pileup.py:226-234)
    if (nt < 6) { r->status = ST_INDEX; return; }
    const uint8_t *b = line + t[4].off;
    const int64_t m = t[4].len;
    r->bases_len = m;
    const unsigned U = up8(r->ref), L = low8(r->ref);
    for (int i = 0; i < 128; i++) { tot[i] = 0; fwd[i] = 0; rev[i] = 0; }
    FullSink s{line + t[5].off, t[5].len, 33 + p.min_base_qual, U, L, tot, fwd, rev, 0};
    if (!walk_streaming(b, m, s)) {
        if (!scratch || scratch_len < m) { r->status = ST_NEED_ARENA; return; }
        for (int i = 0; i < 128; i++) { tot[i] = 0; fwd[i] = 0; rev[i] = 0; }
        s.good = 0;
        int64_t sb = 0, se = 0;
        splice_exact(b, m, scratch, &sb, &se);
        walk_stripped(scratch, sb, se, s);
    }
    if (s.good < 1) return;
    r->has_depth = 1;
    unsigned w = 0;
    uint32_t best = 0;
    for (unsigned c = 0; c < 128; c++) if (tot[c] > best) { best = tot[c]; w = c; }
    r->fail = filter_mask(s.good, best, fwd[w], rev[w], p);
    r->base = (uint8_t)((w == U) ? r->ref : w);
    r->first_is_ref = w == U;
    r->rd = tot[U & 127]; r->rdf = fwd[U & 127]; r->rdr = rev[U & 127];
    uint32_t na = 0;
    for (unsigned c = 0; c < 128; c++) if (c != (U & 127u) && tot[c] > 0) na++;
    r->n_alt = na;
}

// The next ALT allele in most_common_good_bases order (count descending, byte ascending) among the symbols other
// than REF.upper() whose total is still non-zero; zeroes that total so that repeated calls walk the list.
SNP_HD unsigned next_alt(uint32_t *tot, unsigned ref_upper, uint32_t *count) {
    unsigned a = 0;
    uint32_t bc = 0;
    for (unsigned c = 0; c < 128; c++) if (c != (ref_upper & 127u) && tot[c] > bc) { bc = tot[c]; a = c; }
    *count = bc;
    tot[a] = 0;
    return a;
}

}  // namespace snpgpu
