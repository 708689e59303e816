// line_quick3.cuh -- the first-tier parser of the pileup kernel (k1_pileup.cu): one lane walks one line.
//
// It decides the lines whose call does not depend on WHICH other symbols a read carries -- after pileup.py:276-325 has
// stripped "^x" and "$", the '.' / ',' outnumber every other surviving byte together and none of those is the reference
// letter written out, so the reference base is the strict winner of pileup.py:260-266 and the caller (pileup.py:550-588)
// only needs
//     good_depth = surviving bytes,  consensus depth = '.' + ',',  forward = '.',  reverse = ','.
// Every byte of the line is validated on the way (contig name, separators, digits, a printable quality string exactly as
// long as the stripped bases -- the zip() of pileup.py:248); anything else is declined with ST_DETAIL and no side effects:
// an indel token, "^^", a trailing '^', depth 0, another contig, "\r\n", a byte >= 0x80 ...  Declined lines go to the
// follow-up kernel (line_fast.cuh -> line_general.cuh).
//
// Written for the integer pipes of sm_100 (profiles/micro/pipes.cu): LOP3 / SHF / PRMT issue on the ALU pipe, IMAD and
// IDP.4A on the FMA pipe, each at one warp-instruction per two cycles per scheduler -- so the SWAR range tests add their
// constants with IMAD (multiplier `one`, a kernel parameter the compiler cannot fold) and count flags with IDP.4A, which
// leaves the ALU pipe the logic only.  FLO / POPC (quarter rate) stay out of the loops.
//
// Memory: the parser reads the text through a policy object M -- ld(word) / ld4(word, out[4]) for the text's aligned words,
// row(word) / row4(word, out[4]) for the cached contig's name rows, tab16(i) for the filter tables: the warp's slice of
// shared memory in the pileup kernel, the text in global memory + tables in shared memory in its follow-up kernel, plain
// arrays in tests/cpu_sim -- aligned loads only, 32-bit addressing, and the compiler always knows the address space.  Preconditions: '\n' sentinels in bytes
// [limit, limit + QUICK_PAD) of the window.  Any byte values are safe: a byte >= 0x80 ends the column it stands in like a
// separator would (Q3_LOW; the key columns and the quality column test for it directly), and the line is declined.
#pragma once
#include "line_fast.cuh"

namespace snpgpu {

enum : int { ST_DETAIL = 67 };         // "not for this tier": the line goes to the follow-up kernel, no side effects
enum : int { ST_TALLY = 68 };          // declined like ST_DETAIL, but the line is well-formed and out->end is its '\n':
                                       // only the call needs the per-letter tallies of the second tier
enum : int { ST_ZERO = 70 };           // raw depth 0 with columns 1-4 well-formed: out->end is the byte behind the depth column's tab;
                                       // the call is ('-', RawDpth) once the rest of the line shows no CR / VT / FF / byte >= 0x80
enum : int { ST_SIGN = 69 };           // declined like ST_DETAIL by the first look (INDEL = false) over a '+' / '-' / ')' / '/' in
                                       // the bases column -- an indel token, as a rule: columns 1-5 hold no separator but
                                       // their tabs and no byte >= 0x80, out->end is the first byte of the quality column
constexpr uint32_t QUICK_PAD = 32;     // '\n' sentinels the caller keeps behind `limit` (word over-reads land there)

// acc + 128 * (number of bytes of `flags` that are 0x80); flags holds 0x80 / 0x00 bytes only
SNP_HD uint32_t flag_sum(uint32_t flags, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp4a(flags, 0x01010101u, acc);
#else
    return acc + 128u * (uint32_t)__builtin_popcount(flags);
#endif
}

// acc + 128 * (sum of the bytes of w whose byte in `flags` is 0x80); flags holds 0x80 / 0x00 bytes only
SNP_HD uint32_t flag_weigh(uint32_t w, uint32_t flags, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp4a(w, flags, acc);
#else
    for (int b = 0; b < 4; b++) acc += ((w >> (8 * b)) & 0xffu) * ((flags >> (8 * b)) & 0xffu);
    return acc;
#endif
}
SNP_HD uint32_t funnel_l8(uint32_t lo, uint32_t hi) {       // (hi << 8) | (lo >> 24)
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, 8);
#else
    return (hi << 8) | (lo >> 24);
#endif
}

SNP_HD SiteWord load_site_word(const SiteWord *p) {
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    return SiteWord{v.x, v.y, v.z, v.w};
#else
    return *p;
#endif
}

// value of four decimal digit bytes (0..9 each, most significant in the lowest byte)
SNP_HD uint32_t digits4_value(uint32_t w) {
    const uint32_t p = (w * 10u + (w >> 8)) & 0x00ff00ffu;    // byte 0: d0 d1, byte 2: d2 d3
    return (p & 0xffffu) * 100u + (p >> 16);
}

constexpr uint32_t Q3_NAMEW = 20;      // words per alignment row of the cached contig name (names up to 63 bytes; rows 16-byte aligned)
constexpr uint32_t Q3_MASK8_W = 4u * Q3_NAMEW;             // word offset of the [4][8] masks of the first eight words
constexpr uint32_t Q3_MASKC_W = Q3_MASK8_W + 32u;           // ... of the [4][4] masks of words 0, nw - 2, nw - 1 (longer names)
constexpr uint32_t Q3_ROWS_WORDS = Q3_MASKC_W + 16u;
static_assert(Q3_ROWS_WORDS == SITE_Q3ROWS_WORDS, "sites.cuh keeps one block of rows per contig");

// The contig a warp currently expects.  rows: [4][Q3_NAMEW] words of name + '\t' shifted right by a = 0..3 bytes (what
// the aligned words of a line that starts at byte a of a word look like); then [4][8] masks of the bytes that count in the
// first eight words (names of up to 25 bytes are compared as eight masked words, no loop); then [4][4] masks of the words
// that need one when the name is longer: word 0, word nw - 2, word nw - 1 (the words between them count whole).
struct Q3Contig {
    uint32_t rows_w;       // word index of the rows in the warp's slice
    uint32_t len1;         // name length + 1 (the tab); 0: nothing cached, every line is declined
    uint32_t nw;           // words compared per line (warp-uniform): max(3, (len1 + 6) >> 2)
    int32_t  cid;          // index in the site table, -1: a name the table does not hold
    int32_t  max_pos;      // largest site position on the contig, -1 when it holds none
    uint32_t word_base;    // bit_base >> 5: first word of the contig in the site bitmap
};

// word j of alignment row a (and its mask) of a name whose bytes name_at(0 .. L-2) are followed by a tab
template <class F>
SNP_HD void q3_row_word(const F &name_at, uint32_t L, uint32_t a, uint32_t j, uint32_t *w, uint32_t *mk) {
    uint32_t v = 0, m = 0;
    for (uint32_t b = 0; b < 4u; b++) {
        const uint32_t p = 4u * j + b;
        if (p >= a && p - a < L) {
            const uint32_t c = p - a + 1u < L ? (uint32_t)name_at(p - a) : (uint32_t)'\t';
            v |= c << (8u * b);
            m |= 0xffu << (8u * b);
        }
    }
    *w = v; *mk = m;
}

// word idx (0 .. Q3_ROWS_WORDS) of a contig's block of rows: names [4][Q3_NAMEW], masks [4][8], masks [4][4]
template <class F>
SNP_HD uint32_t q3_rows_word(const F &name_at, uint32_t L, uint32_t idx) {
    const uint32_t nw = ((L + 6u) >> 2) < 3u ? 3u : (L + 6u) >> 2;
    uint32_t v = 0, mk = 0;
    if (idx < Q3_MASK8_W) { q3_row_word(name_at, L, idx / Q3_NAMEW, idx % Q3_NAMEW, &v, &mk); return v; }
    if (idx < Q3_MASKC_W) { q3_row_word(name_at, L, (idx - Q3_MASK8_W) >> 3, (idx - Q3_MASK8_W) & 7u, &v, &mk); return mk; }
    const uint32_t a = (idx - Q3_MASKC_W) >> 2, which = (idx - Q3_MASKC_W) & 3u;
    q3_row_word(name_at, L, a, which == 0u || which == 3u ? 0u : nw - 3u + which, &v, &mk);
    return mk;
}

SNP_HD void q3_contig_set(Q3Contig *cc, uint32_t rows_w, uint32_t L, int32_t cid, int64_t max_pos, int64_t bit_base) {
    cc->rows_w = rows_w; cc->len1 = L; cc->cid = cid;
    const uint32_t nw = (L + 6u) >> 2;
    cc->nw = nw < 3u ? 3u : nw;
    cc->max_pos = (int32_t)(max_pos > 0x7fffffff ? 0x7fffffff : max_pos);
    cc->word_base = (uint32_t)(bit_base >> 5);
    if (L == 0u || 4u * cc->nw > 4u * Q3_NAMEW) { cc->len1 = 0; cc->max_pos = -1; }
}

struct Q3Line {
    uint32_t end;          // offset of the line terminator in the window
    uint32_t pos;          // column 2
    uint32_t after;        // offset of the byte behind the tab that ends column 2 (valid when the key columns were taken)
    uint8_t  base;         // consensus character before the '-' substitutions of call_consensus.py:169-176
    uint8_t  fail;         // FAIL_* mask (without FAIL_REGION)
};

// ---- the filters of pileup.py:556-584 as two threshold tables (built once per launch from the very same IEEE double
//      products): fail VarFreq <-> cons < tf[good];  fail StrBias <-> min(fwd, rev) < tb[cons].  Indices below Q3_TABN.
constexpr uint32_t Q3_TABN = 512;
// smallest integer c >= 0 with !((double)c < x), capped at 65535 (x = good * min_cons_freq or cons * min_cons_strand_bias)
SNP_HD uint32_t q3_threshold(double x) {
    if (!(x > 0.0)) return 0u;                            // (also NaN: every comparison with it is false)
    if (x > 65535.0) return 65535u;
    uint32_t c = (uint32_t)x;                             // floor
    while ((double)c < x) c++;
    return c;
}
SNP_HD uint32_t q3_tab_entry(uint32_t idx, const CallParams &p) {     // entry idx of [tf | tb]
    return idx < Q3_TABN ? q3_threshold((double)idx * p.min_cons_freq) : q3_threshold((double)(idx - Q3_TABN) * p.min_cons_strand_bias);
}
template <class M>
SNP_HD uint8_t q3_filter(const M &m, uint32_t good, uint32_t cons, uint32_t fwd, uint32_t rev, const CallParams &p) {
    if (good >= Q3_TABN) return filter_mask(good, cons, fwd, rev, p);
    uint32_t f = 0;
    if (cons < m.tab16(good)) f |= FAIL_VARFREQ;
    if ((int32_t)cons < p.min_cons_depth) f |= FAIL_DEPTH;
    const uint32_t lo = fwd < rev ? fwd : rev;
    if ((int32_t)lo < p.min_cons_strand_depth) f |= FAIL_STRDPTH;
    if (lo < m.tab16(Q3_TABN + cons)) f |= FAIL_STRBIAS;
    return (uint8_t)f;
}

// the SWAR adds, on the FMA pipe: one == 1 at run time
#define Q3_ADD(C, w)  (one * (uint32_t)(C) + (w))                       /* w + C          */
#define Q3_NADD(C, w) (one * (0u - (uint32_t)(C) - 1u) - (w))          /* ~(w + C)       */

// ---- columns 1-2: the cached contig's name + tab, 1..8 digits + tab.  false: not for this tier. ----------------------
template <class M>
SNP_HD bool q3_name(const M &m, uint32_t s, uint32_t limit, const Q3Contig &cc) {
    if (cc.len1 == 0u || s + cc.len1 + 24u > limit + QUICK_PAD) return false;
    const uint32_t a = s & 3u, k0 = s >> 2;
    const uint32_t r = cc.rows_w + a * Q3_NAMEW;
    uint32_t diff;
    if (cc.nw <= 8u) {                                  // (warp-uniform) eight masked words, all loads up front
        uint32_t n[8], mk[8], t[8];
        m.row4(r, n); m.row4(r + 4u, n + 4);
        m.row4(cc.rows_w + Q3_MASK8_W + 8u * a, mk); m.row4(cc.rows_w + Q3_MASK8_W + 8u * a + 4u, mk + 4);
#pragma unroll
        for (uint32_t j = 0; j < 8u; j++) t[j] = m.ld(k0 + j);
#pragma unroll
        for (uint32_t j = 0; j < 8u; j++) t[j] = (t[j] ^ n[j]) & mk[j];
        diff = (t[0] | t[1] | t[2]) | (t[3] | t[4] | t[5]) | (t[6] | t[7]);
    } else {
        const uint32_t q = cc.rows_w + Q3_MASKC_W + 4u * a, nw = cc.nw;
        diff = (m.ld(k0) ^ m.row(r)) & m.row(q);
        for (uint32_t j = 1u; j + 2u < nw; j++) diff |= m.ld(k0 + j) ^ m.row(r + j);
        diff |= (m.ld(k0 + nw - 2u) ^ m.row(r + nw - 2u)) & m.row(q + 1u);
        diff |= (m.ld(k0 + nw - 1u) ^ m.row(r + nw - 1u)) & m.row(q + 2u);
    }
    return diff == 0u;
}

template <class M>
SNP_HD bool q3_key(const M &m, uint32_t s, uint32_t limit, const Q3Contig &cc, uint32_t one, Q3Line *out) {
    const uint32_t H = 0x80808080u;
    out->after = s; out->pos = 0u;
    if (cc.len1 == 0u || s + cc.len1 + 24u > limit + QUICK_PAD) return false;      // (q3_name's own range check, first)
    const uint32_t i = s + cc.len1, k = i >> 2, sh = i << 3;
    const uint32_t a = m.ld(k), b = m.ld(k + 1u), c = m.ld(k + 2u);                  // asked for together with the name's words
    if (!q3_name(m, s, limit, cc)) return false;
    const uint32_t rl = funnel_r(a, b, sh), rh = funnel_r(b, c, sh);
    const uint32_t xl = rl ^ 0x30303030u, xh = rh ^ 0x30303030u;                      // digits -> 0..9
    const uint32_t ndl = Q3_ADD(0x76767676u, xl) & H, ndh = Q3_ADD(0x76767676u, xh) & H;   // bit 7: not a digit
    const bool in_lo = ndl != 0u, eight = (ndl | ndh) == 0u;
    const uint32_t n = in_lo ? (uint32_t)ctz32(ndl) >> 3 : (eight ? 8u : 4u + ((uint32_t)ctz32(ndh) >> 3));
    const uint32_t s8 = n << 3;
    // the digits right-aligned in eight bytes: v_hi = bytes n-4 .. n-1, v_lo = bytes n-8 .. n-5 (zero in front of byte 0)
    const uint32_t v_hi = eight ? xh : funnel_r(in_lo ? 0u : xl, in_lo ? xl : xh, s8);
    const uint32_t v_lo = eight ? xl : (in_lo ? 0u : funnel_r(0u, xl, s8));
    const uint32_t sep = (eight ? funnel_r(c, 0u, sh) : (in_lo ? rl : rh) >> (s8 & 31u)) & 0xffu;   // (eight digits: the byte behind them)
    bool bad = ((rl | rh) & H) != 0u;                     // (a byte >= 0x80 would have fooled the digit test)
    bad |= n == 0u || sep != '\t';
    out->after = i + n + 1u;
    // d0 .. d7 -> value, as dot products (the FMA pipe's): 100 d0 + 10 d1 + d2 and d3 of each half
    out->pos = (flag_weigh(v_lo, 0x00010a64u, 0u) * 10u + flag_weigh(v_lo, 0x01000000u, 0u)) * 10000u
             + flag_weigh(v_hi, 0x00010a64u, 0u) * 10u + flag_weigh(v_hi, 0x01000000u, 0u);
    return !bad;
}

#ifndef Q3_LOOPN
#define Q3_LOOPN 4            // words of the bases column per trip of the first tier's loop
#endif
// acc | (a & b), as ONE three-input logic instruction (the compiler splits it in two in half of the unrolled copies)
SNP_HD uint32_t or_and(uint32_t acc, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
#else
    return acc | (a & b);
#endif
}
// bit 7 <-> byte < 0x21 or byte >= 0x80.  The LOWEST flag of a word is exact whatever the bytes above it are (a byte
// >= 0xa1 carries into its upper neighbour, never downwards), and a word without a flag holds bytes 0x21..0x7f only: the
// SWAR tests below never see a byte that could carry.
#define Q3_LOW(w) ((Q3_NADD(0x5f5f5f5fu, w) | (w)) & H)
// one whole word of the bases column (no separator in it, bytes 0x21..0x7f): classes, anomalies, counts.  car / part /
// prevcar carry noise below bit 7 of every byte; whatever counts with them is masked with H.
#define Q3_BASES_WORD(w) do { \
        const uint32_t car = Q3_ADD(0x22222222u, w) & Q3_NADD(0x21212121u, w); \
        const uint32_t part = funnel_l8(prevcar, car); \
        const uint32_t dol = Q3_ADD(0x5c5c5c5cu, w) & Q3_NADD(0x5b5b5b5bu, w); \
        const uint32_t y2 = Q3_ADD(0x7f7f7f7fu, (w & MFD) ^ 0x2c2c2c2cu); \
        const uint32_t dck = ~(y2 | part) & H; \
        const uint32_t y3 = Q3_ADD(0x7f7f7f7fu, (w & MF9) ^ 0x29292929u); \
        const uint32_t yr = Q3_ADD(0x7f7f7f7fu, (w | 0x20202020u) ^ refb); \
        an0 |= ~(y3 | part); \
        an1 |= ~(yr | part); \
        an2 = or_and(an2, car, part); \
        a_rem = flag_sum((car | part | dol) & H, a_rem); \
        a_dc = flag_sum(dck, a_dc); \
        a_dot = flag_weigh(w, dck, a_dot); \
        prevcar = car; \
    } while (0)

// ---- columns 3-6 of a line whose key columns q3_key() took.  ST_OK (out->end / base / fail filled), ST_TALLY (out->end) or ST_DETAIL. --------
// INDEL (the follow-up kernel's dense second look): indel tokens [+-]<n><n letters> (pileup.py:315-320) of up to 999 bases
// are skipped byte-wise -- the bytes in front of the sign count like any others, the word loop starts again behind the
// token; anything else about a token (no digits, more than three, a symbol that is no letter / '*' inside it) declines.
template <bool INDEL = false, class M>
SNP_HD int q3_rest(const M &m, uint32_t i, uint32_t limit, const CallParams &p, uint32_t one, Q3Line *out) {
    const uint32_t H = 0x80808080u;
    // ---- columns 3-4: one letter, tab, 1..3 digits (not all '0'), tab -- the 8 bytes at offset i -------------------
    uint32_t x0, x1;
    const uint32_t hw1 = m.ld((i >> 2) + 1u), hw2 = m.ld((i >> 2) + 2u);   // (one of them holds the first bases too)
    {
        const uint32_t sh = i << 3;
        const uint32_t a = m.ld(i >> 2);
        x0 = funnel_r(a, hw1, sh);
        x1 = funnel_r(hw1, hw2, sh);
    }
    const unsigned ref = x0 & 0xffu;
    bool bad = ((ref | 0x20u) - 'a') >= 26u;              // a letter: '.'/',' stand for REF / ref (pileup.py:255-256)
    bad |= ((x0 >> 8) & 0xffu) != '\t';
    const uint32_t d = funnel_r(x0, x1, 16);               // the 4 bytes from the depth column's first
    const uint32_t y = d ^ 0x30303030u;                    // digits -> 0..9
    const uint32_t nd = (((y & 0x7f7f7f7fu) + 0x76767676u) | y) & H;   // bit 7: not a digit (exact for any byte)
    const uint32_t n2 = nd ? (uint32_t)ctz32(nd) >> 3 : 4u;
    bad |= n2 == 0u || n2 == 4u;                           // (four digits or more: next tier)
    bad |= ((d >> (8u * (n2 & 3u))) & 0xffu) != '\t';
    if (bad) return ST_DETAIL;
    const uint32_t b0 = i + 3u + n2;                       // first byte of the bases column
    if ((y & ((1u << (8u * (n2 & 3u))) - 1u)) == 0u) {     // depth 0 (pileup.py:226-234): whatever follows is ignored
        out->end = b0;
        return ST_ZERO;
    }
    // ---- column 5: bases, aligned words -----------------------------------------------------------------
    const uint32_t refb = (ref | 0x20u) * 0x01010101u;
    const uint32_t MFD = one * 0xfdfdfdfdu, MF9 = one * 0xf9f9f9f9u;   // (in registers: one LOP3 per masked compare)
    uint32_t k = b0 >> 2;
    uint32_t w = k == (i >> 2) + 1u ? hw1 : hw2;           // b0 = i + 4 .. i + 6: no load, no wait
    {   // the bytes of the first word in front of the column become a symbol that counts nowhere
        const uint32_t mk = 0xffffffffu << ((b0 & 3u) * 8u);
        w = (w & mk) | (0x30303030u & ~mk);
    }
    uint32_t a_rem = 0, a_dc = 0, a_dot = 0;               // 128 x (removed bytes, kept '.'/',', sum of the kept '.'/',' bytes)
    uint32_t an0 = 0, an1 = 0, an2 = 0, prevcar = 0;     // anomaly flags (bit 7 of a byte; other bits: noise)
    uint32_t low;
    uint32_t wn = 0, wn2 = 0, wn3 = 0;                     // the words behind w, asked for a trip ahead (first look only)
    if constexpr (Q3_LOOPN > 1 && !INDEL) wn = m.ld(k + 1u);
    if constexpr (Q3_LOOPN > 2 && !INDEL) wn2 = m.ld(k + 2u);
    if constexpr (Q3_LOOPN > 3 && !INDEL) wn3 = m.ld(k + 3u);
    if constexpr (INDEL) {
        // The second look.  Round by round: all lanes walk their words up to the next sign or the separator (inner loop),
        // meet again, and the lanes that stopped at a sign skip their token TOGETHER -- one pass of the token code per
        // round and warp, not one per lane (a line carries one token as a rule, at a different word in every lane).
        for (;;) {
            uint32_t sign, car, part;
            for (;;) {
                low = Q3_LOW(w);                           // bit 7 <-> byte < 0x21 (or >= 0x80): the column ends here
                const uint32_t first = low & (0u - low);
                const uint32_t valid = low ? (first - 1u) & H : H;                             // the column's bytes of this word
                car = Q3_ADD(0x22222222u, w) & Q3_NADD(0x21212121u, w) & valid;
                part = funnel_l8(prevcar, car);
                const uint32_t y3 = Q3_ADD(0x7f7f7f7fu, (w & MF9) ^ 0x29292929u);
                sign = ~(y3 | part) & valid;                                                   // ) + - / that is no "^x" quality
                if ((sign | low) != 0u) break;
                Q3_BASES_WORD(w);
                w = m.ld(++k);
            }
            if (sign == 0u) break;                         // the separator's word: the common tail takes it
            const uint32_t fs = sign & (0u - sign);
            const uint32_t sb = (uint32_t)ctz32(fs) >> 3;                                      // byte of the sign in the word
            const uint32_t ch = (w >> (8u * sb)) & 0xffu;
            if (ch != '+' && ch != '-') return ST_DETAIL;
            const uint32_t v2 = (fs - 1u) & H;                                                 // the bytes in front of the sign
            const uint32_t car2 = car & v2, part2 = part & v2;
            const uint32_t dol2 = Q3_ADD(0x5c5c5c5cu, w) & Q3_NADD(0x5b5b5b5bu, w) & v2;
            const uint32_t y2 = Q3_ADD(0x7f7f7f7fu, (w & MFD) ^ 0x2c2c2c2cu);
            const uint32_t dck2 = ~(y2 | part) & v2;
            const uint32_t yr = Q3_ADD(0x7f7f7f7fu, (w | 0x20202020u) ^ refb);
            an1 |= ~(yr | part) & v2;
            an2 |= car2 & part2;
            a_rem = flag_sum(car2 | part2 | dol2, a_rem);
            a_dc = flag_sum(dck2, a_dc);
            a_dot = flag_weigh(w, dck2, a_dot);
            uint32_t t = 4u * k + sb + 1u, n = 0, nd = 0;                                      // the token: 1..3 digits, n symbols
            while (nd < 3u && (uint32_t)m.byte(t) - '0' < 10u) { n = n * 10u + ((uint32_t)m.byte(t) - '0'); t++; nd++; }
            if (nd == 0u || (uint32_t)m.byte(t) - '0' < 10u || t + n > limit) return ST_DETAIL;   // a bare sign, a long number
            for (uint32_t x = 0; x < n; x++) {
                const uint32_t c2 = m.byte(t + x);
                if ((c2 | 0x20u) - 'a' >= 26u && c2 != '*') return ST_DETAIL;
            }
            a_rem += 128u * (1u + nd + n);
            const uint32_t wpos = t + n;                                                        // go on behind the token
            k = wpos >> 2;
            w = m.ld(k);
            const uint32_t mk = 0xffffffffu << ((wpos & 3u) * 8u);
            w = (w & mk) | (0x30303030u & ~mk);
            prevcar = 0;
        }
    } else {
        for (;;) {
            low = Q3_LOW(w);                               // bit 7 <-> byte < 0x21 (or >= 0x80): the column ends here
            if (low) break;
            if constexpr (Q3_LOOPN > 1) {                  // several words a trip, each left as soon as it holds the separator:
                Q3_BASES_WORD(w);                          // the later words' loads have whole words' work to arrive
#define Q3_NEXT_WORD(x) w = (x); ++k; low = Q3_LOW(w); if (low) break; Q3_BASES_WORD(w)
                Q3_NEXT_WORD(wn);
                if constexpr (Q3_LOOPN > 2) { Q3_NEXT_WORD(wn2); }
                if constexpr (Q3_LOOPN > 3) { Q3_NEXT_WORD(wn3); }
#undef Q3_NEXT_WORD
                w = m.ld(++k);
                wn = m.ld(k + 1u);
                if constexpr (Q3_LOOPN > 2) wn2 = m.ld(k + 2u);
                if constexpr (Q3_LOOPN > 3) wn3 = m.ld(k + 3u);
            } else {
                Q3_BASES_WORD(w);
                w = m.ld(++k);
            }
        }
    }
    uint32_t q0;
    bool sep_tab;
    const uint32_t qw1 = m.ld(k + 1u), qw2 = m.ld(k + 2u); // the quality column starts in this word or the next: asked for now
    {   // the word that holds the separator: the same, restricted to the bytes in front of it
        const uint32_t first = low & (0u - low);
        const uint32_t valid = (first - 1u) & H;
        const uint32_t j = (uint32_t)ctz32(first) >> 3;
        const uint32_t car = Q3_ADD(0x22222222u, w) & Q3_NADD(0x21212121u, w) & valid;
        const uint32_t part = funnel_l8(prevcar, car);                     // may reach the separator itself
        const uint32_t dol = Q3_ADD(0x5c5c5c5cu, w) & Q3_NADD(0x5b5b5b5bu, w) & valid;
        const uint32_t y2 = Q3_ADD(0x7f7f7f7fu, (w & MFD) ^ 0x2c2c2c2cu);
        const uint32_t dck = ~(y2 | part) & valid;
        const uint32_t y3 = Q3_ADD(0x7f7f7f7fu, (w & MF9) ^ 0x29292929u);
        const uint32_t yr = Q3_ADD(0x7f7f7f7fu, (w | 0x20202020u) ^ refb);
        an0 |= ~(y3 | part) & valid;
        an1 |= ~(yr | part) & valid;
        an2 |= (car & part) | (part & first);                              // "^" + separator: trailing '^'
        a_rem = flag_sum((car | part | dol) & valid, a_rem);
        a_dc = flag_sum(dck, a_dc);
        a_dot = flag_weigh(w, dck, a_dot);
        sep_tab = ((w >> (8u * j)) & 0xffu) == '\t';                       // the column ends in a tab
        q0 = 4u * k + j + 1u;
    }
    const uint32_t bases_len = q0 - 1u - b0;
    if (((an0 | an1 | an2) & H) != 0u || !sep_tab || bases_len == 0u) {
        if (!INDEL && sep_tab && (an0 & H) != 0u) { out->end = q0; return ST_SIGN; }
        return ST_DETAIL;
    }
    const uint32_t nb = bases_len - (a_rem >> 7);          // length of the stripped string
    const uint32_t qe = q0 + nb;                           // where the line has to end
    if (nb < 1u || qe > limit) return ST_DETAIL;
    // ---- column 6: as many printable bytes as bases survived, then the line end (pileup.py:248-250) ----
    {
        uint32_t kq = q0 >> 2;
        const uint32_t ke = qe >> 2;
        uint32_t v = kq == k ? w : qw1;
        const uint32_t mk = 0xffffffffu << ((q0 & 3u) * 8u);
        v = (v & mk) | (0x30303030u & ~mk);
        uint32_t acc = H;                                  // bit 7 stays set while every byte is in 0x21..0x7f
        uint32_t v1 = kq == k ? qw1 : qw2;
        while (kq + 1u < ke) {                             // two words per step, the next two asked for a step ahead
            const uint32_t n0 = m.ld(kq + 2u), n1 = m.ld(kq + 3u);
            acc &= Q3_ADD(0x5f5f5f5fu, v) & ~v;
            acc &= Q3_ADD(0x5f5f5f5fu, v1) & ~v1;
            kq += 2u;
            v = n0; v1 = n1;
        }
        if (kq < ke) {
            acc &= Q3_ADD(0x5f5f5f5fu, v) & ~v;
            v = v1;
            ++kq;
        }
        const uint32_t r = qe & 3u;
        const uint32_t below = (0x80u << (8u * r)) - 1u;   // the bytes in front of the terminator (bit 7 of each)
        const uint32_t good = Q3_ADD(0x5f5f5f5fu, v) & ~v & H;
        if ((acc & H) != H || (good & below) != (below & H) || ((v >> (8u * r)) & 0xffu) != '\n') return ST_DETAIL;
    }
    // ---- call (pileup.py:550-588): the reference base wins outright ---------------------------------
    const uint32_t dc = a_dc >> 7, dot = ((a_dot >> 7) - 0x2cu * dc) >> 1;     // (0x2c nc + 0x2e nd - 0x2c (nc + nd)) / 2
    const uint8_t fail = q3_filter(m, nb, dc, dot, dc - dot, p);
    out->end = qe;
    if (dc <= nb - dc) return ST_TALLY;
    out->base = (uint8_t)ref;
    out->fail = fail;
    return ST_OK;
}

// (rare) the search of q3_find_nl() word by word from offset `from`, once a flagged byte turned out not to be a '\n'
template <class M>
SNP_HD_NOINLINE uint32_t q3_find_nl_slow(const M &m, uint32_t from) {
    uint32_t k = from >> 2;
    uint32_t w = m.ld(k);
    const uint32_t mk = 0xffffffffu << ((from & 3u) * 8u);
    w = (w & mk) | (0x30303030u & ~mk);
    for (;;) {
        const uint32_t t = w ^ 0x0a0a0a0au;
        const uint32_t z = ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;     // exact for any byte values
        if (z) return 4u * k + ((uint32_t)ctz32(z) >> 3);
        w = m.ld(++k);
    }
}

// position of the first '\n' at or behind offset `from` (the sentinels behind `limit` end the search); *odd (in/out) is
// set when a byte 0x0b..0x0d (VT, FF, CR) or a byte >= 0x80 lies in front of it in the words looked at.  Exact for any
// byte values.  Aligned 16-byte chunks: the lanes of a warp step alike, and what follows the loop runs once, converged.
template <class M>
SNP_HD uint32_t q3_find_nl(const M &m, uint32_t from, uint32_t one, uint32_t *odd) {
    const uint32_t H = 0x80808080u;
    uint32_t c = from >> 4;
    uint32_t w[4];
    m.ld4(4u * c, w);
    {   // bytes in front of `from`: made harmless
        const uint32_t fw = (from >> 2) & 3u, mk = 0xffffffffu << ((from & 3u) * 8u);
#pragma unroll
        for (uint32_t j = 0; j < 4u; j++) {
            const uint32_t keep = j < fw ? 0u : (j == fw ? mk : 0xffffffffu);
            w[j] = (w[j] & keep) | (0x30303030u & ~keep);
        }
    }
    uint32_t guard = 0;
    uint32_t in[4];
    for (;;) {
        // bit 7 <-> byte (+ a carry from a byte >= 0x80 below it) in 0x0a..0x0d: a true '\n' is always flagged
#pragma unroll
        for (uint32_t j = 0; j < 4u; j++) in[j] = Q3_ADD(0x76767676u, w[j]) & Q3_NADD(0x72727272u, w[j]);
        if (((in[0] | in[1]) | (in[2] | in[3])) & H) break;
        guard |= w[0] | w[1];
        guard |= w[2] | w[3];
        m.ld4(4u * ++c, w);
    }
    const uint32_t f0 = in[0] & H, f1 = in[1] & H, f2 = in[2] & H;
    const uint32_t j = f0 ? 0u : (f1 ? 1u : (f2 ? 2u : 3u));
    const uint32_t f = f0 ? f0 : (f1 ? f1 : (f2 ? f2 : in[3] & H));
    const uint32_t wj = f0 ? w[0] : (f1 ? w[1] : (f2 ? w[2] : w[3]));
    if (j > 0u) guard |= w[0];
    if (j > 1u) guard |= w[1];
    if (j > 2u) guard |= w[2];
    const uint32_t b = (uint32_t)ctz32(f) >> 3;
    const uint32_t at = 16u * c + 4u * j + b;
    *odd |= (guard | (wj & ((1u << (8u * b)) - 1u))) & H;               // (bytes behind the flagged one are not part of the guard)
    if (((wj >> (8u * b)) & 0xffu) == 0x0au) return at;
    *odd |= H;                                          // VT / FF / CR, or a carry artefact: the exact parser's line
    return q3_find_nl_slow(m, at + 1u);
}

// the site of (cached contig, pos) from the packed word of its 32-position group
SNP_HD void q3_site(const SiteWord &sw, uint32_t bb, int32_t *site, uint32_t *flags) {
    *site = ((sw.any >> bb) & 1u) ? (int32_t)(sw.rank + (uint32_t)popc32(sw.any & ((1u << bb) - 1u))) : -1;
    *flags = (((sw.snp >> bb) & 1u) ? SITE_SNP : 0u) | (((sw.exc >> bb) & 1u) ? SITE_EXCLUDED : 0u);
}

}  // namespace snpgpu
