// cpu_sim.cpp -- TEST HARNESS ONLY (tests/test_cpu_sim.py); never loaded by the product package.
//
// Compiles the very same per-line parsers the kernel runs (line_fast.cuh, line_general.cuh, sites.cuh) with g++ and
// drives them over a text the way k1_pileup.cu does -- '\n'-delimited lines, fast parser first, the exact parser
// for whatever it declines, last-line-wins per site -- so that their logic can be checked against the oracle in
// the build container, which has no GPU.  The tile/scan machinery of the kernel is GPU-only and is covered by the
// -m gpu tests.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "line_fast.cuh"
#include "line_quick3.cuh"
#include "line_general.cuh"
#include "sites_host.h"

using namespace snpgpu;


struct HostWin {                // line_quick3.cuh's memory policy over a plain array: window words, then the name rows
    const uint32_t *p;
    const uint16_t *tab;
    uint32_t ld(uint32_t k) const { return p[k]; }
    uint32_t tab16(uint32_t k) const { return tab[k]; }
    void ld4(uint32_t k, uint32_t *w) const { w[0] = p[k]; w[1] = p[k + 1]; w[2] = p[k + 2]; w[3] = p[k + 3]; }
    uint32_t byte(uint32_t off) const { return reinterpret_cast<const uint8_t *>(p)[off]; }
    uint32_t row(uint32_t k) const { return p[k]; }
    void row4(uint32_t k, uint32_t *w) const { ld4(k, w); }
};

extern "C" {

// counters[0] = lines, [1] = parsed, [2] = lines through the general path, [3] = error offset, [4] = error code,
// [5] = lines decided by the first-tier parser, [6] = of those, lines that needed its token-skipping second look
int cpusim_pileup(const uint8_t *text, size_t nbytes, const char *contig_names, const int32_t *name_off,
                  int32_t n_contigs, const int32_t *snp_contig, const int64_t *snp_pos, size_t n_snp,
                  const int32_t *exc_contig, const int64_t *exc_pos, size_t n_exc, const CallParams *p, int all_positions,
                  int force_general, uint8_t *row_out, uint16_t *line_out, size_t line_out_cap, uint64_t *counters) {
    HostSites h;
    const char *why;
    int rc = build_host_sites(contig_names, name_off, n_contigs, snp_contig, snp_pos, n_snp, exc_contig, exc_pos, n_exc,
                              &h, &why);
    if (rc) return 100 + rc;
    SiteTable t = h.view();
    std::vector<uint64_t> cells(h.n_unique + 1, 0);
    // staging copy: 4-byte aligned, '\n' sentinels behind the text (what the kernel's shared-memory window looks like)
    std::vector<uint32_t> store((nbytes + 64) / 4 + 8 + Q3_ROWS_WORDS, 0x0a0a0a0au);   // ... followed by the name rows
    const uint32_t rows_w = (uint32_t)(((nbytes + 64) / 4 + 7) & ~(size_t)3);
    uint32_t *rows = store.data() + rows_w;
    Q3Contig cc3;
    int cc3_cid = -2;
    std::vector<uint16_t> tab(2 * Q3_TABN);
    for (uint32_t k = 0; k < 2 * Q3_TABN; k++) tab[k] = (uint16_t)q3_tab_entry(k, *p);
    uint8_t *buf = reinterpret_cast<uint8_t *>(store.data());
    memcpy(buf, text, nbytes);
    memset(buf + nbytes, '\n', 48);
    // classic-Mac line ends: what snpgpu_normalize_newlines_dev does before the kernel is run again (api.cu)
    for (size_t i = 0; i < nbytes; i++)
        if (buf[i] == '\r' && buf[i + 1] != '\n') buf[i] = '\n';
    uint64_t n_lines = 0, n_parsed = 0, n_general = 0, n_quick = 0, n_second = 0;
    int hint = 0;
    size_t s = 0;
    counters[3] = ~0ull; counters[4] = 0;
    std::vector<uint8_t> scratch;
    while (s < nbytes) {
        size_t e = s;
        bool high = false, lone_cr = false;
        while (e < nbytes && buf[e] != '\n') {
            if (buf[e] >= 0x80) high = true;
            if (buf[e] == '\r' && e + 1 < nbytes && buf[e + 1] != '\n') lone_cr = true;
            e++;
        }
        const size_t line_idx = n_lines++;
        int st = ST_FALLBACK;
        FastLine fl;
        if (!force_general && p->min_base_qual <= 0) {                   // first tier: the steps of k1_pileup.cu's lane loop
            struct { int32_t site; uint32_t end; uint8_t base, fail, flags; } q = {-1, 0u, 0, 0, 0};
            {
                if (cc3_cid != hint) {                                   // (the kernel follows the tile's first line)
                    const uint32_t L = (uint32_t)(t.name_off[hint + 1] - t.name_off[hint]) + 1u;
                    for (uint32_t k = 0; k < Q3_ROWS_WORDS; k++) rows[k] = t.q3rows[(size_t)hint * SITE_Q3ROWS_WORDS + k];   // (as build_host_sites made them)
                    q3_contig_set(&cc3, rows_w, L, hint, t.max_pos[hint], t.bit_base[hint]);
                    cc3_cid = hint;
                }
                const HostWin m{store.data(), tab.data()};
                uint32_t odd = 0;
                const uint32_t nl = q3_find_nl(m, (uint32_t)s, 1u, &odd);
                if (nl != e) { counters[3] = s; counters[4] = 97; break; }          // harness self-check: the line end
                bool seen_odd = false;
                for (size_t x = s; x < e; x++) seen_odd |= (buf[x] >= 0x0b && buf[x] <= 0x0d);
                if (seen_odd && !odd) { counters[3] = s; counters[4] = 96; break; }  // ... and the odd-byte flag
                Q3Line q3;
                st = ST_DETAIL;
                if (q3_key(m, (uint32_t)s, (uint32_t)nbytes, cc3, 1u, &q3)) {
                    const bool known = (int32_t)q3.pos <= cc3.max_pos;
                    const uint32_t widx = cc3.word_base + (q3.pos >> 5), bb = q3.pos & 31u;
                    if (!all_positions && !(known && ((t.bits[widx] >> bb) & 1u))) st = ST_SKIP;
                    else {
                        SiteWord sw{0u, 0u, 0u, 0u};
                        if (known) sw = t.words[widx];
                        st = q3_rest(m, q3.after, (uint32_t)nbytes, *p, 1u, &q3);
                        bool second = st == ST_DETAIL;                   // the follow-up kernel's second look: indel tokens skipped
                        if (st == ST_TALLY) {                            // well-formed, only the call is left: queued with its end known
                            if (q3.end != e) { counters[3] = s; counters[4] = 94; break; }   // harness self-check: that end
                            st = ST_DETAIL;
                        } else if (st == ST_ZERO) {                      // depth 0: queued; the follow-up kernel's first tier calls it
                            uint32_t odd2 = 0;                           // ('-', RawDpth) unless the pileup kernel saw something odd
                            if (q3_find_nl(m, q3.end, 1u, &odd2) != e) { counters[3] = s; counters[4] = 90; break; }   // harness self-check
                            bool odd_q = false;
                            for (size_t x = q3.end; x < e; x++) odd_q |= (buf[x] >= 0x0b && buf[x] <= 0x0d) || buf[x] >= 0x80;
                            if (odd_q && !odd2) { counters[3] = s; counters[4] = 89; break; }
                            if (odd2) st = ST_DETAIL;
                            else if (q3_rest<true>(m, q3.after, (uint32_t)e, *p, 1u, &q3) != ST_ZERO) { counters[3] = s; counters[4] = 88; break; }
                            else { st = ST_OK; q3.end = (uint32_t)e; q3.base = (uint8_t)'-'; q3.fail = FAIL_RAWDPTH; }
                        } else if (st == ST_SIGN) {                      // the pileup kernel looks for the line end from the quality column on
                            uint32_t odd2 = 0;
                            if (q3_find_nl(m, q3.end, 1u, &odd2) != e) { counters[3] = s; counters[4] = 93; break; }   // harness self-check
                            bool odd_q = false, odd_f = false;
                            for (size_t x = q3.end; x < e; x++) odd_q |= (buf[x] >= 0x0b && buf[x] <= 0x0d) || buf[x] >= 0x80;
                            if (odd_q && !odd2) { counters[3] = s; counters[4] = 92; break; }
                            for (size_t x = s; x < q3.end; x++) odd_f |= (buf[x] < 0x21 && buf[x] != '\t') || buf[x] >= 0x80;
                            if (odd_f) { counters[3] = s; counters[4] = 91; break; }         // ... having seen nothing odd in front of it
                            st = ST_DETAIL;
                            second = true;
                        }
                        if (second) {
                            st = q3_rest<true>(m, q3.after, (uint32_t)e, *p, 1u, &q3);
                            if (st == ST_OK && q3.end != e) st = ST_DETAIL;
                            if (st == ST_TALLY) st = ST_DETAIL;
                            if (st == ST_OK) n_second++;
                        }
                        if (st == ST_OK) {
                            uint32_t fl3;
                            q3_site(sw, bb, &q.site, &fl3);
                            q.flags = (uint8_t)fl3; q.end = q3.end; q.base = q3.base; q.fail = q3.fail;
                        }
                    }
                }
                if (st == ST_SKIP && odd) st = ST_DETAIL;               // (the kernel declines odd lines it would skip)
            }
            if (high && (st == ST_OK || st == ST_SKIP)) { counters[3] = s; counters[4] = 95; break; }   // must decline bytes >= 0x80
            if (st == ST_OK) {
                if (q.end != e) { counters[3] = s; counters[4] = 99; break; }   // harness self-check: the line end
                if (q.flags != (q.site >= 0 ? h.flags[q.site] : 0)) { counters[3] = s; counters[4] = 98; break; }   // ... and the site's flags
                fl.base = q.base; fl.fail = q.fail; fl.site = q.site;
                n_quick++;
            }
        }
        if (high) st = ST_FALLBACK;
        if (!force_general && !high && (st == ST_DETAIL || p->min_base_qual > 0)) {
            if (p->min_base_qual > 0) st = fast_line<true>(buf, (uint32_t)s, (uint32_t)e, t, hint, *p, all_positions != 0, &fl);
            else st = fast_line<false>(buf, (uint32_t)s, (uint32_t)e, t, hint, *p, all_positions != 0, &fl);
        }
        unsigned base = 0, fail = 0;
        int32_t site = -1;
        bool have = false;
        if (st == ST_OK) { base = fl.base; fail = fl.fail; site = fl.site; have = true; }
        else if (st == ST_FALLBACK) {
            n_general++;
            int err = 0;
            LineCall r;
            const uint8_t *line = buf + s;
            int64_t n = (int64_t)(e - s);
            if (lone_cr) err = ST_DOMAIN;
            bool wanted = true;
            if (!err && !all_positions) {
                general_key(line, n, &r);
                err = r.status;
                if (!err) {
                    site = site_find(t, contig_find(t, line + r.chrom_off, r.chrom_len), r.pos);
                    wanted = site >= 0;
                }
            }
            if (!err && wanted) {
                general_line(line, n, *p, nullptr, 0, &r);
                if (r.status == ST_NEED_ARENA) {
                    scratch.resize((size_t)r.bases_len + 16);
                    general_line(line, n, *p, scratch.data(), r.bases_len, &r);
                }
                err = r.status;
                if (!err) {
                    if (all_positions) site = site_find(t, contig_find(t, line + r.chrom_off, r.chrom_len), r.pos);
                    base = r.base; fail = r.fail; have = true;
                }
            }
            if (err) { counters[3] = s; counters[4] = (uint64_t)err; break; }
        }
        if (have) {
            unsigned flags = site >= 0 ? h.flags[site] : 0u;
            if (flags & SITE_EXCLUDED) fail |= FAIL_REGION;
            unsigned cell = (fail || base == '*') ? (unsigned)'-' : base;
            if (flags & SITE_SNP) cells[site] = ((uint64_t)(s + 1) << 8) | cell;
            if (line_out && line_idx < line_out_cap) line_out[line_idx] = (uint16_t)(cell | (fail << 8));
            n_parsed++;
        } else if (line_out && all_positions && line_idx < line_out_cap) {
            line_out[line_idx] = 0;
        }
        s = e + 1;
    }
    for (size_t k = 0; k < n_snp; k++) {
        uint64_t c = cells[h.snp_unique[k]];
        row_out[k] = c ? (uint8_t)(c & 0xff) : (uint8_t)'-';
    }
    counters[0] = n_lines; counters[1] = n_parsed; counters[2] = n_general; counters[5] = n_quick; counters[6] = n_second;
    return (int)counters[4];
}

}  // extern "C"
