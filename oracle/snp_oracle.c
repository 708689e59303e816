/*
 * snp_oracle.c -- CPU restatement of the CFSAN SNP Pipeline hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: it may be built,
 * loaded and called from tests/, from __graft_entry__.smoke() and from the
 * cpu_baseline / --impl reference legs of bench.py, and from nowhere else.  The
 * product (snp_pipeline_b200/ + libsnpgpu.so) never links or imports it.
 *
 * It restates, in plain scalar C, the algorithm of these reference files
 * (paths relative to the upstream repository):
 *   snppipeline/pileup.py:36-37     the two regular expressions
 *   snppipeline/pileup.py:209-274   Record._init_from_split_line
 *   snppipeline/pileup.py:276-325   Record._strip_unwanted_base_patterns
 *   snppipeline/pileup.py:408-429   Reader.__iter__
 *   snppipeline/pileup.py:492-590   ConsensusCaller.call_consensus
 *   snppipeline/call_consensus.py:142-192   the per-sample driver loop + gather
 *   snppipeline/merge_sites.py:94-116 + utils.py:1056-1070   site union
 *   snppipeline/utils.py:1135-1165 + distance.py:90-96   pairwise distance
 *
 * Parity is pinned (tests/test_oracle_*.py): against the reference's doctest
 * known answers, against the bundled lambda-virus / Agona / Listeria golden
 * files, and against vectors produced by running the reference's own Python in
 * the build container (tests/golden/make_golden.py).
 *
 * Domain notes (the reference is Python; this is C):
 *   - text is treated as bytes; any byte >= 0x80 makes a parsed line return
 *     ORACLE_E_DOMAIN (Python would decode UTF-8 there);
 *   - integers follow Python int() for ASCII: optional sign, digits, single
 *     underscores between digits; magnitudes beyond int64 -> ORACLE_E_DOMAIN;
 *   - a reference-base column that is not exactly one byte -> ORACLE_E_DOMAIN
 *     (Python would splice a multi-character string into the bases).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK          0
#define ORACLE_E_VALUE     1  /* ValueError: int() of a malformed column          */
#define ORACLE_E_INDEX     2  /* IndexError: too few columns for Record()          */
#define ORACLE_E_UNPACK    3  /* ValueError: fewer than 2 columns in filter mode   */
#define ORACLE_E_DOMAIN    4  /* input outside the byte/int64 domain stated above  */

#define FAIL_RAWDPTH  1
#define FAIL_VARFREQ  2
#define FAIL_DEPTH    4
#define FAIL_STRDPTH  8
#define FAIL_STRBIAS 16
#define FAIL_REGION  32

typedef struct {
    int    min_base_qual;
    double min_cons_freq;
    int    min_cons_depth;
    int    min_cons_strand_depth;
    double min_cons_strand_bias;
} oracle_params;

typedef struct {
    int     status;
    int     ntok;
    int64_t pos;
    int64_t raw_depth;
    int     chrom_off, chrom_len;       /* first column, as a slice of the line     */
    uint8_t ref;                        /* reference base, case preserved           */
    int     has_bases;                  /* 0 -> "empty record" (pileup.py:226-234)  */
    int     good_depth, fwd_good_depth, rev_good_depth;
    int     total[128], fwd[128], rev[128];
    int     n_common;                   /* 0 -> most_common_good_bases is None      */
    uint8_t common[128];                /* sorted by (-count, byte)                 */
} oracle_record;

/* Python str.isspace() restricted to ASCII. */
static int py_isspace(unsigned c) {
    return c == ' ' || (c >= 0x09 && c <= 0x0d) || (c >= 0x1c && c <= 0x1f);
}
static unsigned up(unsigned c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }
static unsigned low(unsigned c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

/* Python int(str) on an ASCII token without surrounding whitespace. */
static int py_int(const uint8_t *s, int n, int64_t *out) {
    int i = 0, neg = 0;
    if (n > 0 && (s[0] == '+' || s[0] == '-')) { neg = s[0] == '-'; i = 1; }
    if (i >= n) return ORACLE_E_VALUE;
    uint64_t v = 0;
    int prev_digit = 0;
    for (; i < n; i++) {
        unsigned c = s[i];
        if (c >= '0' && c <= '9') {
            if (v > (UINT64_C(0x7fffffffffffffff) - (c - '0')) / 10) return ORACLE_E_DOMAIN;
            v = v * 10 + (c - '0');
            prev_digit = 1;
        } else if (c == '_' && prev_digit && i + 1 < n && s[i + 1] >= '0' && s[i + 1] <= '9') {
            prev_digit = 0;
        } else {
            return ORACLE_E_VALUE;
        }
    }
    *out = neg ? -(int64_t)v : (int64_t)v;
    return ORACLE_OK;
}

/* pileup.py:276-325.  out must hold n bytes.  Returns the stripped length. */
size_t oracle_strip(const uint8_t *in, size_t n, uint8_t *out) {
    /* pass 1: re.subn(r"\^.", "") -- left to right, non-overlapping; '.' never sees '\n' here */
    size_t m = 0;
    for (size_t i = 0; i < n;) {
        if (in[i] == '^' && i + 1 < n && in[i + 1] != '\n') i += 2;
        else out[m++] = in[i++];
    }
    /* pass 2: finditer(r"[+-](\d+)") on the caret-free string, then splice right to left */
    size_t cap = 16, nm = 0;
    size_t *mstart = malloc(cap * sizeof(size_t)), *mend = malloc(cap * sizeof(size_t));
    uint64_t *mnum = malloc(cap * sizeof(uint64_t));
    for (size_t i = 0; i < m;) {
        if ((out[i] == '+' || out[i] == '-') && i + 1 < m && out[i + 1] >= '0' && out[i + 1] <= '9') {
            size_t j = i + 1;
            uint64_t v = 0;
            while (j < m && out[j] >= '0' && out[j] <= '9') {
                if (v < (UINT64_C(1) << 40)) v = v * 10 + (out[j] - '0');
                j++;
            }
            if (nm == cap) {
                cap *= 2;
                mstart = realloc(mstart, cap * sizeof(size_t));
                mend = realloc(mend, cap * sizeof(size_t));
                mnum = realloc(mnum, cap * sizeof(uint64_t));
            }
            mstart[nm] = i; mend[nm] = j; mnum[nm] = v; nm++;
            i = j;
        } else {
            i++;
        }
    }
    for (size_t k = nm; k-- > 0;) {
        /* bases_str = bases_str[:start] + bases_str[end+num:]  (slice ends clamp) */
        size_t a = mstart[k];
        uint64_t b64 = (uint64_t)mend[k] + mnum[k];
        size_t b = b64 > m ? m : (size_t)b64;
        if (a > m) a = m;
        if (b < a) b = a;
        memmove(out + a, out + b, m - b);
        m -= (b - a);
    }
    free(mstart); free(mend); free(mnum);
    /* pass 3: remove '$' */
    size_t w = 0;
    for (size_t i = 0; i < m; i++) if (out[i] != '$') out[w++] = out[i];
    return w;
}

/* Split like line.rstrip().split(): up to max_tok tokens; returns the total token count. */
static int py_split(const uint8_t *s, int n, int *off, int *len, int max_tok) {
    int nt = 0, i = 0;
    while (i < n) {
        while (i < n && py_isspace(s[i])) i++;
        if (i >= n) break;
        int b = i;
        while (i < n && !py_isspace(s[i])) i++;
        if (nt < max_tok) { off[nt] = b; len[nt] = i - b; }
        nt++;
    }
    return nt;
}

/* pileup.py:209-274 on one line (without its terminator). */
void oracle_parse_record(const uint8_t *line, int n, int min_base_qual, oracle_record *r) {
    memset(r, 0, sizeof(*r));
    for (int i = 0; i < n; i++) if (line[i] >= 0x80) { r->status = ORACLE_E_DOMAIN; return; }
    int off[6], len[6];
    int nt = py_split(line, n, off, len, 6);
    r->ntok = nt;
    if (nt < 2) { r->status = ORACLE_E_INDEX; return; }          /* split_line[1] */
    r->chrom_off = off[0]; r->chrom_len = len[0];
    int st = py_int(line + off[1], len[1], &r->pos);
    if (st) { r->status = st; return; }
    if (nt < 3) { r->status = ORACLE_E_INDEX; return; }          /* split_line[2] */
    if (nt < 4) { r->status = ORACLE_E_INDEX; return; }          /* split_line[3] */
    st = py_int(line + off[3], len[3], &r->raw_depth);
    if (st) { r->status = st; return; }
    if (len[2] != 1) { r->status = ORACLE_E_DOMAIN; return; }
    r->ref = line[off[2]];
    if (r->raw_depth == 0 || nt < 5) return;                      /* empty record  */
    if (nt < 6) { r->status = ORACLE_E_INDEX; return; }          /* split_line[5] */
    r->has_bases = 1;

    uint8_t *b = malloc((size_t)len[4] + 1);
    size_t nb = oracle_strip(line + off[4], (size_t)len[4], b);
    const uint8_t *q = line + off[5];
    size_t nq = (size_t)len[5];
    size_t npair = nb < nq ? nb : nq;                             /* zip() truncates */
    unsigned U = up(r->ref), L = low(r->ref);
    for (size_t i = 0; i < npair; i++) {
        if ((int)q[i] - 33 < min_base_qual) continue;
        unsigned c = b[i];
        if (c == '.') c = U;                                      /* replace('.', REF.upper()) */
        if (c == ',') c = L;                                      /* then replace(',', ref.lower()) */
        r->good_depth++;
        r->total[up(c)]++;
        if (c <= 'Z') { r->fwd[c]++; r->fwd_good_depth++; }
        if (c >= 'a') { r->rev[up(c)]++; r->rev_good_depth++; }
    }
    free(b);
    if (r->good_depth >= 1) {
        /* sorted(items, key=(-freq, base)) */
        int nc = 0;
        for (int c = 0; c < 128; c++) if (r->total[c] > 0) r->common[nc++] = (uint8_t)c;
        for (int i = 1; i < nc; i++) {
            uint8_t x = r->common[i];
            int j = i - 1;
            while (j >= 0 && (r->total[r->common[j]] < r->total[x] ||
                              (r->total[r->common[j]] == r->total[x] && r->common[j] > x))) {
                r->common[j + 1] = r->common[j];
                j--;
            }
            r->common[j + 1] = x;
        }
        r->n_common = nc;
    }
}

/* pileup.py:492-590.  Returns the consensus byte; *fail_mask gets the failed filters. */
uint8_t oracle_call(const oracle_record *r, const oracle_params *p, uint8_t *fail_mask) {
    if (r->n_common == 0) { *fail_mask = FAIL_RAWDPTH; return '-'; }
    uint8_t c = r->common[0];
    int good = r->good_depth, cons = r->total[c], f = r->fwd[c], v = r->rev[c];
    uint8_t m = 0;
    if ((double)cons < (double)good * p->min_cons_freq) m |= FAIL_VARFREQ;
    if (cons < p->min_cons_depth) m |= FAIL_DEPTH;
    if (f < p->min_cons_strand_depth || v < p->min_cons_strand_depth) m |= FAIL_STRDPTH;
    double msb = (double)cons * p->min_cons_strand_bias;
    if ((double)f < msb || (double)v < msb) m |= FAIL_STRBIAS;
    if (c == up(r->ref)) c = r->ref;
    *fail_mask = m;
    return c;
}

/* ---- site lookup: open addressing on (chrom index, pos) ------------------------------- */
typedef struct { int64_t *pos; int32_t *chrom; int32_t *val; size_t cap; } site_map;
static uint64_t mix(uint64_t x) { x ^= x >> 33; x *= UINT64_C(0xff51afd7ed558ccd); x ^= x >> 33; return x; }
static void map_init(site_map *m, size_t n) {
    size_t cap = 16; while (cap < 2 * n + 1) cap *= 2;
    m->cap = cap; m->pos = malloc(cap * 8); m->chrom = malloc(cap * 4); m->val = malloc(cap * 4);
    for (size_t i = 0; i < cap; i++) m->chrom[i] = -1;
}
static void map_free(site_map *m) { free(m->pos); free(m->chrom); free(m->val); }
static int32_t *map_slot(site_map *m, int32_t chrom, int64_t pos, int insert) {
    size_t h = mix((uint64_t)pos * 1315423911u + (uint64_t)chrom) & (m->cap - 1);
    for (;;) {
        if (m->chrom[h] == -1) {
            if (!insert) return NULL;
            m->chrom[h] = chrom; m->pos[h] = pos; m->val[h] = 0; return &m->val[h];
        }
        if (m->chrom[h] == chrom && m->pos[h] == pos) return &m->val[h];
        h = (h + 1) & (m->cap - 1);
    }
}

static int find_contig(const uint8_t *names, const int32_t *name_off, int n_contigs, const uint8_t *s, int n) {
    for (int i = 0; i < n_contigs; i++) {
        int ln = name_off[i + 1] - name_off[i];
        if (ln == n && memcmp(names + name_off[i], s, (size_t)n) == 0) return i;
    }
    return -1;
}

/*
 * call_consensus.py:142-188 for one sample.
 *   text/nbytes            the pileup file contents
 *   names/name_off         contig-name table (concatenated, n_contigs+1 offsets) that the site arrays index
 *   snp_*                  snplist.txt entries in file order (duplicates allowed)
 *   exc_*                  positions of the exclude VCF
 *   parse_all              1 = --vcfAllPos (every line parsed), 0 = only lines at snp/excluded positions
 *   row_out[n_snp]         consensus string in snplist order
 *   line_* (nullable)      per PARSED line, in file order: cell byte, fail mask, position
 * Returns 0, or an ORACLE_E_* for the first line (file order) on which the reference raises;
 * *err_line gets that line's 0-based index.
 */
int oracle_pileup_consensus(const uint8_t *text, size_t nbytes,
                            const uint8_t *names, const int32_t *name_off, int n_contigs,
                            const int32_t *snp_chrom, const int64_t *snp_pos, size_t n_snp,
                            const int32_t *exc_chrom, const int64_t *exc_pos, size_t n_exc,
                            const oracle_params *p, int parse_all,
                            uint8_t *row_out,
                            uint8_t *line_cell, uint8_t *line_fail, int64_t *line_pos, size_t max_lines,
                            size_t *n_parsed_out, size_t *err_line)
{
    site_map map;                       /* val: bit0 = in snplist, bit1 = excluded, bits 8.. = cell */
    map_init(&map, n_snp + n_exc);
    for (size_t i = 0; i < n_snp; i++) *map_slot(&map, snp_chrom[i], snp_pos[i], 1) |= 1;
    for (size_t i = 0; i < n_exc; i++) *map_slot(&map, exc_chrom[i], exc_pos[i], 1) |= 2;

    int rc = ORACLE_OK;
    size_t n_parsed = 0, line_no = 0;
    oracle_record *r = malloc(sizeof(oracle_record));
    size_t i = 0;
    while (i < nbytes) {
        /* universal newlines: "\n", "\r\n" or a lone "\r" end a line */
        size_t e = i;
        while (e < nbytes && text[e] != '\n' && text[e] != '\r') e++;
        size_t next = e < nbytes ? e + 1 : e;
        if (e < nbytes && text[e] == '\r' && e + 1 < nbytes && text[e + 1] == '\n') next = e + 2;
        const uint8_t *line = text + i;
        int n = (int)(e - i);
        int32_t *slot = NULL;
        int want = parse_all;
        if (!parse_all) {
            /* pileup.py:423-429: chrom, pos = split_line[:2]; key in set? */
            int off[2], len[2];
            int dom = 0;
            for (int k = 0; k < n; k++) if (line[k] >= 0x80) dom = 1;
            if (dom) { rc = ORACLE_E_DOMAIN; break; }
            int nt = py_split(line, n, off, len, 2);
            if (nt < 2) { rc = ORACLE_E_UNPACK; break; }
            int64_t pos;
            int st = py_int(line + off[1], len[1], &pos);
            if (st) { rc = st; break; }
            int ci = find_contig(names, name_off, n_contigs, line + off[0], len[0]);
            if (ci >= 0) slot = map_slot(&map, ci, pos, 0);
            want = slot != NULL;
        }
        if (want) {
            oracle_parse_record(line, n, p->min_base_qual, r);
            if (r->status) { rc = r->status; break; }
            if (parse_all) {
                int ci = find_contig(names, name_off, n_contigs, line + r->chrom_off, r->chrom_len);
                slot = ci >= 0 ? map_slot(&map, ci, r->pos, 0) : NULL;
            }
            uint8_t fail, base = oracle_call(r, p, &fail);
            if (slot && (*slot & 2)) fail |= FAIL_REGION;
            uint8_t cell = (fail || base == '*') ? '-' : base;
            if (slot && (*slot & 1)) *slot = (*slot & 3) | ((int32_t)cell << 8);
            if (n_parsed < max_lines) {
                if (line_cell) line_cell[n_parsed] = cell;
                if (line_fail) line_fail[n_parsed] = fail;
                if (line_pos) line_pos[n_parsed] = r->pos;
            }
            n_parsed++;
        }
        line_no++;
        i = next;
    }
    free(r);
    if (rc == ORACLE_OK) {
        for (size_t k = 0; k < n_snp; k++) {
            int32_t v = *map_slot(&map, snp_chrom[k], snp_pos[k], 0);
            uint8_t cell = (uint8_t)(v >> 8);
            row_out[k] = cell ? cell : '-';
        }
    }
    map_free(&map);
    if (n_parsed_out) *n_parsed_out = n_parsed;
    if (err_line) *err_line = line_no;
    return rc;
}

/* Full per-line tally for the fuzz tests: fills a flat int32 array
 *   [status, ntok, pos_lo, pos_hi, raw_depth_lo, raw_depth_hi, ref, good, fwd_good, rev_good, n_common,
 *    cons_byte, fail_mask, then total[128], fwd[128], rev[128], common[128]]                       */
void oracle_line_report(const uint8_t *line, int n, const oracle_params *p, int32_t *out) {
    oracle_record *r = malloc(sizeof(oracle_record));
    oracle_parse_record(line, n, p->min_base_qual, r);
    out[0] = r->status; out[1] = r->ntok;
    out[2] = (int32_t)(r->pos & 0xffffffff); out[3] = (int32_t)(r->pos >> 32);
    out[4] = (int32_t)(r->raw_depth & 0xffffffff); out[5] = (int32_t)(r->raw_depth >> 32);
    out[6] = r->ref; out[7] = r->good_depth; out[8] = r->fwd_good_depth; out[9] = r->rev_good_depth;
    out[10] = r->n_common;
    uint8_t fail = 0, base = '-';
    if (r->status == 0) base = oracle_call(r, p, &fail);
    out[11] = base; out[12] = fail;
    for (int c = 0; c < 128; c++) {
        out[13 + c] = r->total[c]; out[13 + 128 + c] = r->fwd[c]; out[13 + 256 + c] = r->rev[c];
        out[13 + 384 + c] = c < r->n_common ? r->common[c] : 0;
    }
    free(r);
}

/* ---- merge_sites.py:94-116 + utils.py:1056-1070 ---------------------------------------
 * keys[i] = (chrom_rank << 32) | pos with chrom_rank assigned in chrom *string* order, listed sample
 * after sample in sorted-sample-directory order (sample_of[i] = that order's index, non-decreasing).
 * Each sample's keys are a set already.  Output: unique keys ascending, per-key sample count and the
 * CSR list of sample indices in input (= sorted sample) order.  Returns the number of unique keys.  */
typedef struct { uint64_t key; uint32_t sample; uint32_t seq; } ks_t;
static int ks_cmp(const void *a, const void *b) {
    const ks_t *x = a, *y = b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}
size_t oracle_merge_sites(const uint64_t *keys, const uint32_t *sample_of, size_t n,
                          uint64_t *uniq_out, uint32_t *count_out, uint32_t *samples_out) {
    ks_t *a = malloc((n ? n : 1) * sizeof(ks_t));
    for (size_t i = 0; i < n; i++) { a[i].key = keys[i]; a[i].sample = sample_of[i]; a[i].seq = (uint32_t)i; }
    qsort(a, n, sizeof(ks_t), ks_cmp);
    size_t u = 0;
    for (size_t i = 0; i < n; i++) {
        if (i == 0 || a[i].key != a[i - 1].key) { uniq_out[u] = a[i].key; count_out[u] = 0; u++; }
        count_out[u - 1]++;
        samples_out[i] = a[i].sample;
    }
    free(a);
    return u;
}

/* ---- utils.py:1135-1165 over all pairs (distance.py:93-96) ----------------------------
 * matrix: n_rows x n_sites bytes, row-major.  dist_out: n_rows x n_rows int32, symmetric, 0 diagonal. */
static int acgt(unsigned c) { c = up(c); return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
/* One pair (utils.py:1147-1163): positions where both bases are in ACGT (either case) and differ.  Written over
 * per-row codes (0 = not ACGT, else the upper-cased letter) so that the loop has no branch and the compiler can
 * vectorise it -- 5000 samples x 200 k sites is 2.5 x 10^12 byte pairs. */
static int32_t pair_distance(const uint8_t *ca, const uint8_t *cb, size_t n_sites) {
    int32_t d = 0;
    for (size_t k = 0; k < n_sites; k++) d += (int32_t)((ca[k] != 0) & (cb[k] != 0) & (ca[k] != cb[k]));
    return d;
}
/* rows row_begin, row_begin + row_step, ... below row_end of the upper triangle (+ mirror): host threads take
 * interleaved rows so that each gets the same share of the triangle */
void oracle_distance_rows_strided(const uint8_t *matrix, size_t n_rows, size_t n_sites, int32_t *dist_out,
                                  size_t row_begin, size_t row_end, size_t row_step) {
    uint8_t *code = (uint8_t *)malloc(n_rows * n_sites + 1);
    for (size_t i = 0; i < n_rows * n_sites; i++) code[i] = acgt(matrix[i]) ? (uint8_t)up(matrix[i]) : 0;
    for (size_t i = row_begin; i < row_end; i += row_step) {
        for (size_t j = i + 1; j < n_rows; j++) {
            const int32_t d = pair_distance(code + i * n_sites, code + j * n_sites, n_sites);
            dist_out[i * n_rows + j] = d;
            dist_out[j * n_rows + i] = d;
        }
        dist_out[i * n_rows + i] = 0;
    }
    free(code);
}
void oracle_distance_rows(const uint8_t *matrix, size_t n_rows, size_t n_sites, int32_t *dist_out,
                          size_t row_begin, size_t row_end) {
    oracle_distance_rows_strided(matrix, n_rows, n_sites, dist_out, row_begin, row_end, 1);
}
void oracle_distance(const uint8_t *matrix, size_t n_rows, size_t n_sites, int32_t *dist_out) {
    oracle_distance_rows(matrix, n_rows, n_sites, dist_out, 0, n_rows);
}

/* ---- collect_metrics.py:322-329: depth_sum += int(line.split()[3]) over the lines of the pileup file (ValueError /
 * IndexError ignored).  Lines end at '\n', '\r' or "\r\n" (Python's universal newlines); returns ORACLE_OK or
 * ORACLE_E_DOMAIN (a byte >= 0x80, an integer beyond int64) with *error_offset = the offset of that line. */
static int dsum_space(unsigned c) { return c == 0x20u || (c - 9u) < 5u || (c - 0x1cu) < 4u; }
int oracle_depth_sum(const uint8_t *text, size_t n, int64_t *sum_out, uint64_t *lines_out, uint64_t *error_offset) {
    int64_t sum = 0;
    uint64_t lines = 0;
    size_t s = 0;
    *error_offset = ~(uint64_t)0;
    while (s < n) {
        size_t e = s;
        while (e < n && text[e] != '\n' && text[e] != '\r') e++;
        for (size_t k = s; k < e; k++)
            if (text[k] >= 0x80u) { *error_offset = s; return ORACLE_E_DOMAIN; }
        size_t i = s;
        int tok = 0;
        while (i < e) {
            while (i < e && dsum_space(text[i])) i++;
            if (i >= e) break;
            size_t b = i;
            while (i < e && !dsum_space(text[i])) i++;
            if (++tok == 4) {
                /* Python int(): optional sign, digits, single underscores between digits */
                size_t j = b;
                int neg = 0, ok = 1, prev_digit = 0;
                uint64_t v = 0;
                if (text[j] == '+' || text[j] == '-') { neg = text[j] == '-'; j++; }
                if (j >= i) ok = 0;
                for (; ok && j < i; j++) {
                    unsigned c = text[j];
                    if (c >= '0' && c <= '9') {
                        if (v > (0x7fffffffffffffffULL - (c - '0')) / 10) { *error_offset = s; return ORACLE_E_DOMAIN; }
                        v = v * 10 + (c - '0');
                        prev_digit = 1;
                    } else if (c == '_' && prev_digit && j + 1 < i && text[j + 1] >= '0' && text[j + 1] <= '9') {
                        prev_digit = 0;
                    } else ok = 0;
                }
                if (ok) { sum += neg ? -(int64_t)v : (int64_t)v; lines++; }
                break;
            }
        }
        s = e + 1;                      /* ("\r\n": the empty line between the two adds nothing) */
    }
    *sum_out = sum;
    *lines_out = lines;
    return ORACLE_OK;
}
