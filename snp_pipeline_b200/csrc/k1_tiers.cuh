// k1_tiers.cuh -- result plumbing of the pileup kernels (site cells, per-line staging, error report) and the second /
// third parser tiers run on the text where it lies in global memory (k1_pileup.cu's follow-up kernel).
#pragma once
#include "internal.h"
#include "line_fast.cuh"
#include "line_general.cuh"

namespace snpgpu {

// exact per-byte mask (0x80 where the byte equals '\n'), any byte values
__device__ __forceinline__ uint32_t nl_mask(uint32_t w) {
    uint32_t t = w ^ 0x0a0a0a0au;
    return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;
}

// What the second and third tier keep between calls.  It lives in local memory (its address is passed to the
// out-of-line tiers); the first tier never touches it.
struct K1Cold {
    int hint;                     // contig of this lane's previous second-tier line (line_fast.cuh moves it)
    uint32_t n_parsed, n_general;
};

__device__ __forceinline__ void k1_report(const PileupArgs &a, unsigned long long goff, int code) {
    atomicMax(&a.st->first_error_inv, ~((goff << 8) | (unsigned long long)code));
}

// call_consensus.py:165-176: Region failure, '-' substitution, keep the cell for the snplist gather.
// Returns the line's result word: matrix cell | fail mask << 8.
__device__ __forceinline__ uint16_t k1_cell_flags(const PileupArgs &a, unsigned base_ch, unsigned fail, int32_t site,
                                                  unsigned flags, unsigned long long goff);
__device__ __forceinline__ uint16_t k1_cell(const PileupArgs &a, unsigned base_ch, unsigned fail, int32_t site,
                                            unsigned long long goff) {
    return k1_cell_flags(a, base_ch, fail, site, site >= 0 ? a.sites.flags[site] : 0u, goff);
}
// (flags: SITE_* of the site, 0 when the line is at none)
__device__ __forceinline__ uint16_t k1_cell_flags(const PileupArgs &a, unsigned base_ch, unsigned fail, int32_t site,
                                                  unsigned flags, unsigned long long goff) {
    if (flags & SITE_EXCLUDED) fail |= FAIL_REGION;
    unsigned cell = (fail || base_ch == '*') ? (unsigned)'-' : base_ch;
    if (flags & SITE_SNP) atomicMax(&a.site_cells[site], ((goff + 1ull) << 8) | (unsigned long long)cell);
    if (a.rec_off) {                                          // the VCF pass wants to know which lines were parsed
        const unsigned long long k = atomicAdd(a.rec_count, 1ull);
        if (k < a.rec_cap) a.rec_off[k] = goff;
    }
    return (uint16_t)(cell | (fail << 8));
}

// Per-line results (all-positions mode with line_out): the k-th line of lane L of a tile (the lane that owns the '\n' in
// front of it) goes to slot [tile][k][L] of the staging array -- the 32 lanes of a warp-step write 64 contiguous bytes;
// the lanes' line counts and the slots are put into file order afterwards by k1_tile_prefix_kernel + k1_lines_kernel.
// No tile has to know where it starts while the pileup kernel runs.  A lane that owns more than K1_LCAP lines (lines of a
// few bytes) appends the rest to an overflow list of (tile * 32 + lane, k, result) entries.
__device__ __forceinline__ void k1_store_slot(const PileupArgs &a, unsigned long long tile_lane, uint32_t k, uint16_t v) {
    if (k < (uint32_t)K1_LCAP) {
        a.stage[((tile_lane >> 5) * (unsigned long long)K1_LCAP + k) * 32ull + (tile_lane & 31ull)] = v;
    } else {
        const unsigned long long n = atomicAdd(&a.st->over_used, 1ull);
        if (n < a.over_cap) a.over[n] = (tile_lane << 32) | ((unsigned long long)(k & 0xffffu) << 16) | (unsigned long long)v;
    }
}

// the line that starts at file offset goff and is the k-th of the lane that owns the byte in front of it
__device__ __forceinline__ void k1_emit(const PileupArgs &a, unsigned base_ch, unsigned fail, int32_t site,
                                        unsigned long long goff, uint32_t k) {
    const uint16_t v = k1_cell(a, base_ch, fail, site, goff);
    if (a.stage) k1_store_slot(a, (goff ? goff - 1ull : 0ull) / (unsigned long long)K1_LANE_BYTES, k, v);   // (lanes tile the text)
}

// third tier: the exact any-input parser, on the text where it lies in global memory
__device__ __noinline__ void k1_general(const PileupArgs &a, K1Cold &cs, unsigned long long goff, uint32_t line_idx) {
    const uint8_t *line = a.text + goff;
    unsigned long long room = a.nbytes - goff;
    int64_t n = 0;
    bool lone_cr = false;
    while ((unsigned long long)n < room && line[n] != '\n') {
        if (line[n] == '\r' && (unsigned long long)(n + 1) < room && line[n + 1] != '\n') lone_cr = true;
        n++;
    }
    cs.n_general++;
    if (lone_cr) { k1_report(a, goff, ST_LONECR); return; }   // classic-Mac line end: the caller normalises and reruns
    LineCall r;
    const bool all = a.mode == SNPGPU_MODE_ALL;
    int32_t site = -1;
    if (!all) {
        general_key(line, n, &r);                             // pileup.py:423-427
        if (r.status) { k1_report(a, goff, r.status); return; }
        int cid = contig_find(a.sites, line + r.chrom_off, r.chrom_len);
        site = site_find(a.sites, cid, r.pos);
        if (site < 0) return;
    }
    general_line(line, n, a.p, nullptr, 0, &r);
    if (r.status == ST_NEED_ARENA) {
        unsigned long long want = ((unsigned long long)r.bases_len + 15ull) & ~15ull;
        unsigned long long off = atomicAdd(&a.arena_st->arena_used, want);
        if (off + want > a.arena_cap) { atomicExch(&a.arena_st->arena_overflow, 1u); return; }
        general_line(line, n, a.p, a.arena + off, r.bases_len, &r);
    }
    if (r.status) { k1_report(a, goff, r.status); return; }
    if (all) {
        int cid = contig_find(a.sites, line + r.chrom_off, r.chrom_len);
        site = site_find(a.sites, cid, r.pos);
    }
    k1_emit(a, r.base, r.fail, site, goff, line_idx);
    cs.n_parsed++;
}

// offset from s of the first '\n' in buf[s, s + cap), cap when there is none; buf 4-byte aligned, whole words are
// read only where all four bytes lie inside the range
__device__ __forceinline__ uint32_t k1_find_nl(const uint8_t *buf, uint32_t s, uint32_t cap) {
    uint32_t i = s;
    const uint32_t end = s + cap;
    for (; i < end && (i & 3u); i++) if (buf[i] == '\n') return i - s;
    for (; i + 4u <= end; i += 4u) {
        const uint32_t m = nl_mask(*reinterpret_cast<const uint32_t *>(buf + i));
        if (m) return i + ((uint32_t)ctz32(m) >> 3) - s;
    }
    for (; i < end; i++) if (buf[i] == '\n') return i - s;
    return cap;
}

// second tier: exact tallies (line_fast.cuh) on the text in global memory (or the thread's staged copy); returns true when the line has to
// go on to k1_general().  The words line_fast reads may reach 7 bytes past the line end, so the last lines of
// the text are left to k1_general(), which reads byte by byte.
template <bool HAS_QUAL, bool ALL>
__device__ __noinline__ bool k1_detail(const PileupArgs &a, K1Cold &cs, unsigned long long goff, uint32_t line_idx,
                                       uint32_t len_hint, const uint8_t *staged) {
    const unsigned long long room = a.nbytes - goff;
    const unsigned long long abase = goff & ~15ull;
    const uint8_t *buf = staged ? staged : a.text + abase;    // (staged: the thread's copy of [abase, line end + 48) in shared memory)
    const uint32_t s = (uint32_t)(goff - abase);
    const uint32_t cap = room < 65536ull ? (uint32_t)room : 65536u;
    const uint32_t n = len_hint && len_hint < cap ? len_hint : k1_find_nl(buf, s, cap);
    if (n == cap || (unsigned long long)n + 8ull > room) return true;          // very long, or at the end of the text
    FastLine fl;
    const int st = fast_line<HAS_QUAL>(buf, s, s + n, a.sites, cs.hint, a.p, ALL, &fl);
    if (st == ST_OK) {
        k1_emit(a, fl.base, fl.fail, fl.site, goff, line_idx);
        cs.n_parsed++;
    }
    return st == ST_FALLBACK;
}

// ---- queue entries of the follow-up kernel ------------------------------------------------------------------------
// entry: index of the line among its lane's lines << 50 | length hint << 38 | file offset.  The hint is the line's length
// without its '\n' when the pileup kernel found it (and it is below 4096), else 0: look for the '\n'.
__device__ __forceinline__ unsigned long long k1_entry(uint32_t line_idx, uint32_t len_hint, unsigned long long goff) {
    return ((unsigned long long)(line_idx & 0x3fffu) << 50) | ((unsigned long long)(len_hint < 4096u ? len_hint : 0u) << 38) | goff;
}
__device__ __forceinline__ unsigned long long k1_entry_goff(unsigned long long e) { return e & ((1ull << 38) - 1ull); }
__device__ __forceinline__ uint32_t k1_entry_len(unsigned long long e) { return (uint32_t)(e >> 38) & 0xfffu; }
__device__ __forceinline__ uint32_t k1_entry_idx(unsigned long long e) { return (uint32_t)(e >> 50); }

}  // namespace snpgpu
