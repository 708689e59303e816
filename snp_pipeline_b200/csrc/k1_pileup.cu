// k1_pileup.cu -- K1: pileup text in HBM -> consensus cells.
//
// Replaces the per-sample hot loop of the reference (call_consensus.py:161-188): pileup.Reader.__iter__
// (pileup.py:408-429), pileup.Record (pileup.py:209-325) and ConsensusCaller.call_consensus (pileup.py:492-590).
//
// Shape of the kernel: every WARP is its own pipeline -- no block-wide barrier anywhere.
//   * a warp strides over 8.5 KiB tiles of the text; a tile owns the lines whose preceding '\n' lies inside it
//     (the first line of the file belongs to tile 0);
//   * the tile (+1 KiB of look-ahead so the last owned line is complete) is staged into the warp's slice of
//     shared memory by ONE 1-D bulk async copy (TMA engine, mbarrier completion); the other resident warps
//     compute while this one waits;
//   * scan: lane L owns K1_LANE_BYTES consecutive bytes of the tile (a stride of 4 mod 8 words, so the 16-byte
//     loads of a quarter warp fall into disjoint banks), tests them for '\n' with SWAR arithmetic and lists its
//     hits; ONE warp prefix sum over the per-lane counts then orders all line starts of the tile;
//   * parse: one lane per line, 32 lines in lock step, through the first-tier parser (line_quick.cuh);
//   * the few lines it declines are queued per warp (by file offset) and, whenever 32 have piled up, run densely
//     through the exact-tally parser (line_fast.cuh) and, from there, the any-input parser (line_general.cuh),
//     both on the text where it lies in global memory (an L2 hit);
//   * results: an atomicMax per hit site keeps the LAST line of a position in file order (the dict overwrite
//     of call_consensus.py:169-176); in all-positions mode one uint16 per line goes straight to its file-order
//     slot: the index of a tile's first line comes from a decoupled look-back over the tiles' line counts.
// Algorithmic traffic: every text byte read once (+12 % look-ahead re-read, an L2 hit), 2 B written per line.
#include "internal.h"
#include <stdio.h>
#include "line_fast.cuh"
#include "line_quick.cuh"
#include "line_general.cuh"

namespace snpgpu {

// Tuning builds (profiles/variant.sh prof -DK1_PROF): cycles lane 0 of every warp spends per phase, summed over all
// warps into the unused tail of the status block and printed by the finish kernel.  Compiled out of the product.
#ifdef K1_PROF
#define PROF_DECL long long prof_t = clock64(); unsigned long long prof_acc[16] = {0}
#define PROF(i) do { const long long prof_n = clock64(); prof_acc[i] += (unsigned long long)(prof_n - prof_t); prof_t = prof_n; } while (0)
#define PROF_FLUSH(st) do { if (lane == 0) for (int i = 0; i < 16; i++) atomicAdd(reinterpret_cast<unsigned long long *>(st) + 8 + i, prof_acc[i]); } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_FLUSH(st)
#endif

struct K1Warp {                                 // one warp's slice of shared memory
    alignas(128) uint8_t buf[K1_TILE + K1_LOOK + K1_PAD];
    uint16_t starts[K1_WCAP];                   // line starts of the current pass, file order: chunk << 5 | flag bit
    uint16_t lanehits[32 * K1_LHCAP];           // the same as each lane found them during the scan, K1_LHCAP per lane
    alignas(16) uint32_t cname[K1_NAMEW];       // name + tab of the contig the warp expects (ContigCache::name4) ...
    alignas(16) uint32_t cmask[K1_NAMEW];       // ... and which of its bytes count (ContigCache::mask4)
    unsigned long long dq[K1_QCAP];             // lines for line_fast.cuh (k1_entry)
    unsigned long long gq[K1_QCAP];             // lines for line_general.cuh, same encoding
    alignas(8) uint64_t bar;
};
static_assert(K1_PAD >= (int)QUICK_PAD && K1_NAMEW % 4 == 0, "line_quick.cuh preconditions");
static_assert(sizeof(K1Warp) * K1_WARPS * K1_CTAS_PER_SM + 1024 * K1_CTAS_PER_SM <= 228 * 1024, "shared memory per SM");
size_t k1_smem_bytes() { return sizeof(K1Warp) * K1_WARPS; }

// exact per-byte mask (0x80 where the byte equals '\n'), any byte values
__device__ __forceinline__ uint32_t nl_mask(uint32_t w) {
    uint32_t t = w ^ 0x0a0a0a0au;
    return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;
}

// non-zero in bit 7 of some byte iff the word holds a '\r' (plus borrow artefacts above one: only an "any" test)
__device__ __forceinline__ uint32_t cr_any(uint32_t w) {
    uint32_t t = w ^ 0x0d0d0d0du;
    return (t - 0x01010101u) & ~t;
}

// Is there a '\r' in buf[off, off + 16) that is not followed by '\n'?  (rare path: only when cr_any fired)
__device__ __noinline__ bool lone_cr_in_chunk(const uint8_t *buf, uint32_t off) {
    const uint4 v = *reinterpret_cast<const uint4 *>(buf + off);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    bool lone = false;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t t = w[k] ^ 0x0d0d0d0du;
        uint32_t m = ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;
        while (m) {
            uint32_t b = (uint32_t)ctz32(m) >> 3;
            m &= m - 1u;
            if (buf[off + 4u * k + b + 1u] != '\n') lone = true;
        }
    }
    return lone;
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

// Per-line results (all-positions mode with line_out): a line's result goes to slot `line_idx` of its tile's row of
// the staging array; the tiles' line counts (PileupArgs::tile_lines) and the rows are put into file order afterwards
// by k1_tile_prefix_kernel + k1_lines_kernel.  No tile has to know where it starts while the pileup kernel runs.
// A tile that owns more than K1_STAGE_CAP lines (only the byte-wise path can: lines of a few bytes) appends the rest
// to an overflow list of (tile, index, result) entries.
__device__ __forceinline__ void k1_store_line(const PileupArgs &a, unsigned long long tile, uint32_t line_idx, uint16_t v) {
    if (line_idx < (uint32_t)K1_STAGE_CAP) {
        a.stage[tile * (unsigned long long)K1_STAGE_CAP + line_idx] = v;
    } else {
        const unsigned long long k = atomicAdd(&a.st->over_used, 1ull);
        if (k < a.over_cap) a.over[k] = (tile << 32) | ((unsigned long long)line_idx << 16) | (unsigned long long)v;
    }
}

// the line that starts at file offset goff and is the line_idx-th of its tile (the tile holding the byte in front of it)
__device__ __forceinline__ void k1_emit(const PileupArgs &a, unsigned base_ch, unsigned fail, int32_t site,
                                        unsigned long long goff, uint32_t line_idx) {
    const uint16_t v = k1_cell(a, base_ch, fail, site, goff);
    if (a.line_out) k1_store_line(a, goff ? (goff - 1ull) / (unsigned long long)K1_TILE : 0ull, line_idx, v);
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
        unsigned long long off = atomicAdd(&a.st->arena_used, want);
        if (off + want > a.arena_cap) { atomicExch(&a.st->arena_overflow, 1u); return; }
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

// second tier: exact tallies (line_fast.cuh) on the text in global memory; returns true when the line has to
// go on to k1_general().  The words line_fast reads may reach 7 bytes past the line end, so the last lines of
// the text are left to k1_general(), which reads byte by byte.
template <bool HAS_QUAL, bool ALL>
__device__ __noinline__ bool k1_detail(const PileupArgs &a, K1Cold &cs, unsigned long long goff, uint32_t line_idx,
                                       uint32_t len_hint) {
    const unsigned long long room = a.nbytes - goff;
    const unsigned long long abase = goff & ~15ull;
    const uint8_t *buf = a.text + abase;
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

// ---- the per-warp queues (all 32 lanes call these together; the fill counts are warp-uniform) -------------
// entry: line index in its tile << 50 | length hint << 38 | file offset.  The hint is the line's length without its
// '\n' when the scan's list knows where the next line starts (and it is below 4096), else 0: look for the '\n'.
static_assert(K1_TILE <= (1 << 14), "queue entries keep a line's index within its tile in 14 bits");
__device__ __forceinline__ unsigned long long k1_entry(uint32_t line_idx, uint32_t len_hint, unsigned long long goff) {
    return ((unsigned long long)line_idx << 50) | ((unsigned long long)(len_hint < 4096u ? len_hint : 0u) << 38) | goff;
}
__device__ __forceinline__ unsigned long long k1_entry_goff(unsigned long long e) { return e & ((1ull << 38) - 1ull); }
__device__ __forceinline__ uint32_t k1_entry_len(unsigned long long e) { return (uint32_t)(e >> 38) & 0xfffu; }
__device__ __forceinline__ uint32_t k1_entry_idx(unsigned long long e) { return (uint32_t)(e >> 50); }
__device__ __forceinline__ uint32_t k1_push(unsigned long long *q, uint32_t n_q, int lane, bool want,
                                            unsigned long long entry) {
    const uint32_t b = __ballot_sync(0xffffffffu, want);
    if (b == 0u) return n_q;
    if (want) q[n_q + (uint32_t)__popc(b & ((1u << lane) - 1u))] = entry;
    __syncwarp();
    return n_q + (uint32_t)__popc(b);
}

// runs queued lines through k1_general(), 32 at a time, while at least 32 wait (all of them when flush)
__device__ __noinline__ uint32_t k1_drain_general(const PileupArgs &a, K1Warp &sm, K1Cold &cs, int lane, uint32_t n_gq,
                                                  bool flush) {
    while (n_gq >= 32u || (flush && n_gq > 0u)) {
        const uint32_t take = n_gq < 32u ? n_gq : 32u;
        n_gq -= take;
        unsigned long long e = 0;
        const bool mine = (uint32_t)lane < take;
        if (mine) e = sm.gq[n_gq + (uint32_t)lane];
        __syncwarp();
        if (mine) k1_general(a, cs, k1_entry_goff(e), k1_entry_idx(e));
        __syncwarp();
    }
    return n_gq;
}

// the same for k1_detail(); what it declines moves to the general queue.  Returns n_dq | n_gq << 16.
template <bool HAS_QUAL, bool ALL>
__device__ __noinline__ uint32_t k1_drain_detail(const PileupArgs &a, K1Warp &sm, K1Cold &cs, int lane, uint32_t n_dq,
                                                 uint32_t n_gq, bool flush) {
    while (n_dq >= 32u || (flush && n_dq > 0u)) {
        const uint32_t take = n_dq < 32u ? n_dq : 32u;
        n_dq -= take;
        unsigned long long e = 0;
        const bool mine = (uint32_t)lane < take;
        if (mine) e = sm.dq[n_dq + (uint32_t)lane];
        __syncwarp();
        bool more = false;
        if (mine) more = k1_detail<HAS_QUAL, ALL>(a, cs, k1_entry_goff(e), k1_entry_idx(e), k1_entry_len(e));
        __syncwarp();
        n_gq = k1_push(sm.gq, n_gq, lane, more, e);
        n_gq = k1_drain_general(a, sm, cs, lane, n_gq, false);
    }
    return n_dq | (n_gq << 16);
}

// A tile with a byte >= 0x80 or a lone CR in its window: every line goes to k1_general().  Byte-wise on purpose
// (no SWAR assumption holds here).  Returns the number of lines the tile owns | n_gq << 16.
__device__ __noinline__ uint32_t k1_slow_tile(const PileupArgs &a, K1Warp &sm, K1Cold &cs, int lane, uint32_t n_gq,
                                              int tile, unsigned long long base, uint32_t wlen) {
    const uint32_t lo = (uint32_t)lane * K1_LANE_BYTES;
    uint32_t hi = lo + K1_LANE_BYTES;
    if (hi > wlen) hi = wlen;
    uint32_t cnt = 0;
    for (uint32_t i = lo; i < hi; i++)
        if (sm.buf[i] == '\n' && i + 1u < wlen) cnt++;
    const bool first = tile == 0 && lane == 0 && wlen > 0;
    if (first) cnt++;
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (a.line_out && lane == 0) { a.tile_lines[tile] = total; atomicAdd(&a.group_lines[tile / K1_ORDER_TILES], (unsigned long long)total); }
    uint32_t idx = incl - cnt, cur = lo;
    bool pending_first = first;
    while (__any_sync(0xffffffffu, cnt > 0u)) {
        unsigned long long entry = 0;
        const bool want = cnt > 0u;
        if (want) {
            uint32_t s = 0;
            if (pending_first) pending_first = false;
            else {
                while (!(sm.buf[cur] == '\n' && cur + 1u < wlen)) cur++;
                s = ++cur;
            }
            entry = k1_entry(idx, 0u, base + s);
            idx++; cnt--;
        }
        n_gq = k1_push(sm.gq, n_gq, lane, want, entry);
        n_gq = k1_drain_general(a, sm, cs, lane, n_gq, false);
    }
    return total | (n_gq << 16);
}

// newline flags of the 16 bytes at buf[off ..]: bit 8*b + w  <->  byte 4*w + b.  Needs every byte < 0x80.
__device__ __forceinline__ uint32_t k1_chunk_mask(const uint4 v, uint32_t off, uint32_t wlen) {
    const uint32_t NL = 0x0a0a0a0au, K = 0x7f7f7f7fu, H = 0x80808080u;
    const uint32_t x0 = (v.x ^ NL) + K, x1 = (v.y ^ NL) + K, x2 = (v.z ^ NL) + K, x3 = (v.w ^ NL) + K;
    uint32_t m = ((~x0 & H) >> 7) | ((~x1 & H) >> 6) | ((~x2 & H) >> 5) | ((~x3 & H) >> 4);
    if (off + 17u > wlen) {                                   // last chunk of the text: a start must be < wlen
#pragma unroll 1
        for (uint32_t pos = wlen - 1u - off; pos < 16u; pos++) m &= ~(1u << (8u * (pos & 3u) + (pos >> 2)));
    }
    return m;
}

// appends the line starts flagged in m (chunk number `chunk`) to the pass's list as (chunk << 5 | flag bit)
// (slots at or above cap are counted but not written)
__device__ __forceinline__ void k1_list_hits(uint16_t *list, uint32_t cap, uint32_t m, uint32_t chunk, uint32_t &idx) {
    if (m == 0u) return;
    const uint32_t chunk5 = chunk << 5;
    if ((m & (m - 1u)) == 0u) {
        if (idx < cap) list[idx] = (uint16_t)(chunk5 | (uint32_t)ctz32(m));
        idx++;
    } else {                                                  // lines shorter than 16 bytes: in byte order
#pragma unroll 1
        for (uint32_t pos = 0; pos < 16u; pos++) {
            const uint32_t f = 8u * (pos & 3u) + (pos >> 2);
            if ((m >> f) & 1u) {
                if (idx < cap) list[idx] = (uint16_t)(chunk5 | f);
                idx++;
            }
        }
    }
}

// the list of a later pass (more than K1_WCAP lines in the tile: very short lines): scans the window again
__device__ __noinline__ void k1_relist(K1Warp &sm, int lane, uint32_t wlen, uint32_t idx) {
    for (int j = 0; j < K1_LANE_CHUNKS; j++) {
        const uint32_t off = (uint32_t)lane * K1_LANE_BYTES + (uint32_t)j * 16u;
        if (off >= wlen) break;
        const uint32_t m = k1_chunk_mask(*reinterpret_cast<const uint4 *>(sm.buf + off), off, wlen);
        k1_list_hits(sm.starts, K1_WCAP, m, (uint32_t)lane * K1_LANE_CHUNKS + (uint32_t)j, idx);
    }
}

// One first-tier step of a pass: the lane's line l of the pass (have: it has one).  Returns true when the line has
// to go on to the second tier; s = offset of the line in the window.
struct K1Pass {
    unsigned long long base;                    // file offset of the window
    unsigned long long tile;
    uint32_t wlen, done;
    bool eof;
};
template <bool ALL>
__device__ __forceinline__ bool k1_quick_step(const PileupArgs &a, K1Warp &sm, const ContigCache &cc, const K1Pass &ps,
                                              bool have, uint32_t l, uint32_t code, uint32_t &s, uint32_t &n_parsed) {
    const uint32_t line_idx = ps.done + l;
    s = 0;                                                    // code: chunk << 5 | flag bit 8*b + w  ->  byte 4*w + b of the chunk
    if (have && code != 0xffffu) s = (code >> 5) * 16u + (code & 7u) * 4u + ((code >> 3) & 3u) + 1u;
    bool to_detail = false;
    if (have) {
        to_detail = true;
        QuickLine q;
        const int st = quick_line(sm.buf, s, ps.wlen, a.sites, cc, a.p, ALL, &q);
        if (st == ST_SKIP) to_detail = false;
        else if (st == ST_OK && !(q.end == ps.wlen && !ps.eof)) {     // (a line that leaves the window goes on)
            const uint16_t v = k1_cell_flags(a, q.base, q.fail, q.site, q.flags, ps.base + s);
            if (ALL && a.line_out) k1_store_line(a, ps.tile, line_idx, v);
            n_parsed++;
            to_detail = false;
        }
    }
    return to_detail;
}

// HAS_QUAL: a minimum base quality is set (call_consensus -q > 0): every line goes straight to line_fast.cuh,
// which pairs each base with its quality.  ALL: all-positions mode (compile-time so that the scan can drop its
// '\r' test: there the parsers look at every byte of every line themselves).
template <bool HAS_QUAL, bool ALL>
__global__ void __launch_bounds__(K1_THREADS, K1_CTAS_PER_SM) k1_pileup_kernel(const __grid_constant__ PileupArgs a) {
    extern __shared__ __align__(128) uint8_t k1_smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = (int)(threadIdx.x >> 5);
    K1Warp &sm = reinterpret_cast<K1Warp *>(k1_smem_raw)[warp];
    if (lane == 0) {
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t parity = 0, n_dq = 0, n_gq = 0, n_parsed = 0;
    PROF_DECL;
    unsigned long long n_lines = 0;
    K1Cold cs{0, 0u, 0u};
    constexpr bool CHECK_CR = !ALL || HAS_QUAL;
    constexpr uint32_t H = 0x80808080u;
    const int n_gwarps = gridDim.x * K1_WARPS;
    ContigCache cc;
    contig_cache_load(a.sites, 0, sm.cname, sm.cmask, K1_NAMEW, &cc);   // every lane writes the same words
    static_assert(K1_PAD == 32, "one sentinel byte per lane");
    sm.buf[K1_TILE + K1_LOOK + lane] = (uint8_t)'\n';        // the '\n' sentinels behind a whole window: no copy ever reaches them
    __syncwarp();
    int ticket = 0;                                           // lane 0: the next tile, when ticket_taken
    bool drained = false;                                     // queued lines ran since the contig cache was last checked
    bool ticket_taken = false;                                // (warp-uniform)
    for (;;) {
        // tiles in increasing order; usually taken while the tile before is parsed
        if (!ticket_taken && lane == 0) ticket = (int)atom_inc_u32(&a.st->next_tile);
        const int tile = __shfl_sync(0xffffffffu, ticket, 0);
        PROF(0);                                              // (waiting for the warp's slowest lane and the ticket)
        ticket_taken = false;
        if (tile >= a.n_tiles) break;
        if (drained) {   // a lane met another contig (line_fast.cuh moved its hint): the warp follows the last such lane
            drained = false;
            const int hint = cs.hint;
            const uint32_t moved = __ballot_sync(0xffffffffu, hint != cc.cid);
            if (moved) {
                const int nh = __shfl_sync(0xffffffffu, hint, 31 - __clz((int)moved));
                cs.hint = nh;
                __syncwarp();
                contig_cache_load(a.sites, nh, sm.cname, sm.cmask, K1_NAMEW, &cc);
            }
        }
        const unsigned long long base = (unsigned long long)tile * K1_TILE;
        const unsigned long long left = a.nbytes - base;
        const uint32_t wlen = left < (unsigned long long)(K1_TILE + K1_LOOK) ? (uint32_t)left : (uint32_t)(K1_TILE + K1_LOOK);
        const uint32_t bulk = wlen & ~15u;
        const bool eof = left <= (unsigned long long)(K1_TILE + K1_LOOK);
        // ---- stage the window ---------------------------------------------------------------------
        __syncwarp();                                         // every lane is done with the previous window
        if (lane == 0 && bulk) {
            fence_proxy_async();
            mbar_arrive_expect_tx(&sm.bar, bulk);
            bulk_g2s(sm.buf, a.text + base, bulk, &sm.bar);
            const unsigned long long nbase = base + (unsigned long long)n_gwarps * K1_TILE;   // the tile a warp will take about one round from now -> L2
            if (nbase < a.nbytes) {
                const unsigned long long nleft = a.nbytes - nbase;
                const uint32_t nb = (nleft < (unsigned long long)(K1_TILE + K1_LOOK) ? (uint32_t)nleft : (uint32_t)(K1_TILE + K1_LOOK)) & ~15u;
                if (nb) bulk_prefetch_l2(a.text + nbase, nb);
            }
        }
        if (wlen != (uint32_t)(K1_TILE + K1_LOOK)) {          // the text's last windows: the bytes behind the last whole 16 and
            for (uint32_t j = bulk + (uint32_t)lane; j < wlen + (uint32_t)K1_PAD; j += 32u)     // the sentinels behind them
                sm.buf[j] = j < wlen ? a.text[base + j] : (uint8_t)'\n';
        }                                                     // (a whole window ends at the sentinels written once, below)
        PROF(1);
        PROF(2);
        if (n_dq >= (uint32_t)K1_DRAIN_AT) {                  // queued lines, while this window loads
            drained = true;
            const uint32_t r = k1_drain_detail<HAS_QUAL, ALL>(a, sm, cs, lane, n_dq, n_gq, true);
            n_dq = r & 0xffffu; n_gq = r >> 16;
        }
        PROF(3);
        if (bulk) { mbar_wait(&sm.bar, parity); parity ^= 1u; }
        PROF(4);
        __syncwarp();
        // ---- scan: every lane lists the line starts of its chunks as it finds them -------------------
        uint32_t hi_acc = 0, cr_acc = 0, cnt = 0;
        uint16_t *myhits = sm.lanehits + lane * K1_LHCAP;
        if (wlen > (uint32_t)K1_TILE) {                       // a whole tile: no chunk needs a bounds test
            uint32_t multi = 0;
            constexpr int B = 7;                              // chunks loaded together, ahead of the stores below
#pragma unroll 1
            for (int j0 = 0; j0 < K1_LANE_CHUNKS; j0 += B) {
                uint4 vv[B];
#pragma unroll
                for (int u = 0; u < B; u++)
                    if (j0 + u < K1_LANE_CHUNKS)
                        vv[u] = *reinterpret_cast<const uint4 *>(sm.buf + (uint32_t)lane * K1_LANE_BYTES + (uint32_t)(j0 + u) * 16u);
#pragma unroll
                for (int u = 0; u < B; u++) {
                    if (j0 + u >= K1_LANE_CHUNKS) continue;
                    const int j = j0 + u;
                    const uint4 v = vv[u];
                    hi_acc |= v.x | v.y | v.z | v.w;
                    if (CHECK_CR) cr_acc |= cr_any(v.x) | cr_any(v.y) | cr_any(v.z) | cr_any(v.w);
                    const uint32_t NL = 0x0a0a0a0au, K = 0x7f7f7f7fu;
                    const uint32_t x0 = (v.x ^ NL) + K, x1 = (v.y ^ NL) + K, x2 = (v.z ^ NL) + K, x3 = (v.w ^ NL) + K;
                    const uint32_t m = ((~x0 & H) >> 7) | ((~x1 & H) >> 6) | ((~x2 & H) >> 5) | ((~x3 & H) >> 4);
                    // no branch: the code is stored whether the chunk has a start or not -- a miss lands on the slot the
                    // next hit overwrites, and the last slot is never counted
                    const uint32_t code = (((uint32_t)lane * K1_LANE_CHUNKS + (uint32_t)j) << 5) | ((uint32_t)__ffs((int)m) - 1u);
                    myhits[cnt < (uint32_t)(K1_LHCAP - 1) ? cnt : (uint32_t)(K1_LHCAP - 1)] = (uint16_t)code;
                    cnt += m != 0u ? 1u : 0u;
                    multi |= m & (m - 1u);                    // two starts within 16 bytes
                }
            }
            if (multi) hi_acc |= 0x80u;                       // tiny lines: the byte-wise path sorts them out
        } else {
#pragma unroll 1
            for (int j = 0; j < K1_LANE_CHUNKS; j++) {
                const uint32_t off = (uint32_t)lane * K1_LANE_BYTES + (uint32_t)j * 16u;
                if (off >= wlen) break;
                const uint4 v = *reinterpret_cast<const uint4 *>(sm.buf + off);
                hi_acc |= v.x | v.y | v.z | v.w;
                if (CHECK_CR) cr_acc |= cr_any(v.x) | cr_any(v.y) | cr_any(v.z) | cr_any(v.w);
                k1_list_hits(myhits, K1_LHCAP, k1_chunk_mask(v, off, wlen), (uint32_t)lane * K1_LANE_CHUNKS + (uint32_t)j, cnt);
            }
        }
        if (cnt > (uint32_t)(K1_LHCAP - 1)) hi_acc |= 0x80u;  // a crowd of short lines: likewise
        PROF(5);
        // look-ahead bytes: only the odd-byte tests
        for (uint32_t off = (uint32_t)K1_TILE + (uint32_t)lane * 16u; off < wlen; off += 512u) {
            const uint4 v = *reinterpret_cast<const uint4 *>(sm.buf + off);
            hi_acc |= v.x | v.y | v.z | v.w;
            if (CHECK_CR) cr_acc |= cr_any(v.x) | cr_any(v.y) | cr_any(v.z) | cr_any(v.w);
        }
        if (CHECK_CR && __any_sync(0xffffffffu, (cr_acc & H) != 0u)) {   // CRs present: fine when each is followed by LF
            for (uint32_t off = (uint32_t)lane * 16u; off < wlen; off += 512u)
                if (lone_cr_in_chunk(sm.buf, off)) hi_acc |= 0x80u;
        }
        if (__any_sync(0xffffffffu, (hi_acc & H) != 0u)) {    // odd bytes around: every line takes the exact path
            const uint32_t r = k1_slow_tile(a, sm, cs, lane, n_gq, tile, base, wlen);
            n_gq = r >> 16;
            n_lines += r & 0xffffu;
            continue;
        }
        const bool file_start = tile == 0 && lane == 0 && wlen > 0;
        if (file_start) cnt++;                                // the first line of the file
        if (!ALL) {
            // ---- filter mode: no per-line output, so no file order is needed and a line is nearly always done after
            //      its key columns -- every lane walks its own list of starts, no prefix sum, no list copy, no sort ----
            const uint32_t n_mine = cnt;                      // (lane 0 of tile 0: the file's first line comes first)
            const uint32_t n_steps = __reduce_max_sync(0xffffffffu, n_mine);
            n_lines += __reduce_add_sync(0xffffffffu, n_mine);
            K1Pass ps;
            ps.base = base; ps.tile = (unsigned long long)tile; ps.wlen = wlen; ps.done = 0; ps.eof = eof;
            if (n_steps) {                                    // the next ticket, its latency hidden behind the parse
                if (lane == 0) ticket = (int)atom_inc_u32(&a.st->next_tile);
                ticket_taken = true;
            }
            const uint32_t shift = file_start ? 1u : 0u;
            for (uint32_t k = 0; k < n_steps; k++) {
                const bool have = k < n_mine;
                uint32_t code = 0, next = 0xffffu;
                if (have) code = (file_start && k == 0u) ? 0xffffu : myhits[k - shift];
                if (k + 1u < n_mine) next = myhits[k + 1u - shift];
                uint32_t s = 0;
                bool to_detail;
                if (!HAS_QUAL) {
                    to_detail = k1_quick_step<false>(a, sm, cc, ps, have, 0u, code, s, n_parsed);
                } else {
                    if (have && code != 0xffffu) s = (code >> 5) * 16u + (code & 7u) * 4u + ((code >> 3) & 3u) + 1u;
                    to_detail = have;
                }
                uint32_t len_hint = 0;                        // up to the '\n' in front of this lane's next start
                if (to_detail && next != 0xffffu) len_hint = (next >> 5) * 16u + (next & 7u) * 4u + ((next >> 3) & 3u) - s;
                n_dq = k1_push(sm.dq, n_dq, lane, to_detail, k1_entry(0u, len_hint, base + s));
                if (n_dq >= 32u) {                            // leaves both queues below 32
                    drained = true;
                    const uint32_t r = k1_drain_detail<HAS_QUAL, ALL>(a, sm, cs, lane, n_dq, n_gq, false);
                    n_dq = r & 0xffffu; n_gq = r >> 16;
                }
            }
            continue;
        }
        // ---- order: one prefix sum over the lanes ---------------------------------------------------
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const uint32_t n_tile_lines = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t first_idx = incl - cnt;
        // per-line results go to the tile's row of the staging array; its line count is all the ordering pass needs
        if (a.line_out && lane == 0) {
            a.tile_lines[tile] = n_tile_lines;
            atomicAdd(&a.group_lines[tile / K1_ORDER_TILES], (unsigned long long)n_tile_lines);   // (for k1_tile_prefix_kernel)
        }
        PROF(6);
        // ---- the warp's list, file order (the first K1_WCAP starts; more -> later passes scan again) -----
        {
            const uint32_t skip = file_start ? 1u : 0u;
            if (file_start) sm.starts[0] = 0xffffu;
            for (uint32_t k = 0; k + skip < cnt; k++) {
                const uint32_t idx = first_idx + skip + k;
                if (idx < (uint32_t)K1_WCAP) sm.starts[idx] = myhits[k];
            }
        }
        PROF(7);
        // ---- parse, K1_WCAP lines per pass -------------------------------------------------------------
        for (uint32_t done = 0; done < n_tile_lines; done += K1_WCAP) {
            __syncwarp();
            if (done) {
                k1_relist(sm, lane, wlen, first_idx + (file_start ? 1u : 0u) - done);   // (wraps below the pass: rejected)
                __syncwarp();
            }
            const uint32_t n_pass = n_tile_lines - done < (uint32_t)K1_WCAP ? n_tile_lines - done : (uint32_t)K1_WCAP;
            // (The lines are parsed in file order.  Sorting a pass by line length, so that the 32 lines of a step run
            //  their loops equally long, was measured: the counting sort costs more than the divergence it removes.)
            // ---- the pass's lines, 32 per step, in file order ---------------------------------------------
            K1Pass ps;
            ps.base = base; ps.tile = (unsigned long long)tile; ps.wlen = wlen; ps.done = done; ps.eof = eof;
            PROF(8);
            if (done + n_pass == n_tile_lines) {              // the tile's last pass: the next ticket, its latency hidden
                if (lane == 0) ticket = (int)atom_inc_u32(&a.st->next_tile);   // behind the parse
                ticket_taken = true;
            }
            for (uint32_t l0 = 0; l0 < n_pass; l0 += 32u) {
                const bool have = l0 + (uint32_t)lane < n_pass;
                const uint32_t l = have ? l0 + (uint32_t)lane : 0u;
                const uint32_t code = have ? sm.starts[l] : 0u;
                uint32_t s = 0;
                bool to_detail;
                PROF(9);
                if (!HAS_QUAL) {
                    to_detail = k1_quick_step<ALL>(a, sm, cc, ps, have, l, code, s, n_parsed);
                } else {
                    if (have && code != 0xffffu) s = (code >> 5) * 16u + (code & 7u) * 4u + ((code >> 3) & 3u) + 1u;
                    to_detail = have;
                }
                PROF(10);
                // a line for the second tier is queued with its length (up to the '\n' in front of the next listed
                // start); the queue is run when 32 wait
                uint32_t len_hint = 0;
                if (to_detail && l + 1u < n_pass) {
                    const uint32_t c1 = sm.starts[l + 1u];
                    len_hint = (c1 >> 5) * 16u + (c1 & 7u) * 4u + ((c1 >> 3) & 3u) - s;
                }
                n_dq = k1_push(sm.dq, n_dq, lane, to_detail, k1_entry(done + l, len_hint, base + s));
                if (n_dq >= 32u) {                            // leaves both queues below 32
                    drained = true;
                    const uint32_t r = k1_drain_detail<HAS_QUAL, ALL>(a, sm, cs, lane, n_dq, n_gq, false);
                    n_dq = r & 0xffffu; n_gq = r >> 16;
                }
                PROF(11);
            }
        }
        PROF(12);
        n_lines += n_tile_lines;
    }
    {
        PROF(13);
        const uint32_t r = k1_drain_detail<HAS_QUAL, ALL>(a, sm, cs, lane, n_dq, n_gq, true);
        PROF(14);
        k1_drain_general(a, sm, cs, lane, r >> 16, true);
    }
    PROF(15);
    PROF_FLUSH(a.st);
    // ---- statistics -------------------------------------------------------------------------------
    uint32_t np = n_parsed + cs.n_parsed, ng = cs.n_general;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        np += __shfl_xor_sync(0xffffffffu, np, d);
        ng += __shfl_xor_sync(0xffffffffu, ng, d);
    }
    if (lane == 0) {
        atomicAdd(&a.st->n_parsed, (unsigned long long)np);
        atomicAdd(&a.st->n_general, (unsigned long long)ng);
        atomicAdd(&a.st->n_lines, n_lines);
    }
}

// ---- per-line results into file order (all-positions mode with line_out) ------------------------------------------
// k1_tile_prefix_kernel: block b owns the b-th group of K1_ORDER_TILES consecutive tiles.  It sums the line totals of the
// groups in front of its own (the pileup kernel keeps them: one atomic add per tile) and scans its own tiles' counts
// -> tile_first[t] = lines the tiles in front of t own.

__global__ void __launch_bounds__(K1_ORDER_TILES) k1_tile_prefix_kernel(const uint32_t *tile_lines,
                                                                        const unsigned long long *group_lines, int n_tiles,
                                                                        unsigned long long *tile_first) {
    __shared__ unsigned long long wsum[K1_ORDER_TILES / 32];
    __shared__ unsigned long long carry_s;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = (int)blockIdx.x * K1_ORDER_TILES;
    unsigned long long sum = 0;                               // lines of the groups of K1_ORDER_TILES tiles in front of this one
    for (int g = tid; g < (int)blockIdx.x; g += K1_ORDER_TILES) sum += group_lines[g];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) wsum[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        unsigned long long c = 0;
        for (int w = 0; w < K1_ORDER_TILES / 32; w++) c += wsum[w];
        carry_s = c;
    }
    __syncthreads();
    const int t = t0 + tid;
    const uint32_t mine = t < n_tiles ? tile_lines[t] : 0u;
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) wsum[warp] = incl;                        // (carry_s was read from wsum before the barrier above)
    __syncthreads();
    unsigned long long wbase = 0;
    for (int w = 0; w < warp; w++) wbase += wsum[w];
    if (t < n_tiles) tile_first[t] = carry_s + wbase + incl - mine;
}

// k1_lines_kernel: one warp per tile copies the tile's row of the staging array to its place in line_out, all its
// loads in flight together; the blocks behind the last tile scatter the overflow list's entries.
__global__ void k1_lines_kernel(const uint16_t *stage, const uint32_t *tile_lines, const unsigned long long *tile_first,
                                int n_tiles, int tile_blocks, const unsigned long long *over, unsigned long long over_cap,
                                const PileupStatusDev *st, uint16_t *line_out, unsigned long long line_out_cap) {
    if ((int)blockIdx.x < tile_blocks) {
        const int tile = (int)blockIdx.x * (int)(blockDim.x >> 5) + (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
        if (tile >= n_tiles) return;
        const unsigned long long first = tile_first[tile];
        uint32_t n = tile_lines[tile];
        if (n > (uint32_t)K1_STAGE_CAP) n = (uint32_t)K1_STAGE_CAP;
        const uint16_t *row = stage + (unsigned long long)tile * K1_STAGE_CAP;
        uint16_t v[K1_STAGE_CAP / 32];
#pragma unroll
        for (int k = 0; k < K1_STAGE_CAP / 32; k++) v[k] = (uint32_t)(lane + 32 * k) < n ? row[lane + 32 * k] : (uint16_t)0;
#pragma unroll
        for (int k = 0; k < K1_STAGE_CAP / 32; k++) {
            const unsigned long long slot = first + (unsigned long long)(lane + 32 * k);
            if ((uint32_t)(lane + 32 * k) < n && slot < line_out_cap) line_out[slot] = v[k];
        }
    } else {
        unsigned long long n = st->over_used;
        if (n > over_cap) n = over_cap;                       // (more than fit: the finish kernel reports it)
        const unsigned long long k = (unsigned long long)((int)blockIdx.x - tile_blocks) * blockDim.x + threadIdx.x;
        if (k >= n) return;
        const unsigned long long e = over[k];
        const unsigned long long slot = tile_first[e >> 32] + ((e >> 16) & 0xffffull);
        if (slot < line_out_cap) line_out[slot] = (uint16_t)(e & 0xffffull);
    }
}

int k1_launch_order(cudaStream_t stream, const PileupArgs &a) {
    if (!a.line_out || a.n_tiles <= 0) return 0;
    k1_tile_prefix_kernel<<<(a.n_tiles + K1_ORDER_TILES - 1) / K1_ORDER_TILES, K1_ORDER_TILES, 0, stream>>>(
        a.tile_lines, a.group_lines, a.n_tiles, a.tile_first);
    const int tile_blocks = (a.n_tiles + 7) / 8, over_blocks = (int)((a.over_cap + 255) / 256);
    k1_lines_kernel<<<tile_blocks + over_blocks, 256, 0, stream>>>(a.stage, a.tile_lines, a.tile_first, a.n_tiles, tile_blocks,
                                                                   a.over, a.over_cap, a.st, a.line_out, a.line_out_cap);
    return 2;
}

// ---- K3: gather the site cells into the consensus row, snplist order (call_consensus.py:187-188); the first
//      thread also turns the device-side status into the caller's snpgpu_pileup_stats ----------------------
__global__ void k1_finish_kernel(const unsigned long long *site_cells, const int32_t *snp_unique, size_t n_snp,
                                 uint8_t *row_out, const PileupStatusDev *st, unsigned long long over_cap,
                                 snpgpu_pileup_stats *out) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_snp) {
        unsigned long long c = site_cells[snp_unique[k]];
        row_out[k] = c ? (uint8_t)(c & 0xffu) : (uint8_t)'-';
    }
#ifdef K1_PROF
    if (k == 0) {
        const unsigned long long *pc = reinterpret_cast<const unsigned long long *>(st) + 8;
        unsigned long long tot = 0;
        for (int i = 0; i < 16; i++) tot += pc[i];
        printf("K1_PROF total %llu Mcycles:", tot / 1000000ull);
        for (int i = 0; i < 16; i++) printf(" p%d %.1f%%", i, 100.0 * (double)pc[i] / (double)tot);
        printf("\n");
    }
#endif
    if (k == 0 && out) {
        out->n_lines = st->n_lines;
        out->n_parsed = st->n_parsed;
        out->n_general = st->n_general;
        out->reserved = 0;
        if (st->arena_overflow || st->over_used > over_cap) {
            out->error_offset = st->arena_overflow ? st->arena_used : 0ull;      // bytes of splice scratch the call needs
            out->reserved = st->over_used > over_cap ? (int32_t)((st->over_used + 1023ull) >> 10) : 0;   // overflow entries / 1024
            out->error_code = SNPGPU_E_NOMEM;
        } else if (st->first_error_inv != 0ull) {
            const unsigned long long e = ~st->first_error_inv;
            out->error_offset = e >> 8;
            out->error_code = (int32_t)(e & 0xffull);
        } else {
            out->error_offset = ~0ull;
            out->error_code = 0;
        }
    }
}

// universal newlines (pileup.py:417): a CR that is not followed by LF ends a line -> make it an LF, in place
__global__ void k1_normalize_newlines_kernel(uint8_t *text, unsigned long long nbytes) {
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (; i < nbytes; i += stride)
        if (text[i] == '\r' && (i + 1 >= nbytes || text[i + 1] != '\n')) text[i] = '\n';
}

int k1_launch_normalize(cudaStream_t stream, uint8_t *text, size_t nbytes) {
    if (!nbytes) return 0;
    k1_normalize_newlines_kernel<<<148 * 8, 256, 0, stream>>>(text, nbytes);
    return 1;
}

template <bool Q, bool A>
static int k1_occupancy() {
    int n = 0;
    cudaFuncSetAttribute(k1_pileup_kernel<Q, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k1_pileup_kernel<Q, A>, K1_THREADS, k1_smem_bytes());
    return n;
}

// resident CTAs per SM (the smallest over the variants, which also sets their shared-memory attribute)
int k1_blocks_per_sm(bool has_qual) {
    int n0 = has_qual ? k1_occupancy<true, false>() : k1_occupancy<false, false>();
    int n1 = has_qual ? k1_occupancy<true, true>() : k1_occupancy<false, true>();
    return n0 < n1 ? n0 : n1;
}

int k1_launch(cudaStream_t stream, const PileupArgs &a, int grid_blocks) {
    if (a.n_tiles <= 0) return 0;
    const int want = (a.n_tiles + K1_WARPS - 1) / K1_WARPS;
    const int grid = want < grid_blocks ? want : grid_blocks;
    const bool q = a.p.min_base_qual > 0, all = a.mode == SNPGPU_MODE_ALL;
    const size_t sh = k1_smem_bytes();
    if (q && all) k1_pileup_kernel<true, true><<<grid, K1_THREADS, sh, stream>>>(a);
    else if (q) k1_pileup_kernel<true, false><<<grid, K1_THREADS, sh, stream>>>(a);
    else if (all) k1_pileup_kernel<false, true><<<grid, K1_THREADS, sh, stream>>>(a);
    else k1_pileup_kernel<false, false><<<grid, K1_THREADS, sh, stream>>>(a);
    return 1;
}

int k1_launch_finish(cudaStream_t stream, const unsigned long long *site_cells, const int32_t *snp_unique, size_t n_snp,
                     uint8_t *row_out_dev, const PileupStatusDev *st, unsigned long long over_cap,
                     snpgpu_pileup_stats *stats_dev) {
    if (!n_snp && !stats_dev) return 0;
    const unsigned grid = (unsigned)((n_snp + 255) / 256);
    k1_finish_kernel<<<grid ? grid : 1u, 256, 0, stream>>>(site_cells, snp_unique, n_snp, row_out_dev, st, over_cap, stats_dev);
    return 1;
}

}  // namespace snpgpu
