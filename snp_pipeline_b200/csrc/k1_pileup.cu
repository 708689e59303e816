// k1_pileup.cu -- K1: pileup text in HBM -> consensus cells, for a batch of samples in one launch.
//
// Replaces the per-sample hot loop of the reference (call_consensus.py:161-188): pileup.Reader.__iter__
// (pileup.py:408-429), pileup.Record (pileup.py:209-325) and ConsensusCaller.call_consensus (pileup.py:492-590), for all
// the samples of a batch (the reference runs one call_consensus process per sample, run.py:709-710).
//
// Shape:
//   * every WARP is its own pipeline, no block-wide barrier.  A warp takes tiles by ticket; the ticket counter spans the
//     batch, so the tail of a launch (last tile, last lines) is paid once per batch, not once per sample;
//   * a tile (32 x K1_LANE_BYTES of text) plus look-ahead is staged into the warp's slice of shared memory by ONE 1-D bulk
//     async copy (TMA engine, mbarrier completion); the other resident warps compute while this one waits;
//   * NO scan pass: lane L owns the lines whose preceding '\n' lies in ITS K1_LANE_BYTES of the tile.  It looks for the
//     first '\n' of its range (a few words) and then walks its lines one after the other -- a line's end is where the
//     parse of its quality column ends (line_quick3.cuh), so nothing has to find, order or list line starts beforehand.
//     The 32 lanes run their k-th lines in lock step; a tile is sized so that a lane owns just under 4 lines;
//   * the first tier (line_quick3.cuh) reads the columns as aligned words; a line it declines costs the lane one search
//     for its '\n' and goes, by file offset, to a global queue that k1_rest_kernel works through densely afterwards
//     (a second look by the first tier with indel tokens skipped, then line_fast.cuh -> line_general.cuh, one thread per
//     line) -- the hot kernel carries none of that code;
//   * results: atomicMax per hit site keeps the LAST line of a position in file order (the dict overwrite of
//     call_consensus.py:169-176); in all-positions mode the uint16 of a lane's k-th line goes to slot [tile][k][lane] of
//     the staging array (one coalesced 64-byte store per warp-step), and the slots are put into file order by two small
//     kernels afterwards (lane counts -> prefix -> copy).
// Algorithmic traffic: every text byte read once (+ the look-ahead re-read, an L2 hit), 2 B written per line.
#include "k1_tiers.cuh"
#include "line_quick3.cuh"

namespace snpgpu {

// one warp's slice of shared memory: the window, its '\n' sentinels, the cached contig's name rows, the copy's mbarrier
constexpr uint32_t K1_SLICE_ROWS  = K1_WIN + K1_PAD;                     // byte offset of the name rows
constexpr uint32_t K1_SLICE_BAR   = K1_SLICE_ROWS + 4u * Q3_ROWS_WORDS;   // ... of the mbarrier
constexpr uint32_t K1_SLICE_BYTES = (K1_SLICE_BAR + 8u + 15u) & ~15u;
constexpr uint32_t K1_TAB_BYTES   = 2u * Q3_TABN * 2u;                    // the CTA's filter tables, in front of the slices
static_assert(K1_PAD >= (int)QUICK_PAD, "line_quick3.cuh precondition");
static_assert(K1_TAB_BYTES + K1_SLICE_BYTES * K1_WARPS <= 227 * 1024 && K1_CTAS_PER_SM == 1, "shared memory per CTA (one CTA per SM)");
size_t k1_smem_bytes() { return (size_t)K1_TAB_BYTES + (size_t)K1_SLICE_BYTES * K1_WARPS; }

extern __shared__ __align__(128) uint8_t k1_smem_raw[];

struct SmemWin {                                // line_quick3.cuh's memory policy: words of the warp's slice
    uint32_t base_w;
    __device__ __forceinline__ uint32_t ld(uint32_t k) const { return reinterpret_cast<const uint32_t *>(k1_smem_raw)[base_w + k]; }
    __device__ __forceinline__ void ld4(uint32_t k, uint32_t *w) const {       // k a multiple of 4: one 16-byte load
        const uint4 v = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(k1_smem_raw) + base_w + k);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    __device__ __forceinline__ uint32_t row(uint32_t k) const { return ld(k); }        // (the name rows live in the slice too)
    __device__ __forceinline__ void row4(uint32_t k, uint32_t *w) const { ld4(k, w); }
    __device__ __forceinline__ uint32_t tab16(uint32_t k) const { return reinterpret_cast<const uint16_t *>(k1_smem_raw)[k]; }
    __device__ __forceinline__ uint32_t byte(uint32_t off) const { return k1_smem_raw[4u * base_w + off]; }
};

// queue entry, second word: sample | flags << 32 | (contig of the line + 1, when its key columns were taken) << 40
enum : unsigned long long { K1_Q_GENERAL = 1ull << 32 };     // odd bytes seen: straight to line_general.cuh
enum : unsigned long long { K1_Q_TALLY = 1ull << 33 };       // the first tier took the whole line, the call needs tallies: second tier
constexpr unsigned long long K1_Q_EMPTY = ~0ull;             // first word of a slot its warp claimed and did not fill
constexpr int K1_QBLOCK = 64;                                // queue entries a warp claims at a time

// The contig of the line at offset s of the window (byte-wise; once per tile at most): loads the warp's cache with it.
// A name the site table does not hold is cached all the same (cid -1, no sites: its lines are parsed / skipped like any
// other); a line without a tab leaves the cache alone.  All 32 lanes call this together.
__device__ __noinline__ void k1_follow_contig(const SiteTable &t, const SmemWin m, uint32_t s, uint32_t limit, Q3Contig *cc) {
    uint32_t n = 0;
    bool clean = true;                                        // printable ASCII only: str.split() sees the same token
    while (s + n < limit && n < 64u && m.byte(s + n) != '\t' && m.byte(s + n) != '\n') {
        clean &= (m.byte(s + n) - 0x21u) < 0x5eu;
        n++;
    }
    if (s + n >= limit || n >= 64u || n == 0u || !clean || m.byte(s + n) != '\t') return;
    const int cid = contig_find(t, k1_smem_raw + 4u * m.base_w + s, (int64_t)n);
    const uint32_t L = n + 1u;
    const int lane = threadIdx.x & 31;
    auto name_at = [&](uint32_t idx) { return m.byte(s + idx); };
    uint32_t w[4];
#pragma unroll
    for (int r = 0; r < 4; r++) w[r] = q3_rows_word(name_at, L, (uint32_t)lane + 32u * r);     // 128 words, four per lane
    __syncwarp();
    uint32_t *rows = reinterpret_cast<uint32_t *>(k1_smem_raw + 4u * m.base_w + K1_SLICE_ROWS);
#pragma unroll
    for (int r = 0; r < 4; r++) rows[(uint32_t)lane + 32u * r] = w[r];
    q3_contig_set(cc, K1_SLICE_ROWS / 4u, L, cid, cid >= 0 ? t.max_pos[cid] : -1, cid >= 0 ? t.bit_base[cid] : 0);
    __syncwarp();
}

template <bool ALL>
__global__ void __launch_bounds__(K1_THREADS, K1_CTAS_PER_SM) k1_pileup_kernel(const __grid_constant__ K1Batch g) {
    const int lane = threadIdx.x & 31;
    const int warp = (int)(threadIdx.x >> 5);
    const SmemWin m{K1_TAB_BYTES / 4u + (uint32_t)warp * (K1_SLICE_BYTES / 4u)};
    uint8_t *buf = k1_smem_raw + K1_TAB_BYTES + (uint32_t)warp * K1_SLICE_BYTES;
    for (uint32_t k = threadIdx.x; k < 2u * Q3_TABN; k += K1_THREADS)       // the filter thresholds (line_quick3.cuh)
        reinterpret_cast<uint16_t *>(k1_smem_raw)[k] = (uint16_t)q3_tab_entry(k, g.p);
    __syncthreads();                                          // (the kernel's only block-wide barrier)
    uint64_t *bar = reinterpret_cast<uint64_t *>(buf + K1_SLICE_BAR);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncwarp();
    constexpr uint32_t R = K1_LANE_BYTES, TILE = K1_TILE, WIN = K1_WIN;
    const int n_gwarps = gridDim.x * K1_WARPS;
    const uint32_t one = g.one;
    uint32_t parity = 0;
    Q3Contig cc;
    q3_contig_set(&cc, K1_SLICE_ROWS / 4u, 0u, -1, -1, 0);       // nothing cached: the first tile's first line fills it
    static_assert(K1_PAD == 32, "one sentinel byte per lane");
    buf[WIN + (uint32_t)lane] = (uint8_t)'\n';              // the '\n' sentinels behind a whole window: no copy ever reaches them
    __syncwarp();
    int ticket = 0;
    if (lane == 0) ticket = (int)atom_inc_u32(g.next_tile);
    int si = 0;                                               // sample of the previous tile (tickets rise: so do samples)
    uint32_t acc_lines = 0, acc_ok = 0;                       // lines / parsed lines of sample si not yet added to its status
    // the follow-up kernel's queue is claimed K1_QBLOCK entries at a time: one atomic per block, not per warp-step (with a
    // large snplist nearly every step of the sites mode queues a line); slots a warp leaves unused are marked empty
    unsigned long long q_base = 0;
    uint32_t q_used = K1_QBLOCK;
    for (;;) {
        const int t = __shfl_sync(0xffffffffu, ticket, 0);
        if (t >= g.total_tiles) break;
        // ---- which sample?  (tile0 ascending; a warp's tickets only rise) -------------------------------------
        if (si + 1 < g.n_samples && g.s[si + 1].tile0 <= t) {
            if (lane == 0 && acc_lines) {
                atomicAdd(&g.s[si].st->n_lines, (unsigned long long)acc_lines);
                if (acc_ok) atomicAdd(&g.s[si].st->n_parsed, (unsigned long long)acc_ok);
            }
            acc_lines = acc_ok = 0;
            do si++; while (si + 1 < g.n_samples && g.s[si + 1].tile0 <= t);
        }
        const K1Samp &S = g.s[si];
        const int tile = t - S.tile0;
        const uint8_t *text = S.text;
        const unsigned long long nbytes = S.nbytes;
        const unsigned long long base = (unsigned long long)tile * TILE;
        const unsigned long long left = nbytes - base;
        const uint32_t wlen = left < (unsigned long long)WIN ? (uint32_t)left : WIN;
        const uint32_t bulk = wlen & ~15u;
        const bool eof = left <= (unsigned long long)WIN;
        // ---- stage the window ---------------------------------------------------------------------------
        __syncwarp();                                         // every lane is done with the previous window
        if (lane == 0 && bulk) {
            fence_proxy_async();
            mbar_arrive_expect_tx(bar, bulk);
            bulk_g2s(buf, text + base, bulk, bar);
#ifndef K1_CFG_PF_ROUNDS
#define K1_CFG_PF_ROUNDS 1
#endif
            const unsigned long long nbase = base + (unsigned long long)(K1_CFG_PF_ROUNDS * n_gwarps) * TILE;   // about one round from now -> L2
            if (nbase < nbytes) {
                const unsigned long long nleft = nbytes - nbase;
                const uint32_t nb = (nleft < (unsigned long long)WIN ? (uint32_t)nleft : WIN) & ~15u;
                if (nb) bulk_prefetch_l2(text + nbase, nb);
            }
        }
        if (wlen != WIN) {                                    // a sample's last windows: the bytes behind the last whole 16
            for (uint32_t j = bulk + (uint32_t)lane; j < wlen + (uint32_t)K1_PAD; j += 32u)     // and the sentinels behind them
                buf[j] = j < wlen ? text[base + j] : (uint8_t)'\n';
        }
        if (lane == 0) ticket = (int)atom_inc_u32(g.next_tile);   // the next tile, its latency hidden behind this one
        // per-sample pointers the lane loop needs (uniform loads, while the copy is in flight)
        unsigned long long *site_cells = S.site_cells;
        uint16_t *stage_lane = (ALL && S.stage) ? S.stage + ((unsigned long long)tile * (K1_LCAP * 32) + (unsigned)lane) : nullptr;
        if (bulk) { mbar_wait(bar, parity); parity ^= 1u; }
        __syncwarp();
        // ---- the lane's first line: the byte behind the first '\n' of its range, looked for 16 bytes at a time -----
        const uint32_t r0 = (uint32_t)lane * R;
        const uint32_t r1 = r0 + R < wlen ? r0 + R : wlen;    // '\n' at [r0, r1) start this lane's lines
        uint32_t s = 0;
        bool have;
        if (tile == 0 && lane == 0) {
            have = left > 0ull;                               // the first line of the file
        } else {
            static_assert(K1_LANE_BYTES % 16 == 0, "the search reads whole 16-byte chunks of the lane's range");
            uint32_t p = r1;
            uint4 vn = *reinterpret_cast<const uint4 *>(buf + r0);
            for (uint32_t c0 = r0; c0 < r1; c0 += 16u) {
                const uint4 v = vn;
                vn = *reinterpret_cast<const uint4 *>(buf + c0 + 16u);          // (the next chunk, a trip ahead; the last one reads into the next lane's range or the look-ahead)
                const uint32_t x0 = v.x ^ 0x0a0a0a0au, x1 = v.y ^ 0x0a0a0a0au, x2 = v.z ^ 0x0a0a0a0au, x3 = v.w ^ 0x0a0a0a0au;
                const uint32_t z0 = (x0 - 0x01010101u) & ~x0 & 0x80808080u, z1 = (x1 - 0x01010101u) & ~x1 & 0x80808080u;
                const uint32_t z2 = (x2 - 0x01010101u) & ~x2 & 0x80808080u, z3 = (x3 - 0x01010101u) & ~x3 & 0x80808080u;
                if (z0 | z1 | z2 | z3) {                      // (the lowest flag of a word is exact: its first zero byte)
                    const uint32_t z = z0 ? z0 : (z1 ? z1 : (z2 ? z2 : z3));
                    p = c0 + (z0 ? 0u : (z1 ? 4u : (z2 ? 8u : 12u))) + ((uint32_t)ctz32(z) >> 3);
                    break;
                }
            }
            s = p + 1u;
            have = p < r1 && (unsigned long long)s < left;    // (a '\n' that ends the text starts no line)
        }
        {   // the tile's first line tells which contig the warp expects
            const uint32_t hb = __ballot_sync(0xffffffffu, have);
            if (hb) {
                const int src = __ffs((int)hb) - 1;
                const uint32_t s0 = __shfl_sync(0xffffffffu, s, src);
                if (!q3_name(m, s0, wlen, cc)) k1_follow_contig(g.sites, m, s0, wlen, &cc);
            }
        }
        uint32_t cnt = 0, n_ok = 0;
        while (__any_sync(0xffffffffu, have)) {
            bool push = false;
            uint32_t push_len = 0, push_flag = 0, push_cid = 0, nxt = 0;
            if (have) {
                uint32_t next;
                Q3Line q;
                const bool keyok = q3_key(m, s, wlen, cc, one, &q);
                const bool known = keyok && (int32_t)q.pos <= cc.max_pos;
                const uint32_t widx = cc.word_base + (q.pos >> 5), bb = q.pos & 31u;
                if (ALL) {
                    SiteWord sw{0u, 0u, 0u, 0u};
                    if (known) sw = load_site_word(g.sites.words + widx);
                    int st = ST_DETAIL;
                    if (keyok && !g.all_rest) st = q3_rest(m, q.after, wlen, g.p, one, &q);
                    if (st == ST_OK && !(q.end == wlen && !eof)) {    // (a line that leaves the window goes on)
                        unsigned fail = q.fail;
                        if ((sw.exc >> bb) & 1u) fail |= FAIL_REGION;
                        const unsigned cell = fail ? (unsigned)'-' : q.base;      // (the base is a letter: never '*')
                        if ((sw.snp >> bb) & 1u) {
                            const uint32_t site = sw.rank + (uint32_t)__popc(sw.any & ((1u << bb) - 1u));
                            atomicMax(&site_cells[site], ((base + s + 1ull) << 8) | (unsigned long long)cell);
                        }
                        if (stage_lane) {
                            const uint16_t v = (uint16_t)(cell | (fail << 8));
                            if (cnt < (uint32_t)K1_LCAP) {
                                stage_lane[cnt * 32u] = v;
                            } else {
                                const unsigned long long n = atomicAdd(&S.st->over_used, 1ull);
                                if (n < g.over_cap)
                                    S.over[n] = ((((unsigned long long)tile << 5) | (unsigned)lane) << 32) | ((unsigned long long)(cnt & 0xffffu) << 16) | (unsigned long long)v;
                            }
                        }
                        n_ok++;
                        next = q.end + 1u;
                    } else if (st == ST_TALLY && q.end < wlen) {      // well-formed, the reference base does not win: its end is known
                        push = true;
                        push_len = q.end - s;
                        push_flag = 2u;
                        push_cid = (uint32_t)(cc.cid + 1);
                        next = q.end + 1u;
                    } else {
                        uint32_t odd = 0;
                        const uint32_t e = q3_find_nl(m, st == ST_SIGN || st == ST_ZERO ? q.end : s, one, &odd);   // (columns 1-5 / 1-4 are clean)
                        push = true;
                        push_len = e < wlen || eof ? e - s : 0u;
                        push_flag = odd ? 1u : 0u;
                        push_cid = keyok ? (uint32_t)(cc.cid + 1) : 0u;
                        next = e + 1u;
                    }
                } else {
                    // filter mode (pileup.py:423-427): only the key columns matter, then the line's end
                    uint32_t odd = 0;
                    const uint32_t e = q3_find_nl(m, keyok ? q.after - 1u : s, one, &odd);
                    const bool at_site = known && ((g.sites.bits[widx] >> bb) & 1u);
                    if (!keyok || odd || at_site || (e >= wlen && !eof)) {
                        push = true;
                        push_cid = keyok ? (uint32_t)(cc.cid + 1) : 0u;
                        push_len = e < wlen || eof ? e - s : 0u;
                        push_flag = (odd || (e >= wlen && !eof)) ? 1u : 0u;
                    }
                    next = e + 1u;
                }
                nxt = next;
            }
            const uint32_t pb = __ballot_sync(0xffffffffu, push);
            if (pb) {                                         // declined lines -> the follow-up kernel's queue
                const uint32_t np = (uint32_t)__popc(pb);
                if (q_used + np > (uint32_t)K1_QBLOCK) {      // the block is full: mark its tail empty, claim the next one
                    for (uint32_t k = q_used + (uint32_t)lane; k < (uint32_t)K1_QBLOCK; k += 32u)
                        if (q_base + k < g.queue_cap) g.queue[2ull * (q_base + k)] = K1_Q_EMPTY;
                    if (lane == 0) q_base = atomicAdd(g.queue_count, (unsigned long long)K1_QBLOCK);
                    q_base = __shfl_sync(0xffffffffu, q_base, 0);
                    q_used = 0;
                }
                const unsigned long long slot = q_base + q_used + (unsigned long long)__popc(pb & ((1u << lane) - 1u));
                if (push && slot < g.queue_cap) {
                    g.queue[2ull * slot] = k1_entry(cnt, push_len, base + s);
                    g.queue[2ull * slot + 1ull] = (unsigned long long)si | ((unsigned long long)push_flag << 32) | ((unsigned long long)push_cid << 40);
                }
                q_used += np;
            }
            if (have) {
                cnt++;
                s = nxt;
                have = nxt - 1u < r1 && (unsigned long long)nxt < left;      // the '\n' in front of the next line is mine
            }
        }
        // ---- the tile's books ------------------------------------------------------------------------------------
        const uint32_t total = __reduce_add_sync(0xffffffffu, cnt);
        acc_lines += total;
        acc_ok += __reduce_add_sync(0xffffffffu, n_ok);
        if (ALL && S.stage) {
            S.lane_lines[((unsigned long long)tile << 5) | (unsigned)lane] = (uint8_t)(cnt < 255u ? cnt : 255u);
            if (lane == 0) {
                S.tile_lines[tile] = total;
                atomicAdd(&S.group_lines[tile / K1_ORDER_TILES], (unsigned long long)total);   // (for k1_tile_prefix_kernel)
            }
        }
    }
    if (lane == 0 && acc_lines) {
        atomicAdd(&g.s[si].st->n_lines, (unsigned long long)acc_lines);
        if (acc_ok) atomicAdd(&g.s[si].st->n_parsed, (unsigned long long)acc_ok);
    }
    for (uint32_t k = q_used + (uint32_t)lane; k < (uint32_t)K1_QBLOCK; k += 32u)      // the unused tail of the warp's last block
        if (q_base + k < g.queue_cap) g.queue[2ull * (q_base + k)] = K1_Q_EMPTY;
}

// ---- the follow-up kernel: the queued lines, one thread per line, through the second and third tier ---------------
__device__ __forceinline__ void k1_sample_args(const K1Batch &g, uint32_t si, PileupArgs *a) {
    const K1Samp &S = g.s[si];
    a->text = S.text; a->nbytes = S.nbytes; a->sites = g.sites; a->p = g.p; a->mode = g.mode;
    a->site_cells = S.site_cells; a->stage = g.mode == SNPGPU_MODE_ALL ? S.stage : nullptr; a->over = S.over; a->over_cap = g.over_cap;
    a->rec_off = S.rec_off; a->rec_count = S.rec_count; a->rec_cap = S.rec_cap; a->st = S.st;
    a->arena = g.arena; a->arena_cap = g.arena_cap; a->arena_st = g.s[0].st;
}

// line_quick3.cuh's memory policy over the text where it lies in global memory (the follow-up kernel): words relative to
// a 16-byte aligned base in front of the line, name rows from the site table, filter tables in shared memory
struct GmemWin {
    const uint32_t *p;
    const uint32_t *rows;
    const uint16_t *tab;
    __device__ __forceinline__ uint32_t ld(uint32_t k) const { return __ldg(p + k); }
    __device__ __forceinline__ void ld4(uint32_t k, uint32_t *w) const {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p + k));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    __device__ __forceinline__ uint32_t byte(uint32_t off) const { return __ldg(reinterpret_cast<const uint8_t *>(p) + off); }
    __device__ __forceinline__ uint32_t row(uint32_t k) const { return __ldg(rows + k); }
    __device__ __forceinline__ void row4(uint32_t k, uint32_t *w) const {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(rows + k));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    __device__ __forceinline__ uint32_t tab16(uint32_t k) const { return tab[k]; }
};

// The follow-up kernel's copy of a line: every thread brings its line into its own K1_RSTAGE bytes of shared memory with
// independent 16-byte loads (one exposed latency instead of one per word the parsers look at), from the 16-byte boundary in
// front of the line to 48 bytes behind it (what the first tier's word loads may touch).  Lines that do not fit, lines of
// unknown length and the last lines of a text are parsed where they lie.
constexpr uint32_t K1_RSTAGE = 208;
struct StagedWin {
    const uint32_t *p;
    const uint32_t *rows;
    const uint16_t *tab;
    __device__ __forceinline__ uint32_t ld(uint32_t k) const { return p[k]; }
    __device__ __forceinline__ void ld4(uint32_t k, uint32_t *w) const {
        const uint4 v = *reinterpret_cast<const uint4 *>(p + k);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    __device__ __forceinline__ uint32_t byte(uint32_t off) const { return reinterpret_cast<const uint8_t *>(p)[off]; }
    __device__ __forceinline__ uint32_t row(uint32_t k) const { return __ldg(rows + k); }
    __device__ __forceinline__ void row4(uint32_t k, uint32_t *w) const {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(rows + k));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    }
    __device__ __forceinline__ uint32_t tab16(uint32_t k) const { return tab[k]; }
};
__device__ __forceinline__ bool k1_rest_stage(const K1Samp &S, unsigned long long goff, uint32_t len, uint8_t *mine) {
    const unsigned long long abase = goff & ~15ull;
    const uint32_t need = (uint32_t)(goff - abase) + len + 48u;
    const uint32_t chunks = (need + 15u) >> 4;
    if (len == 0u || need > K1_RSTAGE || abase + 16ull * chunks > S.nbytes) return false;
    const uint4 *src = reinterpret_cast<const uint4 *>(S.text + abase);
    uint4 *dst = reinterpret_cast<uint4 *>(mine);
    constexpr uint32_t NC = K1_RSTAGE / 16u, HALF = (NC + 1u) / 2u;
    uint4 v[HALF];                                            // (two rounds: half the registers of one)
#pragma unroll
    for (uint32_t r = 0; r < 2u; r++) {
#pragma unroll
        for (uint32_t c = 0; c < HALF; c++) if (r * HALF + c < chunks) v[c] = __ldg(src + r * HALF + c);
#pragma unroll
        for (uint32_t c = 0; c < HALF; c++) if (r * HALF + c < chunks) dst[r * HALF + c] = v[c];
    }
    return true;
}

// A queued line whose key columns the pileup kernel took (default mode: a line at a site; all-positions mode: a line the
// first look declined) through the first tier again, densely -- 32 such lines per warp -- and with indel tokens skipped.
// true: decided (cell / per-line result stored); false: on to the second tier.
template <class WIN>
__device__ __forceinline__ bool k1_rest_quick(const K1Batch &g, const K1Samp &S, const uint16_t *tab, uint32_t cid,
                                              unsigned long long goff, uint32_t len, uint32_t line_idx, uint32_t one,
                                              const uint8_t *staged) {
    if (len == 0u || goff + (unsigned long long)len + 48ull > S.nbytes) return false;     // (the words read reach past the line)
    const unsigned long long abase = goff & ~15ull;
    const WIN m{reinterpret_cast<const uint32_t *>(staged ? staged : S.text + abase), g.sites.q3rows + (size_t)cid * SITE_Q3ROWS_WORDS, tab};
    const uint32_t s = (uint32_t)(goff - abase), limit = s + len;      // (the line's own '\n' stands where the sentinels would)
    Q3Contig cc;
    q3_contig_set(&cc, 0u, (uint32_t)g.sites.len1[cid], (int32_t)cid, g.sites.max_pos[cid], g.sites.bit_base[cid]);
    Q3Line q;
    if (!q3_key(m, s, limit, cc, one, &q)) return false;
    const bool all = g.mode == SNPGPU_MODE_ALL;
    SiteWord sw{0u, 0u, 0u, 0u};
    const uint32_t bb = q.pos & 31u;
    if ((int32_t)q.pos <= cc.max_pos) sw = load_site_word(g.sites.words + cc.word_base + (q.pos >> 5));
    if (!all && !((sw.any >> bb) & 1u)) return false;
    const int st = q3_rest<true>(m, q.after, limit, g.p, one, &q);
    if (st == ST_ZERO) {                                      // depth 0 (pileup.py:226-234): ('-', RawDpth) whatever follows -- the pileup
        q.base = (uint8_t)'-';                                // kernel saw no CR / VT / FF / byte >= 0x80 in the line (else K1_Q_GENERAL)
        q.fail = FAIL_RAWDPTH;
    } else if (st != ST_OK || q.end != limit) return false;
    unsigned fail = q.fail;
    if ((sw.exc >> bb) & 1u) fail |= FAIL_REGION;
    const unsigned cell = fail ? (unsigned)'-' : q.base;
    if ((sw.snp >> bb) & 1u) {
        const uint32_t site = sw.rank + (uint32_t)__popc(sw.any & ((1u << bb) - 1u));
        atomicMax(&S.site_cells[site], ((goff + 1ull) << 8) | (unsigned long long)cell);
    }
    if (S.rec_off) {                                          // the VCF pass wants to know which lines were parsed
        const unsigned long long k = atomicAdd(S.rec_count, 1ull);
        if (k < S.rec_cap) S.rec_off[k] = goff;
    }
    if (all && S.stage) {                                     // slot [tile][k][lane] of the lane that owns the '\n' in front of the line
        const unsigned long long tl = (goff ? goff - 1ull : 0ull) / (unsigned long long)K1_LANE_BYTES;
        const uint16_t v = (uint16_t)(cell | (fail << 8));
        if (line_idx < (uint32_t)K1_LCAP) {
            S.stage[((tl >> 5) * (unsigned long long)K1_LCAP + line_idx) * 32ull + (tl & 31ull)] = v;
        } else {
            const unsigned long long n = atomicAdd(&S.st->over_used, 1ull);
            if (n < g.over_cap) S.over[n] = (tl << 32) | ((unsigned long long)(line_idx & 0xffffu) << 16) | (unsigned long long)v;
        }
    }
    return true;
}

template <bool HAS_QUAL>
__global__ void __launch_bounds__(128) k1_rest_kernel(const __grid_constant__ K1Batch g) {
    __shared__ uint16_t tab[2 * Q3_TABN];
    __shared__ __align__(16) uint8_t stage_s[128 * K1_RSTAGE];
    uint8_t *mine = stage_s + threadIdx.x * K1_RSTAGE;
    for (uint32_t k = threadIdx.x; k < 2u * Q3_TABN; k += blockDim.x) tab[k] = (uint16_t)q3_tab_entry(k, g.p);
    __syncthreads();
    unsigned long long n = *g.queue_count;
    if (n > g.queue_cap) n = g.queue_cap;                     // (more than fit: the finish kernel reports it)
    K1Cold cs{0, 0u, 0u};
    PileupArgs a;
    uint32_t cur = 0xffffffffu;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long e0 = g.queue[2ull * i];
        if (e0 == K1_Q_EMPTY) continue;
        const unsigned long long e1 = g.queue[2ull * i + 1ull];
        const uint32_t cid1 = (uint32_t)(e1 >> 40);
        const bool staged = !(e1 & K1_Q_GENERAL) && k1_rest_stage(g.s[(uint32_t)e1], k1_entry_goff(e0), k1_entry_len(e0), mine);
        bool quick = false;
        if (!HAS_QUAL && cid1 != 0u && !(e1 & (K1_Q_GENERAL | K1_Q_TALLY))) {
            if (staged) quick = k1_rest_quick<StagedWin>(g, g.s[(uint32_t)e1], tab, cid1 - 1u, k1_entry_goff(e0), k1_entry_len(e0), k1_entry_idx(e0), g.one, mine);
            else quick = k1_rest_quick<GmemWin>(g, g.s[(uint32_t)e1], tab, cid1 - 1u, k1_entry_goff(e0), k1_entry_len(e0), k1_entry_idx(e0), g.one, nullptr);
        }
        {   // one add per warp and sample, not one per line
            const uint32_t act = __activemask();
            const uint32_t same = __match_any_sync(act, quick ? (uint32_t)e1 : 0xffffffffu);
            if (quick && (int)(threadIdx.x & 31u) == __ffs((int)same) - 1)
                atomicAdd(&g.s[(uint32_t)e1].st->n_parsed, (unsigned long long)__popc(same));
        }
        if (quick) continue;
        if ((uint32_t)e1 != cur) { cur = (uint32_t)e1; k1_sample_args(g, cur, &a); }
        cs.n_parsed = cs.n_general = 0u;
        bool more = true;
        if (!(e1 & K1_Q_GENERAL)) {
            const uint8_t *st_buf = staged ? mine : nullptr;
            if (g.mode == SNPGPU_MODE_ALL) more = k1_detail<HAS_QUAL, true>(a, cs, k1_entry_goff(e0), k1_entry_idx(e0), k1_entry_len(e0), st_buf);
            else more = k1_detail<HAS_QUAL, false>(a, cs, k1_entry_goff(e0), k1_entry_idx(e0), k1_entry_len(e0), st_buf);
        }
        if (more) k1_general(a, cs, k1_entry_goff(e0), k1_entry_idx(e0));
        if (cs.n_parsed) atomicAdd(&a.st->n_parsed, (unsigned long long)cs.n_parsed);
        if (cs.n_general) atomicAdd(&a.st->n_general, (unsigned long long)cs.n_general);
    }
}

// ---- per-line results into file order (all-positions mode with line_out), batched over the samples (blockIdx.y) ------
// k1_tile_prefix_kernel: block b owns the b-th group of K1_ORDER_TILES consecutive tiles of its sample.  It sums the line
// totals of the groups in front of its own (the pileup kernel keeps them: one atomic add per tile) and scans its own
// tiles' counts -> tile_first[t] = lines the tiles in front of t own.
__global__ void __launch_bounds__(K1_ORDER_TILES) k1_tile_prefix_kernel(const __grid_constant__ K1Batch g) {
    const K1Samp &a = g.s[blockIdx.y];
    if (!a.stage || (int)blockIdx.x * K1_ORDER_TILES >= a.n_tiles) return;
    __shared__ unsigned long long wsum[K1_ORDER_TILES / 32];
    __shared__ unsigned long long carry_s;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = (int)blockIdx.x * K1_ORDER_TILES;
    unsigned long long sum = 0;                               // lines of the groups in front of this one
    for (int gi = tid; gi < (int)blockIdx.x; gi += K1_ORDER_TILES) sum += a.group_lines[gi];
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
    const uint32_t mine = t < a.n_tiles ? a.tile_lines[t] : 0u;
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    const unsigned long long carry = carry_s;
    __syncthreads();
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    unsigned long long wbase = 0;
    for (int w = 0; w < warp; w++) wbase += wsum[w];
    if (t < a.n_tiles) a.tile_first[t] = carry + wbase + incl - mine;
}

// k1_lines_kernel: one warp per tile -- prefix over the lanes' line counts, every lane copies its slots of the staging
// array to its place in line_out; the blocks behind the last tile scatter the overflow list's entries.
__global__ void k1_lines_kernel(const __grid_constant__ K1Batch g, int tile_blocks) {
    const K1Samp &a = g.s[blockIdx.y];
    if (!a.stage) return;
    if ((int)blockIdx.x < tile_blocks) {
        const int tile = (int)blockIdx.x * (int)(blockDim.x >> 5) + (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
        if (tile >= a.n_tiles) return;
        const unsigned long long tl = ((unsigned long long)tile << 5) | (unsigned long long)lane;
        const uint32_t cnt = a.lane_lines[tl];
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        const uint16_t *slots = a.stage + (unsigned long long)tile * K1_LCAP * 32ull + (unsigned long long)lane;
        uint16_t v[K1_LCAP];
#pragma unroll
        for (int k = 0; k < K1_LCAP; k++) v[k] = (uint32_t)k < cnt ? slots[k * 32] : (uint16_t)0;
        const unsigned long long first = a.tile_first[tile] + (unsigned long long)(incl - cnt);
#pragma unroll
        for (int k = 0; k < K1_LCAP; k++) {
            const unsigned long long slot = first + (unsigned long long)k;
            if ((uint32_t)k < cnt && slot < a.line_out_cap) a.line_out[slot] = v[k];
        }
    } else {
        unsigned long long n = a.st->over_used;
        if (n > g.over_cap) n = g.over_cap;                   // (more than fit: the finish kernel reports it)
        const unsigned long long k = (unsigned long long)((int)blockIdx.x - tile_blocks) * blockDim.x + threadIdx.x;
        if (k >= n) return;
        const unsigned long long e = a.over[k];
        const unsigned long long tl = e >> 32;
        const uint32_t lane = (uint32_t)tl & 31u;
        unsigned long long slot = a.tile_first[tl >> 5] + ((e >> 16) & 0xffffull);
        for (uint32_t l = 0; l < lane; l++) slot += a.lane_lines[(tl & ~31ull) + l];
        if (slot < a.line_out_cap) a.line_out[slot] = (uint16_t)(e & 0xffffull);
    }
}

// ---- K3: gather the site cells into the consensus rows, snplist order (call_consensus.py:187-188); the first thread of
//      a sample also turns the device-side status into the caller's snpgpu_pileup_stats -----------------------------
__global__ void k1_finish_kernel(const __grid_constant__ K1Batch g) {
    const K1Samp &S = g.s[blockIdx.y];
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < g.n_snp) {
        const unsigned long long c = S.site_cells[g.snp_unique[k]];
        S.row_out[k] = c ? (uint8_t)(c & 0xffu) : (uint8_t)'-';
    }
    if (!S.stats_out) return;
    // "called consensus positions" (call_consensus.py:184): the snplist positions at least one parsed line fell on
    unsigned called = 0;
    for (size_t u = k; u < S.n_unique; u += (size_t)gridDim.x * blockDim.x)
        called += ((g.sites.flags[u] & SITE_SNP) && S.site_cells[u] != 0ull) ? 1u : 0u;
    called = __reduce_add_sync(0xffffffffu, called);
    if ((threadIdx.x & 31u) == 0u && called) atomicAdd(&S.st->n_called, (unsigned long long)called);
    __syncthreads();
    if (threadIdx.x != 0) return;
    __threadfence();
    if (atomicAdd(&S.st->finish_done, 1u) != gridDim.x - 1u) return;      // (the sample's last block writes the stats)
    __threadfence();
    const volatile PileupStatusDev *st = S.st, *ast = g.s[0].st;
    snpgpu_pileup_stats *out = S.stats_out;
    out->n_lines = st->n_lines;
    out->n_parsed = st->n_parsed;
    out->n_general = st->n_general;
    out->n_called = st->n_called;
    out->reserved = 0;
    const unsigned long long queued = *g.queue_count;
    if (queued > g.queue_cap) {                               // the follow-up queue was too small: entries the batch needs
        out->error_offset = queued;
        out->reserved = -1;
        out->error_code = SNPGPU_E_NOMEM;
    } else if (ast->arena_overflow || st->over_used > g.over_cap) {
        out->error_offset = ast->arena_overflow ? ast->arena_used : 0ull;        // bytes of splice scratch the batch needs
        out->reserved = st->over_used > g.over_cap ? (int32_t)((st->over_used + 1023ull) >> 10) : 0;   // overflow entries / 1024
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

// ---- launchers ---------------------------------------------------------------------------------------------------
int k1_blocks_per_sm() {
    int n0 = 0, n1 = 0;
    cudaFuncSetAttribute(k1_pileup_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes());
    cudaFuncSetAttribute(k1_pileup_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes());
    // the follow-up kernel runs between two pileup kernels: it keeps their shared-memory carve-out (a different split of
    // L1 / shared memory between consecutive kernels costs a reconfiguration each time, and measured as noise)
    cudaFuncSetAttribute(k1_rest_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(k1_rest_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n0, k1_pileup_kernel<false>, K1_THREADS, k1_smem_bytes());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n1, k1_pileup_kernel<true>, K1_THREADS, k1_smem_bytes());
    return n0 < n1 ? n0 : n1;
}

// the pileup kernel and its follow-up kernel over the batch (the two launches bench.py's roofline times together)
int k1_launch(cudaStream_t stream, const K1Batch &g, int grid_blocks, int n_sms) {
    if (g.total_tiles <= 0) return 0;
    const int want = (g.total_tiles + K1_WARPS - 1) / K1_WARPS;
    const int grid = want < grid_blocks ? want : grid_blocks;
    const size_t sh = k1_smem_bytes();
    if (g.mode == SNPGPU_MODE_ALL) k1_pileup_kernel<true><<<grid, K1_THREADS, sh, stream>>>(g);
    else k1_pileup_kernel<false><<<grid, K1_THREADS, sh, stream>>>(g);
#ifndef K1_REST_BLOCKS
#define K1_REST_BLOCKS 8
#endif
    if (g.has_qual) k1_rest_kernel<true><<<n_sms * K1_REST_BLOCKS, 128, 0, stream>>>(g);
    else k1_rest_kernel<false><<<n_sms * K1_REST_BLOCKS, 128, 0, stream>>>(g);
    return 2;
}

// ordering kernels (when any sample wants per-line results), finish kernel; max_tiles: the largest over the batch's samples
int k1_launch_finish(cudaStream_t stream, const K1Batch &g, int max_tiles, bool want_lines) {
    int launches = 0;
    if (want_lines && max_tiles > 0) {
        const dim3 pg((unsigned)((max_tiles + K1_ORDER_TILES - 1) / K1_ORDER_TILES), (unsigned)g.n_samples);
        k1_tile_prefix_kernel<<<pg, K1_ORDER_TILES, 0, stream>>>(g);
        const int tile_blocks = (max_tiles + 7) / 8, over_blocks = (int)((g.over_cap + 255) / 256);
        const dim3 lg((unsigned)(tile_blocks + over_blocks), (unsigned)g.n_samples);
        k1_lines_kernel<<<lg, 256, 0, stream>>>(g, tile_blocks);
        launches += 2;
    }
    const unsigned fx = (unsigned)((g.n_snp + 255) / 256);
    k1_finish_kernel<<<dim3(fx ? fx : 1u, (unsigned)g.n_samples), 256, 0, stream>>>(g);
    return launches + 1;
}

}  // namespace snpgpu
