// k1_pileup.cu -- K1: pileup text in HBM -> consensus cells.
//
// Replaces the per-sample hot loop of the reference (call_consensus.py:161-188): pileup.Reader.__iter__
// (pileup.py:408-429), pileup.Record (pileup.py:209-325) and ConsensusCaller.call_consensus (pileup.py:492-590).
//
// Shape of the kernel:
//   * persistent CTAs stride over 32 KiB tiles of the text; a tile owns the lines whose preceding '\n' lies
//     inside it (the first line of the file belongs to tile 0);
//   * each tile (+2 KiB of look-ahead so the last owned line is complete) is staged into shared memory by ONE
//     1-D bulk async copy (TMA engine, mbarrier completion) -- no register staging, fully coalesced;
//   * scan: every thread tests 16-byte chunks for '\n' with SWAR arithmetic, a ballot turns the hits of a warp
//     into ordered line-start lists (and a tile-wide "byte >= 0x80 present" flag);
//   * parse: one thread per line runs the fast parser (line_fast.cuh) out of shared memory, four bases per
//     32-bit load; lines it declines are queued and run through the exact any-input parser
//     (line_general.cuh) on the text in global memory afterwards, so the common path stays convergent;
//   * results: an atomicMax per hit site keeps the LAST line of a position in file order (the dict overwrite
//     of call_consensus.py:169-176); in all-positions mode a uint16 per line is staged per tile and compacted
//     to file order by a second tiny kernel.
// Algorithmic traffic: every text byte read once (+6 % look-ahead re-read, an L2 hit), 2 B written per line.
#include "internal.h"
#include "line_fast.cuh"
#include "line_general.cuh"

namespace snpgpu {

struct K1Smem {
    alignas(16) uint8_t buf[K1_TILE + K1_LOOK + K1_PAD];
    uint16_t wstarts[K1_WARPS][K1_WCAP];   // line starts (buffer offsets) per warp region, file order
    uint32_t genq[K1_GENQ];                // fallback lines: (tile line index << 16) | buffer offset
    uint32_t wcount[K1_WARPS];             // line starts per warp region (all of them, recorded or not)
    uint32_t full_prefix[K1_WARPS + 1];    // exclusive prefix of wcount: tile-level index of a region's first line
    uint32_t pass_count[K1_WARPS];         // starts recorded in the current pass
    uint32_t pass_prefix[K1_WARPS + 1];
    uint32_t max_wcount;
    uint32_t n_genq;
    uint32_t tile_high;                    // some byte >= 0x80 in the window, or a CR that is not followed by LF
    unsigned long long n_parsed, n_general, n_lines;
    alignas(8) uint64_t bar;
};

size_t k1_smem_bytes() { return sizeof(K1Smem); }

// exact per-byte mask (0x80 where the byte equals '\n'), any byte values
__device__ __forceinline__ uint32_t nl_mask(uint32_t w) {
    uint32_t t = w ^ 0x0a0a0a0au;
    return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t) & 0x80808080u;
}

// the same for words whose bytes are all < 0x80 (no carries between byte lanes): one operation less
__device__ __forceinline__ uint32_t nl_mask7(uint32_t w) {
    return ~((w ^ 0x0a0a0a0au) + 0x7f7f7f7fu) & 0x80808080u;
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

// first '\n' at or after buf[i], or limit
__device__ __forceinline__ uint32_t find_nl(const uint8_t *buf, uint32_t i, uint32_t limit) {
    while (i < limit) {
        uint32_t t = load_u32(buf, i) ^ 0x0a0a0a0au;
        uint32_t z = (t - 0x01010101u) & ~t & 0x80808080u;
        if (z) {
            uint32_t p = i + ((uint32_t)ctz32(z) >> 3);
            return p < limit ? p : limit;
        }
        i += 4u;
    }
    return limit;
}

struct K1Thread {
    const PileupArgs &a;
    K1Smem &sm;
    int tile;
    unsigned long long base;
    uint32_t n_parsed, n_general;

    __device__ __forceinline__ void report(unsigned long long goff, int code) {
        atomicMin(&a.st->first_error, (goff << 8) | (unsigned long long)code);
    }

    // call_consensus.py:165-176: Region failure, '-' substitution, keep the cell for the snplist gather
    __device__ __forceinline__ void emit(unsigned base_ch, unsigned fail, int32_t site, unsigned long long goff,
                                         uint32_t line_idx) {
        unsigned flags = site >= 0 ? a.sites.flags[site] : 0u;
        if (flags & SITE_EXCLUDED) fail |= FAIL_REGION;
        unsigned cell = (fail || base_ch == '*') ? (unsigned)'-' : base_ch;
        if (flags & SITE_SNP) atomicMax(&a.site_cells[site], ((goff + 1ull) << 8) | (unsigned long long)cell);
        if (a.line_stage && line_idx < (uint32_t)K1_MAXLINES)
            a.line_stage[(size_t)tile * K1_MAXLINES + line_idx] = (uint16_t)(cell | (fail << 8));
        n_parsed++;
    }

    // the exact path, on the text where it lies in global memory
    __device__ __noinline__ void general(unsigned long long goff, uint32_t line_idx) {
        const uint8_t *line = a.text + goff;
        unsigned long long room = a.nbytes - goff;
        int64_t n = 0;
        bool lone_cr = false;
        while ((unsigned long long)n < room && line[n] != '\n') {
            if (line[n] == '\r' && (unsigned long long)(n + 1) < room && line[n + 1] != '\n') lone_cr = true;
            n++;
        }
        n_general++;
        if (lone_cr) { report(goff, ST_LONECR); return; }   // classic-Mac line end: the caller normalises and reruns
        LineCall r;
        const bool all = a.mode == SNPGPU_MODE_ALL;
        int32_t site = -1;
        if (!all) {
            general_key(line, n, &r);                         // pileup.py:423-427
            if (r.status) { report(goff, r.status); return; }
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
        if (r.status) { report(goff, r.status); return; }
        if (all) {
            int cid = contig_find(a.sites, line + r.chrom_off, r.chrom_len);
            site = site_find(a.sites, cid, r.pos);
        }
        emit(r.base, r.fail, site, goff, line_idx);
    }
};

// One pass of the newline scan over this warp's 4 KiB region.  Records the starts whose index within the
// region falls in [pass * K1_WCAP, (pass + 1) * K1_WCAP); returns the region's total number of starts.
__device__ __forceinline__ uint32_t k1_scan_region(K1Smem &sm, int warp, int lane, int tile, uint32_t wlen,
                                                   uint32_t pass, uint32_t &hi_acc, uint32_t &cr_acc) {
    uint32_t wtotal = 0;
    const uint32_t lo_idx = pass * K1_WCAP;
    if (tile == 0 && warp == 0 && wlen > 0) {                 // the first line of the file
        if (lane == 0 && pass == 0) sm.wstarts[0][0] = 0;
        wtotal = 1;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll 2
    for (int it = 0; it < K1_WREGION / 512; it++) {
        const uint32_t off = (uint32_t)warp * K1_WREGION + (uint32_t)it * 512u + (uint32_t)lane * 16u;
        uint32_t m = 0;                                       // bit 8*b + w  <->  byte 4*w + b of the chunk
        if (off < wlen) {
            const uint4 v = *reinterpret_cast<const uint4 *>(sm.buf + off);
            hi_acc |= v.x | v.y | v.z | v.w;
            cr_acc |= cr_any(v.x) | cr_any(v.y) | cr_any(v.z) | cr_any(v.w);
            if ((v.x | v.y | v.z | v.w) & 0x80808080u)         // exact form when a byte >= 0x80 is around
                m = (nl_mask(v.x) >> 7) | (nl_mask(v.y) >> 6) | (nl_mask(v.z) >> 5) | (nl_mask(v.w) >> 4);
            else
                m = (nl_mask7(v.x) >> 7) | (nl_mask7(v.y) >> 6) | (nl_mask7(v.z) >> 5) | (nl_mask7(v.w) >> 4);
            if (off + 17u > wlen) {                           // last chunk of the text: a start must be < wlen
                for (uint32_t pos = 0; pos < 16u; pos++)
                    if (off + pos + 1u >= wlen) m &= ~(1u << (8u * (pos & 3u) + (pos >> 2)));
            }
        }
        const uint32_t cnt = (uint32_t)__popc(m);
        const uint32_t b1 = __ballot_sync(0xffffffffu, cnt > 0u);
        const uint32_t b2 = __ballot_sync(0xffffffffu, cnt > 1u);
        uint32_t excl, total;
        if (b2 == 0u) {
            excl = (uint32_t)__popc(b1 & lt_mask);
            total = (uint32_t)__popc(b1);
        } else {
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            excl = incl - cnt;
            total = __shfl_sync(0xffffffffu, incl, 31);
        }
        uint32_t idx = wtotal + excl - lo_idx;                // wraps when below this pass: rejected by < K1_WCAP
        if (cnt == 1u) {
            const uint32_t f = (uint32_t)ctz32(m);
            if (idx < (uint32_t)K1_WCAP) sm.wstarts[warp][idx] = (uint16_t)(off + (f & 7u) * 4u + (f >> 3) + 1u);
        } else if (cnt > 1u) {
            for (uint32_t pos = 0; pos < 16u; pos++) {
                if ((m >> (8u * (pos & 3u) + (pos >> 2))) & 1u) {
                    if (idx < (uint32_t)K1_WCAP) sm.wstarts[warp][idx] = (uint16_t)(off + pos + 1u);
                    idx++;
                }
            }
        }
        wtotal += total;
    }
    return wtotal;
}

template <bool HAS_QUAL>
__global__ void __launch_bounds__(K1_THREADS, 4) k1_pileup_kernel(const PileupArgs a) {
    extern __shared__ __align__(128) uint8_t k1_smem_raw[];
    K1Smem &sm = *reinterpret_cast<K1Smem *>(k1_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&sm.bar, 1);
        mbar_fence_init();
        sm.n_parsed = 0; sm.n_general = 0; sm.n_lines = 0;
    }
    __syncthreads();
    uint32_t parity = 0;
    int hint = 0;
    K1Thread th{a, sm, 0, 0ull, 0u, 0u};
    const bool all = a.mode == SNPGPU_MODE_ALL;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const unsigned long long base = (unsigned long long)tile * K1_TILE;
        const unsigned long long left = a.nbytes - base;
        const uint32_t wlen = left < (unsigned long long)(K1_TILE + K1_LOOK) ? (uint32_t)left : (uint32_t)(K1_TILE + K1_LOOK);
        const uint32_t bulk = wlen & ~15u;
        th.tile = tile; th.base = base;
        // ---- stage the window ---------------------------------------------------------------------
        if (tid == 0) {
            sm.n_genq = 0; sm.tile_high = 0;
            if (bulk) {
                fence_proxy_async();
                mbar_arrive_expect_tx(&sm.bar, bulk);
                bulk_g2s(sm.buf, a.text + base, bulk, &sm.bar);
            }
        }
        for (uint32_t j = bulk + (uint32_t)tid; j < wlen + (uint32_t)K1_PAD; j += K1_THREADS)
            sm.buf[j] = j < wlen ? a.text[base + j] : (uint8_t)'\n';
        if (bulk) { mbar_wait(&sm.bar, parity); parity ^= 1u; }
        __syncthreads();
        // ---- scan: line starts + high-bit flag ----------------------------------------------------
        uint32_t hi_acc = 0, cr_acc = 0;
        uint32_t wtotal = k1_scan_region(sm, warp, lane, tile, wlen, 0u, hi_acc, cr_acc);
        if (tid < K1_LOOK / 16) {                             // look-ahead bytes: only the odd-byte tests
            const uint32_t off = (uint32_t)K1_TILE + (uint32_t)tid * 16u;
            if (off < wlen) {
                const uint4 v = *reinterpret_cast<const uint4 *>(sm.buf + off);
                hi_acc |= v.x | v.y | v.z | v.w;
                if ((cr_any(v.x) | cr_any(v.y) | cr_any(v.z) | cr_any(v.w)) & 0x80808080u)
                    if (lone_cr_in_chunk(sm.buf, off)) hi_acc |= 0x80u;
            }
        }
        if (cr_acc & 0x80808080u) {                           // CRs present: fine when each is followed by LF
            for (int it = 0; it < K1_WREGION / 512; it++) {
                const uint32_t off = (uint32_t)warp * K1_WREGION + (uint32_t)it * 512u + (uint32_t)lane * 16u;
                if (off < wlen && lone_cr_in_chunk(sm.buf, off)) hi_acc |= 0x80u;
            }
        }
        if (lane == 0) sm.wcount[warp] = wtotal;
        if (hi_acc & 0x80808080u) sm.tile_high = 1u;          // every line of the tile takes the exact path
        __syncthreads();
        // ---- parse ----------------------------------------------------------------------------------
        if (tid == 0) {
            uint32_t run = 0, mx = 0;
            for (int w = 0; w < K1_WARPS; w++) {
                sm.full_prefix[w] = run;
                run += sm.wcount[w];
                mx = sm.wcount[w] > mx ? sm.wcount[w] : mx;
            }
            sm.full_prefix[K1_WARPS] = run;
            sm.max_wcount = mx;
        }
        __syncthreads();
        const uint32_t n_tile_lines = sm.full_prefix[K1_WARPS];
        const bool high = sm.tile_high != 0u;
        const uint32_t n_pass = (sm.max_wcount + K1_WCAP - 1u) / K1_WCAP;
        for (uint32_t pass = 0; pass < n_pass; pass++) {
            if (pass > 0) {
                uint32_t dummy = 0, dummy2 = 0;
                k1_scan_region(sm, warp, lane, tile, wlen, pass, dummy, dummy2);
            }
            if (tid == 0) {
                uint32_t run = 0;
                for (int w = 0; w < K1_WARPS; w++) {
                    const uint32_t done = pass * K1_WCAP;
                    const uint32_t rem = sm.wcount[w] > done ? sm.wcount[w] - done : 0u;
                    const uint32_t c = rem < (uint32_t)K1_WCAP ? rem : (uint32_t)K1_WCAP;
                    sm.pass_count[w] = c;
                    sm.pass_prefix[w] = run;
                    run += c;
                }
                sm.pass_prefix[K1_WARPS] = run;
            }
            __syncthreads();
            const uint32_t n_pass_lines = sm.pass_prefix[K1_WARPS];
            for (uint32_t l = (uint32_t)tid; l < n_pass_lines; l += K1_THREADS) {
                int w = 0;
#pragma unroll
                for (int x = 1; x < K1_WARPS; x++) w += (l >= sm.pass_prefix[x]) ? 1 : 0;
                const uint32_t k = l - sm.pass_prefix[w];
                const uint32_t line_idx = sm.full_prefix[w] + pass * K1_WCAP + k;
                const uint32_t s = sm.wstarts[w][k];
                uint32_t e;
                if (k + 1u < sm.pass_count[w]) e = (uint32_t)sm.wstarts[w][k + 1] - 1u;
                else if (pass == 0u && k + 1u == sm.wcount[w] && w + 1 < K1_WARPS && sm.wcount[w + 1] > 0u)
                    e = (uint32_t)sm.wstarts[w + 1][0] - 1u;            // the next region's first line follows
                else e = find_nl(sm.buf, s, wlen);
                const unsigned long long goff = base + s;
                if (high || (e == wlen && base + wlen != a.nbytes)) {   // odd bytes, or the line leaves the window
                    th.general(goff, line_idx);
                    continue;
                }
                FastLine fl;
                int st = fast_line<HAS_QUAL>(sm.buf, s, e, a.sites, hint, a.p, all, &fl);
                if (st == ST_OK) {
                    th.emit(fl.base, fl.fail, fl.site, goff, line_idx);
                } else if (st == ST_FALLBACK) {
                    uint32_t q = atomicAdd(&sm.n_genq, 1u);
                    if (q < (uint32_t)K1_GENQ) sm.genq[q] = (line_idx << 16) | s;
                    else th.general(goff, line_idx);
                }
            }
            __syncthreads();
            const uint32_t nq = sm.n_genq < (uint32_t)K1_GENQ ? sm.n_genq : (uint32_t)K1_GENQ;
            for (uint32_t g = (uint32_t)tid; g < nq; g += K1_THREADS) {
                const uint32_t v = sm.genq[g];
                th.general(base + (v & 0xffffu), v >> 16);
            }
            __syncthreads();
            if (tid == 0) sm.n_genq = 0;
        }
        if (tid == 0) {
            if (a.tile_nlines) a.tile_nlines[tile] = n_tile_lines;
            sm.n_lines += n_tile_lines;
        }
        __syncthreads();                                      // everyone is done with buf before it is refilled
    }
    // ---- statistics -------------------------------------------------------------------------------
    uint32_t np = th.n_parsed, ng = th.n_general;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        np += __shfl_xor_sync(0xffffffffu, np, d);
        ng += __shfl_xor_sync(0xffffffffu, ng, d);
    }
    if (lane == 0) {
        atomicAdd(&sm.n_parsed, (unsigned long long)np);
        atomicAdd(&sm.n_general, (unsigned long long)ng);
    }
    __syncthreads();
    if (tid == 0) {
        atomicAdd(&a.st->n_parsed, sm.n_parsed);
        atomicAdd(&a.st->n_general, sm.n_general);
        atomicAdd(&a.st->n_lines, sm.n_lines);
    }
}

// ---- K3: gather the site cells into the consensus row, snplist order (call_consensus.py:187-188) ---------
__global__ void k1_row_kernel(const unsigned long long *site_cells, const int32_t *snp_unique, size_t n_snp,
                              uint8_t *row_out) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_snp) return;
    unsigned long long c = site_cells[snp_unique[k]];
    row_out[k] = c ? (uint8_t)(c & 0xffu) : (uint8_t)'-';
}

// ---- per-line results: staged per tile -> file order -----------------------------------------------------
__global__ void k1_tile_prefix_kernel(const uint32_t *tile_nlines, int n_tiles, unsigned long long *tile_prefix) {
    __shared__ unsigned long long part[1024];
    const int tid = threadIdx.x;
    const int per = (n_tiles + 1023) / 1024;
    const int lo = tid * per, hi = min(lo + per, n_tiles);
    unsigned long long s = 0;
    for (int i = lo; i < hi; i++) s += tile_nlines[i];
    part[tid] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        unsigned long long v = tid >= d ? part[tid - d] : 0ull;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    unsigned long long run = part[tid] - s;
    for (int i = lo; i < hi; i++) { tile_prefix[i] = run; run += tile_nlines[i]; }
    if (tid == 1023) tile_prefix[n_tiles] = part[1023];
}

__global__ void k1_lines_kernel(const uint16_t *line_stage, const uint32_t *tile_nlines,
                                const unsigned long long *tile_prefix, int n_tiles, uint16_t *line_out, size_t cap) {
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint32_t n = tile_nlines[tile];
        if (n > (uint32_t)K1_MAXLINES) n = K1_MAXLINES;       // more lines than that: one of them raised
        const unsigned long long o = tile_prefix[tile];
        const uint16_t *src = line_stage + (size_t)tile * K1_MAXLINES;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            if (o + i < cap) line_out[o + i] = src[i];
    }
}

__global__ void k1_stats_kernel(const PileupStatusDev *st, snpgpu_pileup_stats *out) {
    out->n_lines = st->n_lines;
    out->n_parsed = st->n_parsed;
    out->n_general = st->n_general;
    if (st->arena_overflow) {
        out->error_offset = st->arena_used;                   // bytes of scratch the call needs
        out->error_code = SNPGPU_E_NOMEM;
    } else if (st->first_error != ~0ull) {
        out->error_offset = st->first_error >> 8;
        out->error_code = (int32_t)(st->first_error & 0xffull);
    } else {
        out->error_offset = ~0ull;
        out->error_code = 0;
    }
    out->reserved = 0;
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

int k1_blocks_per_sm(bool has_qual) {
    int n = 0;
    if (has_qual) {
        cudaFuncSetAttribute(k1_pileup_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Smem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k1_pileup_kernel<true>, K1_THREADS, sizeof(K1Smem));
    } else {
        cudaFuncSetAttribute(k1_pileup_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K1Smem));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k1_pileup_kernel<false>, K1_THREADS, sizeof(K1Smem));
    }
    return n;
}

int k1_launch(cudaStream_t stream, const PileupArgs &a, int grid_blocks) {
    if (a.n_tiles <= 0) return 0;
    int grid = a.n_tiles < grid_blocks ? a.n_tiles : grid_blocks;
    if (a.p.min_base_qual > 0) k1_pileup_kernel<true><<<grid, K1_THREADS, sizeof(K1Smem), stream>>>(a);
    else k1_pileup_kernel<false><<<grid, K1_THREADS, sizeof(K1Smem), stream>>>(a);
    return 1;
}

int k1_launch_row(cudaStream_t stream, const unsigned long long *site_cells, const int32_t *snp_unique, size_t n_snp,
                  uint8_t *row_out_dev) {
    if (!n_snp) return 0;
    k1_row_kernel<<<(unsigned)((n_snp + 255) / 256), 256, 0, stream>>>(site_cells, snp_unique, n_snp, row_out_dev);
    return 1;
}

int k1_launch_lines(cudaStream_t stream, const uint16_t *line_stage, const uint32_t *tile_nlines, int n_tiles,
                    unsigned long long *tile_prefix, uint16_t *line_out_dev, size_t line_out_cap) {
    if (n_tiles <= 0) return 0;
    k1_tile_prefix_kernel<<<1, 1024, 0, stream>>>(tile_nlines, n_tiles, tile_prefix);
    int grid = n_tiles < 148 * 8 ? n_tiles : 148 * 8;
    k1_lines_kernel<<<grid, 128, 0, stream>>>(line_stage, tile_nlines, tile_prefix, n_tiles, line_out_dev, line_out_cap);
    return 2;
}

int k1_launch_stats(cudaStream_t stream, const PileupStatusDev *st, snpgpu_pileup_stats *stats_dev) {
    k1_stats_kernel<<<1, 1, 0, stream>>>(st, stats_dev);
    return 1;
}

}  // namespace snpgpu
