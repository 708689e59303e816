// k5_vcf.cu -- K5: the per-line tallies behind the per-sample consensus VCF.
//
// Replaces what vcf_writer.SingleSampleWriter._make_vcf_record_from_pileup (vcf_writer.py:295-379) reads off a
// pileup.Record: raw depth, the reference base's total / forward / reverse good depth, and the ALT alleles -- every
// other surviving symbol, in most_common_good_bases order (pileup.py:260-266) -- with theirs.  The reference writes
// one VCF record per pileup line its Reader yields (call_consensus.py:161-184): the lines K1 parsed.  K1 (run with
// PileupArgs::rec_off set) lists their file offsets; here they are put in file order and each line is tallied by
// the exact any-input parser (line_general.cuh), one thread per line, on the text in global memory.  A few tens of
// thousands of lines per sample in the default mode; every line with --vcfAllPos.
// The sort is K2's radix sort (k2_merge.cu).  k5_text_*: the records as the text of the VCF's data lines, formatted on the device
// (two passes: sizes + prefix, then the bytes), so that no per-record host loop is left.
#include "internal.h"
#include "line_general.cuh"

namespace snpgpu {

__global__ void k5_tally_kernel(const uint8_t *text, unsigned long long nbytes, SiteTable sites, CallParams p,
                                const unsigned long long *offsets, size_t n_rec, snpgpu_vcf_record *rec_out,
                                snpgpu_vcf_alt *alt_out, unsigned long long alt_cap, unsigned long long *alt_count,
                                PileupStatusDev *st, uint8_t *arena, unsigned long long arena_cap) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rec) return;
    const unsigned long long goff = offsets[i];
    const uint8_t *line = text + goff;
    const unsigned long long room = nbytes - goff;
    int64_t n = 0;
    while ((unsigned long long)n < room && line[n] != '\n') n++;
    uint32_t tot[128], fwd[128], rev[128];
    LineTallyOut r;
    general_tally(line, n, p, nullptr, 0, tot, fwd, rev, &r);
    if (r.status == ST_NEED_ARENA) {
        unsigned long long want = ((unsigned long long)r.bases_len + 15ull) & ~15ull;
        unsigned long long off = atomicAdd(&st->arena_used, want);
        if (off + want > arena_cap) { atomicExch(&st->arena_overflow, 1u); return; }
        general_tally(line, n, p, arena + off, r.bases_len, tot, fwd, rev, &r);
    }
    if (r.status) {                                           // (K1 parsed this line without raising: not expected)
        atomicMax(&st->first_error_inv, ~((goff << 8) | (unsigned long long)r.status));
        return;
    }
    const int cid = contig_find(sites, line + r.chrom_off, r.chrom_len);
    const int32_t site = site_find(sites, cid, r.pos);
    unsigned fail = r.fail;
    if (site >= 0 && (sites.flags[site] & SITE_EXCLUDED)) fail |= FAIL_REGION;      // call_consensus.py:165-168
    snpgpu_vcf_record o;
    o.offset = goff; o.pos = r.pos; o.raw_depth = r.raw_depth;
    o.chrom_off = (uint32_t)r.chrom_off; o.chrom_len = (uint32_t)r.chrom_len;
    o.contig = cid; o.rd = r.rd; o.rdf = r.rdf; o.rdr = r.rdr; o.n_alt = r.n_alt;
    o.ref = r.ref; o.cons = r.base; o.fail = (uint8_t)fail;
    o.flags = (uint8_t)((r.has_depth ? SNPGPU_VCF_HAS_DEPTH : 0) | (r.first_is_ref ? SNPGPU_VCF_FIRST_IS_REF : 0));
    o.alt_index = 0;
    if (r.n_alt) {
        const unsigned long long at = atomicAdd(alt_count, (unsigned long long)r.n_alt);
        o.alt_index = at;
        if (at + r.n_alt <= alt_cap) {
            const unsigned U = up8(r.ref);
            for (uint32_t k = 0; k < r.n_alt; k++) {
                uint32_t ad = 0;
                const unsigned a = next_alt(tot, U, &ad);
                snpgpu_vcf_alt e;
                e.ad = ad; e.adf = fwd[a]; e.adr = rev[a]; e.base = (uint8_t)a; e.pad[0] = e.pad[1] = e.pad[2] = 0;
                alt_out[at + k] = e;
            }
        }
    }
    rec_out[i] = o;
}

int k5_launch_tally(cudaStream_t stream, const uint8_t *text, size_t nbytes, const SiteTable &sites, const CallParams &p,
                    const unsigned long long *offsets, size_t n_rec, snpgpu_vcf_record *rec_out, snpgpu_vcf_alt *alt_out,
                    size_t alt_cap, unsigned long long *alt_count, PileupStatusDev *st, uint8_t *arena, size_t arena_cap) {
    if (!n_rec) return 0;
    k5_tally_kernel<<<(unsigned)((n_rec + 127) / 128), 128, 0, stream>>>(text, nbytes, sites, p, offsets, n_rec, rec_out,
                                                                        alt_out, alt_cap, alt_count, st, arena, arena_cap);
    return 1;
}

// ---- the data lines as text ------------------------------------------------------------------------------------------------
// vcf_writer.py:295-379 builds a PyVCF3 record from the tallies and Writer.write_record prints it:
//   CHROM POS . REF ALT . FILTER NS=1 GT:SDP:RD:AD:RDF:RDR:ADF:ADR:FT gt:sdp:rd:ad:rdf:rdr:adf:adr:ft
// One thread per record, the same code once to count the bytes and once to write them.
constexpr int K5_TEXT_THREADS = 256;
size_t k5_text_blocks(size_t n) { return (n + K5_TEXT_THREADS - 1) / K5_TEXT_THREADS; }

template <bool WRITE>
struct K5Out {
    char *p;
    unsigned long long n;
    __device__ __forceinline__ void ch(char c) { if (WRITE) p[n] = c; n++; }
    __device__ __forceinline__ void str(const char *s) { while (*s) ch(*s++); }
    __device__ __forceinline__ void bytes(const uint8_t *s, uint32_t len) { for (uint32_t i = 0; i < len; i++) ch((char)s[i]); }
    __device__ __forceinline__ void num(unsigned long long v) {
        char d[20];
        int k = 0;
        do { d[k++] = (char)('0' + (int)(v % 10ull)); v /= 10ull; } while (v);
        while (k) ch(d[--k]);
    }
    __device__ __forceinline__ void snum(long long v) {       // (str(int) of a column the pileup may have written with a sign)
        if (v < 0) { ch('-'); num(0ull - (unsigned long long)v); } else num((unsigned long long)v);
    }
};

template <bool WRITE>
__device__ unsigned long long k5_format(const K5TextArgs &a, size_t i, char *dst) {
    const snpgpu_vcf_record r = a.rec[i];
    const snpgpu_vcf_alt *alt = a.alt + r.alt_index;
    const char *ft = a.filter_text + (size_t)(r.fail & (SNPGPU_VCF_FILTER_MASKS - 1)) * SNPGPU_VCF_FILTER_TEXT;
    K5Out<WRITE> o{dst, 0ull};
    o.bytes(a.text + r.offset + r.chrom_off, r.chrom_len);
    o.ch('\t'); o.snum(r.pos); o.str("\t.\t");
    o.ch((char)(a.preserve_ref_case ? r.ref : up8(r.ref)));
    o.ch('\t');
    const bool has_depth = (r.flags & SNPGPU_VCF_HAS_DEPTH) != 0, with_alts = has_depth && r.n_alt != 0;
    if (with_alts) {
        for (uint32_t k = 0; k < r.n_alt; k++) { if (k) o.ch(','); o.ch((char)alt[k].base); }
    } else {
        o.ch('.');
    }
    o.str("\t.\t"); o.str(ft); o.str("\tNS=1\tGT:SDP:RD:AD:RDF:RDR:ADF:ADR:FT\t");
    char gt = '.';                                            // vcf_writer.py:310-339
    if (has_depth) {
        gt = with_alts ? ((r.flags & SNPGPU_VCF_FIRST_IS_REF) ? '0' : '1') : '0';
        if (r.fail) gt = a.failed_snp_gt == '.' ? '.' : (a.failed_snp_gt == '0' ? '0' : '1');
    }
    o.ch(gt); o.ch(':'); o.snum(r.raw_depth); o.ch(':'); o.num(r.rd); o.ch(':');
    if (with_alts) { for (uint32_t k = 0; k < r.n_alt; k++) { if (k) o.ch(','); o.num(alt[k].ad); } } else o.ch('0');
    o.ch(':'); o.num(r.rdf); o.ch(':'); o.num(r.rdr); o.ch(':');
    if (with_alts) { for (uint32_t k = 0; k < r.n_alt; k++) { if (k) o.ch(','); o.num(alt[k].adf); } } else o.ch('0');
    o.ch(':');
    if (with_alts) { for (uint32_t k = 0; k < r.n_alt; k++) { if (k) o.ch(','); o.num(alt[k].adr); } } else o.ch('0');
    o.ch(':'); o.str(ft); o.ch('\n');
    return o.n;
}

// sizes: len[i] and the block's sum
__global__ void __launch_bounds__(K5_TEXT_THREADS) k5_text_len_kernel(const K5TextArgs a) {
    __shared__ unsigned long long ws[K5_TEXT_THREADS / 32];
    const size_t i = (size_t)blockIdx.x * K5_TEXT_THREADS + threadIdx.x;
    unsigned long long n = i < a.n_rec ? k5_format<false>(a, i, nullptr) : 0ull;
    if (i < a.n_rec) a.len[i] = (uint32_t)n;                  // (a line of 4 GiB: not in this world)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0;
        for (int w = 0; w < K5_TEXT_THREADS / 32; w++) s += ws[w];
        a.block_sum[blockIdx.x] = s;
    }
}

// exclusive prefix over the blocks' sums, one block; *total = the text's size
__global__ void __launch_bounds__(1024) k5_text_scan_kernel(unsigned long long *block_sum, size_t nb, unsigned long long *total) {
    __shared__ unsigned long long part[1024];
    const size_t per = (nb + 1023) / 1024;
    const size_t lo = (size_t)threadIdx.x * per, hi = lo + per < nb ? lo + per : nb;
    unsigned long long s = 0;
    for (size_t i = lo; i < hi; i++) s += block_sum[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < 1024; t++) { const unsigned long long v = part[t]; part[t] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (size_t i = lo; i < hi; i++) { const unsigned long long v = block_sum[i]; block_sum[i] = run; run += v; }
}

// the bytes: a record starts where its block starts plus the lengths in front of it inside the block
__global__ void __launch_bounds__(K5_TEXT_THREADS) k5_text_write_kernel(const K5TextArgs a) {
    __shared__ unsigned long long ws[K5_TEXT_THREADS / 32];
    const size_t i = (size_t)blockIdx.x * K5_TEXT_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long mine = i < a.n_rec ? a.len[i] : 0ull;
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    unsigned long long base = a.block_sum[blockIdx.x];
    for (int w = 0; w < warp; w++) base += ws[w];
    if (i < a.n_rec) k5_format<true>(a, i, a.out + base + incl - mine);
}

int k5_launch_text_sizes(cudaStream_t stream, const K5TextArgs &a) {
    const size_t nb = k5_text_blocks(a.n_rec);
    k5_text_len_kernel<<<(unsigned)nb, K5_TEXT_THREADS, 0, stream>>>(a);
    k5_text_scan_kernel<<<1, 1024, 0, stream>>>(a.block_sum, nb, a.total);
    return 2;
}

int k5_launch_text_write(cudaStream_t stream, const K5TextArgs &a) {
    k5_text_write_kernel<<<(unsigned)k5_text_blocks(a.n_rec), K5_TEXT_THREADS, 0, stream>>>(a);
    return 1;
}

}  // namespace snpgpu
