// k5_vcf.cu -- K5: the per-line tallies behind the per-sample consensus VCF.
//
// Replaces what vcf_writer.SingleSampleWriter._make_vcf_record_from_pileup (vcf_writer.py:295-379) reads off a
// pileup.Record: raw depth, the reference base's total / forward / reverse good depth, and the ALT alleles -- every
// other surviving symbol, in most_common_good_bases order (pileup.py:260-266) -- with theirs.  The reference writes
// one VCF record per pileup line its Reader yields (call_consensus.py:161-184): the lines K1 parsed.  K1 (run with
// PileupArgs::rec_off set) lists their file offsets; here they are put in file order and each line is tallied by
// the exact any-input parser (line_general.cuh), one thread per line, on the text in global memory.  A few tens of
// thousands of lines per sample in the default mode; every line with --vcfAllPos.
// The radix sort is a CUB device primitive (library code, like in k2_merge.cu).
#include "internal.h"
#include "line_general.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace snpgpu {

size_t k5_sort_bytes(size_t n) {
    size_t b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b, (const unsigned long long *)nullptr, (unsigned long long *)nullptr, (int64_t)n);
    return b;
}

int k5_sort_offsets(cudaStream_t stream, const unsigned long long *in, unsigned long long *out, size_t n, void *tmp,
                    size_t tmp_bytes) {
    if (n == 0) return 0;
    if (cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, in, out, (int64_t)n, 0, 48, stream) != cudaSuccess) return SNPGPU_E_CUDA;
    return 0;
}

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

}  // namespace snpgpu
