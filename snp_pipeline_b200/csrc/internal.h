// internal.h -- launchers shared between the translation units of libsnpgpu (not part of the C ABI).
#pragma once
#include "common.cuh"

namespace snpgpu {

// k1_pileup.cu
size_t k1_smem_bytes();
int    k1_blocks_per_sm();
// the pileup kernel + its follow-up kernel over a batch; returns the number of kernels launched
int    k1_launch(cudaStream_t stream, const K1Batch &g, int grid_blocks, int n_sms);
int    k1_launch_finish(cudaStream_t stream, const K1Batch &g, int max_tiles, bool want_lines);
int    k1_launch_normalize(cudaStream_t stream, uint8_t *text, size_t nbytes);

// k5_vcf.cu
struct K5TextArgs {
    const uint8_t *text;                  // the staged pileup text (the chromosome column is copied from it)
    const snpgpu_vcf_record *rec;
    const snpgpu_vcf_alt *alt;
    size_t n_rec;
    const char *filter_text;              // [SNPGPU_VCF_FILTER_MASKS][SNPGPU_VCF_FILTER_TEXT]
    char failed_snp_gt;
    bool preserve_ref_case;
    uint32_t *len;                        // per record
    unsigned long long *block_sum, *total;
    char *out;
};
size_t k5_text_blocks(size_t n);
int    k5_launch_text_sizes(cudaStream_t stream, const K5TextArgs &a);
int    k5_launch_text_write(cudaStream_t stream, const K5TextArgs &a);
int    k5_launch_tally(cudaStream_t stream, const uint8_t *text, size_t nbytes, const SiteTable &sites, const CallParams &p,
                       const unsigned long long *offsets, size_t n_rec, snpgpu_vcf_record *rec_out, snpgpu_vcf_alt *alt_out,
                       size_t alt_cap, unsigned long long *alt_count, PileupStatusDev *st, uint8_t *arena, size_t arena_cap);

// k6_metrics.cu
int    k6_launch_depth_sum(cudaStream_t stream, const uint8_t *text, size_t nbytes, unsigned long long *out3);

// k3_sites.cu
size_t k3_scan_bytes(size_t n_words);
int    k3_launch(cudaStream_t stream, const unsigned long long *keys, size_t n, int n_contigs, const int64_t *bit_base,
                 const int64_t *max_pos, uint32_t *bits, uint32_t *rank, size_t n_words, uint8_t *flags,
                 int32_t *snp_unique, SiteWord *words, void *tmp, size_t tmp_bytes);

int    k3_launch_reference_bases(cudaStream_t stream, const uint8_t *seq, size_t seq_len, const long long *pos, size_t n,
                                 uint8_t *out, unsigned long long *first_bad);

// k2_merge.cu
// sorted-unique union of keys with per-key sample lists; all pointers device; tmp: workspace owned by the caller
size_t k2_workspace_bytes(size_t n);
int    k2_launch(cudaStream_t stream, const uint64_t *keys, const uint32_t *sample_of, size_t n, uint64_t *uniq_out,
                 uint32_t *count_out, uint32_t *samples_out, unsigned long long *n_uniq_dev, void *tmp,
                 size_t tmp_bytes, int *launches);

int    k2_sort_pairs(cudaStream_t stream, const uint64_t *keys, const uint32_t *vals, size_t n, uint32_t *vals_out, void *tmp,
                     size_t tmp_bytes, const unsigned long long **sorted_keys, void **spare, uint32_t **hist_out, int *launches);

// k7_regions.cu
int    k7_launch(cudaStream_t stream, const uint64_t *snp_keys, const uint32_t *seg_last, size_t n, const int32_t *max_snps,
                 const int32_t *window, int n_params, const uint64_t *extra_keys, const uint32_t *extra_end, size_t n_extra,
                 uint8_t *removed_out, void *tmp, size_t tmp_bytes, int *launches);
size_t k7_workspace_bytes(size_t n, int n_params, size_t n_extra);

// k4_distance.cu
size_t k4_workspace_bytes(size_t n_rows, size_t n_sites);
int    k4_launch(cudaStream_t stream, const uint8_t *matrix, size_t n_rows, size_t n_sites, size_t row_stride,
                 size_t row_begin, size_t row_end, const uint32_t *tiles, size_t n_tiles, int32_t *dist_out, void *tmp,
                 int n_sms, int *launches);

// synth.cu
int    synth_launch(cudaStream_t stream, const snpgpu_synth_spec &spec, const char *contig_name, uint8_t *text_dev,
                    size_t cap, unsigned long long *nbytes_dev, void *tmp, size_t tmp_bytes, int *launches);
size_t synth_workspace_bytes(uint32_t genome_len);
void   synth_host_sites(const snpgpu_synth_spec &spec, uint32_t *pos_out, size_t cap, size_t *n_out);

}  // namespace snpgpu
