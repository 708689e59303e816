/*
 * snpgpu.h -- C ABI of libsnpgpu.so: the B200 (sm_100a) implementation of the CFSAN SNP Pipeline's
 * pileup -> consensus -> site-union -> SNP-matrix -> pairwise-distance hot path.
 *
 * The reference (CFSAN-Biostatistics/snp-pipeline) is pure Python with no FFI of its own; the surface it
 * offers for this path is four subcommand functions that talk through files.  Each
 * entry point below replaces the arithmetic of one reference function and is what a ctypes binding inside
 * that function would call (INTEGRATION.md shows the stubs).  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 on success or an SNPGPU_E_* code; snpgpu_last_error() gives the text.
 *   - "_dev" pointers are CUDA device pointers in the context's device; everything else is host memory.
 *   - work is enqueued on the context's stream (snpgpu_set_stream); calls that return results to host
 *     memory synchronise that stream before returning, "_dev"/"_async" calls do not.
 *   - one context per host thread / per GPU; a context is not thread-safe.
 */
#ifndef SNPGPU_H
#define SNPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNPGPU_ABI_VERSION 2

/* ---- status codes ----------------------------------------------------------------------------- */
#define SNPGPU_OK            0
#define SNPGPU_E_VALUE       1   /* the reference raises ValueError (malformed integer column)              */
#define SNPGPU_E_INDEX       2   /* the reference raises IndexError (too few columns for pileup.Record)      */
#define SNPGPU_E_UNPACK      3   /* the reference raises ValueError (fewer than 2 columns, filter mode)      */
#define SNPGPU_E_DOMAIN      4   /* input outside the byte / int64 domain (non-ASCII byte, |int| >= 2^63,    */
                                 /* reference-base column not exactly one byte)                              */
#define SNPGPU_E_LONECR      5   /* a line ends in a lone CR (classic Mac).  Python's universal newlines end a   */
                                 /* line there (pileup.py:417); the kernels split on LF only.  The host-buffer   */
                                 /* entry point rewrites such CRs to LF in its device copy and runs again; users */
                                 /* of the _dev entry point call snpgpu_normalize_newlines_dev and repeat         */
#define SNPGPU_E_CUDA        16  /* a CUDA runtime call or kernel failed                                     */
#define SNPGPU_E_ARG         17  /* bad argument (null pointer, misaligned device buffer, size overflow)     */
#define SNPGPU_E_NOMEM       18  /* device or host allocation failed                                         */
#define SNPGPU_E_LENGTH      19  /* distance: a later (sorted) sequence is shorter than an earlier one --    */
                                 /* the reference raises IndexError at utils.py:1158                         */

/* ---- fail-filter bits (pileup.py:556-584, call_consensus.py:165-168) ---------------------------- */
#define SNPGPU_FAIL_RAWDPTH  1
#define SNPGPU_FAIL_VARFREQ  2
#define SNPGPU_FAIL_DEPTH    4
#define SNPGPU_FAIL_STRDPTH  8
#define SNPGPU_FAIL_STRBIAS 16
#define SNPGPU_FAIL_REGION  32

typedef struct snpgpu_ctx snpgpu_ctx;
typedef struct snpgpu_sites snpgpu_sites;

/* ---- context ---------------------------------------------------------------------------------- */
int         snpgpu_abi_version(void);
int         snpgpu_create(int device, snpgpu_ctx **out);
void        snpgpu_destroy(snpgpu_ctx *ctx);
const char *snpgpu_last_error(const snpgpu_ctx *ctx);
/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL = the context's own stream */
int         snpgpu_set_stream(snpgpu_ctx *ctx, void *stream);
int         snpgpu_sync(snpgpu_ctx *ctx);
/* pinned host memory for the streaming entry points */
int         snpgpu_host_alloc(snpgpu_ctx *ctx, size_t nbytes, void **out);
int         snpgpu_host_free(snpgpu_ctx *ctx, void *p);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
uint64_t    snpgpu_launch_count(const snpgpu_ctx *ctx);

/* Per-kernel device timing (bench.py's roofline): when enabled, every launch of the named kernel is bracketed by
 * CUDA events on the context's stream.  snpgpu_kernel_time synchronises the stream, returns the summed device time
 * and the number of launches since the last call, and resets both. */
#define SNPGPU_KERNEL_PILEUP   0   /* k1_pileup_kernel + k1_rest_kernel (one pair per batch of up to 64 samples) */
#define SNPGPU_KERNEL_DISTANCE 1   /* k4_pairs_kernel (+ its pack kernel) */
int         snpgpu_enable_timing(snpgpu_ctx *ctx, int on);
int         snpgpu_kernel_time(snpgpu_ctx *ctx, int kernel, double *ms_out, uint64_t *launches_out);

/* ---- consensus-caller parameters: ConsensusCaller.__init__ (pileup.py:433-471) + Reader's
 *      min_base_quality (pileup.py:389-407) -------------------------------------------------------- */
typedef struct {
    int32_t min_base_qual;          /* call_consensus -q  (cfsan_snp_pipeline.py:397) */
    int32_t min_cons_depth;         /* -D */
    int32_t min_cons_strand_depth;  /* -d */
    int32_t reserved;
    double  min_cons_freq;          /* -c, compared in IEEE double like the reference (pileup.py:564) */
    double  min_cons_strand_bias;   /* -b, pileup.py:580 */
} snpgpu_params;

/* ---- site table: snplist.txt entries (utils.py:1073-1088) plus the exclude VCF's positions
 *      (call_consensus.py:117-125).  snp entries keep file order and may repeat. ------------------- */
int  snpgpu_sites_create(snpgpu_ctx *ctx,
                         const char *contig_names, const int32_t *name_off, int32_t n_contigs,
                         const int32_t *snp_contig, const int64_t *snp_pos, size_t n_snp,
                         const int32_t *exc_contig, const int64_t *exc_pos, size_t n_exc,
                         snpgpu_sites **out);
/* The same table built on the device from K2's output, when merge_sites and call_consensus run in one process: the
 * list that the reference writes to snplist.txt (utils.py:1056-1070) and reads back (utils.py:1073-1088,
 * call_consensus.py:133) stays in HBM.  keys_dev: n_keys sorted unique (chrom_rank << 32 | pos) keys in device memory,
 * as snpgpu_merge_sites_dev leaves them; every key is a snplist entry, in that order; contig_len[c] bounds the
 * positions of contig c (chrom_rank c).  Nothing is copied back, nothing is synchronised; the table is valid for work
 * enqueued afterwards on the context's stream, and snpgpu_sites_destroy hands its memory back in stream order. */
int  snpgpu_sites_create_from_keys_dev(snpgpu_ctx *ctx, const char *contig_names, const int32_t *name_off,
                                       int32_t n_contigs, const int64_t *contig_len, const uint64_t *keys_dev,
                                       size_t n_keys, snpgpu_sites **out);
void snpgpu_sites_destroy(snpgpu_sites *sites);
size_t snpgpu_sites_n_snp(const snpgpu_sites *sites);

/* ---- Reference bases at the snplist positions.  Replaces the gather of utils.write_reference_snp_file
 *      (utils.py:1100-1108: match_dict[id][int(pos) - 1].upper() per snplist entry of that contig).
 *      out[k] = upper(seq[pos[k] - 1]) with Python's indexing (position 0 and negative positions count from the
 *      end); a position outside the sequence -> SNPGPU_E_INDEX (the reference's IndexError), *bad_index = the
 *      first such k.  Host buffers; copies inside the call. ------------------------------------------------ */
int snpgpu_reference_bases(snpgpu_ctx *ctx, const uint8_t *seq, size_t seq_len, const int64_t *pos, size_t n,
                           uint8_t *out, size_t *bad_index);

/* ---- K1: pileup text -> consensus cells.  Replaces pileup.Reader.__iter__ + Record + ConsensusCaller +
 *      the loop body of call_consensus.py:161-188.
 *   mode SNPGPU_MODE_SITES   only lines whose (chrom,pos) is in the site table are parsed (pileup.py:423-429)
 *   mode SNPGPU_MODE_ALL     every line is parsed (--vcfAllPos, pileup.py:419-421); line_out, when given,
 *                            receives one uint16 per line in file order: low byte = matrix cell, high byte =
 *                            fail mask.
 *   row_out[n_snp]           the consensus string in snplist order ('-' where nothing was called).
 *   stats (nullable)         see snpgpu_pileup_stats.
 * On SNPGPU_E_VALUE/INDEX/UNPACK/DOMAIN stats->error_offset is the byte offset of the first offending
 * line in file order -- the line at which the reference's exception would surface. ----------------- */
#define SNPGPU_MODE_SITES 0
#define SNPGPU_MODE_ALL   1

typedef struct {
    uint64_t n_lines;        /* lines seen                                              */
    uint64_t n_parsed;       /* lines that went through tally + consensus call          */
    uint64_t n_general;      /* of those, lines that needed the general (non-fast) path */
    uint64_t error_offset;   /* byte offset of the first raising line, or UINT64_MAX    */
    int32_t  error_code;
    int32_t  reserved;
    uint64_t n_called;       /* snplist positions that received a call: call_consensus.py:184's "called consensus positions" */
} snpgpu_pileup_stats;

/* text in host memory (pinned or pageable); copies are streamed inside the call */
int snpgpu_pileup_consensus(snpgpu_ctx *ctx, const void *text, size_t nbytes,
                            const snpgpu_sites *sites, const snpgpu_params *params, int mode,
                            uint8_t *row_out, uint16_t *line_out, size_t line_out_cap,
                            snpgpu_pileup_stats *stats);

/* The same call in two halves, for callers that stream many samples (the reference runs one call_consensus process per
 * sample, run.py:709-710): _begin enqueues the copy in, the kernels and the copies out on one of the context's two
 * internal lanes and returns at once; _end waits for that call and reports exactly like snpgpu_pileup_consensus.
 * Keeping one call ahead lets the next sample's text cross PCIe while this sample's kernels run and its results go
 * back.  At most two calls in flight; text / row_out / line_out / stats must stay valid until _end and should be
 * page-locked (snpgpu_host_alloc), pageable memory makes the copies synchronous.  snpgpu_pileup_vcf_records does not
 * apply to pipelined calls. */
int snpgpu_pileup_consensus_begin(snpgpu_ctx *ctx, const void *text, size_t nbytes, const snpgpu_sites *sites,
                                  const snpgpu_params *params, int mode, uint8_t *row_out, uint16_t *line_out,
                                  size_t line_out_cap, snpgpu_pileup_stats *stats, int *slot_out);
int snpgpu_pileup_consensus_end(snpgpu_ctx *ctx, int slot);

/* text already in device memory (16-byte aligned).  row_out_dev: n_snp bytes.  line_out_dev: nullable,
 * one uint16 per line in file order, capacity line_out_cap.  stats_dev: nullable device copy of
 * snpgpu_pileup_stats (valid after the stream is synchronised).  Nothing is synchronised here. */
int snpgpu_pileup_consensus_dev(snpgpu_ctx *ctx, const void *text_dev, size_t nbytes,
                                const snpgpu_sites *sites, const snpgpu_params *params, int mode,
                                uint8_t *row_out_dev, uint16_t *line_out_dev, size_t line_out_cap,
                                snpgpu_pileup_stats *stats_dev);

/* The same for a batch of samples that share the site table and the parameters (the reference runs one call_consensus
 * process per sample, run.py:709-710): one launch sequence per up to 64 samples (split evenly), so the fixed costs of a launch (its tail, the
 * small kernels around it) are paid once per batch.  samples: host array; every pointer in it is a device pointer with
 * the meaning it has in snpgpu_pileup_consensus_dev.  Nothing is synchronised. */
typedef struct {
    const void          *text_dev;       /* 16-byte aligned */
    size_t               nbytes;
    uint8_t             *row_out_dev;    /* n_snp bytes */
    uint16_t            *line_out_dev;   /* nullable (SNPGPU_MODE_ALL only) */
    size_t               line_out_cap;
    snpgpu_pileup_stats *stats_dev;      /* nullable */
} snpgpu_pileup_sample;

int snpgpu_pileup_consensus_batch_dev(snpgpu_ctx *ctx, const snpgpu_pileup_sample *samples, size_t n_samples,
                                      const snpgpu_sites *sites, const snpgpu_params *params, int mode);

/* ---- K5: per-line tallies for the per-sample consensus VCF.  Replaces what
 *      vcf_writer.SingleSampleWriter._make_vcf_record_from_pileup (vcf_writer.py:295-379) reads off each
 *      pileup.Record that call_consensus.py:161-184 hands it: one record per pileup line that K1 parsed (lines at
 *      snplist / exclude positions in SNPGPU_MODE_SITES, every line in SNPGPU_MODE_ALL), in file order.
 *      ALT alleles are every surviving symbol other than REF.upper(), in most_common_good_bases order
 *      (pileup.py:260-266); a record's alleles are alt_out[alt_index .. alt_index + n_alt).
 *      Must follow a successful snpgpu_pileup_consensus() on the same context with the same sites / params /
 *      mode: it works on the device copy of that call's text.  When a capacity is too small the call returns
 *      SNPGPU_E_NOMEM with *n_rec / *n_alt set to what is needed. ------------------------------------------ */
#define SNPGPU_VCF_HAS_DEPTH     1   /* most_common_good_bases is not None                   */
#define SNPGPU_VCF_FIRST_IS_REF  2   /* most_common_good_bases[0] == REF.upper()             */

typedef struct {
    uint64_t offset;       /* byte offset of the pileup line                                  */
    int64_t  pos;          /* column 2                                                        */
    int64_t  raw_depth;    /* column 4 (SDP)                                                  */
    uint64_t alt_index;    /* first ALT entry of this record in alt_out                       */
    uint32_t chrom_off;    /* column 1 is text[offset + chrom_off .. + chrom_len)             */
    uint32_t chrom_len;
    int32_t  contig;       /* index of column 1 in the site table's contig names, or -1       */
    uint32_t rd, rdf, rdr; /* good depth of REF.upper(): total, forward, reverse              */
    uint32_t n_alt;
    uint8_t  ref;          /* column 3 as written                                             */
    uint8_t  cons;         /* consensus base (pileup.py:586-588), '-' without good depth      */
    uint8_t  fail;         /* SNPGPU_FAIL_* mask, Region included                             */
    uint8_t  flags;        /* SNPGPU_VCF_*                                                    */
} snpgpu_vcf_record;

typedef struct {
    uint32_t ad, adf, adr; /* good depth of the allele: total, forward, reverse               */
    uint8_t  base;         /* the allele (upper-cased symbol)                                 */
    uint8_t  pad[3];
} snpgpu_vcf_alt;

/* Tell the context that snpgpu_pileup_vcf_records will follow the next snpgpu_pileup_consensus calls (call_consensus
 * --vcfFileName, which run.py:709 always passes): in SNPGPU_MODE_SITES the call then lists the lines it parsed on the
 * way (they go through the follow-up kernel anyway) and K5 does not have to run K1 a second time. */
int snpgpu_pileup_want_vcf_records(snpgpu_ctx *ctx, int on);

int snpgpu_pileup_vcf_records(snpgpu_ctx *ctx, const snpgpu_sites *sites, const snpgpu_params *params, int mode,
                              snpgpu_vcf_record *rec_out, size_t rec_cap, size_t *n_rec,
                              snpgpu_vcf_alt *alt_out, size_t alt_cap, size_t *n_alt);

/* The same records as the text of the VCF's data lines, in file order, formatted on the device: what
 * vcf_writer.py:295-379 (_make_vcf_record_from_pileup) + PyVCF3's Writer.write_record print per pileup record:
 *   CHROM POS . REF ALT . FILTER NS=1 GT:SDP:RD:AD:RDF:RDR:ADF:ADR:FT gt:sdp:rd:ad:rdf:rdr:adf:adr:ft
 * filter_text: SNPGPU_VCF_FILTER_MASKS NUL-terminated strings of SNPGPU_VCF_FILTER_TEXT bytes each -- the FILTER column of
 * every SNPGPU_FAIL_* mask ("PASS" for 0, else the caller's filter names joined with ';', pileup.py:550-588 /
 * call_consensus.py:165-168); failed_snp_gt: '.', '0' or '1' (--vcfFailedSnpGt); preserve_ref_case: --vcfPreserveRefCase.
 * *n_text = bytes of text (SNPGPU_E_NOMEM when text_cap is smaller: the text stays on the device until the next call,
 * which only copies it when it brings *n_text bytes of room). */
#define SNPGPU_VCF_FILTER_MASKS 64
#define SNPGPU_VCF_FILTER_TEXT  64
int snpgpu_pileup_vcf_text(snpgpu_ctx *ctx, const snpgpu_sites *sites, const snpgpu_params *params, int mode,
                           const char *filter_text, int failed_snp_gt, int preserve_ref_case, char *text_out, size_t text_cap,
                           size_t *n_text, size_t *n_rec);

/* ---- K7: which SNPs lie in an abnormal region.  Replaces find_dense_regions (filter_regions.py:17-71),
 *      utils.merge_regions (utils.py:1168-1282) and utils.in_region (utils.py:1285-1318) as filter_regions.py:296-303,
 *      375-383, 386-428 use them.  snp_keys[i] = group << 48 | contig rank << 32 | position, the SNPs of one sample's contig
 *      forming a segment sorted by position (filter_regions.py:421), seg_last[i] = index of the last SNP of i's segment;
 *      group = 0 for every sample in mode "all", the sample's index in mode "each".  For every (max_snps[q], window[q]) pair
 *      a SNP whose window holds more than max_snps[q] SNPs starts a dense region; edge_keys / edge_end give the contigs' edge
 *      regions (same key layout, start in the position field).  removed_out[i] = 1 when SNP i lies in a dense or an edge
 *      region of its (group, contig).  Host buffers; copies inside the call. ------------------------------------ */
int snpgpu_filter_regions(snpgpu_ctx *ctx, const uint64_t *snp_keys, const uint32_t *seg_last, size_t n, const int32_t *max_snps,
                          const int32_t *window, int32_t n_params, const uint64_t *edge_keys, const uint32_t *edge_end,
                          size_t n_edges, uint8_t *removed_out);

/* ---- K6: the pileup by-product of collect_metrics.  Replaces the loop of collect_metrics.py:322-329: the sum over all
 *      lines of int(line.split()[3]) -- the raw-depth column -- where lines without a fourth token or with one that is
 *      no integer add nothing; the caller divides by the reference length and prints "%.2f" (collect_metrics.py:333-338).
 *      lines_out (nullable): how many lines contributed.  SNPGPU_E_DOMAIN (a byte >= 0x80, an integer beyond int64):
 *      *error_offset = byte offset of the first such line. ------------------------------------------------------- */
int snpgpu_pileup_depth_sum(snpgpu_ctx *ctx, const void *text, size_t nbytes, int64_t *sum_out, uint64_t *lines_out,
                            uint64_t *error_offset);
int snpgpu_pileup_depth_sum_dev(snpgpu_ctx *ctx, const void *text_dev, size_t nbytes, int64_t *sum_out, uint64_t *lines_out,
                                uint64_t *error_offset /* host; synchronises */);

/* Rewrites every CR that is not followed by LF to LF, in place (byte offsets and line structure under
 * universal newlines are unchanged).  Needed only after SNPGPU_E_LONECR. */
int snpgpu_normalize_newlines_dev(snpgpu_ctx *ctx, void *text_dev, size_t nbytes);

/* ---- K2: union of variant sites.  Replaces the dict-of-lists loop of merge_sites.py:94-116 and the
 *      sorted() of utils.py:1068.  keys[i] = (chrom_rank << 32) | pos with chrom_rank in chromosome
 *      *string* order; sample_of[i] = index of the owning sample in sorted-sample-directory order, samples
 *      concatenated in that order.  Outputs (caller-allocated, capacity n each): unique keys ascending,
 *      per-key sample count, and sample indices grouped by key (input order kept inside a group). ------ */
int snpgpu_merge_sites(snpgpu_ctx *ctx, const uint64_t *keys, const uint32_t *sample_of, size_t n,
                       uint64_t *uniq_out, uint32_t *count_out, uint32_t *samples_out, size_t *n_uniq_out);
int snpgpu_merge_sites_dev(snpgpu_ctx *ctx, const uint64_t *keys_dev, const uint32_t *sample_of_dev, size_t n,
                           uint64_t *uniq_out_dev, uint32_t *count_out_dev, uint32_t *samples_out_dev,
                           size_t *n_uniq_out /* host; synchronises */);

/* ---- K4: all-pairs SNP distance.  Replaces utils.calculate_sequence_distance (utils.py:1135-1165) over
 *      itertools.combinations (distance.py:93-96).  matrix: n_rows x row_stride bytes, row i holding
 *      n_sites valid bytes; only columns where both bytes are in {A,C,G,T} (case-insensitive) count.
 *      dist_out: n_rows x n_rows int32, symmetric, zero diagonal.
 *      The stripe variant computes rows [row_begin,row_end) of the output only (multi-GPU sharding). ---- */
int snpgpu_pairwise_distance(snpgpu_ctx *ctx, const uint8_t *matrix, size_t n_rows, size_t n_sites,
                             size_t row_stride, int32_t *dist_out);
int snpgpu_pairwise_distance_dev(snpgpu_ctx *ctx, const uint8_t *matrix_dev, size_t n_rows, size_t n_sites,
                                 size_t row_stride, size_t row_begin, size_t row_end, int32_t *dist_out_dev);
/* Whole 64-row tile rows, upper part only: for every entry t of tile_rows (host array) the rows [64 t, 64 t + 64) are
 * computed from their diagonal tile rightwards -- cells to the left of it are left as they are or zeroed -- into 64 consecutive rows of
 * dist_out_dev (n_tile_rows x 64 x n_rows int32); the caller mirrors (the distance is symmetric: utils.py:1135-1165
 * counts positions where the two bases differ).  Multi-GPU drivers deal the tile rows to the ranks in zigzag order, so
 * that the ranks share the triangle's work evenly and no pair is computed twice. */
int snpgpu_pairwise_distance_tiles_dev(snpgpu_ctx *ctx, const uint8_t *matrix_dev, size_t n_rows, size_t n_sites,
                                       size_t row_stride, const uint32_t *tile_rows, size_t n_tile_rows,
                                       int32_t *dist_out_dev);

/* ---- synthetic pileup generator (bench / tests only; SURVEY.md section 8d's input spec).
 *      Writes one sample's pileup text into text_dev (capacity cap bytes) and returns its length.  The text
 *      depends only on (seed, sample, the arguments) so any sample can be regenerated on any rank. -------- */
typedef struct {
    uint64_t seed;
    uint32_t sample;
    uint32_t genome_len;       /* positions 1..genome_len on one contig                              */
    uint32_t mean_depth;       /* ~Poisson-shaped, clipped to [0, 60]                                */
    uint32_t n_pool_sites;     /* size of the global variant-site pool (same for every sample)       */
    float    site_carry_prob;  /* probability that this sample carries a given pool site             */
    float    indel_line_rate;  /* fraction of lines that carry one indel token; 0 = the default 0.001  */
} snpgpu_synth_spec;

int snpgpu_synth_pileup_dev(snpgpu_ctx *ctx, const snpgpu_synth_spec *spec, const char *contig_name,
                            void *text_dev, size_t cap, size_t *nbytes_out /* host; synchronises */);
/* positions (1-based, ascending) of the pool sites this sample carries; returns how many */
int snpgpu_synth_sample_sites(snpgpu_ctx *ctx, const snpgpu_synth_spec *spec, uint32_t *pos_out, size_t cap,
                              size_t *n_out);

#ifdef __cplusplus
}
#endif
#endif /* SNPGPU_H */
