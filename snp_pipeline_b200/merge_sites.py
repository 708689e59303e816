"""`merge_sites` subcommand: union of the samples' variant sites -> snplist.txt (+ the filtered sample list).

Mirror of snppipeline/merge_sites.py:12-133.  Reading the small per-sample VCFs stays on the host; the union itself
(the dict-of-lists loop of merge_sites.py:94-116 and the sort of utils.py:1068) is kernel K2: a stable radix sort
of (chrom rank, position) keys + run-length encode on the GPU.
"""
from __future__ import annotations

import os

import numpy as np

from . import device
from . import utils


def merge_sites(args):
    """args: sampleDirsFile, filteredSampleDirsFile, forceFlag, vcfFileName, maxSnps, snpListFile, verbose."""
    utils.print_log_header()
    utils.print_arguments(args)

    sample_directories_list_path = args.sampleDirsFile
    if utils.verify_non_empty_input_files("File of sample directories", [sample_directories_list_path]) > 0:
        utils.global_error(None)
    with open(sample_directories_list_path, "r") as f:
        unsorted_dirs = [line.rstrip() for line in f]
    unsorted_dirs = [d for d in unsorted_dirs if d]
    sorted_dirs = sorted(unsorted_dirs)

    snp_list_file_path = args.snpListFile
    vcf_files = [os.path.join(d, args.vcfFileName) for d in sorted_dirs]
    bad = utils.verify_non_empty_input_files("VCF file", vcf_files)
    if bad == len(vcf_files):
        utils.global_error("Error: all %d VCF files were missing or empty." % bad)
    elif bad > 0:
        utils.sample_error("Error: %d VCF files were missing or empty." % bad, continue_possible=True)

    if not (args.forceFlag or utils.target_needs_rebuild(vcf_files, snp_list_file_path)):
        utils.verbose_print("SNP list %s has already been freshly built.  Use the -f option to force a rebuild."
                            % snp_list_file_path)
        return

    names, per_sample, excluded_dirs = [], [], set()
    for sample_dir, vcf_path in zip(sorted_dirs, vcf_files):
        if not os.path.isfile(vcf_path) or os.path.getsize(vcf_path) == 0:
            continue
        utils.verbose_print("Processing VCF file %s" % vcf_path)
        sample_name = os.path.basename(os.path.dirname(vcf_path))
        positions = list(dict.fromkeys(utils.read_vcf_positions(vcf_path)))     # a set in the reference
        if args.maxSnps >= 0 and len(positions) > args.maxSnps:
            utils.verbose_print("Excluding sample %s having %d snps." % (sample_name, len(positions)))
            excluded_dirs.add(sample_dir)
            continue
        names.append(sample_name)
        per_sample.append(positions)

    text, n_sites = merged_snplist_text(names, per_sample)
    utils.verbose_print("Found %d snp positions across %d sample vcf files." % (n_sites, len(vcf_files)))
    with open(snp_list_file_path, "w") as f:
        f.write(text)
    with open(args.filteredSampleDirsFile, "w") as f:
        for d in unsorted_dirs:          # original order, so HPC array indices stay aligned (merge_sites.py:125-131)
            if d not in excluded_dirs:
                f.write("%s\n" % d)


def merged_snplist_text(names, per_sample):
    """snplist.txt text from per-sample [(chrom, pos)] lists given in sorted-sample-directory order.
    Keys are (rank of the chrom in string order << 32 | pos): ascending key order is utils.py:1068's tuple order."""
    chroms = sorted({c for s in per_sample for c, _ in s})
    rank = {c: i for i, c in enumerate(chroms)}
    for s in per_sample:
        for _, p in s:
            if not 0 <= p < (1 << 31):                 # (the site table's range: call_consensus accepts what this step writes)
                raise ValueError("VCF position %d outside [0, 2^31)" % p)
    keys = np.array([(rank[c] << 32) | p for s in per_sample for c, p in s], dtype=np.uint64)
    samp = np.array([i for i, s in enumerate(per_sample) for _ in s], dtype=np.uint32)
    uniq, cnt, samples = device.context().merge_sites(keys, samp)
    out, o = [], 0
    for k, c in zip(uniq.tolist(), cnt.tolist()):
        who = [names[i] for i in samples[o:o + c].tolist()]
        o += c
        out.append("%s\t%d\t%d\t%s\n" % (chroms[k >> 32], k & 0xffffffff, c, "\t".join(who)))
    return "".join(out), len(out)
