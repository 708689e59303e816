"""`filter_regions` subcommand: drop the SNPs of abnormal regions from the samples' VCF files.

Mirror of snppipeline/filter_regions.py:74-520 (same arguments, messages, rebuild rules and output files).  Reading and
writing the small per-sample VCFs stays on the host; which SNPs fall into an abnormal region -- the dense-window scan
(find_dense_regions, :17-71), the merge of the regions (utils.merge_regions) and the membership test (utils.in_region) --
is kernel K7 (csrc/k7_regions.cu), one launch chain for all samples.
"""
from __future__ import annotations

import os
import shutil
import sys

import numpy as np

from . import device
from . import utils

_HEADER_KINDS = ("##INFO=", "##FORMAT=", "##FILTER=", "##ALT=", "##contig=")


def removed_flags(samples, contig_length_dict, edge_length, window_size_list, max_num_snps_list, filter_across_samples):
    """Which records lie in an abnormal region.

    samples: [[(chrom, pos) in file order]] -- the non-outgroup samples.  Returns one bool array per sample.
    filter_across_samples True: the regions of every sample apply to all of them (filter_regions.py:256-303), else each
    sample is filtered by its own regions (:352-383).  The per-contig sort of filter_regions.py:421 and the edge regions of
    :411-417 are prepared here; the rest is K7."""
    contig_rank, keys, last, owner = {}, [], [], []
    edge_keys, edge_end, have_edges = [], [], set()
    base = 0
    for s, records in enumerate(samples):
        group = 0 if filter_across_samples else s
        if group >= 1 << 16:
            utils.global_error("Error: filter_regions handles at most 65536 samples in mode each.")
        by_contig = {}
        for r, (chrom, pos) in enumerate(records):
            by_contig.setdefault(chrom, []).append((pos, r))
        for chrom, plist in by_contig.items():
            rank = contig_rank.setdefault(chrom, len(contig_rank))
            if rank >= 1 << 16:
                utils.global_error("Error: filter_regions handles at most 65536 contigs.")
            hi = (group << 48) | (rank << 32)
            if (group, rank) not in have_edges:
                have_edges.add((group, rank))
                length = contig_length_dict.get(chrom, sys.maxsize)
                if length <= edge_length * 2:
                    regions = [(0, length)]
                else:
                    regions = [(0, edge_length), (length - edge_length, length)]
                for start, end in regions:
                    if start < 1 << 32:                     # (a contig the fasta does not hold: its far edge is out of reach)
                        edge_keys.append(hi | start)
                        edge_end.append(min(end, (1 << 32) - 1))
            plist.sort()
            for pos, r in plist:
                if not 0 <= pos < 1 << 32:
                    utils.global_error("Error: VCF position %d is out of range." % pos)
                keys.append(hi | pos)
                owner.append((s, r))
            last.extend([base + len(plist) - 1] * len(plist))
            base += len(plist)
    flags = [np.zeros(len(records), dtype=bool) for records in samples]
    if keys:
        removed = device.context().filter_regions(np.array(keys, dtype=np.uint64), np.array(last, dtype=np.uint32),
                                                  max_num_snps_list, window_size_list, np.array(edge_keys, dtype=np.uint64),
                                                  np.array(edge_end, dtype=np.uint32))
        for (s, r), f in zip(owner, removed):
            flags[s][r] = f
    return flags


class _Vcf(object):
    """The pieces of a sample VCF that vcf.Reader / vcf.Writer pass through (filter_regions.py:460-520)."""

    def __init__(self, path):
        with open(path, "r") as f:
            lines = [ln for ln in (raw.rstrip("\r\n") for raw in f) if ln.strip()]
        meta = [ln for ln in lines if ln.startswith("##")]
        rest = [ln for ln in lines if not ln.startswith("##")]
        if not rest:
            raise StopIteration("vcf file %s holds no header line" % path)
        # PyVCF3's Writer re-emits the template's header grouped by kind: plain metadata, INFO, FORMAT, FILTER, ALT, contig
        head = [ln for ln in meta if not ln.startswith(_HEADER_KINDS)]
        for kind in _HEADER_KINDS:
            head += [ln for ln in meta if ln.startswith(kind)]
        self.header = "".join(ln + "\n" for ln in head + rest[:1])
        self.lines = rest[1:]
        self.records = []
        for ln in self.lines:
            row = utils._VCF_ROW_SPLIT.split(ln.strip())
            if len(row) < 8:
                raise IndexError("list index out of range")
            self.records.append((row[0], int(row[1])))


def _open_vcf(vcf_file_path):
    try:
        return _Vcf(vcf_file_path)
    except (OSError, StopIteration):
        utils.sample_error("Error: Cannot open the input vcf file: %s." % vcf_file_path, continue_possible=True)
        return None


def write_outgroup_preserved_and_removed_vcf_files(vcf_file_path, vcf):
    """filter_regions.py:431-457: the outgroup's SNPs are all kept; its removed file holds the header only."""
    preserved_vcf_file_path = vcf_file_path[:-4] + "_preserved.vcf"
    removed_vcf_file_path = vcf_file_path[:-4] + "_removed.vcf"
    try:
        with open(removed_vcf_file_path, "w") as f:
            f.write(vcf.header)
    except OSError:
        if os.path.exists(removed_vcf_file_path):
            os.remove(removed_vcf_file_path)
        utils.sample_error("Error: Cannot create the file for removed SNPs: %s." % removed_vcf_file_path,
                           continue_possible=True)
        return
    shutil.copyfile(vcf_file_path, preserved_vcf_file_path)


def write_preserved_and_removed_vcf_files(vcf_file_path, vcf, flags):
    """filter_regions.py:460-520 with the in_region answers already in flags."""
    preserved_vcf_file_path = vcf_file_path[:-4] + "_preserved.vcf"
    removed_vcf_file_path = vcf_file_path[:-4] + "_removed.vcf"
    try:
        preserved = open(preserved_vcf_file_path, "w")
    except OSError:
        utils.sample_error("Error: Cannot create the file for preserved SNPs: %s." % preserved_vcf_file_path,
                           continue_possible=True)
        return
    try:
        removed = open(removed_vcf_file_path, "w")
    except OSError:
        preserved.close()
        utils.sample_error("Error: Cannot create the file for removed SNPs: %s." % removed_vcf_file_path,
                           continue_possible=True)
        return
    with preserved, removed:
        preserved.write(vcf.header)
        removed.write(vcf.header)
        for line, bad in zip(vcf.lines, flags):
            (removed if bad else preserved).write(line + "\n")


def filter_regions(args):
    """args: sampleDirsFile, refFastaFile, forceFlag, vcfFileName, edgeLength, windowSizeList, maxSnpsList, outGroupFile,
    mode ("all" | "each"), verbose.  filter_regions.py:74-200."""
    utils.print_log_header()
    utils.print_arguments(args)

    sample_directories_list_path = args.sampleDirsFile
    ref_fasta_path = args.refFastaFile
    vcf_file_name = args.vcfFileName
    out_group_list_path = args.outGroupFile
    filter_across_samples = args.mode == "all"

    if utils.verify_non_empty_input_files("File of sample directories", [sample_directories_list_path]) > 0:
        utils.global_error(None)
    with open(sample_directories_list_path, "r") as f:
        unsorted_dirs = [line.rstrip() for line in f]
    sorted_dirs = sorted(d for d in unsorted_dirs if d)

    list_of_vcf_files = [os.path.join(d, vcf_file_name) for d in sorted_dirs]
    bad = utils.verify_non_empty_input_files("VCF file", list_of_vcf_files)
    if bad == len(list_of_vcf_files):
        utils.global_error("Error: all %d VCF files were missing or empty." % bad)
    elif bad > 0:
        utils.sample_error("Error: %d VCF files were missing or empty." % bad, continue_possible=True)

    if utils.verify_non_empty_input_files("Reference file", [ref_fasta_path]) > 0:
        utils.global_error(None)

    outgroup = []
    if out_group_list_path is not None:
        if utils.verify_non_empty_input_files("File of outgroup samples", [out_group_list_path]) > 0:
            utils.global_error(None)
        try:
            with open(out_group_list_path, "r") as f:
                outgroup = sorted(line.rstrip() for line in f)
        except OSError:
            utils.global_error("Error: Cannot open the file containing the list of outgroup samples!")

    try:
        contig_length_dict = utils.fasta_contig_lengths(ref_fasta_path)
    except (OSError, UnicodeDecodeError):
        utils.global_error("Error: cannot open the reference fastq file, or fail to read the contigs in the reference fastq file.")

    # ---- which samples need a rebuild?  (filter_regions.py:238-254 / :335-350)
    input_file_list = [ref_fasta_path] + ([out_group_list_path] if out_group_list_path else [])
    if filter_across_samples:
        input_file_list = input_file_list + list_of_vcf_files      # any changed input rebuilds every sample
    need_rebuild = {}
    for vcf_file_path in list_of_vcf_files:
        inputs = input_file_list if filter_across_samples else input_file_list + [vcf_file_path]
        need_rebuild[vcf_file_path] = (args.forceFlag
                                       or utils.target_needs_rebuild(inputs, vcf_file_path[:-4] + "_preserved.vcf")
                                       or utils.target_needs_rebuild(inputs, vcf_file_path[:-4] + "_removed.vcf"))
    if not any(need_rebuild.values()):
        utils.verbose_print("All preserved and removed vcf files are already freshly built.  Use the -f option to force a rebuild.")
        return

    # ---- read the samples (mode "all" reads every sample: their regions count even when their files are fresh)
    paths, vcfs = [], []
    for vcf_file_path in list_of_vcf_files:
        if not filter_across_samples and not need_rebuild[vcf_file_path]:
            continue
        vcf = _open_vcf(vcf_file_path)
        if vcf is None:
            continue
        sample_id = utils.sample_id_from_file(vcf_file_path)
        utils.verbose_print("Processing sample %s" % sample_id)
        if sample_id in outgroup:
            write_outgroup_preserved_and_removed_vcf_files(vcf_file_path, vcf)
            continue
        paths.append(vcf_file_path)
        vcfs.append(vcf)

    # ---- the abnormal regions and the SNPs inside them: K7
    flags = removed_flags([v.records for v in vcfs], contig_length_dict, args.edgeLength, args.windowSizeList,
                          args.maxSnpsList, filter_across_samples)

    for vcf_file_path, vcf, f in zip(paths, vcfs, flags):
        if need_rebuild[vcf_file_path]:
            write_preserved_and_removed_vcf_files(vcf_file_path, vcf, f)
