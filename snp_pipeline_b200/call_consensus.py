"""`call_consensus` subcommand: consensus base of one sample at the snplist positions -> consensus.fasta.

Mirror of snppipeline/call_consensus.py:18-192 (same Namespace fields, files, messages and error protocol); the
per-line work (pileup parse, tally, filters, call, gather in snplist order) runs in kernel K1/K3 on the GPU, the
tallies of the optional consensus VCF (--vcfFileName) in K5.
"""
from __future__ import annotations

import os

from . import pileup
from . import utils
from . import vcf_writer


def call_consensus(args):
    """args: allPileupFile, forceFlag, snpListFile, excludeFile, consensusFile, minBaseQual, minConsFreq, minConsDpth,
    minConsStrdDpth, minConsStrdBias, vcfFileName, vcfRefName, vcfAllPos, vcfPreserveRefCase, vcfFailedSnpGt,
    verbose (cfsan_snp_pipeline.py:345-407)."""
    utils.print_log_header()
    utils.print_arguments(args)

    snp_list_file_path = args.snpListFile
    all_pileup_file_path = args.allPileupFile
    sample_directory = os.path.dirname(os.path.abspath(all_pileup_file_path))
    sample_name = os.path.basename(sample_directory)
    consensus_file_path = args.consensusFile
    vcf_file_name = getattr(args, "vcfFileName", None)

    if utils.verify_existing_input_files("Snplist file", [snp_list_file_path]) > 0:
        utils.global_error("Error: cannot call consensus without the snplist file.")
    if utils.verify_non_empty_input_files("Pileup file", [all_pileup_file_path]) > 0:
        utils.sample_error("Error: cannot call consensus without the pileup file.", continue_possible=False)

    source_files = [snp_list_file_path, all_pileup_file_path]
    exclude_file_path = getattr(args, "excludeFile", None)
    if exclude_file_path:
        if utils.verify_existing_input_files("Exclude file", [exclude_file_path]) > 0:
            utils.sample_error("Error: cannot call consensus without the file of excluded positions.",
                               continue_possible=False)
        excluded_positions = utils.convert_vcf_file_to_snp_set(exclude_file_path)
        source_files.append(exclude_file_path)
    else:
        excluded_positions = set()

    if not args.forceFlag and not utils.target_needs_rebuild(source_files, consensus_file_path):
        utils.verbose_print("Consensus call file %s has already been freshly built.  Use the -f option to force a "
                            "rebuild." % consensus_file_path)
        return

    snp_list = utils.read_snp_position_list(snp_list_file_path)
    utils.verbose_print("snp position list length = %d" % len(snp_list))
    utils.verbose_print("excluded snps list length = %d" % len(excluded_positions))
    utils.verbose_print("total snp position list length = %d" % (len(snp_list) + len(excluded_positions)))

    caller = pileup.ConsensusCaller(args.minConsFreq, args.minConsDpth, args.minConsStrdDpth, args.minConsStrdBias)
    if getattr(args, "vcfAllPos", False):
        parse_positions = None
    else:
        parse_positions = set(snp_list).union(excluded_positions)
    reader = pileup.Reader(all_pileup_file_path, args.minBaseQual, parse_positions)
    writer = None
    if vcf_file_name:
        consensus_file_dir = os.path.dirname(os.path.abspath(consensus_file_path))
        vcf_file_path = os.path.join(consensus_file_dir, vcf_file_name)
        writer = vcf_writer.SingleSampleWriter(vcf_file_path, getattr(args, "vcfPreserveRefCase", False))
        filters = caller.get_filter_descriptions()
        filters.append(("Region", "Position is in dense region of snps or near the end of the contig."))
        writer.write_header(sample_name, filters, getattr(args, "vcfRefName", "Unknown reference"))

    try:
        consensus_str, stats = reader.call_consensus(caller, snp_list, excluded_positions, vcf_writer=writer,
                                                     failed_snp_gt=getattr(args, "vcfFailedSnpGt", "."))
    finally:
        if writer:
            writer.close()
    # call_consensus.py:184 counts the entries of its position -> base dict: the snplist positions that received a call
    # (a position the pileup never reaches, or one whose lines all were overwritten by '-', still counts when called)
    utils.verbose_print("called consensus positions = %i" % stats.n_called)
    utils.verbose_print("parsed pileup lines = %i of %i" % (stats.n_parsed, stats.n_lines))

    with open(consensus_file_path, "w") as fasta_file_object:
        fasta_file_object.write(utils.fasta_record_text(sample_name, consensus_str))
