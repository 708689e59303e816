"""`snp_matrix` subcommand: the samples x sites matrix as a multi-fasta file.

Mirror of snppipeline/snp_matrix.py:13-119.  In the reference this step is a verbatim concatenation of the
per-sample consensus.fasta files in sorted-directory order; the gather that fills each matrix row (kernel K3) already
happened inside call_consensus, so this stays file I/O, byte for byte.
"""
from __future__ import annotations

import os

from . import utils


def create_snp_matrix(args):
    """args: sampleDirsFile, forceFlag, consFileName, snpmaFile, verbose."""
    utils.print_log_header()
    utils.print_arguments(args)

    sample_directories_list_filename = args.sampleDirsFile
    if utils.verify_non_empty_input_files("File of sample directories", [sample_directories_list_filename]) > 0:
        utils.global_error(None)
    with open(sample_directories_list_filename, "r") as f:
        dirs = [line.rstrip() for line in f]
    dirs = sorted([d for d in dirs if d])

    consensus_files, bad = [], 0
    for d in dirs:
        path = os.path.join(d, args.consFileName)
        if utils.verify_non_empty_input_files("Consensus fasta file", [path]) == 1:
            bad += 1
        else:
            consensus_files.append(path)
    if bad == len(dirs):
        utils.global_error("Error: all %d consensus fasta files were missing or empty." % bad)
    elif bad > 0:
        utils.sample_error("Error: %d consensus fasta files were missing or empty." % bad, continue_possible=True)

    snpma_file_path = args.snpmaFile
    if not args.forceFlag and not utils.target_needs_rebuild(consensus_files, snpma_file_path):
        utils.verbose_print("SNP matrix %s has already been freshly built.  Use the -f option to force a rebuild."
                            % snpma_file_path)
        return

    with open(snpma_file_path, "w") as output_file:
        for path in consensus_files:
            utils.verbose_print("Merging " + path)
            with open(path, "r") as input_file:
                for line in input_file:
                    output_file.write(line)
