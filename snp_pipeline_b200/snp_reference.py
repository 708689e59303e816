"""`snp_reference` subcommand: the reference bases at the SNP positions as a fasta file (referenceSNP.fasta).

Mirror of snppipeline/snp_reference.py:12-70 and utils.write_reference_snp_file (utils.py:1091-1110): same Namespace
fields, files, messages, freshness rule and error protocol.  The reference parses the fasta with Biopython
(SeqIO.to_dict(SeqIO.parse(..., "fasta"))) and writes one record per reference contig, in sorted id order, holding
upper(seq[pos - 1]) for every snplist line of that contig in snplist order; a contig without SNPs still gets its
header line.  The gather runs on the GPU (snpgpu_reference_bases); the fasta text Biopython would print (id only,
60 columns) is produced directly and pinned by the reference's bundled referenceSNP*.fasta files.
"""
from __future__ import annotations

import numpy as np

from . import device
from . import utils


def read_fasta(path):
    """What SeqIO.to_dict(SeqIO.parse(path, "fasta")) holds, as {id: bytes}: lines in front of the first '>' are
    skipped, the id is the first word of the title, whitespace inside sequence lines is dropped, a repeated id raises
    ValueError like SeqIO.to_dict does."""
    records, order = {}, []
    cur = None
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                title = line[1:].rstrip()
                words = title.split(None, 1)
                if not words:
                    raise IndexError("list index out of range")            # Bio.SeqIO.FastaIO: title.split(None, 1)[0]
                cur = words[0].decode("ascii", "replace")
                if cur in records:
                    raise ValueError("Duplicate key '%s'" % cur)
                records[cur] = []
                order.append(cur)
            elif cur is not None:
                records[cur].append(b"".join(line.split()))
    return {k: b"".join(v) for k, v in records.items()}


def reference_snp_text(reference_file_path, snp_list_file_path):
    """The text of referenceSNP.fasta (utils.py:1096-1110)."""
    with open(snp_list_file_path, "r") as snp_list_file:
        position_list = [line.split()[0:2] for line in snp_list_file]
    match_dict = read_fasta(reference_file_path)
    ctx = device.context()
    out = []
    for ordered_id in sorted(match_dict.keys()):
        pos = np.array([int(p[1]) for p in position_list if p[0] == ordered_id], dtype=np.int64)
        seq = np.frombuffer(match_dict[ordered_id], dtype=np.uint8)
        bases = ctx.reference_bases(seq, pos)                 # IndexError where the reference's indexing raises
        ref_str = bases.tobytes().decode("ascii", "replace")
        out.append(">%s\n" % ordered_id)
        for i in range(0, len(ref_str), 60):
            out.append(ref_str[i:i + 60] + "\n")
    return "".join(out)


def create_snp_reference_seq(args):
    """args: referenceFile, forceFlag, snpListFile, snpRefFile, verbose."""
    utils.print_log_header()
    utils.print_arguments(args)

    reference_file = args.referenceFile
    snp_list_file_path = args.snpListFile
    snp_ref_seq_path = args.snpRefFile

    bad_file_count = utils.verify_existing_input_files("Snplist file", [snp_list_file_path])
    if bad_file_count > 0:
        utils.global_error("Error: cannot create the snp reference sequence without the snplist file.")

    bad_file_count = utils.verify_non_empty_input_files("Reference file", [reference_file])
    if bad_file_count > 0:
        utils.global_error("Error: cannot create the snp reference sequence without the reference fasta file.")

    source_files = [reference_file, snp_list_file_path]
    if args.forceFlag or utils.target_needs_rebuild(source_files, snp_ref_seq_path):
        text = reference_snp_text(reference_file, snp_list_file_path)
        with open(snp_ref_seq_path, "w") as snp_reference_file_object:
            snp_reference_file_object.write(text)
    else:
        utils.verbose_print("SNP reference sequence %s has already been freshly built.  Use the -f option to force a rebuild."
                            % snp_ref_seq_path)
