"""`distance` subcommand: pairwise SNP distances from the multi-fasta SNP matrix.

Mirror of snppipeline/distance.py:14-118.  Parsing the fasta and writing the two TSV files stay on the host; the
O(N^2 S) loop over itertools.combinations (distance.py:93-96, utils.calculate_sequence_distance utils.py:1135-1165)
is kernel K4 on the GPU.
"""
from __future__ import annotations

import numpy as np

from . import device
from . import utils


def read_matrix(path):
    """distance.py:76-84: {id: sequence}, id = header line without its leading '>' characters."""
    seqs = {}
    curr = None
    with open(path) as ifile:
        for line in ifile:
            line = line.rstrip("\n")
            if line.startswith(">"):
                curr = line.lstrip(">")
                seqs[curr] = []
            else:
                seqs[curr].append(line)          # KeyError(None) like the reference if data precedes any header
    return {k: "".join(v) for k, v in seqs.items()}


def distance_matrix(seqs, ids):
    """int32 [n, n] of mismatch counts over columns where both bases are ACGT (case-insensitive).
    Unequal lengths follow the reference's loop `for i in range(len(seq1))` over sorted pairs: a later, shorter
    sequence raises IndexError; a later, longer one is compared over the shorter prefix (padding with '-')."""
    lens = [len(seqs[i]) for i in ids]
    if any(lens[k + 1] < lens[k] for k in range(len(lens) - 1)):     # (some later sequence is shorter than an earlier one)
        raise IndexError("string index out of range")
    width = max(lens) if lens else 0
    m = np.full((len(ids), max(width, 1)), ord("-"), dtype=np.uint8)
    for r, i in enumerate(ids):
        s = seqs[i].encode("utf-8", "replace")
        if len(s) != lens[r]:                    # non-ASCII characters: one '?' per character keeps the columns aligned
            s = seqs[i].encode("ascii", "replace")
        m[r, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    return device.context().pairwise_distance(m[:, :width] if width else m[:, :0])


def calculate_snp_distances(args):
    """args: inputFile, forceFlag, pairwiseFile, matrixFile, verbose."""
    utils.print_log_header()
    utils.print_arguments(args)

    input_file = args.inputFile
    pairwise_file = args.pairwiseFile
    matrix_file = args.matrixFile
    if utils.verify_existing_input_files("SNP matrix file", [input_file]) > 0:
        utils.global_error("Error: cannot calculate sequence distances without the snp matrix file.")
    if not pairwise_file and not matrix_file:
        utils.global_error("Error: no output file specified.")

    rebuild_pairwise = pairwise_file and utils.target_needs_rebuild([input_file], pairwise_file)
    rebuild_matrix = matrix_file and utils.target_needs_rebuild([input_file], matrix_file)
    if not (args.forceFlag or rebuild_pairwise or rebuild_matrix):
        utils.verbose_print("Distance files have already been freshly built.  Use the -f option to force a rebuild.")
        return

    seqs = read_matrix(input_file)
    utils.verbose_print("# %s %s" % (utils.timestamp(), "Calculating all pairwise distances"))
    ids = sorted(seqs.keys())
    d = distance_matrix(seqs, ids) if ids else np.zeros((0, 0), np.int32)

    if pairwise_file:
        with open(pairwise_file, "w") as p_out:
            p_out.write("%s\n" % "\t".join(["Seq1", "Seq2", "Distance"]))
            for a, id1 in enumerate(ids):
                row = d[a].tolist()
                p_out.write("".join("%s\t%s\t%i\n" % (id1, id2, row[b]) for b, id2 in enumerate(ids)))
    if matrix_file:
        with open(matrix_file, "w") as m_out:
            m_out.write("\t%s\n" % "\t".join(ids))
            for a, id1 in enumerate(ids):
                m_out.write("%s\t%s\n" % (id1, "\t".join(map(str, d[a].tolist()))))
