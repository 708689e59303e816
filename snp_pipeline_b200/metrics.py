"""The two by-products of the hot path that the reference's `collect_metrics` step reads back out of its files.

Mirror of snppipeline/collect_metrics.py:
  * :313-342  "avePileupDepth": the sum of the pileup file's raw-depth column over all lines, divided by the reference
    length, printed as "%.2f" -- the loop over the file is kernel K6 (csrc/k6_metrics.cu) on the GPU;
  * :109-128  count_missing_snp_matrix_positions: the number of '-' in one sample's row of the SNP matrix.
The rest of collect_metrics (samtools / VCF / fastq statistics) is not on the hot path.
"""
from __future__ import annotations

import numpy as np

from . import device


def pileup_depth_sum(pileup_file_path):
    """collect_metrics.py:322-329: depth_sum += int(line.split()[3]) for every line (ValueError / IndexError ignored)."""
    text = np.fromfile(pileup_file_path, dtype=np.uint8)
    return device.context().pileup_depth_sum(text)[0]


def mean_pileup_depth(pileup_file_path, reference_length):
    """collect_metrics.py:331-338: the metric's text, or "" where the reference reports "Cannot calculate mean pileup
    depth." (no depth or no reference)."""
    depth_sum = pileup_depth_sum(pileup_file_path)
    if depth_sum > 0 and reference_length > 0:
        return "%.2f" % (float(depth_sum) / float(reference_length))
    return ""


def reference_length(reference_file_path):
    """collect_metrics.py:330-332: the summed length of the fasta records."""
    total = 0
    with open(reference_file_path) as f:
        for line in f:
            if not line.startswith(">"):
                total += len(line.strip())
    return total


def count_missing_snp_matrix_positions(file_path, sample_id):
    """collect_metrics.py:109-128: '-' characters in the record whose id is sample_id (0 when there is none)."""
    cur, n, found = None, 0, False
    with open(file_path) as f:
        for line in f:
            if line.startswith(">"):
                if found:
                    return n
                title = line[1:].split()
                cur = title[0] if title else ""
                found = cur == sample_id
            elif found:
                n += line.count("-")
    return n if found else 0
