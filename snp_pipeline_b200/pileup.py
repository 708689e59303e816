"""Host-side mirror of snppipeline/pileup.py for the GPU path.

The reference's Reader yields one Record per pileup line and ConsensusCaller.call_consensus judges it
(pileup.py:408-429, 209-325, 492-590).  Here the same two objects describe WHAT to compute -- the file, the
minimum base quality, the position filter, the caller's thresholds -- and Reader.call_consensus hands the whole
file to kernel K1 through the C ABI (snpgpu_pileup_consensus).  No line is parsed on the host.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from . import device


class ConsensusCaller(object):
    """Same constructor and filter vocabulary as pileup.ConsensusCaller (pileup.py:433-490)."""

    def __init__(self, min_cons_freq, min_cons_depth, min_cons_strand_depth, min_cons_strand_bias):
        self.min_cons_freq = min_cons_freq
        self.min_cons_depth = min_cons_depth
        self.min_cons_strand_depth = min_cons_strand_depth
        self.min_cons_strand_bias = min_cons_strand_bias
        self.failed_raw_depth_filter = "RawDpth"
        self.failed_freq_filter = "VarFreq" + str(int(100 * min_cons_freq))
        self.failed_depth_filter = "Depth" + str(min_cons_depth)
        self.failed_strand_depth_filter = "StrDpth" + str(min_cons_strand_depth)
        self.failed_strand_bias_filter = "StrBias" + str(int(100 * min_cons_strand_bias))

    def get_filter_descriptions(self):
        """[(filter id, description)] as pileup.py:473-490."""
        return [
            (self.failed_raw_depth_filter, "No read depth"),
            (self.failed_freq_filter, "Variant base frequency below %.2f" % self.min_cons_freq),
            (self.failed_depth_filter, "Less than %i supporting reads" % self.min_cons_depth),
            (self.failed_strand_depth_filter, "Less than %i variant-supporing reads on at least one strand"
             % self.min_cons_strand_depth),
            (self.failed_strand_bias_filter,
             "Fraction of variant supporting reads below %.2f on one strand" % self.min_cons_strand_bias),
        ]

    def fail_names(self, mask):
        """Fail-mask bits of the kernel -> the reference's list of filter names (or None)."""
        names = []
        for bit, name in ((_lib.FAIL_RAWDPTH, self.failed_raw_depth_filter), (_lib.FAIL_VARFREQ, self.failed_freq_filter),
                          (_lib.FAIL_DEPTH, self.failed_depth_filter),
                          (_lib.FAIL_STRDPTH, self.failed_strand_depth_filter),
                          (_lib.FAIL_STRBIAS, self.failed_strand_bias_filter), (_lib.FAIL_REGION, "Region")):
            if mask & bit:
                names.append(name)
        return names or None

    def params(self, min_base_quality):
        return _lib.make_params(min_base_quality, self.min_cons_freq, self.min_cons_depth,
                                self.min_cons_strand_depth, self.min_cons_strand_bias)


class Reader(object):
    """pileup.Reader's constructor (pileup.py:389-407); iteration is replaced by one kernel call."""

    def __init__(self, file_path, min_base_quality, chrom_position_set=None):
        self.file_path = file_path
        self.min_base_quality = min_base_quality
        self.chrom_position_set = chrom_position_set
        f = open(file_path)          # open and close the file to make sure it works (pileup.py:404-406)
        f.close()

    def read_text(self, ctx):
        """The file's bytes in page-locked memory (so the H2D copy inside the C call runs at PCIe speed)."""
        import os
        size = os.path.getsize(self.file_path)
        arr, owner = ctx.pinned_array(size)
        with open(self.file_path, "rb", buffering=0) as f:
            got, view = 0, memoryview(arr)
            while got < size:
                n = f.readinto(view[got:])
                if not n:
                    break
                got += n
        return arr[:got], owner

    def call_consensus(self, caller, snp_list, excluded_positions=(), vcf_writer=None, failed_snp_gt='.'):
        """The loop of call_consensus.py:161-188 for this file: returns (consensus string in snp_list order,
        stats).  chrom_position_set None means every line is parsed (--vcfAllPos), otherwise only lines at
        snp_list / excluded positions, exactly like pileup.py:419-429.
        vcf_writer: an open vcf_writer.SingleSampleWriter (header written) that receives one record per parsed line.
        Raises ValueError / IndexError where the reference does (malformed line at a parsed position)."""
        ctx = device.context()
        sites = ctx.sites(snp_list, excluded_positions)
        text, owner = self.read_text(ctx)
        try:
            mode = _lib.MODE_ALL if self.chrom_position_set is None else _lib.MODE_SITES
            params = caller.params(self.min_base_quality)
            try:
                ctx.want_vcf_records(vcf_writer is not None)     # (K1 then lists the lines it parses on the way)
                row, stats = ctx.pileup_consensus(text, sites, params, mode)[:2]
                if vcf_writer is not None:      # one VCF record per line the reader yields (call_consensus.py:161-184)
                    filter_texts = [";".join(caller.fail_names(m) or ["PASS"]) for m in range(_lib.VCF_FILTER_MASKS)]
                    if all(len(t) < _lib.VCF_FILTER_TEXT for t in filter_texts):
                        data_lines, _ = ctx.pileup_vcf_text(sites, params, mode, filter_texts, failed_snp_gt,
                                                            vcf_writer.preserve_ref_case)
                        vcf_writer.write_text(data_lines)
                    else:                               # (filter names longer than the kernel's table: the numbers, formatted here)
                        records, alts = ctx.pileup_vcf_records(sites, params, mode)
                        vcf_writer.write_records(text, records, alts, caller, failed_snp_gt)
            except _lib.SnpGpuError as e:
                raise translate_error(e, self.file_path)
            finally:
                ctx.want_vcf_records(False)
        finally:
            owner.free()
            sites.close()
        return row.decode("ascii"), stats


def translate_error(e, path):
    """libsnpgpu "the reference raises here" codes -> the exception type the reference raises."""
    where = " (pileup file %s, line at byte offset %s)" % (path, e.offset)
    if e.code == _lib.E_VALUE:
        return ValueError("invalid literal for int() with base 10" + where)
    if e.code == _lib.E_UNPACK:
        return ValueError("not enough values to unpack (expected 2)" + where)
    if e.code == _lib.E_INDEX:
        return IndexError("list index out of range" + where)
    if e.code == _lib.E_DOMAIN:
        return UnicodeDecodeError("utf-8", b"", 0, 1, "byte outside ASCII" + where)
    return e
