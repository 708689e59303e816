"""Per-sample consensus VCF writer.

Mirror of snppipeline/vcf_writer.py (SingleSampleWriter: write_header :85-135, _make_vcf_record_from_pileup :137-379,
write_from_pileup :381-433) for the GPU path.  The reference builds a PyVCF model record per pileup.Record and lets
PyVCF3 (~=1.0.3, setup.py:25; vcf/parser.py Writer) print it; here the numbers come from kernel K5
(snpgpu_pileup_vcf_records: raw depth, reference depths, ALT alleles with their depths, consensus, failed filters) and
the text PyVCF would print is produced directly.  The layout -- header order (fileformat, fileDate, source, reference,
INFO, FORMAT x 9, FILTER lines, tab-separated #CHROM line), '.' for missing ID / QUAL / ALT, "PASS" for an empty
filter list, comma-joined per-allele lists -- is pinned by the reference's bundled lambda consensus*.vcf files
(tests/test_vcf.py).
"""
from __future__ import annotations

import datetime

from . import _lib

__version__ = "2.2.1"          # of the reference pipeline whose file layout this follows (snppipeline/__init__.py:5)

VCF_VERSION = '##fileformat=VCFv4.2\n'
VCF_DATE = '##fileDate=%Y%m%d\n'
VCF_SOURCE = '##source=CFSAN SNP-Pipeline %s\n' % __version__
VCF_INFO = '##INFO=<ID=NS,Number=1,Type=Integer,Description="Number of samples with data">\n'
VCF_FILTER = '##FILTER=<ID=%s,Description="%s">\n'
VCF_FORMAT = '''##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">
##FORMAT=<ID=SDP,Number=1,Type=Integer,Description="Raw read depth">
##FORMAT=<ID=RD,Number=1,Type=Integer,Description="Depth of reference-supporting bases">
##FORMAT=<ID=AD,Number=A,Type=Integer,Description="Depth of variant-supporting bases (comma-separated depth per alt allele)">
##FORMAT=<ID=RDF,Number=1,Type=Integer,Description="Depth of reference-supporting bases on forward strand">
##FORMAT=<ID=RDR,Number=1,Type=Integer,Description="Depth of reference-supporting bases on reverse strand">
##FORMAT=<ID=ADF,Number=A,Type=Integer,Description="Depth of variant-supporting bases on forward strand (comma-separated depth per alt allele)">
##FORMAT=<ID=ADR,Number=A,Type=Integer,Description="Depth of variant-supporting bases on reverse strand (comma-separated depth per alt allele)">
##FORMAT=<ID=FT,Number=1,Type=String,Description="Genotype filters using the same codes as the FILTER data element">
'''
VCF_REFERENCE = '##reference=%s\n'
VCF_HDR_LINE = '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n'
FORMAT_STR = "GT:SDP:RD:AD:RDF:RDR:ADF:ADR:FT"


def header_text(sample_id, filters, reference, now=None):
    """What PyVCF's Writer prints for the template of vcf_writer.py:103-127: simple metadata lines first (in template
    order: fileformat, fileDate, source, reference), then INFO, FORMAT, FILTER lines, then the column line."""
    now = now or datetime.datetime.now()
    out = [VCF_VERSION, datetime.datetime.strftime(now, VCF_DATE), VCF_SOURCE, VCF_REFERENCE % reference, VCF_INFO,
           VCF_FORMAT, VCF_FILTER % ("PASS", "All filters passed")]
    for name, description in filters:
        out.append(VCF_FILTER % (name, description))
    out.append(VCF_HDR_LINE % sample_id)
    return "".join(out)


def record_text(chrom, rec, alts, fail_names, failed_snp_gt, preserve_ref_case):
    """One data line: vcf_writer.py:295-379 on the kernel's numbers.
    rec: one element of the VCF_RECORD_DTYPE array; alts: its slice of the VCF_ALT_DTYPE array; fail_names: the
    reference's list of failed filter names or None."""
    ref = chr(rec["ref"])
    if not preserve_ref_case:
        ref = ref.upper()
    has_depth = bool(rec["flags"] & _lib.VCF_HAS_DEPTH)
    if not has_depth:                                   # most_common_good_bases is None (vcf_writer.py:310-315)
        alt_txt, gt, ad, adf, adr = ".", ".", "0", "0", "0"
    else:
        if len(alts) == 0:                              # :319-323
            alt_txt, gt, ad, adf, adr = ".", "0", "0", "0", "0"
        else:                                           # :324-331
            gt = "0" if rec["flags"] & _lib.VCF_FIRST_IS_REF else "1"
            alt_txt = ",".join(chr(b) for b in alts["base"])
            ad = ",".join(str(int(v)) for v in alts["ad"])
            adf = ",".join(str(int(v)) for v in alts["adf"])
            adr = ",".join(str(int(v)) for v in alts["adr"])
        if fail_names:                                  # :333-339
            gt = "." if failed_snp_gt == "." else ("0" if failed_snp_gt == "0" else "1")
    ft = ";".join(fail_names) if fail_names else "PASS"
    sample = ":".join((gt, str(int(rec["raw_depth"])), str(int(rec["rd"])), ad, str(int(rec["rdf"])), str(int(rec["rdr"])),
                       adf, adr, ft))
    return "\t".join((chrom, str(int(rec["pos"])), ".", ref, alt_txt, ".", ft, "NS=1", FORMAT_STR, sample)) + "\n"


class SingleSampleWriter(object):
    """Same constructor and call sequence as vcf_writer.SingleSampleWriter (write_header, then records, then close);
    records arrive as the arrays Context.pileup_vcf_records returns instead of one pileup.Record at a time."""

    def __init__(self, file, preserve_ref_case=False):
        if isinstance(file, str):
            self.file_path = file
            self.file_handle = open(file, "w")
        else:
            self.file_path = file.name
            self.file_handle = file
        self.preserve_ref_case = preserve_ref_case

    def close(self):
        self.file_handle.close()

    def write_header(self, sample_id, filters, reference):
        self.file_handle.write(header_text(sample_id, filters, reference))

    def write_text(self, data_lines):
        """The records already formatted (Context.pileup_vcf_text: the kernel's K5 text pass)."""
        raw = getattr(self.file_handle, "buffer", None)
        if raw is not None:                             # a text file: the bytes go straight to its binary layer
            self.file_handle.flush()
            raw.write(memoryview(data_lines))
        else:                                           # (a StringIO in the tests)
            self.file_handle.write(bytes(data_lines).decode("ascii", "surrogateescape"))

    def write_records(self, text, records, alts, caller, failed_snp_gt):
        """text: the pileup file's bytes (the chromosome column is copied from it); caller: pileup.ConsensusCaller
        (turns fail masks into the reference's filter names)."""
        mv = memoryview(text)
        w = self.file_handle.write
        names_of = {}
        for rec in records:
            o = int(rec["offset"]) + int(rec["chrom_off"])
            chrom = bytes(mv[o:o + int(rec["chrom_len"])]).decode("ascii")
            mask = int(rec["fail"])
            if mask not in names_of:
                names_of[mask] = caller.fail_names(mask)
            a0 = int(rec["alt_index"])
            w(record_text(chrom, rec, alts[a0:a0 + int(rec["n_alt"])], names_of[mask], failed_snp_gt,
                          self.preserve_ref_case))
