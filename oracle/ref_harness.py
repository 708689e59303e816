"""Run the UNMODIFIED reference Python (from /root/reference) in the build container.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference is mounted (never on the GPU box).  It is
how the oracle is pinned against the reference itself: tests/golden/make_golden.py uses it to produce the
committed golden vectors and tests/test_oracle_vs_reference.py compares live when the mount is present.

The reference imports three third-party packages that are not installed here (Biopython, PyVCF3,
jobrunner).  Minimal stand-ins for exactly the calls the hot path makes are registered in sys.modules
before the import; nothing under /root/reference is modified or copied.
"""
from __future__ import annotations

import importlib
import locale
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("SNP_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "snppipeline", "pileup.py"))


def _install_standins():
    if "Bio" in sys.modules and getattr(sys.modules["Bio"], "_standin", False):
        return
    # ---- Bio ---------------------------------------------------------------------------------
    bio = types.ModuleType("Bio"); bio._standin = True; bio.__path__ = []
    seq_m = types.ModuleType("Bio.Seq")
    rec_m = types.ModuleType("Bio.SeqRecord")
    io_m = types.ModuleType("Bio.SeqIO")

    class Seq(str):
        pass

    class SeqRecord(object):
        def __init__(self, seq, id="<unknown id>", name="<unknown name>", description="<unknown description>"):
            self.seq, self.id, self.name, self.description = seq, id, name, description

        def __getitem__(self, i):
            return self.seq[i]

    def write(records, handle, fmt):
        assert fmt == "fasta"
        n = 0
        for r in records:
            desc = r.description
            if desc and desc.split(None, 1)[0] == r.id:
                title = desc
            elif desc:
                title = "%s %s" % (r.id, desc)
            else:
                title = r.id
            handle.write(">%s\n" % title)
            s = str(r.seq)
            for i in range(0, len(s), 60):
                handle.write(s[i:i + 60] + "\n")
            n += 1
        return n

    def parse(path_or_handle, fmt):
        assert fmt == "fasta"
        h = open(path_or_handle) if isinstance(path_or_handle, str) else path_or_handle
        rid, desc, chunks = None, "", []
        for line in h:
            if line.startswith(">"):
                if rid is not None:
                    yield SeqRecord(Seq("".join(chunks)), id=rid, name=rid, description=desc)
                desc = line[1:].rstrip()
                rid = desc.split(None, 1)[0] if desc.split() else ""
                chunks = []
            else:
                chunks.append(line.strip())
        if rid is not None:
            yield SeqRecord(Seq("".join(chunks)), id=rid, name=rid, description=desc)

    def to_dict(records):
        return {r.id: r for r in records}

    seq_m.Seq = Seq
    rec_m.SeqRecord = SeqRecord
    io_m.write, io_m.parse, io_m.to_dict = write, parse, to_dict
    bio.Seq, bio.SeqRecord, bio.SeqIO = seq_m, rec_m, io_m
    sys.modules.update({"Bio": bio, "Bio.Seq": seq_m, "Bio.SeqRecord": rec_m, "Bio.SeqIO": io_m})

    # ---- vcf (PyVCF3 Reader: only CHROM / POS are consumed by the hot path) ------------------
    vcf_m = types.ModuleType("vcf"); vcf_m.__path__ = []
    row_pattern = re.compile("\t| +")

    class _Rec(object):
        def __init__(self, chrom, pos):
            self.CHROM, self.POS = chrom, pos

    class Reader(object):
        def __init__(self, fsock=None, filename=None, **kw):
            self._reader = fsock if fsock is not None else open(filename)
            self.reader = (line.strip() for line in self._reader if line.strip())
            line = next(self.reader)
            while line.startswith("##"):
                line = next(self.reader)

        def __iter__(self):
            return self

        def __next__(self):
            line = next(self.reader)
            row = row_pattern.split(line.rstrip())
            row[7]
            return _Rec(row[0], int(row[1]))

    vcf_m.Reader = Reader
    model = types.ModuleType("vcf.model")
    parser = types.ModuleType("vcf.parser")
    vcf_m.model, vcf_m.parser = model, parser
    sys.modules.update({"vcf": vcf_m, "vcf.model": model, "vcf.parser": parser})

    # ---- jobrunner (orchestration only; never called on the hot path) -------------------------
    jr = types.ModuleType("jobrunner")
    jr.JobRunner = type("JobRunner", (), {})
    jr.JobRunnerException = type("JobRunnerException", (Exception,), {})
    sys.modules["jobrunner"] = jr

    if not hasattr(locale, "format"):            # removed in py3.12; utils.py:120 calls it unconditionally
        locale.format = locale.format_string


_mods = {}


def ref(name):
    """Import ``snppipeline.<name>`` from the reference tree (stand-ins installed first)."""
    if name in _mods:
        return _mods[name]
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_ROOT)
    _install_standins()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    m = importlib.import_module("snppipeline." + name)
    _mods[name] = m
    return m


def run_command(line, argv0="cfsan_snp_pipeline"):
    """cfsan_snp_pipeline.run_command_from_line(line) with sys.argv set the way the console script sets it."""
    cli = ref("cfsan_snp_pipeline")
    old = sys.argv
    sys.argv = [argv0] + line.split()
    try:
        return cli.run_command_from_line(line)
    finally:
        sys.argv = old
        sys.excepthook = sys.__excepthook__


def record_report(line: str, min_base_qual, min_freq, min_depth, min_strand_depth, min_strand_bias):
    """pileup.Record + ConsensusCaller on one line -> dict shaped like oracle.line_report (or {'raises': name})."""
    pileup = ref("pileup")
    try:
        r = pileup.Record(line, min_base_qual)
        caller = pileup.ConsensusCaller(min_freq, min_depth, min_strand_depth, min_strand_bias)
        base, fails = caller.call_consensus(r)
    except Exception as e:                      # noqa: BLE001 - the class name is the datum
        return {"raises": type(e).__name__}
    return {
        "pos": r.position, "raw_depth": r.raw_depth, "ref": r.reference_base, "good_depth": r.good_depth,
        "fwd_good_depth": r.forward_good_depth, "rev_good_depth": r.reverse_good_depth,
        "most_common": r.most_common_good_bases, "total": dict(r.base_good_depth),
        "fwd": dict(r.forward_base_good_depth), "rev": dict(r.reverse_base_good_depth),
        "cons": base, "fails": fails,
    }
