#!/usr/bin/env python
"""Regenerate tests/golden/* from the reference mounted at /root/reference (build container only).

Produces
  lambda.tar.xz / agona.tar.xz / listeria.tar.xz   the reference's own bundled inputs + expected outputs for
                                                   the hot path (snppipeline/data/*ExpectedResults), repacked
  ref_lines.json.xz      per-line known answers: seeded lines (tests/linegen.py) x parameter sets, each with the
                         output of the reference's pileup.Record + ConsensusCaller (or the exception it raises)
  ref_files.json.xz      whole-file known answers: small synthetic pileups pushed through the reference's
                         `cfsan_snp_pipeline call_consensus` (filter mode, --vcfAllPos mode, with -e exclude
                         files, multi-contig, duplicate positions, CRLF) -> consensus.fasta text or exit code
  references.tar.xz      the three datasets' reference genomes (snp_reference's input)
  doctest_strip.json     the strip doctests of pileup.py:294-309 evaluated by the reference itself

Usage:  python tests/golden/make_golden.py      (needs /root/reference; writes next to this script)
"""
from __future__ import annotations

import io
import json
import lzma
import os
import random
import sys
import tarfile
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness as rh  # noqa: E402
import linegen  # noqa: E402

DATA = os.path.join(rh.REFERENCE_ROOT, "snppipeline", "data")

PARAM_SETS = [
    # (min_base_qual, min_freq, min_depth, min_strand_depth, min_strand_bias)
    (0, 0.6, 1, 0, 0.0),
    (0, 0.6, 3, 0, 0.0),
    (15, 0.6, 3, 1, 0.1),
    (0, 0.55, 2, 2, 0.25),
    (20, 0.9, 10, 0, 0.5),
    (0, 1.0, 0, 0, 0.0),
    (-5, 0.51, 1, 3, 0.33),
]


def pack(name, root, keep):
    path = os.path.join(HERE, name + ".tar.xz")
    with tarfile.open(path, "w:xz", preset=9) as tar:
        for dirpath, _, files in sorted(os.walk(root)):
            for f in sorted(files):
                full = os.path.join(dirpath, f)
                rel = os.path.relpath(full, root)
                if keep(rel):
                    ti = tar.gettarinfo(full, arcname=os.path.join(name, rel))
                    ti.mtime, ti.uid, ti.gid, ti.uname, ti.gname, ti.mode = 0, 0, 0, "", "", 0o644
                    with open(full, "rb") as fh:
                        tar.addfile(ti, fh)
    print("wrote", path, os.path.getsize(path))


def dump_xz(name, obj):
    path = os.path.join(HERE, name)
    with lzma.open(path, "wt", preset=9) as f:
        json.dump(obj, f, separators=(",", ":"), sort_keys=True)
    print("wrote", path, os.path.getsize(path))


def make_datasets():
    hot = ("reads.all.pileup", "var.flt.vcf", "var.flt_preserved.vcf", "var.flt_removed.vcf", "consensus.fasta",
           "consensus_preserved.fasta", "consensus.vcf", "consensus_preserved.vcf")
    top = ("snplist", "snpma", "snp_distance", "referenceSNP")
    keep = lambda rel: (os.path.basename(rel) in hot) or (os.sep not in rel and rel.startswith(top)
                                                          and not rel.endswith(".vcf"))
    pack("lambda", os.path.join(DATA, "lambdaVirusExpectedResults"), keep)
    pack("agona", os.path.join(DATA, "agonaExpectedResults"), keep)
    pack("listeria", os.path.join(DATA, "listeriaExpectedResults"), keep)


def make_references():
    """The reference genomes of the three bundled datasets (inputs of snp_reference), one archive."""
    path = os.path.join(HERE, "references.tar.xz")
    picks = [("lambda", "lambdaVirusInputs/reference/lambda_virus.fasta"), ("agona", "agonaInputs/reference/NC_011149.fasta"),
             ("listeria", "listeriaInputs/reference/CFSAN023463.HGAP.draft.fasta")]
    with tarfile.open(path, "w:xz", preset=9) as tar:
        for name, rel in picks:
            full = os.path.join(DATA, rel)
            ti = tar.gettarinfo(full, arcname=os.path.join("references", name + ".fasta"))
            ti.mtime, ti.uid, ti.gid, ti.uname, ti.gname, ti.mode = 0, 0, 0, "", "", 0o644
            with open(full, "rb") as fh:
                tar.addfile(ti, fh)
    print("wrote", path, os.path.getsize(path))


def make_ref_lines():
    cases = []
    for seed in range(6):
        rng = random.Random(1000 + seed)
        lines = []
        for i in range(260):
            x = rng.random()
            if x < 0.35:
                lines.append(linegen.realistic_line(rng, 100 + i))
            elif x < 0.93:
                lines.append(linegen.nasty_line(rng, 100 + i))
            else:
                lines.append(linegen.broken_line(rng, 100 + i))
        for line in lines:
            for ps in (PARAM_SETS[seed % len(PARAM_SETS)], PARAM_SETS[(seed + 3) % len(PARAM_SETS)]):
                cases.append({"line": line, "params": list(ps), "ref": rh.record_report(line, *ps)})
    dump_xz("ref_lines.json.xz", cases)


def _run_call_consensus(workdir, pileup_text, snplist_text, exclude_text, extra):
    sdir = os.path.join(workdir, "samples", "sampleX")
    os.makedirs(sdir, exist_ok=True)
    with open(os.path.join(sdir, "reads.all.pileup"), "w", newline="") as f:
        f.write(pileup_text)
    with open(os.path.join(workdir, "snplist.txt"), "w") as f:
        f.write(snplist_text)
    cmd = "call_consensus -f -v 0 -l %s/snplist.txt -o %s/consensus.fasta %s" % (workdir, sdir, extra)
    if exclude_text is not None:
        with open(os.path.join(workdir, "exclude.vcf"), "w") as f:
            f.write(exclude_text)
        cmd += " -e %s/exclude.vcf" % workdir
    cmd += " %s/reads.all.pileup" % sdir
    out = os.path.join(sdir, "consensus.fasta")
    if os.path.exists(out):
        os.remove(out)
    os.environ.pop("errorOutputFile", None)
    stderr = sys.stderr
    sys.stderr = io.StringIO()
    try:
        rh.run_command(cmd)
        code = 0
    except SystemExit as e:
        code = e.code
    except Exception as e:      # noqa: BLE001  (uncaught -> excepthook -> exit 100 in the real CLI)
        code = "raises:" + type(e).__name__
    finally:
        sys.stderr = stderr
    fasta = open(out).read() if os.path.exists(out) else None
    return {"exit": code, "fasta": fasta}


VCF_HEAD = "##fileformat=VCFv4.1\n##source=VarScan2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSample1\n"


def vcf_text(sites):
    return VCF_HEAD + "".join("%s\t%d\t.\tA\tG\t.\tPASS\tADP=1\tGT\t1/1\n" % s for s in sites)


def make_ref_files():
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        for seed in range(10):
            rng = random.Random(7000 + seed)
            chroms = ["chrB|x", "chrA", "chr10"] if seed % 3 == 0 else [linegen.CHROM]
            text_parts, all_pos = [], []
            for c in chroms:
                n = rng.randint(150, 400)
                sites = {p: rng.choice("ACGT") for p in rng.sample(range(1, n + 1), 25)}
                t = linegen.pileup_text(seed * 31 + len(c), n, chrom=c, nasty=0.25 if seed % 2 else 0.0, gaps=0.05,
                                        sites=sites)
                text_parts.append(t)
                all_pos += [(c, p) for p in range(1, n + 60)]
            text = "".join(text_parts)
            if seed == 4:                        # duplicate positions: last one wins
                lines = text.splitlines(True)
                text = "".join(lines + lines[40:60][::-1])
            if seed in (5, 8):
                text = text.replace("\n", "\r\n")
            if seed == 6 and text.endswith("\n"):
                text = text[:-1]                 # no trailing newline
            snps = sorted(rng.sample(all_pos, 60), key=lambda s: (s[0], s[1]))
            if seed == 7:
                snps = snps + snps[:5]           # duplicates in the snplist are emitted twice
                rng.shuffle(snps)
            snplist = "".join("%s\t%d\t1\tsampleX\n" % s for s in snps)
            excl = vcf_text(rng.sample(all_pos, 30) + snps[:6]) if seed % 2 == 0 else None
            ps = PARAM_SETS[seed % len(PARAM_SETS)]
            extra = "-q %d -c %s -D %d -d %d -b %s" % ps
            for mode in ("", "--vcfAllPos"):
                res = _run_call_consensus(tmp, text, snplist, excl, extra + " " + mode)
                cases.append({"pileup": text, "snplist": snplist, "exclude": excl, "params": list(ps),
                              "all_pos": bool(mode), "ref": res})
        # error cases: a broken line at / away from a site, both modes
        rng = random.Random(99)
        base = linegen.pileup_text(5, 80)
        snplist = "".join("%s\t%d\t1\tsampleX\n" % (linegen.CHROM, p) for p in (5, 17, 40, 41, 77))
        class _Fixed(object):
            def __init__(self, k):
                self.k = k

            def randrange(self, n):
                return self.k

        for k in range(9):
            for at in (17, 30):
                lines = base.splitlines(True)
                bl = linegen.broken_line(_Fixed(k), at)
                lines[at - 1] = bl
                text = "".join(lines)
                for mode in ("", "--vcfAllPos"):
                    res = _run_call_consensus(tmp, text, snplist, None, "-D 3 " + mode)
                    cases.append({"pileup": text, "snplist": snplist, "exclude": None,
                                  "params": [0, 0.6, 3, 0, 0.0], "all_pos": bool(mode), "ref": res})
        # empty snplist is not an error (regression_tests.sh:3156-3211)
        res = _run_call_consensus(tmp, base, "", None, "")
        cases.append({"pileup": base, "snplist": "", "exclude": None, "params": [0, 0.6, 1, 0, 0.0],
                      "all_pos": False, "ref": res})
    dump_xz("ref_files.json.xz", cases)


def make_doctest_strip():
    pileup = rh.ref("pileup")
    ins = [".,.actg,,,", "^K.,.^Fa,,,^K", "$.,.$*$*,,,*", ".,.+10AAAAAAAAAAa,,,", "+2TT.,.+10AAAAAAAAAAa,,,+2GC",
           ".,.-10AAAAAAAAAAa,,,", "-2TT.,.-10AAAAAAAAAAa,,,-2GC", "^Kc-2TT..$a+10AAAAAAAAAAa,,*,-2GC",
           "^+.,", "^^.,", "^-5,,", ".+", ".-1n,", "^", "+", "+2A", "+2+1AC..", "-1^A..", "$^$^", "+0A", "^+2AA.",
           ".+3AC", "+12ACGTACGTACGT.", "..+1-1A.,", "+1+1+1AAA", "-2^^A."]
    rng = random.Random(5)
    for _ in range(400):
        n = rng.randint(0, 24)
        ins.append("".join(rng.choice("..,,ACGTacgt*^$+-0123459") for _ in range(n)))
    out = [{"in": s, "out": pileup.Record._strip_unwanted_base_patterns(s)} for s in ins]
    with open(os.path.join(HERE, "doctest_strip.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote doctest_strip.json", len(out))


if __name__ == "__main__":
    if not rh.available():
        sys.exit("reference tree not mounted at %s" % rh.REFERENCE_ROOT)
    make_datasets()
    make_references()
    make_doctest_strip()
    make_ref_lines()
    make_ref_files()
