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
  ref_regions.json.xz    filter_regions known answers: seeded samples of (contig, position) records pushed through the
                         reference's collect_dense_regions + utils.merge_regions + utils.in_region in the order
                         filter_regions_across_samples / filter_regions_per_sample call them -> removed flag per record

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


def make_ref_regions():
    fr = rh.ref("filter_regions")
    utils = rh.ref("utils")

    class Rec(object):
        def __init__(self, chrom, pos):
            self.CHROM, self.POS = chrom, pos

    def flags_of(samples, contig_len, edge, windows, maxs, mode, outgroup):
        out = []
        if mode == "all":                                   # filter_regions.py:256-303
            bad = dict()
            for sid, recs in samples:
                if sid not in outgroup:
                    fr.collect_dense_regions([Rec(c, p) for c, p in recs], bad, contig_len, edge, maxs, windows)
            for contig, regions in bad.items():
                bad[contig] = utils.merge_regions(regions)
            for sid, recs in samples:
                out.append(None if sid in outgroup else [bool(utils.in_region(p, bad[c])) for c, p in recs])
        else:                                               # filter_regions.py:352-383
            for sid, recs in samples:
                if sid in outgroup:
                    out.append(None)
                    continue
                bad = dict()
                fr.collect_dense_regions([Rec(c, p) for c, p in recs], bad, contig_len, edge, maxs, windows)
                for contig, regions in bad.items():
                    bad[contig] = utils.merge_regions(regions)
                out.append([bool(utils.in_region(p, bad[c])) for c, p in recs])
        return out

    rng = random.Random(77)
    cases = []
    for k in range(120):
        n_contigs = rng.choice([1, 1, 2, 5])
        contigs = ["ctg%d" % i for i in range(n_contigs)]
        contig_len = {c: rng.choice([300, 900, 5000, 60000, 4000000]) for c in contigs}
        if k % 7 == 3:
            del contig_len[contigs[-1]]                     # a contig the reference fasta does not hold (sys.maxsize)
        edge = rng.choice([1, 50, 500, 500, 2000])
        n_par = rng.choice([1, 1, 2, 3])
        windows = [rng.choice([1, 2, 10, 100, 1000, 1000, 5000]) for _ in range(n_par)]
        maxs = [rng.choice([1, 2, 3, 3, 5, 10]) for _ in range(n_par)]
        samples = []
        for s in range(rng.choice([1, 2, 4, 6])):
            recs = []
            for c in contigs:
                top = min(contig_len.get(c, 100000), 200000)
                n = rng.choice([0, 1, 3, 10, 40, 150])
                if rng.random() < 0.5:                      # clustered
                    centres = [rng.randint(1, top) for _ in range(max(1, n // 8))]
                    pos = [max(1, min(top, int(rng.gauss(rng.choice(centres), rng.choice([3, 60, 400]))))) for _ in range(n)]
                else:
                    pos = [rng.randint(1, top) for _ in range(n)]
                if rng.random() < 0.7:
                    pos = sorted(pos)
                if rng.random() < 0.5:
                    pos = list(dict.fromkeys(pos))
                recs += [(c, p) for p in pos]
            samples.append(("sample%d" % s, recs))
        outgroup = ["sample1"] if k % 5 == 4 else []
        for mode in ("all", "each"):
            cases.append({"samples": samples, "contig_len": contig_len, "edge": edge, "windows": windows, "maxs": maxs,
                          "mode": mode, "outgroup": outgroup,
                          "removed": flags_of(samples, contig_len, edge, windows, maxs, mode, outgroup)})
    doc = []                                                # the doctests of filter_regions.py:38-58 / utils.py:1185-1262
    for maxs_, win, snps in ((3, 1000, []), (3, 1000, [1, 2, 3]), (3, 1000, [1, 2, 3, 1000]), (3, 1000, [1, 2, 3, 999, 1000]),
                             (3, 1000, [1, 2, 3, 1000, 1001, 1002, 1500]), (3, 1000, [1, 2, 3, 1000, 3001, 3002, 3003, 4000]),
                             (1, 1, [5, 5, 6]), (2, 10, [1, 5, 10, 11, 30, 31, 39])):
        doc.append({"max": maxs_, "window": win, "snps": snps, "regions": fr.find_dense_regions(maxs_, win, snps)})
    dump_xz("ref_regions.json.xz", {"cases": cases, "dense": doc})


if __name__ == "__main__":
    if not rh.available():
        sys.exit("reference tree not mounted at %s" % rh.REFERENCE_ROOT)
    if sys.argv[1:] == ["regions"]:
        make_ref_regions()
        sys.exit(0)
    make_datasets()
    make_references()
    make_doctest_strip()
    make_ref_lines()
    make_ref_files()
    make_ref_regions()
