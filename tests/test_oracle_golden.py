"""Pin the oracle (oracle/snp_oracle.c + oracle/oracle.py) against the reference's own known answers.

Sources of truth, none of them ours:
  * doctest values of pileup.py:98-184, 294-309, 513-548
  * golden vectors produced by the reference's Python (tests/golden/make_golden.py)
  * the bundled lambda-virus / Agona / Listeria expected results
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

LAMBDA_SAMPLES = ["sample1", "sample2", "sample3", "sample4"]


def _read(path, mode="r"):
    with open(path, mode) as f:
        return f.read()


# ------------------------------------------------------------------ doctests of pileup.py
def test_doctest_record_tally():
    # pileup.py:98-184 (Record doctests): 'ID 42 G 14 aaaaAAAA...,,, 00001111222333' at min quality 15
    p = orc.make_params(min_base_qual=15, min_cons_freq=0.5, min_cons_depth=0)
    r = orc.line_report(b"ID 42 G 14 aaaaAAAA...,,, 00001111222333", p)
    assert r["status"] == 0 and r["pos"] == 42 and r["raw_depth"] == 14 and r["ref"] == "G"
    assert r["good_depth"] == 14 and r["fwd_good_depth"] == 7 and r["rev_good_depth"] == 7
    assert r["total"] == {"A": 8, "G": 6} and r["fwd"] == {"A": 4, "G": 3} and r["rev"] == {"A": 4, "G": 3}
    assert r["most_common"] == ["A", "G"]
    # tie-break: alphabetical among equal counts (pileup.py:178-184)
    r = orc.line_report(b"ID 42 G 14 aaaAAA....,,,, 00011122223333", orc.make_params(15, 0.5, 0))
    assert r["most_common"] == ["G", "A"]


@pytest.mark.parametrize("line,params,expect", [
    # pileup.py:513-548
    (b"ID 42 G 14 aaaaAAAA...,,, 00001111222333", (15, 0.5, 0, 0, 0.0), ("A", None)),
    (b"ID 42 G 14 aaaaAAAA...,,, 00001111222333", (15, 0.6, 0, 0, 0.0), ("A", ["VarFreq60"])),
    (b"ID 42 G 14 aaaaAAAA...,,, 00001111222333", (15, 0.0, 8, 4, 0.0), ("A", None)),
    (b"ID 42 G 14 aaaaAAAA...,,, 00001111222333", (15, 0.0, 9, 4, 0.0), ("A", ["Depth9"])),
    (b"ID 42 G 14 aaaaAAAA...,,, 00001111222333", (15, 0.0, 0, 5, 0.0), ("A", ["StrDpth5"])),
    (b"ID 42 G 14 aAAAAAAA...,,, 00001111222333", (15, 0.0, 0, 0, 0.2), ("A", ["StrBias20"])),
    (b"ID 42 G 14 aaaAAAAA...,,, 00001111222333", (15, 0.0, 9, 4, 0.4), ("A", ["Depth9", "StrDpth4", "StrBias40"])),
    (b"ID 42 G 14 aaaAAA....,,,, 00011122223333", (15, 0.0, 0, 0, 0.0), ("G", None)),
    (b"ID 42 g 14 aaaAAA....,,,, 00011122223333", (15, 0.0, 0, 0, 0.0), ("g", None)),
    (b"ID 42 g 0", (15, 0.0, 0, 5, 0.0), ("-", ["RawDpth"])),
])
def test_doctest_caller(line, params, expect):
    p = orc.make_params(*params)
    r = orc.line_report(line, p)
    assert r["status"] == 0
    assert (r["cons"], orc.fail_names(r["fail"], p)) == expect


def test_strip_known_answers():
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "doctest_strip.json")))
    assert len(cases) > 400
    for c in cases:
        assert orc.strip_bases(c["in"].encode()).decode() == c["out"], c["in"]


def test_fp64_threshold_edge():
    # SURVEY 7 "fp64 threshold compare": 100 * 0.55 = 55.00000000000001 in IEEE double -> 55 of 100 fails
    line = ("c 1 A 100 " + "G" * 55 + "." * 45 + " " + "I" * 100).encode()
    r = orc.line_report(line, orc.make_params(0, 0.55, 1))
    assert r["cons"] == "G" and r["fail"] == orc.FAIL_VARFREQ


# ------------------------------------------------------------------ vectors from the reference's Python
_EXC = {orc.E_VALUE: "ValueError", orc.E_INDEX: "IndexError", orc.E_UNPACK: "ValueError"}


def test_lines_vs_reference(ref_lines):
    n_ok = n_raise = 0
    for case in ref_lines:
        p = orc.make_params(*case["params"])
        r = orc.line_report(case["line"].encode(), p)
        ref = case["ref"]
        if "raises" in ref:
            assert r["status"] in _EXC and _EXC[r["status"]] == ref["raises"], case["line"]
            n_raise += 1
            continue
        assert r["status"] == 0, case["line"]
        for k in ("pos", "raw_depth", "ref", "good_depth", "fwd_good_depth", "rev_good_depth", "most_common",
                  "total", "fwd", "rev", "cons"):
            assert r[k] == ref[k], (k, case["line"], r[k], ref[k])
        assert orc.fail_names(r["fail"], p) == ref["fails"], case["line"]
        n_ok += 1
    assert n_ok > 2500 and n_raise > 150


def _parse_case(case):
    snps = [(ln.split()[0], int(ln.split()[1])) for ln in case["snplist"].splitlines()]
    excl = []
    if case["exclude"]:
        for ln in case["exclude"].splitlines():
            if not ln.startswith("#"):
                f = ln.split("\t")
                excl.append((f[0], int(f[1])))
    return snps, excl


def test_files_vs_reference(ref_files):
    n_ok = n_raise = 0
    for case in ref_files:
        snps, excl = _parse_case(case)
        p = orc.make_params(*case["params"])
        ref = case["ref"]
        try:
            row = orc.pileup_consensus(case["pileup"].encode(), snps, excl, p, parse_all=case["all_pos"])
        except orc.OracleError as e:
            assert str(ref["exit"]).startswith("raises:"), ref["exit"]
            assert _EXC[e.status] == ref["exit"].split(":")[1]
            n_raise += 1
            continue
        assert ref["exit"] == 0
        assert orc.fasta_text("sampleX", row.decode()) == ref["fasta"]
        n_ok += 1
    assert n_ok >= 20 and n_raise >= 20


# ------------------------------------------------------------------ bundled datasets
@pytest.mark.parametrize("branch", ["", "_preserved"])
def test_lambda_call_consensus(golden_dir, branch):
    root = os.path.join(golden_dir, "lambda")
    snps = orc.read_snp_list(os.path.join(root, "snplist%s.txt" % branch))
    p = orc.make_params(min_cons_depth=3)          # consensus.vcf header: VarFreq60, Depth3, StrDpth0, StrBias0
    for s in LAMBDA_SAMPLES:
        sdir = os.path.join(root, "samples", s)
        text = _read(os.path.join(sdir, "reads.all.pileup"), "rb")
        excl = orc.vcf_sites(os.path.join(sdir, "var.flt_removed.vcf")) if branch else []
        row = orc.pileup_consensus(text, snps, excl, p)
        assert orc.fasta_text(s, row.decode()) == _read(os.path.join(sdir, "consensus%s.fasta" % branch))
        row_all = orc.pileup_consensus(text, snps, excl, p, parse_all=True)
        assert row_all == row


@pytest.mark.parametrize("dataset,vcf,snplist", [
    ("lambda", "var.flt.vcf", "snplist.txt"), ("lambda", "var.flt_preserved.vcf", "snplist_preserved.txt"),
    ("agona", "var.flt.vcf", "snplist.txt"), ("listeria", "var.flt.vcf", "snplist.txt"),
])
def test_merge_sites_golden(golden_dir, dataset, vcf, snplist):
    root = os.path.join(golden_dir, dataset)
    dirs = [os.path.join(root, "samples", d) for d in os.listdir(os.path.join(root, "samples"))]
    text, filt = orc.merge_sites_text(dirs, vcf)
    assert text == _read(os.path.join(root, snplist))
    assert filt == "".join(d + "\n" for d in dirs)
    # the packed-key C restatement agrees with the dict restatement
    names = sorted(os.path.basename(d) for d in dirs)
    per = [orc.vcf_sites(os.path.join(root, "samples", n, vcf)) for n in names]
    chroms = sorted({c for s in per for c, _ in s})
    rank = {c: i for i, c in enumerate(chroms)}
    keys = np.array([(rank[c] << 32) | p for s in per for c, p in s], dtype=np.uint64)
    samp = np.array([i for i, s in enumerate(per) for _ in s], dtype=np.uint32)
    uniq, cnt, samples = orc.merge_sites_keys(keys, samp)
    lines, o = [], 0
    for k, c in zip(uniq, cnt):
        who = [names[i] for i in samples[o:o + c]]
        o += int(c)
        lines.append("%s\t%d\t%d\t%s\n" % (chroms[int(k) >> 32], int(k) & 0xffffffff, c, "\t".join(who)))
    assert "".join(lines) == text


def test_merge_sites_maxsnps(golden_dir):
    # regression_tests.sh:6287-6384: --maxsnps drops whole samples and the filtered list loses them
    root = os.path.join(golden_dir, "lambda")
    dirs = [os.path.join(root, "samples", s) for s in reversed(LAMBDA_SAMPLES)]
    counts = {d: len(orc.vcf_sites(os.path.join(d, "var.flt.vcf"))) for d in dirs}
    cut = sorted(counts.values())[1]
    text, filt = orc.merge_sites_text(dirs, max_snps=cut)
    kept = [d for d in dirs if counts[d] <= cut]
    assert filt == "".join(d + "\n" for d in kept) and 0 < len(kept) < 4
    for ln in text.splitlines():
        assert set(ln.split("\t")[3:]) <= {os.path.basename(d) for d in kept}


@pytest.mark.parametrize("dataset", ["lambda", "agona"])
def test_snp_matrix_golden(golden_dir, dataset):
    root = os.path.join(golden_dir, dataset)
    names = sorted(os.listdir(os.path.join(root, "samples")))
    cat = "".join(_read(os.path.join(root, "samples", n, "consensus.fasta")) for n in names)
    assert cat == _read(os.path.join(root, "snpma.fasta"))


@pytest.mark.parametrize("dataset,suffix", [("lambda", ""), ("lambda", "_preserved"), ("agona", ""),
                                            ("listeria", ""), ("listeria", "_preserved")])
def test_distance_golden(golden_dir, dataset, suffix):
    root = os.path.join(golden_dir, dataset)
    seqs = orc.read_fasta_matrix(os.path.join(root, "snpma%s.fasta" % suffix))
    pair, mat = orc.distance_texts(seqs)
    assert mat == _read(os.path.join(root, "snp_distance_matrix%s.tsv" % suffix))
    pw = os.path.join(root, "snp_distance_pairwise%s.tsv" % suffix)
    if os.path.exists(pw):
        assert pair == _read(pw)


def test_fasta_writer_edges():
    assert orc.fasta_text("s", "") == ">s\n"                      # regression_tests.sh:5920-5934
    assert orc.fasta_text("s", "A" * 60) == ">s\n" + "A" * 60 + "\n"
    assert orc.fasta_text("s", "A" * 61) == ">s\n" + "A" * 60 + "\nA\n"


@pytest.mark.parametrize("dataset", ["lambda", "agona", "listeria"])
@pytest.mark.parametrize("suffix", ["", "_preserved"])
def test_reference_snp_golden(golden_dir, dataset, suffix):
    """snp_reference (scope row f2): the oracle's restatement of utils.write_reference_snp_file against the bundled
    referenceSNP*.fasta of all three datasets."""
    root = os.path.join(golden_dir, dataset)
    ref = os.path.join(golden_dir, "references", dataset + ".fasta")
    got = orc.reference_snp_text(ref, os.path.join(root, "snplist%s.txt" % suffix))
    assert got == open(os.path.join(root, "referenceSNP%s.fasta" % suffix)).read()
