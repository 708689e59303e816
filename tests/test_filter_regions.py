"""filter_regions: oracle against the reference's answers, host layer, and K7 on the GPU.

Modelled on the reference's own cases: the doctests of find_dense_regions / merge_regions / in_region
(filter_regions.py:38-58, utils.py:1185-1262, :1303-1312), the lambda expected results
(var.flt_preserved.vcf / var.flt_removed.vcf) and the argument checks of regression_tests.sh (filter_regions :2295-2766).
tests/golden/ref_regions.json.xz was produced by the reference's functions themselves (tests/golden/make_golden.py).
"""
import os
import random
import shutil
import time

import numpy as np
import pytest

from conftest import load_json_xz
from oracle import oracle as orc
from snp_pipeline_b200 import cfsan_snp_pipeline as cli
from snp_pipeline_b200 import device
from snp_pipeline_b200 import filter_regions as fr

LAMBDA = ["sample1", "sample2", "sample3", "sample4"]


def read(path):
    with open(path) as f:
        return f.read()


def run(line):
    cli.run_command_from_args(cli.parse_command_line(line))


@pytest.fixture(scope="module")
def ref_regions():
    return load_json_xz("ref_regions.json.xz")


def _case_args(c):
    samples = [(sid, [(chrom, pos) for chrom, pos in recs]) for sid, recs in c["samples"]]
    return samples, c["contig_len"], c["edge"], c["windows"], c["maxs"], c["mode"], c["outgroup"]


class _K7Emulation(object):
    """The contract of snpgpu_filter_regions (include/snpgpu.h) in plain Python: lets the CPU suite cover the host side's
    key / segment / edge preparation.  Test double only."""

    def filter_regions(self, snp_keys, seg_last, max_snps, window, edge_keys, edge_end):
        keys = [int(k) for k in snp_keys]
        regions = {}
        for k, e in zip(edge_keys, edge_end):
            regions.setdefault(int(k) >> 32, []).append((int(k) & 0xffffffff, int(e)))
        for i, k in enumerate(keys):
            for m, w in zip(max_snps, window):
                j = i + int(m)
                if j <= int(seg_last[i]) and (k & 0xffffffff) + int(w) - 1 >= (keys[j] & 0xffffffff):
                    regions.setdefault(k >> 32, []).append((k & 0xffffffff, keys[j] & 0xffffffff))
        return np.array([orc.in_region(k & 0xffffffff, regions.get(k >> 32, [])) for k in keys], dtype=bool)


@pytest.fixture
def lambda_work(tmp_path, golden_dir, monkeypatch):
    src = os.path.join(golden_dir, "lambda")
    dst = tmp_path / "lambda"
    for s in LAMBDA:
        os.makedirs(dst / "samples" / s)
        shutil.copy(os.path.join(src, "samples", s, "var.flt.vcf"), dst / "samples" / s / "var.flt.vcf")
    dirs = [str(dst / "samples" / s) for s in (LAMBDA[2], LAMBDA[0], LAMBDA[3], LAMBDA[1])]
    (dst / "sampleDirectories.txt").write_text("".join(d + "\n" for d in dirs))
    ref = os.path.join(golden_dir, "references", "lambda.fasta")
    monkeypatch.setenv("errorOutputFile", str(dst / "error.log"))
    monkeypatch.delenv("StopOnSampleError", raising=False)
    return dst, src, ref


# ------------------------------------------------------------------------------------------ oracle vs the reference
def test_oracle_dense_regions_doctests(ref_regions):
    for d in ref_regions["dense"]:
        assert orc.find_dense_regions(d["max"], d["window"], d["snps"]) == [tuple(r) for r in d["regions"]]
    # utils.py:1185-1262 (merge_regions) and :1303-1312 (in_region)
    assert orc.merge_regions([]) == []
    assert orc.merge_regions([(1, 2)]) == [(1, 2)]
    assert orc.merge_regions([(1, 2), (3, 4)]) == [(1, 4)]                     # adjacent
    assert orc.merge_regions([(1, 2), (4, 5)]) == [(1, 2), (4, 5)]
    assert orc.merge_regions([(1, 10), (2, 3)]) == [(1, 10)]                   # contained
    assert orc.merge_regions([(5, 9), (1, 6)]) == [(1, 9)]                     # unsorted, overlapping
    assert not orc.in_region(1, [])
    assert orc.in_region(5, [(1, 2), (4, 6)]) and not orc.in_region(3, [(1, 2), (4, 6)])


def test_oracle_flags_match_reference(ref_regions):
    assert len(ref_regions["cases"]) == 240
    for c in ref_regions["cases"]:
        samples, clen, edge, windows, maxs, mode, outgroup = _case_args(c)
        assert orc.filter_regions_flags(samples, clen, edge, windows, maxs, mode, outgroup) == c["removed"]


def test_oracle_lambda_files(golden_dir):
    src = os.path.join(golden_dir, "lambda", "samples")
    samples = [(s, [(c, p) for c, p in orc.vcf_sites(os.path.join(src, s, "var.flt.vcf"))]) for s in LAMBDA]
    flags = orc.filter_regions_flags(samples, {samples[0][1][0][0]: 48502})
    for (s, _), f in zip(samples, flags):
        pre, rem = orc.vcf_split_texts(read(os.path.join(src, s, "var.flt.vcf")), f)
        assert pre == read(os.path.join(src, s, "var.flt_preserved.vcf"))
        assert rem == read(os.path.join(src, s, "var.flt_removed.vcf"))


# ------------------------------------------------------------------------------------------ host layer, no GPU
def test_host_keys_against_reference_answers(ref_regions, monkeypatch):
    monkeypatch.setattr(device, "context", lambda: _K7Emulation())
    for c in ref_regions["cases"]:
        samples, clen, edge, windows, maxs, mode, outgroup = _case_args(c)
        kept = [(sid, recs) for sid, recs in samples if sid not in outgroup]
        got = fr.removed_flags([recs for _, recs in kept], clen, edge, windows, maxs, mode == "all")
        want = [f for f in c["removed"] if f is not None]
        assert [list(map(bool, g)) for g in got] == want


def test_subcommand_files_with_emulated_kernel(lambda_work, monkeypatch):
    """The file side (header order, echo of the records, outgroup copy, rebuild rules) without a device."""
    monkeypatch.setattr(device, "context", lambda: _K7Emulation())
    dst, src, ref = lambda_work
    sd = dst / "sampleDirectories.txt"
    run("filter_regions -v 0 %s %s" % (sd, ref))
    for s in LAMBDA:
        for kind in ("preserved", "removed"):
            assert read(dst / "samples" / s / ("var.flt_%s.vcf" % kind)) == read(os.path.join(src, "samples", s, "var.flt_%s.vcf" % kind))
    # fresh outputs are left alone
    stamp = {s: os.path.getmtime(dst / "samples" / s / "var.flt_removed.vcf") for s in LAMBDA}
    time.sleep(0.05)
    run("filter_regions -v 0 %s %s" % (sd, ref))
    assert stamp == {s: os.path.getmtime(dst / "samples" / s / "var.flt_removed.vcf") for s in LAMBDA}
    # a missing output rebuilds that sample only (filter_regions.py:233-236)
    os.remove(dst / "samples" / "sample2" / "var.flt_preserved.vcf")
    run("filter_regions -v 0 %s %s" % (sd, ref))
    assert read(dst / "samples" / "sample2" / "var.flt_preserved.vcf") == read(os.path.join(src, "samples", "sample2", "var.flt_preserved.vcf"))
    assert stamp["sample1"] == os.path.getmtime(dst / "samples" / "sample1" / "var.flt_removed.vcf")
    # outgroup: copied, header-only removed file; its SNPs no longer make regions for the others
    (dst / "og.txt").write_text("sample1\n")
    run("filter_regions -f -v 0 -g %s %s %s" % (dst / "og.txt", sd, ref))
    assert read(dst / "samples" / "sample1" / "var.flt_preserved.vcf") == read(dst / "samples" / "sample1" / "var.flt.vcf")
    removed = read(dst / "samples" / "sample1" / "var.flt_removed.vcf")
    assert removed.endswith("Sample1\n") and all(ln.startswith("#") for ln in removed.splitlines())
    samples = [(s, orc.vcf_sites(str(dst / "samples" / s / "var.flt.vcf"))) for s in LAMBDA]
    flags = orc.filter_regions_flags(samples, {samples[0][1][0][0]: 48502}, outgroup=["sample1"])
    for (s, _), f in zip(samples[1:], flags[1:]):
        pre, rem = orc.vcf_split_texts(read(dst / "samples" / s / "var.flt.vcf"), f)
        assert read(dst / "samples" / s / "var.flt_preserved.vcf") == pre
        assert read(dst / "samples" / s / "var.flt_removed.vcf") == rem


def test_argument_checks(lambda_work, capsys):
    dst, src, ref = lambda_work
    a = cli.parse_command_line("filter_regions dirs ref.fasta")
    assert (a.edgeLength, a.windowSizeList, a.maxSnpsList, a.mode, a.vcfFileName, a.outGroupFile, a.forceFlag) == \
        (500, [1000], [3], "all", "var.flt.vcf", None, False)
    for line, text in (("filter_regions d r -w 1000 100 -m 3", "same number of arguments"),
                       ("filter_regions d r -w 0", "length of the window must be a positive integer, and the input is 0"),
                       ("filter_regions d r -m 0", "maximum number of SNPs allowed must be a positive integer"),
                       ("filter_regions d r -l 0", "length of the edge regions must be a positive integer")):
        with pytest.raises(SystemExit) as e:
            cli.parse_command_line(line)
        assert e.value.code == 100
        assert text in capsys.readouterr().err
    # missing inputs: the reference's error protocol (regression_tests.sh:2295-2420)
    with pytest.raises(SystemExit) as e:
        run("filter_regions -v 0 %s %s" % (dst / "nope.txt", ref))
    assert e.value.code == 100
    assert "File of sample directories" in read(dst / "error.log")
    with pytest.raises(SystemExit) as e:
        run("filter_regions -v 0 %s %s" % (dst / "sampleDirectories.txt", dst / "nope.fasta"))
    assert e.value.code == 100
    assert "Reference file" in read(dst / "error.log")


# ------------------------------------------------------------------------------------------ K7 on the GPU
@pytest.mark.gpu
def test_gpu_flags_match_reference(ref_regions):
    for c in ref_regions["cases"]:
        samples, clen, edge, windows, maxs, mode, outgroup = _case_args(c)
        kept = [recs for sid, recs in samples if sid not in outgroup]
        got = fr.removed_flags(kept, clen, edge, windows, maxs, mode == "all")
        assert [list(map(bool, g)) for g in got] == [f for f in c["removed"] if f is not None]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["all", "each"])
def test_gpu_many_snps_against_oracle(mode):
    """A larger case than the golden ones: 40 samples x 3 contigs, thousands of SNPs, three (max, window) pairs."""
    rng = random.Random(11)
    clen = {"a": 3000000, "b": 40000, "c": 700}
    samples = []
    for s in range(40):
        recs = []
        for c, n in (("a", 3000), ("b", 400), ("c", 12)):
            hot = [rng.randint(1, clen[c]) for _ in range(6)]
            pos = [rng.randint(1, clen[c]) if rng.random() < 0.6 else max(1, min(clen[c], int(rng.gauss(rng.choice(hot), 300))))
                   for _ in range(n)]
            recs += [(c, p) for p in pos]
        rng.shuffle(recs)
        samples.append(("s%d" % s, recs))
    want = orc.filter_regions_flags(samples, clen, 500, [1000, 125, 15], [3, 2, 1], mode)
    got = fr.removed_flags([r for _, r in samples], clen, 500, [1000, 125, 15], [3, 2, 1], mode == "all")
    assert [list(map(bool, g)) for g in got] == want
    assert 0 < sum(map(sum, want)) < sum(len(r) for _, r in samples)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["all", "each"])
def test_gpu_subcommand_lambda_files(lambda_work, mode):
    dst, src, ref = lambda_work
    run("filter_regions -v 0 -M %s %s %s" % (mode, dst / "sampleDirectories.txt", ref))
    samples = [(s, orc.vcf_sites(str(dst / "samples" / s / "var.flt.vcf"))) for s in LAMBDA]
    flags = orc.filter_regions_flags(samples, {samples[0][1][0][0]: 48502}, mode=mode)
    for (s, _), f in zip(samples, flags):
        pre, rem = orc.vcf_split_texts(read(dst / "samples" / s / "var.flt.vcf"), f)
        assert read(dst / "samples" / s / "var.flt_preserved.vcf") == pre
        assert read(dst / "samples" / s / "var.flt_removed.vcf") == rem
        if mode == "all":                                   # the reference's bundled expected files
            assert pre == read(os.path.join(src, "samples", s, "var.flt_preserved.vcf"))
            assert rem == read(os.path.join(src, "samples", s, "var.flt_removed.vcf"))
