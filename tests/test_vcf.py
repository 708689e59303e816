"""The per-sample consensus VCF: oracle restatement pinned against the reference's bundled
lambda consensus*.vcf files, the host-side writer's header, and -- on the GPU -- kernel K5's records against the
oracle plus the subcommand's files against the bundled ones (fileDate / source lines ignored, like
test/test_cfsan_snp_pipeline.py:183 does)."""
import os
import random
import shutil

import pytest

import linegen
from oracle import oracle as orc
from snp_pipeline_b200 import cfsan_snp_pipeline as cli
from snp_pipeline_b200 import pileup as gpu_pileup
from snp_pipeline_b200 import vcf_writer

LAMBDA = ["sample1", "sample2", "sample3", "sample4"]
IGNORE = ("##fileDate", "##source")


def read(path):
    with open(path) as f:
        return f.read()


def body(text):
    return "".join(l for l in text.splitlines(True) if not l.startswith("#"))


def comparable(text):
    return "".join(l for l in text.splitlines(True) if not l.startswith(IGNORE))


# ------------------------------------------------------------------------------------------ no GPU needed
@pytest.mark.parametrize("branch", ["", "_preserved"])
def test_oracle_vcf_matches_bundled_files(golden_dir, branch):
    root = os.path.join(golden_dir, "lambda")
    snps = orc.read_snp_list(os.path.join(root, "snplist%s.txt" % branch))
    for s in LAMBDA:
        d = os.path.join(root, "samples", s)
        text = open(os.path.join(d, "reads.all.pileup"), "rb").read()
        excl = orc.vcf_sites(os.path.join(d, "var.flt_removed.vcf")) if branch else []
        got = orc.consensus_vcf_body(text, snps, excl, orc.make_params(min_cons_depth=3))
        assert got == body(read(os.path.join(d, "consensus%s.vcf" % branch)))


def test_header_matches_bundled_file(golden_dir):
    gold = read(os.path.join(golden_dir, "lambda", "samples", "sample1", "consensus.vcf"))
    caller = gpu_pileup.ConsensusCaller(0.6, 3, 0, 0.0)
    filters = caller.get_filter_descriptions()
    filters.append(("Region", "Position is in dense region of snps or near the end of the contig."))
    hdr = vcf_writer.header_text("sample1", filters, "lambda_virus.fasta")
    want = "".join(l for l in gold.splitlines(True) if l.startswith("#"))
    assert comparable(hdr) == comparable(want)
    assert hdr.splitlines()[1].startswith("##fileDate=") and hdr.splitlines()[2].startswith("##source=CFSAN SNP-Pipeline ")


def test_vcf_doctest_cases():
    """The known answers of vcf_writer.py:170-290 (alt ordering, GT rules, no depth), through the oracle restatement."""
    op = orc.make_params(15, 0.5, 1, 0, 0.0)

    def fields(ref, depth, bases, quals, fails=None, gt="."):
        rep = orc.line_report(("ID\t42\t%s\t%d\t%s\t%s" % (ref, depth, bases, quals)).encode(), op)
        return orc.vcf_record_fields(rep, fails, gt)

    assert fields("G", 14, "aaaaAAAA...,,,", "00001111222333") == ("G", "A", "PASS", "1:14:6:8:3:3:4:4:PASS")
    assert fields("g", 14, "aaaaAAAA...,,,", "00001111222333", ["Fail"], ".") == ("G", "A", "Fail", ".:14:6:8:3:3:4:4:Fail")
    assert fields("g", 14, "aaaaAAAA...,,,", "00001111222333", ["Fail"], "0")[3].startswith("0:")
    assert fields("g", 14, "....,,,,aaaAAA", "00001111222333", ["Fail"], "1") == ("G", "A", "Fail", "1:14:8:6:4:4:3:3:Fail")
    assert fields("G", 14, "gaaaGGGG...,,,", "00001111222333") == ("G", "A", "PASS", "0:14:11:3:7:4:0:3:PASS")
    assert fields("g", 14, "ggggGGGG...,,,", "00001111222333") == ("G", ".", "PASS", "0:14:14:0:7:7:0:0:PASS")
    assert fields("G", 16, "..,,AAaaTTttCCcc", "0000111122223333", ["Fail"]) == \
        ("G", "A,C,T", "Fail", ".:16:4:4,4,4:2:2:2,2,2:2,2,2:Fail")
    assert fields("G", 23, "TttaaAAAcCC.......,,,,,", "00011111222333333333333") == \
        ("G", "A,C,T", "PASS", "0:23:12:5,3,3:7:5:3,2,1:2,1,2:PASS")
    rep = orc.line_report(b"ID\t42\tG\t0", op)
    assert orc.vcf_record_fields(rep, ["Fail"]) == ("G", ".", "Fail", ".:0:0:0:0:0:0:0:Fail")


# ------------------------------------------------------------------------------------------ GPU
def _gpu_vcf_body(ctx, text, snps, excl, ps, all_pos, gt=".", preserve=False):
    import io
    from snp_pipeline_b200 import _lib
    caller = gpu_pileup.ConsensusCaller(ps[1], ps[2], ps[3], ps[4])
    sites = ctx.sites(snps, excl)
    try:
        mode = _lib.MODE_ALL if all_pos else _lib.MODE_SITES
        params = caller.params(ps[0])
        ctx.pileup_consensus(text, sites, params, mode)
        rec, alt = ctx.pileup_vcf_records(sites, params, mode)
        # the same lines formatted on the device (K5's text pass): what the subcommand writes
        filter_texts = [";".join(caller.fail_names(m) or ["PASS"]) for m in range(_lib.VCF_FILTER_MASKS)]
        dev_text, n_dev = ctx.pileup_vcf_text(sites, params, mode, filter_texts, gt, preserve)
    finally:
        sites.close()
    buf = io.StringIO()
    buf.name = "mem.vcf"
    w = vcf_writer.SingleSampleWriter(buf, preserve)
    w.write_records(text, rec, alt, caller, gt)
    assert n_dev == len(rec)
    assert dev_text.tobytes().decode("ascii") == buf.getvalue()
    return buf.getvalue()


@pytest.fixture(scope="module")
def ctx():
    from snp_pipeline_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


PARAM_SETS = [(0, 0.6, 1, 0, 0.0), (0, 0.6, 3, 0, 0.0), (15, 0.6, 3, 1, 0.1), (0, 0.55, 2, 2, 0.25), (20, 0.9, 10, 0, 0.5)]


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(5))
def test_records_match_oracle_realistic(ctx, seed):
    rng = random.Random(900 + seed)
    n = 2500
    sites = {p: rng.choice("ACGT") for p in rng.sample(range(1, n + 1), 150)}
    text = linegen.pileup_text(40 + seed, n, sites=sites, gaps=0.01).encode()
    snps = [(linegen.CHROM, p) for p in sorted(rng.sample(range(1, n + 40), 200))]
    excl = [(linegen.CHROM, p) for p in rng.sample(range(1, n), 40)] if seed % 2 else []
    ps = PARAM_SETS[seed % len(PARAM_SETS)]
    gt = ".01"[seed % 3]
    for all_pos in (False, True):
        want = orc.consensus_vcf_body(text, snps, excl, orc.make_params(*ps), parse_all=all_pos, failed_snp_gt=gt,
                                      preserve_ref_case=bool(seed & 1))
        assert _gpu_vcf_body(ctx, text, snps, excl, ps, all_pos, gt, bool(seed & 1)) == want


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(6))
def test_records_match_oracle_nasty(ctx, seed):
    """Odd alphabets (IUPAC, digits, '>' '<'), indel tokens, caret chains, short quality strings: open-alphabet ALT lists."""
    rng = random.Random(950 + seed)
    n = 500
    raw = linegen.pileup_text(300 + seed, n, nasty=0.6).encode()
    op = orc.make_params()
    keep = [ln + b"\n" for ln in raw.split(b"\n")[:-1] if orc.line_report(ln, op)["status"] == 0 and b"\r" not in ln
            and b"\x0b" not in ln and b"\x1c" not in ln]
    text = b"".join(keep)
    snps = [(linegen.CHROM, p) for p in rng.sample(range(1, n + 10), 250)]
    excl = [(linegen.CHROM, p) for p in rng.sample(range(1, n), 30)]
    ps = PARAM_SETS[seed % len(PARAM_SETS)]
    for all_pos in (False, True):
        want = orc.consensus_vcf_body(text, snps, excl, orc.make_params(*ps), parse_all=all_pos)
        assert _gpu_vcf_body(ctx, text, snps, excl, ps, all_pos) == want


@pytest.mark.gpu
@pytest.mark.parametrize("branch", ["", "_preserved"])
def test_subcommand_writes_bundled_vcf(tmp_path, golden_dir, branch, monkeypatch):
    """call_consensus --vcfFileName, the way run.py:709 always calls it, reproduces the bundled consensus*.vcf."""
    src = os.path.join(golden_dir, "lambda")
    monkeypatch.setenv("errorOutputFile", str(tmp_path / "error.log"))
    for s in LAMBDA:
        sdir = tmp_path / s
        os.makedirs(sdir)
        for f in ("reads.all.pileup", "var.flt_removed.vcf"):
            shutil.copy(os.path.join(src, "samples", s, f), sdir / f)
        extra = "-e %s/var.flt_removed.vcf" % sdir if branch else ""
        line = ("call_consensus -v 0 -l %s/snplist%s.txt -o %s/consensus%s.fasta --minConsDpth 3 --vcfFileName consensus%s.vcf "
                "--vcfRefName lambda_virus.fasta %s %s/reads.all.pileup" % (src, branch, sdir, branch, branch, extra, sdir))
        cli.run_command_from_args(cli.parse_command_line(line))
        assert read(sdir / ("consensus%s.fasta" % branch)) == read(os.path.join(src, "samples", s, "consensus%s.fasta" % branch))
        assert comparable(read(sdir / ("consensus%s.vcf" % branch))) == \
            comparable(read(os.path.join(src, "samples", s, "consensus%s.vcf" % branch)))
