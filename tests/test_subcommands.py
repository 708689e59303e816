"""The reference-facing layer: the four subcommands behind the reference's own argument / file / exit-code contract.

Modelled on the reference's per-step unit tests (test/test_cfsan_snp_pipeline.py:164-224: run one step through
parse_command_line + run_command_from_args and compare the file with the bundled expected result) and on the error-path
cases of test/regression_tests.sh (merge_sites :2772-3044, call_consensus :3047-3330, snp_matrix :3609-3882,
distance :4598-4754, ZeroSnps :5878-5960, ExcessiveSnps :6287-6384).  Steps that compute (merge_sites,
call_consensus, distance) need the GPU; argument handling, error protocol, snp_matrix and the ABI surface do not.
"""
import ctypes
import os
import re
import shutil

import pytest

from snp_pipeline_b200 import cfsan_snp_pipeline as cli
from snp_pipeline_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAMBDA = ["sample1", "sample2", "sample3", "sample4"]


def run(line):
    args = cli.parse_command_line(line)
    cli.run_command_from_args(args)


def read(path):
    with open(path) as f:
        return f.read()


@pytest.fixture
def work(tmp_path, golden_dir, monkeypatch):
    """A scratch copy of the lambda inputs: samples/*/{reads.all.pileup, var.flt*.vcf} + sampleDirectories.txt."""
    src = os.path.join(golden_dir, "lambda")
    dst = tmp_path / "lambda"
    for s in LAMBDA:
        os.makedirs(dst / "samples" / s)
        for f in ("reads.all.pileup", "var.flt.vcf", "var.flt_preserved.vcf", "var.flt_removed.vcf"):
            shutil.copy(os.path.join(src, "samples", s, f), dst / "samples" / s / f)
    dirs = [str(dst / "samples" / s) for s in (LAMBDA[2], LAMBDA[0], LAMBDA[3], LAMBDA[1])]   # not sorted on purpose
    (dst / "sampleDirectories.txt").write_text("".join(d + "\n" for d in dirs))
    monkeypatch.setenv("errorOutputFile", str(dst / "error.log"))
    monkeypatch.delenv("StopOnSampleError", raising=False)
    return dst, src, dirs


# ------------------------------------------------------------------------------------------ no GPU needed
def test_argument_defaults_and_validators():
    a = cli.parse_command_line("call_consensus reads.all.pileup")
    assert (a.snpListFile, a.consensusFile, a.minBaseQual, a.minConsFreq, a.minConsDpth, a.minConsStrdDpth,
            a.minConsStrdBias, a.excludeFile, a.vcfFileName, a.vcfRefName, a.vcfAllPos, a.vcfFailedSnpGt, a.verbose) == \
        ("snplist.txt", "consensus.fasta", 0, 0.6, 1, 0, 0, None, None, "Unknown reference", False, ".", 1)
    a = cli.parse_command_line("merge_sites dirs.txt dirs.filtered")
    assert (a.vcfFileName, a.maxSnps, a.snpListFile, a.forceFlag) == ("var.flt.vcf", -1, "snplist.txt", False)
    a = cli.parse_command_line("snp_matrix dirs.txt")
    assert (a.consFileName, a.snpmaFile) == ("consensus.fasta", "snpma.fasta")
    a = cli.parse_command_line("distance -p p.tsv snpma.fasta")
    assert (a.pairwiseFile, a.matrixFile, a.inputFile) == ("p.tsv", None, "snpma.fasta")
    for bad in ("call_consensus -c 0.5 x", "call_consensus -c 1.01 x", "call_consensus -b 0.6 x",
                "call_consensus --vcfFailedSnpGt 2 x", "nosuchcommand"):
        with pytest.raises(SystemExit) as e:
            cli.parse_command_line(bad)
        assert e.value.code == 2                                    # argparse errors exit 2


def test_abi_exports_every_declared_symbol():
    header = read(os.path.join(ROOT, "include", "snpgpu.h"))
    declared = set(re.findall(r"^\s*(?:int|void|size_t|uint64_t|const char \*)\s*(snpgpu_[a-z0-9_]+)\s*\(", header, re.M))
    assert len(declared) >= 20
    assert os.path.exists(_lib.LIB_PATH), "libsnpgpu.so is not built (run __graft_entry__.build())"
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared
    assert L.snpgpu_abi_version() == 2


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises; it never computes on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.SnpGpuError):
        _lib.Context(0)
    import snp_pipeline_b200
    pkg = os.path.dirname(snp_pipeline_b200.__file__)
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            assert "oracle" not in read(os.path.join(pkg, name)), "%s mentions the oracle" % name


def test_snp_matrix_golden(work):
    dst, src, dirs = work
    for s in LAMBDA:
        shutil.copy(os.path.join(src, "samples", s, "consensus.fasta"), dst / "samples" / s / "consensus.fasta")
    run("snp_matrix -v 0 -o %s/snpma.fasta %s/sampleDirectories.txt" % (dst, dst))
    assert read(dst / "snpma.fasta") == read(os.path.join(src, "snpma.fasta"))
    # fresh target: skipped with the reference's message
    os.utime(dst / "snpma.fasta", (2e9, 2e9))
    (dst / "snpma.fasta").write_text("stale")
    os.utime(dst / "snpma.fasta", (2e9, 2e9))
    run("snp_matrix -v 0 -o %s/snpma.fasta %s/sampleDirectories.txt" % (dst, dst))
    assert read(dst / "snpma.fasta") == "stale"
    run("snp_matrix -f -v 0 -o %s/snpma.fasta %s/sampleDirectories.txt" % (dst, dst))
    assert read(dst / "snpma.fasta") == read(os.path.join(src, "snpma.fasta"))


def test_snp_matrix_missing_inputs(work, monkeypatch):
    dst, src, dirs = work
    with pytest.raises(SystemExit) as e:                           # all consensus files missing -> global error
        run("snp_matrix -v 0 -o %s/snpma.fasta %s/sampleDirectories.txt" % (dst, dst))
    assert e.value.code == 100
    assert "all 4 consensus fasta files were missing or empty" in read(dst / "error.log")
    shutil.copy(os.path.join(src, "samples", "sample1", "consensus.fasta"), dst / "samples" / "sample1" / "consensus.fasta")
    with pytest.raises(SystemExit) as e:                           # some missing, StopOnSampleError unset -> 100
        run("snp_matrix -v 0 -o %s/snpma.fasta %s/sampleDirectories.txt" % (dst, dst))
    assert e.value.code == 100
    monkeypatch.setenv("StopOnSampleError", "false")               # ... false -> the step goes on with what exists
    run("snp_matrix -v 0 -o %s/snpma.fasta %s/sampleDirectories.txt" % (dst, dst))
    assert read(dst / "snpma.fasta") == read(os.path.join(src, "samples", "sample1", "consensus.fasta"))
    with pytest.raises(SystemExit) as e:
        run("snp_matrix -v 0 %s/nosuchfile.txt" % dst)
    assert e.value.code == 100


def test_error_protocol_without_compute(work, monkeypatch):
    dst, src, dirs = work
    pile = dst / "samples" / "sample1" / "reads.all.pileup"
    with pytest.raises(SystemExit) as e:                           # missing snplist: global error (regression :3094+)
        run("call_consensus -v 0 -l %s/nosnplist.txt -o %s/c.fasta %s" % (dst, dst, pile))
    assert e.value.code == 100
    assert "cannot call consensus without the snplist file" in read(dst / "error.log")
    (dst / "snplist.txt").write_text("")
    with pytest.raises(SystemExit) as e:                           # missing pileup: sample error
        run("call_consensus -v 0 -l %s/snplist.txt -o %s/c.fasta %s/nopileup" % (dst, dst, dst))
    assert e.value.code == 100
    monkeypatch.setenv("StopOnSampleError", "false")
    with pytest.raises(SystemExit) as e:
        run("call_consensus -v 0 -l %s/snplist.txt -o %s/c.fasta %s/nopileup" % (dst, dst, dst))
    assert e.value.code == 98
    with pytest.raises(SystemExit) as e:                           # no output named (regression :4690+)
        run("distance -v 0 %s/snplist.txt" % dst)
    assert e.value.code == 100
    with pytest.raises(SystemExit) as e:
        run("distance -v 0 -p %s/p.tsv %s/nosnpma.fasta" % (dst, dst))
    assert e.value.code == 100
    empty = dst / "emptydirs.txt"
    empty.write_text("%s/nodir1\n%s/nodir2\n" % (dst, dst))
    with pytest.raises(SystemExit) as e:                           # every VCF missing
        run("merge_sites -v 0 -o %s/s.txt %s %s.filtered" % (dst, empty, empty))
    assert e.value.code == 100
    assert "all 2 VCF files were missing or empty" in read(dst / "error.log")


# ------------------------------------------------------------------------------------------ GPU: files, byte for byte
@pytest.mark.gpu
@pytest.mark.parametrize("branch", ["", "_preserved"])
def test_lambda_steps_reproduce_expected_files(work, branch):
    """BASELINE config 1 through the drop-in layer: merge_sites -> call_consensus x4 -> snp_matrix -> distance."""
    dst, src, dirs = work
    sd = dst / "sampleDirectories.txt"
    run("merge_sites -v 0 -n var.flt%s.vcf -o %s/snplist%s.txt %s %s.filtered" % (branch, dst, branch, sd, sd))
    assert read(dst / ("snplist%s.txt" % branch)) == read(os.path.join(src, "snplist%s.txt" % branch))
    assert read(str(sd) + ".filtered") == read(sd)
    for s in LAMBDA:
        sdir = dst / "samples" / s
        extra = "-e %s/var.flt_removed.vcf" % sdir if branch else ""
        run("call_consensus -v 0 -l %s/snplist%s.txt -o %s/consensus%s.fasta --minConsDpth 3 %s %s/reads.all.pileup"
            % (dst, branch, sdir, branch, extra, sdir))
        assert read(sdir / ("consensus%s.fasta" % branch)) == read(os.path.join(src, "samples", s, "consensus%s.fasta" % branch))
        run("call_consensus -f -v 0 --vcfAllPos -l %s/snplist%s.txt -o %s/c2.fasta --minConsDpth 3 %s %s/reads.all.pileup"
            % (dst, branch, sdir, extra, sdir))
        assert read(sdir / "c2.fasta") == read(sdir / ("consensus%s.fasta" % branch))
    run("snp_matrix -v 0 -c consensus%s.fasta -o %s/snpma%s.fasta %s" % (branch, dst, branch, sd))
    assert read(dst / ("snpma%s.fasta" % branch)) == read(os.path.join(src, "snpma%s.fasta" % branch))
    run("distance -v 0 -p %s/pw%s.tsv -m %s/mx%s.tsv %s/snpma%s.fasta" % (dst, branch, dst, branch, dst, branch))
    assert read(dst / ("pw%s.tsv" % branch)) == read(os.path.join(src, "snp_distance_pairwise%s.tsv" % branch))
    assert read(dst / ("mx%s.tsv" % branch)) == read(os.path.join(src, "snp_distance_matrix%s.tsv" % branch))


@pytest.mark.gpu
@pytest.mark.parametrize("dataset", ["agona", "listeria"])
def test_real_datasets_merge_matrix_distance(tmp_path, golden_dir, dataset):
    """BASELINE config 3 as far as the bundled data goes (no pileups ship for Agona / Listeria): merge_sites,
    snp_matrix and distance against the reference's expected files."""
    src = os.path.join(golden_dir, dataset)
    names = sorted(os.listdir(os.path.join(src, "samples")))
    sd = tmp_path / "sampleDirectories.txt"
    sd.write_text("".join(os.path.join(src, "samples", n) + "\n" for n in reversed(names)))
    run("merge_sites -v 0 -o %s/snplist.txt %s %s.filtered" % (tmp_path, sd, sd))
    assert read(tmp_path / "snplist.txt") == read(os.path.join(src, "snplist.txt"))
    if os.path.exists(os.path.join(src, "samples", names[0], "consensus.fasta")):
        run("snp_matrix -v 0 -o %s/snpma.fasta %s" % (tmp_path, sd))
        assert read(tmp_path / "snpma.fasta") == read(os.path.join(src, "snpma.fasta"))
    for suffix in ("", "_preserved"):
        if not os.path.exists(os.path.join(src, "snpma%s.fasta" % suffix)):
            continue
        run("distance -f -v 0 -p %s/pw.tsv -m %s/mx.tsv %s/snpma%s.fasta" % (tmp_path, tmp_path, src, suffix))
        assert read(tmp_path / "mx.tsv") == read(os.path.join(src, "snp_distance_matrix%s.tsv" % suffix))
        pw = os.path.join(src, "snp_distance_pairwise%s.tsv" % suffix)
        if os.path.exists(pw):
            assert read(tmp_path / "pw.tsv") == read(pw)


@pytest.mark.gpu
def test_maxsnps_excludes_samples(work):
    """regression_tests.sh:6287-6384: --maxsnps drops whole samples from the union and from the filtered list."""
    from snp_pipeline_b200 import utils
    dst, src, dirs = work
    sd = dst / "sampleDirectories.txt"
    counts = {d: len(utils.convert_vcf_file_to_snp_set(os.path.join(d, "var.flt.vcf"))) for d in dirs}
    cut = sorted(counts.values())[1]
    run("merge_sites -v 0 --maxsnps %d -o %s/snplist.txt %s %s.filtered" % (cut, dst, sd, sd))
    kept = [d for d in dirs if counts[d] <= cut]
    assert 0 < len(kept) < 4
    assert read(str(sd) + ".filtered") == "".join(d + "\n" for d in kept)
    kept_names = {os.path.basename(d) for d in kept}
    lines = read(dst / "snplist.txt").splitlines()
    assert lines and all(set(ln.split("\t")[3:]) <= kept_names for ln in lines)
    assert all(int(ln.split("\t")[2]) == len(ln.split("\t")[3:]) for ln in lines)


@pytest.mark.gpu
def test_call_consensus_edge_files(work, monkeypatch):
    dst, src, dirs = work
    sdir = dst / "samples" / "sample1"
    # empty snplist is not an error: a header-only fasta (regression_tests.sh:3156-3211, :5920-5934)
    (dst / "empty.txt").write_text("")
    run("call_consensus -v 0 -l %s/empty.txt -o %s/e.fasta %s/reads.all.pileup" % (dst, sdir, sdir))
    assert read(sdir / "e.fasta") == ">sample1\n"
    # corrupt snplist -> uncaught exception -> exit 100, or 98 when StopOnSampleError=false (:3047-3092)
    (dst / "corrupt.txt").write_text("gi|9626243|ref|NC_001416.1|\tnot_a_number\t1\tsample1\n")
    with pytest.raises(ValueError):
        run("call_consensus -v 0 -l %s/corrupt.txt -o %s/x.fasta %s/reads.all.pileup" % (dst, sdir, sdir))
    # a malformed pileup line at a snp position raises like the reference (IndexError: five columns)
    snplist = os.path.join(src, "snplist.txt")
    first = read(snplist).split("\t")[1]
    text = read(sdir / "reads.all.pileup").splitlines(True)
    k = [i for i, ln in enumerate(text) if ln.split("\t")[1] == first][0]
    text[k] = "\t".join(text[k].split("\t")[:5]) + "\n"
    (sdir / "broken.pileup").write_text("".join(text))
    with pytest.raises(IndexError):
        run("call_consensus -v 0 -l %s -o %s/b.fasta --minConsDpth 3 %s/broken.pileup" % (snplist, sdir, sdir))
    # freshness: an up-to-date consensus file is left alone
    run("call_consensus -v 0 -l %s -o %s/c.fasta --minConsDpth 3 %s/reads.all.pileup" % (snplist, sdir, sdir))
    (sdir / "c.fasta").write_text("stale")
    os.utime(sdir / "c.fasta", (2e9, 2e9))
    run("call_consensus -v 0 -l %s -o %s/c.fasta --minConsDpth 3 %s/reads.all.pileup" % (snplist, sdir, sdir))
    assert read(sdir / "c.fasta") == "stale"


@pytest.mark.gpu
def test_distance_ragged_and_lowercase(tmp_path):
    (tmp_path / "m.fasta").write_text(">b\nACGTAC\nGT\n>a\nacgtNN\n>>c\nTTTTACGTAA\n")
    run("distance -v 0 -m %s/mx.tsv -p %s/pw.tsv %s/m.fasta" % (tmp_path, tmp_path, tmp_path))
    # ids sorted: a (6), b (8), c (10); a-b over 6 columns: acgtNN vs ACGTAC -> 0; a-c: acgt vs TTTT -> 3; b-c over 8
    assert read(tmp_path / "mx.tsv") == "\ta\tb\tc\na\t0\t0\t3\nb\t0\t0\t3\nc\t3\t3\t0\n"
    (tmp_path / "bad.fasta").write_text(">a\nACGTACGT\n>b\nACG\n")
    with pytest.raises(IndexError):                                # a later sequence is shorter: the reference raises
        run("distance -v 0 -m %s/mx2.tsv %s/bad.fasta" % (tmp_path, tmp_path))


@pytest.mark.gpu
@pytest.mark.parametrize("dataset", ["lambda", "agona", "listeria"])
@pytest.mark.parametrize("suffix", ["", "_preserved"])
def test_snp_reference_golden(tmp_path, golden_dir, dataset, suffix, monkeypatch):
    """`snp_reference` reproduces the bundled referenceSNP*.fasta byte for byte; freshness rule as in the reference."""
    monkeypatch.setenv("errorOutputFile", str(tmp_path / "error.log"))
    root = os.path.join(golden_dir, dataset)
    ref = os.path.join(golden_dir, "references", dataset + ".fasta")
    out = tmp_path / "referenceSNP.fasta"
    line = "snp_reference -v 0 -l %s/snplist%s.txt -o %s %s" % (root, suffix, out, ref)
    cli.run_command_from_args(cli.parse_command_line(line))
    assert out.read_text() == open(os.path.join(root, "referenceSNP%s.fasta" % suffix)).read()
    out.write_text("stale but newer than its inputs\n")
    cli.run_command_from_args(cli.parse_command_line(line))
    assert out.read_text() == "stale but newer than its inputs\n"          # not rebuilt without -f
    cli.run_command_from_args(cli.parse_command_line(line.replace("snp_reference", "snp_reference -f", 1)))
    assert out.read_text() == open(os.path.join(root, "referenceSNP%s.fasta" % suffix)).read()


@pytest.mark.gpu
def test_snp_reference_indexing_and_errors(tmp_path, monkeypatch):
    """Python's indexing in utils.py:1108 (position 0 and negatives count from the end, anything further raises
    IndexError -> exit 100 through handle_global_exception), several contigs, lower case, a contig without SNPs."""
    from oracle import oracle as orc
    from snp_pipeline_b200 import snp_reference
    monkeypatch.setenv("errorOutputFile", str(tmp_path / "error.log"))
    ref = tmp_path / "ref.fasta"
    ref.write_text("; comment in front of the first record\n>b second contig\nacgtnACGTN\nrykm\n>a\nTTGCA\n  gg cc\n>empty\n")
    snps = tmp_path / "snplist.txt"
    snps.write_text("b\t1\t1\ts1\nb\t14\t1\ts1\na\t9\t2\ts1\ts2\nb\t0\t1\ts2\nb\t-3\t1\ts2\nzz\t5\t1\ts1\na\t1\t1\ts1\n")
    want = orc.reference_snp_text(str(ref), str(snps))
    assert want == ">a\nCT\n>b\nAMMR\n>empty\n"
    assert snp_reference.reference_snp_text(str(ref), str(snps)) == want
    snps.write_text("a\t10\t1\ts1\n")
    with pytest.raises(IndexError):
        snp_reference.reference_snp_text(str(ref), str(snps))
    args = cli.parse_command_line("snp_reference -v 0 -l %s -o %s %s" % (snps, tmp_path / "o.fasta", ref))
    with pytest.raises(IndexError):                                # uncaught, like in the reference ...
        cli.run_command_from_args(args)
    with pytest.raises(SystemExit) as e:                           # ... where the step's excepthook turns it into exit 100
        args.excepthook(IndexError, IndexError("index out of range"), None)
    assert e.value.code == 100
