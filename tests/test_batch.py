"""snp_pipeline_b200.batch: the package-level multi-sample / multi-GPU driver against the oracle's three files."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "hotpath_worker.py")


def _run(cmd, tmp_path):
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "HOTPATH OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["sites", "all"])
def test_hot_path_files_one_gpu(tmp_path, mode):
    """19 samples (two launch sequences), two contigs (one with a 45-byte name): snplist.txt, snpma.fasta and
    snp_distance_matrix.tsv byte-identical to the oracle's."""
    _run([sys.executable, WORKER, str(tmp_path), "19", "20000"] + (["all"] if mode == "all" else []), tmp_path)


@pytest.mark.gpu
def test_hot_path_files_two_gpus(tmp_path):
    """The same on two ranks / two GPUs (NCCL): uneven blocks (9 samples -> 5 + 4), one all-gather of the site lists,
    one of the rows, rank 0's files against the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
          "127.0.0.1", "--master-port", "29541", WORKER, str(tmp_path), "9", "30000"], tmp_path)


@pytest.mark.gpu
def test_lambda_sample_dirs_one_pass(tmp_path, golden_dir):
    """BASELINE config 1 through the one-pass driver: the four lambda sample directories in, the reference's expected
    snplist.txt / snpma.fasta / distance files and every consensus.fasta out (run.py:682-732, 770-784)."""
    import shutil
    from snp_pipeline_b200 import batch
    src = os.path.join(golden_dir, "lambda")
    names = sorted(os.listdir(os.path.join(src, "samples")))
    for s in names:
        os.makedirs(tmp_path / "samples" / s)
        for f in ("reads.all.pileup", "var.flt.vcf"):
            shutil.copy(os.path.join(src, "samples", s, f), tmp_path / "samples" / s / f)
    sd = tmp_path / "sampleDirectories.txt"
    sd.write_text("".join(str(tmp_path / "samples" / s) + "\n" for s in reversed(names)))
    batch.main([str(sd), "-o", str(tmp_path / "out"), "-D", "3"])
    for mine, theirs in (("snplist.txt", "snplist.txt"), ("snpma.fasta", "snpma.fasta"),
                         ("snp_distance_matrix.tsv", "snp_distance_matrix.tsv"),
                         ("snp_distance_pairwise.tsv", "snp_distance_pairwise.tsv")):
        assert open(tmp_path / "out" / mine).read() == open(os.path.join(src, theirs)).read(), mine
    for s in names:
        assert open(tmp_path / "samples" / s / "consensus.fasta").read() == \
            open(os.path.join(src, "samples", s, "consensus.fasta")).read(), s
