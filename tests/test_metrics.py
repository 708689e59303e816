"""K6 / metrics.py: the pileup by-products of collect_metrics (collect_metrics.py:109-128, 313-342)."""
import io
import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import linegen
from oracle import oracle as orc


def python_depth_sum(text: bytes, tmp_path):
    """The reference's own loop, run by Python on a real file (universal newlines and all)."""
    p = tmp_path / "x.pileup"
    p.write_bytes(text)
    depth_sum = 0
    with open(p) as f:
        for line in f:
            tokens = line.split()
            try:
                depth_sum += int(tokens[3])
            except (ValueError, IndexError):
                pass
    return depth_sum


def _texts():
    rng = random.Random(9)
    out = [b"", b"\n\n", b"c 1 A 7 ....... IIIIIII", b"c\t1\tA\t+1_0\tx\ty\r\nc\t2\tA\t0x10\t.\tI\rc 3 A 5\n  c\t4\tA\t-3 \n",
           b"a b c\n1 2 3 4 5 6 7\n\x0b\x0c 1 2 3 9\x1c8 \n", b"c 1 A 1_\nc 1 A _1\nc 1 A 1__2\nc 1 A 007\n"]
    for seed in range(4):
        t = linegen.pileup_text(300 + seed, 600, nasty=0.4 if seed % 2 else 0.0).encode()
        if seed == 2:
            t = t.replace(b"\n", b"\r\n")
        if seed == 3:
            t = t.replace(b"\n", b"\r", 40)
        out.append(t)
    return out


def test_oracle_depth_sum_against_python(tmp_path):
    for t in _texts():
        if any(b >= 0x80 for b in t):
            continue
        assert orc.depth_sum(t)[0] == python_depth_sum(t, tmp_path), t[:80]
    assert orc.mean_pileup_depth_text(b"c 1 A 7 . I\nc 2 A 8 . I\n", 4) == "3.75"
    assert orc.mean_pileup_depth_text(b"c 1 A 0 * *\n", 4) == ""
    with pytest.raises(orc.OracleError):
        orc.depth_sum(b"c 1 A 5 . I\nc 2 \xc3\xa9 5 . I\n")


@pytest.mark.gpu
def test_depth_sum_kernel(tmp_path):
    from snp_pipeline_b200 import _lib, metrics
    ctx = _lib.Context(0)
    for t in _texts():
        assert ctx.pileup_depth_sum(t) == orc.depth_sum(t), t[:80]
    big = linegen.pileup_text(77, 20000).encode()
    assert ctx.pileup_depth_sum(big) == orc.depth_sum(big)
    with pytest.raises(_lib.SnpGpuError) as e:
        ctx.pileup_depth_sum(b"c 1 A 5 . I\nc 2 \xc3\xa9 5 . I\n")
    assert e.value.code == _lib.E_DOMAIN and e.value.offset == 12
    with pytest.raises(_lib.SnpGpuError):
        ctx.pileup_depth_sum(b"c 1 A 99999999999999999999 . I\n")
    ctx.close()
    p = tmp_path / "reads.all.pileup"
    p.write_bytes(big)
    ref = tmp_path / "ref.fasta"
    ref.write_text(">c1 x\n" + "ACGT" * 100 + "\n>c2\nAC\n")
    assert metrics.reference_length(str(ref)) == 402
    assert metrics.mean_pileup_depth(str(p), 402) == orc.mean_pileup_depth_text(big, 402)
    fa = tmp_path / "snpma.fasta"
    fa.write_text(">s1\nAC-T\n--\n>s2 d\nA---\n")
    assert metrics.count_missing_snp_matrix_positions(str(fa), "s1") == 3
    assert metrics.count_missing_snp_matrix_positions(str(fa), "s2") == 3
    assert metrics.count_missing_snp_matrix_positions(str(fa), "nope") == 0
