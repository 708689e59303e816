"""The host-side writer of the bench's synthetic pileups (oracle/synth_host.cpp -- the generator's line function,
snp_pipeline_b200/csrc/synth_line.cuh, compiled for the host; the input of `bench.py --impl reference`)."""
import numpy as np

from oracle import oracle as orc

CHROM = "gi|0000000|ref|SYN_5000K.1|"


def test_host_text_is_reproducible_and_well_formed():
    G = 30000
    a = orc.synth_pileup(7, 2, G, 24, 300, 0.3, 0.01, CHROM, threads=1)
    b = orc.synth_pileup(7, 2, G, 24, 300, 0.3, 0.01, CHROM, threads=5)     # (the thread count does not show)
    assert np.array_equal(a, b)
    lines = a.tobytes().split(b"\n")
    assert lines[-1] == b"" and len(lines) == G + 1
    for k in (0, 1, G // 2, G - 1):
        f = lines[k].split(b"\t")
        assert f[0] == CHROM.encode() and int(f[1]) == k + 1 and len(f) == 6
    own = orc.synth_sample_sites(7, 2, G, 24, 300, 0.3)
    assert 30 < own.size < 200 and np.all(np.diff(own.astype(np.int64)) > 0)
    # the oracle takes every line; at the sites the sample carries the call is (nearly always) not the reference base
    snps = [(CHROM, int(p)) for p in own]
    row, (cells, fails, _) = orc.pileup_consensus(a.tobytes(), snps, [], orc.make_params(min_cons_depth=3), parse_all=True,
                                                  want_lines=True)
    assert len(cells) == G
    ref = np.array([lines[int(p) - 1].split(b"\t")[2][0] for p in own], dtype=np.uint8)
    called = np.frombuffer(row, dtype=np.uint8)
    assert np.mean((called != ref) & (called != ord("-"))) > 0.8
