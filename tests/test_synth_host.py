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


def test_bench_reference_arm_runs_without_the_gpu_library():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm): one JSON line with the contract's
    keys, inputs from the host generator, and neither torch nor the GPU library imported by that process."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json, io, contextlib; sys.argv = ['bench.py', '--impl', 'reference', '--samples', '2', '--genome-len', '60000', "
            "'--steps', '1', '--warmup', '1']; import bench; bench.main(); "
            "print(json.dumps({'torch': 'torch' in sys.modules, 'lib': 'snp_pipeline_b200._lib' in sys.modules}))")
    out = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    line, loaded = json.loads(lines[-2]), json.loads(lines[-1])
    assert loaded == {"torch": False, "lib": False}
    assert line["impl"] == "reference" and line["unit"] == "positions/s" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    rp = cb["reference_python"]
    if rp is not None and "unavailable" not in rp:             # (only where oracle/stage_ref.py staged the reference's modules)
        assert rp["positions_per_s_per_core"] > 0 and rp["filter_mode"]["positions_per_s_per_core"] > 0
        assert rp["distance"]["pair_sites_per_s_per_core"] > 0
        assert rp["c1_lambda"]["consensus_identical_to_bundled"] is True
