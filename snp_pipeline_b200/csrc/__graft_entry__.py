"""Driver entry points: build() compiles every native piece, smoke() runs one small hot-path invocation on cuda:0
and checks it against the oracle."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "snp_pipeline_b200", "csrc")


def build():
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo for every .cu (csrc/Makefile) -> libsnpgpu.so in-tree;
    gcc for the oracle's C restatement (building the checker is not using it)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", CSRC, "all", "cpusim"])
    from oracle import oracle as orc
    orc.build()
    import importlib
    pkg = importlib.import_module("snp_pipeline_b200")
    from snp_pipeline_b200 import _lib
    L = _lib.load()
    missing = [n for n in _lib.EXPORTS if not hasattr(L, n)]
    if missing:
        raise RuntimeError("libsnpgpu.so lacks symbols: %s" % missing)
    return pkg


def smoke():
    """One tiny sample through K1 (pileup -> consensus row), K2 (site union) and K4 (distances) on cuda:0, each
    checked against the oracle."""
    import random
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import linegen
    from oracle import oracle as orc
    from snp_pipeline_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        build()
    ctx = _lib.Context(0)
    rng = random.Random(1)
    n = 4000
    variant = {p: rng.choice("ACGT") for p in rng.sample(range(1, n + 1), 60)}
    text = linegen.pileup_text(1, n, sites=variant).encode()
    snps = [(linegen.CHROM, p) for p in sorted(variant)]
    sites = ctx.sites(snps)
    row, stats, lines = ctx.pileup_consensus(text, sites, _lib.make_params(min_cons_depth=3), _lib.MODE_ALL,
                                             want_lines=True)
    want_row, (cells, fails, _) = orc.pileup_consensus(text, snps, [], orc.make_params(min_cons_depth=3),
                                                       parse_all=True, want_lines=True)
    assert row == want_row, "K1 consensus row differs from the oracle"
    assert np.array_equal(lines & 0xff, cells) and np.array_equal(lines >> 8, fails), "K1 per-line calls differ"
    keys = np.array([p for p in sorted(variant)] * 2 + [7, 9], dtype=np.uint64)
    samp = np.array([0] * len(variant) + [1] * len(variant) + [2, 2], dtype=np.uint32)
    got = ctx.merge_sites(keys, samp)
    want = orc.merge_sites_keys(keys, samp)
    assert all(np.array_equal(a, b) for a, b in zip(got, want)), "K2 site union differs from the oracle"
    m = np.frombuffer(b"ACGT-Nacgt", dtype=np.uint8)[np.random.default_rng(0).integers(0, 10, size=(9, 777))]
    assert np.array_equal(ctx.pairwise_distance(m), orc.distance_matrix([bytes(r) for r in m])), "K4 differs"
    print("smoke ok: %d lines, %d parsed, %d via the exact path, %d kernels launched" %
          (stats.n_lines, stats.n_parsed, stats.n_general, ctx.launch_count))
    sites.close()
    ctx.close()


if __name__ == "__main__":
    build()
    if len(sys.argv) > 1 and sys.argv[1] == "smoke":
        smoke()
