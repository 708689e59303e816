"""Parity of the CUDA path (through the C ABI, libsnpgpu.so) with the oracle -- the `-m gpu` tests proper.

Bit-exact: consensus rows, per-line cells and fail masks, first-error codes and offsets, merged site lists, distance
matrices.  Inputs: the reference's bundled lambda-virus / Agona / Listeria files (tests/golden), the golden vectors
produced by the reference's own Python (ref_files), seeded synthetic texts (tests/linegen.py) and device-generated
pileups at BASELINE.json's sample size.
"""
import ctypes
import os
import random

import numpy as np
import pytest

import linegen
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

PARAM_SETS = [(0, 0.6, 1, 0, 0.0), (0, 0.6, 3, 0, 0.0), (15, 0.6, 3, 1, 0.1), (0, 0.55, 2, 2, 0.25),
              (20, 0.9, 10, 0, 0.5), (0, 1.0, 0, 0, 0.0), (-5, 0.51, 1, 3, 0.33)]


@pytest.fixture(scope="module")
def ctx():
    from snp_pipeline_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def _gpu_params(ps):
    from snp_pipeline_b200 import _lib
    return _lib.make_params(*ps)


def _line_offsets(text):
    """byte offset of every line start under universal newlines"""
    offs, i, n = [], 0, len(text)
    while i < n:
        offs.append(i)
        j = i
        while j < n and text[j] not in (10, 13):
            j += 1
        if j < n and text[j] == 13 and j + 1 < n and text[j + 1] == 10:
            j += 1
        i = j + 1
    return offs


def _compare(ctx, text, snps, excl, ps, all_pos):
    from snp_pipeline_b200 import _lib
    op = orc.make_params(*ps)
    want_err = 0
    try:
        want_row, (cells, fails, poss) = orc.pileup_consensus(text, snps, excl, op, parse_all=all_pos, want_lines=True)
    except orc.OracleError as e:
        want_err, err_line = e.status, e.line
    sites = ctx.sites(snps, excl)
    try:
        mode = _lib.MODE_ALL if all_pos else _lib.MODE_SITES
        try:
            out = ctx.pileup_consensus(text, sites, _gpu_params(ps), mode, want_lines=all_pos)
        except _lib.SnpGpuError as e:
            assert want_err and e.code == want_err, (e.code, want_err, str(e))
            assert e.offset == _line_offsets(text)[err_line]
            return None
        assert not want_err, "the oracle raises (%d) and the kernel does not" % want_err
        row, stats = out[0], out[1]
        assert row == want_row
        assert stats.n_lines == len(_line_offsets(text))
        # call_consensus.py:184 "called consensus positions": snplist positions that got at least one parsed line
        chroms = {c for c, _ in snps}
        assert stats.n_called <= len(set(snps))
        if len(chroms) == 1:
            c0 = next(iter(chroms)).encode()
            body = [ln for ln in text.replace(b"\r\n", b"\n").replace(b"\r", b"\n").split(b"\n") if ln.strip()]
            if all(ln.split()[0] == c0 for ln in body):
                assert stats.n_called == len({int(q) for q in poss} & {q for _, q in snps})
        if all_pos:
            lines = out[2]
            assert len(lines) == len(cells)
            assert np.array_equal(lines & 0xff, cells)
            assert np.array_equal(lines >> 8, fails)
        else:
            assert stats.n_parsed == len(cells)
        return stats
    finally:
        sites.close()


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("branch", ["", "_preserved"])
def test_lambda_golden(ctx, golden_dir, branch):
    """BASELINE config 1: the bundled lambda-virus pileups -> consensus.fasta, byte for byte."""
    from snp_pipeline_b200 import _lib
    root = os.path.join(golden_dir, "lambda")
    snps = orc.read_snp_list(os.path.join(root, "snplist%s.txt" % branch))
    p = _lib.make_params(min_cons_depth=3)
    for s in ("sample1", "sample2", "sample3", "sample4"):
        sdir = os.path.join(root, "samples", s)
        text = open(os.path.join(sdir, "reads.all.pileup"), "rb").read()
        excl = orc.vcf_sites(os.path.join(sdir, "var.flt_removed.vcf")) if branch else []
        sites = ctx.sites(snps, excl)
        golden = open(os.path.join(sdir, "consensus%s.fasta" % branch)).read()
        for mode in (_lib.MODE_SITES, _lib.MODE_ALL):
            row, stats = ctx.pileup_consensus(text, sites, p, mode)[:2]
            assert orc.fasta_text(s, row.decode()) == golden
            assert stats.n_lines == 48502
            assert stats.n_general < 200, "lambda lines should stay on the fast path"
        sites.close()
        _compare(ctx, text, snps, excl, (0, 0.6, 3, 0, 0.0), True)


@pytest.mark.parametrize("seed", range(6))
def test_realistic_text(ctx, seed):
    rng = random.Random(seed)
    n = 20000 if seed < 2 else 3000
    sites = {p: rng.choice("ACGT") for p in rng.sample(range(1, n + 1), 200)}
    text = linegen.pileup_text(seed, n, sites=sites, gaps=0.01).encode()
    snps = [(linegen.CHROM, p) for p in sorted(rng.sample(range(1, n + 40), 300))]
    excl = [(linegen.CHROM, p) for p in rng.sample(range(1, n), 40)] if seed % 2 else []
    ps = PARAM_SETS[seed % len(PARAM_SETS)]
    for all_pos in (False, True):
        st = _compare(ctx, text, snps, excl, ps, all_pos)
        assert st.n_general < st.n_lines * 0.02 + 5


def test_fp64_threshold_edges(ctx):
    """SURVEY section 7, "fp64 threshold compare": 100 * 0.55 = 55.00000000000001 in IEEE double, so 55 of 100 fails
    VarFreq while 56 passes -- through the first tier's threshold tables ('.' / ',' lines: the reference base wins) and
    through the exact tiers (a written-out allele wins); and the strand-bias product 25 * 0.28 = 7.000000000000001: 7
    reads on the weaker strand fail StrBias, 8 pass."""
    from snp_pipeline_b200 import _lib
    assert 100 * 0.55 > 55 and 25 * 0.28 > 7
    chrom = linegen.CHROM
    q = "I" * 100
    lines = [
        "%s\t1\tA\t100\t%s\t%s" % (chrom, "G" * 30 + "g" * 25 + "." * 45, q),     # G 55 of 100: fails at 0.55 (exact tiers)
        "%s\t2\tA\t100\t%s\t%s" % (chrom, "G" * 30 + "g" * 26 + "." * 44, q),     # G 56 of 100: passes
        "%s\t3\tA\t100\t%s\t%s" % (chrom, "." * 30 + "," * 25 + "G" * 45, q),     # ref 55 of 100: fails (first tier's table)
        "%s\t4\tA\t100\t%s\t%s" % (chrom, "." * 30 + "," * 26 + "G" * 44, q),     # ref 56 of 100: passes
        "%s\t5\tC\t25\t%s\t%s" % (chrom, "." * 7 + "," * 18, "I" * 25),            # strand bias: 7 < 25 * 0.28
        "%s\t6\tC\t26\t%s\t%s" % (chrom, "." * 8 + "," * 17 + "$", "I" * 25),      # 8 of 25: passes
    ]
    text = ("\n".join(lines) + "\n").encode()
    snps = [(chrom, k) for k in range(1, 7)]
    for ps in ((0, 0.55, 1, 0, 0.0), (0, 0.55, 1, 0, 0.28)):
        for all_pos in (False, True):
            _compare(ctx, text, snps, [], ps, all_pos)
    sites = ctx.sites(snps)
    row, _, per_line = ctx.pileup_consensus(text, sites, _lib.make_params(0, 0.55, 1, 0, 0.0), _lib.MODE_ALL, want_lines=True)
    assert row == b"-G-A" + b"CC"
    assert [int(v) >> 8 for v in per_line] == [_lib.FAIL_VARFREQ, 0, _lib.FAIL_VARFREQ, 0, 0, 0]
    row, _, per_line = ctx.pileup_consensus(text, sites, _lib.make_params(0, 0.5, 1, 0, 0.28), _lib.MODE_ALL, want_lines=True)
    assert [int(v) >> 8 for v in per_line][4:] == [_lib.FAIL_STRBIAS, 0] and row[4:] == b"-C"
    sites.close()


@pytest.mark.parametrize("seed", range(3))
def test_indel_token_corners(ctx, seed):
    """Well- and ill-formed indel tokens at every alignment against the column's end (tests/linegen.py): the lines the
    reference parses in one text, a few of those it stops at (a bare sign: int("")) alone and behind good lines."""
    rng = random.Random(9100 + seed)
    n = 900
    lines = linegen.indel_corner_lines(rng, n)
    ps = PARAM_SETS[seed % len(PARAM_SETS)]
    op = orc.make_params(*ps)
    good, bad = [], []
    for k, line in enumerate(lines):
        try:
            orc.pileup_consensus(line.encode(), [], [], op, parse_all=True, want_lines=True)
            good.append(line)
        except orc.OracleError:
            bad.append((k, line))
    assert len(good) > n // 2 and len(bad) > 10
    snps = [(linegen.CHROM, p) for p in sorted(rng.sample(range(1, n + 1), 200))]
    text = "".join(good).encode()
    for all_pos in (False, True):
        st = _compare(ctx, text, snps, [], ps, all_pos)
        assert st.n_lines == len(good)
    for k, line in bad[:4]:
        assert _compare(ctx, line.encode(), [(linegen.CHROM, 1 + k)], [], ps, True) is None
        assert _compare(ctx, ("".join(good[:300]) + line + "".join(good[300:400])).encode(), snps, [], ps, True) is None


@pytest.mark.parametrize("seed", range(3))
def test_header_shapes(ctx, seed):
    """Contig names of every length, positions of 1..10 digits, depths of 1..4 digits (the word-wise header parse)."""
    from test_cpu_sim import header_shapes_text
    text, snps = header_shapes_text(seed)
    for all_pos in (False, True):
        _compare(ctx, text, snps, [], PARAM_SETS[1], all_pos)
    # one contig over many lines: the first-tier parser takes them
    rng = random.Random(seed)
    for chrom in ("c", "chr%d" % seed, "x" * 31, "y" * 32, "NZ_" + "7" * 30, "z" * 63):
        body = "".join(linegen.realistic_line(rng, 99999990 + k, chrom) for k in range(400))
        snps1 = [(chrom, 99999990 + k) for k in range(0, 400, 7)]
        for all_pos in (False, True):
            st = _compare(ctx, body.encode(), snps1, [], PARAM_SETS[seed % 4], all_pos)
            assert st.n_general <= 1          # (the last line of a text is left to the any-input parser)


@pytest.mark.parametrize("seed", range(10))
def test_nasty_text(ctx, seed):
    """Legal-but-odd lines (caret chains, odd indel tokens, IUPAC, short quality strings, odd separators, lines the
    reference raises on): values, error codes and error offsets."""
    rng = random.Random(100 + seed)
    n = 900
    text = linegen.pileup_text(200 + seed, n, nasty=0.5).encode()
    if seed % 2 == 0:      # keep half of the seeds on the value path: drop the lines the reference raises on
        op = orc.make_params()
        text = b"".join(ln + b"\n" for ln in text.split(b"\n")[:-1] if orc.line_report(ln, op)["status"] == 0)
    if seed % 3 == 0:
        text = text.replace(b"\r\n", b"\n").replace(b"\n", b"\r\n")
    if seed % 4 == 1:
        text = text[:-1]
    snps = [(linegen.CHROM, p) for p in rng.sample(range(1, n + 10), 300)]
    excl = [(linegen.CHROM, p) for p in rng.sample(range(1, n), 30)]
    for ps in (PARAM_SETS[seed % len(PARAM_SETS)], PARAM_SETS[(seed + 2) % len(PARAM_SETS)]):
        for all_pos in (False, True):
            _compare(ctx, text, snps, excl, ps, all_pos)


@pytest.mark.parametrize("seed", range(2))
def test_zero_depth_and_odd_bytes(ctx, seed):
    """Lines of raw depth 0 in every shape (the pileup kernel's first tier calls them itself unless something odd follows
    the depth column) and lines with one byte >= 0x80 / control byte in some column, each between ordinary lines."""
    import test_cpu_sim as tcs
    rng = random.Random(40 + seed)
    lines = [linegen.realistic_line(rng, 1 + k, indel_rate=0.03).encode("latin-1") for k in range(400)]
    snps = [(linegen.CHROM, 1 + k) for k in range(0, 400, 3)]
    clean = []
    for k in range(0, 400, 4):
        if k % 8:
            tail = rng.choice(tcs.ZERO_TAILS)
            ln = ("%s\t%d\t%s\t%s%s%s\n" % (linegen.CHROM, 1 + k, rng.choice("ACGTNacgtn"), rng.choice(["0", "00", "000"]),
                                           rng.choice(["\t", "\t", " "]), tail)).encode("latin-1")
        else:
            f = lines[k].split(b"\t")
            col = rng.choice([4, 4, 5, 1, 0])
            b = bytearray(f[col])
            b[rng.randrange(len(b))] = rng.choice([0x80, 0xa0, 0xa1, 0xff, 0x7f, 0x20, 0x1f, 0x0b, 0x00])
            f[col] = bytes(b)
            ln = b"\t".join(f)
        try:                                                   # lines the reference takes stay in the long text as well
            orc.pileup_consensus(ln, [], [], orc.make_params(), parse_all=True, want_lines=True)
            lines[k] = ln
            clean.append(k)
        except orc.OracleError:
            pass
        for all_pos in (True, False):                          # ... and every line alone between two ordinary ones
            _compare(ctx, lines[k - 1] + ln + lines[k + 1] if k else ln + lines[1], snps, [], PARAM_SETS[k % len(PARAM_SETS)], all_pos)
    assert len(clean) > 20
    for all_pos in (True, False):
        _compare(ctx, b"".join(lines), snps, [], PARAM_SETS[1], all_pos)


def test_reference_file_vectors(ctx, ref_files):
    """Vectors produced by the reference's own `call_consensus` (tests/golden/make_golden.py)."""
    from snp_pipeline_b200 import _lib
    exc = {_lib.E_VALUE: "ValueError", _lib.E_INDEX: "IndexError", _lib.E_UNPACK: "ValueError"}
    n_ok = n_raise = 0
    for case in ref_files:
        snps = [(ln.split()[0], int(ln.split()[1])) for ln in case["snplist"].splitlines()]
        excl = []
        if case["exclude"]:
            for ln in case["exclude"].splitlines():
                if not ln.startswith("#"):
                    f = ln.split("\t")
                    excl.append((f[0], int(f[1])))
        sites = ctx.sites(snps, excl)
        mode = _lib.MODE_ALL if case["all_pos"] else _lib.MODE_SITES
        ref = case["ref"]
        try:
            row = ctx.pileup_consensus(case["pileup"].encode(), sites, _gpu_params(case["params"]), mode)[0]
            assert ref["exit"] == 0
            assert orc.fasta_text("sampleX", row.decode()) == ref["fasta"]
            n_ok += 1
        except _lib.SnpGpuError as e:
            assert exc[e.code] == str(ref["exit"]).split(":")[1]
            n_raise += 1
        sites.close()
    assert n_ok >= 20 and n_raise >= 20


def test_edge_inputs(ctx):
    from snp_pipeline_b200 import _lib
    p = _lib.make_params()
    snps = [(linegen.CHROM, 5), (linegen.CHROM, 7), ("other", 1)]
    sites = ctx.sites(snps)
    # empty file: every cell is '-'
    row, stats = ctx.pileup_consensus(b"", sites, p)
    assert row == b"---" and stats.n_lines == 0
    # a single line without terminator
    line = ("%s\t5\tA\t3\tGGg\tIII" % linegen.CHROM).encode()
    row, stats = ctx.pileup_consensus(line, sites, p)
    assert row == b"G--" and stats.n_lines == 1
    # empty snplist is fine (regression_tests.sh:3156-3211)
    empty = ctx.sites([])
    row, stats = ctx.pileup_consensus(line + b"\n", empty, p)
    assert row == b""
    row, stats, lines = ctx.pileup_consensus(line + b"\n", empty, p, _lib.MODE_ALL, want_lines=True)
    assert len(lines) == 1 and chr(lines[0] & 0xff) == "G"
    empty.close()
    # a very deep line (longer than the kernel's staging look-ahead) goes through the exact path
    deep = ("%s\t7\tC\t9000\t%s\t%s\n" % (linegen.CHROM, "." * 4000 + "t" * 5000, "I" * 9000)).encode()
    text = line + b"\n" + deep
    _compare(ctx, text, snps, [], (0, 0.55, 1, 0, 0.0), True)
    _compare(ctx, text, snps, [], (0, 0.55, 1, 0, 0.0), False)
    # many tiny lines in one tile (more line starts than one scan pass records)
    tiny = b"".join(b"c %d\n" % i for i in range(30000))
    row, stats = ctx.pileup_consensus(tiny, sites, p)
    assert row == b"---" and stats.n_lines == 30000
    # classic-Mac line ends: universal newlines (pileup.py:417)
    mac = linegen.pileup_text(3, 400).replace("\n", "\r").encode()
    _compare(ctx, mac, [(linegen.CHROM, q) for q in range(1, 400, 7)], [], (0, 0.6, 1, 0, 0.0), False)
    _compare(ctx, mac, [(linegen.CHROM, q) for q in range(1, 400, 7)], [], (0, 0.6, 1, 0, 0.0), True)
    # non-ASCII byte: outside the byte domain, reported at its line
    bad = linegen.pileup_text(4, 50).encode()
    offs = _line_offsets(bad)
    bad2 = bad[:offs[20] + 3] + b"\xc3" + bad[offs[20] + 4:]
    with pytest.raises(_lib.SnpGpuError) as ei:
        ctx.pileup_consensus(bad2, sites, p)
    assert ei.value.code == _lib.E_DOMAIN and ei.value.offset == offs[20]
    sites.close()


def test_multi_contig_and_duplicates(ctx):
    rng = random.Random(11)
    chroms = ["chrB|x", "chrA", "chr10"]
    parts, allpos = [], []
    for c in chroms:
        n = 2500
        parts.append(linegen.pileup_text(len(c), n, chrom=c, gaps=0.02))
        allpos += [(c, q) for q in range(1, n + 100)]
    text = "".join(parts)
    lines = text.splitlines(True)
    text = "".join(lines + lines[100:140][::-1]).encode()       # repeated positions: the last line wins
    snps = rng.sample(allpos, 400)
    snps = snps + snps[:7]                                      # duplicates in the snplist are emitted twice
    excl = rng.sample(allpos, 60)
    for all_pos in (False, True):
        _compare(ctx, text, snps, excl, (0, 0.6, 2, 0, 0.0), all_pos)


def _synth(ctx, genome_len, sample, pool, carry=0.05, seed=20261017):
    import torch
    from snp_pipeline_b200 import _lib
    spec = _lib.SynthSpec(seed, sample, genome_len, 24, pool, carry, 0.0)
    cap = int(genome_len) * 112 + 4096
    buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
    n = ctx.synth_pileup_dev(spec, "gi|0000000|ref|SYN_5000K.1|", buf.data_ptr(), cap)
    return spec, buf, n


@pytest.mark.parametrize("genome_len,sample,indel", [(150000, 0, 0.0), (60000, 5, 0.05)])
def test_synthetic_text_host_generator(ctx, genome_len, sample, indel):
    """The reference arm of bench.py parses inputs written on the host (oracle/synth_host.cpp, the generator's own line
    function compiled with g++): they are the bytes the device generator writes, and so are the carried sites."""
    from snp_pipeline_b200 import _lib
    import torch
    spec = _lib.SynthSpec(20261017, sample, genome_len, 24, genome_len // 100, 0.05, indel)
    cap = genome_len * 112 + 4096
    buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
    chrom = "gi|0000000|ref|SYN_5000K.1|"
    n = ctx.synth_pileup_dev(spec, chrom, buf.data_ptr(), cap)
    host = orc.synth_pileup(20261017, sample, genome_len, 24, genome_len // 100, 0.05, indel, chrom, threads=3)
    assert host.size == n and np.array_equal(host, buf[:n].cpu().numpy())
    assert np.array_equal(orc.synth_sample_sites(20261017, sample, genome_len, 24, genome_len // 100, 0.05),
                          ctx.synth_sample_sites(spec))


@pytest.mark.parametrize("genome_len,sample", [(200000, 0), (5000000, 3)])
def test_synthetic_sample_device_resident(ctx, genome_len, sample):
    """BASELINE config 2's unit of work: one synthetic sample (5 Mbp at the full size), text resident in HBM,
    through the device-pointer entry point; every line's cell and fail mask against the oracle."""
    import torch
    from snp_pipeline_b200 import _lib
    spec, buf, n = _synth(ctx, genome_len, sample, pool=genome_len // 100)
    assert 80 * genome_len < n < 100 * genome_len
    chrom = "gi|0000000|ref|SYN_5000K.1|"
    own = ctx.synth_sample_sites(spec)
    assert len(own) > 0
    rng = random.Random(sample)
    extra = rng.sample(range(1, genome_len + 1), 2000)
    snps = sorted({(chrom, int(q)) for q in own} | {(chrom, q) for q in extra}, key=lambda t: t[1])
    sites = ctx.sites(snps)
    p = _lib.make_params(min_cons_depth=3)
    row = torch.empty(len(snps), dtype=torch.uint8, device="cuda")
    lines = torch.empty(genome_len + 8, dtype=torch.int16, device="cuda")
    stats = torch.zeros(6, dtype=torch.int64, device="cuda")
    text = buf[:n].cpu().numpy()
    op = orc.make_params(min_cons_depth=3)
    want_row, (cells, fails, _) = orc.pileup_consensus(text, snps, [], op, parse_all=True, want_lines=True)
    for mode in (_lib.MODE_ALL, _lib.MODE_SITES):
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        ctx.pileup_consensus_dev(buf.data_ptr(), n, sites, p, mode, row.data_ptr(), lines.data_ptr(), genome_len + 8,
                                 stats.data_ptr())
        torch.cuda.synchronize()
        st = stats.cpu().numpy()
        assert st[0] == genome_len and st[4] == 0
        assert row.cpu().numpy().tobytes() == want_row
        if mode == _lib.MODE_ALL:
            got = lines[:genome_len].cpu().numpy().view(np.uint16)
            assert np.array_equal(got & 0xff, cells)
            assert np.array_equal(got >> 8, fails)
            assert st[2] < genome_len * 0.01      # indel-odd lines only
        else:
            assert st[1] == len(snps)
    # the variant cells really are variant: most carried sites call the alternate allele
    own_set = set(int(x) for x in own)
    called = sum(1 for (c, q), b in zip(snps, want_row) if q in own_set and chr(b) in "ACGT")
    assert called > 0.8 * len(own)
    ctx.set_stream(None)
    sites.close()


@pytest.mark.parametrize("mode_all", [True, False])
def test_batch_of_samples_one_launch(ctx, mode_all):
    """snpgpu_pileup_consensus_batch_dev: 19 samples of different lengths (one of them empty) -> two launch sequences;
    every sample's row, per-line results and stats against the oracle run on that sample alone."""
    import torch
    from snp_pipeline_b200 import _lib
    chrom = "gi|0000000|ref|SYN_5000K.1|"
    lens = [30000 + 1777 * i for i in range(18)]
    specs, bufs, ns = [], [], []
    for i, g in enumerate(lens):
        spec, buf, n = _synth(ctx, g, 40 + i, pool=max(g // 50, 1))
        specs.append(spec); bufs.append(buf); ns.append(n)
    bufs.insert(7, torch.empty(64, dtype=torch.uint8, device="cuda")); ns.insert(7, 0); lens.insert(7, 0); specs.insert(7, None)
    snp_pos = sorted({int(q) for sp in specs if sp is not None for q in ctx.synth_sample_sites(sp)} | set(range(5, 60000, 997)))
    snps = [(chrom, q) for q in snp_pos]
    sites = ctx.sites(snps)
    p = _lib.make_params(min_cons_depth=3)
    op = orc.make_params(min_cons_depth=3)
    B = len(bufs)
    rows = torch.zeros((B, len(snps)), dtype=torch.uint8, device="cuda")
    cap = max(lens) + 8
    lines = torch.zeros((B, cap), dtype=torch.int16, device="cuda")
    stats = torch.zeros((B, 6), dtype=torch.int64, device="cuda")
    mode = _lib.MODE_ALL if mode_all else _lib.MODE_SITES
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.pileup_consensus_batch_dev([(bufs[i].data_ptr(), ns[i], rows[i].data_ptr(), lines[i].data_ptr() if mode_all else 0,
                                     cap if mode_all else 0, stats[i].data_ptr()) for i in range(B)], sites, p, mode)
    torch.cuda.synchronize()
    st = stats.cpu().numpy()
    rows_h, lines_h = rows.cpu().numpy(), lines.cpu().numpy().view(np.uint16)
    for i in range(B):
        text = bufs[i][:ns[i]].cpu().numpy()
        want_row, (cells, fails, _) = orc.pileup_consensus(text, snps, [], op, parse_all=mode_all, want_lines=True)
        assert st[i, 0] == lens[i] and st[i, 4] == 0, (i, st[i])
        assert rows_h[i].tobytes() == want_row, i
        if mode_all:
            assert np.array_equal(lines_h[i, :lens[i]] & 0xff, cells), i
            assert np.array_equal(lines_h[i, :lens[i]] >> 8, fails), i
    ctx.set_stream(None)
    sites.close()


# ------------------------------------------------------------------------------------------ K2
def test_pipelined_host_calls(ctx):
    """snpgpu_pileup_consensus_begin / _end with one call kept ahead: the same rows, per-line calls and stats as the plain
    call, also when a sample needs the redo paths (classic-Mac line ends) or raises."""
    from snp_pipeline_b200 import _lib
    rng = random.Random(11)
    n = 1500
    snps = [(linegen.CHROM, p) for p in sorted(rng.sample(range(1, n + 1), 120))]
    sites = ctx.sites(snps)
    p = _lib.make_params(min_cons_depth=3)
    texts = [linegen.pileup_text(60 + k, n).encode() for k in range(5)]
    texts[2] = texts[2].replace(b"\n", b"\r")                               # lone CRs: normalised and redone inside _end
    bad = texts[3].split(b"\n")
    bad[700] = bad[700].replace(b"\t", b"\tx", 1).replace(b"\tx", b"\t", 0)
    f = bad[700].split(b"\t"); f[1] = b"12x"; bad[700] = b"\t".join(f)     # int("12x") raises in the reference
    texts[3] = b"\n".join(bad)
    want = []
    for t in texts:
        try:
            want.append(ctx.pileup_consensus(t, sites, p, _lib.MODE_ALL, want_lines=True))
        except _lib.SnpGpuError as e:
            want.append(e)
    arrs = [np.frombuffer(t, dtype=np.uint8) for t in texts]
    rows = [np.zeros(len(snps), np.uint8) for _ in texts]
    lines = [np.zeros(n + 8, np.uint16) for _ in texts]
    stats = [_lib.PileupStats() for _ in texts]
    got_err = {}
    prev = None
    for k in range(len(texts) + 1):
        cur = None
        if k < len(texts):
            cur = (ctx.pileup_consensus_begin(arrs[k], sites, p, _lib.MODE_ALL, rows[k], lines[k], stats[k]), k)
        if prev is not None:
            try:
                ctx.pileup_consensus_end(prev[0])
            except _lib.SnpGpuError as e:
                got_err[prev[1]] = e
        prev = cur
    for k, w in enumerate(want):
        if isinstance(w, _lib.SnpGpuError):
            assert k in got_err and got_err[k].code == w.code and stats[k].error_offset == w.offset
            continue
        assert k not in got_err
        assert rows[k].tobytes() == w[0]
        assert stats[k].n_lines == w[1].n_lines and stats[k].n_parsed == w[1].n_parsed
        assert np.array_equal(lines[k][:stats[k].n_lines], w[2])
    with pytest.raises(_lib.SnpGpuError):
        ctx.pileup_consensus_end(0)                                          # nothing in flight
    sites.close()


def test_tiny_lines_overflow_the_staging_rows(ctx):
    """Lines of a few bytes: a 10.5 KiB tile then owns far more lines than its row of the per-line staging array holds
    (the overflow list, and the call growing it on the second attempt); per-line results still come out in file order."""
    from snp_pipeline_b200 import _lib
    rng = random.Random(5)
    c = "c"
    body = []
    for k in range(1, 30001):                                  # "c\t<k>\tA\t0\n": 8-12 bytes, > 1000 lines per tile
        body.append("%s\t%d\t%s\t0\n" % (c, k, rng.choice("ACGT")))
    body += [linegen.realistic_line(rng, 30000 + k, c) for k in range(1, 400)]
    text = "".join(body).encode()
    snps = [(c, p) for p in sorted(rng.sample(range(1, 30400), 300))]
    for all_pos in (False, True):
        _compare(ctx, text, snps, [], PARAM_SETS[1], all_pos)
        _compare(ctx, text, snps, [], PARAM_SETS[1], all_pos)  # (second call: the list has its grown size from the start)


def test_site_table_built_on_device(ctx):
    """snpgpu_sites_create_from_keys_dev (K2's keys -> site table, no host round trip) against the host-built table:
    same consensus rows, same per-line calls, on two contigs."""
    import torch
    from snp_pipeline_b200 import _lib
    rng = random.Random(77)
    contigs = ["chrA|1|", "chrB.22"]                       # chromosome-string order = rank order
    lens = [3000, 1200]
    text = (linegen.pileup_text(5, 3000, chrom=contigs[0]) + linegen.pileup_text(6, 1200, chrom=contigs[1])).encode()
    snps = sorted([(0, p) for p in rng.sample(range(1, 3001), 300)] + [(1, p) for p in rng.sample(range(1, 1201), 150)])
    keys = np.array([(c << 32) | p for c, p in snps], dtype=np.uint64)
    keys_dev = torch.from_numpy(keys.view(np.int64)).cuda()
    host = _lib.Sites.from_arrays(ctx, contigs, np.array([c for c, _ in snps], np.int32), np.array([p for _, p in snps], np.int64))
    p = _lib.make_params(min_cons_depth=3)
    for mode in (_lib.MODE_SITES, _lib.MODE_ALL):
        want = ctx.pileup_consensus(text, host, p, mode, want_lines=mode == _lib.MODE_ALL)
        for _ in range(3):                                 # (the second and third tables reuse the first one's memory)
            dev = _lib.Sites.from_keys_dev(ctx, contigs, lens, keys_dev.data_ptr(), keys.size)
            got = ctx.pileup_consensus(text, dev, p, mode, want_lines=mode == _lib.MODE_ALL)
            dev.close()
            assert got[0] == want[0]
            assert got[1].n_parsed == want[1].n_parsed
            if mode == _lib.MODE_ALL:
                assert np.array_equal(got[2], want[2])
    host.close()
    # keys beyond the contig lengths the caller gave (and on a contig it did not name) get no bit: their columns read '-'
    # and every site behind them still lands in its own column
    extra = sorted(snps + [(0, 3500), (0, 9000), (1, 1201), (1, 70000), (2, 5)])
    keys2 = np.array([(c << 32) | q for c, q in extra], dtype=np.uint64)
    keys2_dev = torch.from_numpy(keys2.view(np.int64)).cuda()
    dev = _lib.Sites.from_keys_dev(ctx, contigs, lens, keys2_dev.data_ptr(), keys2.size)
    got = ctx.pileup_consensus(text, dev, p, _lib.MODE_SITES)[0]
    dev.close()
    ref_row = dict(zip(snps, ctx.pileup_consensus(text, _lib.Sites.from_arrays(
        ctx, contigs, np.array([c for c, _ in snps], np.int32), np.array([q for _, q in snps], np.int64)), p, _lib.MODE_SITES)[0]))
    assert bytes(ref_row.get(k, ord("-")) for k in extra) == got
    empty = _lib.Sites.from_keys_dev(ctx, contigs, lens, 0, 0)
    row, stats = ctx.pileup_consensus(text, empty, p, _lib.MODE_SITES)[:2]
    assert row == b"" and stats.n_parsed == 0
    empty.close()


@pytest.mark.parametrize("dataset,vcf", [("lambda", "var.flt.vcf"), ("agona", "var.flt.vcf"),
                                         ("listeria", "var.flt.vcf")])
def test_merge_sites_golden(ctx, golden_dir, dataset, vcf):
    root = os.path.join(golden_dir, dataset)
    names = sorted(os.listdir(os.path.join(root, "samples")))
    per = [orc.vcf_sites(os.path.join(root, "samples", n, vcf)) for n in names]
    chroms = sorted({c for s in per for c, _ in s})
    rank = {c: i for i, c in enumerate(chroms)}
    keys = np.array([(rank[c] << 32) | p for s in per for c, p in s], dtype=np.uint64)
    samp = np.array([i for i, s in enumerate(per) for _ in s], dtype=np.uint32)
    uniq, cnt, samples = ctx.merge_sites(keys, samp)
    lines, o = [], 0
    for k, c in zip(uniq, cnt):
        who = [names[i] for i in samples[o:o + c]]
        o += int(c)
        lines.append("%s\t%d\t%d\t%s\n" % (chroms[int(k) >> 32], int(k) & 0xffffffff, c, "\t".join(who)))
    assert "".join(lines) == open(os.path.join(root, "snplist.txt")).read()


@pytest.mark.parametrize("n_samples,per_sample,seed", [(1, 1, 0), (3, 0, 1), (100, 500, 2), (1000, 5000, 3)])
def test_merge_sites_random(ctx, n_samples, per_sample, seed):
    rng = np.random.default_rng(seed)
    keys, samp = [], []
    for s in range(n_samples):
        k = rng.choice(200000, size=per_sample, replace=False).astype(np.uint64) if per_sample else np.zeros(0, np.uint64)
        k |= rng.integers(0, 3, size=k.size).astype(np.uint64) << np.uint64(32)
        k = np.unique(k)
        rng.shuffle(k)
        keys.append(k)
        samp.append(np.full(k.size, s, dtype=np.uint32))
    keys, samp = np.concatenate(keys), np.concatenate(samp)
    wu, wc, ws = orc.merge_sites_keys(keys, samp)
    gu, gc, gs = ctx.merge_sites(keys, samp)
    assert np.array_equal(wu, gu) and np.array_equal(wc, gc) and np.array_equal(ws, gs)
    assert np.all(gu[1:] > gu[:-1])                              # sortedness, uniqueness
    assert int(gc.sum()) == keys.size


def test_merge_sites_key_bytes(ctx):
    """K2's radix passes follow the key bytes that carry information: all-zero keys (no pass at all), positions up to
    2^31 - 1 with chromosome ranks up to 70 000 (seven key bytes), one key repeated by every sample."""
    rng = np.random.default_rng(12)
    cases = [
        (np.zeros(5000, np.uint64), np.arange(5000, dtype=np.uint32) % 7),
        ((rng.integers(0, 70000, 300000).astype(np.uint64) << np.uint64(32)) | rng.integers(0, 2 ** 31, 300000).astype(np.uint64),
         np.sort(rng.integers(0, 900, 300000)).astype(np.uint32)),
        (np.full(4097, (3 << 32) | 77, np.uint64), np.arange(4097, dtype=np.uint32)),
    ]
    for keys, samp in cases:
        wu, wc, ws = orc.merge_sites_keys(keys, samp)
        gu, gc, gs = ctx.merge_sites(keys, samp)
        assert np.array_equal(wu, gu) and np.array_equal(wc, gc) and np.array_equal(ws, gs)


# ------------------------------------------------------------------------------------------ K4
@pytest.mark.parametrize("dataset,suffix", [("lambda", ""), ("lambda", "_preserved"), ("agona", ""), ("listeria", ""),
                                            ("listeria", "_preserved")])
def test_distance_golden(ctx, golden_dir, dataset, suffix):
    root = os.path.join(golden_dir, dataset)
    seqs = orc.read_fasta_matrix(os.path.join(root, "snpma%s.fasta" % suffix))
    ids = sorted(seqs)
    m = np.frombuffer("".join(seqs[i] for i in ids).encode(), dtype=np.uint8).reshape(len(ids), -1)
    d = ctx.pairwise_distance(m)
    mat = ["\t%s\n" % "\t".join(ids)]
    for i, a in enumerate(ids):
        mat.append("%s\t%s\n" % (a, "\t".join(str(int(x)) for x in d[i])))
    assert "".join(mat) == open(os.path.join(root, "snp_distance_matrix%s.tsv" % suffix)).read()


@pytest.mark.parametrize("n,s,seed", [(1, 10, 0), (2, 0, 1), (2, 1, 2), (5, 31, 3), (7, 32, 4), (64, 33, 5), (65, 1025, 6),
                                      (130, 3000, 7), (300, 20011, 8)])
def test_distance_random(ctx, n, s, seed):
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTACGTACGTacgt-NnRY*.", dtype=np.uint8)
    m = alphabet[rng.integers(0, alphabet.size, size=(n, s))] if s else np.zeros((n, 0), np.uint8)
    if s and n > 2:      # related rows so that distances are not all ~ 3/4 s
        m[1] = m[0]
        m[2, : s // 2] = m[0, : s // 2]
    want = orc.distance_matrix([bytes(r) for r in m]) if s else np.zeros((n, n), np.int32)
    got = ctx.pairwise_distance(m)
    assert np.array_equal(got, want)
    assert np.array_equal(got, got.T) and not got.diagonal().any()


@pytest.mark.parametrize("n,s,stride", [(70, 4097, 4160), (3, 100, 128), (200, 64, 64), (129, 20000, 20032)])
def test_distance_aligned_rows_and_tile_list(ctx, n, s, stride):
    """16-byte aligned rows (K4's coalesced pack kernel), few tiles (the words are dealt to several CTAs), and the
    tile-row form mirrored into the full matrix -- against the oracle."""
    import torch
    rng = np.random.default_rng(n + s)
    alphabet = np.frombuffer(b"ACGTACGTacgt-NnRY*.", dtype=np.uint8)
    m = np.full((n, stride), ord("A"), dtype=np.uint8)                      # (the padding must not count)
    m[:, :s] = alphabet[rng.integers(0, alphabet.size, size=(n, s))]
    want = orc.distance_matrix([bytes(r[:s]) for r in m])
    md = torch.from_numpy(m).cuda()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        out = torch.zeros((n, n), dtype=torch.int32, device="cuda")
        ctx.pairwise_distance_dev(md.data_ptr(), n, s, stride, 0, n, out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want)
        n_tiles = (n + 63) // 64
        tiles = list(range(n_tiles))[::-1]                                   # any order
        part = torch.zeros((n_tiles * 64, n), dtype=torch.int32, device="cuda")
        ctx.pairwise_distance_tiles_dev(md.data_ptr(), n, s, stride, tiles, part.data_ptr())
        torch.cuda.synchronize()
        upper = torch.zeros((n_tiles, 64, n), dtype=torch.int32, device="cuda")
        upper[torch.tensor(tiles, device="cuda")] = part.view(n_tiles, 64, n)
        upper = upper.view(n_tiles * 64, n)[:n]
        row = torch.arange(n, device="cuda")
        full = torch.where(row[None, :] >= (row // 64 * 64)[:, None], upper, upper.t())
        assert np.array_equal(full.cpu().numpy(), want)
    finally:
        ctx.set_stream(None)


def test_distance_stripes(ctx):
    """The multi-GPU sharding of K4: row stripes computed separately equal the full matrix."""
    import torch
    rng = np.random.default_rng(5)
    n, s = 203, 5000
    m = np.frombuffer(b"ACGT-N", dtype=np.uint8)[rng.integers(0, 6, size=(n, s))]
    full = ctx.pairwise_distance(m)
    md = torch.from_numpy(m).cuda()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    for lo, hi in [(0, 70), (70, 128), (128, 203), (13, 14)]:
        out = torch.zeros((hi - lo, n), dtype=torch.int32, device="cuda")
        ctx.pairwise_distance_dev(md.data_ptr(), n, s, s, lo, hi, out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), full[lo:hi])
    ctx.set_stream(None)


def test_distance_full_scale_stripe(ctx):
    """BASELINE configs[4] (5000 samples, ~200 k sites, 12.5 M pairs over 8 GPUs): the stripe one of eight ranks computes,
    at full size, checked through size-independent properties -- zero diagonal, symmetry between two ranks' stripes,
    and sampled pairs against numpy on the same rows."""
    import torch
    n, s, per = 5000, 200_000, 625
    g = torch.Generator(device="cuda").manual_seed(20261017)
    alphabet = torch.tensor(list(b"ACGTACGTACGT-Nacgt"), dtype=torch.uint8, device="cuda")
    clade = alphabet[torch.randint(0, alphabet.numel(), (8, s), generator=g, device="cuda")]       # eight founders ...
    m = clade[torch.randint(0, 8, (n,), generator=g, device="cuda")].clone()                        # ... and their offspring
    flip = torch.rand((n, s), generator=g, device="cuda") < 0.01
    m[flip] = alphabet[torch.randint(0, alphabet.numel(), (int(flip.sum().item()),), generator=g, device="cuda")]
    del flip
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        stripes = []
        for r in (0, 3):
            out = torch.empty((per, n), dtype=torch.int32, device="cuda")
            ctx.pairwise_distance_dev(m.data_ptr(), n, s, s, r * per, (r + 1) * per, out.data_ptr())
            torch.cuda.synchronize()
            stripes.append(out)
        a, b = stripes
        assert not torch.diagonal(a[:, :per]).any() and not torch.diagonal(b[:, 3 * per:4 * per]).any()
        assert torch.equal(a[:, 3 * per:4 * per], b[:, :per].T)                  # rank 0's view of rank 3 == rank 3's of rank 0
        assert torch.equal(a[:, :per], a[:, :per].T)
        # a sampled 64 x 64 block of pairs against the oracle (utils.py:1135-1165 restated in C) on the same rows
        rng = random.Random(3)
        ri = sorted(rng.sample(range(per), 64))
        rj = sorted(rng.sample(range(n), 64))
        rows = [bytes(r) for r in m[torch.tensor(ri + rj, device="cuda")].cpu().numpy()]
        want = orc.distance_matrix(rows)[:64, 64:]
        got = a[torch.tensor(ri, device="cuda")][:, torch.tensor(rj, device="cuda")].cpu().numpy()
        assert np.array_equal(got, want)
        # the tile-row form the multi-GPU driver uses (upper part only, zigzag share of rank 5 of 8) against the stripes
        from snp_pipeline_b200 import sharding
        share = sharding.zigzag_tile_rows(n, 5, 8)
        assert sorted(t for r in range(8) for t in sharding.zigzag_tile_rows(n, r, 8)) == list(range((n + 63) // 64))
        part = torch.full((len(share) * 64, n), -7, dtype=torch.int32, device="cuda")
        ctx.pairwise_distance_tiles_dev(m.data_ptr(), n, s, s, share, part.data_ptr())
        torch.cuda.synchronize()
        for k, t in enumerate(share):
            if t * 64 >= 3 * per and t * 64 + 64 <= 4 * per:                  # a tile row inside rank 3's stripe
                lo = t * 64
                assert torch.equal(part[k * 64:(k + 1) * 64, lo:], b[lo - 3 * per: lo - 3 * per + 64, lo:])
                left = part[k * 64:(k + 1) * 64, :lo]
                assert bool(((left == -7) | (left == 0)).all())               # left of the diagonal tile: untouched or zeroed
    finally:
        ctx.set_stream(None)
