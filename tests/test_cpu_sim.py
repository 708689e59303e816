"""The per-line parsers the kernel runs (csrc/line_fast.cuh, csrc/line_general.cuh), compiled for the host by
csrc/cpu_sim.cpp, against the oracle -- the part of the CUDA path that can be checked without a GPU."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import linegen
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "snp_pipeline_b200", "csrc")
SO = os.path.join(ROOT, "tests", "_build", "libcpusim.so")


class CallParams(ctypes.Structure):
    _fields_ = [("min_base_qual", ctypes.c_int32), ("min_cons_depth", ctypes.c_int32),
                ("min_cons_strand_depth", ctypes.c_int32), ("pad", ctypes.c_int32),
                ("min_cons_freq", ctypes.c_double), ("min_cons_strand_bias", ctypes.c_double)]


@pytest.fixture(scope="module")
def sim():
    """The harness: line_quick3.cuh (k1_pileup.cu's first tier) in front, line_fast.cuh and line_general.cuh behind it."""
    subprocess.check_call(["make", "-s", "-C", CSRC, "cpusim"])
    L = ctypes.CDLL(SO)
    L.cpusim_pileup.restype = ctypes.c_int
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    L.cpusim_pileup.argtypes = [vp, sz, ctypes.c_char_p, vp, ctypes.c_int32, vp, vp, sz, vp, vp, sz,
                                ctypes.POINTER(CallParams), ctypes.c_int, ctypes.c_int, vp, vp, sz, vp]
    return L


def run_sim(L, text, snps, excl, ps, all_pos, force_general=False):
    buf = np.frombuffer(text, dtype=np.uint8)
    names = sorted({c for c, _ in snps} | {c for c, _ in excl})
    idx = {c: i for i, c in enumerate(names)}
    blob, off = orc.contig_table([n.encode() for n in names])
    sc = np.array([idx[c] for c, _ in snps], dtype=np.int32)
    sp = np.array([p for _, p in snps], dtype=np.int64)
    ec = np.array([idx[c] for c, _ in excl], dtype=np.int32)
    ep = np.array([p for _, p in excl], dtype=np.int64)
    p = CallParams(int(ps[0]), int(ps[2]), int(ps[3]), 0, float(ps[1]), float(ps[4]))
    row = np.zeros(max(len(snps), 1), dtype=np.uint8)
    cap = buf.size + 2
    lines = np.zeros(cap, dtype=np.uint16)
    counters = np.zeros(8, dtype=np.uint64)
    ptr = lambda a: ctypes.c_void_p(a.ctypes.data) if a.size else None
    rc = L.cpusim_pileup(ptr(buf), buf.size, blob, ptr(off), len(names), ptr(sc), ptr(sp), len(snps), ptr(ec), ptr(ep),
                         len(excl), ctypes.byref(p), 1 if all_pos else 0, 1 if force_general else 0, ptr(row),
                         ptr(lines), cap, ptr(counters))
    return rc, row[:len(snps)].tobytes(), lines[:int(counters[0])], counters


PARAM_SETS = [(0, 0.6, 1, 0, 0.0), (0, 0.6, 3, 0, 0.0), (15, 0.6, 3, 1, 0.1), (0, 0.55, 2, 2, 0.25),
              (20, 0.9, 10, 0, 0.5), (0, 1.0, 0, 0, 0.0), (-5, 0.51, 1, 3, 0.33)]


def _compare(L, text, snps, excl, ps, all_pos, force_general=False):
    op = orc.make_params(*ps)
    try:
        want_row, (cells, fails, _) = orc.pileup_consensus(text, snps, excl, op, parse_all=all_pos, want_lines=True)
        want_err = 0
    except orc.OracleError as e:
        want_err = e.status
    rc, row, lines, counters = run_sim(L, text, snps, excl, ps, all_pos, force_general)
    assert rc == want_err, (rc, want_err, counters)
    if want_err:
        return counters
    assert row == want_row
    if all_pos:
        assert len(lines) == len(cells)
        assert np.array_equal(lines & 0xff, cells)
        assert np.array_equal(lines >> 8, fails)
    return counters


@pytest.mark.parametrize("seed", range(8))
def test_realistic_text(sim, seed):
    rng = random.Random(seed)
    n = 3000
    sites = {p: rng.choice("ACGT") for p in rng.sample(range(1, n + 1), 120)}
    text = linegen.pileup_text(seed, n, sites=sites, gaps=0.01).encode()
    snps = [(linegen.CHROM, p) for p in sorted(rng.sample(range(1, n + 40), 150))]
    excl = [(linegen.CHROM, p) for p in rng.sample(range(1, n), 40)] if seed % 2 else []
    ps = PARAM_SETS[seed % len(PARAM_SETS)]
    for all_pos in (False, True):
        c = _compare(sim, text, snps, excl, ps, all_pos)
        assert c[2] < c[0] * 0.02 + 5, "the fast path should take nearly every samtools-shaped line"
        if ps[0] <= 0:
            assert c[5] > c[1] * 0.9, "the first-tier parser should decide most parsed lines (%d of %d)" % (c[5], c[1])
        _compare(sim, text, snps, excl, ps, all_pos, force_general=True)


@pytest.mark.parametrize("seed", range(12))
def test_nasty_text(sim, seed):
    rng = random.Random(100 + seed)
    n = 600
    text = linegen.pileup_text(200 + seed, n, nasty=0.5).encode()
    # drop the lines the reference raises on (odd integers): keep this test on the value path
    keep = []
    op = orc.make_params()
    for ln in text.split(b"\n")[:-1]:
        if orc.line_report(ln, op)["status"] == 0:
            keep.append(ln + b"\n")
    text = b"".join(keep)
    if seed % 3 == 0:
        text = text.replace(b"\r\n", b"\n").replace(b"\n", b"\r\n")   # no "\r\r\n": a lone CR is outside the domain
    if seed % 4 == 1:
        text = text[:-1]
    snps = [(linegen.CHROM, p) for p in rng.sample(range(1, n + 10), 200)]
    excl = [(linegen.CHROM, p) for p in rng.sample(range(1, n), 30)]
    for ps in (PARAM_SETS[seed % len(PARAM_SETS)], PARAM_SETS[(seed + 2) % len(PARAM_SETS)]):
        for all_pos in (False, True):
            _compare(sim, text, snps, excl, ps, all_pos)


def header_shapes_text(seed):
    """Contig names of every length 1..45 (all word alignments of the name compare), positions of 1..10 digits
    (with leading zeros), depths of 1..4 digits: the header columns the first-tier parser reads word-wise."""
    rng = random.Random(4000 + seed)
    lines, snps = [], []
    for k in range(1, 46):
        chrom = "".join(rng.choice("abcXYZ019|._-") for _ in range(k))
        for _ in range(6):
            nd = rng.randint(1, 10)
            pos = rng.randint(10 ** (nd - 1), 10 ** nd - 1) if nd < 10 else rng.randint(10 ** 9, 2 ** 31 - 1)
            depth = rng.choice([1, 2, 7, 9, 10, 35, 99, 100, 250, 999, 1000, 1500])
            ln = linegen.realistic_line(rng, pos, chrom, depth=depth)
            f = ln.split("\t")
            if rng.random() < 0.3:
                f[1] = "0" * rng.randint(1, 3) + f[1]
            if rng.random() < 0.2 and f[3] != "0":
                f[3] = "0" + f[3]
            lines.append("\t".join(f))
            if rng.random() < 0.6 and pos < 10 ** 7:          # (the site table is a bitmap over positions)
                snps.append((chrom, pos))
        if k % 7 == 0:                                   # a neighbour that differs only in its last byte / is a prefix
            lines.append(linegen.realistic_line(rng, 5, chrom[:-1] + "Q", depth=8))
            lines.append(linegen.realistic_line(rng, 6, chrom + "x", depth=8))
            snps += [(chrom[:-1] + "Q", 5), (chrom, 6)]
    return "".join(lines).encode(), snps


@pytest.mark.parametrize("seed", range(3))
def test_header_shapes(sim, seed):
    text, snps = header_shapes_text(seed)
    for all_pos in (False, True):
        c = _compare(sim, text, snps, [], PARAM_SETS[1], all_pos)
        if all_pos:
            assert c[5] > c[1] * 0.4, "the first-tier parser should decide many of these lines (%d of %d)" % (c[5], c[1])


def _boundary_line(rng, pos):
    """Lines around the first-tier parser's decision boundary (csrc/line_quick.cuh): reference base winning by a
    hair or tied, the reference letter written out, every kind of '^' partner, markers at word boundaries."""
    ref = rng.choice("ACGTNacgtn")
    depth = rng.choice([1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 17, 24, 31, 32, 33, 40])
    n_oth = rng.choice([0, 0, 1, depth // 2, (depth + 1) // 2, max(0, depth // 2 - 1), depth])
    syms = [rng.choice(".,") for _ in range(depth)]
    for i in rng.sample(range(depth), min(n_oth, depth)):
        syms[i] = rng.choice("ACGTNacgtn*" + ref.upper() + ref.lower() + "><RYk#7")
    toks = []
    for c in syms:
        x = rng.random()
        if x < 0.12:
            toks.append("^" + rng.choice("KIUS!~]+-$^.,*" + ref.upper() + ref.lower()))
        toks.append(c)
        if x > 0.9:
            toks.append("$")
        if 0.5 < x < 0.52:
            toks.append(rng.choice("+-") + rng.choice(["1A", "2ac", "", "0", "3GGG"]))
    bases = "".join(toks)
    if rng.random() < 0.05:
        bases += "^"
    nq = depth + rng.choice([0, 0, 0, 0, 0, 0, 1, -1, 2])
    quals = "".join(chr(rng.randint(33, 126)) for _ in range(max(nq, 0)))
    tail = rng.choice(["\n"] * 8 + ["\r\n", " \n"])
    return "%s\t%d\t%s\t%d\t%s\t%s%s" % (linegen.CHROM, pos, ref, depth, bases, quals, tail)


@pytest.mark.parametrize("seed", range(6))
def test_first_tier_boundary(sim, seed):
    rng = random.Random(700 + seed)
    n = 1500
    op = orc.make_params()
    lines = []
    for pos in range(1, n + 1):
        ln = _boundary_line(rng, pos)
        if orc.line_report(ln.rstrip("\n").encode(), op)["status"] == 0:
            lines.append(ln)
    text = "".join(lines).encode()
    if seed % 3 == 2:
        text = text[:-1]
    snps = [(linegen.CHROM, p) for p in rng.sample(range(1, n + 10), 300)]
    quick = 0
    for ps in (PARAM_SETS[seed % 2], PARAM_SETS[3], PARAM_SETS[5], PARAM_SETS[6]):
        for all_pos in (False, True):
            c = _compare(sim, text, snps, [], ps, all_pos)
            quick += int(c[5])
    assert quick > 0


def test_reference_file_vectors(sim, ref_files):
    n_ok = n_raise = 0
    for case in ref_files:
        snps = [(ln.split()[0], int(ln.split()[1])) for ln in case["snplist"].splitlines()]
        excl = []
        if case["exclude"]:
            for ln in case["exclude"].splitlines():
                if not ln.startswith("#"):
                    f = ln.split("\t")
                    excl.append((f[0], int(f[1])))
        text = case["pileup"].encode()
        rc, row, _, _ = run_sim(sim, text, snps, excl, case["params"], case["all_pos"])
        ref = case["ref"]
        if ref["exit"] == 0:
            assert rc == 0
            assert orc.fasta_text("sampleX", row.decode()) == ref["fasta"]
            n_ok += 1
        else:
            exc = {orc.E_VALUE: "ValueError", orc.E_INDEX: "IndexError", orc.E_UNPACK: "ValueError"}
            assert exc[rc] == ref["exit"].split(":")[1]
            n_raise += 1
    assert n_ok >= 20 and n_raise >= 20


def test_reference_line_vectors(sim, ref_lines):
    """Every golden line (realistic, nasty, broken) x parameter set through the fast+general dispatch."""
    n = 0
    for case in ref_lines:
        line = case["line"]
        if "\r" in line.rstrip("\r\n") or not line.endswith("\n"):
            continue
        ps = case["params"]
        ref = case["ref"]
        text = line.encode()
        rc, _, lines, _ = run_sim(sim, text, [(linegen.CHROM, 1)], [], ps, True)
        if "raises" in ref:
            assert rc in (orc.E_VALUE, orc.E_INDEX), line
            continue
        assert rc == 0, line
        op = orc.make_params(*ps)
        fails = ref["fails"]
        cons = ref["cons"]
        cell = "-" if (fails or cons == "*") else cons
        assert chr(int(lines[0]) & 0xff) == cell, line
        assert orc.fail_names(int(lines[0]) >> 8, op) == fails, line
        n += 1
    assert n > 2000


def test_lambda_golden(sim, golden_dir):
    root = os.path.join(golden_dir, "lambda")
    for branch in ("", "_preserved"):
        snps = orc.read_snp_list(os.path.join(root, "snplist%s.txt" % branch))
        for s in ("sample1", "sample2", "sample3", "sample4"):
            sdir = os.path.join(root, "samples", s)
            text = open(os.path.join(sdir, "reads.all.pileup"), "rb").read()
            excl = orc.vcf_sites(os.path.join(sdir, "var.flt_removed.vcf")) if branch else []
            for all_pos in (False, True):
                rc, row, _, c = run_sim(sim, text, snps, excl, (0, 0.6, 3, 0, 0.0), all_pos)
                assert rc == 0
                assert orc.fasta_text(s, row.decode()) == open(os.path.join(sdir, "consensus%s.fasta" % branch)).read()
                if all_pos:
                    assert c[2] < 200, "lambda pileups should stay on the fast path (got %d general lines)" % c[2]


@pytest.mark.parametrize("seed", range(4))
def test_indel_tokens_in_the_first_tier(sim, seed):
    """Lines with indel tokens [+-]<n><n letters> (samtools' shape, 1..3 digits, several per line, next to '^x' and '$',
    at the start and at the end of the column): decided by the first-tier parser, exactly."""
    rng = random.Random(7000 + seed)
    n = 1500
    lines = [linegen.realistic_line(rng, 1 + k, indel_rate=0.08) for k in range(n)]
    for k in range(0, n, 97):                                          # hand-made corners
        ref = rng.choice("ACGT")
        seq = "".join(rng.choice("ACGTNacgtn") for _ in range(12))
        bases = rng.choice(["+2AC....,,,,", "....,,,,-3acg", ".+1A,-1c.^K.$,,", ".-12" + seq + ",,..", ".+10" + seq[:10] + ".,"])
        nb = len(linegen.strip_bases(bases))
        lines[k] = "%s\t%d\t%s\t%d\t%s\t%s\n" % (linegen.CHROM, 1 + k, ref, nb, bases, "".join(chr(33 + rng.randint(13, 39)) for _ in range(nb)))
    text = "".join(lines).encode()
    snps = [(linegen.CHROM, p) for p in sorted(rng.sample(range(1, n + 1), 150))]
    for all_pos in (False, True):
        c = _compare(sim, text, snps, [], PARAM_SETS[seed % 4], all_pos)
        if all_pos and PARAM_SETS[seed % 4][0] <= 0:
            assert c[6] > 0.5 * n * 0.3, "the token-skipping second look should decide most indel lines (%d)" % c[6]


@pytest.mark.parametrize("seed", range(6))
def test_indel_token_corners(sim, seed):
    """Indel tokens the first tier must either decide exactly or hand on: at every byte alignment against the column's
    end, bare signs, counts of 0 / 4 digits / more bases than the column holds, sequences with non-letters, signs behind
    '^', tokens back to back -- whatever the tier, the result is the oracle's."""
    rng = random.Random(9100 + seed)
    n = 900
    lines = linegen.indel_corner_lines(rng, n)
    ps = PARAM_SETS[seed % len(PARAM_SETS)]
    op = orc.make_params(*ps)
    good, n_bad = [], 0
    for k, line in enumerate(lines):                   # lines the reference stops at (a bare sign: int("")): one by one
        try:
            orc.pileup_consensus(line.encode(), [], [], op, parse_all=True, want_lines=True)
            good.append(line)
        except orc.OracleError:
            n_bad += 1
            _compare(sim, line.encode(), [(linegen.CHROM, 1 + k)], [], ps, True)
    assert len(good) > n // 2 and n_bad > 10
    text = "".join(good).encode()
    snps = [(linegen.CHROM, p) for p in sorted(rng.sample(range(1, n + 1), 200))]
    for all_pos in (False, True):
        c = _compare(sim, text, snps, [], ps, all_pos)
        assert c[0] == (len(good) if all_pos else c[0])


@pytest.mark.parametrize("seed", range(4))
def test_odd_bytes_in_columns(sim, seed):
    """Bytes >= 0x80 (with and without a SWAR carry into their neighbour: 0xa0 / 0xa1), DEL, blanks and control bytes
    dropped into the name, depth, bases and quality columns of otherwise realistic lines: the first tier has to end the
    column there and decline; the result (the domain error included) is the oracle's."""
    for sub in range(60):
        rng = random.Random(1000 * seed + sub)
        lines = [linegen.realistic_line(rng, 1 + k, indel_rate=rng.choice([0.0, 0.05])).encode("latin-1") for k in range(12)]
        k = rng.randrange(len(lines))
        f = lines[k].split(b"\t")
        col = rng.choice([4, 4, 4, 5, 1, 0])
        b = bytearray(f[col])
        for _ in range(rng.choice([1, 1, 2])):
            b[rng.randrange(len(b))] = rng.choice([0x80, 0x81, 0xa0, 0xa1, 0xa2, 0xde, 0xdf, 0xe0, 0xff, 0x7f, 0x20, 0x1f,
                                                   0x0b, 0x0c, 0x1c, 0x00])
        f[col] = bytes(b)
        lines[k] = b"\t".join(f)
        for all_pos in (True, False):
            _compare(sim, b"".join(lines), [(linegen.CHROM, 1 + k)], [], PARAM_SETS[sub % len(PARAM_SETS)], all_pos)


ZERO_TAILS = ["*\t*", "\t", "", "*", "*\t*\textra\tcolumns", "* *", "..,\tII", "*\t*\r", "*\r*", "*\t\x0b", "\x80\t*", "*\t\xa1",
              "+2AC\t!", "\x1c\x1d", " ", "*\t*\t", "^", "\t\t\t"]


@pytest.mark.parametrize("seed", range(3))
def test_zero_depth_lines(sim, seed):
    """Raw depth 0 (pileup.py:226-234): the call is ('-', RawDpth) whatever follows the depth column; the first tier takes
    such lines itself unless the rest holds a CR / VT / FF or a byte >= 0x80 (lone CR ends a line, high bytes are out of
    the domain) -- every shape against the oracle, between ordinary lines."""
    rng = random.Random(seed)
    lines = []
    for k in range(160):
        if k % 2:
            lines.append(linegen.realistic_line(rng, 1 + k).encode("latin-1"))
            continue
        depth = rng.choice(["0", "0", "00", "000", "0000"])
        ref = rng.choice("ACGTNacgtnRr*")
        tail = rng.choice(ZERO_TAILS)
        sep = rng.choice(["\t", "\t", "\t", " ", ""])
        nl = rng.choice(["\n", "\n", "\r\n"])
        lines.append(("%s\t%d\t%s\t%s%s%s%s" % (linegen.CHROM, 1 + k, ref, depth, sep, tail, nl)).encode("latin-1"))
    snps = [(linegen.CHROM, 1 + k) for k in range(0, 160, 3)]
    for all_pos in (True, False):
        for k in range(0, 160, 2):                                 # one odd line at a time between ordinary ones (errors stop a run)
            text = b"".join(lines[max(0, k - 1):k + 2])
            _compare(sim, text, snps, [], PARAM_SETS[k % len(PARAM_SETS)], all_pos)
