"""debug: which line of test_nasty_text[seed] does sites mode miss?  (bisect over prefixes of the text)"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import linegen
from oracle import oracle as orc
from snp_pipeline_b200 import _lib
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2
orc.build()
rng = random.Random(100 + seed)
n = 900
text = linegen.pileup_text(200 + seed, n, nasty=0.5).encode()
if seed % 2 == 0:
    op = orc.make_params()
    text = b"".join(ln + b"\n" for ln in text.split(b"\n")[:-1] if orc.line_report(ln, op)["status"] == 0)
if seed % 3 == 0:
    text = text.replace(b"\r\n", b"\n").replace(b"\n", b"\r\n")
if seed % 4 == 1:
    text = text[:-1]
snps = [(linegen.CHROM, p) for p in rng.sample(range(1, n + 10), 300)]
excl = [(linegen.CHROM, p) for p in rng.sample(range(1, n), 30)]
ctx = _lib.Context(0)
sites = ctx.sites(snps, excl)
ps = (15, 0.6, 3, 1, 0.1)
op = orc.make_params(*ps)
gp = _lib.make_params(*ps) if False else _lib.make_params(min_base_qual=ps[0], min_cons_freq=ps[1], min_cons_depth=ps[2], min_cons_strand_depth=ps[3], min_cons_strand_bias=ps[4])
lines = text.split(b"\n")[:-1]
def counts(k):
    t = b"".join(l + b"\n" for l in lines[:k])
    want = len(orc.pileup_consensus(t, snps, excl, op, parse_all=False, want_lines=True)[1][0])
    got = ctx.pileup_consensus(t, sites, gp, _lib.MODE_SITES)[1].n_parsed
    return want, got
lo, hi = 0, len(lines)
print("full", counts(hi))
while hi - lo > 1:
    mid = (lo + hi) // 2
    w, g = counts(mid)
    if w == g: lo = mid
    else: hi = mid
print("first bad prefix length", hi, counts(hi), "line:", lines[hi - 1][:200])
# the line alone
t = lines[hi - 1] + b"\n"
print("alone", len(orc.pileup_consensus(t, snps, excl, op, parse_all=False, want_lines=True)[1][0]), ctx.pileup_consensus(t, sites, gp, _lib.MODE_SITES)[1].n_parsed)
