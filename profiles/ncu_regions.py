"""Share of executed warp instructions per function of K1 (ncu source page joined with the source files).
Usage: python profiles/ncu_regions.py report.ncu-rep"""
import bisect
import collections
import csv
import os
import re
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "snp_pipeline_b200", "csrc")
DEF = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:SNP_HD_NOINLINE|SNP_HD|__device__|__global__|static|inline)\b[^;(]*?\b(\w+)\s*\($")
DEF2 = re.compile(r"^\s*(?:SNP_HD_NOINLINE|SNP_HD|__device__|__global__)\b.*?\b(\w+)\s*\(")
starts = {}
for name in os.listdir(CSRC):
    if not name.endswith((".cuh", ".cu", ".h")):
        continue
    lst = []
    for i, ln in enumerate(open(os.path.join(CSRC, name)), 1):
        m = DEF2.match(ln)
        if m and not ln.rstrip().endswith(";"):
            lst.append((i, m.group(1)))
    starts[name] = lst

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True, timeout=600).stdout
cur, agg, thr = None, collections.Counter(), collections.Counter()
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 12 or not r[0].isdigit():
        continue
    lst = starts.get(cur, [])
    k = bisect.bisect_right([s[0] for s in lst], int(r[0])) - 1
    f = lst[k][1] if k >= 0 else "?"
    try:
        agg[(cur, f)] += int(r[7]); thr[(cur, f)] += int(r[8])
    except ValueError:
        pass
tot = sum(agg.values()) or 1
print("total warp instructions %d" % tot)
for k, v in agg.most_common(30):
    print("%-18s %-24s %5.1f%%  avg active threads %.1f" % (k[0], k[1], 100 * v / tot, thr[k] / max(v, 1)))
