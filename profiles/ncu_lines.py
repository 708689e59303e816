"""Aggregate an ncu report's source page by CUDA source line: share of warp instructions, average active threads,
share of stall samples.  Usage: python profiles/ncu_lines.py gpurun_out/k1.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True, timeout=600).stdout
rows = list(csv.reader(raw.splitlines()))
cur, agg = None, collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 12 or not r[0].isdigit():
        continue
    try:
        agg[(cur, int(r[0]))] = (int(r[7]), int(r[8]), int(r[6]), r[1].strip()[:100])
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()) or 1
tots = sum(v[2] for v in agg.values()) or 1
thr = sum(v[1] for v in agg.values())
print("warp instructions %d, thread instructions %d (avg %.1f active), stall samples %d" % (tot, thr, thr / tot, tots))
for (f, l), (inst, th, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print("%-16s %4d  inst %5.1f%%  avgthr %5.1f  samp %4.1f%%  %s" % (f, l, 100 * inst / tot, th / max(inst, 1), 100 * s / tots, src))
