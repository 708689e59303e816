"""Lines of one source file with their share of stall samples and instructions (from an exported source page csv).
Usage: python profiles/ncu_file_lines.py file.csv k1_pileup.cu [min_samples]"""
import csv, collections, sys
path, want = sys.argv[1], sys.argv[2]
floor = int(sys.argv[3]) if len(sys.argv) > 3 else 60
cur = None; hdr = None; agg = collections.OrderedDict(); ts = ti = 0
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in csv.reader(open(path)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    ts += num(r[6]); ti += num(r[7])
    if cur == want:
        a = agg.setdefault(int(r[0]), [0, 0, r[1].strip()[:100]]); a[0] += num(r[6]); a[1] += num(r[7])
for ln in sorted(agg):
    s, i, src = agg[ln]
    if s >= floor: print('%4d samp %4.1f%% inst %4.1f%%  %s' % (ln, 100 * s / ts, 100 * i / ti, src))
