"""Per source line: which stall reason its samples carry.  Usage: python profiles/ncu_stalls.py rep.ncu-rep [reason] [top_n]
reason: long_sb | short_sb | wait | branch_resolving | no_inst | math | not_selected ..."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
reason = sys.argv[2] if len(sys.argv) > 2 else "long_sb"
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True, timeout=600).stdout
cur, hdr, agg = None, None, collections.Counter()
src = {}
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    col = hdr.index("stall_" + reason)
    try:
        agg[(cur, int(r[0]))] += int(r[col] or 0)
        src[(cur, int(r[0]))] = r[1].strip()[:110]
    except ValueError:
        pass
tot = sum(agg.values()) or 1
print("stall_%s samples: %d" % (reason, tot))
for k, v in agg.most_common(top_n):
    print("%-16s %4d  %5.1f%%  %s" % (k[0], k[1], 100.0 * v / tot, src[k]))
