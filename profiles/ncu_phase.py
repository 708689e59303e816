"""Share of instructions and stall samples per phase of the pileup kernel (source page of an ncu report exported with
`ncu -i rep --page source --csv --print-source sass,cuda > file.csv`).  Usage: python profiles/ncu_phase.py file.csv"""
import csv, collections, sys
cur = None; hdr = None
inst = collections.Counter(); samp = collections.Counter(); reasons = collections.defaultdict(collections.Counter)
def num(x):
    try: return int(x)
    except ValueError: return 0
for r in csv.reader(open(sys.argv[1])):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    key = cur + ':' + r[1].strip()[:28] if False else cur
    inst[key] += num(r[7]); samp[key] += num(r[6])
    for i, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            reasons[key][h] += num(r[i])
ti = sum(inst.values()); ts = sum(samp.values())
for k, v in inst.most_common(14):
    top = ', '.join('%s %.0f%%' % (a[6:], 100 * b / max(1, samp[k])) for a, b in reasons[k].most_common(5))
    print('%-28s inst %5.1f%%  samples %5.1f%%   %s' % (k, 100 * v / ti, 100 * samp[k] / ts, top))
