"""Which source lines load which issue pipe: per-SASS executed counts of an ncu report joined with nvdisasm's line info.

    python profiles/ncu_pipes.py <report.ncu-rep> <kernel regex> <cubin> <mangled entry> [top_n]

For the all-positions pileup kernel:
    cuobjdump -xelf all snp_pipeline_b200/libsnpgpu.so        (-> k1_pileup.sm_100a.cubin)
    python profiles/ncu_pipes.py gpurun_out/k1_r2d.ncu-rep k1_pileup k1_pileup.sm_100a.cubin \
        _ZN6snpgpu16k1_pileup_kernelILb1EEEvNS_7K1BatchE
Pipes as measured with profiles/micro/pipes.cu: LOP3 / SHF / PRMT / IADD3 / ISETP / SEL / LEA ... -> ALU, IMAD* / IDP -> FMA.
"""
import collections
import csv
import re
import subprocess
import sys

rep, kre, cubin, entry = sys.argv[1:5]
top_n = int(sys.argv[5]) if len(sys.argv) > 5 else 45

ALU = ("LOP3", "SHF", "PRMT", "IADD3", "VIADD", "ISETP", "SEL", "LEA", "PLOP3", "IMNMX", "VIMNMX", "MOV", "CS2R", "IABS", "VABSDIFF",
       "P2R", "R2P", "BMSK", "SGXT", "FSEL")
FMA = ("IMAD", "IDP", "FFMA", "FMUL", "FADD")
XU = ("POPC", "FLO", "BREV", "MUFU", "I2F", "F2I")
LSU = ("LDS", "STS", "LDG", "STG", "LD.", "ST.", "ATOM", "RED", "LDC", "LDL", "STL", "ATOMS", "ATOMG", "UBLKCP", "SYNCS", "LDSM")


def pipe(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in XU:
        return "xu"
    if base in ALU:
        return "alu"
    if any(op.startswith(x) for x in LSU):
        return "lsu"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "BREAK", "NANOSLEEP", "YIELD"):
        return "cbu"
    return "other"


# ---- line info per SASS offset -----------------------------------------------------------------------------------
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True, timeout=600).stdout.splitlines()
line_of, cur, inside = {}, ("?", 0), False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = (".text." + entry + " ") in ln + " " or ln.rstrip("- ").endswith(entry)
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur

# ---- executed counts per SASS instruction ----------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True, timeout=600).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isrc, iex, ith, isamp = (hdr.index(x) for x in ("Address", "Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
ins = [r for r in rows[2:] if len(r) > iex and r[0].startswith("0x")]
base = int(ins[0][ia], 16)
per_line = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in ins:
    off = int(r[ia], 16) - base
    src = r[isrc].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    p = pipe(op)
    n = int(r[iex])
    key = line_of.get(off, ("?", 0))
    per_line[key][p] += n
    per_line[key]["all"] += n
    per_line[key]["thr"] += int(r[ith])
    per_line[key]["samp"] += int(r[isamp])
    tot[p] += n
    tot["all"] += n
print("warp instructions %d: alu %.1f%%  fma %.1f%%  lsu %.1f%%  xu %.1f%%  cbu %.1f%%  other %.1f%%" % (
    tot["all"], *(100.0 * tot[k] / tot["all"] for k in ("alu", "fma", "lsu", "xu", "cbu", "other"))))
files = collections.defaultdict(lambda: collections.Counter())
for (f, l), c in per_line.items():
    files[f].update(c)
for f, c in sorted(files.items(), key=lambda kv: -kv[1]["alu"]):
    print("%-22s all %5.1f%%  alu %5.1f%% of alu  fma %5.1f%% of fma" % (f, 100.0 * c["all"] / tot["all"], 100.0 * c["alu"] / tot["alu"], 100.0 * c["fma"] / max(tot["fma"], 1)))
print("top lines by ALU-pipe instructions (share of all ALU instructions; counts per line: alu / fma / lsu / all in M; avg threads)")
for (f, l), c in sorted(per_line.items(), key=lambda kv: -kv[1]["alu"])[:top_n]:
    print("%-18s %4d  alu %5.2f%%  %6.2f / %6.2f / %6.2f / %6.2f  thr %4.1f  samp %4.1f%%" % (
        f, l, 100.0 * c["alu"] / tot["alu"], c["alu"] / 1e6, c["fma"] / 1e6, c["lsu"] / 1e6, c["all"] / 1e6, c["thr"] / max(c["all"], 1),
        100.0 * c["samp"] / max(sum(x["samp"] for x in per_line.values()), 1)))
