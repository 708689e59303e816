"""Turn the round's ncu artefacts (gpurun_out/) into the tracked summaries under profiles/.

    python profiles/summarize.py <round tag> <k1 full capture .ncu-rep> <launch list .csv>

Writes profiles/<tag>_k1_metrics.csv (selected raw metrics of the dominant kernel), profiles/<tag>_k1_lines.txt (share
of instructions / stall samples per source line), profiles/<tag>_k1_regions.txt (per function), profiles/<tag>_launches.txt
(per-kernel share of device time in the launch list) and profiles/k1_traffic.json (DRAM bytes per launch, read by
bench.py for roofline.traffic)."""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
tag, rep, launches = sys.argv[1], sys.argv[2], sys.argv[3]

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=300).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
KEEP += ["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
with open(os.path.join(HERE, "%s_k1_metrics.csv" % tag), "w") as f:
    f.write("metric,unit,value\n")
    for krow in rows[2:]:                              # every kernel of the capture (the pileup kernel, its follow-up kernel)
        mk = {h: (u, v) for h, u, v in zip(hdr, units, krow)}
        f.write("kernel,,%s\n" % mk.get("Kernel Name", ("", "?"))[1].replace(",", ";"))
        for k in hdr:
            if k in KEEP or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                f.write("%s,%s,%s\n" % (k, mk[k][0], mk[k][1]))


def to_bytes(key):
    u, v = m[key]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    return float(v) * scale


traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
json.dump({"dram_bytes_per_launch": traffic, "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
           "dram_bytes_write": to_bytes("dram__bytes_write.sum"), "source": os.path.basename(rep),
           "workload": "k1_pileup_kernel over a batch of TWO synthetic 5 Mbp samples, all-positions mode (profiles/run_k1.py, "
                       "BATCH=2): divide by 2 for the per-sample figure",
           "dram_bytes_per_sample": traffic / 2},
          open(os.path.join(HERE, "k1_traffic.json"), "w"), indent=1)

for script, out in (("ncu_lines.py", "%s_k1_lines.txt"), ("ncu_regions.py", "%s_k1_regions.txt")):
    txt = subprocess.run([sys.executable, os.path.join(HERE, script), rep], capture_output=True, text=True, timeout=900).stdout
    open(os.path.join(HERE, out % tag), "w").write(txt)

per = collections.Counter()
n = collections.Counter()
for r in csv.DictReader(l for l in open(launches) if l.startswith('"')):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        name = r["Kernel Name"].split("(")[0]
        per[name] += float(r["Metric Value"])
        n[name] += 1
tot = sum(per.values()) or 1
with open(os.path.join(HERE, "%s_launches.txt" % tag), "w") as f:
    f.write("# per-kernel device time in the launch list %s (ncu --metrics gpu__time_duration.sum; cold-cache,\n"
            "# serialised: compare shares, not absolutes)\n" % os.path.basename(launches))
    for k, v in per.most_common():
        f.write("%-70s launches %4d  total %10.3f ms  share %5.1f%%  avg %9.3f us\n" % (k, n[k], v / 1e6, 100 * v / tot, v / n[k] / 1e3))
print("wrote summaries for", tag, "; DRAM traffic per K1 launch = %.1f MB" % (traffic / 1e6))
