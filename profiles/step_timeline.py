"""GPU timeline of the bench's device-resident step (where the step's time outside K1 goes): two steps of bench.device_step
under torch.profiler (CUPTI sees the library's kernels too), chrome trace -> gpurun_out/step_trace.json; run on the GPU
box, summarise with `python profiles/step_timeline.py summarize gpurun_out/step_trace.json`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def summarize(path):
    ev = json.load(open(path))["traceEvents"]
    k = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "ts" in e]
    k.sort(key=lambda e: e["ts"])
    marks = [e for e in ev if e.get("name", "").startswith("step_") and e.get("ph") == "X" and e.get("cat") in ("user_annotation", "gpu_user_annotation")]
    t0 = k[0]["ts"]
    end_prev = None
    tot = {}
    for e in k:
        name = e["name"].split("(")[0][:60]
        gap = e["ts"] - end_prev if end_prev is not None else 0.0
        tot.setdefault(name, [0, 0.0, 0.0])
        tot[name][0] += 1; tot[name][1] += e["dur"]; tot[name][2] += max(gap, 0.0)
        end_prev = max(end_prev or 0, e["ts"] + e["dur"])
    span = k[-1]["ts"] + k[-1]["dur"] - t0
    busy = sum(e["dur"] for e in k)
    print("span %.1f us, busy %.1f us, idle %.1f us over %d GPU activities" % (span, busy, span - busy, len(k)))
    for name, (n, dur, gap) in sorted(tot.items(), key=lambda kv: -kv[1][1] - kv[1][2]):
        print("%-62s n %4d  busy %10.1f us  idle in front %9.1f us" % (name, n, dur, gap))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "summarize":
        summarize(sys.argv[2])
        sys.exit(0)
    import torch
    import bench
    from snp_pipeline_b200 import _lib
    sys.argv = [sys.argv[0]]
    args = bench.parse_args()
    ctx = _lib.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    w = bench.Workload(ctx, torch, args, 0)
    for _ in range(3):
        bench.device_step(w, None, 1)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            bench.device_step(w, None, 1)
        torch.cuda.synchronize()
    prof.export_chrome_trace(os.path.join(ROOT, "gpurun_out", "step_trace.json"))
    summarize(os.path.join(ROOT, "gpurun_out", "step_trace.json"))
