"""Profiling driver: one synthetic 5 Mbp sample resident in HBM, K1 (all-positions mode) launched a few times.
Usage (on the GPU box):  ncu --set full --clock-control none --import-source on -k regex:k1_pileup_kernel -s 2 -c 1 \
                             -o gpurun_out/k1 python profiles/run_k1.py [sites|all] [n_launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from snp_pipeline_b200 import _lib

mode = _lib.MODE_SITES if len(sys.argv) > 1 and sys.argv[1] == "sites" else _lib.MODE_ALL
n_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
G = int(os.environ.get("GENOME_LEN", "5000000"))
ctx = _lib.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
spec = _lib.SynthSpec(20261017, 0, G, 24, G // 100, 0.05, float(os.environ.get("INDEL_RATE", "0")))   # 0: the default 0.1 % of lines
cap = G * 112 + 4096
buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
n = ctx.synth_pileup_dev(spec, "gi|0000000|ref|SYN_5000K.1|", buf.data_ptr(), cap)
pos = ctx.synth_sample_sites(spec)
extra = int(os.environ.get("EXTRA_SITES", "0"))          # a union as large as a big batch's: more lines at sites
if extra:
    pos = np.union1d(pos, np.random.default_rng(1).choice(G, extra, replace=False).astype(pos.dtype) + 1)
sites = _lib.Sites.from_arrays(ctx, ["gi|0000000|ref|SYN_5000K.1|"], np.zeros(pos.size, np.int32), pos.astype(np.int64))
B = int(os.environ.get("BATCH", "8"))                     # samples per launch (the same text, separate outputs)
row = torch.empty((B, max(pos.size, 1)), dtype=torch.uint8, device="cuda")
lines = torch.empty((B, G + 64), dtype=torch.int16, device="cuda")
stats = torch.zeros((B, 6), dtype=torch.int64, device="cuda")
p = _lib.make_params(min_cons_depth=3)
ctx.enable_timing(True)
batch = [(buf.data_ptr(), n, row[i].data_ptr(), lines[i].data_ptr(), G + 64, stats[i].data_ptr()) for i in range(B)]
ctx.pileup_consensus_batch_dev(batch, sites, p, mode)           # (warm-up: buffers grow on the first call)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(stream)
for _ in range(n_launch):
    ctx.pileup_consensus_batch_dev(batch, sites, p, mode)
ev1.record(stream)
torch.cuda.synchronize()
ms, k = ctx.kernel_time(0)
per = ms / (k * B)
whole = ev0.elapsed_time(ev1) / (n_launch * B)                  # pileup + follow-up + ordering + finish kernels, per sample
print("text bytes %d, K1 avg %.4f ms per sample over %d launches of %d samples -> %.1f GB/s; whole call %.4f ms per sample; stats %s"
      % (n, per, k, B, n / per / 1e6, whole, stats[0].tolist()))
