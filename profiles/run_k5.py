"""K5 alone: one synthetic 5 Mbp sample through the host-buffer call, then the consensus VCF's data lines as text.
Usage (on the GPU box):  python profiles/run_k5.py [sites|all] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from snp_pipeline_b200 import _lib
from snp_pipeline_b200 import pileup as gpu_pileup

mode = _lib.MODE_SITES if len(sys.argv) > 1 and sys.argv[1] == "sites" else _lib.MODE_ALL
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
G = int(os.environ.get("GENOME_LEN", "5000000"))
ctx = _lib.Context(0)
spec = _lib.SynthSpec(20261017, 0, G, 24, G // 100, 0.05, 0.0)
cap = G * 112 + 4096
buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
n = ctx.synth_pileup_dev(spec, "gi|0000000|ref|SYN_5000K.1|", buf.data_ptr(), cap)
pos = ctx.synth_sample_sites(spec)
extra = int(os.environ.get("EXTRA_SITES", "45000"))
pos = np.union1d(pos, np.random.default_rng(1).choice(G, extra, replace=False).astype(pos.dtype) + 1)
sites = _lib.Sites.from_arrays(ctx, ["gi|0000000|ref|SYN_5000K.1|"], np.zeros(pos.size, np.int32), pos.astype(np.int64))
host, owner = ctx.pinned_array(n)
host[:] = buf[:n].cpu().numpy()
caller = gpu_pileup.ConsensusCaller(0.6, 3, 0, 0.0)
params = caller.params(0)
ftexts = [";".join(caller.fail_names(m) or ["PASS"]) for m in range(_lib.VCF_FILTER_MASKS)]
ctx.want_vcf_records(True)
for k in range(reps + 1):
    t0 = time.perf_counter()
    ctx.pileup_consensus(host, sites, params, mode)
    t1 = time.perf_counter()
    text, n_rec = ctx.pileup_vcf_text(sites, params, mode, ftexts)
    t2 = time.perf_counter()
    if k:
        print("mode %s: pileup_consensus %.2f ms, vcf text %.2f ms for %d records / %.1f MB of text (%.1f M records/s)"
              % (sys.argv[1] if len(sys.argv) > 1 else "all", (t1 - t0) * 1e3, (t2 - t1) * 1e3, n_rec, text.size / 1e6,
                 n_rec / (t2 - t1) / 1e6))
print(text[:300].tobytes().decode())
