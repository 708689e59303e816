"""The path's small kernels at the sizes of BASELINE configs[3] (K2 site union over 10 M pairs, K3 site table from keys, K5
consensus-VCF records + text, K6 depth sum, K7 abnormal regions of 1000 samples), for profiler captures of the kernels
that have none yet.  Runs as a plain script in seconds.  NOT yet captured: `ncu --set full -k regex:'k2_|k3_|k5_|k6_|k7_' -c 90`
over this script did not finish within 500 s on the box (round 2, last session: ncu saves and restores the process's ~1.5 GB of
device memory around every replay pass of every kernel) -- capture one kernel family per call with a narrow `-k`, or use
`--section SpeedOfLight --section MemoryWorkloadAnalysis` instead of the full set."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from snp_pipeline_b200 import _lib
from snp_pipeline_b200 import pileup as gpu_pileup

G = 5_000_000
CH = "gi|0000000|ref|SYN_5000K.1|"
ctx = _lib.Context(0)
rng = np.random.default_rng(3)
# ---- K2: 1000 samples x 10 k sites out of a 200 k pool = 10 M (key, sample) pairs -------------------------------------
pool = np.sort(rng.choice(G, 200_000, replace=False).astype(np.uint64) + 1)
keys = np.concatenate([np.sort(rng.choice(pool, 10_000, replace=False)) for _ in range(1000)])
samp = np.repeat(np.arange(1000, dtype=np.uint32), 10_000)
kd = torch.from_numpy(keys.view(np.int64)).cuda()
sd = torch.from_numpy(samp.view(np.int32)).cuda()
uq = torch.empty(keys.size, dtype=torch.int64, device="cuda")
cn = torch.empty(keys.size, dtype=torch.int32, device="cuda")
so = torch.empty(keys.size, dtype=torch.int32, device="cuda")
n_uniq = ctx.merge_sites_dev(kd.data_ptr(), sd.data_ptr(), keys.size, uq.data_ptr(), cn.data_ptr(), so.data_ptr())
print("K2: %d pairs -> %d sites" % (keys.size, n_uniq))
# ---- K3: the site table from K2's keys -----------------------------------------------------------------------------
sites = _lib.Sites.from_keys_dev(ctx, [CH], [G], uq.data_ptr(), n_uniq)
torch.cuda.synchronize()
# ---- K5 / K6: one 5 Mbp sample ---------------------------------------------------------------------------------------
spec = _lib.SynthSpec(20261017, 0, G, 24, G // 100, 0.05, 0.0)
cap = G * 112 + 4096
buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
n = ctx.synth_pileup_dev(spec, CH, buf.data_ptr(), cap)
host, owner = ctx.pinned_array(n)
host[:] = buf[:n].cpu().numpy()
caller = gpu_pileup.ConsensusCaller(0.6, 3, 0, 0.0)
params = caller.params(0)
ftexts = [";".join(caller.fail_names(m) or ["PASS"]) for m in range(_lib.VCF_FILTER_MASKS)]
ctx.want_vcf_records(True)
ctx.pileup_consensus(host, sites, params, _lib.MODE_SITES)
text, n_rec = ctx.pileup_vcf_text(sites, params, _lib.MODE_SITES, ftexts)
print("K5: %d records, %.1f MB of text" % (n_rec, text.size / 1e6))
ctx.want_vcf_records(False)
total, lines = ctx.pileup_depth_sum(host)
print("K6: depth sum %d over %d lines" % (total, lines))
# ---- K7: 1000 samples x 2400 SNPs each, mode "each" ---------------------------------------------------------------------
ns, per = 1000, 2400
pos = np.concatenate([np.sort(rng.choice(G, per, replace=False)).astype(np.uint64) + 1 for _ in range(ns)])
grp = np.repeat(np.arange(ns, dtype=np.uint64), per)
snp_keys = (grp << np.uint64(48)) | pos
seg_last = (np.arange(ns, dtype=np.uint32) + 1) * per - 1
seg_last = np.repeat(seg_last, per)
edge_keys = np.concatenate([(np.arange(ns, dtype=np.uint64) << np.uint64(48)) | np.uint64(1),
                            (np.arange(ns, dtype=np.uint64) << np.uint64(48)) | np.uint64(G - 499)])
edge_end = np.concatenate([np.full(ns, 500, np.uint32), np.full(ns, G, np.uint32)])
order = np.argsort(edge_keys, kind="stable")
removed = ctx.filter_regions(snp_keys, seg_last, [3], [1000], edge_keys[order], edge_end[order])
print("K7: %d of %d SNPs in abnormal regions" % (int(removed.sum()), removed.size))
