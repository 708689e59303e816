"""Profiling driver for K4: a synthetic n_rows x n_sites SNP matrix resident in HBM, one rank's stripe of the distance
matrix.  Usage: python profiles/run_k4.py [n_rows] [n_sites] [stripe_rows] [n_launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from snp_pipeline_b200 import _lib

n_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
n_sites = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
stripe = int(sys.argv[3]) if len(sys.argv) > 3 else 625
n_launch = int(sys.argv[4]) if len(sys.argv) > 4 else 3
ctx = _lib.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
g = torch.Generator(device="cuda").manual_seed(5)
ref = torch.randint(0, 4, (n_sites,), device="cuda", generator=g)
alt = (ref + torch.randint(1, 4, (n_sites,), device="cuda", generator=g)) % 4
carry = torch.rand((n_rows, n_sites), device="cuda", generator=g) < 0.05
code = torch.where(carry, alt.expand(n_rows, -1), ref.expand(n_rows, -1))
lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
m = lut[code]
m[torch.rand((n_rows, n_sites), device="cuda", generator=g) < 0.03] = ord("-")
del carry, code
d = torch.empty((stripe, n_rows), dtype=torch.int32, device="cuda")
ctx.enable_timing(True)
for _ in range(n_launch):
    ctx.pairwise_distance_dev(m.data_ptr(), n_rows, n_sites, n_sites, 0, stripe, d.data_ptr())
torch.cuda.synchronize()
ms, k = ctx.kernel_time(1)
per = ms / k
print("K4 stripe %d x %d rows x %d sites: %.3f ms per launch -> %.2f T pair-sites/s" % (stripe, n_rows, n_sites, per, stripe * n_rows * n_sites / per / 1e9))
# a sampled block against a straightforward torch expression
i = torch.arange(0, min(stripe, 48), device="cuda"); j = torch.arange(n_rows - 48, n_rows, device="cuda")
a, b = m[i][:, None, :], m[j][None, :, :]
valid = (a != ord("-")) & (b != ord("-"))
want = ((a != b) & valid).sum(-1).to(torch.int32)
assert torch.equal(d[i][:, j], want), "sampled block differs"
print("sampled 48 x 48 block identical")
