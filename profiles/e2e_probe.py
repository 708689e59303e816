"""Where the end-to-end leg loses time against the plain H2D copy: the pipelined host-buffer calls of bench.py's host_step
on one synthetic sample, with and without per-line results, against one call at a time and a bare cudaMemcpy loop.
    python profiles/e2e_probe.py [n_calls]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from snp_pipeline_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 40
G = 5000000
ctx = _lib.Context(0)
spec = _lib.SynthSpec(20261017, 0, G, 24, G // 100, 0.05, 0.0)
cap = G * 112 + 4096
buf = torch.empty(cap, dtype=torch.uint8, device="cuda")
n = ctx.synth_pileup_dev(spec, "gi|0000000|ref|SYN_5000K.1|", buf.data_ptr(), cap)
pos = ctx.synth_sample_sites(spec)
sites = _lib.Sites.from_arrays(ctx, ["gi|0000000|ref|SYN_5000K.1|"], np.zeros(pos.size, np.int32), pos.astype(np.int64))
pool = []
for k in range(4):
    arr, owner = ctx.pinned_array(n)
    arr[:] = buf[:n].cpu().numpy()
    pool.append((arr, owner))
rows_arr, ro = ctx.pinned_array(N * 4096)
rows = rows_arr.reshape(N, -1)
lines_arr, lo = ctx.pinned_array(2 * 2 * (G + 64))
lines = lines_arr.view(np.uint16).reshape(2, -1)
stats = (_lib.PileupStats(), _lib.PileupStats())
p = _lib.make_params(min_cons_depth=3)


def pipelined(want_lines, mode):
    in_flight = None
    for i in range(N + 1):
        nxt = None
        if i < N:
            slot = ctypes.c_int(-1)
            rc = ctx.lib.snpgpu_pileup_consensus_begin(
                ctx.handle, ctypes.c_void_p(pool[i % 4][0].ctypes.data), n, sites.handle, ctypes.byref(p), mode,
                ctypes.c_void_p(rows[i].ctypes.data), ctypes.c_void_p(lines[i % 2].ctypes.data) if want_lines else None,
                lines[i % 2].size if want_lines else 0, ctypes.byref(stats[i % 2]), ctypes.byref(slot))
            ctx._check(rc)
            nxt = slot.value
        if in_flight is not None:
            ctx._check(ctx.lib.snpgpu_pileup_consensus_end(ctx.handle, in_flight))
        in_flight = nxt


def one_at_a_time(mode):
    for i in range(N):
        ctx.pileup_consensus(pool[i % 4][0], sites, p, mode)


def bare_copy():
    cudart = ctypes.CDLL("libcudart.so.12")
    for i in range(N):
        cudart.cudaMemcpy(ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(pool[i % 4][0].ctypes.data), ctypes.c_size_t(n), 1)


for name, fn in (("bare cudaMemcpy H2D", bare_copy),
                 ("pipelined, all-positions, per-line results", lambda: pipelined(True, _lib.MODE_ALL)),
                 ("pipelined, all-positions, no per-line results", lambda: pipelined(False, _lib.MODE_ALL)),
                 ("pipelined, default mode", lambda: pipelined(False, _lib.MODE_SITES)),
                 ("one call at a time, default mode", lambda: one_at_a_time(_lib.MODE_SITES))):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%-48s %6.2f ms per sample  %5.1f GB/s" % (name, dt / N * 1e3, n * N / dt / 1e9))
