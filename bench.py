#!/usr/bin/env python
"""bench.py -- the hot path (site union -> pileup parse + tally + consensus -> SNP matrix -> pairwise distance) on
BASELINE.json's configuration 2: synthetic 100 samples x 5 Mbp reference, ~50 k SNP sites, per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the whole path over the rank's batch of samples:
    K2 union of the samples' variant-site lists          (merge_sites)
    K1 every pileup line parsed, tallied, called         (call_consensus --vcfAllPos: all-positions mode, so no
       + K3 gather into the samples x sites matrix        line is skipped), 2 B of result per line + the matrix row
    K4 all-pairs distance over the matrix                 (distance)
`value` times that with the pileup text resident in HBM (100 samples = 44.5 GB, far above L2, so no flush is
needed); `e2e` times the same path through the host-buffer C-ABI calls (pinned host text -> H2D inside the timed
region, results read back every step).  Multi-GPU: samples are sharded per rank (weak scaling: 100 per GPU), one NCCL
all-gather of the per-rank site lists and one of the matrix rows; each rank computes its stripe of the distances.
`--impl reference` times the CPU restatement of the reference (oracle/, all host threads) on a bounded sample of
the same workload.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONTIG = "gi|0000000|ref|SYN_5000K.1|"
SEED = 20261017
METRIC = "pileup positions/sec through site-union + parse/tally/consensus + SNP-matrix + pairwise distance"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", type=int, default=100, help="samples per GPU")
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--pool-sites", type=int, default=50_000)
    ap.add_argument("--carry", type=float, default=0.05)
    ap.add_argument("--host-pool", type=int, default=4, help="distinct samples kept in pinned host memory for e2e")
    ap.add_argument("--cpu-samples", type=int, default=2, help="samples the cpu_baseline leg parses")
    ap.add_argument("--config", default="c2", choices=["c2", "c4", "c5"],
                    help="c2: BASELINE configs[1] (the default, the metric's configuration); c4: configs[3], 1000 samples "
                         "sharded over the ranks, files + SHA-256 parity with the oracle; c5: configs[4], distance only")
    ap.add_argument("--samples-total", type=int, default=0, help="c4 / c5: samples over all ranks (default 1000 / 5000)")
    ap.add_argument("--out-dir", default="/tmp/snp_pipeline_b200_bench_files")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ workload
class Workload(object):
    """The rank's batch: synthetic pileups resident in HBM + each sample's variant-site list."""

    def __init__(self, ctx, torch, args, rank):
        from snp_pipeline_b200 import _lib
        self.ctx, self.torch, self.args, self.rank = ctx, torch, args, rank
        self.lib = _lib
        self.n = args.samples
        self.texts, self.nbytes, self.site_pos = [], [], []
        cap = args.genome_len * 112 + 4096
        scratch = torch.empty(cap, dtype=torch.uint8, device="cuda")
        for i in range(self.n):
            spec = self.spec(i)
            nb = ctx.synth_pileup_dev(spec, CONTIG, scratch.data_ptr(), cap)
            t = torch.empty(nb + 64, dtype=torch.uint8, device="cuda")
            t[:nb].copy_(scratch[:nb])
            self.texts.append(t)
            self.nbytes.append(nb)
            self.site_pos.append(ctx.synth_sample_sites(spec))
        del scratch
        self.total_text = int(sum(self.nbytes))
        keys = np.concatenate([p.astype(np.uint64) for p in self.site_pos])          # chrom rank 0
        samp = np.concatenate([np.full(p.size, i, dtype=np.uint32) for i, p in enumerate(self.site_pos)])
        self.keys_host, self.samp_host = keys, samp
        self.keys_dev = torch.from_numpy(keys.view(np.int64)).cuda()
        self.samp_dev = torch.from_numpy(samp.view(np.int32)).cuda()
        nk = keys.size
        self.uniq_dev = torch.empty(max(nk, 1), dtype=torch.int64, device="cuda")
        self.cnt_dev = torch.empty(max(nk, 1), dtype=torch.int32, device="cuda")
        self.sout_dev = torch.empty(max(nk, 1), dtype=torch.int32, device="cuda")
        self.lines_dev = torch.empty((self.n, args.genome_len + 64), dtype=torch.int16, device="cuda")
        self.stats_dev = torch.zeros((self.n, 6), dtype=torch.int64, device="cuda")
        self.params = _lib.make_params(min_cons_depth=3)
        # the batch's descriptors (snpgpu_pileup_sample), built once: a step only fills in where its matrix rows lie
        self.batch_arr = (_lib.PileupSample * self.n)()
        for i in range(self.n):
            self.batch_arr[i] = _lib.PileupSample(self.texts[i].data_ptr(), int(self.nbytes[i]), None, self.lines_dev[i].data_ptr(),
                                                  args.genome_len + 64, self.stats_dev[i].data_ptr())

    def spec(self, i):
        a = self.args
        return self.lib.SynthSpec(SEED, self.rank * self.n + i, a.genome_len, 24, a.pool_sites, a.carry, 0.0)


def build_sites(ctx, positions):
    """Site table from the merged positions (one contig): snpgpu_sites_create via the array fast path."""
    from snp_pipeline_b200 import _lib
    return _lib.Sites.from_arrays(ctx, [CONTIG], np.zeros(positions.size, dtype=np.int32),
                                  positions.astype(np.int64))


def device_step(w, dist, world):
    """One device-resident pass.  Returns (n_sites, matrix tensor, distance tensor)."""
    from snp_pipeline_b200 import sharding
    torch, ctx = w.torch, w.ctx
    nk = w.keys_host.size
    n_uniq = ctx.merge_sites_dev(w.keys_dev.data_ptr(), w.samp_dev.data_ptr(), nk, w.uniq_dev.data_ptr(),
                                 w.cnt_dev.data_ptr(), w.sout_dev.data_ptr())
    local = w.uniq_dev[:n_uniq]
    if world > 1:
        # the path's one exchange step: all-gather of the per-rank sorted-unique site lists (NCCL over NVLink)
        parts = sharding.allgather_varlen(local, dist, world)
        counts_h = [int(p.numel()) for p in parts]
        allkeys = torch.cat(parts)
        owner = torch.cat([torch.full((int(counts_h[r]),), r, dtype=torch.int32, device="cuda") for r in range(world)])
        gu = torch.empty_like(allkeys); gc = torch.empty(allkeys.numel(), dtype=torch.int32, device="cuda")
        gs = torch.empty(allkeys.numel(), dtype=torch.int32, device="cuda")
        n_uniq = ctx.merge_sites_dev(allkeys.data_ptr(), owner.data_ptr(), allkeys.numel(), gu.data_ptr(),
                                     gc.data_ptr(), gs.data_ptr())
        local = gu[:n_uniq]
    # the merged list never leaves HBM: K2's sorted unique keys -> the table K1 probes (snpgpu_sites_create_from_keys_dev)
    sites = w.lib.Sites.from_keys_dev(ctx, [CONTIG], [w.args.genome_len], local.data_ptr(), n_uniq)
    matrix = torch.empty((w.n, (max(n_uniq, 1) + 63) // 64 * 64), dtype=torch.uint8, device="cuda")    # 16-byte aligned rows
    # one launch sequence per <= 64 samples (snpgpu_pileup_consensus_batch_dev); every sample has its own per-line results
    row0, pitch = matrix.data_ptr(), matrix.shape[1]
    for i in range(w.n):
        w.batch_arr[i].row_out_dev = row0 + i * pitch
    ctx.pileup_consensus_batch_dev(w.batch_arr, sites, w.params, w.lib.MODE_ALL)
    full = sharding.allgather_rows(matrix, dist, world)
    lo = w.rank * w.n
    d = torch.empty((w.n, full.shape[0]), dtype=torch.int32, device="cuda")
    ctx.pairwise_distance_dev(full.data_ptr(), full.shape[0], n_uniq, full.shape[1], lo, lo + w.n, d.data_ptr())
    sites.close()
    w.full_matrix = full                                       # (every rank's rows: the N > 1 parity check reads it)
    return n_uniq, matrix, d


def w_last_sites(w, ctx, dist, world, torch):
    """The global sorted-unique site keys as this rank sees them (the exchange of device_step, repeated)."""
    from snp_pipeline_b200 import sharding
    n_uniq = ctx.merge_sites_dev(w.keys_dev.data_ptr(), w.samp_dev.data_ptr(), w.keys_host.size, w.uniq_dev.data_ptr(),
                                 w.cnt_dev.data_ptr(), w.sout_dev.data_ptr())
    parts = sharding.allgather_varlen(w.uniq_dev[:n_uniq], dist, world)
    allkeys = torch.cat(parts)
    owner = torch.cat([torch.full((int(p.numel()),), r, dtype=torch.int32, device="cuda") for r, p in enumerate(parts)])
    gu = torch.empty_like(allkeys); gc = torch.empty(allkeys.numel(), dtype=torch.int32, device="cuda")
    gs = torch.empty(allkeys.numel(), dtype=torch.int32, device="cuda")
    n = ctx.merge_sites_dev(allkeys.data_ptr(), owner.data_ptr(), allkeys.numel(), gu.data_ptr(), gc.data_ptr(), gs.data_ptr())
    return gu[:n].clone()


def host_step(w, pool, pool_n, bufs, dist=None, world=1):
    """One end-to-end pass through the host-buffer C-ABI calls (what a ctypes user of the library makes).  With more
    than one rank the two exchange steps of device_step() happen here too, from and to host memory."""
    from snp_pipeline_b200 import sharding
    torch, ctx, lib = w.torch, w.ctx, w.lib
    t_0 = time.perf_counter()
    uniq, cnt, samples = ctx.merge_sites(w.keys_host, w.samp_host)
    h2d = w.keys_host.nbytes + w.samp_host.nbytes
    d2h = uniq.nbytes + cnt.nbytes + samples.nbytes
    if world > 1:                                              # the global union: all-gather of the per-rank lists
        local = torch.from_numpy(uniq.view(np.int64)).cuda()
        parts = sharding.allgather_varlen(local, dist, world)
        allkeys = torch.cat(parts).cpu().numpy().view(np.uint64)
        owner = np.concatenate([np.full(int(p.numel()), r, dtype=np.uint32) for r, p in enumerate(parts)])
        h2d += uniq.nbytes + allkeys.nbytes + owner.nbytes
        uniq, cnt, samples = ctx.merge_sites(allkeys, owner)
        d2h += allkeys.nbytes + uniq.nbytes + cnt.nbytes + samples.nbytes
    t_1 = time.perf_counter()
    sites = build_sites(ctx, uniq)
    t_2 = time.perf_counter()
    n_sites = uniq.size
    rows, lines, stats = bufs
    # one call kept ahead (snpgpu_pileup_consensus_begin / _end): sample i+1's text crosses PCIe while sample i's
    # kernels run and its results come back
    in_flight = None
    worst = (0.0, -1, "")
    for i in range(w.n + 1):
        nxt = None
        t_b = time.perf_counter()
        if i < w.n:
            k = i % len(pool)
            slot = ctypes.c_int(-1)
            rc = ctx.lib.snpgpu_pileup_consensus_begin(
                ctx.handle, ctypes.c_void_p(pool[k].ctypes.data), pool_n[k], sites.handle, ctypes.byref(w.params),
                lib.MODE_ALL, ctypes.c_void_p(rows[i].ctypes.data), ctypes.c_void_p(lines[i % 2].ctypes.data),
                lines[i % 2].size, ctypes.byref(stats[i % 2]), ctypes.byref(slot))
            ctx._check(rc)
            h2d += pool_n[k]
            nxt = (slot.value, i % 2)
        t_m = time.perf_counter()
        if in_flight is not None:
            ctx._check(ctx.lib.snpgpu_pileup_consensus_end(ctx.handle, in_flight[0]))
            d2h += n_sites + 2 * stats[in_flight[1]].n_lines + ctypes.sizeof(stats[0])
        t_e = time.perf_counter()
        if t_m - t_b > worst[0]:
            worst = (t_m - t_b, i, "begin")
        if t_e - t_m > worst[0]:
            worst = (t_e - t_m, i, "end")
        in_flight = nxt
    t_3 = time.perf_counter()
    m = rows[:, :n_sites]
    if world == 1:
        d = np.zeros((w.n, w.n), dtype=np.int32)
        ctx._check(ctx.lib.snpgpu_pairwise_distance(ctx.handle, ctypes.c_void_p(rows.ctypes.data), w.n, n_sites,
                                                    rows.shape[1], ctypes.c_void_p(d.ctypes.data)))
        h2d += w.n * rows.shape[1]
        d2h += d.nbytes
    else:                                                      # every rank needs every row: all-gather, then its stripe
        block = torch.from_numpy(rows).cuda()
        full = sharding.allgather_rows(block, dist, world)
        dd = torch.empty((w.n, full.shape[0]), dtype=torch.int32, device="cuda")
        lo = w.rank * w.n
        ctx.pairwise_distance_dev(full.data_ptr(), full.shape[0], n_sites, full.shape[1], lo, lo + w.n, dd.data_ptr())
        d = dd.cpu().numpy()
        h2d += rows.nbytes
        d2h += d.nbytes
    t_4 = time.perf_counter()
    sites.close()
    w.e2e_breakdown = {"merge_sites_ms": (t_1 - t_0) * 1e3, "site_table_ms": (t_2 - t_1) * 1e3,
                       "pileup_calls_ms": (t_3 - t_2) * 1e3, "distance_ms": (t_4 - t_3) * 1e3,
                       "site_table_free_ms": (time.perf_counter() - t_4) * 1e3,
                       "slowest_call": {"ms": worst[0] * 1e3, "sample": worst[1], "which": worst[2]}}
    w.e2e_all = getattr(w, "e2e_all", []) + [w.e2e_breakdown]
    return m, d, h2d, d2h


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_k1(orc, text, snps):
    return orc.pileup_consensus(text, snps, [], orc.make_params(min_cons_depth=3), parse_all=True, want_lines=True)


def cpu_baseline(w, matrix_host, n_samples):
    """The oracle (CPU port of the reference's algorithm), one core, on a bounded sample of the same workload."""
    from oracle import oracle as orc
    orc.build()
    uniq_t0 = time.perf_counter()
    uniq, cnt, samples = orc.merge_sites_keys(w.keys_host, w.samp_host)
    t_k2 = time.perf_counter() - uniq_t0
    snps = [(CONTIG, int(p)) for p in uniq]
    t_k1 = 0.0
    for i in range(n_samples):
        text = w.texts[i][:w.nbytes[i]].cpu().numpy()
        t0 = time.perf_counter()
        row, _ = cpu_k1(orc, text, snps)
        t_k1 += time.perf_counter() - t0
        assert row == matrix_host[i].tobytes(), "cpu_baseline: the oracle's row differs from the GPU's"
    t0 = time.perf_counter()
    d = orc.distance_matrix([bytes(r) for r in matrix_host])
    t_k4 = time.perf_counter() - t0
    t_step = t_k1 / n_samples * w.n + t_k2 + t_k4
    return {"value": w.n * w.args.genome_len / t_step, "unit": "positions/s", "cores": 1, "kind": "port",
            "sample": "%d of %d samples through the oracle's K1 (all-positions mode, %.1f s) scaled x%g, plus K2 "
                      "(%.3f s) and K4 (%.2f s) on the full batch" % (n_samples, w.n, t_k1, w.n / n_samples, t_k2, t_k4),
            "k1_positions_per_s_per_core": n_samples * w.args.genome_len / t_k1}, d


def reference_python_leg(w, n_lines=150_000):
    """cpu_baseline.reference_python of the GPU arm: reference_python_legs() on the rank's first synthetic sample."""
    text = w.texts[0][:w.nbytes[0]].cpu().numpy().tobytes()
    return reference_python_legs(text, w.site_pos[0], n_lines)


def reference_python_legs(text, own_sites, n_lines=150_000):
    """The reference's OWN Python (staged under oracle/_ref by oracle/stage_ref.py, unmodified), one core, on the GPU box's
    host (SURVEY.md section 8d "CPU baseline timing"):
      (ii)  pileup.Reader over every one of the first n_lines lines of one synthetic sample + ConsensusCaller.call_consensus
            per record -- the loop of call_consensus.py:161-176 without the file writing (--vcfAllPos);
      (iii) the same file in the default filter mode (Reader with the sample's own sites as the position set);
      (iv)  utils.calculate_sequence_distance over all pairs of 40 x 10 000-site rows (distance.py:93-96);
      (i)   the reference's four subcommands on the bundled lambda data (tests/golden/lambda.tar.xz), wall time per stage.
    None when the modules are not staged."""
    import tempfile
    from oracle import stage_ref
    root = stage_ref.staged_root()
    if root is None:
        return None
    os.environ["SNP_REFERENCE_ROOT"] = root
    from oracle import ref_harness
    ref_harness.REFERENCE_ROOT = root
    try:
        pileup = ref_harness.ref("pileup")
    except Exception as e:  # noqa: BLE001
        return {"unavailable": repr(e)[:200]}
    end, k = 0, 0
    while k < n_lines:
        nxt = text.find(b"\n", end)
        if nxt < 0:
            break
        end, k = nxt + 1, k + 1
    with tempfile.NamedTemporaryFile("wb", suffix=".pileup", delete=False) as f:
        f.write(text[:end])
        path = f.name
    out = {"cores": 1}
    try:
        caller = pileup.ConsensusCaller(0.6, 3, 0, 0.0)
        t0 = time.perf_counter()
        n = 0
        for record in pileup.Reader(path, 0, None):
            caller.call_consensus(record)
            n += 1
        dt = time.perf_counter() - t0
        out.update({"positions_per_s_per_core": n / dt, "lines_timed": n, "seconds": dt,
                    "what": "snppipeline/pileup.py (unmodified, staged by oracle/stage_ref.py): Reader(all positions) + "
                            "ConsensusCaller.call_consensus per record"})
        wanted = set((CONTIG, int(p)) for p in own_sites)
        t0 = time.perf_counter()
        n_f = hits = 0
        for record in pileup.Reader(path, 0, wanted):
            caller.call_consensus(record)
            hits += 1
        n_f = k
        dt = time.perf_counter() - t0
        out["filter_mode"] = {"positions_per_s_per_core": n_f / dt, "lines_timed": n_f, "lines_at_sites": hits, "seconds": dt,
                              "what": "the default mode of run.py:709: Reader(position set = the sample's own sites), "
                                      "records only at the sites"}
    finally:
        os.unlink(path)
    try:                                                        # (iv) the distance loop of the reference
        utils = ref_harness.ref("utils")
        rng = np.random.default_rng(5)
        rows = ["".join(r) for r in np.array(list("ACGTN-acgt"))[rng.integers(0, 10, size=(40, 10_000))]]
        t0 = time.perf_counter()
        for i in range(len(rows)):
            for j in range(i + 1, len(rows)):
                utils.calculate_sequence_distance(rows[i], rows[j])
        dt = time.perf_counter() - t0
        pairs = len(rows) * (len(rows) - 1) // 2
        out["distance"] = {"pair_sites_per_s_per_core": pairs * 10_000 / dt, "pairs": pairs, "sites": 10_000, "seconds": dt,
                           "what": "utils.calculate_sequence_distance (utils.py:1135-1165) over all pairs of 40 rows"}
    except Exception as e:  # noqa: BLE001
        out["distance"] = {"unavailable": repr(e)[:200]}
    try:
        out["c1_lambda"] = reference_c1_stages(ref_harness)
    except BaseException as e:  # noqa: BLE001  (the reference's error paths end in sys.exit)
        out["c1_lambda"] = {"unavailable": repr(e)[:200]}
    return out


def reference_c1_stages(ref_harness):
    """BASELINE configs[0]: the reference's own merge_sites -> call_consensus x 4 -> snp_matrix -> distance on the bundled
    lambda-virus samples (4 x 48 502 positions, 166 sites), in-process through its command-line parser, wall seconds per
    stage; the consensus files are compared with the bundled expected ones."""
    import contextlib
    import io
    import tarfile
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        with tarfile.open(os.path.join(ROOT, "tests", "golden", "lambda.tar.xz")) as tar:
            tar.extractall(tmp, filter="data")
        base = os.path.join(tmp, "lambda")
        dirs = sorted(os.path.join(base, "samples", d) for d in os.listdir(os.path.join(base, "samples")))
        work = os.path.join(tmp, "work")
        os.makedirs(work)
        sdf = os.path.join(work, "sampleDirectories.txt")
        with open(sdf, "w") as f:
            f.write("".join(d + "\n" for d in dirs))
        expected = {d: open(os.path.join(d, "consensus.fasta")).read() for d in dirs}
        for d in dirs:
            os.rename(os.path.join(d, "consensus.fasta"), os.path.join(d, "consensus.expected"))
        snplist, snpma = os.path.join(work, "snplist.txt"), os.path.join(work, "snpma.fasta")
        stages = {}
        sink = io.StringIO()

        def run(stage, line):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sink), contextlib.redirect_stderr(sink):
                ref_harness.run_command(line)
            stages[stage] = stages.get(stage, 0.0) + time.perf_counter() - t0

        run("merge_sites", "merge_sites -f -n var.flt.vcf -o %s %s %s" % (snplist, sdf, os.path.join(work, "filtered.txt")))
        for d in dirs:
            run("call_consensus", "call_consensus -f -l %s -o %s %s" % (snplist, os.path.join(d, "consensus.fasta"),
                                                                     os.path.join(d, "reads.all.pileup")))
        run("snp_matrix", "snp_matrix -f -c consensus.fasta -o %s %s" % (snpma, sdf))
        run("distance", "distance -f -p %s -m %s %s" % (os.path.join(work, "pairwise.tsv"), os.path.join(work, "matrix.tsv"), snpma))
        same = all(open(os.path.join(d, "consensus.fasta")).read() == expected[d] for d in dirs)
        return {"stage_seconds": stages, "samples": len(dirs), "positions_per_sample": 48502,
                "consensus_identical_to_bundled": same,
                "what": "the reference's own subcommands (run_command_from_line, default parameters, no consensus.vcf) on "
                        "tests/golden/lambda.tar.xz, one process, one core"}


def run_reference(args, rank):
    """--impl reference: the CPU restatement of the reference's path with every host thread, bounded sample."""
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc
    orc.build()
    orc.lib()
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    n_ref = max(1, min(cores, 32, args.samples))             # one sample per host thread (32 x 0.46 GB of text at most)
    # the inputs: the same synthetic samples the GPU arm parses, written on the host by the generator's own line function
    # (oracle/synth_host.cpp) -- this arm loads neither torch nor libsnpgpu.so
    texts = [orc.synth_pileup(SEED, i, args.genome_len, 24, args.pool_sites, args.carry, 0.0, CONTIG, threads=cores)
             for i in range(n_ref)]
    site_pos = [orc.synth_sample_sites(SEED, i, args.genome_len, 24, args.pool_sites, args.carry) for i in range(n_ref)]
    keys = np.concatenate([p.astype(np.uint64) for p in site_pos])
    samp = np.concatenate([np.full(p.size, i, dtype=np.uint32) for i, p in enumerate(site_pos)])

    def step():
        uniq, cnt, samples = orc.merge_sites_keys(keys, samp)
        snps = [(CONTIG, int(p)) for p in uniq]
        with ThreadPoolExecutor(max_workers=min(cores, n_ref)) as ex:     # ctypes drops the GIL inside the C call
            rows = list(ex.map(lambda t: cpu_k1(orc, t, snps)[0], texts))
        orc.distance_matrix(rows)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = n_ref * args.genome_len / dt
    ref_py = None
    if not args.no_cpu:
        try:                                                 # the reference's own Python beside its C restatement (one core)
            ref_py = reference_python_legs(texts[0].tobytes(), site_pos[0])
        except Exception as e:  # noqa: BLE001
            ref_py = {"unavailable": repr(e)[:200]}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "positions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "positions/s", "cores": min(cores, n_ref), "kind": "port",
                         "sample": "%d of %d samples per step, one thread each (oracle/snp_oracle.c, all-positions "
                                   "mode) + K2 and K4 on those samples" % (n_ref, args.samples),
                         "reference_python": ref_py},
        "e2e": {"value": value, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": "BASELINE configs[1]: synthetic %d samples x %.1f Mbp per GPU, ~%dk-site pool, depth ~24; "
                        "call_consensus in all-positions mode (every line parsed)" %
                        (args.samples, args.genome_len / 1e6, args.pool_sites // 1000),
            "samples_per_gpu": args.samples, "genome_len": args.genome_len, "pool_sites": args.pool_sites,
            "l2": "inputs (%.1f GB of text per GPU) exceed L2; no flush needed" % (args.samples * args.genome_len * 89e-9),
            "host_pool_samples": args.host_pool, "parallelism": "samples sharded x%d" % world}


# ------------------------------------------------------------------------------------------ configs[3]: 1000 samples, files, parity
def run_c4(args, torch, dist, ctx, rank, world, local_rank, barrier):
    """BASELINE configs[3] through the package-level driver (snp_pipeline_b200/batch.py, the reference's merge_sites ->
    call_consensus x N -> snp_matrix -> distance, run.py:691-732, 775-776): samples sharded over the ranks, one all-gather of
    the site lists, one of the rows; then -- outside the timed region -- the same files from the CPU oracle run over the
    same synthetic inputs, compared by SHA-256."""
    import hashlib
    from concurrent.futures import ThreadPoolExecutor
    from snp_pipeline_b200 import _lib, batch, sharding
    from oracle import oracle as orc
    total = args.samples_total or 1000
    pool = args.pool_sites if args.pool_sites != 50_000 else 200_000
    lo, hi = sharding.shard_bounds(total, rank, world)
    names = ["s%05d" % i for i in range(total)]
    cap = args.genome_len * 112 + 4096
    scratch = torch.empty(cap, dtype=torch.uint8, device="cuda")
    texts, site_pos = [], []
    for i in range(lo, hi):
        spec = _lib.SynthSpec(SEED, i, args.genome_len, 24, pool, args.carry, 0.0)
        nb = ctx.synth_pileup_dev(spec, CONTIG, scratch.data_ptr(), cap)
        t = torch.empty(nb + 64, dtype=torch.uint8, device="cuda")
        t[:nb].copy_(scratch[:nb])
        texts.append(t[:nb])
        site_pos.append(ctx.synth_sample_sites(spec))
    del scratch
    sites = [{CONTIG: sp} for sp in site_pos]
    params = _lib.make_params(min_cons_depth=3)

    def step(out_dir=None):
        return batch.run_hot_path(ctx, names[lo:hi], texts, sites, [CONTIG], [args.genome_len], params, out_dir=out_dir,
                                  mode=_lib.MODE_SITES, dist=dist, rank=rank, world=world)

    for _ in range(args.warmup):
        res = step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.enable_timing(True)
    ctx.kernel_time(0); ctx.kernel_time(1)
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        res = step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    k1_ms, _ = ctx.kernel_time(0)
    k4_ms, k4_n = ctx.kernel_time(1)
    ctx.enable_timing(False)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt.item()) / args.steps
    res = step(args.out_dir)                                   # once more, with the files (rank 0 writes them)
    # ---- the oracle over the same inputs: every rank its own samples, rank 0 the union and the distances ----------
    orc.build()
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = max(1, cores // world)
    t0 = time.perf_counter()
    parts = [None] * world
    if world > 1:
        dist.all_gather_object(parts, [sp.astype(np.uint32) for sp in site_pos])
    else:
        parts = [site_pos]
    all_pos = [sp for blk in parts for sp in blk]
    keys = np.concatenate([sp.astype(np.uint64) for sp in all_pos])
    samp = np.concatenate([np.full(sp.size, i, dtype=np.uint32) for i, sp in enumerate(all_pos)])
    uniq, cnt, grouped = orc.merge_sites_keys(keys, samp)
    snps = [(CONTIG, int(p)) for p in uniq]
    op = orc.make_params(min_cons_depth=3)

    def one(t):
        return orc.pileup_consensus(t.cpu().numpy(), snps, [], op, parse_all=False)

    with ThreadPoolExecutor(max_workers=threads) as ex:
        rows = list(ex.map(one, texts))
    per = (total + world - 1) // world
    blk = torch.full((per, max(len(snps), 1)), ord("-"), dtype=torch.uint8, device="cuda")
    if rows:
        blk[:len(rows), :len(snps)] = torch.from_numpy(np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(len(rows), len(snps)).copy()).cuda()
    full = sharding.allgather_rows(blk, dist, world)
    parity = None
    if rank == 0:
        rows_all = [bytes(r) for r in full[:total, :len(snps)].cpu().numpy()]
        want = orc.hot_path_texts(names, None, [[(CONTIG, int(p)) for p in sp] for sp in all_pos], op, threads=cores, rows=rows_all)[:3]
        parity = {"oracle_seconds": None, "files": {}}
        ok = True
        for key, text in zip(("snplist", "snpma", "distance_matrix"), want):
            got = hashlib.sha256(open(res.files[key], "rb").read()).hexdigest()
            exp = hashlib.sha256(text.encode()).hexdigest()
            ok &= got == exp
            parity["files"][os.path.basename(res.files[key])] = {"sha256": got, "oracle_sha256": exp, "identical": got == exp,
                                                                  "bytes": os.path.getsize(res.files[key])}
        parity["identical"] = bool(ok)
        parity["oracle_seconds"] = time.perf_counter() - t0
        parity["oracle"] = "oracle/snp_oracle.c: K1 per sample on %d host threads per rank, K2 and K4 on rank 0 (%d threads)" % (threads, cores)
        assert ok, "C4: the files differ from the oracle's: %r" % parity
    if rank == 0:
        positions = total * args.genome_len
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        text_local = float(sum(int(t.numel()) for t in texts))
        k1_per = k1_ms / max(args.steps * len(texts), 1)
        ach = (text_local / max(len(texts), 1) + res.n_sites) / (k1_per * 1e-3) / 1e9 if k1_per else None
        line = {
            "metric": METRIC, "value": positions / (ms_step * 1e-3), "unit": "positions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3]: synthetic %d samples x %.1f Mbp sharded over %d GPUs, ~%dk-site pool; "
                                   "merge_sites -> call_consensus (default mode: lines at snplist positions) -> snp_matrix -> "
                                   "distance through snp_pipeline_b200.batch.run_hot_path" %
                                   (total, args.genome_len / 1e6, world, pool // 1000),
                       "samples_total": total, "samples_per_gpu": len(texts), "genome_len": args.genome_len, "pool_sites": pool,
                       "l2": "inputs (%.1f GB of text per GPU) exceed L2; no flush needed" % (text_local / 1e9),
                       "parallelism": "samples sharded x%d, all-gather of site lists + rows" % world},
            "clocks": clocks, "e2e": None, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k1_pileup_kernel (sites mode)", "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": ach / peak if ach else None, "avg_launch_ms": k1_per, "share_of_step": k1_ms / ms if ms else None,
                         "traffic": None, "k4": {"ms_per_step": k4_ms / max(args.steps, 1), "launches": k4_n}},
            "cpu_baseline": None, "n_sites": int(res.n_sites), "parity": parity,
        }
        print(json.dumps(line))


# ------------------------------------------------------------------------------------------ configs[4]: 5000 samples, distance only
def run_c5(args, torch, dist, ctx, rank, world, local_rank, barrier):
    """BASELINE configs[4]: the pairwise SNP-distance matrix of 5000 samples x ~200 k sites (12.5 M pairs) over the ranks.
    Every rank holds the matrix (in the pipeline: after the row all-gather); the triangle's 64-row tile rows are dealt in
    zigzag order (sharding.zigzag_tile_rows), each rank computes its share with K4, rank 0 gathers and mirrors."""
    from snp_pipeline_b200 import sharding
    from oracle import oracle as orc
    n = args.samples_total or 5000
    s_sites = args.pool_sites if args.pool_sites != 50_000 else 200_000
    stride = (s_sites + 63) // 64 * 64
    g = torch.Generator(device="cuda").manual_seed(SEED)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    ref = torch.randint(0, 4, (s_sites,), device="cuda", generator=g)
    alt = (ref + torch.randint(1, 4, (s_sites,), device="cuda", generator=g)) % 4
    m = torch.empty((n, stride), dtype=torch.uint8, device="cuda")
    for lo in range(0, n, 500):                                # (in blocks: the temporaries stay small)
        hi = min(lo + 500, n)
        carry = torch.rand((hi - lo, s_sites), device="cuda", generator=g) < 0.05
        blk = lut[torch.where(carry, alt.expand(hi - lo, -1), ref.expand(hi - lo, -1))]
        blk[torch.rand((hi - lo, s_sites), device="cuda", generator=g) < 0.03] = ord("-")
        blk[torch.rand((hi - lo, s_sites), device="cuda", generator=g) < 0.01] = ord("N")
        m[lo:hi, :s_sites] = blk
    del carry, blk
    share = [sharding.zigzag_tile_rows(n, r, world) for r in range(world)]
    most = max(len(x) for x in share)
    part = torch.zeros((max(most, 1) * 64, n), dtype=torch.int32, device="cuda")
    parts = [torch.empty_like(part) for _ in range(world)] if rank == 0 and world > 1 else None
    n_tiles = (n + 63) // 64

    def step():
        if share[rank]:
            ctx.pairwise_distance_tiles_dev(m.data_ptr(), n, s_sites, stride, share[rank], part.data_ptr())
        if world > 1:
            dist.gather(part, parts, dst=0)
        if rank != 0:
            return None
        upper = torch.zeros((n_tiles, 64, n), dtype=torch.int32, device="cuda")
        for r in range(world):
            if share[r]:
                src = parts[r] if world > 1 else part
                upper[torch.tensor(share[r], device="cuda")] = src.view(-1, 64, n)[:len(share[r])]
        upper = upper.view(n_tiles * 64, n)[:n]
        row = torch.arange(n, device="cuda")
        return torch.where(row[None, :] >= (row // 64 * 64)[:, None], upper, upper.t())

    for _ in range(args.warmup):
        d = step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.enable_timing(True)
    ctx.kernel_time(1)
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        d = step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    k4_ms, k4_n = ctx.kernel_time(1)
    ctx.enable_timing(False)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    tt = torch.tensor([ms, k4_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step = float(tt[0].item()) / args.steps
    k4_step = float(tt[1].item()) / args.steps                 # the slowest rank's K4 time per step
    if rank == 0:
        assert bool((d == d.t()).all()) and not bool(torch.diagonal(d).any()), "C5: not symmetric with a zero diagonal"
        rng = np.random.default_rng(7)                          # a sampled 64 x 64 block of pairs against the oracle
        ri, rj = np.sort(rng.choice(n, 64, replace=False)), np.sort(rng.choice(n, 64, replace=False))
        rows = [bytes(r) for r in m[torch.from_numpy(np.concatenate([ri, rj])).cuda(), :s_sites].cpu().numpy()]
        orc.build()
        want = orc.distance_matrix(rows)[:64, 64:]
        got = d[torch.from_numpy(ri).cuda()][:, torch.from_numpy(rj).cuda()].cpu().numpy()
        assert np.array_equal(got, want), "C5: a sampled block differs from the oracle"
        pairs = n * (n - 1) // 2
        n_sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        peak = n_sms * 16 * 32 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12   # POPC: 16 lanes / clk / SM, 32 sites per word
        my_pair_sites = sum((n_tiles - t) for t in share[0]) * 64 * 64 * s_sites  # (rank 0's share, tiles incl. the diagonal ones)
        line = {
            "metric": "pairwise SNP-distance matrix: sample pairs x sites per second (BASELINE configs[4])",
            "value": pairs * s_sites / (ms_step * 1e-3), "unit": "pair-sites/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]: synthetic %d samples x %d sites, %d pairs, tile rows of the triangle dealt "
                                   "zigzag over %d GPUs, gathered and mirrored on rank 0" % (n, s_sites, pairs, world),
                       "samples_total": n, "n_sites": s_sites, "pairs": pairs,
                       "l2": "matrix %.2f GB + planes %.2f GB per GPU exceed L2; no flush needed" % (n * stride / 1e9, 3 * n * stride / 8e9),
                       "parallelism": "tile rows dealt zigzag x%d, one gather" % world},
            "clocks": clocks, "e2e": None, "gpu_launches": int(launches), "pairs_per_s": pairs / (ms_step * 1e-3),
            "roofline": {"bound": "issue (POPC: one per pair and 32 sites)", "kernel": "k4_pack_kernel + k4_pairs_kernel",
                         "achieved": my_pair_sites / (k4_step * 1e-3) / 1e12 if k4_step else None, "peak": peak,
                         "unit": "T pair-sites/s per GPU", "frac": (my_pair_sites / (k4_step * 1e-3) / 1e12) / peak if k4_step else None,
                         "k4_ms_per_step": k4_step, "share_of_step": k4_step / ms_step if ms_step else None, "traffic": None,
                         "peak_source": "n_sms x 16 POPC lanes/clk (profiles/micro/pipes.cu) x 32 sites x SM clock"},
            "cpu_baseline": None,
            "parity": {"sampled_block_vs_oracle": "64 x 64 pairs identical", "symmetric_zero_diagonal": True},
        }
        print(json.dumps(line))


# ------------------------------------------------------------------------------------------ main
def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    from snp_pipeline_b200 import _lib, device
    numa_cpus = device.bind_to_gpu_cpus(local_rank)           # pinned host buffers next to the GPU (e2e leg)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":   # its banner goes to stdout, in front of the JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = _lib.Context(local_rank)
    # one explicit stream for torch's ops, the library's kernels and the timing events (the legacy default stream has
    # handle 0, which snpgpu_set_stream reads as "the context's own stream")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if args.config in ("c4", "c5"):
        (run_c4 if args.config == "c4" else run_c5)(args, torch, dist, ctx, rank, world, local_rank, barrier)
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return
    w = Workload(ctx, torch, args, rank)

    # ---- device-resident: value ------------------------------------------------------------------
    for _ in range(args.warmup):
        n_sites, matrix, d = device_step(w, dist, world)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.enable_timing(True)
    ctx.kernel_time(0); ctx.kernel_time(1)
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        n_sites, matrix, d = device_step(w, dist, world)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    k1_ms, k1_n = ctx.kernel_time(0)
    k4_ms, k4_n = ctx.kernel_time(1)
    ctx.enable_timing(False)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    positions = world * w.n * args.genome_len
    value = positions / (ms_step * 1e-3)
    stats = w.stats_dev.cpu().numpy()
    assert (stats[:, 0] == args.genome_len).all() and (stats[:, 4] == 0).all(), "K1 reported an error"
    matrix_host = matrix[:, :n_sites].cpu().numpy()
    d_host = d.cpu().numpy()

    # ---- multi-GPU: size-independent checks of the two exchange steps (outside the timed region) -------
    if world > 1:
        keys_now = w_last_sites(w, ctx, dist, world, torch)
        sig = torch.stack([keys_now.sum(), keys_now.numel() * torch.ones((), dtype=torch.int64, device="cuda"),
                           (keys_now * torch.arange(1, keys_now.numel() + 1, device="cuda")).sum()])
        sigs = torch.empty((world, 3), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(sigs, sig)
        assert bool((sigs == sigs[0]).all()), "the ranks disagree about the merged site list"
        full_d = torch.empty((world * w.n, world * w.n), dtype=torch.int32, device="cuda")
        dist.all_gather_into_tensor(full_d, d.contiguous())            # the ranks' row stripes, in rank order
        assert bool((full_d == full_d.T).all()) and not bool(torch.diagonal(full_d).any()), \
            "the distance matrix assembled from the ranks' stripes is not symmetric with a zero diagonal"
        if rank == 0 and not args.no_cpu:
            # ... and against the oracle: rank 0's first two rows under the GLOBAL site list, and a block of distances
            # between its first 16 rows and 64 rows sampled from all ranks
            from oracle import oracle as orc
            orc.build()
            snps_g = [(CONTIG, int(p)) for p in keys_now.cpu().numpy()]
            for i in range(min(2, w.n)):
                row = orc.pileup_consensus(w.texts[i][:w.nbytes[i]].cpu().numpy(), snps_g, [], orc.make_params(min_cons_depth=3), parse_all=True)
                assert row == matrix_host[i].tobytes(), "N > 1: the oracle's row %d differs from the GPU's" % i
            cols = np.sort(np.random.default_rng(11).choice(world * w.n, min(64, world * w.n), replace=False))
            fm = w.full_matrix[:, :n_sites].cpu().numpy()
            rows_o = [bytes(r) for r in fm[:16]] + [bytes(fm[c]) for c in cols]
            want = orc.distance_matrix(rows_o)[:16, 16:]
            assert np.array_equal(d_host[:16][:, cols], want), "N > 1: a sampled block of distances differs from the oracle's"

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = (w.total_text + 2 * w.n * args.genome_len + w.n * n_sites) / w.n      # per launch: text + 2 B/line + row
    k1_avg_ms = k1_ms / max(args.steps * w.n, 1)              # per sample: the batches' launches / the samples they covered
    achieved = alg_bytes / (k1_avg_ms * 1e-3) / 1e9
    traffic = None
    try:
        # (the ncu capture ran a batch of two samples: per sample, the unit `achieved` and `alg_bytes_per_launch` use)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json"))).get("dram_bytes_per_sample")
    except (OSError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": "k1_pileup_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650",
                "alg_bytes_per_launch": alg_bytes, "avg_launch_ms": k1_avg_ms, "launches_timed": k1_n,
                "launch_unit": "one sample (a launch of k1_pileup_kernel + k1_rest_kernel covers a batch of up to 64 "
                               "samples; its time is divided by the samples it covered)",
                "share_of_step": k1_ms / ms if ms else None,
                "k4": {"avg_launch_ms": k4_ms / max(k4_n, 1), "pair_sites_per_s":
                       (w.n * (world * w.n) * n_sites) / (k4_ms / max(k4_n, 1) * 1e-3) if k4_ms else None}}

    # ---- the same kernel in the pipeline's default mode (only lines at snplist positions are parsed, no per-line
    #      output: call_consensus without --vcfAllPos, run.py:709), reported next to the all-positions roofline ----------
    sites_d = _lib.Sites.from_keys_dev(ctx, [CONTIG], [args.genome_len], w.uniq_dev.data_ptr(), n_sites) if world == 1 else None   # (N = 1 only)
    if sites_d is not None:
        for timed in (False, True):
            ctx.enable_timing(timed)
            ctx.kernel_time(0)
            ctx.pileup_consensus_batch_dev([(w.texts[i].data_ptr(), w.nbytes[i], matrix[i].data_ptr(), 0, 0,
                                             w.stats_dev[i].data_ptr()) for i in range(w.n)], sites_d, w.params,
                                           w.lib.MODE_SITES)
            torch.cuda.synchronize()
        d_ms, d_n = ctx.kernel_time(0)
        d_n = w.n
        ctx.enable_timing(False)
        sites_d.close()
        d_bytes = (w.total_text + w.n * n_sites) / w.n
        d_ach = d_bytes / (d_ms / max(d_n, 1) * 1e-3) / 1e9
        roofline["default_mode"] = {"what": "k1_pileup_kernel with only the lines at snplist positions parsed (no --vcfAllPos)",
                                    "avg_launch_ms": d_ms / max(d_n, 1), "alg_bytes_per_launch": d_bytes, "achieved": d_ach,
                                    "frac": d_ach / peak, "launches_timed": d_n}
        assert np.array_equal(w.stats_dev.cpu().numpy()[:, 0], stats[:, 0]), "default mode saw another number of lines"

    # ---- end to end through the host-buffer C ABI ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        pool, pool_n, owners = [], [], []
        for k in range(min(args.host_pool, w.n)):
            arr, owner = ctx.pinned_array(w.nbytes[k])
            arr[:] = w.texts[k][:w.nbytes[k]].cpu().numpy()
            pool.append(arr); pool_n.append(w.nbytes[k]); owners.append(owner)
        rows_arr, rows_owner = ctx.pinned_array(w.n * ((n_sites + 63) // 64 * 64 + 64))
        rows = rows_arr.reshape(w.n, -1)
        lines_arr, lines_owner = ctx.pinned_array(2 * 2 * (args.genome_len + 64))
        bufs = (rows, lines_arr.view(np.uint16).reshape(2, -1), (_lib.PileupStats(), _lib.PileupStats()))
        for _ in range(max(1, min(args.warmup, 1))):
            m, dd, h2d, d2h = host_step(w, pool, pool_n, bufs, dist, world)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        step_ms = []
        for _ in range(e2e_steps):
            ts = time.perf_counter()
            m, dd, h2d, d2h = host_step(w, pool, pool_n, bufs, dist, world)
            step_ms.append((time.perf_counter() - ts) * 1e3)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item()) / e2e_steps
        for k in range(w.n):
            assert m[k].tobytes() == matrix_host[k % len(pool)].tobytes(), "e2e row differs from the device-resident row"
        e2e = {"value": positions / e2e_s, "unit": "positions/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "h2d_gbs": h2d / e2e_s / 1e9, "step_breakdown": getattr(w, "e2e_all", [])[-e2e_steps:], "step_ms": step_ms}
        # ---- the per-sample call the reference's run.py:709 makes: default mode + consensus.vcf (K1 sites mode, then K5:
        #      sort of the parsed lines' offsets, tallies, the data lines' text on the device), host buffers, N = 1 only
        if world == 1:
            from snp_pipeline_b200 import pileup as gpu_pileup
            caller = gpu_pileup.ConsensusCaller(0.6, 3, 0, 0.0)
            ftexts = [";".join(caller.fail_names(mm) or ["PASS"]) for mm in range(_lib.VCF_FILTER_MASKS)]
            sites_v = _lib.Sites.from_keys_dev(ctx, [CONTIG], [args.genome_len], w.uniq_dev.data_ptr(), n_sites)
            ctx.want_vcf_records(True)
            t_call = t_vcf = 0.0
            n_lines = n_bytes = 0
            reps = min(3, len(pool))
            for k in range(reps + 1):                             # (first pass: warm-up)
                t0 = time.perf_counter()
                ctx.pileup_consensus(pool[k % len(pool)][:pool_n[k % len(pool)]], sites_v, w.params, w.lib.MODE_SITES)
                t1 = time.perf_counter()
                vtext, vrec = ctx.pileup_vcf_text(sites_v, w.params, w.lib.MODE_SITES, ftexts)
                t2 = time.perf_counter()
                if k:
                    t_call += t1 - t0; t_vcf += t2 - t1; n_lines += vrec; n_bytes += int(vtext.size)
            ctx.want_vcf_records(False)
            sites_v.close()
            e2e["production_call"] = {
                "what": "call_consensus of one sample as run.py:709 runs it (snplist = the batch's union, --vcfFileName): "
                        "consensus row + the consensus.vcf data lines, host buffers, wall clock",
                "pileup_consensus_ms": t_call / reps * 1e3, "vcf_text_ms": t_vcf / reps * 1e3,
                "vcf_records_per_sample": n_lines // reps, "vcf_bytes_per_sample": n_bytes // reps}
        for o in owners + [rows_owner, lines_owner]:
            o.free()

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, d_cpu = cpu_baseline(w, matrix_host, min(args.cpu_samples, w.n))
        assert np.array_equal(d_cpu, d_host), "cpu_baseline: the oracle's distance matrix differs from the GPU's"
        try:
            cpu["reference_python"] = reference_python_leg(w)
        except Exception as e:  # noqa: BLE001  (a reported baseline: its failure must not cost the run its line)
            cpu["reference_python"] = {"unavailable": repr(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "positions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "host_cpus_bound": len(numa_cpus) if numa_cpus else None,
            "n_sites": int(n_sites), "matrix_cells_per_s": world * w.n * n_sites / (ms_step * 1e-3),
            "text_gb_per_gpu": w.total_text / 1e9,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
