"""Worker of tests/test_batch.py (also runnable under torchrun): a few small synthetic samples through
snp_pipeline_b200.batch.run_hot_path on WORLD_SIZE GPUs; rank 0 compares the three files with the oracle's texts.
    python tests/hotpath_worker.py <out_dir> <n_samples> <genome_len> [all]
Prints "HOTPATH OK ..." and exits 0 when the files are byte-identical."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

CONTIGS = ["gi|0000000|ref|SYN_TEST.1|", "NC_000000.2_plasmid_with_a_rather_long_name_x"]


def main():
    out_dir, n_samples, genome_len = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    parse_all = len(sys.argv) > 4 and sys.argv[4] == "all"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from snp_pipeline_b200 import _lib, batch, sharding
    ctx = _lib.Context(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    lo, hi = sharding.shard_bounds(n_samples, rank, world)
    names_all = ["sample_%03d" % i for i in range(n_samples)]
    texts, sites = {}, {}
    cap = genome_len * 112 + 4096
    scratch = torch.empty(cap, dtype=torch.uint8, device="cuda")

    def make(i):
        """sample i: two contigs, the second one's text appended (its own seed)"""
        parts, pos = [], []
        for c, name in enumerate(CONTIGS):
            g = genome_len if c == 0 else max(genome_len // 7, 50)
            spec = _lib.SynthSpec(777 + c, i, g, 20, max(g // 40, 1), 0.3, 0.002)
            nb = ctx.synth_pileup_dev(spec, name, scratch.data_ptr(), cap)
            parts.append(scratch[:nb].clone())
            pos += [(name, int(p)) for p in ctx.synth_sample_sites(spec)]
        return torch.cat(parts), pos

    for i in range(lo, hi):
        texts[i], sites[i] = make(i)
    params = _lib.make_params(min_cons_depth=2)
    res = batch.run_hot_path(ctx, names_all[lo:hi], [texts[i] for i in range(lo, hi)], [sites[i] for i in range(lo, hi)],
                             CONTIGS, [genome_len, max(genome_len // 7, 50)], params, out_dir=out_dir,
                             mode=_lib.MODE_ALL if parse_all else _lib.MODE_SITES, dist=dist, rank=rank, world=world,
                             pairwise=True)
    ok = True
    if rank == 0:
        from oracle import oracle as orc
        orc.build()
        all_texts, all_sites = [], []
        for i in range(n_samples):                    # rank 0 regenerates every sample for the oracle
            t, s = texts[i] if i in texts else make(i)[0], sites[i] if i in sites else make(i)[1]
            all_texts.append(t.cpu().numpy())
            all_sites.append(s)
        snplist, snpma, dmat, _ = orc.hot_path_texts(names_all, all_texts, all_sites, orc.make_params(min_cons_depth=2),
                                                     parse_all=parse_all, threads=4)
        for key, want in (("snplist", snplist), ("snpma", snpma), ("distance_matrix", dmat)):
            got = open(res.files[key]).read()
            same = got == want
            ok &= same
            print("%s %s sha256 %s" % (key, "identical" if same else "DIFFERENT", hashlib.sha256(got.encode()).hexdigest()[:16]))
        assert res.n_sites > 0 and res.n_samples == n_samples
        print("HOTPATH %s world=%d samples=%d sites=%d" % ("OK" if ok else "MISMATCH", world, n_samples, res.n_sites))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
