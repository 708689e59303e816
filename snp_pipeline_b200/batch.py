"""The hot path over MANY samples on one or more GPUs, files out -- the package-level driver.

What the reference does with four subcommands and one process per sample (run.py:691-692 merge_sites, :709-710
call_consensus through run_array over sampleDirectories.txt, :724-725 snp_matrix, :775-776 distance), as one pass:

    K2   union of the samples' variant sites                      (merge_sites.py:94-116, utils.py:1056-1070)
    K1   every sample's pileup parsed / tallied / called          (call_consensus.py:161-188)
    K3   gather into the samples x sites matrix                   (snp_matrix.py:114-119)
    K4   all-pairs SNP distance                                   (distance.py:90-96, utils.py:1135-1165)

and the same three files: snplist.txt, snpma.fasta, snp_distance_matrix.tsv (+ the pairwise TSV on request), byte for
byte what the subcommands write.

Multi-GPU (one process per GPU, torch.distributed): ranks own consecutive blocks of the SORTED sample list, so the
matrix row order is the concatenation of the rank blocks.  The path has one exchange step before K1 -- the union of
variant sites must be global before any sample can be called: ONE variable-length all-gather of the per-sample site
lists ((chrom rank, pos) keys + owning sample), after which every rank runs K2 over the same global list and holds the
same site table -- and one after it: the all-gather of the matrix rows, so that every rank can compute its share of
the distance matrix.  Rank 0 writes the files.  No other collective exists on the path.
"""
from __future__ import annotations

import os

import numpy as np

from . import _lib, sharding, utils


class HotPathResult(object):
    """What run_hot_path() leaves on every rank (files: rank 0 only)."""

    def __init__(self):
        self.n_samples = 0
        self.n_sites = 0
        self.sample_names = []        # all samples, matrix row order
        self.matrix = None            # torch uint8 [n_samples, n_sites] (every rank, after the row all-gather)
        self.distance = None          # torch int32 [n_samples, n_samples] on rank 0, else None
        self.stats = None             # numpy int64 [n_local, 6]: n_lines, n_parsed, n_general, error_offset, error_code, n_called
        self.files = {}


def _chrom_ranks(contigs):
    order = sorted(set(contigs))
    return order, {c: i for i, c in enumerate(order)}


def site_keys(positions, rank_of):
    """[(chrom, pos)] -- or {chrom: integer array of positions} -- -> uint64 (chrom rank in string order << 32 | pos),
    duplicates dropped (a set in the reference, utils.py:1127-1131)."""
    if isinstance(positions, dict):
        parts = []
        for c, pos in positions.items():
            pos = np.asarray(pos, dtype=np.int64)
            if pos.size > 1 and not bool(np.all(pos[1:] > pos[:-1])):      # (VCF order: already strictly rising)
                pos = np.unique(pos)
            if pos.size and not (0 <= int(pos.min()) and int(pos.max()) < (1 << 31)):
                raise ValueError("VCF position outside [0, 2^31)")
            parts.append((np.uint64(rank_of[c]) << np.uint64(32)) | pos.astype(np.uint64))
        return np.concatenate(parts) if parts else np.zeros(0, np.uint64)
    seen = dict.fromkeys(positions)
    for _, p in seen:
        if not 0 <= p < (1 << 31):
            raise ValueError("VCF position %d outside [0, 2^31)" % p)
    return np.array([(rank_of[c] << 32) | p for c, p in seen], dtype=np.uint64)


def _site_items(site_lists):
    """(chrom, largest position) pairs of site lists in either form"""
    for s in site_lists:
        if isinstance(s, dict):
            for c, pos in s.items():
                yield c, (int(np.max(pos)) if len(pos) else 0)
        else:
            for c, p in s:
                yield c, p


def snplist_text(chroms, uniq, cnt, samples, names):
    """snplist.txt (utils.write_list_of_snps, utils.py:1056-1070): chrom, pos, count, the samples' names."""
    out, o = [], 0
    samples = samples.tolist()
    for k, c in zip(uniq.tolist(), cnt.tolist()):
        out.append("%s\t%d\t%d\t%s\n" % (chroms[k >> 32], k & 0xffffffff, c, "\t".join([names[i] for i in samples[o:o + c]])))
        o += c
    return "".join(out)


def distance_matrix_text(ids, d):
    """snp_distance_matrix.tsv (distance.py:108-115); ids sorted like distance.py:90."""
    order = sorted(range(len(ids)), key=lambda i: ids[i])
    dd = d[np.ix_(order, order)] if len(ids) else d
    sid = [ids[i] for i in order]
    rows = ["\t%s\n" % "\t".join(sid)]
    for a, id1 in enumerate(sid):
        rows.append("%s\t%s\n" % (id1, "\t".join(map(str, dd[a].tolist()))))
    return "".join(rows)


def distance_pairwise_text(ids, d):
    """snp_distance_pairwise.tsv (distance.py:99-106)."""
    order = sorted(range(len(ids)), key=lambda i: ids[i])
    dd = d[np.ix_(order, order)] if len(ids) else d
    sid = [ids[i] for i in order]
    rows = ["%s\n" % "\t".join(["Seq1", "Seq2", "Distance"])]
    for a, id1 in enumerate(sid):
        row = dd[a].tolist()
        rows.append("".join("%s\t%s\t%i\n" % (id1, id2, row[b]) for b, id2 in enumerate(sid)))
    return "".join(rows)


def run_hot_path(ctx, local_names, local_texts, local_sites, contigs, contig_len, params, out_dir=None,
                 mode=_lib.MODE_SITES, dist=None, rank=0, world=1, n_total=None, pairwise=False, distance=True):
    """One pass of the path over this rank's block of the sorted sample list.

    local_names   sample names of the block (global order = rank 0's block, rank 1's block, ...)
    local_texts   per sample: a CUDA uint8 tensor holding the pileup file's bytes (16-byte aligned, resident in HBM)
    local_sites   per sample: [(chrom, pos)] of its variant sites (the positions of var.flt.vcf, utils.py:1113-1132),
                  or {chrom: array of positions}
    contigs       every contig name of the reference, contig_len their lengths (bounds of the site bitmap); None: the
                  contigs the site lists name, each as long as its largest site position
    out_dir       rank 0 writes snplist.txt, snpma.fasta, snp_distance_matrix.tsv there (None: no files)
    Returns a HotPathResult.  Raises SnpGpuError where the reference's call_consensus would raise on a sample."""
    import torch
    res = HotPathResult()
    dev = torch.device("cuda", torch.cuda.current_device())
    if contigs is None:                                       # (the site lists bound the table: no reference needed)
        seen = {}
        for c, p in _site_items(local_sites):
            seen[c] = max(seen.get(c, 0), p)
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, seen)
            seen = {}
            for d in parts:
                for c, p in d.items():
                    seen[c] = max(seen.get(c, 0), p)
        contigs, contig_len = list(seen), [seen[c] for c in seen]
    chroms, rank_of = _chrom_ranks(contigs)
    len_of = dict(zip(contigs, contig_len))
    for c, _ in _site_items(local_sites):
        if c not in rank_of:
            raise ValueError("variant site on contig %r, which the reference does not hold" % c)
    n_local = len(local_names)
    per = n_local
    if world > 1:
        t = torch.tensor([n_local], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per = int(t.item())                                  # rows per rank block (the last blocks may be shorter)
    first = rank * per                                        # global index of this block's first sample
    # ---- the exchange in front of K1: every sample's site list, to every rank ------------------------------------
    keys = [site_keys(s, rank_of) for s in local_sites]
    k_local = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
    s_local = np.concatenate([np.full(k.size, first + i, dtype=np.int64) for i, k in enumerate(keys)]) if keys else np.zeros(0, np.int64)
    k_dev = torch.from_numpy(k_local.view(np.int64)).to(dev)
    s_dev = torch.from_numpy(s_local).to(dev)
    if world > 1:
        all_names = [None] * world
        dist.all_gather_object(all_names, list(local_names))
        k_dev = torch.cat(sharding.allgather_varlen(k_dev, dist, world))
        s_dev = torch.cat(sharding.allgather_varlen(s_dev, dist, world))
        names = [None] * (world * per)
        for r, blk in enumerate(all_names):
            names[r * per: r * per + len(blk)] = blk
    else:
        names = list(local_names)
    n_rows = len(names)                                       # incl. the padding rows of short blocks (name None)
    # ---- K2: the global union (every rank the same) -----------------------------------------------------------------
    nk = int(k_dev.numel())
    samp32 = s_dev.to(torch.int32)
    uniq = torch.empty(max(nk, 1), dtype=torch.int64, device=dev)
    cnt = torch.empty(max(nk, 1), dtype=torch.int32, device=dev)
    grouped = torch.empty(max(nk, 1), dtype=torch.int32, device=dev)
    n_sites = ctx.merge_sites_dev(k_dev.data_ptr(), samp32.data_ptr(), nk, uniq.data_ptr(), cnt.data_ptr(), grouped.data_ptr()) if nk else 0
    res.n_sites = n_sites
    # ---- K1 + K3: this block's rows ---------------------------------------------------------------------------------
    sites = _lib.Sites.from_keys_dev(ctx, chroms, [len_of[c] for c in chroms], uniq.data_ptr(), n_sites)
    width = (max(n_sites, 1) + 63) // 64 * 64                # row stride: 16-byte aligned rows (K4's coalesced pack kernel)
    block = torch.full((per, width), ord("-"), dtype=torch.uint8, device=dev)
    stats = torch.zeros((max(n_local, 1), 6), dtype=torch.int64, device=dev)
    if n_local:
        ctx.pileup_consensus_batch_dev([(t.data_ptr(), int(t.numel()), block[i].data_ptr(), 0, 0, stats[i].data_ptr())
                                        for i, t in enumerate(local_texts)], sites, params, mode)
    st = stats[:n_local].cpu().numpy()                        # (synchronises)
    res.stats = st
    for i in range(n_local):
        if st[i, 4]:
            sites.close()
            raise _lib.SnpGpuError(int(st[i, 4]), "sample %s: the reference raises on the pileup line at byte offset %d"
                                   % (local_names[i], st[i, 3]), int(st[i, 3]))
    # ---- the exchange behind it: every row to every rank; K4 on this rank's share ------------------------------------
    matrix = sharding.allgather_rows(block, dist, world)
    res.matrix = matrix
    d_host = None
    if distance:
        if world == 1:
            full = torch.zeros((n_rows, n_rows), dtype=torch.int32, device=dev)
            if n_sites and n_rows:
                ctx.pairwise_distance_dev(matrix.data_ptr(), n_rows, n_sites, width, 0, n_rows, full.data_ptr())
            res.distance = full
        else:
            # the triangle's 64-row tile rows, dealt in zigzag order: every pair is computed once, on one rank
            n_tiles = (n_rows + 63) // 64
            share = [sharding.zigzag_tile_rows(n_rows, r, world) for r in range(world)]
            most = max(len(x) for x in share)
            part = torch.zeros((max(most, 1) * 64, n_rows), dtype=torch.int32, device=dev)
            if n_sites and share[rank]:
                ctx.pairwise_distance_tiles_dev(matrix.data_ptr(), n_rows, n_sites, width, share[rank], part.data_ptr())
            parts = [torch.empty_like(part) for _ in range(world)] if rank == 0 else None
            dist.gather(part, parts, dst=0)
            if rank == 0:
                upper = torch.zeros((n_tiles, 64, n_rows), dtype=torch.int32, device=dev)
                for r in range(world):
                    if share[r]:
                        upper[torch.tensor(share[r], device=dev)] = parts[r].view(-1, 64, n_rows)[:len(share[r])]
                upper = upper.view(n_tiles * 64, n_rows)[:n_rows]
                row = torch.arange(n_rows, device=dev)
                res.distance = torch.where(row[None, :] >= (row // 64 * 64)[:, None], upper, upper.t())
    sites.close()
    # ---- files (rank 0) ---------------------------------------------------------------------------------------------
    keep = [i for i, nm in enumerate(names) if nm is not None]
    res.sample_names = [names[i] for i in keep]
    res.n_samples = len(keep)
    if rank == 0 and out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        uniq_h = uniq[:n_sites].cpu().numpy().view(np.uint64)
        cnt_h = cnt[:n_sites].cpu().numpy()
        grouped_h = grouped[:nk].cpu().numpy()
        res.files["snplist"] = os.path.join(out_dir, "snplist.txt")
        with open(res.files["snplist"], "w") as f:
            f.write(snplist_text(chroms, uniq_h, cnt_h, grouped_h, names))
        m_host = matrix.cpu().numpy()
        res.files["snpma"] = os.path.join(out_dir, "snpma.fasta")
        with open(res.files["snpma"], "w") as f:              # snp_matrix.py:114-119: the samples' consensus.fasta files, in order
            for i in keep:
                f.write(utils.fasta_record_text(names[i], m_host[i, :n_sites].tobytes().decode("ascii")))
        if distance:
            d_host = res.distance.cpu().numpy()[np.ix_(keep, keep)]
            res.files["distance_matrix"] = os.path.join(out_dir, "snp_distance_matrix.tsv")
            with open(res.files["distance_matrix"], "w") as f:
                f.write(distance_matrix_text(res.sample_names, d_host))
            if pairwise:
                res.files["distance_pairwise"] = os.path.join(out_dir, "snp_distance_pairwise.tsv")
                with open(res.files["distance_pairwise"], "w") as f:
                    f.write(distance_pairwise_text(res.sample_names, d_host))
    return res


# ------------------------------------------------------------------------------------------ sample directories in, files out
def run_sample_dirs(args):
    """`python -m snp_pipeline_b200.batch` (under torchrun for more than one GPU): the reference's merge_sites ->
    call_consensus (every sample) -> snp_matrix -> distance sequence (run.py:682-732, 770-784) over the sample
    directories listed in args.sampleDirsFile -- <dir>/reads.all.pileup and <dir>/var.flt.vcf in, <dir>/consensus.fasta
    per sample and snplist.txt / snpma.fasta / snp_distance_matrix.tsv / snp_distance_pairwise.tsv in args.outDir out."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    with open(args.sampleDirsFile) as f:
        dirs = sorted(d for d in (line.rstrip() for line in f) if d)
    lo, hi = sharding.shard_bounds(len(dirs), rank, world)
    mine = dirs[lo:hi]
    ctx = _lib.Context(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    contigs, lens = None, None
    if args.referenceFile:                                    # contig names and lengths of the reference fasta
        contigs, lens = [], []
        with open(args.referenceFile) as f:
            for line in f:
                if line.startswith(">"):
                    contigs.append(line[1:].split()[0] if line[1:].split() else "")
                    lens.append(0)
                elif contigs:
                    lens[-1] += len(line.strip())
    texts, sites = [], []
    for d in mine:
        raw = np.fromfile(os.path.join(d, args.pileupFileName), dtype=np.uint8)
        t = torch.empty(raw.size + 64, dtype=torch.uint8, device="cuda")
        t[:raw.size].copy_(torch.from_numpy(raw))
        texts.append(t[:raw.size])
        sites.append(utils.read_vcf_positions(os.path.join(d, args.vcfFileName)))
    params = _lib.make_params(args.minBaseQual, args.minConsFreq, args.minConsDpth, args.minConsStrdDpth, args.minConsStrdBias)
    names = [os.path.basename(os.path.normpath(d)) for d in mine]
    res = run_hot_path(ctx, names, texts, sites, contigs, lens, params, out_dir=args.outDir, dist=dist, rank=rank,
                       world=world, pairwise=True)
    per = (len(dirs) + world - 1) // world
    block = res.matrix[rank * per: rank * per + len(mine)].cpu().numpy()
    for i, d in enumerate(mine):                              # call_consensus.py:189-192
        with open(os.path.join(d, args.consensusFileName), "w") as f:
            f.write(utils.fasta_record_text(names[i], block[i, :res.n_sites].tobytes().decode("ascii")))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return res


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(prog="python -m snp_pipeline_b200.batch", description=run_sample_dirs.__doc__)
    ap.add_argument("sampleDirsFile")
    ap.add_argument("-r", "--referenceFile", default=None, help="reference fasta (bounds the site table; default: the VCFs do)")
    ap.add_argument("-o", "--outDir", default=".")
    ap.add_argument("--pileupFileName", default="reads.all.pileup")
    ap.add_argument("--vcfFileName", default="var.flt.vcf")
    ap.add_argument("--consensusFileName", default="consensus.fasta")
    ap.add_argument("-q", "--minBaseQual", type=int, default=0)          # defaults of cfsan_snp_pipeline.py:397-407
    ap.add_argument("-c", "--minConsFreq", type=float, default=0.60)
    ap.add_argument("-D", "--minConsDpth", type=int, default=1)
    ap.add_argument("-d", "--minConsStrdDpth", type=int, default=0)
    ap.add_argument("-b", "--minConsStrdBias", type=float, default=0)
    run_sample_dirs(ap.parse_args(argv))


if __name__ == "__main__":
    main()
