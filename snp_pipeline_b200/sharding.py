"""Multi-GPU sharding of the path: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in the
CPU tests).  Samples are independent through K1/K3 -- the reference already runs them as one process per sample
(run.py:709-710) -- so ranks own consecutive blocks of the sorted sample list, which keeps the matrix row order equal
to the concatenation of the rank blocks.  The path has ONE exchange step: the union of variant sites must be global
before any sample can be called (merge_sites precedes call_consensus, run.py:691-710).  It is a variable-length
all-gather of each rank's sorted-unique site keys (counts first, then keys padded to the longest list) followed by a
local K2 over the gathered keys, so every rank ends with the identical site list.  K4 then needs every row: one
all-gather of the equal-sized row blocks; each rank computes its stripe of the distance matrix.
"""
from __future__ import annotations


def shard_bounds(n_items, rank, world):
    """Block assignment: ceil(n/world) consecutive items per rank (the last ranks may get fewer or none)."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def zigzag_tile_rows(n_rows, rank, world, tile=64):
    """K4's share of a rank: the 64-row tile rows t with zigzag(t) == rank, where zigzag runs 0..world-1, world-1..0, ...
    Tile row t of the upper triangle holds (T - t) tiles, so a plain block split would leave the last rank nearly
    idle; pairing a long tile row with a short one gives every rank the same number of tiles to within one row."""
    n_tiles = (n_rows + tile - 1) // tile
    out = []
    for t in range(n_tiles):
        k = t % (2 * world)
        if (k if k < world else 2 * world - 1 - k) == rank:
            out.append(t)
    return out


def allgather_varlen(local, dist, world, pad_value=-1):
    """All-gather of 1-D int64 tensors of different lengths.  Returns the list of per-rank tensors (rank order).
    Two collectives: the counts, then the payload padded to the longest list."""
    import torch
    dev = local.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([local.numel()], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine)
    counts_h = counts.cpu().tolist()
    mx = max(max(counts_h), 1)
    padded = torch.full((mx,), pad_value, dtype=torch.int64, device=dev)
    padded[:local.numel()] = local
    gathered = torch.empty(world * mx, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, padded)
    return [gathered[r * mx: r * mx + int(counts_h[r])] for r in range(world)]


def allgather_rows(block, dist, world):
    """All-gather of equal-shaped [rows, sites] uint8 blocks -> [world * rows, sites] in rank order."""
    import torch
    if world == 1:
        return block
    full = torch.empty((world * block.shape[0], block.shape[1]), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(full, block.contiguous())
    return full
