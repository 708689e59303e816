"""The N > 1 path's host logic on CPU: world_size 2 over gloo (127.0.0.1).  Each rank owns a block of samples, the
ranks all-gather their sorted-unique site lists (variable length) and their matrix row blocks; the merged result must
equal the single-process union / matrix.  The device kernels are exercised by the -m gpu tests; here the oracle's
packed-key union stands in for K2 so that the exchange logic can be checked without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from snp_pipeline_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sample_keys(n_samples, seed):
    rng = np.random.default_rng(seed)
    out = []
    for s in range(n_samples):
        k = np.unique(rng.choice(5000, size=rng.integers(0, 400), replace=False).astype(np.int64)
                      | (rng.integers(0, 2, size=1).astype(np.int64)[0] << 32))
        out.append(k)
    return out


def _worker(rank, world, port, n_samples, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        samples = _sample_keys(n_samples, 7)
        lo, hi = sharding.shard_bounds(n_samples, rank, world)
        mine = samples[lo:hi]
        local = np.unique(np.concatenate(mine)) if mine and sum(m.size for m in mine) else np.zeros(0, np.int64)
        parts = sharding.allgather_varlen(torch.from_numpy(local), dist, world)
        merged = np.unique(np.concatenate([p.numpy() for p in parts]))
        # rows: every rank fills its block of the matrix (cell = 1 where the sample carries the site)
        per = (n_samples + world - 1) // world
        block = np.zeros((per, max(merged.size, 1)), dtype=np.uint8)
        for i, s in enumerate(mine):
            block[i, np.searchsorted(merged, s)] = 1
        full = sharding.allgather_rows(torch.from_numpy(block), dist, world).numpy()
        np.save(os.path.join(tmp, "merged_%d.npy" % rank), merged)
        np.save(os.path.join(tmp, "full_%d.npy" % rank), full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [7, 2, 1])
def test_two_rank_exchange(tmp_path, n_samples):
    from oracle import oracle as orc
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_samples, str(tmp_path)), nprocs=world, join=True)
    samples = _sample_keys(n_samples, 7)
    keys = np.concatenate(samples).astype(np.uint64) if samples else np.zeros(0, np.uint64)
    samp = np.concatenate([np.full(s.size, i, np.uint32) for i, s in enumerate(samples)])
    want_uniq, want_cnt, _ = orc.merge_sites_keys(keys, samp)
    m0, m1 = (np.load(tmp_path / ("merged_%d.npy" % r)) for r in range(world))
    assert np.array_equal(m0, m1), "every rank must end with the identical site list"
    assert np.array_equal(m0.astype(np.uint64), want_uniq)
    f0, f1 = (np.load(tmp_path / ("full_%d.npy" % r)) for r in range(world))
    assert np.array_equal(f0, f1)
    per = (n_samples + world - 1) // world
    for i, s in enumerate(samples):        # row i of the gathered matrix is sample i (rank blocks concatenate in order)
        r, k = divmod(i, per)
        row = f0[r * per + k]
        assert row.sum() == s.size and (row[np.searchsorted(m0, s)] == 1).all()
    assert int(f0[:, :max(m0.size, 1)].sum(axis=0).astype(np.int64)[: m0.size].sum()) == int(want_cnt.sum())


def test_shard_bounds():
    for n in (0, 1, 7, 100, 1000):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(0 <= hi - lo <= (n + world - 1) // world for lo, hi in spans)
