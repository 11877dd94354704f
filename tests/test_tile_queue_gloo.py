"""Host-side logic of the multi-GPU path on CPU: two ``gloo`` ranks pull
row-block tiles from the shared store counter (the dynamic queue used by
``bench.py --gpus N`` under torchrun); every tile must be taken exactly once,
the tiles must cover every pair of the upper triangle exactly once, and the
max-over-ranks timing reduction must work."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graphdot_b200.kernel.marginalized._backend_b200 import PairJobs
from graphdot_b200.kernel.marginalized._tiles import (LocalTileQueue,
                                                      StoreTileQueue,
                                                      row_tiles, tile_pairs)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, n, rows, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    store = dist.distributed_c10d._get_default_store()
    tiles = row_tiles(n, rows)
    mine = []
    for step in range(2):                       # the key is unique per pass
        dist.barrier()
        for i0, i1 in StoreTileQueue(store, tiles, f'tiles{step}'):
            mine.append((step, i0, i1))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)    # max-over-ranks timing
    pairs = torch.tensor([sum(tile_pairs(a, b, n) for _, a, b in mine)],
                         dtype=torch.float64)
    dist.all_reduce(pairs, op=dist.ReduceOp.SUM)
    np.save(os.path.join(out_dir, f'rank{rank}.npy'), np.array(mine))
    if rank == 0:
        np.save(os.path.join(out_dir, 'reduced.npy'),
                np.array([t.item(), pairs.item()]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_store_tile_queue_two_ranks(tmp_path):
    n, rows, world = 53, 8, 2
    mp.spawn(_rank_main, args=(world, _free_port(), n, rows, str(tmp_path)),
             nprocs=world, join=True)
    taken = [tuple(x) for r in range(world)
             for x in np.load(tmp_path / f'rank{r}.npy').reshape(-1, 3)]
    tiles = row_tiles(n, rows)
    for step in range(2):
        got = sorted((a, b) for s, a, b in taken if s == step)
        assert got == sorted(tiles)             # each tile exactly once
    tmax, pairs = np.load(tmp_path / 'reduced.npy')
    assert tmax == world and pairs == 2 * n * (n + 1) // 2


def test_tiles_cover_the_upper_triangle_exactly_once():
    n, rows = 37, 5
    seen = np.zeros((n, n), dtype=int)
    total = 0
    for i0, i1 in LocalTileQueue(row_tiles(n, rows)):
        jobs = np.asarray(PairJobs.triu(i0, i1, n))
        assert len(jobs) == tile_pairs(i0, i1, n)
        seen[jobs['i'], jobs['j']] += 1
        total += len(jobs)
    assert total == n * (n + 1) // 2
    assert np.array_equal(seen, np.triu(np.ones((n, n), dtype=int)))
