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


def _rank_shared(rank, world, port, nx, ny, cols, path, out_dir):
    """The C5 arrangement of bench.py on CPU: rank 0 'packs' and broadcasts a
    payload, both ranks pull COLUMN tiles of the X-by-Y rectangle from the
    store counter and write them straight into ONE shared host matrix (a
    memory-mapped file); rank 0 then holds the assembled result."""
    from graphdot_b200.kernel.marginalized._tiles import col_tiles
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                      RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    store = dist.distributed_c10d._get_default_store()
    payload = [dict(blob=b'x' * 1000, cuts=list(range(5)))
               if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    assert payload[0]['cuts'] == list(range(5))
    if rank == 0:
        with open(path, 'wb') as f:
            f.truncate(nx * ny * 4)
    dist.barrier()
    K = np.memmap(path, dtype=np.float32, mode='r+').reshape((nx, ny),
                                                              order='F')
    done = 0
    for j0, j1 in StoreTileQueue(store, col_tiles(ny, cols), 'c5'):
        jobs = np.asarray(PairJobs.rect(0, nx, nx + j0, nx + j1))
        assert len(jobs) == nx * (j1 - j0)
        # stand-in for the solve: entry (i, j) = 1000 i + j
        K[jobs['i'], jobs['j'] - nx] = 1000.0 * jobs['i'] + (jobs['j'] - nx)
        done += len(jobs)
    K.flush()
    total = torch.tensor([float(done)], dtype=torch.float64)
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, 'total.npy'), np.array([total.item()]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_column_tiles_fill_one_shared_host_matrix(tmp_path):
    nx, ny, cols, world = 19, 23, 4, 2
    path = str(tmp_path / 'gram.bin')
    mp.spawn(_rank_shared, args=(world, _free_port(), nx, ny, cols, path,
                                 str(tmp_path)), nprocs=world, join=True)
    K = np.fromfile(path, dtype=np.float32).reshape((nx, ny), order='F')
    want = 1000.0 * np.arange(nx)[:, None] + np.arange(ny)[None, :]
    assert np.array_equal(K, want)
    assert np.load(tmp_path / 'total.npy')[0] == nx * ny
