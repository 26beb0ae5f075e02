"""The channel-parallel layer on CPU: partitioning logic and a world_size-2 gloo run of
the plane gather (the only inter-rank step; the hot path itself has no collective)."""
import os
import socket
import sys

import numpy as np
import pytest

from katsdpimager_b200 import distributed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('channels,world', [(64, 8), (4096, 8), (10, 4), (3, 8), (7, 2), (1, 1)])
def test_channel_blocks_partition(channels, world):
    blocks = [distributed.channel_block(channels, r, world) for r in range(world)]
    assert blocks[0][0] == 0 and blocks[-1][1] == channels
    for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in blocks]
    assert max(sizes) - min(sizes) <= 1
    for c in range(channels):
        r = distributed.owner_of(c, channels, world)
        assert blocks[r][0] <= c < blocks[r][1]
    with pytest.raises(ValueError):
        distributed.channel_block(channels, world, world)


def test_gather_single_process():
    planes = {2: np.full((2, 4, 4), 2.0, np.float32), 0: np.zeros((2, 4, 4), np.float32)}
    cube = distributed.gather_planes(planes, [0, 2], 3)
    assert cube.shape == (3, 2, 4, 4)
    assert np.all(cube[2] == 2.0) and np.all(np.isnan(cube[1]))


def _worker(rank, world, port, queue):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        num_channels = 5
        start, stop = distributed.channel_block(num_channels, rank, world)
        planes = {c: np.full((2, 8, 8), float(c + 1), np.float32) for c in range(start, stop)}
        cube = distributed.gather_planes(planes, range(start, stop), num_channels, dist)
        if rank == 0:
            ok = cube.shape == (5, 2, 8, 8) and all(np.all(cube[c] == c + 1) for c in range(5))
            queue.put(bool(ok))
        else:
            queue.put(cube is None)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results == [True, True]
