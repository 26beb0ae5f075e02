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


def _cube_worker(rank, world, name, num_channels):
    sys.path.insert(0, ROOT)
    from katsdpimager_b200 import io
    start, stop = distributed.channel_block(num_channels, rank, world)
    cube = io.FitsCube(name)
    for channel in range(start, stop):
        plane = np.full(cube.shape[1:], float(channel + 1), np.float32)
        plane[:, :, 0] = -float(channel + 1)            # marks the l = 0 column (flipped on disk)
        cube.store(channel, plane)
    cube.close()


def test_cube_written_by_two_worker_processes(tmp_path):
    """The multi-GPU 'gather': every worker process maps the same FITS cube and stores the
    planes of its own channel block (katsdpimager_b200/io.py FitsCube; on a GPU box the store
    is a device copy into the page-locked mapping).  No collective, no copy through rank 0."""
    import torch.multiprocessing as mp
    from katsdpimager_b200 import io, parameters as prm
    fixed = prm.FixedImageParameters([1, 2], np.float32)
    ip = prm.ImageParameters(fixed, wavelength=0.21, pixels=16, pixel_size=1e-4)
    name = str(tmp_path / 'cube.fits')
    num_channels = 5
    io.FitsCube.create(name, num_channels, ip, 856e6, 1e6).close()
    ctx = mp.get_context('spawn')
    procs = [ctx.Process(target=_cube_worker, args=(r, 2, name, num_channels)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    header, data = io.read_fits(name)
    assert data.shape == (num_channels, 2, 16, 16)
    for channel in range(num_channels):
        assert np.all(data[channel, :, :, :-1] == channel + 1)
        assert np.all(data[channel, :, :, -1] == -(channel + 1))
