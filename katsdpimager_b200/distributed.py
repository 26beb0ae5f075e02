"""Channel-parallel multi-GPU layout (new functionality; the reference processes
channels sequentially in one process, reference frontend.py:749-767, and scales out by
launching separate processes on disjoint ``--start-channel/--stop-channel`` ranges,
frontend.py:281-284).

Spectral channels share nothing mutable (each has its own image/grid parameters, kernel
table, weights, PSF, CLEAN state and output plane), so the hot path needs no collective:
every rank images a contiguous block of channels on its own GPU, and only the finished
``[polarizations, N, N]`` planes are gathered on the host.  These helpers hold the
partitioning and the gather so that they can be tested on CPU with the gloo backend.
"""
import numpy as np


def channel_block(num_channels, rank, world_size):
    """Contiguous block [start, stop) of `num_channels` owned by `rank`.

    Blocks rather than round-robin so each worker streams a contiguous slab of the
    (channel, w_slice, vis) store; block sizes differ by at most one channel.
    """
    if not 0 <= rank < world_size:
        raise ValueError('rank {} out of range for world size {}'.format(rank, world_size))
    base, extra = divmod(num_channels, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def owner_of(channel, num_channels, world_size):
    """Rank that owns `channel` under :func:`channel_block`."""
    base, extra = divmod(num_channels, world_size)
    boundary = extra * (base + 1)
    if channel < boundary:
        return channel // (base + 1)
    return extra + (channel - boundary) // base


def gather_planes(planes, channels, num_channels, dist=None, dst=0):
    """Assemble per-rank image planes into a cube on rank `dst`.

    `planes` maps channel -> array ``[polarizations, N, N]`` for the channels this rank
    imaged.  `dist` is ``torch.distributed`` (an initialised process group, any backend
    with CPU tensors, e.g. gloo) or None for a single process.  Returns the cube
    ``[num_channels, polarizations, N, N]`` on `dst` (channels nobody imaged are NaN) and
    None elsewhere.
    """
    channels = list(channels)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        gathered = [(channels, [planes[c] for c in channels])]
    else:
        import torch
        payload = (channels, [torch.from_numpy(np.ascontiguousarray(planes[c])) for c in channels])
        out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
        dist.gather_object(payload, out, dst=dst)
        if dist.get_rank() != dst:
            return None
        gathered = [(chs, [t.numpy() for t in tensors]) for chs, tensors in out]
    cube = None
    for chs, arrays in gathered:
        for c, a in zip(chs, arrays):
            if cube is None:
                cube = np.full((num_channels,) + a.shape, np.nan, a.dtype)
            cube[c] = a
    return cube


def image_channel_block(template, command_queue, channels, load_channel, cube, major, vis_block,
                        weight_parameters, restore=None):
    """Image the channels of this rank's block and store each finished plane in `cube`
    (a :class:`~.io.FitsCube` mapped by every rank): the channel loop of reference
    frontend.py:749-767 restricted to ``channels``, one :func:`~.pipeline.process_channel`
    per channel, visibilities uploaded once per channel.

    ``load_channel(channel)`` returns ``(image_parameters, grid_parameters, slices)`` with
    `slices` the preprocessed records per W slice.  Returns {channel: statistics}.
    """
    from . import pipeline
    stats = {}
    pols = len(template.fixed_image_parameters.polarizations)
    for channel in channels:
        ip, gp, slices = load_channel(channel)
        imager = template.instantiate(command_queue, ip, gp, vis_block, 0, major)
        imager.ensure_all_bound()
        vis = pipeline.ResidentVisibilities(command_queue, slices, pols)
        stats[channel] = pipeline.process_channel(
            imager, vis, ip, gp, template.clean_parameters, weight_parameters, major, vis_block,
            restore=restore)
        cube.store_device(channel, imager.buffer('dirty'), command_queue)
        command_queue.finish()
    return stats
